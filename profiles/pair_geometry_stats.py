"""Rows and irregular trees of candidate pair geometries (s, r) per cluster count, counted on the GPU by cutting
every tree (bamse_debug_pair_stats).  usage (GPU box): python profiles/pair_geometry_stats.py > gpurun_out/pair_geometry.txt"""
import sys
import time
sys.path.insert(0, ".")
from bamse_b200 import cuda_shim

CASES = {5: [(3, 3), (3, 2), (2, 3)], 6: [(3, 4), (4, 3), (3, 3)], 7: [(4, 4), (4, 3), (5, 3)], 8: [(4, 5), (5, 4), (5, 5), (4, 4)],
         9: [(5, 5), (5, 6), (6, 5), (4, 5)], 10: [(5, 5), (6, 5), (5, 6), (6, 6)]}
with cuda_shim.Context(0) as ctx:
    for K, cases in CASES.items():
        for s, r in cases:
            t0 = time.time()
            try:
                st = ctx.pair_stats(K, s, r)
            except Exception as e:
                print("K=%d s=%d r=%d: %s" % (K, s, r, e))
                continue
            T = K ** (K - 1)
            print("K=%2d s=%d r=%d  trees %11d  inside rows %8d  context rows %9d of %9d  irregular %10d (%.4f)  rows/tree %.5f  [%.1f s]"
                  % (K, s, r, T, st["nf"], st["nt_used"], st["nt_full"], st["irregular"], st["irregular"] / T,
                     (st["nf"] + st["nt_used"]) / T, time.time() - t0), flush=True)
