#!/usr/bin/env python
"""The numbers of profiles/r02_pair_geometry.txt (counted by cutting ALL K^(K-1) trees on the GPU) recomputed with the
HOST instances of the same code (bamse_decode_tree + bamse_debug_classify, no GPU): pack rows / context rows in use
(closed under the recipes' parent links) and irregular trees per geometry.   usage: host_cut_stats.py K,s,r [K,s,r ...]
(1.5 us per tree: K = 9 takes a minute, K = 10 half an hour).  Results: profiles/r02_host_cut_stats.txt"""
import sys, time, ctypes
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ('tests', 'oracle', ''):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
from test_pairs_cpu import pair_geometry, ctx_recipes
from bamse_b200 import cuda_shim
lib = cuda_shim.load_library()
def stats(K, s, r):
    s, r, nf, nt = pair_geometry(lib, K, s, r)
    crec = ctx_recipes(lib, K, s, r, nt)
    T = K ** (K - 1)
    used_fg = np.zeros(nf, dtype=bool); used_t = np.zeros(nt, dtype=bool)
    par = np.zeros(K, dtype=np.int8); out = np.zeros(3, dtype=np.int32)
    pp, po = par.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p)
    irregular = 0
    dec, cla = lib.bamse_decode_tree, lib.bamse_debug_classify
    for i in range(T):
        dec(K, i, pp)
        cla(K, pp, s, r, po)
        if out[0]:
            used_fg[out[1]] = True; used_t[out[2]] = True
        else:
            irregular += 1
    # close the context rows in use under the recipes' parent links (rows are ordered by node count)
    for i in range(nt - 1, -1, -1):
        if used_t[i] and crec[i, 0] >= 0:
            used_t[crec[i, 0]] = True
    return dict(K=K, s=s, r=r, trees=T, inside_rows=nf, pack_rows_used=int(used_fg.sum()), context_rows_used=int(used_t.sum()), context_rows_full=nt, irregular=irregular)
if __name__ == "__main__":
    for K, s, r in [tuple(int(x) for x in a.split(',')) for a in sys.argv[1:]]:
        t0 = time.time(); print(stats(K, s, r), "%.1fs" % (time.time() - t0), flush=True)
