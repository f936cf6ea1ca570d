"""Throughput of the fp32 enumeration kernel across data regimes: the same K = 6, M = 10 problem at
coverages from 20 (flat tables: most terms of a convolution matter) to 20000 (sharp peaks: a handful do).
    python profiles/window_sweep.py        (on a B200)"""
import sys, time, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bamse_b200 import cuda_shim
rng = np.random.default_rng(1)
ctx = cuda_shim.Context(0)
for cov in (20, 100, 1000, 20000):
    K, M, C = 6, 10, 8
    var = np.zeros((C, K, M), dtype=np.int64); ref = np.zeros_like(var)
    for c in range(C):
        frac = np.sort(rng.dirichlet(np.ones(K + 1))[:K])[::-1]
        cm = np.clip(np.cumsum(frac[::-1])[::-1][:, None] * rng.uniform(0.6, 1.0, size=(K, M)), 0, 1)
        a = rng.binomial(cov, cm * 0.497 + 0.003); var[c] = a; ref[c] = cov - a
    ctx.upload_problem([0.003], [K] * C, var, ref)
    const = np.zeros((1, C))
    ctx.score_topk(const, topk=5, precision=cuda_shim.FP32)
    ctx.kernel_time(True); ctx.kernel_stats(True)
    for _ in range(3): rec, _, _ = ctx.score_topk(const, topk=5, precision=cuda_shim.FP32)
    ms, n = ctx.kernel_time(True); st = ctx.kernel_stats(True)
    # accuracy vs fp64 on a sample
    s32, _ = ctx.score_range(0, 0, 0, 400, cuda_shim.FP32); s64, _ = ctx.score_range(0, 0, 0, 400, cuda_shim.FP64)
    print("cov %6d: %.2f M evals/s  terms/conv %.0f  max rel err %.2e" % (cov, C * 7776 * 3 / ms / 1e3, st["conv_terms"] / max(1, st["convs"]) , np.max(np.abs(s32 - s64) / np.abs(s64))))
