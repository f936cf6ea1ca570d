import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import numpy as np
from bamse_b200 import cuda_shim, workloads
from bamse_b200.clustering import _pack_problem
ctx = cuda_shim.Context(0)
work = workloads.build("c2", ctx=ctx)
clusts = work["clusts"]
Ks, var, ref, Kmax = _pack_problem(clusts)
const = np.zeros((1, len(clusts)))
import torch
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); c = cuda_shim.Context(0); t1 = time.perf_counter()
    c.upload_problem([0.001], Ks, var, ref); t2 = time.perf_counter()
    c.score_topk(const, None, None, None, topk=5, precision=cuda_shim.FP32); t3 = time.perf_counter()
    c.upload_problem([0.001], Ks, var, ref); t4 = time.perf_counter()
    c.score_topk(const, None, None, None, topk=5, precision=cuda_shim.FP32); t5 = time.perf_counter()
    c.close(); t6 = time.perf_counter()
    print("create %.1f  upload1 %.1f  topk1 %.1f  upload2 %.1f  topk2 %.1f  close %.1f ms" % tuple(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5)), flush=True)
