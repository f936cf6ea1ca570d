"""Dry run of bench.py's own arm WITHOUT a GPU: the CUDA runtime and the library's context are replaced by
stand-ins that return made-up numbers, so that every line of the driver-facing script executes here (argument
defaults, leg order, counters, JSON assembly).  It checks the SHAPE of the bench line, nothing about performance or
parity -- the numbers it prints are meaningless by construction, and the line is marked as a dry run.

    python tests/bench_dryrun.py [bench.py arguments]

Used by tests/test_host_cpu.py::test_bench_script_executes_end_to_end_with_stand_ins.
"""
import os
import runpy
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bamse_b200 import cuda_shim  # noqa: E402

_real_device = torch.device


class _Stream(object):
    cuda_stream = 1

    def __init__(self, *a, **k):
        pass


class _Event(object):
    def __init__(self, *a, **k):
        self.t = None

    def record(self, *a):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max(1e-3, 1e3 * (other.t - self.t))


class _Context(object):
    """stand-in for cuda_shim.Context: shapes and call order only."""
    n_created = 0

    def __init__(self, device=0):
        _Context.n_created += 1
        self.P = self.C = self.M = self.Kmax = 0
        self.launch_count = 0
        self.fast_variant = 0
        self.stats_on = False
        self.last_tables = True
        self._tab = [0.0, 0]
        self._ker = [0.0, 0]
        self._stat = np.zeros(4, dtype=np.uint64)

    def close(self):
        pass

    def set_stream(self, s):
        pass

    def set_shard(self, rank, world):
        assert 0 <= rank < world

    def set_early_exit(self, on):
        pass

    def samples_evaluated(self, reset=True):
        return 1

    def set_option(self, option, value):
        if option == cuda_shim.OPT_STATS:
            self.stats_on = bool(value)

    def kernel_time(self, reset=True):
        out = tuple(self._ker)
        if reset:
            self._ker = [0.0, 0]
        return out

    def table_time(self, reset=True):
        out = tuple(self._tab)
        if reset:
            self._tab = [0.0, 0]
        return out

    def kernel_stats(self, reset=True):
        out = dict(mufu_ops=int(self._stat[0]), conv_terms=int(self._stat[1]), convs=int(self._stat[2]), scans=int(self._stat[3]))
        if reset:
            self._stat[:] = 0
        return out

    def upload_problem(self, err, K_of_c, var, ref):
        var = np.asarray(var)
        assert var.ndim == 3 and var.shape == np.asarray(ref).shape and var.shape[0] == len(K_of_c)
        self.P, (self.C, self.Kmax, self.M) = len(np.atleast_1d(err)), var.shape

    def rebuild_tables(self):
        self._tab[0] += 3.0
        self._tab[1] += 1
        self.launch_count += 8
        if self.stats_on:
            self._stat += np.array([900, 800, 7, 3], dtype=np.uint64)

    def score_list(self, clust, par, nod, precision=cuda_shim.FP64):
        n = len(clust)
        assert np.asarray(par).shape == (n, self.Kmax) and np.asarray(nod).shape == (n, self.Kmax)
        rng = np.random.default_rng(n)
        return -100.0 - rng.random((self.P, n))

    def score_topk(self, const, thr=None, tb=None, te=None, topk=1, precision=cuda_shim.FP32):
        assert np.asarray(const).size == self.P * self.C
        rec = np.zeros((self.P, topk), dtype=cuda_shim.RECORD_DTYPE)
        rec["total"] = -1.0
        return rec, np.zeros((self.P, topk, self.Kmax), dtype=np.int8), np.full(self.P, topk, dtype=np.int32)

    def enumerate_topk_dev(self, d_const, d_thr, d_tb, d_te, topk, precision, slack, d_out):
        self._ker[0] += 1.0
        self._ker[1] += 1
        self.launch_count += 3
        self._stat += np.array([100, 0, 0, 0], dtype=np.uint64)

    def merge_topk_dev(self, *a):
        self.launch_count += 1


def main():
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.Stream = _Stream
    torch.cuda.set_stream = lambda *a, **k: None
    torch.cuda.current_stream = lambda *a, **k: _Stream()
    torch.cuda.Event = _Event
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    torch.device = lambda *a, **k: _real_device("cpu")
    cuda_shim.Context = _Context
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:           # under torchrun: gloo stands in for NCCL
        import torch.distributed as dist
        real_init = dist.init_process_group
        dist.init_process_group = lambda backend=None, **k: real_init("gloo")
    sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[1:]
    print("DRY RUN (stand-ins for the CUDA runtime and the library context; numbers are meaningless)", file=sys.stderr)
    runpy.run_path(os.path.join(ROOT, "bench.py"), run_name="__main__")


if __name__ == "__main__":
    main()
