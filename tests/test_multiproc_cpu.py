"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: tree-range sharding, the
all-gather of per-rank top-k records and the deterministic merge."""
import os
import socket

import numpy as np
import pytest

from bamse_b200 import multigpu
from bamse_b200.cuda_shim import RECORD_DTYPE


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_scores(K_of_c, P):
    """deterministic pseudo-scores for every (p, c, tree)."""
    out = []
    for p in range(P):
        for c, K in enumerate(K_of_c):
            T = K ** (K - 1)
            t = np.arange(T, dtype=np.uint64)
            s = -100.0 - ((t * np.uint64(2654435761) + np.uint64(97 * c + 13 * p)) % np.uint64(1000)).astype(np.float64) / 7.0
            out.append((p, c, t, s))
    return out


def _local_topk(K_of_c, P, rank, world, topk):
    rec = np.zeros((P, topk), dtype=RECORD_DTYPE)
    rec["clust"] = -1
    for p in range(P):
        rows = []
        for (pp, c, t, s) in _fake_scores(K_of_c, P):
            if pp != p:
                continue
            sel = np.isin(t, multigpu.shard_indices(len(t), rank, world))
            for ti, si in zip(t[sel], s[sel]):
                rows.append((si + c, si, c, p, ti))
        rows.sort(key=lambda r: (-r[0], r[2], r[4]))
        for j, r in enumerate(rows[:topk]):
            rec[p, j] = r
    return rec


def _worker(rank, world, port, K_of_c, P, topk, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        local = _local_topk(K_of_c, P, rank, world, topk)
        merged = multigpu.allgather_topk(local, topk)
        q.put((rank, merged.tobytes()))
    finally:
        dist.destroy_process_group()


def test_interleaved_shards_partition_the_index_space():
    """bamse_set_shard's scheme: rank r takes r, r + W, r + 2W, ...; the shards are disjoint, cover every
    index once, and their sizes differ by at most one."""
    Ks = [1, 2, 3, 5, 7]
    for world in (1, 2, 3, 4, 8):
        for K in Ks:
            T = K ** (K - 1)
            parts = [multigpu.shard_indices(T, r, world) for r in range(world)]
            assert [len(p) for p in parts] == [multigpu.shard_size(T, r, world) for r in range(world)]
            allidx = np.concatenate(parts)
            assert len(allidx) == T and len(np.unique(allidx)) == T
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert multigpu.shard_size(12 ** 11, 7, 8) == (12 ** 11 - 7 + 7) // 8


def test_merge_records_order_and_invalid_slots():
    lists = np.zeros((2, 1, 3), dtype=RECORD_DTYPE)
    lists["clust"] = -1
    lists[0, 0, 0] = (5.0, 5.0, 1, 0, 7)
    lists[0, 0, 1] = (4.0, 4.0, 0, 0, 9)
    lists[1, 0, 0] = (5.0, 5.0, 0, 0, 8)      # tie on total: lower clust first
    lists[1, 0, 1] = (5.0, 5.0, 1, 0, 3)      # tie on total and clust: lower tree first
    lists[1, 0, 2] = (np.nan, 0.0, 2, 0, 1)   # NaN totals are dropped
    out = multigpu.merge_records(lists, 4)
    assert [(int(r["clust"]), int(r["tree"])) for r in out[0]] == [(0, 8), (1, 3), (1, 7), (0, 9)]
    out = multigpu.merge_records(lists, 6)
    assert int(out[0, 4]["clust"]) == -1 and out[0, 4]["total"] == -np.inf


@pytest.mark.timeout(120)
def test_two_rank_allgather_topk_equals_single_rank():
    import torch.multiprocessing as mp
    K_of_c, P, topk, world = [3, 4, 5], 2, 6, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, K_of_c, P, topk, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=90) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    want = _local_topk(K_of_c, P, 0, 1, topk)           # one rank owning every tree
    for r in range(world):
        merged = np.frombuffer(got[r], dtype=RECORD_DTYPE).reshape(P, topk)
        for f in ("total", "score", "clust", "tree"):
            np.testing.assert_array_equal(merged[f], want[f])


@pytest.mark.timeout(300)
def test_bench_script_under_torchrun_with_two_ranks_and_stand_ins():
    """bench.py as the driver launches it for N > 1 (torch.distributed.run, one rank per GPU), on the stand-ins of
    tests/bench_dryrun.py with gloo in place of NCCL: both ranks walk the same legs (every collective is matched, nobody
    hangs), rank 0 alone prints the one line, and the line says n_gpus = 2.  Shape only."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(root, "tests", "bench_dryrun.py"), "--gpus", "2", "--config", "c1",
           "--steps", "2", "--warmup", "3"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=280)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["steps"] == 2 and d["warmup"] == 3 and d["scaling"] == "strong"
    assert "cpu_baseline" not in d and "also_measured" not in d          # rank 0 at N = 1 only
    assert d["config"]["parallelism"].startswith("x2:")
    assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0
