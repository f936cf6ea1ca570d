"""Multi-GPU plumbing.

Two ways to use several GPUs of one box, both on the library's sharding plan (bamse_set_shard before the
upload: whole (parameter set, clustering) pairs per rank by estimated cost, the most expensive ones split over all
ranks: by pack node sets on the pair path, by interleaved tree index otherwise) and the same deterministic merge of the per-rank top-k records:

* one process (rank) per GPU under torchrun, one NCCL all-gather of each rank's records (bench.py);
  torch.distributed is only the transport (NCCL over NVLink on the GPU box, gloo in the CPU tests);
* ONE process driving all GPUs -- `MultiContext`, what the drop-in command line uses (`bamse.py --gpus N`;
  the reference's bamse.py:91-106 is a single-process loop): one library context per device, one host thread per
  context for the blocking C-ABI calls (ctypes releases the GIL), records merged on the host.  No NCCL needed:
  the exchange is t records per GPU.

No score is ever computed here.
"""
import numpy as np

from .cuda_shim import RECORD_DTYPE


def shard_size(T, rank, world):
    """number of tree indices of a T-tree clustering that `rank` of `world` scores under the library's
    interleaved sharding (bamse_set_shard): indices rank, rank + world, rank + 2 world, ... below T."""
    T, rank, world = int(T), int(rank), int(world)
    return (T - rank + world - 1) // world if T > rank else 0


def shard_indices(T, rank, world):
    """the tree indices themselves (tests / small T only).  Interleaving, not contiguous ranges: the high
    Pruefer digits select the tree shape, so contiguous ranges differ in cost (low indices are star-like)."""
    return np.arange(int(rank), int(T), int(world), dtype=np.uint64)


def merge_records(lists, topk):
    """lists: array [G][P][k] of bamse_record -> [P][topk]; order: total desc, clust asc,
    tree asc; invalid records (clust < 0 or NaN total) dropped; unused slots clust = -1.
    Same rule as the device merge kernel (k_merge_topk)."""
    lists = np.asarray(lists, dtype=RECORD_DTYPE)
    G, P, k = lists.shape
    out = np.zeros((P, topk), dtype=RECORD_DTYPE)
    out["clust"] = -1
    out["total"] = -np.inf
    out["score"] = -np.inf
    for p in range(P):
        cand = lists[:, p, :].reshape(-1)
        cand = cand[(cand["clust"] >= 0) & ~np.isnan(cand["total"])]
        order = np.lexsort((cand["tree"], cand["clust"], -cand["total"]))
        sel = cand[order[:topk]]
        out[p, :len(sel)] = sel
        out["param"][p] = p
    return out


def allgather_records(local, group=None):
    """all-gathers a [P][k] record array over the default (or given) process group and
    returns [world][P][k].  Works on gloo (CPU tensors) and nccl (CUDA tensors)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    local = np.ascontiguousarray(local, dtype=RECORD_DTYPE)
    words = torch.from_numpy(local.view(np.int64).reshape(-1).copy())
    backend = dist.get_backend(group)
    if backend == "nccl":
        words = words.cuda()
    gathered = torch.empty(world * words.numel(), dtype=torch.int64, device=words.device)
    dist.all_gather_into_tensor(gathered, words, group=group)
    arr = gathered.cpu().numpy().view(RECORD_DTYPE)
    return arr.reshape((world,) + local.shape)


def allgather_topk(local, topk, group=None):
    """local top-k records of this rank -> global top-k (identical on every rank)."""
    return merge_records(allgather_records(local, group), topk)


class MultiContext(object):
    """N library contexts (one per device ordinal in `devices`; an ordinal may repeat: contexts are independent)
    behind the interface of cuda_shim.Context that the host pipeline uses (clustering.analyse_clusterings):
    upload_problem / score_list / score_topk / set_early_exit.  Context r is rank r of N of the sharding plan."""

    def __init__(self, devices):
        from concurrent.futures import ThreadPoolExecutor
        from . import cuda_shim
        self._shim = cuda_shim
        self.devices = list(devices)
        if not self.devices:
            raise ValueError("MultiContext: no devices")
        self.ctxs = [cuda_shim.Context(d) for d in self.devices]
        for r, c in enumerate(self.ctxs):
            c.set_shard(r, len(self.ctxs))
        self._pool = ThreadPoolExecutor(max_workers=len(self.ctxs))
        self._resident_token = None

    # -- the attributes the host pipeline reads
    P = property(lambda self: self.ctxs[0].P)
    C = property(lambda self: self.ctxs[0].C)
    M = property(lambda self: self.ctxs[0].M)
    Kmax = property(lambda self: self.ctxs[0].Kmax)
    K_of_c = property(lambda self: self.ctxs[0].K_of_c)

    def _all(self, fn):
        return [f.result() for f in [self._pool.submit(fn, c) for c in self.ctxs]]

    def upload_problem(self, err, K_of_c, var_counts, ref_counts):
        self._all(lambda c: c.upload_problem(err, K_of_c, var_counts, ref_counts))

    def score_list(self, item_clust, item_parent, item_nodes, precision=None):
        """explicit items (reorder, greedy candidates, thresholds: a few hundred small trees) on the first GPU;
        the check-mode tables every rank holds cover all clusterings."""
        precision = self._shim.FP64 if precision is None else precision
        if precision == self._shim.FP32:
            raise ValueError("MultiContext.score_list: fp32 list scoring needs the tables of every clustering on one "
                             "GPU; use the check mode (FP64) for explicit items")
        return self.ctxs[0].score_list(item_clust, item_parent, item_nodes, precision)

    def set_early_exit(self, on):
        for c in self.ctxs:
            c.set_early_exit(on)

    def set_shard(self, rank, world):
        if (rank, world) != (0, 1):
            raise ValueError("MultiContext shards over its own devices")

    def score_topk(self, clust_const, threshold=None, tree_begin=None, tree_end=None, topk=1, precision=0):
        parts = self._all(lambda c: c.score_topk(clust_const, threshold, tree_begin, tree_end, topk=topk, precision=precision))
        rec = merge_records(np.stack([p[0] for p in parts]), topk)
        P, Kmax = rec.shape[0], self.Kmax
        par = np.full((P, topk, Kmax), -2, dtype=np.int8)
        cnt = np.zeros(P, dtype=np.int32)
        for p in range(P):
            for j in range(topk):
                if rec[p, j]["clust"] >= 0:
                    K = int(self.K_of_c[int(rec[p, j]["clust"])])
                    par[p, j, :K] = self._shim.decode_tree(K, int(rec[p, j]["tree"]))
                    cnt[p] += 1
        return rec, par, cnt

    @property
    def last_topk_verified(self):
        return all(c.last_topk_verified for c in self.ctxs)

    @property
    def last_eval_count(self):
        return sum(c.last_eval_count for c in self.ctxs)

    def close(self):
        for c in self.ctxs:
            c.close()
        self._pool.shutdown()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
