"""Multi-GPU plumbing: one process (rank) per GPU, trees sharded by interleaved index, one
all-gather of each rank's top-k records, identical deterministic merge on every rank.

torch.distributed is used only as the transport (NCCL over NVLink on the GPU box, gloo in
the CPU tests); no score is ever computed here.
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
