"""ctypes binding of libbamse_cuda.so (include/bamse_cuda.h).

This is the only door between the Python host and the CUDA kernels.  There is NO CPU
fallback: if the shared library is missing or no B200 is visible, construction raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# BAMSE_CUDA_LIB: load another build of the library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("BAMSE_CUDA_LIB") or os.path.join(_HERE, "libbamse_cuda.so")

FP32 = 0
FP64 = 1
FP32_LOGDOMAIN = 2
OPT_EARLY_EXIT = 1
OPT_SHARD_PLAN = 2
OPT_STATS = 3
KMAX = 12
TOPK_MAX = 64
GRID = 512

RECORD_DTYPE = np.dtype([("total", "<f8"), ("score", "<f8"), ("clust", "<i4"), ("param", "<i4"), ("tree", "<u8")])
assert RECORD_DTYPE.itemsize == 32


class BamseCudaError(RuntimeError):
    def __init__(self, code, message):
        RuntimeError.__init__(self, "libbamse_cuda error %d: %s" % (code, message))
        self.code = code


_lib = None


def load_library():
    """Loads libbamse_cuda.so once and declares every prototype of include/bamse_cuda.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(make -C bamse_b200/csrc).  There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    c_p = ctypes.c_void_p
    i32, i64, u64, dbl = ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_double
    lib.bamse_abi_version.restype = i32
    lib.bamse_abi_version.argtypes = []
    lib.bamse_create.restype = i32
    lib.bamse_create.argtypes = [ctypes.POINTER(c_p), i32]
    lib.bamse_destroy.restype = None
    lib.bamse_destroy.argtypes = [c_p]
    lib.bamse_last_error.restype = ctypes.c_char_p
    lib.bamse_last_error.argtypes = [c_p]
    lib.bamse_set_stream.restype = i32
    lib.bamse_set_stream.argtypes = [c_p, c_p]
    lib.bamse_set_shard.restype = i32
    lib.bamse_set_shard.argtypes = [c_p, i32, i32]
    lib.bamse_get_shard_plan.restype = i32
    lib.bamse_get_shard_plan.argtypes = [c_p, c_p]
    lib.bamse_fast_variant.restype = i32
    lib.bamse_fast_variant.argtypes = [c_p]
    lib.bamse_set_option.restype = i32
    lib.bamse_set_option.argtypes = [c_p, i32, i32]
    lib.bamse_samples_evaluated.restype = i64
    lib.bamse_samples_evaluated.argtypes = [c_p, i32]
    lib.bamse_last_topk_verified.restype = i32
    lib.bamse_last_topk_verified.argtypes = [c_p]
    lib.bamse_last_topk_preselected.restype = i32
    lib.bamse_last_topk_preselected.argtypes = [c_p]
    lib.bamse_launch_count.restype = i64
    lib.bamse_launch_count.argtypes = [c_p]
    lib.bamse_last_eval_count.restype = i64
    lib.bamse_last_eval_count.argtypes = [c_p]
    lib.bamse_decode_tree.restype = i32
    lib.bamse_decode_tree.argtypes = [i32, u64, c_p]
    lib.bamse_upload_problem.restype = i32
    lib.bamse_upload_problem.argtypes = [c_p, i32, c_p, i32, i32, i32, c_p, c_p, c_p]
    lib.bamse_rebuild_tables.restype = i32
    lib.bamse_rebuild_tables.argtypes = [c_p]
    lib.bamse_score_list.restype = i32
    lib.bamse_score_list.argtypes = [c_p, i64, c_p, c_p, c_p, i32, c_p]
    lib.bamse_score_range.restype = i32
    lib.bamse_score_range.argtypes = [c_p, i32, i32, u64, i64, i32, c_p, c_p]
    lib.bamse_score_topk.restype = i32
    lib.bamse_score_topk.argtypes = [c_p, c_p, c_p, c_p, c_p, i32, i32, c_p, c_p, c_p]
    lib.bamse_enumerate_topk_dev.restype = i32
    lib.bamse_enumerate_topk_dev.argtypes = [c_p, c_p, c_p, c_p, c_p, i32, i32, dbl, c_p]
    lib.bamse_merge_topk_dev.restype = i32
    lib.bamse_merge_topk_dev.argtypes = [c_p, c_p, i32, i32, i32, c_p]
    lib.bamse_kernel_time.restype = i32
    lib.bamse_kernel_time.argtypes = [c_p, i32, ctypes.POINTER(dbl), ctypes.POINTER(i64)]
    lib.bamse_table_time.restype = i32
    lib.bamse_table_time.argtypes = [c_p, i32, ctypes.POINTER(dbl), ctypes.POINTER(i64)]
    lib.bamse_kernel_stats.restype = i32
    lib.bamse_kernel_stats.argtypes = [c_p, i32, c_p]
    lib.bamse_debug_forest_geometry.restype = i32
    lib.bamse_debug_forest_geometry.argtypes = [i32, c_p, c_p, c_p, c_p, c_p]
    lib.bamse_debug_forest_recipes.restype = i32
    lib.bamse_debug_forest_recipes.argtypes = [i32, i32, c_p]
    lib.bamse_debug_program.restype = i32
    lib.bamse_debug_program.argtypes = [i32, c_p, c_p, i32, i32, i32, c_p]
    lib.bamse_debug_pair_geometry.restype = i32
    lib.bamse_debug_pair_geometry.argtypes = [i32, c_p, c_p, c_p, c_p]
    lib.bamse_debug_ctx_recipes.restype = i32
    lib.bamse_debug_ctx_recipes.argtypes = [i32, i32, i32, c_p]
    lib.bamse_debug_classify.restype = i32
    lib.bamse_debug_classify.argtypes = [i32, c_p, i32, i32, c_p]
    lib.bamse_debug_shard_plan.restype = i32
    lib.bamse_debug_shard_plan.argtypes = [i32, i32, c_p, i32, c_p, c_p, c_p]
    lib.bamse_debug_pair_stats.restype = i32
    lib.bamse_debug_pair_stats.argtypes = [c_p, i32, i32, i32, c_p]
    if lib.bamse_abi_version() != 1:
        raise RuntimeError("libbamse_cuda ABI version mismatch")
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    "bamse_abi_version", "bamse_create", "bamse_destroy", "bamse_last_error", "bamse_set_stream", "bamse_set_shard", "bamse_get_shard_plan", "bamse_fast_variant", "bamse_set_option", "bamse_samples_evaluated",
    "bamse_launch_count", "bamse_last_eval_count", "bamse_last_topk_verified", "bamse_last_topk_preselected", "bamse_decode_tree", "bamse_upload_problem", "bamse_rebuild_tables",
    "bamse_score_list", "bamse_score_range", "bamse_score_topk", "bamse_enumerate_topk_dev",
    "bamse_merge_topk_dev", "bamse_kernel_time", "bamse_table_time", "bamse_kernel_stats",
    "bamse_debug_forest_geometry", "bamse_debug_forest_recipes", "bamse_debug_program",
    "bamse_debug_pair_geometry", "bamse_debug_ctx_recipes", "bamse_debug_classify", "bamse_debug_pair_stats", "bamse_debug_shard_plan",
]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def decode_tree(K, index):
    """index -> parent vector (python ints) through the library's host decoder."""
    lib = load_library()
    out = np.empty(K, dtype=np.int8)
    rc = lib.bamse_decode_tree(int(K), int(index), _ptr(out))
    if rc != 0:
        raise BamseCudaError(rc, "bamse_decode_tree(K=%d, index=%d)" % (K, index))
    return [int(x) for x in out]


class Context(object):
    """One bamse_ctx (one CUDA device, one host thread)."""

    def __init__(self, device=0):
        self._lib = load_library()
        h = ctypes.c_void_p()
        rc = self._lib.bamse_create(ctypes.byref(h), int(device))
        if rc != 0:
            raise BamseCudaError(rc, self._lib.bamse_last_error(None).decode())
        self._h = h
        self.device = int(device)
        self.P = self.C = self.M = self.Kmax = 0
        self.K_of_c = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bamse_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise BamseCudaError(rc, self._lib.bamse_last_error(self._h).decode())

    # ---------------------------------------------------------------- plumbing
    def set_stream(self, cuda_stream):
        self._check(self._lib.bamse_set_stream(self._h, ctypes.c_void_p(cuda_stream or 0)))

    def set_shard(self, rank, world):
        """interleaved multi-GPU sharding of every enumeration of this context."""
        self._check(self._lib.bamse_set_shard(self._h, int(rank), int(world)))

    def set_early_exit(self, on):
        """BAMSE_OPT_EARLY_EXIT: exact sample-level early exit in fp32 top-k enumerations."""
        self._check(self._lib.bamse_set_option(self._h, 1, 1 if on else 0))

    def samples_evaluated(self, reset=True):
        return int(self._lib.bamse_samples_evaluated(self._h, 1 if reset else 0))

    @property
    def fast_variant(self):
        """0 = exponential-path kernel, 1 = kernel with the framed path, -1 = undecided."""
        return int(self._lib.bamse_fast_variant(self._h))

    @property
    def launch_count(self):
        return int(self._lib.bamse_launch_count(self._h))

    @property
    def last_eval_count(self):
        return int(self._lib.bamse_last_eval_count(self._h))

    def kernel_time(self, reset=True):
        """(summed device ms of the enumeration kernel, launches) since the last reset."""
        ms, n = ctypes.c_double(), ctypes.c_int64()
        self._check(self._lib.bamse_kernel_time(self._h, 1 if reset else 0, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def table_time(self, reset=True):
        """(summed device ms of the table-building kernels, builds) since the last reset."""
        ms, n = ctypes.c_double(), ctypes.c_int64()
        self._check(self._lib.bamse_table_time(self._h, 1 if reset else 0, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def set_option(self, option, value):
        self._check(self._lib.bamse_set_option(self._h, int(option), int(value)))

    def kernel_stats(self, reset=True):
        """dict of executed-work counters of the fp32 enumeration kernel since the last reset."""
        out = np.zeros(4, dtype=np.uint64)
        self._check(self._lib.bamse_kernel_stats(self._h, 1 if reset else 0, _ptr(out)))
        return dict(mufu_ops=int(out[0]), conv_terms=int(out[1]), convs=int(out[2]), scans=int(out[3]))

    # ---------------------------------------------------------------- problem
    def upload_problem(self, err, K_of_c, var_counts, ref_counts):
        """err: [P] floats; K_of_c: [C]; var/ref_counts: [C][Kmax][M] integer arrays."""
        err = np.ascontiguousarray(np.atleast_1d(err), dtype=np.float64)
        K_of_c = np.ascontiguousarray(K_of_c, dtype=np.int32)
        var = np.ascontiguousarray(var_counts, dtype=np.int64)
        ref = np.ascontiguousarray(ref_counts, dtype=np.int64)
        if var.ndim != 3 or var.shape != ref.shape or var.shape[0] != len(K_of_c):
            raise ValueError("var_counts/ref_counts must be [C][Kmax][M] and match K_of_c")
        C, Kmax, M = var.shape
        self._check(self._lib.bamse_upload_problem(self._h, len(err), _ptr(err), C, M, Kmax, _ptr(K_of_c),
                                                   _ptr(var), _ptr(ref)))
        self.P, self.C, self.M, self.Kmax = len(err), C, M, Kmax
        self.K_of_c = K_of_c.copy()

    def shard_plan(self):
        """[P][C] owning rank of every (parameter set, clustering), -1 = split over all ranks."""
        out = np.empty((self.P, self.C), dtype=np.int32)
        self._check(self._lib.bamse_get_shard_plan(self._h, _ptr(out)))
        return out

    def pair_stats(self, K, s, r):
        """dict(nf, nt_full, nt_used, irregular) of the pair geometry (s, r) at K clusters (cuts every tree on the GPU)."""
        out = np.zeros(4, dtype=np.int64)
        self._check(self._lib.bamse_debug_pair_stats(self._h, int(K), int(s), int(r), _ptr(out)))
        return dict(nf=int(out[0]), nt_full=int(out[1]), nt_used=int(out[2]), irregular=int(out[3]))

    def rebuild_tables(self):
        """re-derives every table of the resident problem from the counts in device memory (asynchronous)."""
        self._check(self._lib.bamse_rebuild_tables(self._h))

    # ---------------------------------------------------------------- scoring
    def score_list(self, item_clust, item_parent, item_nodes, precision=FP64):
        """Returns [P][n] raw get_tree_score values.  item_parent/item_nodes: [n][Kmax] int8
        (parent padding -2, nodes padding -1)."""
        clust = np.ascontiguousarray(item_clust, dtype=np.int32)
        par = np.ascontiguousarray(item_parent, dtype=np.int8)
        nod = np.ascontiguousarray(item_nodes, dtype=np.int8)
        n = len(clust)
        if par.shape != (n, self.Kmax) or nod.shape != (n, self.Kmax):
            raise ValueError("item_parent / item_nodes must be [n][Kmax=%d]" % self.Kmax)
        out = np.empty((self.P, n), dtype=np.float64)
        self._check(self._lib.bamse_score_list(self._h, n, _ptr(clust), _ptr(par), _ptr(nod), int(precision), _ptr(out)))
        return out

    def score_range(self, p, c, tree_begin, n, precision=FP64):
        """Dense scores of trees [tree_begin, tree_begin+n): (get_tree_score, log scre)."""
        score = np.empty(n, dtype=np.float64)
        logscre = np.empty(n, dtype=np.float64)
        self._check(self._lib.bamse_score_range(self._h, int(p), int(c), int(tree_begin), int(n), int(precision),
                                                _ptr(score), _ptr(logscre)))
        return score, logscre

    def score_topk(self, clust_const, threshold=None, tree_begin=None, tree_end=None, topk=1, precision=FP32):
        """Host-buffer enumeration + global top-k.  Returns (records [P][topk] structured array,
        parents [P][topk][Kmax] int8, counts [P])."""
        cc = np.ascontiguousarray(clust_const, dtype=np.float64).reshape(self.P, self.C)
        th = None if threshold is None else np.ascontiguousarray(threshold, dtype=np.float64).reshape(self.P, self.C)
        tb = None if tree_begin is None else np.ascontiguousarray(tree_begin, dtype=np.uint64)
        te = None if tree_end is None else np.ascontiguousarray(tree_end, dtype=np.uint64)
        rec = np.empty((self.P, topk), dtype=RECORD_DTYPE)
        par = np.empty((self.P, topk, self.Kmax), dtype=np.int8)
        cnt = np.empty(self.P, dtype=np.int32)
        self._check(self._lib.bamse_score_topk(self._h, _ptr(cc), _ptr(th), _ptr(tb), _ptr(te), int(topk),
                                               int(precision), _ptr(rec), _ptr(par), _ptr(cnt)))
        if not self.last_topk_verified:
            import warnings
            warnings.warn("libbamse_cuda: more than %d trees are tied within rounding of the top-%d cut; the "
                          "ranking below the ties is best effort" % (self.last_topk_preselected, topk))
        return rec, par, cnt

    def score_list_topk(self, item_clust, item_parent, item_nodes, clust_const, threshold=None, topk=1,
                        precision=FP64):
        """Explicit-list counterpart of score_topk (SURVEY 8b's `K_enum_mode = 0`): the top-`topk` of a GIVEN list of
        full trees (nodes 0..K-1 in order: `scre` depends on the labels) instead of an index range.  The scores come from the GPU (bamse_score_list); `log scre`
        (tree_tools.canonizeF, integer), the constants, the threshold rule (`score >= thr - 1e-12 |thr|`) and the
        order (total desc, clustering asc, position in the list asc) are the ones of bamse_score_topk, in host fp64.
        Returns (records [P][topk] -- `tree` holds the item's position in the list --, counts [P])."""
        from .tree_tools import canonizeF
        clust = np.ascontiguousarray(item_clust, dtype=np.int32)
        par = np.ascontiguousarray(item_parent, dtype=np.int8)
        n = len(clust)
        if n and (clust.min() < 0 or clust.max() >= self.C):
            raise ValueError("score_list_topk: clustering index out of range")
        cc = np.ascontiguousarray(clust_const, dtype=np.float64).reshape(self.P, self.C)
        th = None if threshold is None else np.ascontiguousarray(threshold, dtype=np.float64).reshape(self.P, self.C)
        raw = self.score_list(clust, par, item_nodes, precision) if n else np.empty((self.P, 0))
        logscre = np.empty(n)
        for i in range(n):
            K = int(self.K_of_c[clust[i]])
            tree = [int(x) for x in par[i, :K]]
            if [int(x) for x in np.asarray(item_nodes)[i][:K]] != list(range(K)) or (K < par.shape[1] and par[i, K] != -2):
                raise ValueError("score_list_topk: item %d is not a full tree of its clustering" % i)
            logscre[i] = np.log(canonizeF(tree)['scre'])
        rec = np.zeros((self.P, topk), dtype=RECORD_DTYPE)
        rec["clust"], rec["total"], rec["score"] = -1, -np.inf, -np.inf
        cnt = np.zeros(self.P, dtype=np.int32)
        for p in range(self.P):
            rec["param"][p] = p
            score = raw[p] + logscre
            keep = ~np.isnan(score)
            if th is not None:
                t = th[p, clust]
                keep &= score >= t - 1e-12 * np.abs(t)
            idx = np.nonzero(keep)[0]
            total = score[idx] + cc[p, clust[idx]]
            order = idx[np.lexsort((idx, clust[idx], -total))][:topk]
            cnt[p] = len(order)
            rec["total"][p, :len(order)] = score[order] + cc[p, clust[order]]
            rec["score"][p, :len(order)] = score[order]
            rec["clust"][p, :len(order)] = clust[order]
            rec["tree"][p, :len(order)] = order
        return rec, cnt

    @property
    def last_topk_verified(self):
        """True when the last score_topk proved its preselection cut safe (identical to a full fp64 ranking)."""
        return bool(self._lib.bamse_last_topk_verified(self._h))

    @property
    def last_topk_preselected(self):
        return int(self._lib.bamse_last_topk_preselected(self._h))

    # ---------------------------------------------------------------- device-resident
    def enumerate_topk_dev(self, d_clust_const, d_threshold, d_tree_begin, d_tree_end, topk, precision, slack,
                           d_out_records):
        """All d_* are integer device addresses (e.g. torch.Tensor.data_ptr()); asynchronous."""
        self._check(self._lib.bamse_enumerate_topk_dev(
            self._h, ctypes.c_void_p(d_clust_const or 0), ctypes.c_void_p(d_threshold or 0),
            ctypes.c_void_p(d_tree_begin or 0), ctypes.c_void_p(d_tree_end or 0), int(topk), int(precision),
            float(slack), ctypes.c_void_p(d_out_records)))

    def merge_topk_dev(self, d_lists, n_lists, P, topk, d_out_records):
        self._check(self._lib.bamse_merge_topk_dev(self._h, ctypes.c_void_p(d_lists), int(n_lists), int(P), int(topk),
                                                   ctypes.c_void_p(d_out_records)))
