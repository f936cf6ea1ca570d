"""Host-side mirror of the reference's `clustering` module for the tree-scoring hot path.

Same names and argument meaning as the reference (clustering.py) so that tests and the
CLI read like the reference's own code, but every marginal-likelihood evaluation goes
through libbamse_cuda (hand-written sm_100a kernels) -- there is no CPU scoring code
in this package.  What stays on the host, in fp64 Python, is what the north star keeps
there: read-count aggregation, the per-clustering constants that enter the ranking key
(maxlikelihood, cluster_prior), the comparator-driven relabelling and the greedy
candidate's control flow (their scores are batched GPU calls).

`analyse_clusterings` is the batch-level replacement of the loop bamse.py:91-106.
"""
import functools
import math
import warnings

import numpy as np
from scipy.special import gammaln

from . import cuda_shim
from .tree_tools import canonizeF, get_matrix

# reference clustering.py:16-18 (number of unlabeled rooted trees; not A000081 from index 7 on)
tree_numbers = np.array([1, 1, 1, 2, 4, 9, 20, 47, 108, 252, 582, 1345, 3086, 7072, 16121, 36667, 83099, 187885,
                         423610, 953033, 2139158, 4792126, 10714105, 23911794, 53273599, 118497834, 263164833,
                         583582570])

_default_ctx = None


def get_context(device=0):
    """process-wide default CUDA context (created on first use; raises without a B200)."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = cuda_shim.Context(device)
    return _default_ctx


def part_func(n, kmax):
    """reference clustering.py:397-408."""
    table = [np.ones(min(k, kmax)) for k in range(1, n + 1)]
    if n >= 4:
        for m in range(4, n + 1):
            row, prev = table[m - 1], table[m - 2]
            for k in range(1, min(m, kmax)):
                row[k] = prev[k - 1]
                if m - 1 - k > k:
                    row[k] += table[m - 2 - k][k]
    return table[n - 1]


def log_conf_penalty(assign, pf):
    """model prior of a clustering (reference clustering.py:32-35)."""
    sizes = np.unique(assign, return_counts=True)[1]
    shape = np.unique(sizes, return_counts=True)[1]
    k = len(sizes)
    return (np.sum(gammaln(sizes + 1.)) - gammaln(np.sum(sizes) + 1) + np.sum(gammaln(shape + 1.))
            - np.log(pf[k - 1]) + np.log(tree_numbers[k]) - (k - 1) * np.log(k))


def _aggregate(am, bm, assign, K):
    As = np.array([np.sum(am[assign == x], 0) for x in range(K)])
    Bs = np.array([np.sum(bm[assign == x], 0) for x in range(K)])
    return As, Bs


def _pack_problem(clusts):
    Kmax = max(c.As.shape[0] for c in clusts)
    M = clusts[0].As.shape[1]
    var = np.zeros((len(clusts), Kmax, M), dtype=np.int64)
    ref = np.zeros((len(clusts), Kmax, M), dtype=np.int64)
    Ks = np.zeros(len(clusts), dtype=np.int32)
    for i, c in enumerate(clusts):
        K = c.As.shape[0]
        if c.As.shape[1] != M:
            raise ValueError("all clusterings of a batch must have the same samples")
        var[i, :K], ref[i, :K], Ks[i] = c.As, c.Bs, K
    return Ks, var, ref, Kmax


def _pack_items(Kmax, items):
    n = len(items)
    clust = np.zeros(n, dtype=np.int32)
    par = np.full((n, Kmax), -2, dtype=np.int8)
    nod = np.full((n, Kmax), -1, dtype=np.int8)
    for i, (c, tree, nodes) in enumerate(items):
        clust[i] = c
        par[i, :len(tree)] = np.asarray(tree, dtype=np.int8)
        nod[i, :len(nodes)] = np.asarray(nodes, dtype=np.int8)
    return clust, par, nod


class clustering:
    """A clustering of mutations with its error model (reference clustering.py:39-183)."""

    def __init__(self, variant_reads, normal_reads, assignments, pf, err, ctx=None):
        self.am = np.array(variant_reads)
        self.bm = np.array(normal_reads)
        self.assign = np.array(assignments).astype(int)
        self.pf = pf
        self.err = err
        self._ctx = ctx
        K = int(max(assignments)) + 1
        self.As, self.Bs = _aggregate(self.am, self.bm, self.assign, K)
        # log maximum likelihood of the data for this clustering (reference clustering.py:66-68)
        self.maxlikelihood = np.sum(gammaln(self.am + self.bm + 2.) - gammaln(self.am + 1.) - gammaln(self.bm + 1.)) \
            - np.sum(gammaln(self.As + self.Bs + 2) - gammaln(self.As + 1) - gammaln(self.Bs + 1))
        self.cluster_prior = log_conf_penalty(self.assign, self.pf)
        self.order = list(range(K))
        self.trees = []

    # -------------------------------------------------------------- GPU residency
    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = get_context()
        return self._ctx

    def _make_resident(self):
        ctx = self.ctx
        token = (id(self), self.As.tobytes(), self.Bs.tobytes(), float(self.err))
        if getattr(ctx, "_resident_token", None) != token:
            Ks, var, ref, _ = _pack_problem([self])
            ctx.upload_problem([self.err], Ks, var, ref)
            ctx._resident_token = token
        return ctx

    def _scores(self, items, precision=cuda_shim.FP64):
        """items: [(tree, nodes)] -> fp64 array of get_tree_score values (one GPU batch)."""
        ctx = self._make_resident()
        clust, par, nod = _pack_items(ctx.Kmax, [(0, t, n) for t, n in items])
        return ctx.score_list(clust, par, nod, precision)[0]

    # -------------------------------------------------------------- reference API
    def get_tree_score(self, tree, nodes):
        """log marginal likelihood of `tree` over clusters `nodes` (reference clustering.py:87-90)."""
        return float(self._scores([(list(tree), list(nodes))])[0])

    def compare_clusters(self, clustA, clustB):
        """reference clustering.py:92-96."""
        s = self._scores([([-1, 0], [clustB, clustA]), ([-1, 0], [clustA, clustB])])
        return int(np.sign(s[0] - s[1]))

    def _pair_items(self):
        K = self.As.shape[0]
        return [([-1, 0], [a, b]) for a in range(K) for b in range(K) if a != b]

    def _apply_order(self, pair_scores):
        """sorted(..., key=cmp_to_key(compare_clusters)) over a pre-computed score table."""
        K = self.As.shape[0]
        table = {}
        i = 0
        for a in range(K):
            for b in range(K):
                if a != b:
                    table[(a, b)] = pair_scores[i]
                    i += 1

        def cmp(A, B):
            return int(np.sign(table[(B, A)] - table[(A, B)])) if A != B else 0

        self.order = sorted(np.unique(self.assign), key=functools.cmp_to_key(cmp))
        self.assign = np.array([self.order.index(x) for x in self.assign])
        self.As, self.Bs = _aggregate(self.am, self.bm, self.assign, K)
        self.means = (self.As + 0.) / (self.As + self.Bs)

    def reorder(self):
        """relabel clusters by pairwise parent-likelihood (reference clustering.py:72-85); all
        K(K-1) two-node scores are one GPU batch, then the same Python `sorted`."""
        self._apply_order(self._scores(self._pair_items()))

    def get_candidate_tree(self):
        """greedy initial tree (reference clustering.py:104-114); each round is one GPU batch."""
        tree = [-1]
        for n in range(1, len(self.order)):
            s = self._scores([(tree + [x], list(range(n + 1))) for x in range(n)])
            tree = tree + [int(np.argmax(s))]
        self.trees = [{'tree': tree, 'root': 0}]
        return [{'tree': tree, 'root': 0}]

    def get_trees2(self, candidate_tree, max_trees=cuda_shim.TOPK_MAX, precision=cuda_shim.FP32):
        """every labeled rooted tree with score + log scre >= the candidate's (the set the
        reference's branch-and-bound, clustering.py:116-148, searches for), found by
        exhaustive GPU enumeration.  Deviation from the reference: at most `max_trees`
        (<= cuda_shim.TOPK_MAX = 64) best ones are returned, best first; a warning says when
        the list was cut (the reference returns every qualifying tree, in search order)."""
        if not 1 <= int(max_trees) <= cuda_shim.TOPK_MAX:
            raise ValueError("get_trees2: max_trees must be in [1, %d]" % cuda_shim.TOPK_MAX)
        K = self.As.shape[0]
        cand = list(candidate_tree['tree'])
        thr = self.get_tree_score(cand, list(range(K))) + np.log(canonizeF(cand)['scre'])
        ctx = self._make_resident()
        rec, par, cnt = ctx.score_topk(np.zeros((1, 1)), np.array([[thr]]), topk=int(max_trees),
                                       precision=precision)
        if int(cnt[0]) == int(max_trees):
            warnings.warn("get_trees2: %d or more trees reach the candidate's score; only the best %d are kept"
                          % (int(max_trees), int(max_trees)))
        self.trees = []
        for j in range(int(cnt[0])):
            tree = [int(x) for x in par[0, j, :K]]
            self.trees.append({'tree': tree, 'root': tree.index(-1), 'score': float(rec[0, j]['score'])})
        return self.trees

    def analyse_trees_regular2(self, solve_fractions=True):
        """tree matrices, score assembly and ML clone fractions (reference clustering.py:170-183)."""
        K = self.As.shape[0]
        missing = [t for t in self.trees if 'score' not in t]
        if missing:
            s = self._scores([(t['tree'], list(range(K))) for t in missing])
            for t, v in zip(missing, s):
                t['score'] = float(v) + math.log(canonizeF(t['tree'])['scre'])
        for tree in self.trees:
            B = get_matrix(tree['tree'])[0]
            tree['Amatrix'] = np.linalg.inv(B)
            tree['Bmatrix'] = B
            tree['assign'] = self.assign
            if solve_fractions:
                from .fractions import get_max_values_r2
                get_max_values_r2(tree, self.As, self.Bs, self.err)
                tree['vaf'] = tree['Bmatrix'].dot(tree['clone_proportions'])
            tree['totalscore'] = tree['score'] + self.cluster_prior + self.maxlikelihood


# ------------------------------------------------------------------ batch driver
def analyse_clusterings(clusts, top_trees, ctx=None, precision=cuda_shim.FP32, solve_fractions=True,
                        tree_shard=None):
    """Replaces the loop bamse.py:91-106 for a list of `clustering` objects that share one
    error rate: reorder -> greedy candidate -> threshold for all clusterings in batched GPU
    calls, then ONE enumeration of every labeled rooted tree of every clustering with a
    global top-`top_trees`.  Returns the ranked list of tree dicts (reference keys).
    tree_shard = (rank, world) restricts the enumeration to this rank's interleaved share."""
    if not clusts:
        return []
    ctx = ctx or get_context()
    err = clusts[0].err
    if any(c.err != err for c in clusts):
        raise ValueError("analyse_clusterings: one error rate per batch")

    def upload():
        Ks, var, ref, Kmax = _pack_problem(clusts)
        ctx.upload_problem([err], Ks, var, ref)
        ctx._resident_token = None
        return Kmax

    # 1. reorder(): all two-node scores of all clusterings in one call
    Kmax = upload()
    items, spans = [], []
    for ci, c in enumerate(clusts):
        pi = c._pair_items()
        spans.append((len(items), len(pi)))
        items += [(ci, t, n) for t, n in pi]
    if items:
        s = ctx.score_list(*_pack_items(Kmax, items), cuda_shim.FP64)[0]
    for c, (o, n) in zip(clusts, spans):
        c._apply_order(s[o:o + n] if n else [])
    # 2. greedy candidates, one batched call per round (counts changed -> upload again)
    Kmax = upload()
    cand = [[-1] for _ in clusts]
    for n in range(1, Kmax):
        items, owners = [], []
        for ci, c in enumerate(clusts):
            if c.As.shape[0] > n:
                owners.append((ci, len(items)))
                items += [(ci, cand[ci] + [x], list(range(n + 1))) for x in range(n)]
        s = ctx.score_list(*_pack_items(Kmax, items), cuda_shim.FP64)[0]
        for ci, o in owners:
            cand[ci] = cand[ci] + [int(np.argmax(s[o:o + n]))]
    # 3. thresholds = candidate score + log scre (clustering.py:121-122)
    items = [(ci, cand[ci], list(range(c.As.shape[0]))) for ci, c in enumerate(clusts)]
    s = ctx.score_list(*_pack_items(Kmax, items), cuda_shim.FP64)[0]
    thr = np.array([[s[ci] + math.log(canonizeF(cand[ci])['scre']) for ci in range(len(clusts))]])
    const = np.array([[c.cluster_prior + c.maxlikelihood for c in clusts]])
    # 4. exhaustive enumeration + global top-k
    if int(top_trees) > cuda_shim.TOPK_MAX:
        warnings.warn("analyse_clusterings: -t %d exceeds the scorer's top-k capacity; returning the best %d"
                      % (int(top_trees), cuda_shim.TOPK_MAX))
    if tree_shard is not None:
        ctx.set_shard(*tree_shard)
    # the search only needs the trees that can still qualify: exact sample-level early exit
    ctx.set_early_exit(True)
    try:
        rec, par, cnt = ctx.score_topk(const, thr, None, None, topk=min(max(int(top_trees), 1), cuda_shim.TOPK_MAX),
                                       precision=precision)
    finally:
        ctx.set_early_exit(False)
        if tree_shard is not None:
            ctx.set_shard(0, 1)          # the context is process-wide: never leave it sharded
    out = []
    for c in clusts:
        c.trees = []
    for j in range(int(cnt[0])):
        c = clusts[int(rec[0, j]['clust'])]
        K = c.As.shape[0]
        tree = [int(x) for x in par[0, j, :K]]
        t = {'tree': tree, 'root': tree.index(-1), 'score': float(rec[0, j]['score']),
             'tree_index': int(rec[0, j]['tree']), 'clustering_index': int(rec[0, j]['clust'])}
        c.trees.append(t)                # a clustering keeps all of its ranked trees
        out.append(t)
    for c in clusts:
        if c.trees:
            c.analyse_trees_regular2(solve_fractions=solve_fractions)
    for c, cd in zip(clusts, cand):
        c.candidate = cd
    return out
