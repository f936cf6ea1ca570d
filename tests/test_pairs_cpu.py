"""CPU tests of the inside x outside factorisation (bamse_b200/csrc/pairs.cuh): context recipes and the cut of
every tree come out of libbamse_cuda's HOST instances of the code the kernels run (bamse_debug_* entry points,
no GPU needed); the tables are built here in fp64 with the numpy oracle's operators, and the dot product of the
two rows a tree is cut into must reproduce the oracle's tree score."""
import ctypes
import math

import numpy as np
import pytest
from scipy.special import logsumexp

import bamse_oracle as orc
from bamse_b200 import cuda_shim
from test_forest_cpu import Tables, _tables, recipes

LOGL = math.log(512)


def pair_geometry(lib, K, s=0, r=-1):
    a, b, nf, nt = ctypes.c_int32(s), ctypes.c_int32(r), ctypes.c_int32(), ctypes.c_int32()
    assert lib.bamse_debug_pair_geometry(K, ctypes.byref(a), ctypes.byref(b), ctypes.byref(nf), ctypes.byref(nt)) == 0
    return a.value, b.value, nf.value, nt.value


def ctx_recipes(lib, K, s, r, nt):
    out = np.zeros((nt, 3), dtype=np.int32)
    assert lib.bamse_debug_ctx_recipes(K, s, r, out.ctypes.data_as(ctypes.c_void_p)) == 0
    return out


def classify(lib, tree, s, r):
    par = np.asarray(tree, dtype=np.int8)
    out = np.zeros(3, dtype=np.int32)
    assert lib.bamse_debug_classify(len(tree), par.ctypes.data_as(ctypes.c_void_p), s, r, out.ctypes.data_as(ctypes.c_void_p)) == 0
    return bool(out[0]), int(out[1]), int(out[2])


class Contexts(object):
    """numpy build of the context rows T for one sample, following the recipes (rows with fewer nodes first:
    a recipe's parent row always has a smaller index base, so evaluate lazily with memoisation)."""

    def __init__(self, D, p0, FG, rec):
        self.D, self.p0, self.FG, self.rec = D, p0, FG, rec
        self.rows = {}

    def row(self, i):
        if i in self.rows:
            return self.rows[i]
        par, h, v = (int(x) for x in self.rec[i])
        L = self.D.shape[1]
        o = np.full(L, -LOGL) if par < 0 else self.row(par)          # outside message at the marked node
        a = np.logaddexp.accumulate((o + self.D[v])[::-1])[::-1]     # A[y] = sum_{x >= y} O[x] d[x]
        if h < 0:
            t = a + self.p0 - LOGL
        else:
            g = self.FG[h]
            t = np.array([logsumexp(g[:L - z] + a[z:]) for z in range(L)]) + self.p0 - 2 * LOGL
        self.rows[i] = t
        return t


@pytest.mark.parametrize("K,s,r", [(2, 1, 1), (3, 2, 1), (4, 2, 2), (5, 3, 3), (5, 3, 2), (6, 4, 3), (6, 3, 3), (6, 3, 4), (7, 4, 4),
                                   (7, 4, 3), (8, 4, 4)])
def test_pairs_reproduce_the_oracle(built_library, K, s, r):
    """all trees for K <= 5, a stratified sample above: log sum 2^(FG[pack] + T[context]) == oracle score."""
    lib = cuda_shim.load_library()
    s, r, nf, nt = pair_geometry(lib, K, s, r)
    rec = recipes(lib, K, s, nf)
    crec = ctx_recipes(lib, K, s, r, nt)
    assert (crec[:, 0] < np.arange(nt)).all()            # a parent context always sits in an earlier row
    assert (crec[:, 1] < nf).all()
    D = _tables(K, 7 * K + s)
    p0 = orc.prior_const(K)
    tab = Tables(D, p0, rec, 0)
    ctxs = Contexts(D, p0, tab.FG, crec)
    T = K ** (K - 1)
    idx = range(T) if T <= 700 else list(range(0, T, max(1, T // 200))) + [T - 1]
    n_reg = 0
    memo = {}                        # the oracle's sub-tree messages for THIS D (pure-function cache, oracle/bamse_oracle.py)
    for i in idx:
        tree = orc.decode_tree(K, i)
        reg, fg, t = classify(lib, tree, s, r)
        if not reg:
            continue
        n_reg += 1
        assert 0 <= fg < nf and 0 <= t < nt
        want = orc.tree_score_samples(tree, D[:, None, :], p0, memo)[0]
        got = logsumexp(tab.FG[fg] + ctxs.row(t))
        assert got == pytest.approx(want, rel=1e-9), (K, s, r, i, tree, fg, t)
    if (K, s, r) in ((2, 1, 1), (3, 2, 1), (4, 2, 2), (5, 3, 3), (6, 4, 3), (6, 3, 4), (7, 4, 4)):
        assert n_reg == len(idx)                          # these geometries cut every tree
    else:
        assert n_reg > 0.85 * len(idx)


@pytest.mark.parametrize("K", [9, 10])
def test_pair_cuts_for_large_k(built_library, K):
    """default geometry at K = 9, 10: most random trees are regular, rows are in range, the pack has <= s
    nodes and the context <= r nodes (checked through the recipes' node counts), extreme shapes included."""
    lib = cuda_shim.load_library()
    s, r, nf, nt = pair_geometry(lib, K)
    assert nf < (1 << 20)
    rec = recipes(lib, K, s, nf)
    rng = np.random.default_rng(K)
    trees = [orc.decode_tree(K, int(rng.integers(0, K ** (K - 1)))) for _ in range(600)]
    trees += [[-1] + list(range(K - 1)), [-1] + [0] * (K - 1), [-1, 0] + [1] * (K - 2), [-1] + [i // 2 for i in range(K - 1)]]
    n_reg = 0
    for t in trees:
        reg, fg, tr = classify(lib, t, s, r)
        if reg:
            n_reg += 1
            n_pack = bin(int(rec[fg][4])).count("1")
            assert K - r <= n_pack <= s and 0 <= tr < nt
    assert n_reg > 0.8 * len(trees)


class LazyTables(object):
    """inside rows FG / FSG on demand (memoised recursion over the recipes): what a handful of trees needs out of
    the 10^5 rows of K >= 8."""

    def __init__(self, D, p0, rec):
        self.D, self.p0, self.rec = D, p0, rec
        self._fg, self._fsg = {}, {}

    def fg(self, i):
        if i not in self._fg:
            kind, u, a, b, mask = (int(x) for x in self.rec[i])
            if kind == 0:
                v = orc.log_cumsum(self.D[u]) + self.p0
            elif kind == 1:
                v = self.D[u] + self.fsg(b)
            else:
                v = orc.log_convolve(self.fg(a), self.fg(b))
            self._fg[i] = v
        return self._fg[i]

    def fsg(self, i):
        if i not in self._fsg:
            self._fsg[i] = orc.log_cumsum(self.fg(i)) + self.p0
        return self._fsg[i]

    class _View(object):
        def __init__(self, owner):
            self.owner = owner

        def __getitem__(self, i):
            return self.owner.fg(int(i))

    @property
    def FG(self):
        return LazyTables._View(self)


@pytest.mark.parametrize("K", [8, 9, 10])
def test_pairs_at_the_default_geometry_of_large_k_reproduce_the_oracle(built_library, K):
    """K = 8 (4, 5), K = 9 and 10 (5, 6) -- the cluster counts of BASELINE configs 3 and 4: the cut of a tree
    (host instance of classify_tree), the inside row of its pack and the context row of the rest, built here in
    fp64 from the recipes the kernels follow (only the rows these trees need), reproduce the oracle's score.
    Random trees + the extreme shapes; K = 10 also meets irregular trees (2 %), which are skipped here (their
    programs are checked in test_forest_cpu.py)."""
    lib = cuda_shim.load_library()
    s, r, nf, nt = pair_geometry(lib, K)
    assert (s, r) == {8: (4, 5), 9: (5, 6), 10: (5, 6)}[K]
    rec = recipes(lib, K, s, nf)
    crec = ctx_recipes(lib, K, s, r, nt)
    D = _tables(K, 31 * K)
    p0 = orc.prior_const(K)
    tab = LazyTables(D, p0, rec)
    ctxs = Contexts(D, p0, tab.FG, crec)
    rng = np.random.default_rng(900 + K)
    trees = [orc.decode_tree(K, int(rng.integers(0, K ** (K - 1)))) for _ in range(36)]
    trees += [[-1] + list(range(K - 1)), [-1] + [0] * (K - 1), [-1, 0] + [1] * (K - 2), [-1] + [i // 2 for i in range(K - 1)],
              [K - 1] + [-1] + [1] * (K - 3) + [K - 2]]
    n_reg = 0
    memo = {}
    for tree in trees:
        reg, fg, t = classify(lib, tree, s, r)
        if not reg:
            continue
        n_reg += 1
        assert 0 <= fg < nf and 0 <= t < nt
        want = orc.tree_score_samples(tree, D[:, None, :], p0, memo)[0]
        got = logsumexp(tab.FG[fg] + ctxs.row(t))
        assert got == pytest.approx(want, rel=1e-9), (K, s, r, tree, fg, t)
    assert n_reg >= len(trees) - (4 if K == 10 else 0)


def test_default_geometry_is_consistent(built_library):
    lib = cuda_shim.load_library()
    for K in range(1, 13):
        s, r, nf, nt = pair_geometry(lib, K)
        assert r == 0 or (s + r >= K and r - 1 <= s)


@pytest.mark.parametrize("K,s,r,ctx_used,ctx_full,irregular", [(5, 3, 3, 85, 315, 0), (5, 3, 2, 45, 45, 60), (6, 3, 4, 846, 4446, 0),
                                                               (6, 3, 3, 606, 606, 360), (7, 4, 4, 2821, 9996, 0),
                                                               (7, 4, 3, 1036, 1036, 13230), (8, 4, 5, 40552, 194552, 0),
                                                               (8, 4, 4, 19552, 19552, 166320)])
def test_host_cut_of_every_tree_reproduces_the_gpu_counts(built_library, K, s, r, ctx_used, ctx_full, irregular):
    """profiles/r02_pair_geometry.txt was counted on the GPU by cutting ALL K^(K-1) trees (k_build_pair_table); the
    host instance of the same code (decode_tree + classify_tree) over all trees gives the same number of irregular
    trees and of context rows in use -- host and device agree on the cut of every tree.  (K = 9 and K = 10 by
    profiles/host_cut_stats.py: profiles/r02_host_cut_stats.txt.)"""
    import importlib.util
    import os
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("host_cut_stats", os.path.join(here, "profiles", "host_cut_stats.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    got = mod.stats(K, s, r)
    assert (got["context_rows_used"], got["context_rows_full"], got["irregular"]) == (ctx_used, ctx_full, irregular)
    line = "K=%2d s=%d r=%d  trees %11d  inside rows %8d  context rows %9d of %9d  irregular %10d" % (
        K, s, r, K ** (K - 1), got["inside_rows"], ctx_used, ctx_full, irregular)
    with open(os.path.join(here, "profiles", "r02_pair_geometry.txt")) as f:
        assert any(l.startswith(line) for l in f), line
