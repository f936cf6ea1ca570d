"""CPU tests of the data-independent machinery of the sub-forest memoisation (bamse_b200/csrc/forest.cuh):
the table recipes and the table-driven evaluation programs come out of libbamse_cuda's HOST instances of the
same code the kernels run (bamse_debug_* entry points, no GPU needed) and are interpreted here with the numpy
oracle's operators.  Every program must reproduce the oracle's tree score; the live-vector bound that selects
the 2-slot kernel must hold."""
import ctypes
import math

import numpy as np
import pytest
from scipy.special import logsumexp

import bamse_oracle as orc
from bamse_b200 import cuda_shim

FOP_PUSH_G, FOP_PUSH_SG, FOP_SCAN, FOP_CONV, FOP_ADDD = range(5)
FIN_LSE, FIN_DOT_D, FIN_DOT_SD, FIN_DOT_FH = range(4)
LOGL = math.log(512)


def geometry(lib, K, s=0, sh=0):
    a, b, nf, nfh, slots = (ctypes.c_int32(s), ctypes.c_int32(sh), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32())
    assert lib.bamse_debug_forest_geometry(K, ctypes.byref(a), ctypes.byref(b), ctypes.byref(nf), ctypes.byref(nfh),
                                           ctypes.byref(slots)) == 0
    return a.value, b.value, nf.value, nfh.value, slots.value


def recipes(lib, K, s, nf):
    out = np.zeros((nf, 5), dtype=np.int32)
    assert lib.bamse_debug_forest_recipes(K, s, out.ctypes.data_as(ctypes.c_void_p)) == 0
    return out


def program(lib, tree, nodes, Kc, s, sh):
    par = np.asarray(tree, dtype=np.int8)
    nod = None if nodes is None else np.asarray(nodes, dtype=np.int8)
    out = np.zeros(46, dtype=np.int32)
    rc = lib.bamse_debug_program(len(tree), par.ctypes.data_as(ctypes.c_void_p),
                                 None if nod is None else nod.ctypes.data_as(ctypes.c_void_p), Kc, s, sh,
                                 out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, (tree, nodes)
    n = int(out[0])
    return dict(ops=[(int(out[6 + i]), int(out[26 + i])) for i in range(n)], fin=int(out[1]), fin_arg=int(out[2]),
                fin_r=int(out[3]), scre=int(out[4]), max_sp=int(out[5]))


class _Rows(object):
    def __init__(self, get):
        self.get = get

    def __getitem__(self, i):
        return self.get(int(i))


class Tables(object):
    """numpy build of FG / FSG / FH / SD for one sample, following the recipes; rows are built when first read
    (a row's operands have fewer nodes, so the recursion ends), because a sample of trees touches a small part
    of the 10^4 .. 10^5 rows of K >= 7."""

    def __init__(self, D, p0, rec, nfh):
        self.D, self.p0, self.rec = D, p0, rec
        K = D.shape[0]
        self._fg, self._fsg = {}, {}
        self.FG, self.FSG = _Rows(self._get_fg), _Rows(self._get_fsg)
        self.masks = [int(r[4]) for r in rec]
        self.nfh = nfh
        self._fh = {}
        # SD[r][i] = e^p0 / L^2 * sum_{x >= i} d_r[x]
        self.SD = [np.logaddexp.accumulate(D[r][::-1])[::-1] + p0 - 2 * LOGL for r in range(K)]

    def _get_fg(self, i):
        if i not in self._fg:
            kind, u, a, b, mask = (int(x) for x in self.rec[i])
            if kind == 0:
                v = orc.log_cumsum(self.D[u]) + self.p0
            elif kind == 1:
                v = self.D[u] + self._get_fsg(b)
            else:
                v = orc.log_convolve(self._get_fg(a), self._get_fg(b))
            self._fg[i] = v
        return self._fg[i]

    def _get_fsg(self, i):
        if i not in self._fsg:
            self._fsg[i] = orc.log_cumsum(self._get_fg(i)) + self.p0
        return self._fsg[i]

    def FH(self, r, f):
        assert f < self.nfh and not (self.masks[f] >> r) & 1
        if (r, f) not in self._fh:
            d, sg = self.D[r], self.FSG[f]
            L = len(d)
            self._fh[(r, f)] = np.array([logsumexp(d[i:] + sg[:L - i]) for i in range(L)]) - 2 * LOGL
        return self._fh[(r, f)]

    def run(self, prog):
        st = []
        for op, arg in prog["ops"]:
            if op == FOP_PUSH_G:
                st.append(self.FG[arg])
            elif op == FOP_PUSH_SG:
                st.append(self.FSG[arg])
            elif op == FOP_SCAN:
                st[-1] = orc.log_cumsum(st[-1]) + self.p0
            elif op == FOP_CONV:
                b = st.pop()
                st[-1] = orc.log_convolve(st[-1], b)
            else:
                st[-1] = st[-1] + self.D[arg]
            assert len(st) <= prog["max_sp"]
        assert len(st) == 1
        top, r = st[0], prog["fin_r"]
        if prog["fin"] == FIN_LSE:
            return logsumexp(top) - LOGL
        if prog["fin"] == FIN_DOT_D:
            return logsumexp(top + self.D[r]) - LOGL
        if prog["fin"] == FIN_DOT_SD:
            return logsumexp(top + self.SD[r])
        return logsumexp(top + self.FH(r, prog["fin_arg"]))


def _tables(K, seed):
    rng = np.random.default_rng(seed)
    As = rng.integers(3, 400, size=K)
    Bs = rng.integers(3, 400, size=K)
    return np.array([orc.distrib(As[k], Bs[k], 0.003) for k in range(K)])


@pytest.mark.parametrize("K,s,sh", [(3, 1, 1), (4, 2, 2), (5, 3, 2), (5, 2, 1), (6, 4, 2), (6, 3, 3), (6, 2, 2), (7, 4, 3)])
def test_programs_reproduce_the_oracle(built_library, K, s, sh):
    """all trees for K <= 5, a stratified sample above: table-driven program == oracle score (1e-9 relative),
    canonizeF weight carried through, live vectors within the kernel's slot bound."""
    lib = cuda_shim.load_library()
    s, sh, nf, nfh, slots = geometry(lib, K, s, sh)
    rec = recipes(lib, K, s, nf)
    D = _tables(K, 10 * K + s)
    p0 = orc.prior_const(K)
    tab = Tables(D, p0, rec, nfh)
    T = K ** (K - 1)
    idx = range(T) if T <= 700 else list(range(0, T, max(1, T // 220))) + [T - 1]
    convs = 0
    memo = {}                        # the oracle's sub-tree messages for THIS D (pure-function cache, oracle/bamse_oracle.py)
    for i in idx:
        tree = orc.decode_tree(K, i)
        prog = program(lib, tree, None, K, s, sh)
        want = orc.tree_score_samples(tree, D[:, None, :], p0, memo)[0]
        got = tab.run(prog)
        assert got == pytest.approx(want, rel=1e-9), (K, s, sh, i, tree, prog)
        assert prog["scre"] == orc.canon_weight(tree)[1]
        assert prog["max_sp"] <= slots
        convs += sum(1 for op, _ in prog["ops"] if op == FOP_CONV)
    if (K, s) in ((5, 3), (6, 4)):
        assert convs == 0            # these geometries leave no convolution to the enumeration


def test_sub_forest_items_and_permuted_nodes(built_library):
    """bamse_score_list items: trees on node SUBSETS of a K = 7 clustering with arbitrary root position and
    node order (what reorder / get_candidate_tree submit) go through the same builder."""
    lib = cuda_shim.load_library()
    K = 7
    s, sh, nf, nfh, slots = geometry(lib, K)
    rec = recipes(lib, K, s, nf)
    D = _tables(K, 99)
    p0 = orc.prior_const(K)
    tab = Tables(D, p0, rec, nfh)
    rng = np.random.default_rng(4)
    for k in (1, 2, 2, 3, 4, 5, 6, 7, 7):
        for _ in range(6):
            nodes = [int(x) for x in rng.permutation(K)[:k]]
            perm = rng.permutation(k)
            tree = [-2] * k
            tree[perm[0]] = -1
            for i in range(1, k):
                tree[perm[i]] = int(perm[rng.integers(0, i)])
            prog = program(lib, tree, nodes, K, s, sh)
            want = orc.tree_score_samples(tree, D[nodes][:, None, :], p0)[0]
            assert tab.run(prog) == pytest.approx(want, rel=1e-9), (tree, nodes, prog)


@pytest.mark.parametrize("K", [8, 9, 10, 11, 12])
def test_program_bounds_for_large_k(built_library, K):
    """random trees + extreme shapes at K = 8 ... 12 with the default geometry: programs fit 20 ops, row ids fit
    16 bits, live vectors stay within the slot bound, and the star / chain / caterpillar evaluate correctly."""
    lib = cuda_shim.load_library()
    s, sh, nf, nfh, slots = geometry(lib, K)
    assert nf <= 65535
    rng = np.random.default_rng(K)
    trees = [orc.decode_tree(K, int(rng.integers(0, K ** (K - 1)))) for _ in range(400)]
    trees += [[-1] + list(range(K - 1)), [-1] + [0] * (K - 1), [-1, 0] + [1] * (K - 2),
              [-1] + [i // 2 for i in range(K - 1)], [-1] + [max(0, i - 1) if i % 2 else max(0, i - 2) for i in range(K - 1)]]
    worst_sp = worst_ops = 0
    for t in trees:
        prog = program(lib, t, None, K, s, sh)
        worst_sp, worst_ops = max(worst_sp, prog["max_sp"]), max(worst_ops, len(prog["ops"]))
        assert all(0 <= a < nf for op, a in prog["ops"] if op in (FOP_PUSH_G, FOP_PUSH_SG))
        assert prog["fin"] != FIN_DOT_FH or prog["fin_arg"] < nfh
    assert worst_sp <= slots and worst_ops <= 20
    if K <= 9:
        rec = recipes(lib, K, s, nf)
        D = _tables(K, K)
        p0 = orc.prior_const(K)
        tab = Tables(D, p0, rec, nfh)
        memo = {}
        for t in trees[-5:] + trees[:3]:
            want = orc.tree_score_samples(t, D[:, None, :], p0, memo)[0]
            assert tab.run(program(lib, t, None, K, s, sh)) == pytest.approx(want, rel=1e-9)


def _step_down(s, sh):
    """reduce_forest_geometry (bamse_b200/csrc/forest_host.h): the next smaller geometry, or None."""
    if sh > 1 and sh >= s - 1:
        return s, sh - 1
    if s > 2:
        return s - 1, min(sh, s - 1)
    if sh > 1:
        return s, sh - 1
    return None


@pytest.mark.parametrize("K", list(range(3, 13)))
def test_stepped_down_geometries_stay_within_the_kernel_limits(built_library, K):
    """Under memory pressure (BAMSE_TABLE_GB, or a problem larger than the device) plan_tables walks the geometry
    of a cluster count down to (s, sh) = (2, 1).  On every rung the programs must still fit the kernel: at most
    forest_slots_needed live vectors (the kernel traps above its slot count), at most 20 operations, 16-bit row
    ids.  All trees for K <= 6, 20 000 random ones + the extreme shapes above (exhaustive runs of this check up to
    K = 8, 1.5 M random trees per rung for K = 9 ... 12, gave the same maxima: 2 vectors up to (K <= 8, any s),
    (K <= 10, s >= 3); 3 vectors at K >= 9 with s = 2 and at K >= 11 -- a fourth would need a 19-node tree)."""
    lib = cuda_shim.load_library()
    T = K ** (K - 1)
    rng = np.random.default_rng(100 + K)
    idx = range(T) if T <= 8000 else [int(x) for x in rng.integers(0, T, size=20000)]
    trees = [orc.decode_tree(K, i) for i in idx]
    trees += [[-1] + list(range(K - 1)), [-1] + [0] * (K - 1), [-1, 0] + [1] * (K - 2),
              [-1] + [i // 2 for i in range(K - 1)], [-1] + [max(0, i - 1) if i % 2 else max(0, i - 2) for i in range(K - 1)]]
    if K >= 9:      # the smallest tree that keeps three vectors alive at s = 2: two 4-node "need two" subtrees under a root
        trees.append(([-1, 0, 1, 1, 2, 0, 5, 5, 6] + [0] * (K - 9))[:K])
    s, sh, nf, nfh, slots = geometry(lib, K)
    rungs = [(s, sh)]
    while _step_down(*rungs[-1]):
        rungs.append(_step_down(*rungs[-1]))
    assert rungs[-1] == (2, 1) or K <= 3
    seen_sp = {}
    for s, sh in rungs:
        s2, sh2, nf, nfh, slots = geometry(lib, K, s, sh)
        assert (s2, sh2) == (s, sh) and nf <= 65535
        worst_sp = worst_ops = 0
        for t in trees:
            prog = program(lib, t, None, K, s, sh)
            worst_sp, worst_ops = max(worst_sp, prog["max_sp"]), max(worst_ops, len(prog["ops"]))
            assert all(0 <= a < nf for op, a in prog["ops"] if op in (FOP_PUSH_G, FOP_PUSH_SG))
            assert prog["fin"] != FIN_DOT_FH or prog["fin_arg"] < nfh
        assert worst_sp <= slots <= 3 and worst_ops <= 20, (K, s, sh, worst_sp, slots, worst_ops)
        seen_sp[(s, sh)] = worst_sp
    if K >= 9:
        assert seen_sp[(2, 1)] == 3


@pytest.mark.parametrize("K,s,sh", [(7, 2, 1), (8, 2, 1), (8, 3, 1), (9, 2, 1)])
def test_programs_of_the_smallest_geometries_reproduce_the_oracle(built_library, K, s, sh):
    """the rungs only memory pressure reaches: a stratified sample of trees, program == oracle (1e-9 relative)."""
    lib = cuda_shim.load_library()
    s, sh, nf, nfh, slots = geometry(lib, K, s, sh)
    rec = recipes(lib, K, s, nf)
    D = _tables(K, 7 * K + s)
    p0 = orc.prior_const(K)
    tab = Tables(D, p0, rec, nfh)
    T = K ** (K - 1)
    trees = [orc.decode_tree(K, i) for i in list(range(0, T, max(1, T // 40))) + [T - 1]]
    trees += [[-1] + [i // 2 for i in range(K - 1)], [-1, 0] + [1] * (K - 2)]
    if K >= 9:
        trees.append([-1, 0, 1, 1, 2, 0, 5, 5, 6][:K])
    memo = {}
    for t in trees:
        prog = program(lib, t, None, K, s, sh)
        want = orc.tree_score_samples(t, D[:, None, :], p0, memo)[0]
        assert tab.run(prog) == pytest.approx(want, rel=1e-9), (K, s, sh, t, prog)
        assert prog["scre"] == orc.canon_weight(t)[1] and prog["max_sp"] <= slots


@pytest.mark.parametrize("K", [2, 3, 4, 5, 6, 7])
def test_host_canon_weight_of_every_tree(built_library, K):
    """the canonizeF weight the device code carries (`scre` of the program built by the host instance of
    build_program_forest) equals the oracle's canon_weight for EVERY labeled rooted tree of K <= 7 clusters
    (117 649 at K = 7: the cluster counts of BASELINE configs 2 and 3) -- integer parity, no sampling."""
    lib = cuda_shim.load_library()
    s, sh, nf, nfh, slots = geometry(lib, K)
    out = np.zeros(46, dtype=np.int32)
    po = out.ctypes.data_as(ctypes.c_void_p)
    hist = {}
    for i in range(K ** (K - 1)):
        tree = orc.decode_tree(K, i)
        par = np.asarray(tree, dtype=np.int8)
        assert lib.bamse_debug_program(K, par.ctypes.data_as(ctypes.c_void_p), None, K, s, sh, po) == 0
        w = orc.canon_weight(tree)[1]
        assert int(out[4]) == w, (K, i, tree)
        hist[w] = hist.get(w, 0) + 1
    assert sum(hist.values()) == K ** (K - 1) and hist.get(1, 0) > 0
