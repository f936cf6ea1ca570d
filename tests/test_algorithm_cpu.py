"""The mathematical facts the fp32 production kernel (bamse_b200/csrc/score_fast.cuh) relies on,
checked on the CPU against the oracle's exact operators (oracle/bamse_oracle.py, which restates
clustering.py:359-395 of the reference).  No CUDA code runs here; the numpy restatements below follow
the kernel's algorithm (merge path, window, root shortcut, pair ownership of outputs) step by step."""
import math

import numpy as np
import pytest

import bamse_oracle as orc

L = orc.GRID
THETA_BITS = 26.0                      # score_fast.cuh: THETA


def _tables(K, M, cov, seed, err=0.003):
    rng = np.random.default_rng(seed)
    frac = np.sort(rng.dirichlet(np.ones(K + 1))[:K])[::-1]
    cm = np.clip(np.cumsum(frac[::-1])[::-1][:, None] * rng.uniform(0.6, 1.0, size=(K, M)), 0, 1)
    a = rng.binomial(cov, cm * (0.5 - err) + err)
    b = cov - a
    return np.array([[orc.distrib(a[k, m], b[k, m], err) for m in range(M)] for k in range(K)])


def _messages(K, cov, seed, n_trees=6):
    """Pairs (a, b) of vectors that the evaluation of random trees actually convolves."""
    D = _tables(K, 1, cov, seed)[:, 0, :]
    p0 = orc.prior_const(K)
    pairs, msgs = [], []
    real = orc.log_convolve

    def spy(a, b):
        if not (np.allclose(a, a[0]) or np.allclose(b, b[0])):
            pairs.append((a.copy(), b.copy()))
        out = real(a, b)
        msgs.append(out)
        return out

    orc.log_convolve = spy
    try:
        rng = np.random.default_rng(seed + 1)
        for _ in range(n_trees):
            tree = list(orc.decode_tree(K, int(rng.integers(0, K ** (K - 1)))))
            orc.eval_node(tree, tree.index(-1), D, p0)
    finally:
        orc.log_convolve = real
    return D, p0, pairs, msgs


def _merge_path(a, b):
    """jstar[x] of warp_merge_path: per 16-output run a binary search for the split, then a walk that
    always advances along the larger slope (ties go to a)."""
    ext = lambda v: np.concatenate([v, np.full(4, -np.inf)])          # the 4 trailing -inf entries of a slot
    a, b = ext(a), ext(b)
    jstar = np.empty(L, dtype=np.int64)
    with np.errstate(invalid="ignore"):
        for lane in range(32):
            x0 = 16 * lane
            lo, hi = 0, x0
            while lo < hi:
                mid = (lo + hi) >> 1
                if a[mid + 1] - a[mid] >= b[x0 - mid] - b[x0 - mid - 1]:
                    lo = mid + 1
                else:
                    hi = mid
            i, j = lo, x0 - lo
            for t in range(16):
                jstar[x0 + t] = j
                if (b[j + 1] - b[j]) > (a[i + 1] - a[i]):
                    j += 1
                else:
                    i += 1
    return jstar


def _terms(a, b):
    xs = np.arange(L)
    idx = xs[:, None] - xs[None, :]
    return np.where(idx >= 0, b[None, :] + a[np.where(idx >= 0, idx, 0)], -np.inf)


@pytest.mark.parametrize("cov", [30, 800, 20000])
def test_every_message_is_log_concave(cov):
    """binomial tables are concave in the grid index; scans and convolutions keep that."""
    D, _, _, msgs = _messages(6, cov, seed=cov)
    for v in list(D) + msgs:
        d2 = np.diff(v, 2)
        assert d2.max() <= 1e-9 * max(1.0, np.abs(v).max())


@pytest.mark.parametrize("cov", [30, 800, 20000])
def test_merge_path_finds_the_maximising_split(cov):
    _, _, pairs, _ = _messages(6, cov, seed=7 + cov)
    assert pairs
    for a, b in pairs[:10]:
        T = _terms(a, b)
        js = _merge_path(a, b)
        xs = np.arange(L)
        assert (js <= xs).all() and (js >= 0).all()
        assert np.all(np.diff(js) >= 0) and np.all(np.diff(js) <= 1)          # the walk advances by 0 or 1
        got, want = T[xs, js], T.max(axis=1)
        # equal up to the rounding noise of the slopes (the kernel only needs M_x within a factor of 2^0.5:
        # its sanity check accepts sums in [0.7, 2048])
        assert np.all(got >= want - 1e-6 * np.maximum(1.0, np.abs(want)))


@pytest.mark.parametrize("cov", [30, 800, 20000])
def test_window_is_contiguous_and_carries_the_sum(cov):
    """terms within 2^-THETA of an output's maximum form one window around j*; the rest is < 4e-7 relative."""
    _, _, pairs, _ = _messages(6, cov, seed=11 + cov)
    cut = THETA_BITS * math.log(2.0)
    for a, b in pairs[:10]:
        T = _terms(a, b)
        mx = T.max(axis=1)
        inside = T >= (mx[:, None] - cut)
        first = inside.argmax(axis=1)
        last = L - 1 - inside[:, ::-1].argmax(axis=1)
        assert np.all(inside.sum(axis=1) == last - first + 1)                   # no holes
        full = np.exp(T - mx[:, None]).sum(axis=1)
        win = np.where(inside, np.exp(T - mx[:, None]), 0.0).sum(axis=1)
        assert np.all((full - win) <= 4e-7 * full)


def test_scan_commutes_with_convolution():
    """scan(E1) (*) E2 == scan(E1 (*) E2): any child may carry the prior's scan, and a leaf child brings it
    from the SEL2 table (build_program_fast)."""
    _, p0, pairs, _ = _messages(5, 500, seed=3)
    a, b = pairs[0]
    scan = lambda v: orc.log_convolve(np.full(L, p0), v)
    lhs = orc.log_convolve(scan(a), b)
    rhs = scan(orc.log_convolve(a, b))
    np.testing.assert_allclose(lhs, rhs, rtol=1e-12, atol=1e-9)


def test_root_shortcut_is_a_dot_product():
    """The last operation of a tree, logsumexp_x(D[r][x] + scan(G)[x]), is a dot product of G with the
    suffix sums of D[r] (the SD2 table): no scan of the final 512-vector is needed."""
    D, p0, _, msgs = _messages(5, 500, seed=5)
    G = msgs[-1]
    d = D[0]
    log_l = math.log(L)
    scanned = np.logaddexp.accumulate(G) + p0 - log_l
    want = np.logaddexp.reduce(d + scanned) - log_l
    suffix = np.logaddexp.accumulate(d[::-1])[::-1]                             # log sum_{x >= i} exp(d[x])
    got = np.logaddexp.reduce(G + suffix) + p0 - log_l - log_l
    assert got == pytest.approx(want, rel=1e-12)


def test_pair_ownership_reuses_every_loaded_element():
    """conv_pass_pair: a lane owns outputs x (even) and x + 1; for the term j = j0 + k it needs
    (a[p - k], a[p - k + 1]), p = x - j0 -- the aligned pair q_{k/2} for even k and (q_{(k+1)/2}.hi,
    q_{(k-1)/2}.lo) for odd k, where q_i = (a[p - 2i], a[p - 2i + 1])."""
    rng = np.random.default_rng(0)
    a = rng.normal(size=64)
    x, j0 = 40, 8
    p = x - j0
    q = [(a[p - 2 * i], a[p - 2 * i + 1]) for i in range(5)]
    for k in range(8):
        want = (a[x - (j0 + k)], a[x + 1 - (j0 + k)])
        got = q[k // 2] if k % 2 == 0 else (q[(k + 1) // 2][1], q[(k - 1) // 2][0])
        assert got == want


def test_per_sample_scores_are_not_positive():
    """the early exit over samples (BAMSE_OPT_EARLY_EXIT) needs every per-sample score <= 0."""
    K, M = 5, 4
    for cov in (0, 5, 300, 40000):
        D = _tables(K, M, cov, seed=cov + 1)
        rng = np.random.default_rng(cov)
        for _ in range(4):
            tree = list(orc.decode_tree(K, int(rng.integers(0, K ** (K - 1)))))
            s = orc.tree_score_samples(tree, D, orc.prior_const(K))
            assert np.all(s <= 1e-12)


def test_window_statistics_of_a_32_output_pass():
    """What a half-warp pass exponentiates: the union of its 32 outputs' windows (one common j-range), against
    the outputs' own windows and against windows aligned per lane to the lane's own j* (DESIGN.md section 8,
    item 1).  own <= aligned <= union; on peaked data the union is ~1.3x the own windows."""
    _, _, pairs, _ = _messages(6, 16000, seed=21, n_trees=8)
    cut = THETA_BITS * math.log(2.0)
    own, union, aligned = [], [], []
    for a, b in pairs[:12]:
        T = _terms(a, b)
        mx, js = T.max(axis=1), T.argmax(axis=1)
        inside = T >= (mx[:, None] - cut)
        lo = inside.argmax(axis=1)
        hi = L - 1 - inside[:, ::-1].argmax(axis=1)
        own.append((hi - lo + 1).mean())
        for x0 in range(0, L, 32):
            s = slice(x0, x0 + 32)
            union.append(hi[s].max() - lo[s].min() + 1)
            aligned.append((js[s] - lo[s]).max() + (hi[s] - js[s]).max() + 1)
    own, union, aligned = np.mean(own), np.mean(union), np.mean(aligned)
    assert own <= aligned <= union
    assert 1.1 < union / own < 2.0
