"""Parity of the CUDA path (through the C ABI) against the oracle and the golden vectors.
Runs on the B200 box:  python -m pytest tests -m gpu"""
import math

import numpy as np
import pytest

import bamse_oracle as orc

pytestmark = pytest.mark.gpu

FP32_RTOL = 1e-6      # north_star: log marginal likelihoods within 1e-6 relative in fp32 mode
FP64_RTOL = 1e-10     # ... and 1e-10 in fp64 check mode


@pytest.fixture(scope="module")
def shim():
    from bamse_b200 import cuda_shim
    return cuda_shim


@pytest.fixture()
def ctx(shim):
    c = shim.Context(0)
    yield c
    c.close()


def pack_items(Kmax, items):
    """items: list of (clust, tree, nodes) -> arrays for score_list."""
    n = len(items)
    clust = np.zeros(n, dtype=np.int32)
    par = np.full((n, Kmax), -2, dtype=np.int8)
    nod = np.full((n, Kmax), -1, dtype=np.int8)
    for i, (c, tree, nodes) in enumerate(items):
        clust[i] = c
        par[i, :len(tree)] = tree
        nod[i, :len(nodes)] = nodes
    return clust, par, nod


def test_golden_tree_scores(ctx, shim, golden_scores):
    """every golden case (reference-produced) through bamse_score_list, both precisions."""
    for case in golden_scores:
        As = np.array(case["var"])[None]
        Bs = np.array(case["ref"])[None]
        K = As.shape[1]
        assert case["K_prior"] == K
        ctx.upload_problem([case["err"]], [K], As, Bs)
        clust, par, nod = pack_items(K, [(0, case["tree"], case["nodes"])])
        s64 = ctx.score_list(clust, par, nod, shim.FP64)[0, 0]
        s32 = ctx.score_list(clust, par, nod, shim.FP32)[0, 0]
        s32l = ctx.score_list(clust, par, nod, shim.FP32_LOGDOMAIN)[0, 0]
        assert abs(s64 - case["total"]) <= FP64_RTOL * abs(case["total"]), (case["id"], s64, case["total"])
        assert abs(s32 - case["total"]) <= FP32_RTOL * abs(case["total"]), (case["id"], s32, case["total"])
        assert abs(s32l - case["total"]) <= FP32_RTOL * abs(case["total"]), (case["id"], s32l, case["total"])


def _random_problem(rng, K, M, cov):
    frac = np.sort(rng.dirichlet(np.ones(K + 1))[:K])[::-1]
    cm = np.clip(np.cumsum(frac[::-1])[::-1][:, None] * rng.uniform(0.5, 1.0, size=(K, M)), 0, 1)
    As = rng.binomial(cov, cm * 0.497 + 0.003)
    return As, cov - As


@pytest.mark.parametrize("K,M,cov,err", [(1, 2, 100, 0.003), (2, 3, 500, 0.01), (3, 2, 50, 0.003),
                                         (4, 3, 3000, 0.001), (5, 2, 200, 0.003)])
def test_dense_range_matches_oracle(ctx, shim, K, M, cov, err):
    """all K^(K-1) labeled rooted trees of one clustering, by index, vs the numpy oracle."""
    rng = np.random.default_rng(100 + K)
    As, Bs = _random_problem(rng, K, M, cov)
    ctx.upload_problem([err], [K], As[None], Bs[None])
    T = K ** (K - 1)
    D = np.array([[orc.distrib(As[k, m], Bs[k, m], err) for m in range(M)] for k in range(K)])
    p0 = orc.prior_const(K)
    trees = [orc.decode_tree(K, i) for i in range(T)]
    want = np.array([orc.tree_score_samples(t, D, p0).sum() for t in trees])
    want_w = np.array([math.log(orc.canon_weight(t)[1]) for t in trees])
    s64, w64 = ctx.score_range(0, 0, 0, T, shim.FP64)
    s32, w32 = ctx.score_range(0, 0, 0, T, shim.FP32)
    np.testing.assert_allclose(s64, want, rtol=FP64_RTOL, atol=0)
    np.testing.assert_allclose(s32, want, rtol=FP32_RTOL, atol=0)
    np.testing.assert_allclose(w64, want_w, rtol=1e-15, atol=0)   # integer weight: exact
    np.testing.assert_array_equal(w64, w32)
    # library decoder == oracle decoder (bit-exact index work)
    for i in range(0, T, max(1, T // 50)):
        assert shim.decode_tree(K, i) == trees[i]


def test_topk_matches_oracle_ranking(ctx, shim):
    """two clusterings (K=3 and K=4) + two error rates in one launch: global top-k per
    parameter set, with thresholds, identical trees/clusterings to the oracle ranking."""
    rng = np.random.default_rng(5)
    M, Kmax = 2, 4
    Ks = [3, 4]
    errs = [0.003, 0.02]
    var = np.zeros((2, Kmax, M), dtype=np.int64)
    ref = np.zeros((2, Kmax, M), dtype=np.int64)
    for c, K in enumerate(Ks):
        a, b = _random_problem(rng, K, M, 400)
        var[c, :K], ref[c, :K] = a, b
    ctx.upload_problem(errs, Ks, var, ref)
    const = np.array([[1.5, -2.0], [0.25, 3.0]])
    full = {}
    for p, err in enumerate(errs):
        for c, K in enumerate(Ks):
            D = np.array([[orc.distrib(var[c, k, m], ref[c, k, m], err) for m in range(M)] for k in range(K)])
            for i in range(K ** (K - 1)):
                t = orc.decode_tree(K, i)
                full[(p, c, i)] = orc.tree_score_samples(t, D, orc.prior_const(K)).sum() + math.log(orc.canon_weight(t)[1])
    # thresholds: the median score of each (p, c) -- roughly half the trees are dropped
    thr = np.array([[np.median([v for (pp, cc, _), v in full.items() if pp == p and cc == c]) for c in range(2)]
                    for p in range(2)])
    topk = 7
    for precision in (shim.FP64, shim.FP32):
        rec, par, cnt = ctx.score_topk(const, thr, topk=topk, precision=precision)
        for p in range(2):
            cands = [(-(v + const[p, c]), c, i, v) for (pp, c, i), v in full.items() if pp == p and v >= thr[p, c]]
            cands.sort()
            want = cands[:topk]
            assert cnt[p] == len(want)
            for j, (neg_total, c, i, v) in enumerate(want):
                assert (rec[p, j]["clust"], rec[p, j]["tree"]) == (c, i), (precision, p, j)
                assert rec[p, j]["total"] == pytest.approx(-neg_total, rel=FP64_RTOL)
                assert rec[p, j]["score"] == pytest.approx(v, rel=FP64_RTOL)
                assert [int(x) for x in par[p, j, :Ks[c]]] == orc.decode_tree(Ks[c], i)
    assert ctx.last_eval_count == 2 * (9 + 64)


def test_tree_range_sharding_equals_single_range(ctx, shim):
    """shard-merge equivalence on one GPU: scoring index ranges
    separately and merging gives the single-range result."""
    rng = np.random.default_rng(9)
    K, M = 5, 2
    As, Bs = _random_problem(rng, K, M, 300)
    ctx.upload_problem([0.003], [K], As[None], Bs[None])
    T = K ** (K - 1)
    const = np.zeros((1, 1))
    whole, _, _ = ctx.score_topk(const, topk=5, precision=shim.FP64)
    parts = []
    for r in range(4):
        b, e = T * r // 4, T * (r + 1) // 4
        rec, _, cnt = ctx.score_topk(const, None, [b], [e], topk=5, precision=shim.FP64)
        parts.extend(rec[0, :cnt[0]].tolist())
    parts.sort(key=lambda r: (-r[0], r[2], r[4]))
    assert [r[4] for r in parts[:5]] == whole[0]["tree"].tolist()
    assert [r[0] for r in parts[:5]] == whole[0]["total"].tolist()
    # interleaved shards (bamse_set_shard), fp32 selection + fp64 re-score, host merge
    from bamse_b200 import multigpu
    lists = []
    evals = 0
    for r in range(3):
        ctx.set_shard(r, 3)
        rec, _, _ = ctx.score_topk(const, topk=5, precision=shim.FP32)
        evals += ctx.last_eval_count
        lists.append(rec)
    ctx.set_shard(0, 1)
    assert evals == T
    merged = multigpu.merge_records(np.stack(lists), 5)
    assert merged[0]["tree"].tolist() == whole[0]["tree"].tolist()
    np.testing.assert_allclose(merged[0]["total"], whole[0]["total"], rtol=1e-12)


def test_errors_are_reported_not_thrown(ctx, shim):
    with pytest.raises(shim.BamseCudaError) as e:
        ctx.score_range(0, 0, 0, 1)
    assert e.value.code == -4  # BAMSE_ESTATE: nothing uploaded
    with pytest.raises(shim.BamseCudaError):
        ctx.upload_problem([0.0], [2], np.ones((1, 2, 1)), np.ones((1, 2, 1)))   # err must be > 0
    ctx.upload_problem([0.01], [2], np.ones((1, 2, 1)), np.ones((1, 2, 1)))
    clust, par, nod = pack_items(2, [(0, [0, 1], [0, 1])])                       # a 2-cycle, not a tree
    with pytest.raises(shim.BamseCudaError) as e:
        ctx.score_list(clust, par, nod)
    assert e.value.code == -1


def test_index_ranges_beyond_32_bits(ctx, shim):
    """K = 12 has 12^11 = 7.4e11 labeled rooted trees: a range starting above 2^32 (and one of a
    K = 9 clustering in the same problem) through the top-k path equals the dense check-mode
    scores of the same indices."""
    rng = np.random.default_rng(2026)
    M, Kmax = 2, 12
    Ks = [12, 9]
    var = np.zeros((2, Kmax, M), dtype=np.int64)
    ref = np.zeros((2, Kmax, M), dtype=np.int64)
    for c, K in enumerate(Ks):
        a, b = _random_problem(rng, K, M, 600)
        var[c, :K], ref[c, :K] = a, b
    ctx.upload_problem([0.004], Ks, var, ref)
    begin = np.array([(1 << 33) + 12345, 9 ** 8 - 700], dtype=np.uint64)
    n = [900, 700]
    end = begin + np.array(n, dtype=np.uint64)
    const = np.array([[0.0, 25.0]])
    rec, par, cnt = ctx.score_topk(const, None, begin, end, topk=10, precision=shim.FP32)
    assert cnt[0] == 10 and ctx.last_eval_count == sum(n)
    full = []
    for c in range(2):
        s, w = ctx.score_range(0, c, int(begin[c]), n[c], shim.FP64)
        full += [(-(s[i] + w[i] + const[0, c]), c, int(begin[c]) + i) for i in range(n[c])]
    full.sort()
    for j in range(10):
        assert (int(rec[0, j]["clust"]), int(rec[0, j]["tree"])) == full[j][1:], j
        assert rec[0, j]["total"] == pytest.approx(-full[j][0], rel=1e-10)
        K = Ks[full[j][1]]
        assert [int(x) for x in par[0, j, :K]] == orc.decode_tree(K, full[j][2])
