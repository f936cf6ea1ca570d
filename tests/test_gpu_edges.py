"""Edge cases of the scoring path through the C ABI (the reference has no tests of its own, so
the cases are the ones the domain offers): one sample and BAMSE_MMAX samples, clusters without variant reads /
without any reads / without reference reads, error rates at both ends of (0, 0.5), more ranks requested than
trees exist, empty index ranges, thresholds nothing reaches, empty item lists, duplicated clusterings (exact
ties across clusterings), a context re-used for problems of other shapes.  Small cases against the numpy oracle,
larger ones fp32 against the fp64 check mode.  Runs on the B200 box:  python -m pytest tests -m gpu"""
import math

import numpy as np
import pytest

import bamse_oracle as orc

pytestmark = pytest.mark.gpu

FP32_RTOL = 1e-6      # north_star: 1e-6 relative in fp32 mode (absolute 1e-6 below |score| = 1)
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


def oracle_all_trees(var, ref, err):
    """(get_tree_score, log scre) of all K^(K-1) trees of one clustering, numpy oracle."""
    K, M = var.shape
    D = np.array([[orc.distrib(var[k, m], ref[k, m], err) for m in range(M)] for k in range(K)])
    p0 = orc.prior_const(K)
    trees = [orc.decode_tree(K, i) for i in range(K ** (K - 1))]
    s = np.array([orc.tree_score_samples(t, D, p0).sum() for t in trees])
    w = np.array([math.log(orc.canon_weight(t)[1]) for t in trees])
    return s, w


def close32(got, want):
    return np.all(np.abs(got - want) <= FP32_RTOL * np.maximum(np.abs(want), 1.0))


def close64(got, want):
    return np.all(np.abs(got - want) <= FP64_RTOL * np.maximum(np.abs(want), 1.0))


def check_dense_against_oracle(ctx, shim, p, c, want, want_w):
    n = len(want)
    s64, w64 = ctx.score_range(p, c, 0, n, shim.FP64)
    s32, w32 = ctx.score_range(p, c, 0, n, shim.FP32)
    s32l, _ = ctx.score_range(p, c, 0, n, shim.FP32_LOGDOMAIN)
    assert close64(s64, want), np.max(np.abs(s64 - want))
    assert close32(s32, want), np.max(np.abs(s32 - want))
    assert close32(s32l, want), np.max(np.abs(s32l - want))
    np.testing.assert_allclose(w64, want_w, rtol=1e-14, atol=0)
    np.testing.assert_allclose(w32, want_w, rtol=1e-14, atol=0)


def ranking(scores_by_clust, const, topk):
    """[(clust, tree)] of the best `topk` by (total desc, clust asc, tree asc)."""
    full = []
    for c, (s, w) in enumerate(scores_by_clust):
        full += [(-(s[i] + w[i] + const[c]), c, i) for i in range(len(s))]
    full.sort()
    return [(c, i) for _, c, i in full[:topk]]


def test_one_sample_and_the_maximum_number_of_samples(ctx, shim):
    """M = 1 and M = BAMSE_MMAX = 64 (tail splitting over samples, sample-major pair path, the per-sample result
    buffer): all trees of a K = 3 clustering against the oracle; K = 5 fp32 against the check mode."""
    rng = np.random.default_rng(64)
    for M in (1, 64):
        var = rng.integers(0, 200, size=(3, M)).astype(np.int64)
        ref = rng.integers(100, 400, size=(3, M)).astype(np.int64)
        ctx.upload_problem([0.003], [3], var[None], ref[None])
        want, want_w = oracle_all_trees(var, ref, 0.003)
        check_dense_against_oracle(ctx, shim, 0, 0, want, want_w)
        rec, par, cnt = ctx.score_topk(np.zeros((1, 1)), topk=4, precision=shim.FP32)
        assert cnt[0] == 4 and ctx.last_topk_verified
        assert [(int(r["clust"]), int(r["tree"])) for r in rec[0]] == ranking([(want, want_w)], [0.0], 4)
    M = 64
    var = rng.integers(0, 300, size=(5, M)).astype(np.int64)
    ref = rng.integers(200, 600, size=(5, M)).astype(np.int64)
    ctx.upload_problem([0.003], [5], var[None], ref[None])
    s64, w64 = ctx.score_range(0, 0, 0, 625, shim.FP64)
    s32, _ = ctx.score_range(0, 0, 0, 625, shim.FP32)
    assert close32(s32, s64)
    for early in (False, True):
        ctx.set_early_exit(early)
        rec, _, cnt = ctx.score_topk(np.zeros((1, 1)), topk=6, precision=shim.FP32)
        assert cnt[0] == 6 and ctx.last_topk_verified
        assert [(int(r["clust"]), int(r["tree"])) for r in rec[0]] == ranking([(s64, w64)], [0.0], 6)
    ctx.set_early_exit(False)
    with pytest.raises(shim.BamseCudaError) as e:      # one sample too many
        ctx.upload_problem([0.003], [2], np.ones((1, 2, 65), dtype=np.int64), np.ones((1, 2, 65), dtype=np.int64))
    assert e.value.code == -1


def test_clusters_without_variant_reads_without_reads_and_without_reference_reads(ctx, shim):
    """zero counts: a cluster with no variant read in any sample, one with no read at all in one sample, one with
    variant reads only in one sample (ref = 0).  All 64 trees in every precision against the oracle."""
    var = np.array([[120, 80, 95], [0, 0, 0], [30, 0, 12], [7, 55, 3]], dtype=np.int64)
    ref = np.array([[380, 420, 405], [250, 310, 275], [170, 0, 188], [193, 0, 197]], dtype=np.int64)
    ctx.upload_problem([0.002], [4], var[None], ref[None])
    want, want_w = oracle_all_trees(var, ref, 0.002)
    assert np.all(np.isfinite(want))
    check_dense_against_oracle(ctx, shim, 0, 0, want, want_w)
    for prec in (shim.FP32, shim.FP64):
        rec, par, cnt = ctx.score_topk(np.zeros((1, 1)), topk=5, precision=prec)
        assert [(int(r["clust"]), int(r["tree"])) for r in rec[0]] == ranking([(want, want_w)], [0.0], 5)
        assert close64(rec[0]["score"], np.array([want[t] + want_w[t] for t in rec[0]["tree"]]))
        for j in range(5):
            assert [int(x) for x in par[0, j]] == orc.decode_tree(4, int(rec[0, j]["tree"]))


def test_error_rates_at_both_ends_in_one_sweep(ctx, shim):
    """e = 1e-6 (log cf spans 13 nats: the widest dynamic range of the tables) and e = 0.499 (the grid is almost
    constant: the flattest tables) as two parameter sets of one upload."""
    rng = np.random.default_rng(499)
    var = rng.integers(0, 400, size=(4, 2)).astype(np.int64)
    ref = rng.integers(300, 900, size=(4, 2)).astype(np.int64)
    errs = [1e-6, 0.499]
    ctx.upload_problem(errs, [4], var[None], ref[None])
    wants = []
    for p, e in enumerate(errs):
        want, want_w = oracle_all_trees(var, ref, e)
        check_dense_against_oracle(ctx, shim, p, 0, want, want_w)
        wants.append((want, want_w))
    rec, _, cnt = ctx.score_topk(np.zeros((2, 1)), topk=3, precision=shim.FP32)
    for p in range(2):
        assert cnt[p] == 3
        assert [(int(r["clust"]), int(r["tree"])) for r in rec[p]] == ranking([wants[p]], [0.0], 3)
        assert all(int(r["param"]) == p for r in rec[p])
    for bad in (0.0, 0.5, -0.1, float("nan")):
        with pytest.raises(shim.BamseCudaError):
            ctx.upload_problem([bad], [4], var[None], ref[None])


def test_more_ranks_requested_than_trees_exist(ctx, shim):
    """K = 2 and K = 1 hold three trees together: topk = 5 returns three records, then invalid slots
    (clust -1, totals -inf, parents padded with -2)."""
    var = np.zeros((2, 2, 2), dtype=np.int64)
    ref = np.zeros((2, 2, 2), dtype=np.int64)
    var[0], ref[0] = [[40, 10], [5, 30]], [[160, 190], [195, 170]]
    var[1, :1], ref[1, :1] = [[22, 31]], [[178, 169]]
    ctx.upload_problem([0.003], [2, 1], var, ref)
    by_clust = [oracle_all_trees(var[0], ref[0], 0.003), oracle_all_trees(var[1, :1], ref[1, :1], 0.003)]
    const = [0.0, -3.5]
    for prec in (shim.FP32, shim.FP64):
        rec, par, cnt = ctx.score_topk(np.array([const]), topk=5, precision=prec)
        assert cnt[0] == 3 and ctx.last_eval_count == 3
        assert [(int(r["clust"]), int(r["tree"])) for r in rec[0, :3]] == ranking(by_clust, const, 3)
        assert all(int(r["clust"]) == -1 and r["total"] == -np.inf for r in rec[0, 3:])
        assert np.all(par[0, 3:] == -2)
        for j in range(3):
            K = [2, 1][int(rec[0, j]["clust"])]
            assert [int(x) for x in par[0, j, :K]] == orc.decode_tree(K, int(rec[0, j]["tree"]))
            assert np.all(par[0, j, K:] == -2)


def test_empty_ranges_unreachable_thresholds_and_empty_lists(ctx, shim):
    rng = np.random.default_rng(7)
    var = rng.integers(0, 150, size=(2, 4, 2)).astype(np.int64)
    ref = rng.integers(100, 300, size=(2, 4, 2)).astype(np.int64)
    errs = [0.003, 0.01]
    ctx.upload_problem(errs, [4, 4], var, ref)
    s0, w0 = ctx.score_range(0, 1, 0, 64, shim.FP64)
    # the first clustering contributes nothing: begin == end
    begin = np.array([17, 0], dtype=np.uint64)
    end = np.array([17, 64], dtype=np.uint64)
    for prec in (shim.FP32, shim.FP64):
        rec, _, cnt = ctx.score_topk(np.zeros((2, 2)), None, begin, end, topk=4, precision=prec)
        assert cnt.tolist() == [4, 4] and ctx.last_eval_count == 2 * 64
        assert all(int(r["clust"]) == 1 for r in rec.ravel())
        assert [int(r["tree"]) for r in rec[0]] == [i for _, i in ranking([(s0, w0)], [0.0], 4)]
        # both ranges empty
        rec, _, cnt = ctx.score_topk(np.zeros((2, 2)), None, begin, begin, topk=4, precision=prec)
        assert cnt.tolist() == [0, 0] and all(int(r["clust"]) == -1 for r in rec.ravel())
        # thresholds nothing reaches, for the second parameter set only
        thr = np.array([[-np.inf, -np.inf], [1e30, 1e30]])
        rec, _, cnt = ctx.score_topk(np.zeros((2, 2)), thr, topk=4, precision=prec)
        assert cnt.tolist() == [4, 0]
        assert all(int(r["clust"]) == -1 for r in rec[1])
        # a threshold exactly at the best tree's score keeps that tree (>=), one ulp above it keeps none of clustering 1
        best = float(np.max(s0 + w0))
        thr = np.array([[1e30, best], [1e30, 1e30]])
        rec, _, cnt = ctx.score_topk(np.zeros((2, 2)), thr, topk=4, precision=prec)
        assert cnt.tolist() == [1, 0] and int(rec[0, 0]["tree"]) == int(np.argmax(s0 + w0))
    # empty item list / empty dense range: nothing to do, no error
    out = ctx.score_list(np.zeros(0, dtype=np.int32), np.zeros((0, 4), dtype=np.int8), np.zeros((0, 4), dtype=np.int8))
    assert out.shape == (2, 0)
    s, w = ctx.score_range(0, 0, 5, 0, shim.FP32)
    assert len(s) == 0 and len(w) == 0
    with pytest.raises(shim.BamseCudaError):
        ctx.score_range(0, 0, 60, 5, shim.FP32)            # beyond K^(K-1) = 64
    with pytest.raises(shim.BamseCudaError):
        ctx.score_topk(np.zeros((2, 2)), topk=shim.TOPK_MAX + 1)


def test_duplicated_clusterings_tie_exactly_and_rank_by_clustering_index(ctx, shim):
    """the same clustering three times in one problem (KMeans restarts that converge to the same assignment are
    de-duplicated by the host, but nothing in the ABI forbids it): every tree is tied exactly with its two copies;
    the order is (total desc, clustering asc, tree asc) in both precisions and the cut is still proved."""
    rng = np.random.default_rng(33)
    v = rng.integers(0, 250, size=(5, 3)).astype(np.int64)
    r = rng.integers(150, 500, size=(5, 3)).astype(np.int64)
    var, ref = np.stack([v, v, v]), np.stack([r, r, r])
    ctx.upload_problem([0.003], [5, 5, 5], var, ref)
    s64, w64 = ctx.score_range(0, 0, 0, 625, shim.FP64)
    for c in (1, 2):
        s, _ = ctx.score_range(0, c, 0, 625, shim.FP64)
        np.testing.assert_array_equal(s, s64)
    want = ranking([(s64, w64)] * 3, [0.0, 0.0, 0.0], 12)
    for prec in (shim.FP32, shim.FP64):
        rec, _, cnt = ctx.score_topk(np.zeros((1, 3)), topk=12, precision=prec)
        assert cnt[0] == 12 and ctx.last_topk_verified
        assert [(int(x["clust"]), int(x["tree"])) for x in rec[0]] == want
    # a constant that prefers the last copy turns the order of the copies around
    rec, _, _ = ctx.score_topk(np.array([[0.0, 1e-4, 2e-4]]), topk=3, precision=shim.FP32)
    assert [int(x["clust"]) for x in rec[0]] == [2, 1, 0] and len({int(x["tree"]) for x in rec[0]}) == 1


def test_a_context_is_reused_for_problems_of_other_shapes(shim):
    """buffers, plans and data-independent tables survive uploads: A (K = 5, 4; M = 3), then B (K = 3; M = 1), then
    a larger C (K = 6, 2; M = 4; two error rates), then A again -- every answer equals a fresh context's."""
    rng = np.random.default_rng(5)

    def problem(Ks, M):
        Kmax = max(Ks)
        var = np.zeros((len(Ks), Kmax, M), dtype=np.int64)
        ref = np.zeros((len(Ks), Kmax, M), dtype=np.int64)
        for c, K in enumerate(Ks):
            var[c, :K] = rng.integers(0, 200, size=(K, M))
            ref[c, :K] = rng.integers(100, 400, size=(K, M))
        return Ks, var, ref

    probs = {"A": ([0.003], problem([5, 4], 3)), "B": ([0.02], problem([3], 1)),
             "C": ([0.001, 0.004], problem([6, 2], 4))}

    def answer(c, name):
        errs, (Ks, var, ref) = probs[name]
        c.upload_problem(errs, Ks, var, ref)
        const = np.zeros((len(errs), len(Ks)))
        rec, par, cnt = c.score_topk(const, topk=7, precision=shim.FP32)
        dense = c.score_range(len(errs) - 1, 0, 0, Ks[0] ** (Ks[0] - 1), shim.FP32)[0]
        return rec, par, cnt, dense

    fresh = {}
    for name in probs:
        with shim.Context(0) as c:
            fresh[name] = answer(c, name)
    with shim.Context(0) as c:
        for name in ("A", "B", "C", "A", "C", "B"):
            rec, par, cnt, dense = answer(c, name)
            want = fresh[name]
            np.testing.assert_array_equal(cnt, want[2])
            np.testing.assert_array_equal(rec["tree"], want[0]["tree"])
            np.testing.assert_array_equal(rec["clust"], want[0]["clust"])
            np.testing.assert_array_equal(rec["total"], want[0]["total"])
            np.testing.assert_array_equal(par, want[1])
            np.testing.assert_allclose(dense, want[3], rtol=1e-13, atol=0)
