"""Round-2 parity tests (B200 box: python -m pytest tests -m gpu): the chain reference -> oracle -> CUDA
closed where BASELINE configs 3 and 4 live (K = 8 ... 12, reference-produced goldens), on the
reference's own sample input bamse.dat, at full config-3 / config-5 size, and at the preselection cut
(ties, extreme coverage)."""
import math
import warnings

import numpy as np
import pytest

import bamse_oracle as orc
from conftest import load_golden

pytestmark = pytest.mark.gpu

FP32_RTOL = 1e-6      # north_star: fp32 mode within 1e-6 relative
FP64_RTOL = 1e-10     # fp64 check mode within 1e-10 relative


@pytest.fixture(scope="module")
def shim():
    from bamse_b200 import cuda_shim
    return cuda_shim


@pytest.fixture(scope="module")
def ctx(shim):
    c = shim.Context(0)
    yield c
    c.close()


def _pack(Kmax, items):
    n = len(items)
    clust = np.zeros(n, dtype=np.int32)
    par = np.full((n, Kmax), -2, dtype=np.int8)
    nod = np.full((n, Kmax), -1, dtype=np.int8)
    for i, (c, tree, nodes) in enumerate(items):
        clust[i] = c
        par[i, :len(tree)] = tree
        nod[i, :len(nodes)] = nodes
    return clust, par, nod


# ------------------------------------------------------------------ reference goldens at K = 8 .. 12
def test_highk_reference_goldens(ctx, shim):
    """35 get_tree_score values produced by the unmodified reference at K = 8, 9, 10, 11, 12 (random trees,
    chain, star, wide middle node; coverage 200 ... 100 000): fp64 check mode <= 1e-10, fp32 production and
    fp32 log-domain kernels <= 1e-6, and the device canonizeF weight of the same tree index is exact."""
    cases = load_golden("tree_scores_highk.json")
    assert len(cases) >= 30 and {c["K_prior"] for c in cases} == {8, 9, 10, 11, 12}
    worst = {"fp64": 0.0, "fp32": 0.0, "fp32l": 0.0}
    for case in cases:
        As = np.array(case["var"])[None]
        Bs = np.array(case["ref"])[None]
        K = As.shape[1]
        ctx.upload_problem([case["err"]], [K], As, Bs)
        clust, par, nod = _pack(K, [(0, case["tree"], case["nodes"])])
        want = case["total"]
        got = {"fp64": ctx.score_list(clust, par, nod, shim.FP64)[0, 0],
               "fp32": ctx.score_list(clust, par, nod, shim.FP32)[0, 0],
               "fp32l": ctx.score_list(clust, par, nod, shim.FP32_LOGDOMAIN)[0, 0]}
        for k, v in got.items():
            worst[k] = max(worst[k], abs(v - want) / abs(want))
        assert abs(got["fp64"] - want) <= FP64_RTOL * abs(want), (case["id"], got, want)
        assert abs(got["fp32"] - want) <= FP32_RTOL * abs(want), (case["id"], got, want)
        assert abs(got["fp32l"] - want) <= FP32_RTOL * abs(want), (case["id"], got, want)
        # the same tree by INDEX through the enumeration kernels: score and canonizeF weight
        idx = orc.encode_tree(case["tree"])
        assert shim.decode_tree(K, idx) == [int(x) for x in case["tree"]]
        s32, w32 = ctx.score_range(0, 0, idx, 1, shim.FP32)
        s64, w64 = ctx.score_range(0, 0, idx, 1, shim.FP64)
        assert w32[0] == w64[0] == math.log(case["scre"])
        assert abs(s64[0] - want) <= FP64_RTOL * abs(want)
        assert abs(s32[0] - want) <= FP32_RTOL * abs(want)
    print("worst relative errors vs the reference:", worst)


@pytest.mark.parametrize("K", [6, 7, 8, 9, 10, 11, 12])
def test_device_canon_weight_on_random_indices(ctx, shim, K):
    """canonizeF 'scre' (tree_tools.py:47-63) computed on the device for ~300 tree indices per K (random
    starts x consecutive runs, incl. the ends of the index space) equals the oracle's, bit for bit; the
    library decoder equals the oracle decoder on the same indices."""
    rng = np.random.default_rng(900 + K)
    M = 1
    As = rng.integers(5, 300, size=(K, M))
    Bs = rng.integers(5, 300, size=(K, M))
    ctx.upload_problem([0.003], [K], As[None], Bs[None])
    T = K ** (K - 1)
    run = 30
    starts = [0, T - run] + [int(x) for x in rng.integers(0, T - run, size=8)]
    for b in starts:
        _, w = ctx.score_range(0, 0, b, run, shim.FP32)
        want = [math.log(orc.canon_weight(orc.decode_tree(K, b + i))[1]) for i in range(run)]
        np.testing.assert_array_equal(w, np.array(want))
        for i in (0, run - 1):
            assert shim.decode_tree(K, b + i) == orc.decode_tree(K, b + i)


# ------------------------------------------------------------------ bamse.dat + a K = 5 kept set
def test_bamse_dat_and_k5_search_goldens(ctx, shim):
    """the reference's own sample input (bamse.dat, coverage 10 000: the hard dynamic-range case) at its
    KMeans-4 and KMeans-5 clusterings, and a K = 5 simdata input, through reorder -> get_candidate_tree ->
    threshold -> get_trees2 -> score assembly: labels, candidate, threshold, kept set (reference-kept subset
    of ours, equal scores), totalscore and the best tree equal what the reference's own methods returned."""
    from bamse_b200.clustering import clustering
    cases = load_golden("search_big.json")
    assert [c["spec"].get("source") for c in cases[:2]] == ["bamse.dat", "bamse.dat"]
    for case in cases:
        for precision in (shim.FP32, shim.FP64):
            c = clustering(case["am"], case["bm"], case["assign_in"], np.array(case["pf"]), case["spec"]["err"], ctx=ctx)
            assert c.cluster_prior == pytest.approx(case["cluster_prior"], rel=1e-13)
            assert c.maxlikelihood == pytest.approx(case["maxlikelihood"], rel=1e-13)
            c.reorder()
            assert [int(x) for x in c.order] == case["order"]
            assert c.assign.tolist() == case["assign_out"]
            np.testing.assert_array_equal(c.As, case["As"])
            np.testing.assert_array_equal(c.Bs, case["Bs"])
            cand = c.get_candidate_tree()[0]
            assert cand["tree"] == case["candidate"]
            K = c.As.shape[0]
            thr = c.get_tree_score(cand["tree"], list(range(K))) + math.log(orc.canon_weight(cand["tree"])[1])
            assert thr == pytest.approx(case["threshold"], rel=1e-10)
            trees = c.get_trees2(cand, precision=precision)
            kept = {tuple(t["tree"]): t for t in trees}
            for ref in case["kept"]:
                assert tuple(ref["tree"]) in kept
                assert kept[tuple(ref["tree"])]["score"] == pytest.approx(ref["score"] + math.log(ref["scre"]), rel=1e-10)
            c.analyse_trees_regular2(solve_fractions=False)
            best_ref = max(case["kept"], key=lambda r: r["totalscore"])
            assert tuple(c.trees[0]["tree"]) == tuple(best_ref["tree"])
            assert c.trees[0]["totalscore"] == pytest.approx(best_ref["totalscore"], rel=1e-10)
            assert c.trees[0]["root"] == best_ref["root"]


# ------------------------------------------------------------------ full-size configs 3 and 5
def test_config3_full_size_fp32_equals_check_mode(ctx, shim):
    """BASELINE config 3 at full size (20 samples, 2000 mutations, -nc 8 -n 100: ~9e6 labeled rooted trees)
    through the fp32 production path, and -- because a complete fp64 pass over 1.8e8 tree-samples takes
    minutes -- the interleaved third of it (what rank 0 of 3 GPUs scores: 3e6 trees, every clustering, all 20
    samples) through both paths: identical top-t trees, clusterings and assignments; totals are fp64 in both."""
    from bamse_b200 import workloads
    from bamse_b200.clustering import analyse_clusterings
    w32 = workloads.build("c3", ctx=ctx)
    assert w32["M"] == 20 and w32["n_evals_per_param"] > 5_000_000
    full = analyse_clusterings(w32["clusts"], 5, ctx=ctx, precision=shim.FP32, solve_fractions=False)
    assert ctx.last_topk_verified and len(full) >= 1 and ctx.last_eval_count == w32["n_evals_per_param"]
    t32 = analyse_clusterings(workloads.build("c3", ctx=ctx)["clusts"], 5, ctx=ctx, precision=shim.FP32,
                              solve_fractions=False, tree_shard=(0, 3))
    assert ctx.last_topk_verified
    w64 = workloads.build("c3", ctx=ctx)
    t64 = analyse_clusterings(w64["clusts"], 5, ctx=ctx, precision=shim.FP64, solve_fractions=False, tree_shard=(0, 3))
    assert len(t32) == len(t64) >= 1
    assert full[0]["totalscore"] >= t32[0]["totalscore"]
    for a, b in zip(t32, t64):
        assert a["tree"] == b["tree"] and a["clustering_index"] == b["clustering_index"]
        assert a["assign"].tolist() == b["assign"].tolist()
        assert a["totalscore"] == b["totalscore"]               # same fp64 list-scoring kernel in both modes


def test_config5_full_size_parameter_sweep(ctx, shim):
    """BASELINE config 5 at full size: 64 (e, s) pairs x 10 samples x 500 mutations, -nc 6, ONE upload and
    ONE enumeration launch; the top-5 of every parameter set is identical in fp32 and fp64 mode, and equals
    a single-parameter upload of the same e (the sparsity s never reaches a score)."""
    from bamse_b200 import workloads
    from bamse_b200.clustering import _pack_problem
    w = workloads.build("c5", ctx=ctx)
    clusts, errs = w["clusts"], w["errs"]
    assert len(errs) == 64
    for c in clusts:
        c.reorder()
    Ks, var, ref, Kmax = _pack_problem(clusts)
    ctx.upload_problem(errs, Ks, var, ref)
    const = np.array([[c.cluster_prior + c.maxlikelihood for c in clusts] for _ in errs])
    l0 = ctx.launch_count
    r32, p32, n32 = ctx.score_topk(const, topk=5, precision=shim.FP32)
    assert ctx.last_topk_verified
    r64, p64, n64 = ctx.score_topk(const, topk=5, precision=shim.FP64)
    assert ctx.last_eval_count == 64 * w["n_evals_per_param"] and ctx.launch_count - l0 < 40
    assert np.all(n32 == 5) and np.all(n64 == 5)
    for f in ("tree", "clust", "param"):
        np.testing.assert_array_equal(r32[f], r64[f])
    np.testing.assert_array_equal(r32["total"], r64["total"])
    np.testing.assert_array_equal(p32, p64)
    for p in (0, 31, 63):
        ctx.upload_problem([errs[p]], Ks, var, ref)
        one, _, _ = ctx.score_topk(const[p:p + 1], topk=5, precision=shim.FP32)
        assert one["tree"][0].tolist() == r32["tree"][p].tolist()
        np.testing.assert_allclose(one["total"][0], r32["total"][p], rtol=1e-12)
    assert r32["total"][0, 0] != r32["total"][63, 0]


# ------------------------------------------------------------------ the preselection cut
def test_verified_cut_with_exact_ties(ctx, shim):
    """two identical clusters: every tree has a twin (labels swapped) with the same score up to rounding, so
    the fp32 selection cannot separate them.  The cut is proven (growing the short list when needed) and
    fp32 and fp64 mode return the same records for every t, including cuts that fall between twins."""
    rng = np.random.default_rng(42)
    K, M = 6, 4
    As = rng.integers(20, 900, size=(K, M))
    Bs = rng.integers(20, 900, size=(K, M))
    As[4], Bs[4] = As[3], Bs[3]                         # clusters 3 and 4 are indistinguishable
    As[5], Bs[5] = As[3], Bs[3]                         # ... and 5 too: classes of up to 6 tied trees
    ctx.upload_problem([0.002], [K], As[None], Bs[None])
    const = np.zeros((1, 1))
    for topk in (1, 2, 3, 5, 8, 13, 40, 64):
        r32, p32, n32 = ctx.score_topk(const, topk=topk, precision=shim.FP32)
        assert ctx.last_topk_verified and ctx.last_topk_preselected > topk
        r64, p64, n64 = ctx.score_topk(const, topk=topk, precision=shim.FP64)
        assert ctx.last_topk_verified
        assert n32[0] == n64[0] == topk
        np.testing.assert_array_equal(r32["tree"], r64["tree"])
        np.testing.assert_array_equal(r32["total"], r64["total"])
        assert np.all(np.diff(r32["total"][0]) <= 0)
    # ranking against dense check-mode scores of the whole space
    s, lw = ctx.score_range(0, 0, 0, K ** (K - 1), shim.FP64)
    full = s + lw
    r32, _, _ = ctx.score_topk(const, topk=64, precision=shim.FP32)
    order = np.lexsort((np.arange(len(full)), -full))[:64]
    np.testing.assert_allclose(r32["total"][0], full[order], rtol=1e-12)
    # the set can only differ inside a group tied to rounding with the 64th value
    lo = full[order[-1]]
    sure = {int(i) for i in np.nonzero(full > lo + 1e-9 * abs(lo))[0]}
    assert sure <= {int(t) for t in r32["tree"][0]}


def test_all_trees_tied_reports_unverified(ctx, shim):
    """K identical clusters with zero reads: all K^(K-1) scores coincide (up to the canonizeF weight), more
    trees are tied than the internal short list can hold -> the call still returns a valid ranking but says
    so (bamse_last_topk_verified = 0, a Python warning)."""
    K, M = 6, 2
    As = np.zeros((K, M), dtype=np.int64)
    Bs = np.zeros((K, M), dtype=np.int64)
    ctx.upload_problem([0.01], [K], As[None], Bs[None])
    with warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        rec, par, cnt = ctx.score_topk(np.zeros((1, 1)), topk=4, precision=shim.FP32)
    assert cnt[0] == 4 and not ctx.last_topk_verified and ctx.last_topk_preselected == 256
    assert any("tied" in str(w.message) for w in wlist)
    assert np.all(np.isfinite(rec["total"]))
    # a separable problem right after it verifies again
    rng = np.random.default_rng(1)
    As = rng.integers(20, 900, size=(K, M))
    Bs = rng.integers(20, 900, size=(K, M))
    ctx.upload_problem([0.01], [K], As[None], Bs[None])
    ctx.score_topk(np.zeros((1, 1)), topk=4, precision=shim.FP32)
    assert ctx.last_topk_verified


def test_extreme_coverage_many_samples(ctx, shim):
    """coverage 5e6 per cluster and sample, M = 20, K = 5: |score| reaches 1e6-1e7, where an fp32 ulp of a
    log value is ~0.5 and the fixed margins of round 1 were unproven.  Scores stay within 1e-6 relative of
    the check mode, and the verified cut returns the fp64 ranking."""
    rng = np.random.default_rng(5_000_000)
    K, M, cov = 5, 20, 5_000_000
    frac = np.sort(rng.dirichlet(np.ones(K + 1))[:K])[::-1]
    cm = np.clip(np.cumsum(frac[::-1])[::-1][:, None] * rng.uniform(0.6, 1.0, size=(K, M)), 0, 1)
    As = rng.binomial(cov, cm * 0.499 + 0.001)
    Bs = cov - As
    ctx.upload_problem([0.001], [K], As[None], Bs[None])
    T = K ** (K - 1)
    s64, _ = ctx.score_range(0, 0, 0, T, shim.FP64)
    s32, _ = ctx.score_range(0, 0, 0, T, shim.FP32)
    np.testing.assert_allclose(s32, s64, rtol=FP32_RTOL)
    const = np.zeros((1, 1))
    for topk in (1, 5, 20):
        r32, p32, n32 = ctx.score_topk(const, topk=topk, precision=shim.FP32)
        assert ctx.last_topk_verified
        r64, p64, n64 = ctx.score_topk(const, topk=topk, precision=shim.FP64)
        np.testing.assert_array_equal(r32["tree"], r64["tree"])
        np.testing.assert_array_equal(r32["total"], r64["total"])
    D = np.array([[orc.distrib(As[k, m], Bs[k, m], 0.001) for m in range(M)] for k in range(K)])
    for i in (0, 311, 624):
        want = orc.tree_score_samples(orc.decode_tree(K, i), D, orc.prior_const(K)).sum()
        assert s64[i] == pytest.approx(want, rel=FP64_RTOL)


# ------------------------------------------------------------------ host-side state
def test_shard_is_restored_and_all_ranked_trees_are_kept(ctx, shim, golden_search):
    """analyse_clusterings(tree_shard=...) must not leave the process-wide context sharded, and a clustering
    with several ranked trees keeps all of them."""
    from bamse_b200.clustering import analyse_clusterings, clustering
    case = [c for c in golden_search if c["spec"]["K"] == 4][0]
    mk = lambda: clustering(case["am"], case["bm"], case["assign_in"], np.array(case["pf"]), case["spec"]["err"], ctx=ctx)
    whole = analyse_clusterings([mk()], 6, ctx=ctx, solve_fractions=False)
    parts = []
    for r in range(2):
        parts += analyse_clusterings([mk()], 6, ctx=ctx, solve_fractions=False, tree_shard=(r, 2))
        c = mk()
        c.reorder()
        c.get_trees2(c.get_candidate_tree()[0])
        assert ctx.last_eval_count == 64                # everything again: the shard did not stick
    parts.sort(key=lambda t: (-t["totalscore"], t["tree_index"]))
    assert [t["tree_index"] for t in parts[:len(whole)]] == [t["tree_index"] for t in whole]
    c = mk()
    out = analyse_clusterings([c], 6, ctx=ctx, solve_fractions=False)
    assert len(c.trees) == len(out) and [t["tree"] for t in c.trees] == [t["tree"] for t in out]
    assert all("totalscore" in t for t in c.trees)
