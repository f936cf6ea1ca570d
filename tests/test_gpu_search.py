"""The callers either side of the scoring path, through the host mirror of the reference's
`clustering` class (bamse_b200.clustering) on the GPU: reorder, greedy candidate, threshold,
kept set, score assembly -- against results of the reference's own methods
(tests/golden/search.json) -- plus robustness cases of the fp32 production kernel."""
import math

import numpy as np
import pytest

import bamse_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def shim():
    from bamse_b200 import cuda_shim
    return cuda_shim


@pytest.fixture(scope="module")
def ctx(shim):
    c = shim.Context(0)
    yield c
    c.close()


def test_reorder_candidate_threshold_kept_match_reference(ctx, golden_search):
    from bamse_b200.clustering import clustering
    for case in golden_search:
        c = clustering(case["am"], case["bm"], case["assign_in"], np.array(case["pf"]), case["spec"]["err"], ctx=ctx)
        assert c.cluster_prior == pytest.approx(case["cluster_prior"], rel=1e-13)
        assert c.maxlikelihood == pytest.approx(case["maxlikelihood"], rel=1e-13)
        c.reorder()
        assert [int(x) for x in c.order] == case["order"]
        assert c.assign.tolist() == case["assign_out"]
        np.testing.assert_array_equal(c.As, case["As"])
        cand = c.get_candidate_tree()[0]
        assert cand["tree"] == case["candidate"]
        K = c.As.shape[0]
        thr = c.get_tree_score(cand["tree"], list(range(K))) + math.log(orc.canon_weight(cand["tree"])[1])
        assert thr == pytest.approx(case["threshold"], rel=1e-10)
        trees = c.get_trees2(cand)
        kept = {tuple(t["tree"]): t for t in trees}
        for ref in case["kept"]:                       # reference-kept subset of GPU-kept, equal scores
            assert tuple(ref["tree"]) in kept
            assert kept[tuple(ref["tree"])]["score"] == pytest.approx(ref["score"] + math.log(ref["scre"]), rel=1e-10)
        c.analyse_trees_regular2(solve_fractions=False)
        for t in c.trees:
            for ref in case["kept"]:
                if tuple(ref["tree"]) == tuple(t["tree"]):
                    assert t["totalscore"] == pytest.approx(ref["totalscore"], rel=1e-10)
                    assert t["root"] == ref["root"]
        # the reference's best solution is our best solution
        best_ref = max(case["kept"], key=lambda r: r["totalscore"])
        assert tuple(c.trees[0]["tree"]) == tuple(best_ref["tree"])


def test_batch_driver_equals_per_clustering_calls(ctx, golden_search):
    """analyse_clusterings (the replacement of bamse.py:91-106) over several clusterings at
    once returns the global ranking the per-clustering reference results imply."""
    from bamse_b200.clustering import analyse_clusterings, clustering
    cases = [c for c in golden_search if c["spec"]["err"] == 0.003]
    assert len(cases) >= 2
    # same error rate, different data sets are fine for the ranking logic as long as M matches
    cases = [c for c in cases if len(c["am"][0]) == len(cases[0]["am"][0])]
    clusts = [clustering(c["am"], c["bm"], c["assign_in"], np.array(c["pf"]), 0.003, ctx=ctx) for c in cases]
    out = analyse_clusterings(clusts, top_trees=4, ctx=ctx, solve_fractions=True)
    ref_all = sorted([(r["totalscore"], ci, tuple(r["tree"])) for ci, c in enumerate(cases) for r in c["kept"]], reverse=True)
    got = [(t["totalscore"], t["clustering_index"], tuple(t["tree"])) for t in out]
    # every reference solution appears, with its score, in the GPU ranking (the GPU list may
    # additionally hold qualifying trees that the reference's heuristic bound pruned)
    for tot, ci, tree in ref_all[:2]:
        match = [g for g in got if g[1] == ci and g[2] == tree]
        assert match and match[0][0] == pytest.approx(tot, rel=1e-10)
    assert got[0][1:] == ref_all[0][1:]
    assert all(got[i][0] >= got[i + 1][0] for i in range(len(got) - 1))
    t = out[0]
    assert t["clone_proportions"].shape == (len(t["tree"]), len(cases[0]["am"][0]))
    assert np.all(t["vaf"] <= 1 + 1e-6) and np.all(t["clone_proportions"] >= -1e-9)


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


@pytest.mark.parametrize("framed", ["0", "1"])
@pytest.mark.parametrize("cov", [0, 3, 40, 2000, 60000, 5000000])
def test_fp32_kernel_over_the_dynamic_range(ctx, shim, cov, framed, monkeypatch):
    """from empty / nearly flat tables (every term of a convolution matters) to 5e6 reads per
    cluster (peaks far narrower than a grid cell, steps of thousands of nats): the fp32
    production kernel stays within 1e-6 of the fp64 check mode on every tree of K = 4, and the
    check mode agrees with the numpy oracle."""
    monkeypatch.setenv("BAMSE_FRAMED", framed)      # both instantiations of the fp32 kernel
    rng = np.random.default_rng(cov + 1)
    K, M = 4, 3
    prev = np.array([[0.9, 0.8, 0.95], [0.5, 0.1, 0.6], [0.3, 0.6, 0.02], [0.1, 0.05, 0.3]])
    As = rng.binomial(cov, prev * 0.499 + 0.001) if cov else np.zeros((K, M), dtype=int)
    Bs = cov - As
    if cov == 3:
        As[1, 0], Bs[1, 0] = 0, 3          # a == 0 and b == 0 corners
        As[2, 1], Bs[2, 1] = 3, 0
    ctx.upload_problem([0.001], [K], As[None], Bs[None])
    s64, _ = ctx.score_range(0, 0, 0, 64, shim.FP64)
    s32, _ = ctx.score_range(0, 0, 0, 64, shim.FP32)
    s32l, _ = ctx.score_range(0, 0, 0, 64, shim.FP32_LOGDOMAIN)
    assert np.all(np.isfinite(s64))
    np.testing.assert_allclose(s32, s64, rtol=1e-6, atol=2e-6)
    np.testing.assert_allclose(s32l, s64, rtol=1e-6, atol=2e-6)
    D = np.array([[orc.distrib(As[k, m], Bs[k, m], 0.001) for m in range(M)] for k in range(K)])
    for i in (0, 7, 21, 42, 63):
        want = orc.tree_score_samples(orc.decode_tree(K, i), D, orc.prior_const(K)).sum()
        assert s64[i] == pytest.approx(want, rel=1e-10, abs=1e-12)


def test_framed_variant_on_a_whole_clustering(ctx, shim, monkeypatch):
    """all 7776 trees of a K = 6 clustering with flat-ish tables (where the framed path is used for
    most blocks): framed and exponential instantiations agree with each other to fp32 accuracy and
    return the same top-k (fp64 re-scored) as the check mode."""
    rng = np.random.default_rng(123)
    K, M = 6, 4
    frac = np.sort(rng.dirichlet(np.ones(K + 1))[:K])[::-1]
    cm = np.clip(np.cumsum(frac[::-1])[::-1][:, None] * rng.uniform(0.6, 1.0, size=(K, M)), 0, 1)
    As = rng.binomial(150, cm * 0.497 + 0.003)
    Bs = 150 - As
    out = {}
    for framed in ("0", "1"):
        monkeypatch.setenv("BAMSE_FRAMED", framed)
        ctx.upload_problem([0.003], [K], As[None], Bs[None])
        out[framed] = (ctx.score_range(0, 0, 0, K ** (K - 1), shim.FP32)[0],
                       ctx.score_topk(np.zeros((1, 1)), topk=8, precision=shim.FP32)[0])
    monkeypatch.delenv("BAMSE_FRAMED")
    np.testing.assert_allclose(out["1"][0], out["0"][0], rtol=5e-7)
    ctx.upload_problem([0.003], [K], As[None], Bs[None])
    ref, _, _ = ctx.score_topk(np.zeros((1, 1)), topk=8, precision=shim.FP64)
    for framed in ("0", "1"):
        assert out[framed][1]["tree"].tolist() == ref["tree"].tolist()
        np.testing.assert_allclose(out[framed][1]["total"], ref["total"], rtol=1e-10)
    s64 = ctx.score_range(0, 0, 0, 1000, shim.FP64)[0]
    np.testing.assert_allclose(out["1"][0][:1000], s64, rtol=1e-6)


def test_large_shapes_fp32_vs_check_mode(ctx, shim):
    """BASELINE config-3/4 shapes (K = 8, M = 20 and K = 10, M = 5): random trees, fp32 vs fp64."""
    rng = np.random.default_rng(77)
    for K, M, cov, n in [(8, 20, 3000, 48), (10, 5, 800, 48), (12, 2, 500, 24)]:
        frac = np.sort(rng.dirichlet(np.ones(K + 1))[:K])[::-1]
        cm = np.clip(np.cumsum(frac[::-1])[::-1][:, None] * rng.uniform(0.5, 1.0, size=(K, M)), 0, 1)
        As = rng.binomial(cov, cm * 0.497 + 0.003)
        Bs = cov - As
        ctx.upload_problem([0.003], [K], As[None], Bs[None])
        items = []
        for _ in range(n):
            perm = rng.permutation(K)
            parent = [-2] * K
            parent[perm[0]] = -1
            for i in range(1, K):
                parent[perm[i]] = int(perm[rng.integers(0, i)])
            items.append((0, parent, list(range(K))))
        items.append((0, [-1] + list(range(K - 1)), list(range(K))))          # the chain: scans only
        items.append((0, [-1] + [0] * (K - 1), list(range(K))))              # the star: K-2 convolutions
        clust, par, nod = _pack(K, items)
        s64 = ctx.score_list(clust, par, nod, shim.FP64)[0]
        s32 = ctx.score_list(clust, par, nod, shim.FP32)[0]
        np.testing.assert_allclose(s32, s64, rtol=1e-6)


def test_relabelling_invariance(ctx, shim):
    """a tree and its node table permuted together give the same score (child order only
    changes the association order of the convolutions): property test at K = 7."""
    rng = np.random.default_rng(11)
    K, M = 7, 4
    As = rng.integers(5, 4000, size=(K, M))
    Bs = rng.integers(5, 4000, size=(K, M))
    ctx.upload_problem([0.002], [K], As[None], Bs[None])
    parent = [-1, 0, 0, 1, 1, 2, 5]
    items = [(0, parent, list(range(K)))]
    for _ in range(6):
        perm = [int(x) for x in rng.permutation(K)]       # local position i holds cluster perm[i]
        inv = {g: i for i, g in enumerate(perm)}
        p2 = [(-1 if parent[perm[i]] == -1 else inv[parent[perm[i]]]) for i in range(K)]
        items.append((0, p2, perm))
    clust, par, nod = _pack(K, items)
    s64 = ctx.score_list(clust, par, nod, shim.FP64)[0]
    s32 = ctx.score_list(clust, par, nod, shim.FP32)[0]
    np.testing.assert_allclose(s64, s64[0], rtol=1e-12)
    np.testing.assert_allclose(s32, s64[0], rtol=1e-6)


def test_parameter_sweep_equals_separate_uploads(ctx, shim):
    """config-5 style: several error rates in one upload == one upload per error rate; the
    sparsity parameter never reaches the scorer (the reference ignores it)."""
    rng = np.random.default_rng(3)
    K, M = 5, 3
    As = rng.integers(20, 900, size=(K, M))
    Bs = rng.integers(20, 900, size=(K, M))
    errs = [1e-4, 1e-3, 5e-3, 1e-2]
    ctx.upload_problem(errs, [K], As[None], Bs[None])
    all_p = [ctx.score_range(p, 0, 0, 625, shim.FP32)[0] for p in range(len(errs))]
    rec_all, _, _ = ctx.score_topk(np.zeros((len(errs), 1)), topk=3, precision=shim.FP32)
    for p, e in enumerate(errs):
        ctx.upload_problem([e], [K], As[None], Bs[None])
        one, _ = ctx.score_range(0, 0, 0, 625, shim.FP32)
        np.testing.assert_array_equal(one, all_p[p])
        rec, _, _ = ctx.score_topk(np.zeros((1, 1)), topk=3, precision=shim.FP32)
        assert rec["tree"][0].tolist() == rec_all["tree"][p].tolist()
    assert not np.allclose(all_p[0], all_p[-1])


def test_config1_pipeline_fp32_equals_check_mode(ctx, shim):
    """BASELINE config 1 (3 samples, 50 mutations, -nc 5 -n 10 -t 5, e = 0.001) end to end:
    identical top-t trees and cluster assignments in fp32 and in the fp64 check mode."""
    from bamse_b200 import workloads
    from bamse_b200.clustering import analyse_clusterings
    w32 = workloads.build("c1", ctx=ctx)
    t32 = analyse_clusterings(w32["clusts"], 5, ctx=ctx, precision=shim.FP32, solve_fractions=False)
    w64 = workloads.build("c1", ctx=ctx)
    t64 = analyse_clusterings(w64["clusts"], 5, ctx=ctx, precision=shim.FP64, solve_fractions=False)
    assert len(t32) == len(t64) >= 1
    for a, b in zip(t32, t64):
        assert a["tree"] == b["tree"] and a["clustering_index"] == b["clustering_index"]
        assert a["assign"].tolist() == b["assign"].tolist()
        assert a["totalscore"] == pytest.approx(b["totalscore"], rel=1e-10)   # both re-scored in fp64
    # and the best one satisfies the reference's keep rule for its clustering
    c = w64["clusts"][t64[0]["clustering_index"]]
    thr = c.get_tree_score(c.candidate, list(range(c.As.shape[0]))) + math.log(orc.canon_weight(c.candidate)[1])
    assert t64[0]["score"] >= thr - 1e-9 * abs(thr)


def test_cli_end_to_end(tmp_path, monkeypatch, ctx):
    """bamse.py -e .. -nc .. -n .. -t .. inputfile -> the reference's four kinds of output files."""
    import bamse_b200.clustering as C
    from bamse_b200 import cli, simdata
    d = simdata.simulate(3, 2, 30, seed=21)
    inp = tmp_path / "in.dat"
    simdata.write_input_file(str(inp), d["As"], d["Bs"])
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(C, "_default_ctx", ctx)
    assert cli.main(["-e", "0.003", "-nc", "4", "-n", "3", "-t", "2", "--seed", "5", str(inp)]) == 0
    txt = (tmp_path / "bamse_output.txt").read_text()
    assert txt.startswith("Solution Number 1\ntree = [") and "logscore = " in txt
    assert "Mutations Subclone Membership = ['" in txt
    for name in ("tree_number_0_subclones.dot", "tree_number_0_samples.dot", "tree_number_0.txt"):
        assert (tmp_path / name).read_text()
    assert (tmp_path / "tree_number_0_subclones.dot").read_text().startswith("digraph {\nSubcloneA [")


def test_early_exit_returns_the_same_topk(ctx, shim):
    """BAMSE_OPT_EARLY_EXIT: identical records with and without the exact sample-level early exit,
    while a good share of the (tree, sample) evaluations is skipped."""
    from bamse_b200 import workloads
    from bamse_b200.clustering import _pack_problem
    w = workloads.build("c1", ctx=ctx)
    clusts = w["clusts"]
    Ks, var, ref, Kmax = _pack_problem(clusts)
    ctx.upload_problem([0.001], Ks, var, ref)
    const = np.array([[c.cluster_prior + c.maxlikelihood for c in clusts]])
    # thresholds: the best score of each clustering minus 5 nats (what a good greedy candidate gives)
    thr = np.zeros((1, len(clusts)))
    for ci, K in enumerate(Ks):
        s, lw = ctx.score_range(0, ci, 0, int(K) ** (int(K) - 1), shim.FP64)
        thr[0, ci] = (s + lw).max() - 5.0
    total_samples = sum(int(K) ** (int(K) - 1) for K in Ks) * var.shape[2]
    out = {}
    for on in (False, True):
        ctx.set_early_exit(on)
        ctx.samples_evaluated(reset=True)
        out[on] = (ctx.score_topk(const, thr, topk=6, precision=shim.FP32), ctx.samples_evaluated(reset=True))
    ctx.set_early_exit(False)
    (rec0, par0, cnt0), n0 = out[False]
    (rec1, par1, cnt1), n1 = out[True]
    assert cnt0[0] == cnt1[0] >= 1
    assert rec0["tree"].tolist() == rec1["tree"].tolist() and rec0["clust"].tolist() == rec1["clust"].tolist()
    np.testing.assert_array_equal(rec0["total"], rec1["total"])
    assert n0 == total_samples and n1 < 0.8 * total_samples


def test_config2_full_size_fp32_equals_check_mode(ctx, shim):
    """BASELINE config 2 at full size (79 clusterings, 397 530 labeled rooted trees, 10 samples):
    the fp32 production path (with the exact early exit the CLI driver uses) and a complete fp64
    check-mode enumeration return identical top-t trees, clusterings and assignments."""
    from bamse_b200 import workloads
    from bamse_b200.clustering import analyse_clusterings
    w32 = workloads.build("c2", ctx=ctx)
    assert w32["n_evals_per_param"] == 397530
    t32 = analyse_clusterings(w32["clusts"], 5, ctx=ctx, precision=shim.FP32, solve_fractions=False)
    w64 = workloads.build("c2", ctx=ctx)
    t64 = analyse_clusterings(w64["clusts"], 5, ctx=ctx, precision=shim.FP64, solve_fractions=False)
    assert len(t32) == len(t64) == 5
    for a, b in zip(t32, t64):
        assert a["tree"] == b["tree"] and a["clustering_index"] == b["clustering_index"]
        assert a["assign"].tolist() == b["assign"].tolist()
        assert a["totalscore"] == pytest.approx(b["totalscore"], rel=1e-10)
    assert t64[0]["totalscore"] > t64[-1]["totalscore"]


def test_config4_shape_index_subranges(ctx, shim):
    """BASELINE config 4 shape (K = 10: 10^9 labeled rooted trees per clustering, M = 5, 1000
    mutations): 3 x 300 000-tree index sub-ranges.  Properties at that size: the exact early exit does
    not change the top-k; splitting the ranges in two and merging gives the same records; and on a
    20 000-tree sub-range the fp32 path returns what a complete fp64 check-mode pass returns."""
    from bamse_b200 import multigpu, workloads
    from bamse_b200.clustering import _pack_problem
    w = workloads.build("c4", ctx=ctx)
    clusts = w["clusts"]
    assert all(c.As.shape == (10, 5) for c in clusts)
    Ks, var, ref, Kmax = _pack_problem(clusts)
    ctx.upload_problem([0.001], Ks, var, ref)
    C = len(clusts)
    const = np.array([[c.cluster_prior + c.maxlikelihood for c in clusts]])
    begin = np.array([123456789 + 1000 * c for c in range(C)], dtype=np.uint64)
    n = 300000
    end = begin + np.uint64(n)
    plain, _, cnt = ctx.score_topk(const, None, begin, end, topk=8, precision=shim.FP32)
    assert cnt[0] == 8 and ctx.last_eval_count == C * n
    ctx.set_early_exit(True)
    ctx.samples_evaluated(reset=True)
    early, _, _ = ctx.score_topk(const, None, begin, end, topk=8, precision=shim.FP32)
    done = ctx.samples_evaluated(reset=True)
    ctx.set_early_exit(False)
    assert early["tree"].tolist() == plain["tree"].tolist() and early["clust"].tolist() == plain["clust"].tolist()
    np.testing.assert_array_equal(early["total"], plain["total"])
    assert done <= C * n * 5
    halves = []
    for lo, hi in ((begin, begin + np.uint64(n // 2)), (begin + np.uint64(n // 2), end)):
        halves.append(ctx.score_topk(const, None, lo, hi, topk=8, precision=shim.FP32)[0])
    merged = multigpu.merge_records(np.stack(halves), 8)
    assert merged["tree"].tolist() == plain["tree"].tolist()
    np.testing.assert_allclose(merged["total"], plain["total"], rtol=1e-12)
    small_end = begin + np.uint64(20000)
    f32, _, _ = ctx.score_topk(const, None, begin, small_end, topk=5, precision=shim.FP32)
    f64, _, _ = ctx.score_topk(const, None, begin, small_end, topk=5, precision=shim.FP64)
    assert f32["tree"].tolist() == f64["tree"].tolist() and f32["clust"].tolist() == f64["clust"].tolist()
    np.testing.assert_allclose(f32["total"], f64["total"], rtol=1e-10)


def test_program_tables_equal_in_kernel_decoding(ctx, shim, monkeypatch):
    """The fp32 enumeration fetches ready-made evaluation programs for K <= 8 instead of decoding the
    tree index in the kernel: both routes must give bit-identical scores, log scre and top-k, for whole
    small search spaces, a mixed-K problem and a K = 8 sub-range (K = 9 always decodes in the kernel)."""
    rng = np.random.default_rng(2024)

    def problem(Ks, M, cov):
        Kmax = max(Ks)
        As = np.zeros((len(Ks), Kmax, M), np.int64)
        Bs = np.zeros_like(As)
        for ci, K in enumerate(Ks):
            frac = np.sort(rng.dirichlet(np.ones(K + 1))[:K])[::-1]
            cm = np.clip(np.cumsum(frac[::-1])[::-1][:, None] * rng.uniform(0.6, 1.0, size=(K, M)), 0, 1)
            As[ci, :K] = rng.binomial(cov, cm * 0.497 + 0.003)
            Bs[ci, :K] = cov - As[ci, :K]
        return As, Bs

    for Ks, M, cov, ranges in [([3, 4, 5, 6, 2, 1], 3, 900, None), ([7], 2, 400, [(0, 0, 30000), (0, 7 ** 6 - 5000, 5000)]),
                               ([8, 5], 2, 2500, [(0, 8 ** 7 - 3000, 3000), (0, 12345, 2000), (1, 0, 625)])]:
        As, Bs = problem(Ks, M, cov)
        got = {}
        monkeypatch.setenv("BAMSE_FRAMED", "1" if Ks == [7] else "0")     # both kernel instantiations
        for mode in ("0", "1"):
            monkeypatch.setenv("BAMSE_PROGTAB", mode)
            ctx.upload_problem([0.003], Ks, As, Bs)
            rs = ranges or [(c, 0, K ** max(K - 1, 0)) for c, K in enumerate(Ks)]
            dense = [ctx.score_range(0, c, b, n, shim.FP32) for c, b, n in rs]
            top = ctx.score_topk(np.zeros((1, len(Ks))), topk=6, precision=shim.FP32)
            got[mode] = (dense, top)
        monkeypatch.delenv("BAMSE_PROGTAB")
        monkeypatch.delenv("BAMSE_FRAMED")
        for (s0, l0), (s1, l1) in zip(got["0"][0], got["1"][0]):
            assert np.array_equal(s0, s1) and np.array_equal(l0, l1)
        assert got["0"][1][0].tobytes() == got["1"][1][0].tobytes()
        assert np.array_equal(got["0"][1][1], got["1"][1][1])
