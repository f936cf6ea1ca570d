"""Pins oracle/bamse_oracle.py (numpy restatement) to golden vectors produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import math

import numpy as np
import pytest

import bamse_oracle as orc


def test_tree_scores_match_reference(golden_scores):
    assert len(golden_scores) >= 25
    for case in golden_scores:
        As = np.array(case["var"])
        Bs = np.array(case["ref"])
        K, M = As.shape
        D = np.array([[orc.distrib(As[k, m], Bs[k, m], case["err"]) for m in range(M)] for k in range(K)])
        s = orc.tree_score_samples(case["tree"], D[case["nodes"]], orc.prior_const(case["K_prior"]))
        np.testing.assert_allclose(s, case["per_sample"], rtol=1e-12, atol=0, err_msg=case["id"])
        assert abs(s.sum() - case["total"]) <= 1e-12 * abs(case["total"])


def test_survey_known_answers(golden_scores):
    """SURVEY.md section 8a quotes these sums to full precision; they must be what the
    reference produced when the golden file was generated."""
    quoted = {"G1": -145.76370519862147, "G2": -991.505659145346, "G3": -672.8709598555804,
              "G4": -11.991470855683604, "G5": -19.867300435107396, "G6": -81.22858030512461,
              "G7": -24.78865926010896, "G8": -328.295525392031}
    by_id = {c["id"]: c for c in golden_scores}
    for k, v in quoted.items():
        assert by_id[k]["total"] == pytest.approx(v, rel=1e-13)


def test_vector_operators(golden_vectors):
    for d in golden_vectors["distrib"]:
        y = orc.distrib(d["a"], d["b"], d["err"])
        np.testing.assert_allclose(y[d["idx"]], d["y"], rtol=1e-14)
    for c, t in zip(golden_vectors["convolve"], golden_vectors["trap_int"]):
        u = orc.distrib(c["a1"], c["b1"], c["err"])
        v = orc.distrib(c["a2"], c["b2"], c["err"])
        w = orc.log_convolve(u, v)
        np.testing.assert_allclose(w[c["idx"]], c["y"], rtol=1e-13)
        np.testing.assert_allclose(orc.log_cumsum(w)[t["idx"]], t["y"], rtol=1e-13)


def test_leaf_convolution_is_a_scan():
    """my_convolve(D, flat prior) == logaddexp.accumulate(D) + p0 - log L."""
    D = orc.distrib(300, 1700, 0.003)
    p0 = orc.prior_const(5)
    a = orc.log_convolve(D, np.full(512, p0))
    b = np.logaddexp.accumulate(D) + p0 - math.log(512)
    np.testing.assert_allclose(a, b, rtol=1e-13)


def test_canon_weight(golden_canon):
    assert golden_canon["alltrees_counts"][1:] == [1, 1, 2, 4, 9, 20, 48, 115, 286, 719]
    for K, items in golden_canon["alltrees_scre"].items():
        for it in items:
            s, w = orc.canon_weight(it["tree"])
            assert (s, w) == (it["canstr"], it["scre"])
    for it in golden_canon["random"]:
        s, w = orc.canon_weight(it["tree"])
        assert (s, w) == (it["canstr"], it["scre"]), it["tree"]
    # quirk examples quoted in SURVEY 8a row 9
    assert orc.canon_weight([-1, 0, 0, 0])[1] == 3
    assert orc.canon_weight([-1, 0, 0, 0, 1])[1] == 2
    assert orc.canon_weight([-1, 0, 0, 0, 2])[1] == 1
    assert orc.canon_weight([-1, 0, 0, 1, 1, 2, 2])[1] == 8


def test_unlabeled_enumerator_matches_alltrees(golden_canon):
    for K in range(1, 11):
        mine = sorted(orc.canon_weight(t)[0] for t in orc.unlabeled_rooted_trees(K))
        assert mine == golden_canon["alltrees_canstr"][str(K)]


def test_constants(golden_constants):
    for p in golden_constants["part_func"]:
        np.testing.assert_array_equal(orc.part_func(p["n"], p["kmax"]), p["pf"])
    for p in golden_constants["penalty"]:
        pf = orc.part_func(p["n"], p["kmax"])
        assert orc.log_conf_penalty(np.array(p["assign"]), pf) == pytest.approx(p["value"], rel=1e-14)
    for c in golden_constants["clustering"]:
        o = orc.OracleClustering(c["am"], c["bm"], c["assign"], np.array(c["pf"]), c["err"])
        np.testing.assert_array_equal(o.As, c["As"])
        np.testing.assert_array_equal(o.Bs, c["Bs"])
        assert o.maxlikelihood == pytest.approx(c["maxlikelihood"], rel=1e-13)
        assert o.cluster_prior == pytest.approx(c["cluster_prior"], rel=1e-13)
    # SURVEY 8a known answers
    np.testing.assert_array_equal(orc.part_func(80, 5), [1, 40, 533, 3689, 16019])
    np.testing.assert_array_equal(orc.part_func(50, 5), [1, 25, 208, 920, 2611])
    assert orc.log_conf_penalty(np.array([0, 0, 1, 1, 1, 2, 3, 3, 3, 3]), orc.part_func(10, 4)) == \
        pytest.approx(-14.41126539251557, rel=1e-14)


@pytest.mark.parametrize("K", [1, 2, 3, 4, 5, 6])
def test_tree_index_bijection(K):
    """Cayley: the index decoder hits each of the K**(K-1) labeled rooted trees once
    (reference test.py:24-25 known answer)."""
    seen = set()
    for i in range(orc.n_trees(K)):
        p = orc.decode_tree(K, i)
        assert p.count(-1) == 1 and len(p) == K
        # acyclic / connected: every node reaches the root
        for v in range(K):
            steps, w = 0, v
            while p[w] != -1:
                w = p[w]
                steps += 1
                assert steps <= K
        assert orc.encode_tree(p) == i
        seen.add(tuple(p))
    assert len(seen) == K ** (K - 1)


def test_symmetry_class_sizes(golden_canon):
    """sum over unlabeled topologies of (#labelings) = K^(K-1)."""
    for K in range(1, 7):
        from collections import Counter
        cnt = Counter(orc.canon_weight(orc.decode_tree(K, i))[0] for i in range(orc.n_trees(K)))
        assert sorted(cnt) == golden_canon["alltrees_canstr"][str(K)]
        assert sum(cnt.values()) == K ** (K - 1)


def test_search_path(golden_search):
    """reorder / candidate / threshold / kept set against the reference's own
    clustering.reorder, get_candidate_tree and get_trees2."""
    for case in golden_search:
        o = orc.OracleClustering(case["am"], case["bm"], case["assign_in"], np.array(case["pf"]), case["spec"]["err"])
        o.reorder()
        assert [int(x) for x in o.order] == case["order"]
        assert o.assign.tolist() == case["assign_out"]
        np.testing.assert_array_equal(o.As, case["As"])
        cand = o.get_candidate_tree()
        assert cand == case["candidate"]
        thr = o.threshold(cand)
        assert thr == pytest.approx(case["threshold"], rel=1e-12)
        # semantic target of get_trees2: exhaustive set with score + log scre >= threshold.
        K = o.K
        full = {}
        for i in range(orc.n_trees(K)):
            t = orc.decode_tree(K, i)
            full[tuple(t)] = o.full_score(t)
        kept_exh = {t for t, s in full.items() if s >= thr - 1e-9 * abs(thr)}
        kept_ref = {tuple(k["tree"]) for k in case["kept"]}
        assert kept_ref <= kept_exh                      # reference-kept is a subset (its B&B bound is heuristic)
        for k in case["kept"]:
            assert full[tuple(k["tree"])] == pytest.approx(k["score"] + math.log(k["scre"]), rel=1e-12)
            assert o.total_score(k["tree"]) == pytest.approx(k["totalscore"], rel=1e-12)
        # the best tree overall is always among the reference-kept ones
        best = max(full, key=full.get)
        assert best in kept_ref


def test_tree_scores_highk_match_reference():
    """round 2: reference-produced scores at K = 8 ... 12 (tests/golden/tree_scores_highk.json)."""
    from conftest import load_golden
    cases = load_golden("tree_scores_highk.json")
    assert len(cases) >= 30
    for case in cases:
        As = np.array(case["var"])
        Bs = np.array(case["ref"])
        K, M = As.shape
        D = np.array([[orc.distrib(As[k, m], Bs[k, m], case["err"]) for m in range(M)] for k in range(K)])
        s = orc.tree_score_samples(case["tree"], D[case["nodes"]], orc.prior_const(case["K_prior"]))
        np.testing.assert_allclose(s, case["per_sample"], rtol=1e-12, atol=0, err_msg=case["id"])
        assert orc.canon_weight(case["tree"])[1] == case["scre"]
        assert orc.decode_tree(K, orc.encode_tree(case["tree"])) == case["tree"]


def test_search_path_on_bamse_dat_and_k5():
    """round 2: reorder / candidate / threshold / kept set on the reference's sample input bamse.dat
    (KMeans-4 and KMeans-5 clusterings) and on a K = 5 simdata input (tests/golden/search_big.json)."""
    from conftest import load_golden
    for case in load_golden("search_big.json"):
        o = orc.OracleClustering(case["am"], case["bm"], case["assign_in"], np.array(case["pf"]), case["spec"]["err"])
        o.reorder()
        assert [int(x) for x in o.order] == case["order"]
        assert o.assign.tolist() == case["assign_out"]
        np.testing.assert_array_equal(o.As, case["As"])
        cand = o.get_candidate_tree()
        assert cand == case["candidate"]
        thr = o.threshold(cand)
        assert thr == pytest.approx(case["threshold"], rel=1e-12)
        for k in case["kept"]:
            assert o.full_score(k["tree"]) == pytest.approx(k["score"] + math.log(k["scre"]), rel=1e-12)
            assert o.full_score(k["tree"]) >= thr - 1e-9 * abs(thr)
            assert o.total_score(k["tree"]) == pytest.approx(k["totalscore"], rel=1e-12)


def test_tree_scores_at_the_edges_of_the_domain_match_reference():
    """round 2: reference-produced scores for the inputs of tests/test_gpu_edges.py (tests/golden/
    tree_scores_edges.json): zero variant / reference / total reads, e = 1e-6 and 0.499, M = 1 and 64, K = 1, 2,
    all counts zero.  reference == oracle here, oracle == CUDA there."""
    from conftest import load_golden
    cases = load_golden("tree_scores_edges.json")
    assert len(cases) >= 120
    seen = set()
    for case in cases:
        seen.add(case["id"].split("-")[0])
        As = np.array(case["var"])
        Bs = np.array(case["ref"])
        K, M = As.shape
        assert np.all(np.isfinite(case["per_sample"]))
        D = np.array([[orc.distrib(As[k, m], Bs[k, m], case["err"]) for m in range(M)] for k in range(K)])
        s = orc.tree_score_samples(case["tree"], D, orc.prior_const(K))
        np.testing.assert_allclose(s, case["per_sample"], rtol=1e-12, atol=1e-12, err_msg=case["id"])
        assert orc.canon_weight(case["tree"])[1] == case["scre"]
        assert orc.decode_tree(K, case["tree_index"]) == case["tree"]
    assert {"zero", "err1e", "err0.499", "M1", "M64", "K2", "K1", "allzero"} <= seen


def test_tree_scores_at_the_shapes_of_the_baseline_configs_match_reference():
    """round 2: reference-produced scores on the aggregated tables of the largest clustering of each bench.py
    workload c1 .. c5 (K = 4 / 6 / 7 / 10 / 5, M = 3 / 10 / 20 / 5 / 10; c5 at both ends of its error-rate sweep):
    chain, star and three random trees each (tests/golden/tree_scores_configs.json)."""
    from conftest import load_golden
    cases = load_golden("tree_scores_configs.json")
    assert {c["config"] for c in cases} == {"c1", "c2", "c3", "c4", "c5"} and len(cases) == 30
    shapes = {c["config"]: np.array(c["var"]).shape for c in cases}
    assert shapes == {"c1": (4, 3), "c2": (6, 10), "c3": (7, 20), "c4": (10, 5), "c5": (5, 10)}
    tables = {}
    for case in cases:
        As, Bs = np.array(case["var"]), np.array(case["ref"])
        K, M = As.shape
        key = (case["config"], case["err"])
        if key not in tables:
            tables[key] = np.array([[orc.distrib(As[k, m], Bs[k, m], case["err"]) for m in range(M)] for k in range(K)])
        s = orc.tree_score_samples(case["tree"], tables[key], orc.prior_const(K))
        np.testing.assert_allclose(s, case["per_sample"], rtol=1e-12, atol=0, err_msg=case["id"])
        assert float(np.sum(s)) == pytest.approx(case["total"], rel=1e-13)
        assert orc.canon_weight(case["tree"])[1] == case["scre"]
