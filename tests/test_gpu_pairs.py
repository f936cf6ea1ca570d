"""The inside x outside pair path (bamse_b200/csrc/pairs.cuh) on the GPU, through the C ABI: against the oracle,
against the table-driven programs (BAMSE_PAIRS=0 builds the same problem without context tables) and against the
fp64 check mode, incl. cluster counts whose irregular trees take the flat-list route (K >= 8)."""
import math
import os

import numpy as np
import pytest

import bamse_oracle as orc

pytestmark = pytest.mark.gpu

FP32_RTOL = 1e-6


@pytest.fixture(scope="module")
def shim():
    from bamse_b200 import cuda_shim
    return cuda_shim


@pytest.fixture()
def ctx(shim):
    c = shim.Context(0)
    yield c
    c.close()


def _problem(rng, K, M, cov):
    frac = np.sort(rng.dirichlet(np.ones(K + 1))[:K])[::-1]
    n = rng.integers(20, 120, size=K)
    var = np.zeros((K, M), dtype=np.int64)
    ref = np.zeros((K, M), dtype=np.int64)
    for k in range(K):
        for m in range(M):
            tot = int(n[k]) * cov
            v = rng.binomial(tot, min(0.49, frac[k] * rng.uniform(0.6, 1.2) * 0.5 + 0.003))
            var[k, m], ref[k, m] = v, tot - v
    return var, ref


class programs_only(object):
    """uploads made inside run without the pair path (the library reads BAMSE_PAIRS at upload time)."""

    def __enter__(self):
        self.old = os.environ.get("BAMSE_PAIRS")
        os.environ["BAMSE_PAIRS"] = "0"

    def __exit__(self, *a):
        if self.old is None:
            del os.environ["BAMSE_PAIRS"]
        else:
            os.environ["BAMSE_PAIRS"] = self.old


@pytest.mark.parametrize("K,M,cov", [(2, 2, 50), (3, 3, 200), (4, 2, 1000), (5, 3, 100), (6, 2, 300)])
def test_all_trees_pairs_vs_oracle_and_programs(ctx, shim, K, M, cov):
    """every tree of a clustering: pair path == oracle (K <= 5) == programs-only build == fp64 check mode."""
    rng = np.random.default_rng(100 * K + M)
    var, ref = _problem(rng, K, M, cov)
    T = K ** (K - 1)
    ctx.upload_problem([0.003], [K], var[None], ref[None])
    s_pair, w_pair = ctx.score_range(0, 0, 0, T, shim.FP32)
    s64, w64 = ctx.score_range(0, 0, 0, T, shim.FP64)
    with programs_only():
        ctx.upload_problem([0.003], [K], var[None], ref[None])
        s_prog, w_prog = ctx.score_range(0, 0, 0, T, shim.FP32)
    np.testing.assert_array_equal(w_pair, w64)
    np.testing.assert_array_equal(w_prog, w64)
    assert np.max(np.abs(s_pair - s64) / np.abs(s64)) <= FP32_RTOL
    assert np.max(np.abs(s_prog - s64) / np.abs(s64)) <= FP32_RTOL
    if K <= 5:
        D = np.array([[orc.distrib(var[k, m], ref[k, m], 0.003) for m in range(M)] for k in range(K)])
        for i in range(0, T, max(1, T // 60)):
            want = orc.tree_score_samples(orc.decode_tree(K, i), D, orc.prior_const(K)).sum()
            assert abs(s_pair[i] - want) <= FP32_RTOL * abs(want)


@pytest.mark.parametrize("K,M,cov", [(7, 3, 200), (8, 2, 400), (9, 2, 100), (10, 2, 2000)])
def test_random_ranges_pairs_vs_check_mode(ctx, shim, K, M, cov):
    """runs of consecutive indices all over the index space (regular and irregular trees mixed): pair path +
    flat list == fp64 check mode within 1e-6, canonizeF weights exact."""
    rng = np.random.default_rng(7 * K)
    var, ref = _problem(rng, K, M, cov)
    ctx.upload_problem([0.003], [K], var[None], ref[None])
    T = K ** (K - 1)
    run = 700
    for b in [0, T - run] + [int(x) for x in rng.integers(0, T - run, size=3)]:
        s32, w32 = ctx.score_range(0, 0, b, run, shim.FP32)
        s64, w64 = ctx.score_range(0, 0, b, run, shim.FP64)
        np.testing.assert_array_equal(w32, w64)
        assert np.max(np.abs(s32 - s64) / np.abs(s64)) <= FP32_RTOL, (K, b)


def test_topk_pairs_equals_programs_and_check_mode(ctx, shim):
    """a mixed problem (K = 3 ... 8, two parameter sets, thresholds): bamse_score_topk returns the same records
    with the pair path, without it and in the check mode; early exit does not change them."""
    rng = np.random.default_rng(5)
    Ks = [3, 5, 6, 7, 8, 4]
    M, Kmax = 3, 8
    var = np.zeros((len(Ks), Kmax, M), dtype=np.int64)
    ref = np.zeros((len(Ks), Kmax, M), dtype=np.int64)
    for c, K in enumerate(Ks):
        v, r = _problem(rng, K, M, 150)
        var[c, :K], ref[c, :K] = v, r
    errs = [0.001, 0.01]
    const = rng.uniform(-5, 5, size=(2, len(Ks)))
    out = {}
    for mode in ("pairs", "programs", "check", "pairs_early"):
        if mode == "programs":
            with programs_only():
                ctx.upload_problem(errs, Ks, var, ref)
        else:
            ctx.upload_problem(errs, Ks, var, ref)
        ctx.set_early_exit(mode == "pairs_early")
        rec, par, cnt = ctx.score_topk(const, None, None, None, topk=6, precision=shim.FP64 if mode == "check" else shim.FP32)
        ctx.set_early_exit(False)
        assert ctx.last_topk_verified
        out[mode] = rec
    for mode in ("programs", "check", "pairs_early"):
        np.testing.assert_array_equal(out[mode]["tree"], out["pairs"]["tree"])
        np.testing.assert_array_equal(out[mode]["clust"], out["pairs"]["clust"])
        np.testing.assert_allclose(out[mode]["total"], out["pairs"]["total"], rtol=1e-12)


def test_rebuild_tables_is_idempotent(ctx, shim):
    """bamse_rebuild_tables re-derives the same tables from the resident counts."""
    rng = np.random.default_rng(11)
    var, ref = _problem(rng, 6, 3, 200)
    ctx.upload_problem([0.003], [6], var[None], ref[None])
    a, _ = ctx.score_range(0, 0, 0, 6 ** 5, shim.FP32)
    ctx.rebuild_tables()
    b, _ = ctx.score_range(0, 0, 0, 6 ** 5, shim.FP32)
    np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_shard_plan_ranks_together_equal_one_gpu(shim, world):
    """the sharding plan on ONE device: `world` contexts, each planned as rank r of `world` (whole clusterings per
    rank, the most expensive ones split by interleaved index), score disjoint shares whose union is every tree;
    the merged top-k equals the unsharded one, record for record, and every rank's eval counts add up."""
    from bamse_b200.multigpu import merge_records
    rng = np.random.default_rng(77)
    Ks = [6, 6, 6, 5, 5, 4, 7, 3, 6, 5, 2]
    M, Kmax = 3, 7
    var = np.zeros((len(Ks), Kmax, M), dtype=np.int64)
    ref = np.zeros((len(Ks), Kmax, M), dtype=np.int64)
    for c, K in enumerate(Ks):
        v, r = _problem(rng, K, M, 120)
        var[c, :K], ref[c, :K] = v, r
    errs = [0.002, 0.02]
    const = rng.uniform(-3, 3, size=(2, len(Ks)))
    total = 2 * sum(K ** (K - 1) for K in Ks)
    with shim.Context(0) as one:
        one.upload_problem(errs, Ks, var, ref)
        want, _, _ = one.score_topk(const, None, None, None, topk=7, precision=shim.FP32)
        assert one.last_eval_count == total
    parts, evals, owners = [], 0, []
    for r in range(world):
        with shim.Context(0) as c:
            c.set_shard(r, world)
            c.upload_problem(errs, Ks, var, ref)
            owners.append(c.shard_plan())
            rec, _, _ = c.score_topk(const, None, None, None, topk=7, precision=shim.FP32)
            evals += c.last_eval_count
            parts.append(rec)
    for o in owners[1:]:
        np.testing.assert_array_equal(o, owners[0])          # every rank computes the same plan
    assert (owners[0] >= 0).any()                             # some whole (p, c) pairs are owned ...
    assert set(np.unique(owners[0])) <= set(range(-1, world))
    assert evals == total
    got = merge_records(np.stack(parts), 7)
    np.testing.assert_array_equal(got["tree"], want["tree"])
    np.testing.assert_array_equal(got["clust"], want["clust"])
    np.testing.assert_allclose(got["total"], want["total"], rtol=1e-12)


def test_changing_the_shard_of_a_planned_problem_needs_a_new_upload(ctx, shim):
    rng = np.random.default_rng(3)
    Ks = [5, 5, 4, 5]
    var = np.zeros((4, 5, 2), dtype=np.int64)
    ref = np.zeros((4, 5, 2), dtype=np.int64)
    for c, K in enumerate(Ks):
        v, r = _problem(rng, K, 2, 100)
        var[c, :K], ref[c, :K] = v, r
    ctx.set_shard(1, 2)
    ctx.upload_problem([0.003], Ks, var, ref)
    ctx.score_topk(np.zeros((1, 4)), topk=3)
    ctx.set_shard(0, 1)
    with pytest.raises(shim.BamseCudaError):
        ctx.score_topk(np.zeros((1, 4)), topk=3)
    ctx.upload_problem([0.003], Ks, var, ref)
    ctx.score_topk(np.zeros((1, 4)), topk=3)


def _run_cli(tmp_path, monkeypatch, sub, extra):
    import bamse_b200.clustering as C
    from bamse_b200 import cli, simdata
    d = simdata.simulate(4, 3, 60, seed=33)
    out = tmp_path / sub
    out.mkdir()
    inp = out / "in.dat"
    simdata.write_input_file(str(inp), d["As"], d["Bs"])
    monkeypatch.chdir(out)
    monkeypatch.setattr(C, "_default_ctx", None)
    assert cli.main(["-e", "0.003", "-nc", "6", "-n", "4", "-t", "3", "--seed", "9"] + extra + [str(inp)]) == 0
    names = ["bamse_output.txt"] + ["tree_number_%d%s" % (t, sfx) for t in range(3) for sfx in ("_subclones.dot", "_samples.dot", ".txt")]
    return {n: (out / n).read_bytes() for n in names if (out / n).exists()}


def test_cli_multi_gpu_output_is_byte_identical(tmp_path, monkeypatch, shim):
    """`bamse.py --gpus 2` (one process, one context per GPU, sharding plan, host merge) writes byte-identical
    output files to the single-GPU run.  On a one-GPU box the second context sits on the same device."""
    import torch
    from bamse_b200 import multigpu
    one = _run_cli(tmp_path, monkeypatch, "one", [])
    assert "bamse_output.txt" in one and len(one) >= 4
    if torch.cuda.device_count() >= 2:
        two = _run_cli(tmp_path, monkeypatch, "two", ["--gpus", "2"])
    else:
        real = multigpu.MultiContext
        monkeypatch.setattr(multigpu, "MultiContext", lambda devs: real([0 for _ in devs]))
        two = _run_cli(tmp_path, monkeypatch, "two", ["--gpus", "2"])
    assert one.keys() == two.keys()
    for n in one:
        assert one[n] == two[n], n


def test_trivial_cluster_counts_in_a_mixed_problem(ctx, shim):
    """K = 1 (one tree, no pair path) and K = 2, 3 next to K = 5: every tree of every clustering is ranked once,
    scores equal the check mode."""
    rng = np.random.default_rng(21)
    Ks = [1, 2, 3, 5, 1]
    M, Kmax = 2, 5
    var = np.zeros((len(Ks), Kmax, M), dtype=np.int64)
    ref = np.zeros((len(Ks), Kmax, M), dtype=np.int64)
    for c, K in enumerate(Ks):
        v, r = _problem(rng, K, M, 80)
        var[c, :K], ref[c, :K] = v, r
    ctx.upload_problem([0.003], Ks, var, ref)
    total = sum(K ** (K - 1) for K in Ks)
    assert total <= shim.TOPK_MAX * 20
    rec32, _, cnt32 = ctx.score_topk(np.zeros((1, len(Ks))), topk=64, precision=shim.FP32)
    rec64, _, cnt64 = ctx.score_topk(np.zeros((1, len(Ks))), topk=64, precision=shim.FP64)
    assert cnt32[0] == cnt64[0] == 64
    np.testing.assert_array_equal(rec32["tree"], rec64["tree"])
    np.testing.assert_array_equal(rec32["clust"], rec64["clust"])
    for c, K in enumerate(Ks):
        s32, _ = ctx.score_range(0, c, 0, K ** (K - 1), shim.FP32)
        s64, _ = ctx.score_range(0, c, 0, K ** (K - 1), shim.FP64)
        assert np.max(np.abs(s32 - s64) / np.abs(s64)) <= FP32_RTOL


@pytest.mark.parametrize("K,world", [(8, 3), (9, 2), (7, 8)])
def test_one_big_clustering_is_split_by_node_sets(shim, K, world):
    """a problem that is ONE clustering: the plan splits it over all ranks; the pair path shards by the pack's node
    set (every rank builds only its share of the top-level table rows), the irregular trees by interleaved index.
    Shares add up to every tree exactly once and the merged top-k equals the unsharded one."""
    from bamse_b200.multigpu import merge_records
    rng = np.random.default_rng(1000 + K)
    var, ref = _problem(rng, K, 2, 150)
    const = np.zeros((1, 1))
    T = K ** (K - 1)
    with shim.Context(0) as one:
        one.upload_problem([0.003], [K], var[None], ref[None])
        want, _, _ = one.score_topk(const, topk=9, precision=shim.FP32)
        assert one.last_eval_count == T
    parts, evals = [], 0
    for r in range(world):
        with shim.Context(0) as c:
            c.set_shard(r, world)
            c.upload_problem([0.003], [K], var[None], ref[None])
            assert c.shard_plan().tolist() == [[-1]]
            rec, _, _ = c.score_topk(const, topk=9, precision=shim.FP32)
            evals += c.last_eval_count
            parts.append(rec)
            with pytest.raises(shim.BamseCudaError):
                c.score_range(0, 0, 0, 10, shim.FP32)      # dense fp32 ranges need all tables on one rank
    assert evals == T
    got = merge_records(np.stack(parts), 9)
    np.testing.assert_array_equal(got["tree"], want["tree"])
    np.testing.assert_allclose(got["total"], want["total"], rtol=1e-12)
