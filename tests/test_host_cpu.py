"""CPU-only tests of the host side: the C ABI loads and exports what include/*.h declares,
argument/error behaviour without a GPU, tree utilities and clustering constants against the
reference-generated golden vectors, writers, CLI parsing."""
import ctypes
import math
import os
import re
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, has_cuda

from bamse_b200 import cuda_shim, tree_tools
from bamse_b200 import clustering as C


def header_functions():
    text = open(os.path.join(ROOT, "include", "bamse_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bamse_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_library):
    lib = cuda_shim.load_library()
    declared = header_functions()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), "libbamse_cuda.so does not export %s" % name
    assert sorted(cuda_shim.EXPORTED_SYMBOLS) == declared
    assert lib.bamse_abi_version() == 1


def test_record_layout_matches_header():
    assert cuda_shim.RECORD_DTYPE.itemsize == 32
    assert [cuda_shim.RECORD_DTYPE.fields[n][1] for n in ("total", "score", "clust", "param", "tree")] == [0, 8, 16, 20, 24]


@pytest.mark.skipif(has_cuda(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly_no_cpu_fallback(built_library):
    with pytest.raises(cuda_shim.BamseCudaError) as e:
        cuda_shim.Context(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    c = C.clustering(np.ones((4, 2), dtype=int), np.ones((4, 2), dtype=int), [0, 0, 1, 1], C.part_func(4, 2), 0.01)
    with pytest.raises(cuda_shim.BamseCudaError):
        c.get_tree_score([-1, 0], [0, 1])


def test_host_decoder_is_a_bijection(built_library):
    import bamse_oracle as orc
    for K in range(1, 7):
        seen = set()
        for i in range(K ** (K - 1)):
            p = cuda_shim.decode_tree(K, i)
            assert p == orc.decode_tree(K, i)
            seen.add(tuple(p))
        assert len(seen) == K ** (K - 1)
    lib = cuda_shim.load_library()
    buf = np.zeros(12, dtype=np.int8)
    assert lib.bamse_decode_tree(3, 9, buf.ctypes.data_as(ctypes.c_void_p)) == -1      # index out of range
    assert lib.bamse_decode_tree(13, 0, buf.ctypes.data_as(ctypes.c_void_p)) == -1     # K > KMAX
    assert cuda_shim.decode_tree(12, 12 ** 11 - 1).count(-1) == 1                      # 64-bit indices


def test_tree_tools_against_reference_golden(golden_canon):
    for K, items in golden_canon["alltrees_scre"].items():
        for it in items:
            r = tree_tools.canonizeF(it["tree"])
            assert (r["canstr"], r["scre"]) == (it["canstr"], it["scre"])
    for it in golden_canon["random"]:
        r = tree_tools.canonizeF(it["tree"])
        assert (r["canstr"], r["scre"]) == (it["canstr"], it["scre"])
    for K in range(1, 11):
        mine = sorted(tree_tools.canonizeF(t)["canstr"] for t in tree_tools.unlabeled_rooted_trees(K))
        assert mine == golden_canon["alltrees_canstr"][str(K)]


def test_tree_tools_small_known_answers():
    B, anc = tree_tools.get_matrix([-1, 0, 0, 1])
    np.testing.assert_array_equal(B, [[1, 1, 1, 1], [0, 1, 0, 1], [0, 0, 1, 0], [0, 0, 0, 1]])
    assert anc[3] == [3, 1, 0]
    assert tree_tools.get_childs([2, 2, -1, 0, 0], 2) == [0, 1]
    assert tree_tools.get_tree_depth([2, 2, -1, 0, 0]) == [2, 2, 1, 3, 3]
    assert tree_tools.reorder_tree([-1, 0, 1], [5, 3, 4], [3, 4, 5]) == [2, 0, -1]
    assert list(tree_tools.reorder_assignmens([2, 2, 0, 1, 0])) == [0, 0, 1, 2, 1]
    assert list(tree_tools.collapse_node(np.array([-1, 0, 1, 2]), 1)) == [-1, 0, 1]
    assert list(tree_tools.collapse_assign([0, 1, 2, 3, 1], 1, 2)) == [0, 1, 1, 2, 1]


def test_remove_zeros_collapses_empty_single_child_nodes():
    t = {'tree': np.array([-1, 0, 1, 1]), 'assign': np.array([0, 1, 2, 3, 1]),
         'clone_proportions': np.array([[0.3, 0.2], [0.0, 0.0004], [0.4, 0.1], [0.1, 0.2]])}
    out = tree_tools.remove_zeros(t)      # node 1 is empty but has two children: kept
    assert list(out['tree']) == [-1, 0, 1, 1]
    t = {'tree': np.array([-1, 0, 1, 2]), 'assign': np.array([0, 1, 2, 3, 1]),
         'clone_proportions': np.array([[0.3, 0.2], [0.0, 0.0004], [0.4, 0.1], [0.1, 0.2]])}
    out = tree_tools.remove_zeros(t)
    assert list(out['tree']) == [-1, 0, 1]
    assert list(out['assign']) == [0, 1, 1, 2, 1]
    assert out['clone_proportions'].shape == (3, 2) and out['Bmatrix'].shape == (3, 3)


def test_clustering_constants_against_reference_golden(golden_constants):
    for p in golden_constants["part_func"]:
        np.testing.assert_array_equal(C.part_func(p["n"], p["kmax"]), p["pf"])
    for p in golden_constants["penalty"]:
        assert C.log_conf_penalty(np.array(p["assign"]), C.part_func(p["n"], p["kmax"])) == pytest.approx(p["value"], rel=1e-14)
    for c in golden_constants["clustering"]:
        o = C.clustering(c["am"], c["bm"], c["assign"], np.array(c["pf"]), c["err"])
        np.testing.assert_array_equal(o.As, c["As"])
        np.testing.assert_array_equal(o.Bs, c["Bs"])
        assert o.maxlikelihood == pytest.approx(c["maxlikelihood"], rel=1e-13)
        assert o.cluster_prior == pytest.approx(c["cluster_prior"], rel=1e-13)
    assert list(C.tree_numbers[:10]) == [1, 1, 1, 2, 4, 9, 20, 47, 108, 252]


def test_writers_are_byte_compatible_with_the_reference():
    """tests/golden/out_* were written by the reference's own tree_io functions."""
    import json
    from bamse_b200 import tree_io
    spec = json.load(open(os.path.join(GOLDEN, "out_spec.json")))
    tree = dict(tree=np.array(spec["tree"]), root=spec["root"], assign=np.array(spec["assign"]),
                vaf=np.array(spec["vaf"]), clone_proportions=np.array(spec["clone_proportions"]),
                totalscore=spec["totalscore"], sample_names=spec["sample_names"])
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        a, b, c = (os.path.join(d, n) for n in ("t.dot", "s.dot", "o.txt"))
        tree_io.write_dot_files(tree, spec["sample_colors"], spec["cluster_colors"], a, b)
        tree_io.write_text_output([tree, tree], c)
        assert open(a).read() == open(os.path.join(GOLDEN, "out_subclones.dot")).read()
        assert open(b).read() == open(os.path.join(GOLDEN, "out_samples.dot")).read()
        assert open(c).read() == open(os.path.join(GOLDEN, "out_text.txt")).read()


def test_ascii_tree_shape():
    from bamse_b200.tree_io import get_ascii_tree
    assert get_ascii_tree([-1, 0, 1, 0], ["A", "B", "C", "D"]) == "A\n +-- B\n |   +-- C\n +-- D"


def test_cli_keeps_the_reference_options():
    from bamse_b200.cli import build_parser
    a = build_parser().parse_args(["x.dat"])
    assert (a.e, a.nc, a.s, a.n, a.t, a.infile) == (0.001, 12, -1, 3, 1, ["x.dat"])
    a = build_parser().parse_args(["-e", "0.01", "-nc", "5", "-s", "0.2", "-n", "10", "-t", "5", "a", "b"])
    assert (a.e, a.nc, a.s, a.n, a.t, a.infile[0]) == (0.01, 5, 0.2, 10, 5, "a")


def test_input_reader_matches_bamse_dat_layout(tmp_path):
    from bamse_b200.cli import read_input
    from bamse_b200 import simdata
    d = simdata.simulate(3, 2, 20, seed=1)
    p = tmp_path / "in.dat"
    simdata.write_input_file(str(p), d["As"], d["Bs"])
    As, Bs, names = read_input(str(p))
    assert names == ["S1", "S2"]
    np.testing.assert_array_equal(As, d["As"])
    np.testing.assert_array_equal(Bs, d["Bs"])


def test_simdata_shapes_and_seed():
    from bamse_b200 import simdata
    a = simdata.simulate(4, 3, 50, seed=12345)
    b = simdata.simulate(4, 3, 50, seed=12345)
    assert a["As"].shape == (50, 3) and (a["As"] + a["Bs"] == 200).all()
    np.testing.assert_array_equal(a["As"], b["As"])
    assert sorted(set(a["assign"])) == [0, 1, 2, 3] and np.bincount(a["assign"]).min() >= 2


def test_fraction_solver_recovers_planted_fractions():
    from bamse_b200.fractions import solve4
    B = tree_tools.get_matrix([-1, 0, 0])[0]
    x_true = np.array([[0.2, 0.5], [0.3, 0.1], [0.4, 0.3]])
    prev = B.dot(x_true)
    cov, er = 200000, 0.001
    As = np.rint(cov * (er + prev * (0.5 - er))).astype(int)
    res = solve4(B, As, cov - As, er, [np.ones(3)] * 2)
    np.testing.assert_allclose(res['s'], x_true, atol=2e-3)


def _barrier_newton(B, a, b, er):
    """independent solve of solve4's concave program for one sample: log-barrier Newton in prevalence space
    y = B x (the likelihood is separable there), constraints A y >= 0, 0 <= y <= 1 with A = B^-1."""
    K = len(a)
    A = np.linalg.inv(B)
    h = 0.5 - er
    scale = float(np.sum(a + b))

    def f(y):
        q = er + y * h
        return -(np.sum(a * np.log(q)) + np.sum(b * np.log1p(-q))) / scale

    y = B.dot(np.full(K, 0.5 / K))                      # strictly feasible: x > 0, sum x < 1
    mu = 1e-2
    for _ in range(60):
        for _ in range(50):
            q = er + y * h
            x = A.dot(y)
            g = -(a / q - b / (1 - q)) * h / scale - mu * (A.T.dot(1 / x) + 1 / y - 1 / (1 - y))
            H = np.diag((a / q ** 2 + b / (1 - q) ** 2) * h * h / scale + mu * (1 / y ** 2 + 1 / (1 - y) ** 2)) \
                + mu * A.T.dot(np.diag(1 / x ** 2)).dot(A)
            step = -np.linalg.solve(H, g)
            t = 1.0
            phi = lambda z: f(z) - mu * (np.sum(np.log(A.dot(z))) + np.sum(np.log(z)) + np.sum(np.log(1 - z)))
            while True:
                z = y + t * step
                if (z > 0).all() and (z < 1).all() and (A.dot(z) > 0).all() and phi(z) <= phi(y) + 1e-4 * t * g.dot(step):
                    break
                t *= 0.5
                if t < 1e-12:
                    break
            y = y + t * step
            if np.abs(g.dot(step)) < 1e-18:
                break
        mu *= 0.3
        if mu < 1e-16:
            break
    return A.dot(y), f(y) * scale


def test_fraction_solver_against_an_independent_solve():
    """the SLSQP solve of solve4's concave program (reference clustering.py:410-426; cvxpy is absent
    here, so its digits are unpinned) against a second, independent solver (a log-barrier Newton method written
    here) on 40 random trees x counts incl. the bamse.dat KMeans-4 tables: every solution is feasible, objective
    values agree to 1e-9 relative and fractions to 2e-4 absolute.  That is the tolerance INTEGRATION.md states for
    `clone_proportions` / `vaf`."""
    import bamse_oracle as orc
    from bamse_b200.fractions import solve4
    rng = np.random.default_rng(8)
    cases = []
    for _ in range(39):
        K = int(rng.integers(2, 8))
        tree = orc.decode_tree(K, int(rng.integers(0, K ** (K - 1))))
        cov = int(rng.choice([60, 200, 5000]))
        x = rng.dirichlet(np.ones(K + 1))[:K]
        B = tree_tools.get_matrix(tree)[0]
        prev = np.clip(B.dot(x), 0, 1)
        n = rng.integers(3, 60, size=K) * cov
        a = rng.binomial(n, 0.003 + prev * (0.5 - 0.003))
        cases.append((B, a[:, None], (n - a)[:, None], 0.003))
    As = np.array([[17647, 28962, 31830], [65454, 115303, 57305], [2259, 8349, 3164], [13461, 6383, 11212]])
    Bs = np.array([[52353, 41038, 38170], [324546, 274697, 332695], [37741, 31651, 36836], [286539, 293617, 288788]])
    cases.append((tree_tools.get_matrix([-1, 0, 1, 2])[0], As, Bs, 0.001))
    worst_x = worst_f = 0.0
    for B, a, b, er in cases:
        K, M = a.shape
        res = solve4(B, a, b, er, [np.ones(K)] * M)
        total = 0.0
        for m in range(M):
            am, bm = a[:, m].astype(float), b[:, m].astype(float)
            x_s = res['s'][:, m]
            assert (x_s >= -1e-9).all() and (B.dot(x_s) <= 1 + 1e-9).all() and (B.dot(x_s) >= -1e-9).all()
            q = er + B.dot(x_s) * (0.5 - er)
            f1 = -(np.sum(am * np.log(q)) + np.sum(bm * np.log1p(-q)))
            x2, f2 = _barrier_newton(B, am, bm, er)
            worst_f = max(worst_f, abs(f1 - f2) / abs(f2))
            worst_x = max(worst_x, float(np.max(np.abs(x_s - x2))))
            total += -f1
        assert res['value'] == pytest.approx(total, rel=1e-12)
    assert worst_f <= 1e-9 and worst_x <= 2e-4, (worst_f, worst_x)


def test_shard_plan_assignment():
    """the assignment step of the multi-GPU sharding plan (pure host arithmetic behind bamse_set_shard +
    bamse_upload_problem): whole (parameter set, clustering) pairs by longest processing time, the most expensive
    ones split over all ranks (-1) when that lowers the estimated makespan; deterministic."""
    import ctypes
    lib = cuda_shim.load_library()

    def plan(P, Ks, W, build, enum):
        Ks = np.asarray(Ks, dtype=np.int32)
        b = np.zeros(13)
        e = np.zeros(13)
        for k, v in build.items():
            b[k] = v
        for k, v in enum.items():
            e[k] = v
        out = np.zeros(P * len(Ks), dtype=np.int32)
        assert lib.bamse_debug_shard_plan(P, len(Ks), Ks.ctypes.data_as(ctypes.c_void_p), W, b.ctypes.data_as(ctypes.c_void_p),
                                          e.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p)) == 0
        return out.reshape(P, len(Ks)), b, e

    # config-2-like: 49 + 26 + 4 clusterings, tables dominate -> whole pairs, 7 of the K = 6 ones on the fullest rank
    Ks = [6] * 49 + [5] * 26 + [4] * 4
    own, b, e = plan(1, Ks, 8, {6: 1000.0, 5: 150.0, 4: 40.0}, {6: 31.0, 5: 2.5, 4: 0.3})
    assert (own >= 0).all() and set(own.ravel()) == set(range(8))
    loads = np.zeros(8)
    for c, K in enumerate(Ks):
        loads[own[0, c]] += b[K] + e[K]
    assert loads.max() <= 7 * 1031.0 + 1e-9 and loads.max() / loads.mean() < 1.15
    own2, _, _ = plan(1, Ks, 8, {6: 1000.0, 5: 150.0, 4: 40.0}, {6: 31.0, 5: 2.5, 4: 0.3})
    np.testing.assert_array_equal(own, own2)
    # config-4-like: three huge clusterings on 8 ranks, enumeration dominates -> all split
    own, _, _ = plan(1, [10, 10, 10], 8, {10: 2.0e6}, {10: 3.0e7})
    assert (own == -1).all()
    # one rank: everything is owned by rank 0
    own, _, _ = plan(2, [5, 6], 1, {5: 1.0, 6: 2.0}, {5: 1.0, 6: 1.0})
    assert (own == 0).all()
    # a sweep: 64 parameter sets x 3 clusterings on 8 ranks -> 24 pairs per rank, nothing split
    own, _, _ = plan(64, [3, 4, 5], 8, {3: 1.0, 4: 5.0, 5: 30.0}, {3: 0.1, 4: 0.3, 5: 2.5})
    assert (own >= 0).all() and np.bincount(own.ravel(), minlength=8).tolist() == [24] * 8


def test_workloads_are_reproducible():
    """the BASELINE.json configurations as work lists: seeded, same clusterings every time."""
    from bamse_b200 import workloads
    a = workloads.build("c1")
    b = workloads.build("c1")
    assert a["n_evals_per_param"] == b["n_evals_per_param"] == 523
    assert [c.assign.tolist() for c in a["clusts"]] == [c.assign.tolist() for c in b["clusts"]]
    assert a["M"] == 3 and a["data"]["As"].shape == (50, 3) and len(a["errs"]) == 1
    c2 = workloads.build("c2")
    assert c2["n_evals_per_param"] == 397530 and len(c2["clusts"]) == 79 and c2["M"] == 10
    assert len(workloads.CONFIGS["c5"]["errs"]) == 64
    # the reference-produced goldens at the configurations' shapes (tests/golden/tree_scores_configs.json) were
    # taken on exactly these tables: the largest clustering of the work list
    from conftest import load_golden
    gold = {c["config"]: c for c in load_golden("tree_scores_configs.json")}
    for name, work in (("c1", a), ("c2", c2)):
        big = max(work["clusts"], key=lambda x: x.As.shape[0])
        assert np.array_equal(np.array(gold[name]["var"]), big.As) and np.array_equal(np.array(gold[name]["ref"]), big.Bs)


def test_algorithmic_op_count_uses_the_mean_leaf_count():
    """bench.py's roofline numerator: E[leaves] = K (1 - 1/K)^(K-1) over all labeled rooted trees
    (SURVEY section 8d), checked by enumeration."""
    import importlib.util
    import bamse_oracle as orc
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for K in range(2, 7):
        leaves = 0
        for i in range(K ** (K - 1)):
            p = orc.decode_tree(K, i)
            leaves += sum(1 for v in range(K) if v not in p)
        assert leaves / K ** (K - 1) == pytest.approx(bench.mean_leaves(K), rel=1e-12)
    n_c, xu, flops = bench.algorithmic_ops_per_eval(7, 10)
    assert n_c == pytest.approx(10 * (bench.mean_leaves(7) - 1) * 131328)
    assert xu > n_c and flops > 4 * n_c


def test_committed_bench_lines_follow_the_contract():
    """the bench lines kept under profiles/ carry every key of the bench.py contract."""
    import glob
    import json
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "scale", "r02_scale_c*.json")) +
                   glob.glob(os.path.join(ROOT, "profiles", "scale", "r01", "scale_c2_*.json")))
    assert len(files) >= 12
    for fn in files:
        d = json.load(open(fn))
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
            assert k in d, (fn, k)
        assert d["unit"] == "evals/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
        assert "workload" in d["config"] and "model" not in d["config"]
        assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"])
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
        assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"])
        assert d["gpu_launches"] > 0 and d["value"] > 0 and d["e2e"]["value"] > 0
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_bench_script_executes_end_to_end_with_stand_ins(built_library):
    """bench.py's own arm, every leg, on stand-ins for the CUDA runtime and the library context (tests/bench_dryrun.py):
    the driver-facing script parses its arguments, runs its legs in order and assembles a line with every key of the
    contract.  Shape only -- the numbers of a dry run mean nothing."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bench_dryrun.py"), "--config", "c1", "--steps", "3",
                          "--warmup", "3", "--no-cpu-baseline", "--also", "c2"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert k in d, k
    assert d["steps"] == 3 and d["warmup"] == 3 and d["n_gpus"] == 1 and d["config"]["evals_per_step"] == 523
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "cold_single_shot_ms")) <= set(d["e2e"])
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic", "executed")) <= set(d["roofline"])
    # the second configuration runs the REAL script in a child process: without a GPU it must fail there (no CPU
    # fallback) and leave the headline intact
    if not has_cuda():
        assert d["also_measured"]["config"] == "c2" and "no CUDA device" in d["also_measured"]["error"]
    groups = d["roofline"]["executed"]["by_kernel_group"]
    # the stand-in counts 900 builder ops with the counters on and 100 per enumeration: the split is by difference
    assert groups["table_builders"]["mufu_ops_per_step"] == 900 and groups["enumeration"]["mufu_ops_per_step"] == 100
    assert abs(groups["table_builders"]["share_of_kernel_time"] + groups["enumeration"]["share_of_kernel_time"] - 1) < 1e-12


def test_default_workload_of_the_bench_is_config_3():
    """no flags = BASELINE.json configs[2] at every N (one strong-scaling curve), with its own DRAM-traffic summary."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0
    import re
    assert re.search(r"--config\s*\{c1,c2,c3,c4,c5\}", out.stdout)
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert 'add_argument("--config", default="c3"' in src
    tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic_c3.json")))
    assert tj["config"] == "c3" and tj["n_gpus"] == 1
    assert tj["dram_bytes_read"] == sum(k["dram_bytes_read"] for k in tj["kernels"].values())
    assert tj["dram_bytes_write"] == sum(k["dram_bytes_write"] for k in tj["kernels"].values())


def test_the_binding_printed_in_integration_md_matches_the_library(built_library):
    """INTEGRATION.md section 2 prints the ctypes module a maintainer of the reference would add.  It must load
    against the built library as printed: every symbol it binds is exported, its record dtype is the ABI's, and
    without a GPU its constructor reports the library's own error (no CPU fallback) instead of crashing."""
    import re
    with open(os.path.join(ROOT, "INTEGRATION.md")) as f:
        text = f.read()
    block = re.search(r"```python\n(# bamse_cuda\.py.*?)```", text, re.S).group(1)
    assert 'ctypes.CDLL("libbamse_cuda.so")' in block
    ns = {}
    exec(compile(block.replace('"libbamse_cuda.so"', repr(built_library)), "INTEGRATION.md", "exec"), ns)
    assert ns["RECORD"] == cuda_shim.RECORD_DTYPE and ns["RECORD"].itemsize == 32
    for name in re.findall(r"_lib\.(bamse_\w+)", block):
        assert hasattr(ns["_lib"], name), name
    if not has_cuda():
        with pytest.raises(RuntimeError) as e:
            ns["Scorer"](0)
        assert "no CPU fallback" in str(e.value)


def test_synthetic_generator_draws_what_the_reference_draws():
    """bamse_b200/simdata.py against tests/golden/simdata.json = the reference's own sim_data_readcounts
    (simdata.py:18-31) at the (K_true, M, N, seed) of the bench workloads: given the same topology list the
    same tree is drawn and assign / As / Bs are identical (digests), i.e. the legacy global RNG is consumed in
    the same order."""
    import hashlib
    from conftest import load_golden
    from bamse_b200 import simdata
    digest = lambda a: hashlib.sha256(np.ascontiguousarray(a, dtype=np.int64).tobytes()).hexdigest()
    cases = load_golden("simdata.json")
    assert len(cases) == 7
    for g in cases:
        np.random.seed(g["seed"])
        d = simdata.sim_data_readcounts(g["K"], g["M"], {g["K"]: g["topologies"]}, g["N"], coverage=g["coverage"], error=g["error"])
        assert [int(x) for x in d["tree"]] == g["tree"]
        assert digest(d["assign"]) == g["assign"] and digest(d["As"]) == g["As"] and digest(d["Bs"]) == g["Bs"]
        assert d["As"][:3].tolist() == g["As_head"] and float(np.sum(d["vaf"])) == pytest.approx(g["vaf_sum"], rel=1e-13)
        # the same SET of topologies as our own enumerator (AllTrees.txt is not shipped)
        from bamse_b200.tree_tools import unlabeled_rooted_trees
        assert len(unlabeled_rooted_trees(g["K"])) == len(g["topologies"])


def test_candidate_generation_follows_the_reference_script():
    """workloads.candidate_clusterings / unique_Kmeans_clustering against tests/golden/candidates.json = the
    reference's own bamse.py:18-27 and :74-90 executed on bamse.dat and on the config-1 input with numpy's global
    RNG seeded: same partition-function row, same clusterings in the same order (i.e. the same RNG draws, the
    same use of the argsort indices as cluster counts), same prior / maxlikelihood constants."""
    from conftest import load_golden
    from bamse_b200 import workloads
    cases = load_golden("candidates.json")
    assert [c["id"] for c in cases] == ["bamse.dat", "bamse.dat-nc4", "c1"]
    for g in cases:
        As, Bs = np.array(g["var"]), np.array(g["ref"])
        np.random.seed(g["seed"])
        clusts, pf = workloads.candidate_clusterings(As, Bs, g["max_clusts"], g["n_trees"], g["err"])
        np.testing.assert_allclose(pf, g["pf"], rtol=0, atol=0)
        assert [[int(a) for a in c.assign] for c in clusts] == g["assignments"], g["id"]
        np.testing.assert_allclose([c.cluster_prior for c in clusts], g["cluster_prior"], rtol=1e-13)
        np.testing.assert_allclose([c.maxlikelihood for c in clusts], g["maxlikelihood"], rtol=1e-13)
        assert sorted({int(c.As.shape[0]) for c in clusts}) == sorted({k for k in g["candidate_numbers"] if k >= 1})


def test_remove_zeros_against_reference_golden():
    """tree_tools.remove_zeros against tests/golden/remove_zeros.json = the reference's own remove_zeros
    (tree_tools.py:159-175) on 56 random trees with random empty clones, 19 of which collapse at least one node
    (chains of empty nodes, an empty root with one child, empty leaves and empty branching nodes that must stay)."""
    from conftest import load_golden
    cases = load_golden("remove_zeros.json")
    assert len(cases) == 56 and sum(1 for c in cases if len(c["tree_out"]) < len(c["tree_in"])) >= 15
    for c in cases:
        t = {"tree": np.array(c["tree_in"]), "assign": np.array(c["assign_in"]), "clone_proportions": np.array(c["props_in"])}
        out = tree_tools.remove_zeros(t)
        assert [int(x) for x in out["tree"]] == c["tree_out"], c
        assert [int(x) for x in out["assign"]] == c["assign_out"], c
        np.testing.assert_array_equal(np.asarray(out["clone_proportions"]), np.array(c["props_out"]))
        np.testing.assert_array_equal(np.asarray(out["Bmatrix"]), np.array(c["Bmatrix"]))
        np.testing.assert_allclose(np.asarray(out["Amatrix"]).dot(np.array(c["Bmatrix"])), np.eye(len(c["tree_out"])), atol=1e-12)


def test_writers_are_byte_compatible_on_a_spread_of_trees(tmp_path):
    """tests/golden/outputs_more.json: the reference's own write_dot_files / write_text_output on 14 tree dicts
    (K = 1 .. 8 clones, 1 .. 6 samples, any root, absent subclones); ours writes the same bytes."""
    from conftest import load_golden
    from bamse_b200 import tree_io
    cases = load_golden("outputs_more.json")
    assert len(cases) == 14
    for i, g in enumerate(cases):
        spec = g["spec"]
        tree = dict(tree=np.array(spec["tree"]), root=spec["root"], assign=np.array(spec["assign"]),
                    vaf=np.array(spec["vaf"]), clone_proportions=np.array(spec["clone_proportions"]),
                    totalscore=spec["totalscore"], sample_names=spec["sample_names"])
        a, b, c = (str(tmp_path / ("%d_%s" % (i, n))) for n in ("t.dot", "s.dot", "o.txt"))
        tree_io.write_dot_files(tree, spec["sample_colors"], spec["cluster_colors"], a, b)
        tree_io.write_text_output([tree], c)
        K, M = len(spec["tree"]), len(spec["sample_names"])
        assert open(a).read() == g["subclones"], (K, M)
        assert open(b).read() == g["samples"], (K, M)
        assert open(c).read() == g["text"], (K, M)


def test_explicit_list_topk_ranks_like_the_enumeration_rule():
    """Context.score_list_topk (K_enum_mode = 0) with the GPU call replaced by the oracle's scores: constants,
    `log scre`, threshold rule and tie order are host work and must reproduce a brute-force ranking of the same list."""
    import bamse_oracle as orc
    from bamse_b200.clustering import _pack_items
    rng = np.random.default_rng(5)
    K_of_c = [3, 4, 3]
    tables = [np.array([[orc.distrib(int(a), int(b), 0.002)] for a, b in zip(rng.integers(5, 300, K), rng.integers(5, 300, K))])
              for K in K_of_c]
    tables[2] = tables[0]                                    # a duplicated clustering: exact ties, broken by the clustering index
    ctx = object.__new__(cuda_shim.Context)
    ctx.P, ctx.C, ctx.Kmax, ctx.K_of_c = 2, 3, 4, np.array(K_of_c, dtype=np.int32)
    items = [(c, orc.decode_tree(K, i), list(range(K))) for c, K in enumerate(K_of_c) for i in range(K ** (K - 1))]
    raw = np.array([orc.tree_score_samples(t, tables[c], orc.prior_const(len(t))).sum() for c, t, _ in items])
    ctx.score_list = lambda clust, par, nod, precision: np.stack([raw, raw - 0.25 * np.asarray(clust)])
    const = np.array([[0.0, 1.5, 0.0], [0.3, 0.0, 0.3]])
    thr = np.array([[-np.inf, np.median(raw[9:]), -np.inf], [-np.inf, -np.inf, 1e9]])
    rec, cnt = ctx.score_list_topk(*_pack_items(4, items), const, thr, topk=7)
    for p in range(2):
        rows = []
        for i, (c, t, _) in enumerate(items):
            s = (raw[i] if p == 0 else raw[i] - 0.25 * c) + math.log(orc.canon_weight(t)[1])
            if s >= thr[p, c] - 1e-12 * abs(thr[p, c]):
                rows.append((-(s + const[p, c]), c, i, s))
        rows.sort()
        assert cnt[p] == min(7, len(rows))
        assert [(int(r["clust"]), int(r["tree"])) for r in rec[p, :cnt[p]]] == [(c, i) for _, c, i, _ in rows[:7]]
        np.testing.assert_allclose(rec["score"][p, :cnt[p]], [s for _, _, _, s in rows[:7]], rtol=0, atol=0)
    assert (rec["clust"][1] != 2).all()                      # an unreachable threshold removes a clustering
    # clusterings 0 and 2 are the same tables with the same constant: their trees tie exactly, the lower clustering first
    r0 = rec[0, :cnt[0]]
    tied = [j for j in range(len(r0) - 1) if r0["total"][j] == r0["total"][j + 1]]
    assert tied and all((int(r0["clust"][j]), int(r0["clust"][j + 1])) == (0, 2) for j in tied)
    with pytest.raises(ValueError):
        ctx.score_list_topk(*_pack_items(4, [(0, [-1, 0], [0, 1])]), const, None, topk=1)      # not a full tree
