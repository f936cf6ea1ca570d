"""The routes a problem takes when the defaults do not fit or are switched off: a table budget that forces the
geometry down (fewer context rows -> irregular trees on the flat list -> programs only, down to (s, sh) = (2, 1)),
no pair table (the cut of every tree computed in the pair kernel), no sorted pair table, tree-major instead of
sample-major order.  Each must return the same top-k (the records are re-scored in fp64, so they are equal bit
for bit) and fp32 scores within 1e-6 of the check mode.  Runs on the B200 box:  python -m pytest tests -m gpu"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FP32_RTOL = 1e-6

KS = [7, 6, 8]
M = 3
# index ranges per clustering: all trees of K = 7 and K = 6, a slice of the 2 097 152 trees of K = 8 that starts
# at a non-zero root digit and is not a multiple of the chunk sizes
BEGIN = np.array([0, 0, 1234567], dtype=np.uint64)
END = np.array([7 ** 6, 6 ** 5, 1234567 + 60001], dtype=np.uint64)
DENSE = (2, 1234567 + 111, 3001)      # (clustering, first tree, count) of the dense comparison

VARIANTS = [
    {"BAMSE_PAIRTAB_KMAX": "0"},                  # no pair tables: decode -> cut -> rows in the pair kernel
    {"BAMSE_PAIR_SORT": "0"},                     # pair tables in index order
    {"BAMSE_SAMPLE_MAJOR": "0"},                  # tree-major pair kernel (what K >= 9 uses)
    {"BAMSE_PAIR_R": "4"},                        # K = 8 at (4, 4): 7.9 % irregular trees -> flat list + programs
    {"BAMSE_PAIR_R": "4", "BAMSE_PAIRTAB_KMAX": "0"},   # ... listed by the kernel's own cut, programs decoded in the kernel
    {"BAMSE_TABLE_GB": "0.1"},                    # budget below the default tables (~0.4 GB): step down part of the way
    {"BAMSE_TABLE_GB": "0.005"},                  # ... and all the way to programs on (s, sh) = (2, 1)
]


@pytest.fixture(scope="module")
def shim():
    from bamse_b200 import cuda_shim
    return cuda_shim


def _problem():
    rng = np.random.default_rng(808)
    Kmax = max(KS)
    var = np.zeros((len(KS), Kmax, M), dtype=np.int64)
    ref = np.zeros((len(KS), Kmax, M), dtype=np.int64)
    for c, K in enumerate(KS):
        frac = np.sort(rng.dirichlet(np.ones(K + 1))[:K])[::-1]
        n = rng.integers(20, 120, size=K)
        for k in range(K):
            for m in range(M):
                tot = int(n[k]) * 150
                v = rng.binomial(tot, min(0.49, frac[k] * rng.uniform(0.6, 1.2) * 0.5 + 0.003))
                var[c, k, m], ref[c, k, m] = v, tot - v
    return var, ref


def _level_constants(shim, var, ref):
    """per-clustering constants that bring the best trees of the three clusterings within 0.1 of each other, so
    that the global top-k mixes the clusterings (best total of clustering c alone, by three top-1 calls)."""
    const = np.zeros((1, len(KS)))
    with shim.Context(0) as ctx:
        ctx.upload_problem([0.003], KS, var, ref)
        for c in range(len(KS)):
            end = BEGIN.copy()
            end[c] = END[c]                           # every other range is empty
            rec, _, cnt = ctx.score_topk(np.zeros((1, len(KS))), None, BEGIN, end, topk=1, precision=shim.FP32)
            assert cnt[0] == 1 and int(rec[0, 0]["clust"]) == c
            const[0, c] = -float(rec[0, 0]["total"]) + (0.0, 0.05, -0.05)[c]
    return const


def _answers(shim, var, ref, const):
    with shim.Context(0) as ctx:
        ctx.upload_problem([0.003], KS, var, ref)
        rec, par, cnt = ctx.score_topk(const, None, BEGIN, END, topk=10, precision=shim.FP32)
        assert ctx.last_topk_verified
        evals = ctx.last_eval_count
        c, b, n = DENSE
        s32, w32 = ctx.score_range(0, c, b, n, shim.FP32)
        s7, _ = ctx.score_range(0, 0, 100000, 2000, shim.FP32)
        s64, w64 = ctx.score_range(0, c, b, n, shim.FP64)
        s7_64, _ = ctx.score_range(0, 0, 100000, 2000, shim.FP64)
    return dict(rec=rec, par=par, cnt=cnt, evals=evals, s32=s32, w32=w32, s64=s64, w64=w64, s7=s7, s7_64=s7_64)


def test_fallback_routes_return_the_default_answers(shim, monkeypatch):
    var, ref = _problem()
    names = ("BAMSE_PAIRTAB_KMAX", "BAMSE_PAIR_SORT", "BAMSE_SAMPLE_MAJOR", "BAMSE_PAIR_R", "BAMSE_TABLE_GB",
             "BAMSE_FOREST_S", "BAMSE_FOREST_SH", "BAMSE_PAIRS")
    for name in names:
        monkeypatch.delenv(name, raising=False)
    const = _level_constants(shim, var, ref)
    want = _answers(shim, var, ref, const)
    assert want["cnt"][0] == 10 and want["evals"] == int(np.sum(END - BEGIN))
    assert np.max(np.abs(want["s32"] - want["s64"]) / np.abs(want["s64"])) <= FP32_RTOL
    assert np.max(np.abs(want["s7"] - want["s7_64"]) / np.abs(want["s7_64"])) <= FP32_RTOL
    assert int(want["rec"][0, 0]["clust"]) == 1 and len({int(c) for c in want["rec"][0]["clust"]}) >= 2   # mixed
    for env in VARIANTS:
        for name, value in env.items():
            monkeypatch.setenv(name, value)
        got = _answers(shim, var, ref, const)
        for name in env:
            monkeypatch.delenv(name)
        assert got["evals"] == want["evals"], env
        np.testing.assert_array_equal(got["cnt"], want["cnt"], err_msg=str(env))
        np.testing.assert_array_equal(got["rec"]["tree"], want["rec"]["tree"], err_msg=str(env))
        np.testing.assert_array_equal(got["rec"]["clust"], want["rec"]["clust"], err_msg=str(env))
        np.testing.assert_array_equal(got["rec"]["total"], want["rec"]["total"], err_msg=str(env))
        np.testing.assert_array_equal(got["par"], want["par"], err_msg=str(env))
        np.testing.assert_array_equal(got["s64"], want["s64"], err_msg=str(env))     # the check mode does not depend on any of it
        np.testing.assert_array_equal(got["w32"], want["w64"], err_msg=str(env))
        assert np.max(np.abs(got["s32"] - want["s64"]) / np.abs(want["s64"])) <= FP32_RTOL, env
        assert np.max(np.abs(got["s7"] - want["s7_64"]) / np.abs(want["s7_64"])) <= FP32_RTOL, env


def test_a_budget_below_the_smallest_tables_is_an_error_not_a_crash(shim, monkeypatch):
    var, ref = _problem()
    monkeypatch.setenv("BAMSE_TABLE_GB", "0.0001")
    with shim.Context(0) as ctx:
        with pytest.raises(shim.BamseCudaError) as e:
            ctx.upload_problem([0.003], KS, var, ref)
        assert e.value.code == -5 and "budget" in str(e.value)          # BAMSE_ENOMEM
        monkeypatch.delenv("BAMSE_TABLE_GB")
        ctx.upload_problem([0.003], KS, var, ref)                       # the context is still usable
        rec, _, cnt = ctx.score_topk(np.zeros((1, 3)), None, BEGIN, END, topk=3, precision=shim.FP32)
        assert cnt[0] == 3
