"""Generates tests/golden/*.json by calling the UNMODIFIED reference functions
(imported from /root/reference through oracle/ref_shim.py).  Run in the build
container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py                      # scores vectors canon constants search outputs
    python tests/golden/make_golden.py highk search_big     # round 2: K = 8..12, bamse.dat / K = 5 search (minutes)
    python tests/golden/make_golden.py edges configs simdata candidates remove_zeros outputs_more

Everything is seeded; re-running reproduces the committed files bit for bit
(up to json float repr, which round-trips fp64 exactly).
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402

ref = ref_shim.load()
cl = ref.clustering
tt = ref.tree_tools


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f, indent=0, separators=(",", ":"))
    print("wrote", name)


def ref_tables(As, Bs, err, K_prior):
    K, M = As.shape
    D = np.array([[cl.distrib(As[k, m], Bs[k, m], err) for m in range(M)] for k in range(K)])
    P = np.array([[cl.prior_fractions(K_prior) for m in range(M)] for k in range(K)])
    return D, P


def random_tree(rng, k):
    """random labeled rooted tree as a parent vector with arbitrary root position."""
    perm = rng.permutation(k)
    parent = [-2] * k
    parent[perm[0]] = -1
    for i in range(1, k):
        parent[perm[i]] = int(perm[rng.integers(0, i)])
    return parent


# ---------------------------------------------------------------- 1. tree scores
def gen_scores():
    cases = []
    # SURVEY 8a G1..G8 inputs, recomputed here (not transcribed) -------------------
    g_as = np.array([[17647, 28962, 31830], [65454, 115303, 57305], [2259, 8349, 3164], [13461, 6383, 11212]])
    g_bs = np.array([[52353, 41038, 38170], [324546, 274697, 332695], [37741, 31651, 36836], [286539, 293617, 288788]])
    fixed = [
        ("G1", g_as, g_bs, 0.001, [-1, 0, 1, 2]),
        ("G2", g_as, g_bs, 0.001, [-1, 0, 0, 0]),
        ("G3", g_as, g_bs, 0.001, [-1, 0, 0, 1]),
        ("G4", np.array([[90, 40]]), np.array([[110, 160]]), 0.003, [-1]),
        ("G5", np.array([[95, 60], [30, 10]]), np.array([[105, 140], [170, 190]]), 0.003, [-1, 0]),
        ("G6", np.array([[95, 60], [30, 10]]), np.array([[105, 140], [170, 190]]), 0.003, [1, -1]),
        ("G7", np.array([[800], [300], [250]]), np.array([[1200], [1700], [1750]]), 0.003, [-1, 0, 0]),
        ("G8", np.array([[400, 380], [150, 30], [120, 200], [60, 10], [20, 90]]),
         np.array([[600, 620], [850, 970], [880, 800], [940, 990], [980, 910]]), 0.01, [2, 2, -1, 0, 0]),
    ]
    for name, As, Bs, err, tree in fixed:
        D, P = ref_tables(As, Bs, err, As.shape[0])
        s = cl.tree_integrate_multisample(tree, D, P)
        cases.append(dict(id=name, var=As.tolist(), ref=Bs.tolist(), err=err, K_prior=int(As.shape[0]),
                          tree=tree, nodes=list(range(As.shape[0])), per_sample=s.tolist(), total=float(np.sum(s))))
    # random cases: varied K, coverage, err, root position, node subsets -------------
    rng = np.random.default_rng(20261019)
    n = 0
    for K in (1, 2, 3, 4, 5, 6, 7):
        for rep in range(3 if K <= 5 else 2):
            M = int(rng.integers(1, 4))
            cov = int(rng.choice([30, 200, 2000, 40000, 400000]))
            frac = rng.dirichlet(np.ones(K + 1))[:K]
            frac = np.sort(frac)[::-1]
            cm = np.clip(np.cumsum(frac[::-1])[::-1][:, None] * rng.uniform(0.6, 1.0, size=(K, M)), 0, 1)
            err = float(rng.choice([0.0005, 0.001, 0.003, 0.01]))
            As = rng.binomial(cov, cm * (0.5 - err) + err)
            Bs = cov - As
            # a node subset (partial forest component) half of the time
            k = K if (rep % 2 == 0 or K == 1) else int(rng.integers(1, K + 1))
            nodes = sorted(rng.choice(K, size=k, replace=False).tolist()) if k < K else list(range(K))
            if rep == 1 and k > 1:
                nodes = [int(x) for x in rng.permutation(nodes)]
            tree = random_tree(rng, k)
            D, P = ref_tables(As, Bs, err, K)
            s = cl.tree_integrate_multisample(np.array(tree), D[nodes], P[nodes])
            cases.append(dict(id="R%d" % n, var=As.tolist(), ref=Bs.tolist(), err=err, K_prior=K, tree=tree,
                              nodes=[int(x) for x in nodes], per_sample=s.tolist(), total=float(np.sum(s))))
            n += 1
    dump("tree_scores.json", cases)


def gen_remove_zeros():
    """round 2: tree_tools.remove_zeros (tree_tools.py:159-175, with collapse_node / collapse_assign) of the unmodified
    reference on random trees with random empty clones (numpy-array trees, as bamse.py hands them over): the tree,
    the assignment and the clone proportions after collapsing every empty single-child node."""
    import copy
    rng = np.random.default_rng(159)
    cases = []
    for K in (2, 3, 4, 5, 6, 7, 8):
        for rep in range(8):
            tree = np.array(random_tree(rng, K))
            props = rng.uniform(0.05, 0.4, size=(K, 3))
            empty = [int(x) for x in np.nonzero(rng.uniform(size=K) < 0.4)[0]]
            for n in empty:
                props[n] = rng.choice([0.0, 0.0002], size=3)
            assign = rng.integers(0, K, size=3 * K)
            t = {"tree": tree.copy(), "assign": assign.copy(), "clone_proportions": props.copy()}
            out = tt.remove_zeros(copy.deepcopy(t))
            cases.append(dict(tree_in=tree.tolist(), assign_in=assign.tolist(), props_in=props.tolist(), empty=empty,
                              tree_out=[int(x) for x in out["tree"]], assign_out=[int(x) for x in out["assign"]],
                              props_out=np.asarray(out["clone_proportions"]).tolist(),
                              Bmatrix=np.asarray(out["Bmatrix"]).tolist()))
    print(sum(1 for c in cases if len(c["tree_out"]) < len(c["tree_in"])), "of", len(cases), "cases collapse something")
    dump("remove_zeros.json", cases)


def gen_candidates():
    """round 2: the reference's candidate generation -- `unique_Kmeans_clustering` (bamse.py:18-27) and the model
    selection loop (bamse.py:75-90), taken from the reference's bamse.py AS SOURCE TEXT (the file is a script, not a
    module) and executed here against the reference's own clustering / tree_tools -- on bamse.dat and on the
    config-1 simdata input, with numpy's global RNG seeded (sklearn's random_state=None draws from it)."""
    import ast
    import contextlib
    import io
    import pandas as pd
    from sklearn.cluster import KMeans
    src = open("/root/reference/bamse.py").read()
    lines = src.split("\n")
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "unique_Kmeans_clustering"][0]
    fn_src = "\n".join(lines[fn.lineno - 1:fn.end_lineno])
    first = next(i for i, l in enumerate(lines) if l.startswith("clusterings = []"))
    last = next(i for i, l in enumerate(lines) if "clusts += list(unique_Kmeans_clustering" in l)
    loop_src = "\n".join(lines[first:last + 1])
    assert first + 1 == 74 and last + 1 == 90, (first, last)
    table = pd.read_table("/root/reference/bamse.dat")
    names = sorted({c[:-4] for c in table.columns if c.endswith(".var")})
    dat = (table[[x + ".var" for x in names]].values, table[[x + ".ref" for x in names]].values)
    sys.path.insert(0, ROOT)
    from bamse_b200 import simdata as our_sim          # pinned to the reference's generator by simdata.json
    d = our_sim.simulate(4, 3, 50, 12345)
    cases = []
    for tag, (As, Bs), max_clusts, n_trees, err, seed in [("bamse.dat", dat, 6, 10, 0.001, 11), ("bamse.dat-nc4", dat, 4, 5, 0.003, 12),
                                                          ("c1", (d["As"], d["Bs"]), 5, 10, 0.001, 13346)]:
        ns = dict(np=np, KMeans=KMeans, clustering=cl.clustering, reorder_assignmens=tt.reorder_assignmens)
        exec(compile(fn_src, "bamse.py", "exec"), ns)
        max_clusts = min(len(As), max_clusts)
        pf = cl.part_func(len(As), max_clusts)
        ns.update(X={"As": As, "Bs": Bs}, max_clusts=max_clusts, pf=pf, err=err, n_trees=n_trees)
        np.random.seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            exec(compile(loop_src, "bamse.py", "exec"), ns)
        cases.append(dict(id=tag, var=np.asarray(As).tolist(), ref=np.asarray(Bs).tolist(), max_clusts=int(max_clusts), n_trees=n_trees,
                          err=err, seed=seed, pf=[float(x) for x in pf], raw_scores=[float(x) for x in ns["raw_scores"]],
                          candidate_numbers=[int(x) for x in ns["candidate_numbers"]],
                          assignments=[[int(a) for a in c.assign] for c in ns["clusts"]],
                          cluster_prior=[float(c.cluster_prior) for c in ns["clusts"]],
                          maxlikelihood=[float(c.maxlikelihood) for c in ns["clusts"]]))
        print(tag, cases[-1]["candidate_numbers"], len(cases[-1]["assignments"]))
    dump("candidates.json", cases)


def gen_simdata():
    """round 2: the reference's own synthetic generator (simdata.py:18-31, imported unmodified with matplotlib /
    seaborn stubbed out -- they are only used by its plotting helpers) at the (K_true, M, N, seed) of the five
    bench workloads and two more: the topology list handed in, the tree drawn, and digests of assign / As / Bs."""
    import hashlib
    import pickle
    import types
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, "/root/reference")
    import simdata as rsim
    with open("/root/reference/AllTrees.txt", "rb") as f:
        T, S = pickle.load(f, encoding="latin1")
    digest = lambda a: hashlib.sha256(np.ascontiguousarray(a, dtype=np.int64).tobytes()).hexdigest()
    cases = []
    for K, M, N, seed, cov, err in [(4, 3, 50, 12345, 200, 0.003), (6, 10, 500, 2, 200, 0.003), (7, 20, 2000, 3, 200, 0.003),
                                    (9, 5, 1000, 4, 200, 0.003), (5, 10, 500, 5, 200, 0.003), (3, 2, 40, 77, 10000, 0.001),
                                    (8, 4, 300, 8, 60, 0.01)]:
        np.random.seed(seed)
        d = rsim.sim_data_readcounts(K, M, T, N, coverage=cov, error=err)
        cases.append(dict(K=K, M=M, N=N, seed=seed, coverage=cov, error=err,
                          topologies=[[int(x) for x in t] for t in T[K]], tree=[int(x) for x in d["tree"]],
                          assign=digest(d["assign"]), As=digest(d["As"]), Bs=digest(d["Bs"]),
                          As_head=np.asarray(d["As"])[:3].tolist(), vaf_sum=float(np.sum(d["vaf"]))))
    dump("simdata.json", cases)


def gen_scores_configs():
    """round 2: reference-produced scores at the SHAPES of BASELINE.json's configurations -- the aggregated tables
    of the largest clustering of each of bench.py's workloads c1 .. c5 (K = 4 / 6 / 7 / 10 / 5 clusters, M = 3 /
    10 / 20 / 5 / 10 samples; c5 at the first and the last error rate of its sweep), the chain, the star and three
    random labeled rooted trees each."""
    sys.path.insert(0, ROOT)
    import bamse_oracle as orc
    from bamse_b200 import workloads
    cases = []
    for name in ("c1", "c2", "c3", "c4", "c5"):
        work = workloads.build(name)
        c = max(work["clusts"], key=lambda x: x.As.shape[0])
        As, Bs = np.array(c.As), np.array(c.Bs)
        K = As.shape[0]
        rng = np.random.default_rng(abs(hash(name)) % 1000 if False else {"c1": 1, "c2": 2, "c3": 3, "c4": 4, "c5": 5}[name])
        trees = [[-1] + list(range(K - 1)), [-1] + [0] * (K - 1)] + \
                [orc.decode_tree(K, int(rng.integers(0, K ** (K - 1)))) for _ in range(3)]
        errs = [work["errs"][0]] if len(work["errs"]) == 1 else [work["errs"][0], work["errs"][-1]]
        for e in errs:
            D, P = ref_tables(As, Bs, e, K)
            for j, tree in enumerate(trees):
                t0 = time.time()
                sc = cl.tree_integrate_multisample(tree, D, P)
                cases.append(dict(id="%s-e%g-t%d" % (name, e, j), config=name, var=As.tolist(), ref=Bs.tolist(), err=float(e),
                                  K_prior=K, tree=tree, nodes=list(range(K)), per_sample=[float(x) for x in sc],
                                  total=float(np.sum(sc)), scre=int(tt.canonizeF(tree)['scre'])))
                print(cases[-1]["id"], K, As.shape[1], "%.1fs" % (time.time() - t0), cases[-1]["total"])
    dump("tree_scores_configs.json", cases)


def gen_scores_edges():
    """round 2: the inputs of tests/test_gpu_edges.py through the UNMODIFIED reference -- clusters without variant
    reads / without any reads / without reference reads (all 64 trees), error rates 1e-6 and 0.499, one sample
    and 64 samples, K = 1 and K = 2.  Closes the chain reference == oracle (here) == CUDA (test_gpu_edges.py) at
    the edges of the domain."""
    import bamse_oracle as orc
    cases = []

    def add(tag, As, Bs, err, tree_indices):
        K = As.shape[0]
        D, P = ref_tables(As, Bs, err, K)
        for i in tree_indices:
            tree = orc.decode_tree(K, int(i))
            s = cl.tree_integrate_multisample(tree, D, P)
            cases.append(dict(id="%s-%d" % (tag, i), var=As.tolist(), ref=Bs.tolist(), err=err, K_prior=K, tree=tree,
                              tree_index=int(i), nodes=list(range(K)), per_sample=[float(x) for x in s],
                              total=float(np.sum(s)), scre=int(tt.canonizeF(tree)['scre'])))

    var = np.array([[120, 80, 95], [0, 0, 0], [30, 0, 12], [7, 55, 3]], dtype=np.int64)
    ref_ = np.array([[380, 420, 405], [250, 310, 275], [170, 0, 188], [193, 0, 197]], dtype=np.int64)
    add("zero", var, ref_, 0.002, range(64))
    rng = np.random.default_rng(499)          # test_error_rates_at_both_ends_in_one_sweep
    var = rng.integers(0, 400, size=(4, 2)).astype(np.int64)
    ref_ = rng.integers(300, 900, size=(4, 2)).astype(np.int64)
    for e in (1e-6, 0.499):
        add("err%g" % e, var, ref_, e, range(0, 64, 3))
    rng = np.random.default_rng(64)           # test_one_sample_and_the_maximum_number_of_samples
    for M in (1, 64):
        var = rng.integers(0, 200, size=(3, M)).astype(np.int64)
        ref_ = rng.integers(100, 400, size=(3, M)).astype(np.int64)
        add("M%d" % M, var, ref_, 0.003, range(9) if M == 1 else (0, 4, 8))
    add("K2", np.array([[40, 10], [5, 30]]), np.array([[160, 190], [195, 170]]), 0.003, range(2))
    add("K1", np.array([[22, 31]]), np.array([[178, 169]]), 0.003, range(1))
    add("allzero", np.zeros((3, 2), dtype=np.int64), np.zeros((3, 2), dtype=np.int64), 0.003, range(9))
    dump("tree_scores_edges.json", cases)


# ------------------------------------------------------------- 2. vector operators
def gen_vectors():
    rng = np.random.default_rng(7)
    out = {"distrib": [], "convolve": [], "trap_int": []}
    for a, b, err in [(0, 0, 0.001), (1, 0, 0.001), (0, 7, 0.01), (90, 110, 0.003), (17647, 52353, 0.001),
                      (324546, 65454, 0.001), (2000000, 10, 0.0001)]:
        y = cl.distrib(a, b, err)
        out["distrib"].append(dict(a=a, b=b, err=err, idx=list(range(0, 512, 7)) + [511],
                                   y=y[list(range(0, 512, 7)) + [511]].tolist()))
    for (a1, b1, a2, b2, err) in [(90, 110, 30, 170, 0.003), (17647, 52353, 2259, 37741, 0.001),
                                  (5, 3, 900, 100, 0.01), (65454, 324546, 13461, 286539, 0.001)]:
        u = cl.distrib(a1, b1, err)
        v = cl.distrib(a2, b2, err)
        c = cl.my_convolve(u, v)
        t = cl.trap_int(c)
        sel = list(range(0, 512, 5)) + [511]
        out["convolve"].append(dict(a1=a1, b1=b1, a2=a2, b2=b2, err=err, idx=sel, y=c[sel].tolist()))
        out["trap_int"].append(dict(a1=a1, b1=b1, a2=a2, b2=b2, err=err, idx=sel, y=t[sel].tolist()))
    dump("vectors.json", out)


# --------------------------------------------------------------- 3. canon weights
def gen_canon():
    import pickle
    with open("/root/reference/AllTrees.txt", "rb") as f:
        T, S = pickle.load(f, encoding="latin1")
    rng = np.random.default_rng(11)
    out = {"alltrees_counts": [len(t) for t in T], "alltrees_scre": {}, "random": []}
    for K in range(1, 8):
        out["alltrees_scre"][str(K)] = [dict(tree=[int(x) for x in t], scre=int(tt.canonizeF(t)["scre"]),
                                             canstr=tt.canonizeF(t)["canstr"]) for t in T[K]]
    for K in range(2, 11):
        for _ in range(40):
            t = random_tree(rng, K)
            r = tt.canonizeF(t)
            out["random"].append(dict(tree=t, scre=int(r["scre"]), canstr=r["canstr"]))
    # canonical strings of every AllTrees topology (for the set-equality test of our enumerator)
    out["alltrees_canstr"] = {str(K): sorted(tt.canonizeF(t)["canstr"] for t in T[K]) for K in range(1, 11)}
    dump("canon.json", out)


# ----------------------------------------------------------- 4. clustering constants
def gen_constants():
    rng = np.random.default_rng(5)
    out = {"part_func": [], "penalty": [], "clustering": []}
    for n, kmax in [(1, 5), (3, 5), (4, 3), (10, 4), (50, 5), (80, 5), (500, 7), (80, 12), (6, 12)]:
        out["part_func"].append(dict(n=n, kmax=kmax, pf=cl.part_func(n, kmax).tolist()))
    for n, kmax, K in [(10, 4, 4), (50, 5, 3), (80, 5, 5), (80, 12, 9), (200, 8, 1)]:
        pf = cl.part_func(n, kmax)
        assign = np.concatenate([np.arange(K), rng.integers(0, K, n - K)])
        rng.shuffle(assign)
        out["penalty"].append(dict(n=n, kmax=kmax, assign=assign.tolist(),
                                   value=float(cl.log_conf_penalty(assign, pf))))
    for (N, M, K, cov) in [(12, 2, 3, 200), (30, 3, 4, 1000), (50, 3, 5, 200)]:
        pf = cl.part_func(N, 6)
        assign = np.concatenate([np.arange(K), rng.integers(0, K, N - K)])
        am = rng.binomial(cov, rng.uniform(0.02, 0.45, size=(N, M)))
        bm = cov - am
        c = cl.clustering(am, bm, assign, pf, 0.002)
        out["clustering"].append(dict(am=am.tolist(), bm=bm.tolist(), assign=assign.tolist(), pf=pf.tolist(), err=0.002,
                                      As=c.As.tolist(), Bs=c.Bs.tolist(), maxlikelihood=float(c.maxlikelihood),
                                      cluster_prior=float(c.cluster_prior)))
    dump("constants.json", out)


# ------------------------------------------------------------------ 5. search path
def simdata(K, M, N, seed, T, coverage=200, error=0.003):
    """simdata.py:18-31 through the reference's own get_matrix (legacy global RNG)."""
    np.random.seed(seed)
    ind = np.random.randint(len(T[K]))
    true_tree = T[K][ind]
    true_sm = np.array([np.random.dirichlet(np.ones(K + 1))[:K] for m in range(M)])
    B = tt.get_matrix(true_tree)[0]
    true_cm = np.array([B.dot(x) for x in true_sm]).transpose()
    ps = np.random.dirichlet(np.ones(K))
    assign = np.random.choice(range(K), N - 2 * K, p=ps)
    assign = np.concatenate((np.repeat(np.arange(K), 2), assign))
    covs = coverage * np.ones((N, M)).astype(int)
    As = np.random.binomial(covs, true_cm[assign] * (0.5 - error) + error)
    return As, covs - As, assign, [int(x) for x in true_tree]


def gen_search():
    """reorder -> candidate -> threshold -> get_trees2 -> analysed scores, on small inputs
    (each reference eval costs 0.05-1 s, so K <= 4 here)."""
    import contextlib
    import io
    import pickle
    with open("/root/reference/AllTrees.txt", "rb") as f:
        T, S = pickle.load(f, encoding="latin1")
    out = []
    specs = [dict(K=3, M=2, N=24, seed=101, err=0.003, cov=200),
             dict(K=4, M=3, N=50, seed=12345, err=0.001, cov=200),
             dict(K=4, M=2, N=40, seed=77, err=0.003, cov=60),
             dict(K=2, M=3, N=20, seed=3, err=0.01, cov=500)]
    for sp in specs:
        t0 = time.time()
        am, bm, assign, true_tree = simdata(sp["K"], sp["M"], sp["N"], sp["seed"], T, coverage=sp["cov"])
        # clustering = the true assignment, first-seen relabelled (what bamse.py:23 does to KMeans labels)
        assign = tt.reorder_assignmens(assign)
        pf = cl.part_func(sp["N"], 6)
        c = cl.clustering(am, bm, tuple(int(x) for x in assign), pf, sp["err"])
        with contextlib.redirect_stdout(io.StringIO()):
            c.reorder()
            cand = c.get_candidate_tree()[0]
            K = c.As.shape[0]
            thr = c.get_tree_score(cand["tree"], list(range(K))) + np.log(tt.canonizeF(cand["tree"])["scre"])
            trees = c.get_trees2(cand)
        kept = []
        for t in trees:
            sc = c.get_tree_score(t["tree"], list(range(K)))
            scre = int(tt.canonizeF(t["tree"])["scre"])
            kept.append(dict(tree=[int(x) for x in t["tree"]], root=int(t["root"]), score=float(sc), scre=scre,
                             totalscore=float(sc + np.log(scre) + c.cluster_prior + c.maxlikelihood)))
        # the 2-node scores that drive reorder(), in ORIGINAL labels (before reorder)
        out.append(dict(spec=sp, am=am.tolist(), bm=bm.tolist(), assign_in=[int(x) for x in assign],
                        true_tree=true_tree, pf=pf.tolist(),
                        order=[int(x) for x in c.order], assign_out=[int(x) for x in c.assign],
                        As=c.As.tolist(), Bs=c.Bs.tolist(), candidate=[int(x) for x in cand["tree"]],
                        threshold=float(thr), kept=kept, cluster_prior=float(c.cluster_prior),
                        maxlikelihood=float(c.maxlikelihood)))
        print("search case", sp, "kept", len(kept), "%.1fs" % (time.time() - t0))
    dump("search.json", out)


# ------------------------------------------------- 6. tree scores at K = 8 .. 12 (round 2)
def gen_scores_highk():
    """reference-produced get_tree_score values where configs 3/4 live (K = 8, 9, 10, 11, 12): random labeled
    rooted trees + the chain and the star of every K, simdata-like and high-coverage tables, full node sets.
    One reference eval costs (K + leaves - 1) my_convolve calls per sample (~15 ms each)."""
    cases = []
    rng = np.random.default_rng(20261020)
    n = 0
    for K in (8, 9, 10, 11, 12):
        shapes = [random_tree(rng, K) for _ in range(4)]
        shapes.append([-1] + list(range(K - 1)))            # chain: scans only
        shapes.append([-1] + [0] * (K - 1))                 # star: K - 2 convolutions
        # a deep tree with a wide node in the middle
        shapes.append([-1, 0, 1] + [2] * (K - 3))
        for rep, tree in enumerate(shapes):
            M = 3 if K <= 10 else 2
            cov = int([200, 1000, 30000, 200, 2000, 200, 100000][rep])
            frac = rng.dirichlet(np.ones(K + 1))[:K]
            frac = np.sort(frac)[::-1]
            cm = np.clip(np.cumsum(frac[::-1])[::-1][:, None] * rng.uniform(0.6, 1.0, size=(K, M)), 0, 1)
            err = float(rng.choice([0.001, 0.003]))
            As = rng.binomial(cov, cm * (0.5 - err) + err)
            Bs = cov - As
            D, P = ref_tables(As, Bs, err, K)
            t0 = time.time()
            s = cl.tree_integrate_multisample(np.array(tree), D, P)
            scre = int(tt.canonizeF(tree)["scre"])
            cases.append(dict(id="H%d" % n, var=As.tolist(), ref=Bs.tolist(), err=err, K_prior=K, tree=[int(x) for x in tree],
                              nodes=list(range(K)), per_sample=s.tolist(), total=float(np.sum(s)), scre=scre))
            print("highk case", n, "K", K, "%.1fs" % (time.time() - t0), flush=True)
            n += 1
    dump("tree_scores_highk.json", cases)


# ------------------------------------------- 7. search path on bamse.dat and at K = 5 (round 2)
def _search_case(am, bm, assign, pf, err, spec, extra=None):
    import contextlib
    import io
    t0 = time.time()
    c = cl.clustering(am, bm, tuple(int(x) for x in assign), pf, err)
    with contextlib.redirect_stdout(io.StringIO()):
        c.reorder()
        cand = c.get_candidate_tree()[0]
        K = c.As.shape[0]
        thr = c.get_tree_score(cand["tree"], list(range(K))) + np.log(tt.canonizeF(cand["tree"])["scre"])
        trees = c.get_trees2(cand)
    kept = []
    for t in trees:
        sc = c.get_tree_score(t["tree"], list(range(K)))
        scre = int(tt.canonizeF(t["tree"])["scre"])
        kept.append(dict(tree=[int(x) for x in t["tree"]], root=int(t["root"]), score=float(sc), scre=scre,
                         totalscore=float(sc + np.log(scre) + c.cluster_prior + c.maxlikelihood)))
    out = dict(spec=spec, am=np.asarray(am).tolist(), bm=np.asarray(bm).tolist(), assign_in=[int(x) for x in assign],
               pf=pf.tolist(), order=[int(x) for x in c.order], assign_out=[int(x) for x in c.assign],
               As=c.As.tolist(), Bs=c.Bs.tolist(), candidate=[int(x) for x in cand["tree"]], threshold=float(thr),
               kept=kept, cluster_prior=float(c.cluster_prior), maxlikelihood=float(c.maxlikelihood))
    if extra:
        out.update(extra)
    print("search case", spec, "kept", len(kept), "%.1fs" % (time.time() - t0), flush=True)
    return out


def gen_search_big():
    """(a) the reference's own sample input bamse.dat (80 mutations x 3 samples, coverage 10 000: the hard
    dynamic-range case) at its KMeans-4 and KMeans-5 clusterings, through reorder -> get_candidate_tree ->
    get_trees2 exactly as bamse.py:91-98 drives them; (b) one K = 5 simdata case (kept set at K = 5)."""
    import pickle
    import pandas as pd
    from sklearn.cluster import KMeans
    with open("/root/reference/AllTrees.txt", "rb") as f:
        T, S = pickle.load(f, encoding="latin1")
    out = []
    tab = pd.read_table("/root/reference/bamse.dat")
    names = ["S1", "S2", "S3"]
    am = tab[[x + ".var" for x in names]].values
    bm = tab[[x + ".ref" for x in names]].values
    pf = cl.part_func(len(tab), 10)
    Y = (am + 0.0) / (am + bm)
    for K in (4, 5):
        labels = KMeans(n_clusters=K, n_init=100, random_state=0).fit(Y).labels_
        assign = tt.reorder_assignmens(labels)                       # bamse.py:23
        out.append(_search_case(am, bm, assign, pf, 0.001, dict(source="bamse.dat", K=K, err=0.001)))
    sp = dict(K=5, M=3, N=60, seed=2026, err=0.003, cov=200)
    am, bm, assign, true_tree = simdata(sp["K"], sp["M"], sp["N"], sp["seed"], T, coverage=sp["cov"])
    assign = tt.reorder_assignmens(assign)
    out.append(_search_case(am, bm, assign, cl.part_func(sp["N"], 6), sp["err"], sp, dict(true_tree=true_tree)))
    dump("search_big.json", out)


def gen_outputs():
    """dot / text files written by the reference's own tree_io on a fixed tree dict.
    asciitree is absent here, so its LeftAligned drawer is supplied by our restatement
    (bamse_b200.tree_io.LeftAligned): everything except that drawer is pinned."""
    sys.path.insert(0, ROOT)
    from bamse_b200.tree_io import LeftAligned
    ref.tree_io.LeftAligned = LeftAligned
    rng = np.random.default_rng(3)
    tree = [2, 2, -1, 0, 0, 1]
    K, M = 6, 3
    B = tt.get_matrix(tree)[0]
    s = rng.dirichlet(np.ones(K + 1), size=M)[:, :K].T
    s[3, :] = 0.0                      # an absent subclone (exercises the <0.04 pruning of the tiling)
    spec = dict(tree=tree, root=2, assign=[int(x) for x in rng.integers(0, K, 40)], vaf=B.dot(s).tolist(),
                clone_proportions=s.tolist(), totalscore=-1234.5678901234567, sample_names=["S1", "S2", "tumorB"],
                sample_colors=["red", "firebrick4", "brown4", "darkgoldenrod4"],
                cluster_colors=["plum", "lightyellow", "palegreen", "pink", "paleturquoise", "tan", "moccasin"])
    t = dict(tree=np.array(spec["tree"]), root=spec["root"], assign=np.array(spec["assign"]), vaf=np.array(spec["vaf"]),
             clone_proportions=np.array(spec["clone_proportions"]), totalscore=spec["totalscore"],
             sample_names=spec["sample_names"])
    ref.tree_io.write_dot_files(t, spec["sample_colors"], spec["cluster_colors"],
                                os.path.join(HERE, "out_subclones.dot"), os.path.join(HERE, "out_samples.dot"))
    ref.tree_io.write_text_output([t, t], os.path.join(HERE, "out_text.txt"))
    dump("out_spec.json", spec)


def gen_outputs_more():
    """round 2: the reference's writers on a spread of tree dicts (K = 1 .. 8 clones, 1 .. 6 samples, random trees with
    any root, absent subclones, the colour lists of bamse.py:59-61) -> tests/golden/outputs_more.json holds the three
    files of every case.  asciitree's drawer is ours, as in gen_outputs."""
    import tempfile
    sys.path.insert(0, ROOT)
    from bamse_b200.tree_io import LeftAligned
    ref.tree_io.LeftAligned = LeftAligned
    cluster_colors = ["plum", "lightyellow", "palegreen", "pink", "paleturquoise", "tan", "moccasin", "lightcyan3", "oldlace",
                      "palevioletred", "olivedrab1", "lightgray", "steelblue1"]
    sample_colors = ["red", "firebrick4", "brown4", "darkgoldenrod4", "deeppink1", "darkgreen", "cyan4", "darkorchid", "blue", "dimgray"]
    rng = np.random.default_rng(77)
    cases, failed = [], []
    for K, M in [(1, 1), (1, 3), (2, 2), (3, 1), (3, 4), (4, 2), (4, 6), (5, 3), (5, 5), (6, 2), (7, 3), (7, 6), (8, 4), (8, 1)]:
        tree = random_tree(rng, K)
        B = tt.get_matrix(tree)[0]
        sm = rng.dirichlet(np.ones(K + 1), size=M)[:, :K].T
        if K >= 3:
            sm[int(rng.integers(0, K)), :] = 0.0
        if K >= 5:
            sm[int(rng.integers(0, K)), int(rng.integers(0, M))] = 0.0
        spec = dict(tree=tree, root=tree.index(-1), assign=[int(x) for x in rng.integers(0, K, 10 * K)], vaf=B.dot(sm).tolist(),
                    clone_proportions=sm.tolist(), totalscore=float(-rng.uniform(10, 1e5)),
                    sample_names=["S%d" % (i + 1) for i in range(M)], sample_colors=sample_colors, cluster_colors=cluster_colors)
        t = dict(tree=np.array(spec["tree"]), root=spec["root"], assign=np.array(spec["assign"]), vaf=np.array(spec["vaf"]),
                 clone_proportions=np.array(spec["clone_proportions"]), totalscore=spec["totalscore"], sample_names=spec["sample_names"])
        with tempfile.TemporaryDirectory() as d:
            a, b, c = (os.path.join(d, n) for n in ("t.dot", "s.dot", "o.txt"))
            try:
                ref.tree_io.write_dot_files(t, sample_colors, cluster_colors, a, b)
                ref.tree_io.write_text_output([t], c)
            except Exception as e:
                failed.append((K, M, repr(e)))
                continue
            cases.append(dict(spec=spec, subclones=open(a).read(), samples=open(b).read(), text=open(c).read()))
    print(len(cases), "cases;", "the reference itself raised on:", failed)
    dump("outputs_more.json", cases)


if __name__ == "__main__":
    which = sys.argv[1:] or ["scores", "vectors", "canon", "constants", "search", "outputs"]
    t0 = time.time()
    if "scores" in which:
        gen_scores()
    if "vectors" in which:
        gen_vectors()
    if "canon" in which:
        gen_canon()
    if "constants" in which:
        gen_constants()
    if "search" in which:
        gen_search()
    if "outputs" in which:
        gen_outputs()
    if "highk" in which:          # round 2 additions: not in the default list (minutes of reference CPU time)
        gen_scores_highk()
    if "search_big" in which:
        gen_search_big()
    if "edges" in which:
        gen_scores_edges()
    if "configs" in which:
        gen_scores_configs()
    if "simdata" in which:
        gen_simdata()
    if "candidates" in which:
        gen_candidates()
    if "remove_zeros" in which:
        gen_remove_zeros()
    if "outputs_more" in which:
        gen_outputs_more()
    print("done in %.1fs" % (time.time() - t0))
