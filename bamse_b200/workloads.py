"""The five BASELINE.json configurations as reproducible synthetic work lists.

A work list is what bamse.py:53-90 hands to the scoring loop: simdata-style read counts
(simdata.py:18-31 restated in bamse_b200/simdata.py) -> KMeans model selection and
`-n` KMeans restarts per candidate cluster count (bamse.py:75-90, including its use of the
argsort *indices* as cluster counts) -> a list of `clustering` objects.  Every labeled
rooted tree (K^(K-1)) of every clustering is then scored.
"""
import numpy as np

from . import simdata
from .clustering import clustering, part_func
from .tree_tools import reorder_assignmens

# name -> (K_true, M samples, N mutations, max_clusts -nc, n_trees -n, seed, error rates, force_K)
# seeds follow SURVEY.md section 8d.  force_K pins the candidate cluster counts where the
# BASELINE.json line itself names the tree space (c4: "-nc 10 full tree enumeration").
CONFIGS = {
    "c1": dict(K_true=4, M=3, N=50, nc=5, n=10, seed=12345, errs=[0.001], top=5,
               desc="3 samples, 50 mutations, -nc 5 -n 10 -t 5, e=0.001"),
    "c2": dict(K_true=6, M=10, N=500, nc=7, n=50, seed=2, errs=[0.001], top=5,
               desc="10 samples, 500 mutations, -nc 7 -n 50"),
    "c3": dict(K_true=7, M=20, N=2000, nc=8, n=100, seed=3, errs=[0.001], top=5,
               desc="20 samples, 2000 mutations, -nc 8 -n 100"),
    "c4": dict(K_true=9, M=5, N=1000, nc=10, n=3, seed=4, errs=[0.001], top=5, force_K=[10],
               desc="5 samples, 1000 mutations, -nc 10 full tree enumeration"),
    "c5": dict(K_true=5, M=10, N=500, nc=6, n=3, seed=5, top=5,
               errs=[float(x) for x in np.logspace(-4, -2, 64)],
               desc="64 (e,s) pairs x 10 samples x 500 mutations, -nc 6"),
}


def unique_Kmeans_clustering(As, Bs, num_clusters, n_init, pf, err, ctx=None):
    """bamse.py:18-27."""
    from sklearn.cluster import KMeans
    seen, result = set(), []
    data = (As + 0.0) / (As + Bs)
    for _ in range(n_init):
        assign = tuple(int(x) for x in reorder_assignmens(KMeans(n_clusters=num_clusters, n_init=1).fit(data).labels_))
        if assign not in seen:
            seen.add(assign)
            result.append(clustering(As, Bs, assign, pf, err, ctx=ctx))
    return result


def candidate_clusterings(As, Bs, max_clusts, n_trees, err, ctx=None, model_select_n_init=100, force_K=None,
                          verbose=False):
    """bamse.py:54-90: model selection over K = 1..max_clusts, then the unique KMeans
    restarts for the selected cluster counts."""
    from sklearn.cluster import KMeans
    N = len(As)
    max_clusts = min(N, max_clusts)
    pf = part_func(N, max_clusts)
    if force_K is not None:
        candidate_numbers = list(force_K)
    else:
        Y = (As + 0.0) / (Bs + As)
        raw = []
        for K in range(1, max_clusts + 1):
            assign = KMeans(n_clusters=K, n_init=model_select_n_init).fit(Y).labels_
            c = clustering(As, Bs, assign, pf, err, ctx=ctx)
            raw.append(c.cluster_prior + c.maxlikelihood)
        raw_scores = np.array(raw)
        if verbose:
            print(raw_scores)
        if np.any(np.diff(raw_scores) < 0):
            candidate_numbers = [int(x) for x in raw_scores.argsort()[-3:]]
        else:
            candidate_numbers = [max_clusts]
        # bamse.py:84,90 passes argsort *indices* (K-1) as cluster counts; a 0 there would
        # make KMeans raise, so the reference effectively never sees it -- skip it.
        candidate_numbers = [k for k in candidate_numbers if k >= 1]
    if verbose:
        print('possible number of clusters:', candidate_numbers)
    clusts = []
    for k in candidate_numbers:
        clusts += unique_Kmeans_clustering(As, Bs, k, n_trees, pf, err, ctx=ctx)
    return clusts, pf


def build(name, ctx=None, max_clusterings=None):
    """Returns dict(name, desc, errs, clusts (for errs[0]), data, n_evals_per_param)."""
    cfg = CONFIGS[name]
    data = simdata.simulate(cfg["K_true"], cfg["M"], cfg["N"], cfg["seed"])
    np.random.seed(cfg["seed"] + 1000)   # sklearn's random_state=None draws from the global RNG
    clusts, pf = candidate_clusterings(data["As"], data["Bs"], cfg["nc"], cfg["n"], cfg["errs"][0], ctx=ctx,
                                       force_K=cfg.get("force_K"))
    if max_clusterings is not None:
        clusts = clusts[:max_clusterings]
    evals = sum(int(c.As.shape[0]) ** (int(c.As.shape[0]) - 1) for c in clusts)
    return dict(name=name, desc=cfg["desc"], errs=list(cfg["errs"]), clusts=clusts, data=data, pf=pf,
                top=cfg["top"], n_evals_per_param=evals, M=cfg["M"])
