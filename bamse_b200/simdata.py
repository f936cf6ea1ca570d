"""Synthetic read-count generator, restating the reference's sim_data_readcounts
(reference simdata.py:18-31; same draws from numpy's legacy global RNG in the same order)
plus a writer for the tab-separated input format of bamse.py (bamse.dat's header style).

The true tree is drawn from `unlabeled_rooted_trees(K)` -- the same SET of topologies as
AllTrees.txt's T[K], in our own deterministic order (AllTrees.txt itself is not shipped).
"""
import numpy as np

from .tree_tools import get_matrix, unlabeled_rooted_trees


def sim_data_readcounts(K, M, alltrees, num_mut, coverage=200, min_in_clust=2, error=0.003):
    """K subclones, M samples, num_mut mutations at fixed coverage.  `alltrees[K]` is the list
    of candidate topologies (pass None to use unlabeled_rooted_trees)."""
    topologies = alltrees[K] if alltrees is not None else unlabeled_rooted_trees(K)
    ind = np.random.randint(len(topologies))
    true_tree = topologies[ind]
    true_sm = np.array([np.random.dirichlet(np.ones(K + 1))[:K] for m in range(M)])
    B = get_matrix(true_tree)[0]
    true_cm = np.array([B.dot(x) for x in true_sm]).transpose()
    ps = np.random.dirichlet(np.ones(K))
    assign = np.random.choice(range(K), num_mut - min_in_clust * K, p=ps)
    assign = np.concatenate((np.repeat(np.arange(K), 2), assign))
    covs = coverage * np.ones((num_mut, M)).astype(int)
    As = np.random.binomial(covs, true_cm[assign] * (0.5 - error) + error)
    Bs = covs - As
    return {'As': As, 'Bs': Bs, 'vaf': true_cm, 'assign': assign, 'tree': true_tree,
            's': np.array(true_sm).transpose(), 'Bmatrix': B, 'Amatrix': np.linalg.inv(B)}


def simulate(K, M, num_mut, seed, coverage=200, error=0.003):
    np.random.seed(seed)
    return sim_data_readcounts(K, M, None, num_mut, coverage=coverage, error=error)


def write_input_file(path, As, Bs, sample_names=None):
    """mutations x (S<i>.var ..., S<i>.ref ...) tab-separated, as bamse.py:53-72 reads it."""
    N, M = As.shape
    names = sample_names or ["S%d" % (i + 1) for i in range(M)]
    with open(path, "w") as f:
        f.write("\t".join([n + ".var" for n in names] + [n + ".ref" for n in names]) + "\n")
        for i in range(N):
            f.write("\t".join(str(int(x)) for x in list(As[i]) + list(Bs[i])) + "\n")
