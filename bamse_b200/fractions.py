"""Maximum-likelihood clone fractions for the few winning trees (the role of the reference's
cvxpy-based solve4 / get_max_values_r2, clustering.py:410-432).  cvxpy is not available
here, so the same concave program is solved per sample with scipy's SLSQP:

    maximise  sum_k As[k] log(q_k) + Bs[k] log(1 - q_k),  q = er + (B x)(1/2 - er)
    s.t.      x >= 0,  0 <= B x <= 1,  x <= profile

Host-side, runs only on the <= t returned trees; not part of the scored hot path.
Parity with the original solver's digits is unpinned.
"""
import numpy as np
from scipy.optimize import minimize


def solve4(B, am, bm, er, profile):
    n = len(B)
    M = am.shape[1]
    res, value = [], 0.0
    for m in range(M):
        a = am[:, m].astype(float)
        b = bm[:, m].astype(float)

        def negll(x):
            q = er + B.dot(x) * (0.5 - er)
            q = np.clip(q, 1e-300, 1 - 1e-16)
            return -(np.sum(a * np.log(q)) + np.sum(b * np.log1p(-q)))

        def grad(x):
            q = er + B.dot(x) * (0.5 - er)
            q = np.clip(q, 1e-300, 1 - 1e-16)
            g = (a / q - b / (1 - q)) * (0.5 - er)
            return -B.T.dot(g)

        cons = [{'type': 'ineq', 'fun': lambda x: B.dot(x), 'jac': lambda x: B},
                {'type': 'ineq', 'fun': lambda x: 1.0 - B.dot(x), 'jac': lambda x: -B}]
        ub = np.asarray(profile[m], dtype=float)
        # start from the per-cluster MLE of the prevalences mapped back through A = B^-1
        prev = np.clip(((a / np.maximum(a + b, 1.0)) - er) / (0.5 - er), 0.0, 1.0)
        x0 = np.clip(np.linalg.solve(B, prev), 0.0, ub)
        scale = max(1.0, float(np.sum(a + b)))
        r = minimize(lambda x: negll(x) / scale, x0, jac=lambda x: grad(x) / scale, method='SLSQP',
                     bounds=[(0.0, float(u)) for u in ub], constraints=cons,
                     options={'maxiter': 500, 'ftol': 1e-14})
        res.append(r.x)
        value += -negll(r.x)
    return {'s': np.array(res).transpose(), 'value': value}


def get_max_values_r2(regular_tree, As, Bs, err):
    B = regular_tree['Bmatrix']
    K, M = As.shape
    result = solve4(B, As, Bs, err, [np.ones(K) for _ in range(M)])
    regular_tree['clone_proportions'] = result['s']
    return regular_tree['clone_proportions']
