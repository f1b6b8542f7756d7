"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the pseudo-spectral PCE the reference obtains from
chaospy 4.2.3 (requirements.txt:5; absent here, so its published algorithm is restated):
``cp.generate_quadrature(q, J(Uniform...), rule="gaussian")`` -> tensor Gauss-Legendre rule with q+1
nodes per dimension; ``cp.generate_expansion(p, dist)`` -> total-degree-p orthogonal (Legendre)
basis; ``cp.fit_quadrature`` -> c_k = sum_q w_q f(x_q) phi_k(x_q) / ||phi_k||^2; ``cp.E`` = c_0,
``cp.Var`` = sum_{k>0} c_k^2 ||phi_k||^2 (src/gpc/chaospy_wrapper.py:12-29).  Written with the
orthonormal basis, for which the norms are 1.  Parity is pinned by the reference's own closed forms
(tests/utils.py:14-27), restated below as analytical_mean / analytical_var.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.
"""
import itertools

import numpy as np


def analytical_mean(a, constant=0):                       # tests/utils.py:14-17
    a = a if isinstance(a, list) else [a]
    return np.prod([((1 - np.cos(a_i)) / a_i) for a_i in a]) + constant


def analytical_var(a):                                     # tests/utils.py:20-27
    a = a if isinstance(a, list) else [a]
    m = analytical_mean(a, constant=0)
    term1 = np.prod([0.5 - (np.sin(2 * a_i) / (4 * a_i)) for a_i in a])
    term2 = m ** 2
    term3 = 2 * m * np.prod([(np.cos(a_i) - 1) / a_i for a_i in a]) * ((-1) ** (len(a) - 1))
    return term1 + term2 + term3


def multi_index(dim, order):
    out = []
    for total in range(order + 1):
        level = [k for k in itertools.product(range(total + 1), repeat=dim) if sum(k) == total]
        out.extend(sorted(level, reverse=True))
    return np.asarray(out, dtype=np.int64).reshape(-1, dim)


def tensor_grid(n_per_dim, lower, upper):
    lower, upper = np.asarray(lower, float), np.asarray(upper, float)
    t, w = np.polynomial.legendre.leggauss(n_per_dim)
    d = len(lower)
    grids = np.meshgrid(*[lower[i] + (t + 1) / 2 * (upper[i] - lower[i]) for i in range(d)], indexing="ij")
    nodes = np.stack([g.ravel() for g in grids], axis=1)
    wgrid = np.meshgrid(*[w / 2] * d, indexing="ij")
    wts = np.prod(np.stack([g.ravel() for g in wgrid], axis=1), axis=1)
    return nodes, wts


def legendre_orthonormal(t, pmax):
    """(len(t), pmax+1): sqrt(2n+1) P_n(t), t in [-1, 1] (NumPy's Legendre series evaluation)."""
    cols = []
    for n in range(pmax + 1):
        c = np.zeros(n + 1)
        c[n] = 1.0
        cols.append(np.sqrt(2 * n + 1.0) * np.polynomial.legendre.legval(t, c))
    return np.stack(cols, axis=1)


def project(nodes, weights, values, lower, upper, mi):
    """c_k = sum_q w_q f_q prod_i phi_{k_i}(x_qi) -- the dense restatement of fit_quadrature."""
    nodes = np.asarray(nodes, float)
    lower, upper = np.asarray(lower, float), np.asarray(upper, float)
    pmax = int(mi.max()) if mi.size else 0
    Q, d = nodes.shape
    phi = np.ones((Q, mi.shape[0]))
    for i in range(d):
        t = 2 * (nodes[:, i] - lower[i]) / (upper[i] - lower[i]) - 1
        L = legendre_orthonormal(t, pmax)
        phi *= L[:, mi[:, i]]
    return phi.T @ (np.asarray(weights).ravel() * np.asarray(values).ravel())


def pce_mean_var(function, lower, upper, polynomial_order, quadrature_order):
    nodes, wts = tensor_grid(quadrature_order + 1, lower, upper)
    mi = multi_index(len(lower), polynomial_order)
    c = project(nodes, wts, np.asarray(function(nodes)).ravel(), lower, upper, mi)
    return c[0], float(np.sum(c[1:] ** 2)), c
