"""CPU tests: pin the oracle as far as it can be pinned without GPy.

Parity is UNPINNED against the reference's own outputs (GPy cannot run here); these tests check the
restatement against finite differences, closed forms, an independently written formulation, and the
golden vectors that CAN be produced by running reference code (the delay iterators)."""
import os

import numpy as np
import pytest
from scipy import linalg as sla

from oracle import gpy_oracle as go
from oracle import mfgp_oracle as mo
from tests import util

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("kind,d,E", [(go.KIND_COMPOSITE, 2, 5), (go.KIND_COMPOSITE, 1, 1),
                                      (go.KIND_RBF, 2, 5), (go.KIND_RBF, 3, 0)])
def test_gradient_matches_finite_differences(kind, d, E):
    X, Y, th = util.random_case(0, 40, d, E, kind)
    res = go.inference(kind, X, Y, d, th)
    for i in range(len(th)):
        h = 1e-6 * th[i]
        tp, tm = th.copy(), th.copy()
        tp[i] += h
        tm[i] -= h
        fd = (go.inference(kind, X, Y, d, tp, want_grad=False)["lml"]
              - go.inference(kind, X, Y, d, tm, want_grad=False)["lml"]) / (2 * h)
        assert abs(res["grad"][i] - fd) <= 1e-6 * max(1.0, abs(fd))


def test_closed_form_one_point_gp():
    # one training point: K_y = s + noise + 1e-8; LML and predictive moments in closed form
    s, l, noise, y, x, xs = 1.7, 0.4, 0.05, 0.9, 0.3, 0.55
    X, Y, th = np.array([[x]]), np.array([[y]]), np.array([s, l, noise])
    res = go.inference(go.KIND_RBF, X, Y, 1, th)
    ky = s + noise + 1e-8
    assert np.isclose(res["lml"], -0.5 * (np.log(2 * np.pi) + np.log(ky) + y * y / ky), rtol=1e-13)
    mu, var = go.posterior_predict(go.KIND_RBF, X, 1, th, res["L"], res["alpha"], np.array([[xs]]))
    k = s * np.exp(-0.5 * (x - xs) ** 2 / l ** 2)
    assert np.isclose(mu[0, 0], k * y / ky, rtol=1e-13)
    assert np.isclose(var[0, 0], s - k * k / ky + noise, rtol=1e-12)


@pytest.mark.parametrize("kind", [go.KIND_COMPOSITE, go.KIND_RBF])
def test_against_independent_formulation(kind):
    # second implementation: explicit loops for the kernel, slogdet / solve instead of Cholesky calls
    d, E = 2, 3
    X, Y, th = util.random_case(3, 25, d, E, kind)
    N = X.shape[0]
    K = np.zeros((N, N))
    for i in range(N):
        for j in range(N):
            rx = np.sum((X[i, :d] - X[j, :d]) ** 2)
            rz = np.sum((X[i, d:] - X[j, d:]) ** 2)
            if kind == go.KIND_COMPOSITE:
                K[i, j] = (th[0] * np.exp(-rz / (2 * th[1] ** 2)) * th[2] * np.exp(-rx / (2 * th[3] ** 2))
                           + th[4] * np.exp(-rx / (2 * th[5] ** 2)))
            else:
                K[i, j] = th[0] * np.exp(-(rx + rz) / (2 * th[1] ** 2))
    Ky = K + (th[-1] + 1e-8) * np.eye(N)
    sign, logdet = np.linalg.slogdet(Ky)
    alpha = np.linalg.solve(Ky, Y)
    lml = -0.5 * (N * np.log(2 * np.pi) + logdet + float(Y.T @ alpha))
    res = go.inference(kind, X, Y, d, th)
    assert np.isclose(res["lml"], lml, rtol=1e-11)
    assert util.rel_err(res["alpha"], alpha) < 1e-9
    assert util.rel_err(go.assemble_Ky(kind, X, d, th, form="direct"), Ky) < 1e-14
    assert util.rel_err(go.assemble_Ky(kind, X, d, th, form="gpy"), Ky) < 1e-12
    Xs = np.random.default_rng(5).uniform(size=(7, d + E))
    mu, var = go.posterior_predict(kind, X, d, th, res["L"], res["alpha"], Xs)
    for m in range(7):
        ks = np.array([go.kernel_K(kind, X[i:i + 1], Xs[m:m + 1], d, th[:-1])[0, 0] for i in range(N)])
        assert np.isclose(mu[m, 0], ks @ alpha[:, 0], rtol=1e-9, atol=1e-12)
        v = go.kernel_Kdiag(kind, th[:-1], 1)[0] - ks @ np.linalg.solve(Ky, ks) + th[-1]
        assert np.isclose(var[m, 0], v, rtol=1e-8)


def test_jitchol_retry_schedule():
    # singular matrix (duplicated rows, no noise) -> succeeds only with jitter mean(diag)*1e-6*10^k
    X = np.repeat(np.random.default_rng(0).uniform(size=(4, 2)), 3, axis=0)
    K = go.kernel_K(go.KIND_RBF, X, None, 2, np.array([1.0, 1.0]))
    K -= 1e-9 * np.eye(len(K))
    L, jit = go.jitchol(K)
    assert jit in [np.diag(K).mean() * 1e-6 * 10.0 ** k for k in range(5)]
    assert np.allclose(np.tril(L) @ np.tril(L).T, K + jit * np.eye(len(K)), atol=1e-10)
    with pytest.raises(np.linalg.LinAlgError):
        go.jitchol(-np.eye(3))


def test_logexp_transform_roundtrip():
    f = np.array([1e-8, 1e-3, 0.5, 1.0, 20.0, 40.0, 500.0])
    assert np.allclose(go.logexp_f(go.logexp_finv(f)), f, rtol=1e-10)
    x = np.array([-5.0, 0.0, 3.0])
    h = 1e-6
    num = (go.logexp_f(x + h) - go.logexp_f(x - h)) / (2 * h)
    assert np.allclose(go.logexp_gradfactor(go.logexp_f(x)), num, rtol=1e-6)


def test_delay_offsets_match_reference_golden():
    # golden produced by running the reference's iterators (tests/golden/make_iterator_golden.py)
    g = np.load(os.path.join(GOLD, "augm_offsets.npz"))
    for n in range(4):
        for dim in range(1, 5):
            assert np.array_equal(mo.backward_offsets(n, dim), g["backward_n%d_d%d" % (n, dim)])
            assert np.array_equal(mo.even_offsets(n, dim), g["even_n%d_d%d" % (n, dim)])


def test_augmentation_layout():
    X = np.array([[0.2, 0.4], [0.6, 0.8]])
    offs = mo.backward_offsets(2, 2)
    Xa = mo.augment(X, offs, 0.001, util.lf_2d)
    assert Xa.shape == (2, 2 + 5)
    assert np.allclose(Xa[:, :2], X)
    assert np.isclose(Xa[1, 2], util.lf_2d(X[1:2])[0, 0])
    assert np.isclose(Xa[1, 3], util.lf_2d(X[1:2] + np.array([[-0.001, 0.0]]))[0, 0])
    assert np.isclose(Xa[1, 6], util.lf_2d(X[1:2] + np.array([[0.0, -0.002]]))[0, 0])


def _nargp(rng_seed=0, n_l=30, n_h=8, d=1):
    rng = np.random.default_rng(rng_seed)
    lf_X = rng.uniform(size=(n_l, d))
    lf_Y = util.f_low_1d(lf_X)
    hf_X = rng.uniform(size=(n_h, d))
    m = mo.OracleMFGP(d, 0, 0, util.f_high_1d, lf_X=lf_X, lf_Y=lf_Y, lf_theta=[1.0, 0.12, 1e-4])
    m.fit(hf_X, theta=[1.0, 0.8, 1.0, 0.5, 0.1, 0.3, 1e-3])
    return m


def test_mc_with_one_sample_and_zero_noise_is_predict():
    m = _nargp()
    Xt = np.linspace(0, 1, 17)[:, None]
    mean, var = m.predict_mc(Xt, np.zeros((17, 1, 1)))
    mu, v = m.predict(Xt)
    assert np.allclose(mean, mu, rtol=1e-13) and np.allclose(var, v, rtol=1e-13)


def test_mc_moments():
    m = _nargp()
    Xt = np.linspace(0, 1, 5)[:, None]
    eps = np.random.default_rng(1).standard_normal((5, 64, 1))
    mean, var, mu_s, v_s = m.predict_mc(Xt, eps, return_samples=True)
    assert np.allclose(mean[:, 0], mu_s.mean(1))
    assert np.allclose(var[:, 0], v_s.mean(1) + ((mu_s - mu_s.mean(1, keepdims=True)) ** 2).mean(1))


def test_candidate_argmax_lowest_index_on_ties():
    pred = lambda c: (None, np.array([0.1, 0.7, 0.7, 0.3])[:, None])
    i, x, fopt, gap = mo.candidate_argmax(pred, np.arange(8.0).reshape(4, 2))
    assert i == 1 and fopt == -0.7 and gap == 0.0


def test_gauss_legendre_pce_mean_closed_form():
    # closed form of the reference's test function (tests/utils.py:14-17): E prod sin(a_i x_i)
    a = [np.pi] * 4
    nodes, w = mo.gauss_legendre_grid(7, 4)
    assert nodes.shape == (8 ** 4, 4) and np.isclose(w.sum(), 1.0)
    est = np.sum(w * (util.hf_4d(nodes)[:, 0] - 5.0))
    exact = np.prod([(1 - np.cos(ai)) / ai for ai in a])
    assert np.isclose(est, exact, rtol=1e-9)


def test_oracle_fit_recipe_improves_likelihood():
    X, Y, _ = util.random_case(2, 20, 1, 1, go.KIND_COMPOSITE)
    m = go.OracleGPRegression(X, Y, go.KIND_COMPOSITE, d=1)
    before = m.log_likelihood()
    go.ard_recipe(m, 3, rng=np.random.RandomState(0))
    assert m.log_likelihood() > before


def test_pce_oracle_reproduces_reference_closed_forms():
    """oracle/pce_oracle.py against the closed-form mean / variance of prod sin(a_i x_i) on [0,1]^d that
    the reference's own gPC tests are written around (tests/utils.py:14-27, tests/test_mfgp_adapt_2d.py:9)."""
    from oracle import pce_oracle as po
    a2 = [2.2 * np.pi, np.pi]
    m, v, c = po.pce_mean_var(lambda x: np.prod(np.sin(x * np.array(a2)), axis=1), [0, 0], [1, 1], 14, 14)
    assert abs(m - po.analytical_mean(a2)) < 1e-14
    assert abs(v - po.analytical_var(a2)) < 1e-9 * po.analytical_var(a2)
    a4 = [np.pi] * 4                                      # tests/test_mfgp_adapt_4d.py:10,13-15 (constant 5)
    m, v, c = po.pce_mean_var(lambda x: np.prod(np.sin(x * np.pi), axis=1) + 5.0, [0] * 4, [1] * 4, 8, 8)
    assert abs(m - po.analytical_mean(a4, 5.0)) < 1e-12
    assert abs(v - po.analytical_var(a4)) < 1e-4 * po.analytical_var(a4)      # truncation at order 8
    assert c.shape == (495,)                               # C(8 + 4, 4) terms of total degree <= 8
