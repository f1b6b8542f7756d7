"""CPU oracle for the reference's own orchestration on the hot path.

TEST INFRASTRUCTURE ONLY -- see the header of ``oracle/gpy_oracle.py`` (the GP arithmetic
underneath is a restatement of GPy 1.9.9, which cannot be run here: that part is unpinned).
THIS module's part IS pinned: tests/golden/reference_runs.npz holds what the reference's own
classes compute when /root/reference/src is executed unmodified over the same GP arithmetic
(tests/golden/make_reference_run_golden.py), and tests/test_oracle.py requires this restatement
to reproduce those runs.  The parts below follow the *reference's* code, cited line by line:

  A1  delay augmentation        src/MFDataFusion.py:177-208,
                                src/augm_iterators/backward_augm_iterator.py:20-37,
                                src/augm_iterators/even_augm_iterator.py:20-48
  A5  fit                       src/MFDataFusion.py:75-100
  A6  predict                   src/MFDataFusion.py:141-156
  A7  low-fidelity level        src/abstractMFGP.py:82-106
  A8  Monte-Carlo propagation   (extension; README.md:13 -> Perdikaris et al. 2017, sec. 2(c))
  A9  acquisition               src/abstractMFGP.py:124-129,
                                src/adaptation_maximizers/scipydirect_wrapper.py:16-31
  A10 adaptation loop           src/abstractMFGP.py:317-359
"""
import numpy as np

from . import gpy_oracle as go


# --------------------------------------------------------------------------------------
# A1: delay offsets and augmentation
# --------------------------------------------------------------------------------------
def backward_offsets(n, dim):
    """Offsets 0, -e_1..-e_dim, -2e_1.., ..., -n e_dim  (backward_augm_iterator.py:20-37)."""
    out = [np.zeros(dim)]
    for i in range(1, n + 1):
        for k in range(dim):
            v = np.zeros(dim)
            v[k] = -i
            out.append(v)
    return np.array(out)


def even_offsets(n, dim):
    """Offsets 0, then for i=1..n: -i e_1..-i e_dim, +i e_1..+i e_dim
    (even_augm_iterator.py:20-48)."""
    out = [np.zeros(dim)]
    for i in range(1, n + 1):
        for sign in (-1.0, 1.0):
            for k in range(dim):
                v = np.zeros(dim)
                v[k] = sign * i
                out.append(v)
    return np.array(out)


def augment(X, offsets, tau, f_low_batched):
    """X_aug[i] = [x_i, f_low(x_i + o_0 tau), ..., f_low(x_i + o_{E-1} tau)]
    (src/MFDataFusion.py:193-206).  ``f_low_batched`` maps (rows, d) -> (rows, 1)."""
    M, d = X.shape
    E = offsets.shape[0]
    loc = X[:, None, :] + offsets[None, :, :] * tau               # (M, E, d)  :193
    vals = f_low_batched(loc.reshape(M * E, d)).reshape(M, E)     # :197-201
    return np.concatenate([X, vals], axis=1)                      # :204


# --------------------------------------------------------------------------------------
# the model
# --------------------------------------------------------------------------------------
class OracleMFGP:
    """CPU restatement of MultifidelityDataFusion (src/MFDataFusion.py:56-208) without
    plotting.  ``maximizer(predict, lb, ub) -> (x, fopt)`` mirrors AbstractMaximizer."""

    def __init__(self, input_dim, num_derivatives, tau, f_exact, f_low=None, lf_X=None,
                 lf_Y=None, use_composite_kernel=True, add_noise=False, lower_bound=None,
                 upper_bound=None, form="gpy", lf_theta=None, rng=None, offsets=None):
        self.input_dim = input_dim
        self.tau = tau
        self.f_exact = f_exact
        self.offsets = backward_offsets(num_derivatives, input_dim)   # MFDataFusion.py:67
        if offsets is not None:                                        # e.g. even_offsets(n, dim)
            self.offsets = np.asarray(offsets, dtype=np.float64)
        self.kind = go.KIND_COMPOSITE if use_composite_kernel else go.KIND_RBF
        P = 7 if use_composite_kernel else 3
        self.theta = np.ones(P)            # kernel object persists across fits (:96 warm start)
        self.add_noise = add_noise
        self.form = form
        self.rng = rng
        self.lower_bound = np.zeros(input_dim) if lower_bound is None else lower_bound
        self.upper_bound = np.ones(input_dim) if upper_bound is None else upper_bound
        assert (f_low is not None) ^ (lf_X is not None and lf_Y is not None)   # abstractMFGP.py:93-95
        self.data_driven_lf_approach = f_low is None
        if self.data_driven_lf_approach:
            self.lf_model = go.OracleGPRegression(lf_X, lf_Y, go.KIND_RBF, form=form)
            if lf_theta is None:
                self.lf_model.optimize()                              # abstractMFGP.py:103
            else:
                self.lf_model.theta = np.array(lf_theta, dtype=np.float64)
            self.f_low = lambda t: self.lf_model.predict(t)[0]        # abstractMFGP.py:104
        else:
            self.f_low = f_low

    def augment(self, X):
        return augment(X, self.offsets, self.tau, self.f_low)

    def fit(self, hf_X, theta=None, num_restarts=6):
        """theta given -> skip the optimiser (parity at fixed theta)."""
        assert hf_X.ndim == 2 and hf_X.shape[1] == self.input_dim
        self.hf_X = hf_X
        self.hf_Y = self.f_exact(hf_X)
        assert self.hf_Y.shape == (hf_X.shape[0], 1)
        self.hf_model = go.OracleGPRegression(self.augment(hf_X), self.hf_Y, self.kind,
                                              d=self.input_dim, theta=self.theta, form=self.form)
        if theta is None:
            go.ard_recipe(self.hf_model, num_restarts, rng=self.rng)
        else:
            self.hf_model.theta = np.array(theta, dtype=np.float64)
        self.theta = self.hf_model.theta.copy()

    def predict(self, X_test):
        assert X_test.ndim == 2 and X_test.shape[1] == self.input_dim
        Xa = self.augment(X_test)
        if self.add_noise:                                            # MFDataFusion.py:154-155
            self.hf_model.theta[-1] = 1e-6
            self.hf_model._post = None
        return self.hf_model.predict(Xa)

    def get_mse(self, X_test, Y_test):
        return float(np.mean((Y_test - self.predict(X_test)[0]) ** 2))

    def adapt(self, adapt_steps, maximizer, eps=1e-8, fit_kwargs=None):
        """adapt_and_plot loop body (src/abstractMFGP.py:317-359), plotting removed."""
        fit_kwargs = fit_kwargs or {}
        acquired = []
        for i in range(adapt_steps):
            x, fopt = maximizer(self.predict, self.lower_bound, self.upper_bound)
            acquired.append((np.array(x), float(fopt)))
            self.fit(np.vstack((self.hf_X, x)), **fit_kwargs)
            if np.abs(fopt) < eps:
                break
        return acquired

    # -- A8: Monte-Carlo propagation of the LF posterior ------------------------------
    def lf_marginals(self, X_test, include_noise=True, jitter=0.0):
        """LF posterior at the E augmented locations of every test point:
        mu (M,E), cov (M,E,E).  The E x E block is the LF GP's joint predictive
        covariance at x_i + o_k tau (noise on its diagonal when include_noise)."""
        assert self.data_driven_lf_approach
        lf = self.lf_model
        M, d = X_test.shape
        E = self.offsets.shape[0]
        loc = (X_test[:, None, :] + self.offsets[None, :, :] * self.tau).reshape(M * E, d)
        L, alpha = lf.posterior()
        theta_k, noise = go.split_theta(lf.kind, lf.theta)
        mu, var, tmp = go.posterior_predict(lf.kind, lf.X, lf.d, lf.theta, L, alpha, loc,
                                            include_noise=False, form=self.form, return_tmp=True)
        mu = mu.reshape(M, E)
        tmp = tmp.T.reshape(M, E, -1)                                  # (M, E, N_l)
        loc = loc.reshape(M, E, d)
        cov = np.empty((M, E, E))
        for a in range(E):
            for b in range(E):
                r2 = np.sum((loc[:, a, :] - loc[:, b, :]) ** 2, axis=1)
                kab = theta_k[0] * np.exp(-0.5 * r2 / theta_k[1] ** 2)
                cov[:, a, b] = kab - np.sum(tmp[:, a, :] * tmp[:, b, :], axis=1)
        idx = np.arange(E)
        cov[:, idx, idx] = np.clip(cov[:, idx, idx], go.VAR_CLIP, np.inf)
        if include_noise:
            cov[:, idx, idx] += noise
        cov[:, idx, idx] += jitter
        return mu, cov

    def predict_mc(self, X_test, eps, include_lf_noise=True, jitter=0.0, return_samples=False):
        """eps: (M, S, E) standard normals.  z_s = mu_l + chol(cov_l) eps_s;
        (mu_s, v_s) = HF predict at [x, z_s];  mean = mean_s mu_s;
        var = mean_s v_s + var_s(mu_s) (ddof=0).  S=1, eps=0 reproduces predict()."""
        M, S, E = eps.shape
        assert E == self.offsets.shape[0]
        mu_l, cov_l = self.lf_marginals(X_test, include_lf_noise, jitter)
        C = np.linalg.cholesky(cov_l) if E > 1 else np.sqrt(cov_l)
        z = mu_l[:, None, :] + np.einsum("mab,msb->msa", C, eps)       # (M, S, E)
        Xa = np.concatenate([np.repeat(X_test[:, None, :], S, axis=1), z], axis=2)
        if self.add_noise:
            self.hf_model.theta[-1] = 1e-6
            self.hf_model._post = None
        mu_s, v_s = self.hf_model.predict(Xa.reshape(M * S, -1))
        mu_s = mu_s.reshape(M, S)
        v_s = v_s.reshape(M, S)
        mean = mu_s.mean(axis=1, keepdims=True)
        var = v_s.mean(axis=1, keepdims=True) + mu_s.var(axis=1, keepdims=True)
        if return_samples:
            return mean, var, mu_s, v_s
        return mean, var


    def predict_mc_joint(self, X_test, eps, include_lf_noise=True, jitter=0.0):
        """MC propagation with the LF posterior sampled JOINTLY across the test points (E = 1):
        Sigma = K_l(X*, X*) - tmp^T tmp (+ noise I, + jitter I), z_s = mu_l + chol(Sigma) eps[:, s];
        eps (M, S).  Returns (mean (M,1), var (M,1), mu_s (M,S))."""
        assert self.offsets.shape[0] == 1
        lf = self.lf_model
        L, alpha = lf.posterior()
        theta_k, noise = go.split_theta(lf.kind, lf.theta)
        mu, _, tmp = go.posterior_predict(lf.kind, lf.X, lf.d, lf.theta, L, alpha, X_test,
                                          include_noise=False, form=self.form, return_tmp=True)
        Sigma = go.kernel_K(lf.kind, X_test, None, lf.d, theta_k, "direct") - tmp.T.dot(tmp)
        Sigma[np.diag_indices(X_test.shape[0])] += (noise if include_lf_noise else 0.0) + jitter
        z = mu + np.linalg.cholesky(Sigma).dot(eps)                         # (M, S)
        M, S = eps.shape
        Xa = np.concatenate([np.repeat(X_test[:, None, :], S, axis=1), z[:, :, None]], axis=2)
        mu_s, v_s = self.hf_model.predict(Xa.reshape(M * S, -1))
        mu_s, v_s = mu_s.reshape(M, S), v_s.reshape(M, S)
        return (mu_s.mean(axis=1, keepdims=True), v_s.mean(axis=1, keepdims=True) + mu_s.var(axis=1, keepdims=True),
                mu_s)


def predict_mc_chain(levels, X_test, eps, include_lower_noise=True):
    """MC propagation through a chain of GP levels (recursive NARGP; Perdikaris et al. 2017, sec. 2(c)):
    levels[0] a GP on x, levels[t >= 1] GPs on [x, z].  eps (L-1, M, S) standard normals:
    z_1 = mu_0 + sd_0 eps_1; (mu_t, v_t) = level t at [x, z_t]; z_{t+1} = mu_t + sqrt(v_t) eps_{t+1};
    mean = mean_s mu_top, var = mean_s v_top + var_s(mu_top).  levels: OracleGPRegression objects."""
    Lm1, M, S = eps.shape
    assert Lm1 == len(levels) - 1
    mu, v = levels[0].predict(X_test, include_noise=include_lower_noise)     # (M,1)
    z = mu + np.sqrt(v) * eps[0]                                             # (M,S)
    Xrep = np.repeat(X_test[:, None, :], S, axis=1)
    for t in range(1, len(levels)):
        top = t == len(levels) - 1
        Xa = np.concatenate([Xrep, z[:, :, None]], axis=2).reshape(M * S, -1)
        mu_s, v_s = levels[t].predict(Xa, include_noise=True if top else include_lower_noise)
        mu_s, v_s = mu_s.reshape(M, S), v_s.reshape(M, S)
        if not top:
            z = mu_s + np.sqrt(v_s) * eps[t]
    return mu_s.mean(axis=1, keepdims=True), v_s.mean(axis=1, keepdims=True) + mu_s.var(axis=1, keepdims=True)


# --------------------------------------------------------------------------------------
# A9: candidate-set acquisition
# --------------------------------------------------------------------------------------
def candidate_argmax(predict, candidates):
    """i* = argmax_i var(c_i), lowest index on ties (np.argmax).  Returns
    (i*, c_{i*}, -var_{i*}, top-2 relative gap)."""
    var = predict(candidates)[1].ravel()
    i = int(np.argmax(var))
    if var.size > 1:
        second = np.partition(var, -2)[-2]
        gap = (var[i] - second) / abs(var[i]) if var[i] != 0 else 0.0
    else:
        gap = np.inf
    return i, candidates[i].copy(), -float(var[i]), float(gap)


def make_candidate_maximizer(candidates):
    def maximizer(predict, lb, ub):
        i, x, fopt, _ = candidate_argmax(predict, candidates)
        return x, fopt
    return maximizer


# --------------------------------------------------------------------------------------
# PCE mean consumer (src/gpc/chaospy_wrapper.py:13,19-24): tensor Gauss-Legendre nodes
# --------------------------------------------------------------------------------------
def gauss_legendre_grid(order, dim, lower=0.0, upper=1.0):
    """(order+1)^dim tensor Gauss-Legendre nodes/weights for U[lower,upper]^dim
    (weights sum to 1), node ordering = C-order over the per-dimension nodes."""
    x, w = np.polynomial.legendre.leggauss(order + 1)
    x = 0.5 * (upper - lower) * (x + 1.0) + lower
    w = 0.5 * w
    grids = np.meshgrid(*([x] * dim), indexing="ij")
    nodes = np.stack([g.ravel() for g in grids], axis=1)
    wg = np.meshgrid(*([w] * dim), indexing="ij")
    weights = np.prod(np.stack([g.ravel() for g in wg], axis=1), axis=1)
    return nodes, weights
