"""CPU oracle for the GP engine the reference delegates to (GPy 1.9.9 / paramz 0.9.5).

TEST INFRASTRUCTURE ONLY.  Nothing under ``multifidelity_datafusion_gps_b200/`` may
import this module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and only as the checker / the timed
CPU baseline.

PARITY UNPINNED.  The reference (``/root/reference``) owns no arithmetic on this path:
it calls ``GPy.models.GPRegression`` (``src/MFDataFusion.py:93-100``,
``src/abstractMFGP.py:100-104,131-137``).  GPy 1.9.9 (``requirements.txt:9``) and
paramz 0.9.5 (``requirements.txt:23``) are not vendored in the reference, not installed
in this image and not installable (no network), and the reference's tests hold no
golden vectors for this boundary.  This file therefore restates
GPy's *published algorithm* from memory of the upstream sources named below; it is
validated by finite differences, closed-form 1-/2-point GP cases and a second,
independently written direct-solve formulation and -- since round 2 -- scikit-learn's exact
GP as a third-party implementation of the same mathematics (LML, every hyper-parameter gradient,
predictive mean and variance, RBF and composite kernel; ``tests/test_oracle.py``), not by the
reference's own outputs: GPy's constants (the 1e-8 added to the diagonal, the jitchol schedule,
the 1e-15 variance clip) remain recalled.  (What the reference itself owns -- the orchestration around this
engine -- IS pinned by executing its code over this module: see ``oracle/mfgp_oracle.py``.)

Upstream files followed, by name (GPy 1.9.9):
  GPy/kern/src/stationary.py   Stationary._unscaled_dist / _scaled_dist / K /
                               update_gradients_full, dK_dr_via_X
  GPy/kern/src/rbf.py          RBF.K_of_r = variance * exp(-0.5 r^2); dK_dr = -r K
  GPy/kern/src/prod.py         Prod.K = prod of parts; gradient via dL_dK * other part
  GPy/kern/src/add.py          Add.K = sum of parts; every part sees dL_dK
  GPy/likelihoods/gaussian.py  Gaussian.exact_inference_gradients = sum(diag dL_dK);
                               predictive_values adds the noise variance
  GPy/inference/latent_function_inference/exact_gaussian_inference.py
                               Ky = K + (variance + 1e-8) I; pdinv; dpotrs; LML; dL_dK
  GPy/inference/latent_function_inference/posterior.py
                               Posterior._raw_predict (woodbury_vector / woodbury_chol path)
  GPy/util/linalg.py           jitchol, pdinv, dpotrs, dpotri, dtrtrs, tdot
paramz 0.9.5:
  paramz/transformations.py    Logexp (f, finv, gradfactor)
  paramz/model.py              optimize, optimize_restarts, _objective_grads
  paramz/optimization/optimization.py   opt_lbfgsb -> scipy.optimize.fmin_l_bfgs_b
  paramz/core/parameter_core.py         randomize (N(0,1) in the transformed space)
"""
import numpy as np
from scipy import linalg as sla
from scipy.linalg import lapack
from scipy import optimize as sopt

LOG_2_PI = np.log(2.0 * np.pi)
JITTER_CONST = 1e-8          # exact_gaussian_inference.py: diag.add(Ky, variance + 1e-8)
VAR_CLIP = 1e-15             # posterior.py: var = np.clip(var, 1e-15, np.inf)   [GPy-recall]

KIND_RBF = 0                 # GPy.kern.RBF(D)             (src/abstractMFGP.py:59-60)
KIND_COMPOSITE = 1           # RBF(aug)*RBF(x) + RBF(x)    (src/abstractMFGP.py:62-80)


# --------------------------------------------------------------------------------------
# kernels
# --------------------------------------------------------------------------------------
def unscaled_dist(X, X2=None, form="gpy"):
    """Euclidean distance matrix.

    form="gpy": GPy Stationary._unscaled_dist -- expansion |x|^2+|y|^2-2x.y, diagonal
    forced to zero for the symmetric case, clipped at zero, then sqrt.
    form="direct": sqrt(sum (x-y)^2) -- what the CUDA kernels compute (more accurate).
    """
    if form == "direct":
        Y = X if X2 is None else X2
        r2 = np.zeros((X.shape[0], Y.shape[0]))
        for k in range(X.shape[1]):
            diff = X[:, k][:, None] - Y[:, k][None, :]
            r2 += diff * diff
        return np.sqrt(r2)
    if X2 is None:
        Xsq = np.sum(np.square(X), 1)
        r2 = -2.0 * X.dot(X.T) + (Xsq[:, None] + Xsq[None, :])
        r2[np.diag_indices(X.shape[0])] = 0.0
        r2 = np.clip(r2, 0, np.inf)
        return np.sqrt(r2)
    X1sq = np.sum(np.square(X), 1)
    X2sq = np.sum(np.square(X2), 1)
    r2 = -2.0 * X.dot(X2.T) + (X1sq[:, None] + X2sq[None, :])
    r2 = np.clip(r2, 0, np.inf)
    return np.sqrt(r2)


def rbf_K(X, X2, variance, lengthscale, form="gpy"):
    """GPy RBF (non-ARD): variance * exp(-0.5 * (dist/lengthscale)^2)."""
    r = unscaled_dist(X, X2, form) / lengthscale
    return variance * np.exp(-0.5 * r ** 2), r


def split_theta(kind, theta):
    """theta layout (SURVEY.md section 8c param order):
    composite: [s1, l1 | s2, l2 | s3, l3 | noise], rbf: [s, l | noise]."""
    theta = np.asarray(theta, dtype=np.float64)
    if kind == KIND_COMPOSITE:
        assert theta.shape == (7,)
    else:
        assert theta.shape == (3,)
    return theta[:-1], theta[-1]


def kernel_K(kind, X, X2, d, theta_k, form="gpy", want_parts=False):
    """Covariance of the reference's two kernels.

    kind=COMPOSITE: k1(z,z')*k2(x,x') + k3(x,x'), x = first d columns, z = the rest
    (src/abstractMFGP.py:72-80); kind=RBF: one RBF over all columns (src/abstractMFGP.py:59-60).
    """
    if kind == KIND_RBF:
        K, r = rbf_K(X, X2, theta_k[0], theta_k[1], form)
        return (K, dict(K1=K, r1=r)) if want_parts else K
    Xx, Xz = X[:, :d], X[:, d:]
    X2x, X2z = (None, None) if X2 is None else (X2[:, :d], X2[:, d:])
    K1, r1 = rbf_K(Xz, X2z, theta_k[0], theta_k[1], form)
    K2, r2 = rbf_K(Xx, X2x, theta_k[2], theta_k[3], form)
    K3, r3 = rbf_K(Xx, X2x, theta_k[4], theta_k[5], form)
    K = K1 * K2 + K3
    if want_parts:
        return K, dict(K1=K1, K2=K2, K3=K3, r1=r1, r2=r2, r3=r3)
    return K


def kernel_Kdiag(kind, theta_k, n):
    if kind == KIND_RBF:
        return np.full(n, theta_k[0])
    return np.full(n, theta_k[0] * theta_k[2] + theta_k[4])


def assemble_Ky(kind, X, d, theta, form="gpy"):
    """K + (noise + 1e-8) I  -- what ExactGaussianInference factorises."""
    theta_k, noise = split_theta(kind, theta)
    K = kernel_K(kind, X, None, d, theta_k, form)
    Ky = K.copy()
    Ky[np.diag_indices(K.shape[0])] += noise + JITTER_CONST
    return Ky


# --------------------------------------------------------------------------------------
# linear algebra the way GPy.util.linalg does it
# --------------------------------------------------------------------------------------
class NotPD(np.linalg.LinAlgError):
    pass


def jitchol(A, maxtries=5):
    """GPy.util.linalg.jitchol: dpotrf(lower=1); on failure add jitter
    mean(diag)*1e-6*10^k, k = 0..maxtries-1.  Returns (L, jitter_used)."""
    A = np.ascontiguousarray(A)
    L, info = lapack.dpotrf(A, lower=1)
    if info == 0:
        return L, 0.0
    diagA = np.diag(A)
    if np.any(diagA <= 0.0):
        raise NotPD("not pd: non-positive diagonal elements")
    jitter = diagA.mean() * 1e-6
    for _ in range(maxtries):
        L, info = lapack.dpotrf(A + np.eye(A.shape[0]) * jitter, lower=1)
        if info == 0:
            return L, jitter
        jitter *= 10.0
    raise NotPD("not positive definite, even with jitter.")


def inference(kind, X, Y, d, theta, form="gpy", want_grad=True):
    """ExactGaussianInference.inference + kernel/likelihood gradient reductions.

    Returns dict(lml, grad (P,), L, alpha, Ki, jitter).  grad is d LML / d theta in the
    *untransformed* parameter space, ordered as theta.
    """
    theta_k, noise = split_theta(kind, theta)
    N = X.shape[0]
    K, parts = kernel_K(kind, X, None, d, theta_k, form, want_parts=True)
    Ky = K.copy()
    Ky[np.diag_indices(N)] += noise + JITTER_CONST
    L, jitter = jitchol(Ky)
    alpha, _ = lapack.dpotrs(L, Y, lower=1)
    logdet = 2.0 * np.sum(np.log(np.diag(L)))
    lml = 0.5 * (-Y.size * LOG_2_PI - Y.shape[1] * logdet - np.sum(alpha * Y))
    out = dict(lml=float(lml), L=np.tril(L), alpha=alpha, jitter=jitter)
    if not want_grad:
        return out
    Ki, _ = lapack.dpotri(L, lower=1)
    Ki = np.tril(Ki) + np.tril(Ki, -1).T
    dL_dK = 0.5 * (alpha.dot(alpha.T) - Y.shape[1] * Ki)
    g = np.zeros(len(theta))

    def rbf_grads(Kpart, r, G, variance, lengthscale):
        gv = np.sum(Kpart * G) / variance
        dL_dr = (-r * Kpart) * G
        gl = -np.sum(dL_dr * r) / lengthscale
        return gv, gl

    if kind == KIND_RBF:
        g[0], g[1] = rbf_grads(parts["K1"], parts["r1"], dL_dK, theta_k[0], theta_k[1])
    else:
        K1, K2, K3 = parts["K1"], parts["K2"], parts["K3"]
        g[0], g[1] = rbf_grads(K1, parts["r1"], dL_dK * K2, theta_k[0], theta_k[1])
        g[2], g[3] = rbf_grads(K2, parts["r2"], dL_dK * K1, theta_k[2], theta_k[3])
        g[4], g[5] = rbf_grads(K3, parts["r3"], dL_dK, theta_k[4], theta_k[5])
    g[-1] = np.trace(dL_dK)
    out.update(grad=g, Ki=Ki)
    return out


def posterior_predict(kind, X, d, theta, L, alpha, Xnew, include_noise=True, form="gpy",
                      return_tmp=False):
    """GP.predict -> Posterior._raw_predict (woodbury_chol path) + Gaussian.predictive_values.

    mu = Kx^T alpha;  var = Kxx - sum((L^-1 Kx)^2, 0), clipped at 1e-15, + noise.
    Returns (mu (M,1), var (M,1)).
    """
    theta_k, noise = split_theta(kind, theta)
    Kx = kernel_K(kind, X, Xnew, d, theta_k, form)               # (N, M)
    mu = Kx.T.dot(alpha)
    tmp, _ = lapack.dtrtrs(L, Kx, lower=1)
    var = (kernel_Kdiag(kind, theta_k, Xnew.shape[0]) - np.square(tmp).sum(0))[:, None]
    var = np.clip(var, VAR_CLIP, np.inf)
    if include_noise:
        var = var + noise
    if return_tmp:
        return mu, var, tmp
    return mu, var


# --------------------------------------------------------------------------------------
# paramz: transformed parameter space and the optimiser recipe
# --------------------------------------------------------------------------------------
_LIM_VAL = 36.0
_LOG_LIM_VAL = np.log(np.finfo(np.float64).max)


def logexp_f(x):
    x = np.asarray(x, dtype=np.float64)
    return np.where(x > _LIM_VAL, x, np.log1p(np.exp(np.clip(x, -_LOG_LIM_VAL, _LIM_VAL))))


def logexp_finv(f):
    f = np.asarray(f, dtype=np.float64)
    return np.where(f > _LIM_VAL, f, np.log(np.expm1(f)))


def logexp_gradfactor(f):
    f = np.asarray(f, dtype=np.float64)
    return np.where(f > _LIM_VAL, 1.0, -np.expm1(-f))


class OracleGPRegression:
    """Minimal stand-in for GPy.models.GPRegression as the reference uses it.

    Parameters live in ``theta`` (untransformed); ``fixed`` masks parameters removed from
    the optimiser array (``.fix()``).  ``optimize`` / ``optimize_restarts`` follow paramz.
    """

    def __init__(self, X, Y, kind=KIND_RBF, d=None, theta=None, form="gpy"):
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.Y = np.ascontiguousarray(Y, dtype=np.float64)
        assert self.Y.ndim == 2 and self.Y.shape[1] == 1
        self.kind = kind
        self.d = self.X.shape[1] if d is None else d
        P = 7 if kind == KIND_COMPOSITE else 3
        self.theta = np.ones(P) if theta is None else np.array(theta, dtype=np.float64)
        self.fixed = np.zeros(P, dtype=bool)
        self.form = form
        self._fail_count = 0
        self.n_evals = 0
        self._post = None

    # -- inference -------------------------------------------------------------------
    def _infer(self, want_grad=True):
        return inference(self.kind, self.X, self.Y, self.d, self.theta, self.form, want_grad)

    def log_likelihood(self):
        return self._infer(False)["lml"]

    def objective_and_grad(self, x):
        """paramz Model._objective_grads: x is the transformed, un-fixed array."""
        free = ~self.fixed
        try:
            self.theta[free] = logexp_f(x)
            res = self._infer(True)
            self._fail_count = 0
            self.n_evals += 1
            grads = -(res["grad"][free] * logexp_gradfactor(self.theta[free]))
            return -res["lml"], grads
        except (np.linalg.LinAlgError, ZeroDivisionError, ValueError):
            if self._fail_count >= 10:
                raise
            self._fail_count += 1
            return np.inf, np.zeros(int(free.sum()))

    def optimize(self, max_iters=1000):
        free = ~self.fixed
        x0 = logexp_finv(self.theta[free])
        x_opt, f_opt, info = sopt.fmin_l_bfgs_b(self.objective_and_grad, x0,
                                               maxfun=max_iters, maxiter=max_iters)
        # paramz opt_lbfgsb.opt: self.f_opt = f_fp(self.x_opt)[0] -- re-evaluated at the returned point; this
        # is the value optimize_restarts compares  [paramz-recall]
        f_opt = self.objective_and_grad(x_opt)[0]
        self.theta[free] = logexp_f(x_opt)
        self._post = None
        return x_opt, f_opt

    def optimize_restarts(self, num_restarts, max_iters=1000, rng=None):
        rng = np.random if rng is None else rng
        free = ~self.fixed
        runs = []
        for i in range(num_restarts):
            try:
                if i > 0:
                    self.theta[free] = logexp_f(rng.normal(size=int(free.sum())))
                runs.append(self.optimize(max_iters))
            except Exception:                                   # paramz robust=False re-raises;
                raise                                            # kept explicit on purpose
        best = int(np.argmin([r[1] for r in runs]))
        self.theta[free] = logexp_f(runs[best][0])
        self._post = None
        return runs

    # -- prediction ------------------------------------------------------------------
    def posterior(self):
        if self._post is None:
            res = self._infer(False)
            self._post = (res["L"], res["alpha"])
        return self._post

    def predict(self, Xnew, include_noise=True):
        L, alpha = self.posterior()
        return posterior_predict(self.kind, self.X, self.d, self.theta, L, alpha,
                                 np.ascontiguousarray(Xnew, dtype=np.float64),
                                 include_noise, self.form)


def ard_recipe(model, num_restarts=6, rng=None, max_iters_1=500, max_iters_2=1000):
    """AbstractMFGP.ARD (src/abstractMFGP.py:131-137)."""
    model.theta[-1] = model.Y.var() * 0.01
    model.fixed[-1] = True
    model.optimize(max_iters=max_iters_1)
    model.fixed[-1] = False
    model.optimize_restarts(num_restarts, max_iters=max_iters_2, rng=rng)
