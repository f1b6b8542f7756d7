"""GPU parity tests, kernel by kernel, through the C-ABI against the CPU oracle.

Tolerances (BASELINE.json north_star): predictive mean 1e-8 relative, variance and LML 1e-6 relative,
arg-max index bit-exact.  The CUDA kernels compute distances as sum (x-y)^2; the oracle offers that
form ("direct", tight bound) and GPy's expansion form ("gpy", the parity bound)."""
import numpy as np
import pytest
import scipy.linalg as sla

from oracle import gpy_oracle as go
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T(gpu):
    import torch
    from multifidelity_datafusion_gps_b200 import ops

    class NS:
        pass
    ns = NS()
    ns.torch, ns.ops = torch, ops
    ns.dev = "cuda:0"
    ns.up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to("cuda:0")
    return ns


# (kind, d, E): the shapes with compile-time kernels (NARGP 1-D / 4-D, GPDFC and GPDF 2-D with delays, a
# plain 3-D level) and two that take the runtime-width kernels (composite with E = 2; D = 11 > 8)
CASES = [(go.KIND_COMPOSITE, 1, 1), (go.KIND_COMPOSITE, 2, 5), (go.KIND_COMPOSITE, 4, 1),
         (go.KIND_RBF, 2, 5), (go.KIND_RBF, 3, 0), (go.KIND_COMPOSITE, 3, 2), (go.KIND_RBF, 2, 9)]


@pytest.mark.parametrize("kind,d,E", CASES)
@pytest.mark.parametrize("N", [1, 7, 64, 65, 200, 333])
def test_assemble_matches_oracle(T, kind, d, E, N):
    X, _, th = util.random_case(N, N, d, E, kind)
    for uplo in (0, 1):
        K = T.ops.assemble(T.up(X), kind, d, th, uplo=uplo).cpu().numpy()
        ref_d = go.assemble_Ky(kind, X, d, th, form="direct")
        ref_g = go.assemble_Ky(kind, X, d, th, form="gpy")
        if uplo == 0:
            K, ref_d, ref_g = np.tril(K), np.tril(ref_d), np.tril(ref_g)
        assert util.rel_err(K, ref_d) < 5e-15
        assert util.rel_err(K, ref_g) < 1e-11


@pytest.mark.parametrize("theta", [
    [1e6, 0.02, 1e3, 0.05, 1e-8, 0.01, 1e-2],      # huge / tiny variances, length-scales far below the data spacing
    [1e-150, 0.3, 1e-150, 0.3, 1e-300, 0.3, 1e-2],  # variances whose product underflows
    [1.0, 1e-3, 1.0, 1e-3, 1.0, 1e-4, 0.0],         # exponents down to -1e8: every off-diagonal element underflows
])
def test_assemble_elementwise_accuracy_at_extreme_hyperparameters(T, theta):
    """exp2s (fastmath.cuh) against the oracle element by element: relative error where the value is
    representable, < 1e-300 absolute where it underflows (the clamp returns a tiny positive number)."""
    X, _, _ = util.random_case(3, 300, 4, 1, go.KIND_COMPOSITE)
    th = np.array(theta)
    K = np.tril(T.ops.assemble(T.up(X), go.KIND_COMPOSITE, 4, th, uplo=0).cpu().numpy())
    ref = np.tril(go.assemble_Ky(go.KIND_COMPOSITE, X, 4, th, form="direct"))
    assert np.all(np.isfinite(K))
    big = np.abs(ref) > 1e-290
    # argument error |x| * 2^-52 for exponents up to ~700, plus ~2 ulp of the exponential itself
    assert np.max(np.abs(K[big] - ref[big]) / np.abs(ref[big])) < 1e-12
    assert np.max(np.abs(K[~big] - ref[~big]), initial=0.0) < 1e-289


def test_assemble_odd_leading_dimension_and_jitter(T):
    X, _, th = util.random_case(1, 37, 2, 5, go.KIND_COMPOSITE)
    K = T.ops.assemble(T.up(X), go.KIND_COMPOSITE, 2, th, uplo=1, jitter=0.25, ld=41).cpu().numpy()
    ref = go.assemble_Ky(go.KIND_COMPOSITE, X, 2, th, form="direct") + 0.25 * np.eye(37)
    assert util.rel_err(K[:, :37], ref) < 5e-15
    assert np.all(K[:, 37:] == 0.0)


@pytest.mark.parametrize("n", [128, 200, 256, 384, 640, 1000])
def test_potrf_trtri_lauum_match_lapack(T, n):
    rng = np.random.default_rng(n)
    B = rng.standard_normal((n, n))
    A = B @ B.T / n + np.eye(n)
    Ap = T.ops.pad_spd(T.up(A))
    W, info = T.ops.potrf(Ap)
    assert info == 0
    L = np.tril(Ap.cpu().numpy())[:n, :n]
    Lref = sla.cholesky(A, lower=True)
    assert util.rel_err(L, Lref) < 1e-12
    T.ops.trtri(Ap, W)
    Wn = np.tril(W.cpu().numpy())
    assert util.rel_err(Wn[:n, :n], np.linalg.inv(Lref)) < 1e-11
    assert np.allclose(Wn[n:, n:], np.eye(Wn.shape[0] - n))      # identity pad survives
    Kinv = np.tril(T.ops.lauum(W).cpu().numpy())[:n, :n]
    assert util.rel_err(Kinv, np.tril(np.linalg.inv(A))) < 1e-11


@pytest.mark.parametrize("n", [4096, 5000, 6272])
def test_potrf_lookahead_path_matches_library_cholesky(T, n):
    # >= 4096 takes the panel/look-ahead driver (two streams); 5000 and 6272 give ragged last panels.
    # torch.linalg.cholesky (cuSOLVER) is the checker here, never the product path.
    torch = T.torch
    g = torch.Generator(device="cuda").manual_seed(n)
    B = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g)
    A = B @ B.T / n + torch.eye(n, dtype=torch.float64, device="cuda")
    Ap = T.ops.pad_spd(A.clone())
    W, info = T.ops.potrf(Ap)
    assert info == 0
    Lref = torch.linalg.cholesky(A)
    L = torch.tril(Ap)[:n, :n]
    assert ((L - Lref).abs().max() / Lref.abs().max()).item() < 1e-12
    T.ops.trtri(Ap, W)
    Wn = torch.tril(W)[:n, :n]
    I = torch.eye(n, dtype=torch.float64, device="cuda")
    assert (Wn @ L - I).abs().max().item() < 1e-10
    # repeatability: the two-stream schedule must not change a bit
    Ap2 = T.ops.pad_spd(A.clone())
    T.ops.potrf(Ap2)
    assert torch.equal(torch.tril(Ap2), torch.tril(Ap))


def test_potrf_reports_first_bad_pivot(T):
    n = 300
    A = np.eye(n)
    A[170, 170] = -1.0
    Ap = T.ops.pad_spd(T.up(A))
    _, info = T.ops.potrf(Ap)
    assert info == 171          # LAPACK convention: 1-based index of the failing pivot


@pytest.mark.parametrize("kind,d,E", CASES)
@pytest.mark.parametrize("N", [1, 2, 10, 30, 96, 97, 128, 200, 700])   # <= 96: fused single-CTA kernel
def test_lml_and_gradient_match_oracle(T, kind, d, E, N):
    X, Y, th = util.random_case(100 + N, N, d, E, kind)
    buf = T.ops.FactorBuffers(N, T.dev)
    lml, g, info = T.ops.lml_grad(T.up(X), T.up(Y.ravel()), kind, d, th, buf)
    assert info == 0
    for form, tol in (("direct", 1e-9), ("gpy", 1e-6)):
        ref = go.inference(kind, X, Y, d, th, form=form)
        assert abs(lml - ref["lml"]) <= tol * abs(ref["lml"])
        assert util.rel_err(g, ref["grad"]) < max(tol, 1e-8)
        assert util.rel_err(buf.alpha.cpu().numpy()[:N], ref["alpha"].ravel()) < max(tol, 1e-8)
        Ki = np.tril(buf.A.cpu().numpy())[:N, :N]
        assert util.rel_err(Ki, np.tril(ref["Ki"])) < max(tol, 1e-8)
    # the posterior left behind serves predictions (W = L^-1, alpha), whichever kernel produced it
    lvl = T.ops.LevelRef(T.up(X), kind, d, th, buf)
    Xs = np.random.default_rng(N).uniform(size=(33, d + E))
    mean, var = T.ops.predict(lvl, T.up(Xs), True, True)
    ref = go.inference(kind, X, Y, d, th, form="direct", want_grad=False)
    mu_ref, var_ref = go.posterior_predict(kind, X, d, th, ref["L"], ref["alpha"], Xs, True, form="direct")
    assert util.rel_err(mean.cpu().numpy(), mu_ref.ravel(), max(np.max(np.abs(mu_ref)), 1e-3)) < 1e-8
    assert util.rel_err(var.cpu().numpy(), var_ref.ravel(), 1.5) < 1e-6


@pytest.mark.parametrize("kind,d,E", CASES)
def test_factorize_and_predict_match_oracle(T, kind, d, E):
    N, M = 150, 777
    X, Y, th = util.random_case(7, N, d, E, kind)
    Xs = np.random.default_rng(8).uniform(size=(M, d + E))
    Xs[:5] = X[:5]                                      # test points on training points: tiny variance
    buf = T.ops.FactorBuffers(N, T.dev)
    dX = T.up(X)
    lml, logdet, yta, info = T.ops.factorize(dX, T.up(Y.ravel()), kind, d, th, buf)
    assert info == 0
    ref = go.inference(kind, X, Y, d, th, form="gpy", want_grad=False)
    assert abs(lml - ref["lml"]) <= 1e-6 * abs(ref["lml"])
    lvl = T.ops.LevelRef(dX, kind, d, th, buf)
    for include_noise in (True, False):
        mean, var = T.ops.predict(lvl, T.up(Xs), True, include_noise)
        mu_ref, var_ref = go.posterior_predict(kind, X, d, th, ref["L"], ref["alpha"], Xs, include_noise)
        scale = np.max(np.abs(mu_ref))
        assert util.rel_err(mean.cpu().numpy(), mu_ref.ravel(), scale) < 1e-8
        kdiag = go.kernel_Kdiag(kind, th[:-1], 1)[0]
        assert util.rel_err(var.cpu().numpy(), var_ref.ravel(), kdiag) < 1e-6
    mean_only, none = T.ops.predict(lvl, T.up(Xs), want_var=False)
    assert none is None and T.torch.equal(mean_only, mean)


def test_predict_chunking_is_bit_invariant(T):
    # the chunk size of the cross-covariance scratch must not change a single bit of the outputs
    kind, d, E, N, M = go.KIND_COMPOSITE, 2, 5, 300, 1000
    X, Y, th = util.random_case(9, N, d, E, kind)
    Xs = np.random.default_rng(10).uniform(size=(M, d + E))
    buf = T.ops.FactorBuffers(N, T.dev)
    dX = T.up(X)
    T.ops.factorize(dX, T.up(Y.ravel()), kind, d, th, buf)
    lvl = T.ops.LevelRef(dX, kind, d, th, buf)
    from multifidelity_datafusion_gps_b200 import _ffi
    h = _ffi.get_handle(0)
    m1, v1 = T.ops.predict(lvl, T.up(Xs))
    m2, v2 = T.ops.predict(lvl, T.up(Xs), ws_bytes=h.lib.mfgp_predict_ws_bytes(N, 128))
    assert T.torch.equal(m1, m2) and T.torch.equal(v1, v2)
    m3, v3 = T.ops.predict(lvl, T.up(Xs[500:]))
    assert T.torch.equal(m1[500:], m3) and T.torch.equal(v1[500:], v3)


def test_empty_inputs(T):
    kind, d, E, N = go.KIND_RBF, 2, 0, 20
    X, Y, th = util.random_case(11, N, d, E, kind)
    buf = T.ops.FactorBuffers(N, T.dev)
    dX = T.up(X)
    T.ops.factorize(dX, T.up(Y.ravel()), kind, d, th, buf)
    lvl = T.ops.LevelRef(dX, kind, d, th, buf)
    mean, var = T.ops.predict(lvl, T.torch.empty((0, d), dtype=T.torch.float64, device=T.dev))
    assert mean.numel() == 0 and var.numel() == 0


def test_argmax_matches_numpy_including_ties(T):
    rng = np.random.default_rng(0)
    for n in (1, 5, 257, 100000, 1 << 20):
        v = rng.standard_normal(n)
        if n > 4:
            v[[n // 3, n // 2, n - 1]] = v.max() + 1.0          # three-way tie: lowest index wins
        val, idx = T.ops.argmax(T.up(v))
        assert idx == int(np.argmax(v)) and val == v[idx]


def test_philox_normals_are_standard_and_counter_based(T):
    a = T.ops.fill_normal(2, 0, 1 << 20, T.dev)
    b = T.ops.fill_normal(2, 1000, 4096, T.dev)
    assert T.torch.equal(a[1000:1000 + 4096], b)        # value depends on (seed, counter) only
    x = a.cpu().numpy()
    assert abs(x.mean()) < 5e-3 and abs(x.var() - 1.0) < 5e-3
    assert abs(np.mean(x ** 3)) < 2e-2 and abs(np.mean(x ** 4) - 3.0) < 5e-2
    c = T.ops.fill_normal(3, 0, 4096, T.dev)
    assert not T.torch.equal(a[:4096], c)


def test_bad_arguments_are_rejected(T):
    from multifidelity_datafusion_gps_b200 import _ffi
    X, _, th = util.random_case(1, 10, 2, 5, go.KIND_COMPOSITE)
    with pytest.raises(_ffi.MfgpError):
        T.ops.assemble(T.up(X), go.KIND_COMPOSITE, 2, th[:3])            # wrong P
    bad = th.copy()
    bad[1] = -1.0
    with pytest.raises(_ffi.MfgpError):
        T.ops.assemble(T.up(X), go.KIND_COMPOSITE, 2, bad)               # negative lengthscale


def test_large_factorization_properties(T):
    # size-independent properties at a size the oracle would take minutes for:
    # W L = I, K^-1 K_y = I (sampled columns), LML consistent with logdet/quadratic form
    kind, d, E, N = go.KIND_COMPOSITE, 4, 1, 4096
    rng = np.random.default_rng(1)
    X = rng.uniform(size=(N, d + E))
    Y = util.hf_4d(X[:, :4])
    th = np.array([1.0, 0.3, 1.0, 0.3, 0.1, 0.3, 0.01 * Y.var()])
    dX, dy = T.up(X), T.up(Y.ravel())
    buf = T.ops.FactorBuffers(N, T.dev)
    lml, logdet, yta, info = T.ops.factorize(dX, dy, kind, d, th, buf)
    assert info == 0
    torch = T.torch
    L = torch.tril(buf.A)
    W = torch.tril(buf.W)
    I = torch.eye(buf.npad, dtype=torch.float64, device=T.dev)
    assert (W @ L - I).abs().max().item() < 1e-9
    Ky = T.ops.assemble(dX, kind, d, th, uplo=1)
    assert ((L[:N, :N] @ L[:N, :N].T) - Ky).abs().max().item() < 1e-11
    assert (Ky @ buf.alpha[:N] - dy).abs().max().item() < 1e-8 * Y.max()
    assert abs(lml - 0.5 * (-N * np.log(2 * np.pi) - logdet - yta)) < 1e-9 * abs(lml)
    lml2, g, info = T.ops.lml_grad(dX, dy, kind, d, th, buf)
    assert info == 0 and abs(lml2 - lml) <= 1e-12 * abs(lml)
    Ki = torch.tril(buf.A) + torch.tril(buf.A, -1).T
    cols = torch.arange(0, N, 97, device=T.dev)
    R = Ki[:N, :N] @ Ky[:, cols]
    assert (R - I[:N, cols]).abs().max().item() < 1e-7
    # gradient against central finite differences of the GPU LML itself
    for i in (1, 3, 6):
        hstep = 1e-5 * th[i]
        tp, tm = th.copy(), th.copy()
        tp[i] += hstep
        tm[i] -= hstep
        lp = T.ops.factorize(dX, dy, kind, d, tp, buf)[0]
        lm = T.ops.factorize(dX, dy, kind, d, tm, buf)[0]
        fd = (lp - lm) / (2 * hstep)
        assert abs(g[i] - fd) <= 1e-5 * max(1.0, abs(fd))


@pytest.mark.parametrize("kind,d,E", [(go.KIND_COMPOSITE, 2, 1), (go.KIND_RBF, 3, 0)])
@pytest.mark.parametrize("N0", [5, 127, 128, 200])
def test_bordered_update_matches_refactorisation(T, kind, d, E, N0):
    """mfgp_append_point through gp.GPRegression.append_point: three points appended one by one at fixed
    theta (crossing a 128-tile boundary for N0 = 127, 128) against the oracle's factorisation of all points."""
    from multifidelity_datafusion_gps_b200 import gp
    X, Y, th = util.random_case(11, N0 + 3, d, E, kind)
    kern = gp.NARGPKernel(d, E) if kind == go.KIND_COMPOSITE else gp.RBF(d + E)
    m = gp.GPRegression(X[:N0], Y[:N0], kernel=kern)
    m._set_params(th)
    m.predict(X[:2])                                   # factorise the first N0 points
    for i in range(N0, N0 + 3):
        m.append_point(X[i], Y[i])
    assert m.N == N0 + 3 and m.npad == (N0 + 3 + 127) // 128 * 128
    ref = go.OracleGPRegression(X, Y, kind=kind, d=d, theta=th, form="direct")
    L_ref, alpha_ref = ref.posterior()
    n = N0 + 3
    assert util.rel_err(m._dalpha.cpu().numpy()[:n], alpha_ref.ravel()) < 1e-9
    assert np.all(m._dalpha.cpu().numpy()[n:] == 0.0)
    assert util.rel_err(np.tril(m._dA.cpu().numpy()[:n, :n]), L_ref) < 1e-11
    W = np.tril(m._dW.cpu().numpy()[:n, :n])             # blocks above the diagonal are trtri scratch
    assert util.rel_err(W @ L_ref, np.eye(n)) < 1e-9
    assert np.isclose(m._lml, ref.log_likelihood(), rtol=1e-9)
    Xq = np.random.default_rng(12).uniform(size=(64, d + E))
    mu, var = m.predict(Xq)
    mu_ref, var_ref = ref.predict(Xq)
    assert util.rel_err(mu, mu_ref) < 1e-8 and util.rel_err(var, var_ref, 1.5) < 1e-6


def test_full_size_lml_grad_properties_n16384(T):
    """BASELINE.json configs[3] at full size (N_h = 16384, look-ahead Cholesky, level-batched inverse):
    properties that do not need the O(N^3) oracle -- K^-1 K_y = I and L L^T = K_y on sampled columns,
    alpha solves the system, analytic gradient against central differences of the GPU LML."""
    kind, d, E, N = go.KIND_COMPOSITE, 4, 1, 16384
    torch = T.torch
    rng = np.random.default_rng(1)
    X4 = rng.uniform(size=(N, 4))
    X = np.concatenate([X4, util.lf_4d(X4)], axis=1)
    Y = util.hf_4d(X4)
    th = np.array([1.0, 0.3, 1.0, 0.3, 0.1, 0.3, 0.01 * Y.var()])
    dX, dy = T.up(X), T.up(Y.ravel())
    buf = T.ops.FactorBuffers(N, T.dev)
    lml, logdet, yta, info = T.ops.factorize(dX, dy, kind, d, th, buf)
    assert info == 0 and np.isfinite(lml)
    cols = torch.arange(5, N, 1021, device=T.dev)
    Ky_cols = T.ops.assemble(dX, kind, d, th, uplo=1)[:, cols]          # (N, 17) of the full matrix
    L = torch.tril(buf.A)
    assert ((L @ L[cols, :].T) - Ky_cols).abs().max().item() < 1e-10
    W = torch.tril(buf.W)
    E_cols = torch.zeros((N, cols.numel()), dtype=torch.float64, device=T.dev)
    E_cols[cols, torch.arange(cols.numel(), device=T.dev)] = 1.0
    assert (W @ L[:, cols] - E_cols).abs().max().item() < 1e-9           # W L = I
    del L, W
    alpha = buf.alpha[:N].clone()
    lml2, g, info = T.ops.lml_grad(dX, dy, kind, d, th, buf)
    assert info == 0 and abs(lml2 - lml) <= 1e-12 * abs(lml)
    Ki = torch.tril(buf.A) + torch.tril(buf.A, -1).T
    assert (Ki @ Ky_cols - E_cols).abs().max().item() < 1e-6             # K^-1 K_y = I
    assert (Ki @ dy - alpha).abs().max().item() < 1e-7 * alpha.abs().max().item()
    del Ki
    for i in (1, 6):
        hstep = 1e-5 * th[i]
        tp, tm = th.copy(), th.copy()
        tp[i] += hstep
        tm[i] -= hstep
        lp = T.ops.factorize(dX, dy, kind, d, tp, buf)[0]
        lm = T.ops.factorize(dX, dy, kind, d, tm, buf)[0]
        fd = (lp - lm) / (2 * hstep)
        assert abs(g[i] - fd) <= 2e-5 * max(1.0, abs(fd))


def _config4_case(N):
    rng = np.random.default_rng(1)
    X4 = rng.uniform(size=(N, 4))
    X = np.concatenate([X4, util.lf_4d(X4)], axis=1)
    Y = util.hf_4d(X4)
    th = np.array([1.0, 0.3, 1.0, 0.3, 0.1, 0.3, 0.01 * Y.var()])
    return X, Y, th


def _lml_grad_against_oracle(T, N):
    kind, d = go.KIND_COMPOSITE, 4
    X, Y, th = _config4_case(N)
    buf = T.ops.FactorBuffers(N, T.dev)
    lml, g, info = T.ops.lml_grad(T.up(X), T.up(Y.ravel()), kind, d, th, buf)
    assert info == 0
    alpha = buf.alpha[:N].cpu().numpy()
    del buf
    T.torch.cuda.empty_cache()
    ref = go.inference(kind, X, Y, d, th)                      # GPy-form oracle: the parity bound
    assert abs(lml - ref["lml"]) <= 1e-6 * abs(ref["lml"])
    assert np.max(np.abs(g - ref["grad"])) <= 1e-6 * np.max(np.abs(ref["grad"]))
    assert util.rel_err(alpha, ref["alpha"].ravel()) <= 1e-8
    return abs(lml - ref["lml"]) / abs(ref["lml"]), np.max(np.abs(g - ref["grad"])) / np.max(np.abs(ref["grad"]))


def test_lml_and_gradient_match_oracle_n8192(T):
    """Config 4 at half size against the O(N^3) oracle itself (look-ahead Cholesky, level-batched inverse,
    one-launch W^T W, fused gradient reduction): LML, all seven gradient entries and alpha."""
    _lml_grad_against_oracle(T, 8192)


@pytest.mark.skipif(not __import__("os").environ.get("MFGP_SLOW_TESTS"),
                    reason="opt-in (MFGP_SLOW_TESTS=1): the NumPy oracle needs ~1 min and ~40 GB at N = 16384")
def test_lml_and_gradient_match_oracle_n16384(T):
    """BASELINE.json configs[3] at FULL size against the oracle (bench.py's lml_grad block also runs this
    comparison and reports the differences in every bench line)."""
    _lml_grad_against_oracle(T, 16384)


@pytest.mark.parametrize("kind,d,E", [(go.KIND_COMPOSITE, 4, 1), (go.KIND_COMPOSITE, 2, 5), (go.KIND_RBF, 2, 5),
                                      (go.KIND_RBF, 3, 0), (go.KIND_COMPOSITE, 3, 2)])
@pytest.mark.parametrize("N", [1, 5, 16, 17, 30, 48, 96, 97, 128])
def test_batched_objective_matches_single_evaluations_and_oracle(T, kind, d, E, N):
    """mfgp_lml_grad_batch (one CTA per hyper-parameter vector, scalars only; what the optimiser calls at
    the reference's own sizes): every vector of a batch of 20 (two slices of the 16-vector launch) must give
    the bits it gives alone, agree with mfgp_lml_grad to round-off, and with the oracle at the parity bound."""
    from multifidelity_datafusion_gps_b200 import gp
    X, Y, th = util.random_case(40 + N, N, d, E, kind)
    kern = gp.NARGPKernel(d, E) if kind == go.KIND_COMPOSITE else gp.RBF(d + E)
    m = gp.GPRegression(X, Y, kernel=kern)
    rng = np.random.default_rng(N)
    thetas = th[None, :] * np.exp(0.4 * rng.standard_normal((20, len(th))))
    lml, grad, info = m.lml_and_grad_batch(thetas)
    assert not info.any()
    for b in (0, 7, 16, 19):
        l1, g1, i1 = m.lml_and_grad_batch(thetas[b][None, :])
        assert i1[0] == 0 and l1[0] == lml[b] and np.array_equal(g1[0], grad[b])      # batch-invariant bits
        l2, g2 = m.lml_and_grad(thetas[b])                                           # mfgp_lml_grad
        assert abs(l2 - lml[b]) <= 1e-11 * max(1.0, abs(l2))
        assert np.max(np.abs(g2 - grad[b])) <= 1e-9 * max(1.0, np.max(np.abs(g2)))
        ref = go.inference(kind, X, Y, d, thetas[b])
        assert abs(lml[b] - ref["lml"]) <= 1e-6 * max(1.0, abs(ref["lml"]))
        assert np.max(np.abs(grad[b] - ref["grad"])) <= 1e-6 * max(1.0, np.max(np.abs(ref["grad"])))
    lml_only, _, _ = m.lml_and_grad_batch(thetas, want_grad=False)
    assert np.array_equal(lml_only, lml)


def test_batched_objective_reports_failed_factorisations(T):
    from multifidelity_datafusion_gps_b200 import gp
    rng = np.random.default_rng(3)
    X = rng.uniform(size=(12, 3))
    Y = rng.standard_normal((12, 1))
    good = np.array([1.0, 0.5, 1.0, 0.5, 0.1, 0.5, 1e-2])
    m = gp.GPRegression(X, Y, kernel=gp.NARGPKernel(2, 1))
    lml, grad, info = m.lml_and_grad_batch(np.array([good, 2.0 * good, good]))
    assert not info.any() and lml[0] == lml[2] and np.all(np.isfinite(lml))
    Xn = X.copy()
    Xn[7, 1] = np.nan                                       # NaN covariances: pivot 8 is the first not > 0
    mn = gp.GPRegression(Xn, Y, kernel=gp.NARGPKernel(2, 1))
    lml, grad, info = mn.lml_and_grad_batch(np.array([good, 2.0 * good]))
    assert list(info) == [8, 8]
