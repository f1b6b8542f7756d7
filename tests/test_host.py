"""CPU tests of the host side: C-ABI exports, public surface, sharding / arg-max plumbing (gloo)."""
import ctypes
import inspect
import os
import re
import socket

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from multifidelity_datafusion_gps_b200 import _ffi, build
    build.build()
    header = open(os.path.join(ROOT, "include", "mfgp_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(mfgp_[a-z_0-9]+)\s*\(", header)))
    assert len(declared) >= 18
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_ffi.EXPORTS) == declared
    lib.mfgp_padded_n.argtypes = [ctypes.c_int]
    assert [lib.mfgp_padded_n(n) for n in (1, 128, 129, 16384)] == [128, 128, 256, 16384]
    assert lib.mfgp_version() == 100


def test_mc_scratch_size_helper_is_pure_host_logic():
    """mfgp_predict_mc_ws_bytes: three doubles per (point, sample) column when every upper level is small enough for
    the fused kernel, four 128-column tiles per SM of padded rows otherwise; never less than one point's samples."""
    from multifidelity_datafusion_gps_b200 import _ffi
    lib = _ffi.load_library()
    f = lib.mfgp_predict_mc_ws_bytes
    M, S, d = 1 << 20, 100, 4
    small, general = f(100, 30, d, M, S), f(100, 1024, d, M, S)
    assert small >= 16 * M + 24 * M * S                                  # whole batch in one launch
    assert general >= 16 * M + 148 * 128 * 4 * (1024 + d + 4) * 8        # four tiles per SM
    assert f(4096, 1024, d, M, S) >= 16 * M + (4096 + 1) * 148 * 128 * 4 * 8   # the LF stage gets the same
    assert f(100, 30, d, 1 << 26, 1000) <= (12 << 30)                    # capped; the call then walks in chunks
    assert f(100, 65, d, 10, 7) >= (7 + 256) * (128 + d + 4) * 8         # minimum: one point's samples


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import multifidelity_datafusion_gps_b200 as pkg
    from multifidelity_datafusion_gps_b200 import _ffi
    with pytest.raises(_ffi.MfgpError):
        pkg.NARGP(1, lambda x: x, lambda x: x)
    h = ctypes.c_void_p()
    lib = _ffi.load_library()
    assert lib.mfgp_create(0, ctypes.byref(h)) != 0 and not h.value
    assert b"no CUDA device" in lib.mfgp_last_error(None)


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "multifidelity_datafusion_gps_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "/root/reference" not in src, f


def test_public_surface_matches_reference_signatures():
    import multifidelity_datafusion_gps_b200 as pkg
    sig = lambda f: list(inspect.signature(f).parameters)
    # src/MFDataFusion.py:56-59
    ref = ["name", "input_dim", "num_derivatives", "tau", "f_exact", "lower_bound", "upper_bound", "f_low",
           "lf_X", "lf_Y", "lf_hf_adapt_ratio", "use_composite_kernel", "adapt_maximizer", "eps", "add_noise"]
    ours = sig(pkg.MultifidelityDataFusion.__init__)[1:]
    assert ours[:len(ref)] == ref                      # the reference's parameters, same order
    assert ours[len(ref):] == ["augm_iterator"]        # extensions come after them, keyword-style
    # src/models/NARGP.py:15-17
    assert sig(pkg.NARGP.__init__)[1:12] == ["input_dim", "f_exact", "f_low", "name", "lower_bound",
                                             "upper_bound", "lf_X", "lf_Y", "lf_hf_adapt_ratio", "eps", "add_noise"]
    # src/models/GPDF.py:15-17, src/models/GPDFC.py:16-18
    for cls in (pkg.GPDF, pkg.GPDFC):
        assert sig(cls.__init__)[1:14] == ["input_dim", "tau", "num_derivatives", "f_exact", "f_low", "name",
                                           "lower_bound", "upper_bound", "lf_X", "lf_Y", "lf_hf_adapt_ratio",
                                           "eps", "add_noise"]
    # src/MFDataFusion.py:75,102,141,158
    assert sig(pkg.MultifidelityDataFusion.fit)[:2] == ["self", "hf_X"]
    assert sig(pkg.MultifidelityDataFusion.adapt) == ["self", "adapt_steps", "plot_mode", "X_test", "Y_test", "eps"]
    assert sig(pkg.MultifidelityDataFusion.predict) == ["self", "X_test"]
    assert sig(pkg.MultifidelityDataFusion.get_mse) == ["self", "X_test", "Y_test"]
    # src/adaptation_maximizers/abstract_maximizer.py:14
    assert sig(pkg.AbstractMaximizer.maximize) == ["self", "model_predict", "lower_bound", "upper_bound"]
    for name in ("fit", "adapt", "predict", "get_mse"):
        assert name in pkg.AbstractMFGP.__abstractmethods__


def test_iterators_follow_reference_protocol():
    import multifidelity_datafusion_gps_b200 as pkg
    g = np.load(os.path.join(ROOT, "tests", "golden", "augm_offsets.npz"))
    for n in range(4):
        for dim in range(1, 5):
            for cls, key in ((pkg.BackwardAugmentation, "backward"), (pkg.EvenAugmentation, "even")):
                it = cls(n, dim=dim)
                first = np.array(list(it)).reshape(-1, dim)
                second = np.array(list(it)).reshape(-1, dim)       # reusable after StopIteration
                assert np.array_equal(first, g["%s_n%d_d%d" % (key, n, dim)])
                assert np.array_equal(first, second)
                assert len(first) == it.new_entries_count()
                assert np.array_equal(it.offset_table(), first)


def test_candidate_maximizer_with_foreign_predict():
    import multifidelity_datafusion_gps_b200 as pkg
    cands = np.random.default_rng(0).uniform(size=(1000, 3))
    var = np.random.default_rng(1).uniform(size=1000)
    mx = pkg.CandidateSetMaximizer(candidates=cands)
    x, fopt = mx.maximize(lambda c: (None, var[:, None]), np.zeros(3), np.ones(3))
    assert np.array_equal(x, cands[np.argmax(var)]) and fopt == -var.max()
    mx2 = pkg.CandidateSetMaximizer(n_candidates=50, seed=3)
    lb, ub = np.array([1.0, -1.0]), np.array([2.0, 1.0])
    mx2.maximize(lambda c: (None, np.arange(len(c), dtype=float)[:, None]), lb, ub)
    assert mx2.candidates.shape == (50, 2) and (mx2.candidates >= lb).all() and (mx2.candidates <= ub).all()


def test_shard_ranges_and_argmax_combine():
    from multifidelity_datafusion_gps_b200 import dist
    for n in (0, 1, 7, 100, 1 << 20):
        for world in (1, 2, 3, 8):
            spans = [dist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert dist.combine_argmax([0.5, 0.9, 0.9, -np.inf], [3, 40, 17, -1]) == (0.9, 17)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as tdist
    from multifidelity_datafusion_gps_b200 import dist
    from oracle import gpy_oracle as go
    tdist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    X, Y = rng.uniform(size=(20, 2)), rng.standard_normal((20, 1))
    model = go.OracleGPRegression(X, Y, theta=[1.0, 0.3, 0.01])
    cands = np.random.default_rng(1).uniform(size=(5001, 2))
    lo, hi = dist.shard_range(len(cands), rank, world)
    var = model.predict(cands[lo:hi])[1].ravel()
    i = int(np.argmax(var))
    val, idx = dist.gather_argmax(var[i], lo + i)
    total = dist.allreduce_sum_scalar(float(var.sum()))
    full = model.predict(cands)[1].ravel()
    q.put((rank, idx, val, int(np.argmax(full)), float(full.max()), total, float(full.sum())))
    tdist.destroy_process_group()


def test_sharded_argmax_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, idx, val, ref_idx, ref_val, total, ref_total in res:
        assert idx == ref_idx and val == ref_val           # shard-count invariant, bit-exact
        assert np.isclose(total, ref_total, rtol=1e-12)


# ---- restart-parallel fit: host protocol on gloo ----------------------
def _stub_gp():
    """gp.GPRegression with the GPU objective replaced by a multi-modal host function of the
    transformed parameters: the restart protocol (RNG order, shares, gather, winner) is what is tested."""
    import scipy.optimize as sopt
    from multifidelity_datafusion_gps_b200 import gp

    class Stub(gp.GPRegression):
        def __init__(self):                      # no device, no kernels
            self._theta = np.ones(3)
            self._fixed = np.zeros(3, dtype=bool)
            self.optimization_runs = []

        @property
        def param_array(self):
            return self._theta.copy()

        def _set_params(self, theta):
            self._theta = np.asarray(theta, dtype=np.float64).copy()

        def optimize(self, optimizer=None, max_iters=1000, messages=False, **kw):
            f = lambda x: (float(np.sum(np.sin(3.0 * x) + 0.1 * x ** 2)), 3.0 * np.cos(3.0 * x) + 0.2 * x)
            x0 = gp.logexp_finv(self.param_array)
            x_opt, f_opt, _ = sopt.fmin_l_bfgs_b(f, x0, maxfun=max_iters, maxiter=max_iters)
            self._set_params(gp.logexp_f(x_opt))
            self.optimization_runs.append((x_opt, float(f_opt)))
            return self
    return Stub()


def _restart_worker(rank, world, port, q):
    import torch.distributed as tdist
    tdist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    np.random.seed(123)                          # same RNG state on every rank (documented requirement)
    m = _stub_gp()
    m.optimize_restarts(7, parallel=True, max_iters=200)
    q.put((rank, m.param_array, [f for _, f in m.optimization_runs]))
    tdist.destroy_process_group()


def test_restart_parallel_fit_world_size_2_gloo_equals_serial():
    import torch.multiprocessing as mp
    from multifidelity_datafusion_gps_b200 import dist
    assert dist.restart_share(7, 0, 2) == [0, 2, 4, 6] and dist.restart_share(7, 1, 2) == [1, 3, 5]
    assert dist.best_run([(0, 2.0, "a"), (1, 1.0, "b"), (2, 1.0, "c"), (3, float("nan"), "d")])[2] == "b"
    np.random.seed(123)
    serial = _stub_gp()
    serial.optimize_restarts(7, max_iters=200)   # world size 1: the reference's serial loop
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_restart_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, theta, objs in res:
        assert np.array_equal(theta, serial.param_array)              # same winner, bit for bit
        assert objs == [f for _, f in serial.optimization_runs]       # same runs, in run order


# ---- the stepped L-BFGS-B driver must be SciPy's optimiser, bit for bit --------------------------------------
def test_stepped_lbfgsb_reproduces_scipy_bit_for_bit():
    """gp.GPRegression drives SciPy's compiled L-BFGS-B step directly (no ScalarFunction layers; resumable
    at every evaluation so that independent restarts can share one batched GPU launch).  Its iterates must
    equal scipy.optimize.fmin_l_bfgs_b's (what paramz calls) on smooth, failing (+inf) and budget-limited
    runs; otherwise the package falls back to fmin_l_bfgs_b (selfcheck)."""
    from scipy import optimize as sopt
    from multifidelity_datafusion_gps_b200 import _lbfgsb as lb
    assert lb.selfcheck()

    def make(seed):
        rng = np.random.default_rng(seed)
        A = rng.standard_normal((7, 7))
        Q, b = A @ A.T + 0.1 * np.eye(7), rng.standard_normal(7)

        def f(x):
            if np.abs(x).max() > 6.0:
                return np.inf, np.zeros(7)
            return float(0.5 * x @ Q @ x - b @ x + np.sum(np.sin(2 * x))), Q @ x - b + 2 * np.cos(2 * x)
        return f
    for seed in range(12):
        f = make(seed)
        x0 = 2.0 * np.random.default_rng(100 + seed).standard_normal(7)
        for budget in (1000, 15, 5):
            a = sopt.fmin_l_bfgs_b(f, x0, maxfun=budget, maxiter=budget)
            b = lb.minimize(f, x0, maxfun=budget, maxiter=budget)
            assert np.array_equal(a[0], b[0]) and a[1] == b[1]
            assert (a[2]["funcalls"], a[2]["nit"], a[2]["warnflag"]) == (b[2]["funcalls"], b[2]["nit"], b[2]["warnflag"])
    f = make(0)
    x0s = [np.random.default_rng(200 + s).standard_normal(7) for s in range(6)]
    calls = []

    def batch(xs):
        calls.append(len(xs))
        return [f(x) for x in xs]
    res = lb.minimize_lockstep(batch, x0s, 1000, 1000)
    for (x, fx, info), x0 in zip(res, x0s):
        a = sopt.fmin_l_bfgs_b(f, x0, maxfun=1000, maxiter=1000)
        assert np.array_equal(x, a[0]) and fx == a[1] and info["funcalls"] == a[2]["funcalls"]
    assert calls[0] == 6 and sum(calls) == sum(r[2]["funcalls"] for r in res)    # one batch per round
