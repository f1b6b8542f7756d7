#!/usr/bin/env python
"""bench.py -- headline benchmark of the multi-fidelity GP hot path on B200.

Metric (BASELINE.json): MC predictive samples/s = test points x MC samples pushed through the
NARGP 2-level posterior per second, plus (N=1 only) LML+gradient evaluations/s at N=16384.

A "step" is one pass of the hot path over one batch: predict_mc on the full test batch
(M = 32^4 = 1 048 576 tensor Gauss-Legendre nodes x S = 100 low-fidelity posterior samples, PCE mean
reduced at the end).  Multi-GPU: weak scaling -- every rank processes its own batch of M points (global
point indices rank*M ..), no data-path collective; the only collectives are the one-off NCCL broadcast
of the factorised state (untimed, reported) and one all_reduce of the PCE-mean scalar per step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                  [--nh 1024 --nl 4096 --m 1048576 --s 100] [--no-lml]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PI = np.pi


def hf_4d(x):                    # reference tests/test_mfgp_adapt_4d.py:13-15
    return (np.prod(np.sin(x[:, :4] * PI), axis=1) + 5.0)[:, None]


def lf_4d(x):                    # reference tests/test_mfgp_adapt_4d.py:18-21
    return hf_4d(x) - 0.25 * (np.sin(x[:, 0] * PI * 0.1) + np.sin(x[:, 1] * PI * 0.05)
                              + np.sin(x[:, 2] * 0.15 * PI) + np.sin(x[:, 3] * 0.2 * PI))[:, None]


def gauss_legendre_grid(n_per_dim, dim):
    x, w = np.polynomial.legendre.leggauss(n_per_dim)
    x, w = 0.5 * (x + 1.0), 0.5 * w
    grids = np.meshgrid(*([x] * dim), indexing="ij")
    nodes = np.stack([g.ravel() for g in grids], axis=1)
    wg = np.meshgrid(*([w] * dim), indexing="ij")
    weights = np.prod(np.stack([g.ravel() for g in wg], axis=1), axis=1)
    return np.ascontiguousarray(nodes), np.ascontiguousarray(weights)


def workload(args):
    """Synthetic config-5 model: d = 4, U[0,1]^4 training inputs (default_rng(1)),
    y_l = lf_4d, y_h = hf_4d, fixed hyper-parameters (config 4), test set = tensor Gauss-Legendre grid."""
    rng = np.random.default_rng(1)
    Xh = rng.uniform(size=(args.nh, 4))
    Xl = rng.uniform(size=(args.nl, 4))
    yh, yl = hf_4d(Xh), lf_4d(Xl)
    n1 = int(round(args.m ** 0.25))
    if n1 ** 4 == args.m:
        Xt, w = gauss_legendre_grid(n1, 4)
    else:
        Xt = np.random.default_rng(3).uniform(size=(args.m, 4))
        w = np.full(args.m, 1.0 / args.m)
    lf_theta = np.array([1.0, 0.3, 0.01 * yl.var()])
    hf_theta = np.array([1.0, 0.3, 1.0, 0.3, 0.1, 0.3, 0.01 * yh.var()])
    return dict(Xh=Xh, Xl=Xl, yh=yh, yl=yl, Xt=Xt, w=w, lf_theta=lf_theta, hf_theta=hf_theta)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def oracle_mc_sample(wl, args, n_points, threads_note):
    """CPU baseline: the oracle's vectorised predict_mc on a bounded sample of the same workload.
    Returns (samples_per_s, seconds, description).  Factorisations are setup (untimed), as on the GPU."""
    from oracle import mfgp_oracle as mo
    o = mo.OracleMFGP(4, 0, 0, hf_4d, lf_X=wl["Xl"], lf_Y=wl["yl"], lf_theta=wl["lf_theta"])
    o.fit(wl["Xh"], theta=wl["hf_theta"])
    o.lf_model.posterior()
    o.hf_model.posterior()
    rng = np.random.default_rng(2)
    chunk = 128
    t0 = time.perf_counter()
    acc = 0.0
    for lo in range(0, n_points, chunk):
        X = wl["Xt"][lo:lo + chunk]
        eps = rng.standard_normal((X.shape[0], args.s, 1))
        mean, var = o.predict_mc(X, eps)
        acc += float(np.sum(wl["w"][lo:lo + chunk] * mean[:, 0]))
    dt = time.perf_counter() - t0
    return n_points * args.s / dt, dt, "first %d of %d test points x %d samples, %s" % (
        n_points, args.m, args.s, threads_note)


def run_reference(args):
    """--impl reference: the reference's CPU path for this metric.  The reference itself (GPy) cannot be
    installed here, so this is the oracle port on all host cores; each step = a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload(args)
    cores = os.cpu_count()
    note = "NumPy/SciPy OpenBLAS, %d host threads" % cores
    n_pts = args.ref_points
    for _ in range(args.warmup):
        oracle_mc_sample(wl, args, min(n_pts, 128), note)
    vals, secs = [], []
    for _ in range(args.steps):
        v, dt, sample = oracle_mc_sample(wl, args, n_pts, note)
        vals.append(v); secs.append(dt)
    value = float(n_pts * args.s * args.steps / sum(secs))
    line = {
        "impl": "reference", "metric": "mc_predictive_samples_per_s", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(secs) / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference (GPy 1.9.9) not installable offline; oracle port timed on host cores",
    }
    print(json.dumps(line))


def config_dict(args):
    return {"workload": "NARGP 2-level MC prediction: M=%d test points (32^4 Gauss-Legendre nodes) x S=%d "
                        "LF posterior samples -> PCE mean; N_h=%d, N_l=%d, d=4 (BASELINE configs[4])"
                        % (args.m, args.s, args.nh, args.nl),
            "M": args.m, "S": args.s, "N_h": args.nh, "N_l": args.nl, "d": 4,
            "parallelism": "test points sharded, %d rank(s), weak" % args.gpus,
            "l2": "inputs_larger_than_L2 (cross-covariance chunks of >300 MB stream through HBM each step)"}


def measure_fp64_peak(torch):
    """cuBLAS DGEMM 8192^3 as the FP64-tensor measuring stick (not on any product path)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    best = 1e30
    for i in range(6):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); torch.matmul(a, b); e.record(); torch.cuda.synchronize()
        if i > 0:
            best = min(best, s.elapsed_time(e))
    del a, b
    return 2.0 * n ** 3 / best / 1e9


def lml_grad_section(torch, pkg_ops, peaks, fp64_peak, n=16384):
    """Secondary metric: LML+gradient evaluations/s at N = 16384 on one B200 (config 4)."""
    from multifidelity_datafusion_gps_b200 import _ffi
    rng = np.random.default_rng(1)
    X = rng.uniform(size=(n, 4))
    Xa = np.concatenate([X, lf_4d(X)], axis=1)
    y = hf_4d(X)
    theta = np.array([1.0, 0.3, 1.0, 0.3, 0.1, 0.3, 0.01 * y.var()])
    dX = torch.from_numpy(Xa).cuda()
    dy = torch.from_numpy(y.ravel().copy()).cuda()
    buf = pkg_ops.FactorBuffers(n, "cuda")
    pkg_ops.lml_grad(dX, dy, _ffi.KIND_COMPOSITE, 4, theta, buf)          # warm-up
    torch.cuda.synchronize()
    reps, stages, total = 3, np.zeros(6), 0.0
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        lml, g, info, ms = pkg_ops.lml_grad(dX, dy, _ffi.KIND_COMPOSITE, 4, theta, buf, timed=True)
        e.record(); torch.cuda.synchronize()
        total += s.elapsed_time(e); stages += ms
    ms_eval = total / reps
    stages /= reps
    names = ["assemble", "potrf", "trtri", "solve", "lauum", "grad_reduce"]
    hbm = peaks.get("hbm_gbs", 6650.0)
    asm_bytes = 4.0 * n * (n + 1) + 8.0 * n * 5          # lower triangle written + inputs read
    grad_bytes = 4.0 * n * (n + 1) + 8.0 * n * 5         # lower triangle of K^-1 read + inputs
    out = {
        "n": n, "evals_per_s": 1e3 / ms_eval, "ms_per_eval": ms_eval, "lml": lml, "info": int(info),
        "stages_ms": {k: float(v) for k, v in zip(names, stages)},
        "roofline_eval": {"bound": "tensor", "achieved": n ** 3 / ms_eval / 1e9, "peak": fp64_peak,
                          "unit": "TFLOP/s", "frac": n ** 3 / ms_eval / 1e9 / fp64_peak,
                          "flops": float(n) ** 3},
        "roofline_potrf": {"bound": "tensor", "achieved": n ** 3 / 3 / stages[1] / 1e9, "peak": fp64_peak,
                           "unit": "TFLOP/s", "frac": n ** 3 / 3 / stages[1] / 1e9 / fp64_peak},
        "roofline_assemble": {"bound": "hbm", "achieved": asm_bytes / stages[0] / 1e6, "peak": hbm,
                              "unit": "GB/s", "frac": asm_bytes / stages[0] / 1e6 / hbm,
                              "bytes": asm_bytes, "note": "lower triangle only: 4N(N+1)+8ND"},
        "roofline_grad_reduce": {"bound": "hbm", "achieved": grad_bytes / stages[5] / 1e6, "peak": hbm,
                                 "unit": "GB/s", "frac": grad_bytes / stages[5] / 1e6 / hbm},
        "scaling": "replicas only (single-GPU Cholesky; SURVEY.md section 8e)",
    }
    del buf, dX, dy
    torch.cuda.empty_cache()
    # CPU baseline for this metric: the oracle's LML+gradient on the host cores at a
    # size it finishes in seconds, next to the GPU path at the same size
    try:
        from oracle import gpy_oracle as go
        nc = 4096
        Xc, yc = Xa[:nc], y[:nc]
        t0 = time.perf_counter()
        ref = go.inference(go.KIND_COMPOSITE, Xc, yc, 4, theta)
        cpu_ms = 1e3 * (time.perf_counter() - t0)
        dXc, dyc = torch.from_numpy(Xc.copy()).cuda(), torch.from_numpy(yc.ravel().copy()).cuda()
        bufc = pkg_ops.FactorBuffers(nc, "cuda")
        pkg_ops.lml_grad(dXc, dyc, _ffi.KIND_COMPOSITE, 4, theta, bufc)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lml_c, g_c, _ = pkg_ops.lml_grad(dXc, dyc, _ffi.KIND_COMPOSITE, 4, theta, bufc)
        torch.cuda.synchronize()
        gpu_ms = 1e3 * (time.perf_counter() - t0)
        out["cpu_baseline"] = {"n": nc, "cpu_ms_per_eval": cpu_ms, "gpu_ms_per_eval": gpu_ms, "cores": os.cpu_count(),
                               "kind": "port", "lml_rel_diff": abs(lml_c - ref["lml"]) / abs(ref["lml"]),
                               "grad_rel_diff": float(np.max(np.abs(g_c - ref["grad"])) / np.max(np.abs(ref["grad"])))}
    except Exception as exc:      # the baseline is a report, not a dependency of the metric
        out["cpu_baseline"] = {"error": repr(exc)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--nh", type=int, default=1024)
    ap.add_argument("--nl", type=int, default=4096)
    # (--points / --samples rather than --m / --s: torchrun's own parser treats "--m" as ambiguous)
    ap.add_argument("--points", "--m", dest="m", type=int, default=32 ** 4)
    ap.add_argument("--samples", "--s", dest="s", type=int, default=100)
    ap.add_argument("--ref-points", type=int, default=1024, dest="ref_points")
    ap.add_argument("--cpu-points", type=int, default=1024, dest="cpu_points")
    ap.add_argument("--no-lml", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as tdist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a B200 (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = "cuda:%d" % local

    import multifidelity_datafusion_gps_b200 as pkg
    from multifidelity_datafusion_gps_b200 import _ffi, gp, ops
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    wl = workload(args)
    # ---- setup (untimed): fit state on rank 0 at fixed theta, NCCL broadcast of the factorised state
    # The constructor trains the LF GP with L-BFGS-B (reference src/abstractMFGP.py:100-103); the bench
    # runs at FIXED hyper-parameters, so it is built on an 8-point stub and the real LF level is
    # installed directly (rank 0) or received by the broadcast (other ranks).
    model = pkg.NARGP(4, hf_4d, None, lf_X=wl["Xl"][:8], lf_Y=wl["yl"][:8])
    t_fit0 = time.perf_counter()
    if rank == 0:
        model.lf_X, model.lf_Y = wl["Xl"], wl["yl"]
        model.lf_model = gp.GPRegression(wl["Xl"], wl["yl"])
        model.lf_model._set_params(wl["lf_theta"])
        model.lf_model._ensure_posterior()
        model.fit(wl["Xh"], theta=wl["hf_theta"])
        model.hf_model._ensure_posterior()
    torch.cuda.synchronize()
    fit_s = time.perf_counter() - t_fit0
    bcast_ms = 0.0
    if world > 1:
        tdist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.broadcast_state(src=0)
        torch.cuda.synchronize(); tdist.barrier()
        bcast_ms = 1e3 * (time.perf_counter() - t0)

    M, S = args.m, args.s
    m0 = rank * M                                   # weak scaling: every rank owns its own batch
    Xt_pin = torch.from_numpy(wl["Xt"]).pin_memory()
    w_pin = torch.from_numpy(wl["w"]).pin_memory()
    dX = Xt_pin.to(dev)
    dw = w_pin.to(dev)
    h = _ffi.get_handle(local)

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def step_device():
        mean, var, wsum = model.predict_mc_device(dX, S, None, 2, m0, dw)
        if world > 1:
            t = torch.tensor([wsum], dtype=torch.float64, device=dev)
            tdist.all_reduce(t)
            wsum = float(t.item())
        return wsum

    def step_e2e():
        # public API, host buffers: pinned NumPy views in, NumPy out (H2D + D2H inside the timed region)
        mean, var = model.predict_mc(Xt_pin.numpy(), n_samples=S, seed=2, weights=w_pin.numpy(), m0=m0)
        wsum = model.last_pce_mean
        if world > 1:
            t = torch.tensor([wsum], dtype=torch.float64, device=dev)
            tdist.all_reduce(t)
            wsum = float(t.item())
        return wsum, mean, var

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        pce = step_device()
    # ---- timed: device-resident inputs
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    h.profile_enable(True)
    launches0 = h.launches
    s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_ev.record()
    for _ in range(args.steps):
        pce = step_device()
    e_ev.record()
    barrier()
    launches = h.launches - launches0
    prof = h.profile_read()
    h.profile_enable(False)
    ms_dev = max_over_ranks(s_ev.elapsed_time(e_ev)) / args.steps
    clocks = sampler.stop() if sampler else None

    # ---- timed: end to end through the public API with host buffers
    step_e2e()
    barrier()
    s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s_ev.record()
    for _ in range(args.steps):
        pce_e2e, mean_h, var_h = step_e2e()
    e_ev.record()
    barrier()
    ms_e2e = max_over_ranks(max(s_ev.elapsed_time(e_ev), 1e3 * (time.perf_counter() - t0))) / args.steps
    h2d = wl["Xt"].nbytes + wl["w"].nbytes
    d2h = mean_h.nbytes + var_h.nbytes + 8

    total_samples = float(world) * M * S
    value = total_samples / (ms_dev * 1e-3)
    e2e_value = total_samples / (ms_e2e * 1e-3)

    if rank == 0:
        fp64_peak = measure_fp64_peak(torch)
        trmm_ms, trmm_cnt = prof["trmm_sumsq"]
        hf_launches_per_step = max(trmm_cnt / args.steps, 1.0)
        cols_per_launch = M * S / hf_launches_per_step
        flops_per_launch = cols_per_launch * float(args.nh) ** 2       # N_h^2 per (point, sample)
        achieved = flops_per_launch / (trmm_ms * 1e-3) / 1e12 if trmm_ms > 0 else 0.0
        roofline = {"bound": "tensor", "kernel": "dg::trmm_sumsq_kernel (tmp = W Kx, fused column sum of squares)",
                    "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": achieved / fp64_peak if fp64_peak else None,
                    # DRAM read+write per launch from the ncu --set full capture of this kernel
                    # (profiles/r01_ncu_mc_hf_kernels_v6.txt: 1.651 GB for 75 776 columns at N_h = 1024),
                    # scaled to this run's columns per launch
                    "traffic": (1.651e9 / 75776.0 * cols_per_launch) if args.nh == 1024 else None,
                    "avg_launch_ms": trmm_ms, "launches_per_step": hf_launches_per_step,
                    "flops_per_launch": flops_per_launch,
                    "peak_source": "cuBLAS DGEMM 8192^3 measured live in this run (MEASURED_PEAKS.json has no FP64 "
                                   "figure; tcgen05 has no FP64 kind, FP64 MMA on sm_100a is DMMA.8x8x4)",
                    "share_of_step": trmm_ms * hf_launches_per_step / ms_dev,
                    "other_kernels_ms_per_launch": {k: v[0] for k, v in prof.items() if v[1] > 0}}
        line = {
            "metric": "mc_predictive_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args),
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "pce_mean": pce, "pce_mean_e2e": pce_e2e,
            "setup": {"fit_fixed_theta_s": fit_s, "nccl_broadcast_state_ms": bcast_ms},
        }
        if not args.no_cpu:
            v, dt, sample = oracle_mc_sample(wl, args, args.cpu_points,
                                             "NumPy/SciPy OpenBLAS, %d host threads" % os.cpu_count())
            line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": sample, "seconds": dt}
        if world == 1 and not args.no_lml:
            # acquisition throughput (candidates/s for A9): arg-max of the predictive
            # variance over the same M points taken as candidates, device-resident
            acq_ms = []
            for i in range(3):
                s_a, e_a = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s_a.record()
                _, var_c = model._predict_device(dX)
                c_val, c_idx = ops.argmax(var_c)
                e_a.record(); torch.cuda.synchronize()
                if i > 0:
                    acq_ms.append(s_a.elapsed_time(e_a))
            line["acquisition"] = {"candidates": int(M), "N_h": args.nh, "N_l": args.nl,
                                   "ms": float(np.mean(acq_ms)),
                                   "candidates_per_s": float(M / (np.mean(acq_ms) * 1e-3)),
                                   "argmax_index": int(c_idx), "max_variance": float(c_val)}
            if not args.no_cpu:
                # CPU baseline: the oracle's vectorised predict + np.argmax on a bounded sample of the candidates
                from oracle import mfgp_oracle as mo
                o = mo.OracleMFGP(4, 0, 0, hf_4d, lf_X=wl["Xl"], lf_Y=wl["yl"], lf_theta=wl["lf_theta"])
                o.fit(wl["Xh"], theta=wl["hf_theta"])
                o.lf_model.posterior(); o.hf_model.posterior()
                nc = 8192
                t0 = time.perf_counter()
                v_ref = o.predict(wl["Xt"][M - nc:])[1].ravel()
                i_ref = int(np.argmax(v_ref))
                dt = time.perf_counter() - t0
                line["acquisition"]["cpu_baseline"] = {
                    "candidates_per_s": nc / dt, "cores": os.cpu_count(), "kind": "port",
                    "sample": "last %d of %d candidates, NumPy/SciPy oracle" % (nc, M), "seconds": dt,
                    "argmax_agrees_on_sample": bool(M - nc + i_ref == int(c_idx)) if int(c_idx) >= M - nc else None}
            del dX
            gp._ws_pool.clear()
            torch.cuda.empty_cache()
            line["lml_grad"] = lml_grad_section(torch, ops, peaks, fp64_peak)
        print(json.dumps(line))
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
