"""LML+gradient evaluations at the reference's own sizes (N_h = 10, 30): wall time per evaluation through
the C-ABI, and a few launches for `ncu -k regex:small_gp` (profiles/r02_ncu_small_gp.txt)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multifidelity_datafusion_gps_b200 import _ffi, ops  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for n in (10, 30, 96):
    rng = np.random.default_rng(n)
    X = rng.uniform(size=(n, 5))
    y = np.sin(3.0 * X.sum(axis=1))
    theta = np.array([1.0, 0.5, 1.0, 0.6, 0.1, 0.5, 1e-3])
    dX, dy = torch.from_numpy(X).cuda(), torch.from_numpy(y.copy()).cuda()
    buf = ops.FactorBuffers(n, "cuda")
    for _ in range(5):
        lml, g, info = ops.lml_grad(dX, dy, _ffi.KIND_COMPOSITE, 4, theta, buf)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        lml, g, info = ops.lml_grad(dX, dy, _ffi.KIND_COMPOSITE, 4, theta, buf)
    dt = time.perf_counter() - t0
    print("N=%d: %.1f us per LML+gradient evaluation (wall, %d reps), lml=%.9f" % (n, 1e6 * dt / reps, reps, lml))

# the optimiser's objective: scalars-only batched kernel (mfgp_lml_grad_batch), B = 1 and B = 6 (one restart round)
from multifidelity_datafusion_gps_b200 import gp  # noqa: E402
for n in (10, 30, 96, 128):
    rng = np.random.default_rng(n)
    X = rng.uniform(size=(n, 5))
    Y = np.sin(3.0 * X.sum(axis=1))[:, None]
    m = gp.GPRegression(X, Y, kernel=gp.NARGPKernel(4, 1))
    theta = np.array([1.0, 0.5, 1.0, 0.6, 0.1, 0.5, 1e-3])
    for B in (1, 6, 16):
        th = np.tile(theta, (B, 1)) * (1.0 + 0.01 * np.arange(B))[:, None]
        for _ in range(5):
            m.lml_and_grad_batch(th)
        t0 = time.perf_counter()
        for _ in range(reps):
            lml, g, info = m.lml_and_grad_batch(th)
        dt = time.perf_counter() - t0
        print("N=%d B=%d: %.1f us per launch, %.1f us per evaluation (wall), lml[0]=%.9f" %
              (n, B, 1e6 * dt / reps, 1e6 * dt / reps / B, lml[0]))
