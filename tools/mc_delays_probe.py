"""Throughput of the MC propagation for a model with delays (GPDF 2-D, tau = 1e-3, n = 2 -> E = 5, D = 7):
N_l = 4096, N_h = 1024, M test points x S samples; prints samples/s and the per-class kernel times."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multifidelity_datafusion_gps_b200 as pkg  # noqa: E402
from multifidelity_datafusion_gps_b200 import _ffi, gp  # noqa: E402

A2 = [2.2 * np.pi, np.pi]
hf = lambda x: (np.sin(x[:, 0] * A2[0]) * np.sin(x[:, 1] * A2[1]))[:, None]
lf = lambda x: hf(x) - 1.2 * (np.sin(x[:, 0] * np.pi * 0.1) + np.sin(x[:, 1] * np.pi * 0.1))[:, None]

M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
S = int(sys.argv[2]) if len(sys.argv) > 2 else 100
rng = np.random.default_rng(1)
Xl, Xh = rng.uniform(size=(4096, 2)), rng.uniform(size=(1024, 2))
m = pkg.GPDF(2, 1e-3, 2, hf, None, lf_X=Xl[:8], lf_Y=lf(Xl[:8]))
m.lf_X, m.lf_Y = Xl, lf(Xl)
m.lf_model = gp.GPRegression(Xl, lf(Xl))
m.lf_model._set_params(np.array([1.0, 0.3, 1e-4]))
m.fit(Xh, theta=np.array([1.0, 0.5, 1e-4]))
dX = gp.to_device(rng.uniform(size=(M, 2)), 0)
h = _ffi.get_handle(0)
for rep in range(3):
    torch.cuda.synchronize()
    h.profile_enable(True)
    t0 = time.perf_counter()
    mean, var, _ = m.predict_mc_device(dX, S, None, 3, 0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    prof = h.profile_read()
    h.profile_enable(False)
    print("M=%d S=%d E=5: %.1f ms  %.2f M samples/s  finite=%s  classes(ms/launch, launches)=%s" % (
        M, S, dt * 1e3, M * S / dt / 1e6, bool(torch.isfinite(mean).all() and torch.isfinite(var).all()),
        {k: (round(v[0], 3), v[1]) for k, v in prof.items() if v[1] > 0}))
