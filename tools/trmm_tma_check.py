"""A/B of the two trmm_sumsq variants (MFGP_TRMM_TMA=0|1, read once per process): prints a checksum of the
predictive variances (must be identical across the variants: same products, same order) and the average launch
time.  Run each variant in its own process, the TMA one under `timeout`."""
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multifidelity_datafusion_gps_b200 import _ffi, gp  # noqa: E402

h = _ffi.get_handle(0)
for n, m in ((300, 1000), (1024, 75776), (4096, 18944)):
    rng = np.random.default_rng(n)
    X = rng.uniform(size=(n, 5))
    Y = np.sin(3.0 * X.sum(axis=1))[:, None]
    model = gp.GPRegression(X, Y, kernel=gp.NARGPKernel(4, 1))
    model._set_params(np.array([1.0, 0.5, 1.0, 0.6, 0.1, 0.5, 1e-3]))
    dXq = torch.from_numpy(rng.uniform(size=(m, 5))).cuda()
    mean, var = model.predict_device(dXq)
    torch.cuda.synchronize()
    h.profile_enable(True)
    for _ in range(5):
        mean, var = model.predict_device(dXq)
    prof = h.profile_read()
    h.profile_enable(False)
    v = var.cpu().numpy()
    print("N=%d M=%d variant=%s sha=%s trmm_sumsq %.4f ms/launch (%d launches) min var %.3e" % (
        n, m, os.environ.get("MFGP_TRMM_TMA", "default"), hashlib.sha256(v.tobytes()).hexdigest()[:16],
        prof["trmm_sumsq"][0], prof["trmm_sumsq"][1], v.min()), flush=True)
