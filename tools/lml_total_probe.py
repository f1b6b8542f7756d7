"""LML+gradient at N: total time of the production schedule (no stage events: solves under K^-1 = W^T W) against the
sequential schedule with stage events; the results must be identical."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multifidelity_datafusion_gps_b200 import _ffi, ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
rng = np.random.default_rng(1)
X = rng.uniform(size=(n, 4))
z = np.prod(np.sin(X * np.pi), axis=1) + 5.0
Xa = np.concatenate([X, (z - 0.25 * np.sin(X[:, 0] * 0.3))[:, None]], axis=1)
theta = np.array([1.0, 0.3, 1.0, 0.3, 0.1, 0.3, 0.01 * z.var()])
dX, dy = torch.from_numpy(Xa).cuda(), torch.from_numpy(z.copy()).cuda()
buf = ops.FactorBuffers(n, "cuda")
ops.lml_grad(dX, dy, _ffi.KIND_COMPOSITE, 4, theta, buf)
res = {}
for timed in (False, True, False, True):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    out = ops.lml_grad(dX, dy, _ffi.KIND_COMPOSITE, 4, theta, buf, timed=timed)
    e.record(); torch.cuda.synchronize()
    res[timed] = out
    print("N=%d timed=%s total %.2f ms lml=%.9f" % (n, timed, s.elapsed_time(e), out[0]))
print("identical:", res[False][0] == res[True][0] and np.array_equal(res[False][1], res[True][1]),
      "alpha sum %.15g" % float(buf.alpha[:n].sum().item()))
