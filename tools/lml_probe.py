"""One LML+gradient evaluation at N (default 16384) with stage timings; used plain and under ncu."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multifidelity_datafusion_gps_b200 import _ffi, ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rng = np.random.default_rng(1)
X = rng.uniform(size=(n, 4))
z = np.prod(np.sin(X * np.pi), axis=1) + 5.0
Xa = np.concatenate([X, (z - 0.25 * np.sin(X[:, 0] * 0.3))[:, None]], axis=1)
y = z
theta = np.array([1.0, 0.3, 1.0, 0.3, 0.1, 0.3, 0.01 * y.var()])
dX, dy = torch.from_numpy(Xa).cuda(), torch.from_numpy(y.copy()).cuda()
buf = ops.FactorBuffers(n, "cuda")
for r in range(reps):
    lml, g, info, ms = ops.lml_grad(dX, dy, _ffi.KIND_COMPOSITE, 4, theta, buf, timed=True)
    print("N=%d rep %d lml=%.6f info=%d total=%.2f ms  stages(assemble,potrf,trtri,solve,lauum,grad)=%s" %
          (n, r, lml, info, ms.sum(), np.round(ms, 3)))
if os.environ.get("MFGP_PROBE_CHECKSUM"):
    # bit-level fingerprint of the factors and the gradient (A/B runs of schedule-only changes must agree exactly)
    import hashlib
    W = buf.W if hasattr(buf, "W") else None
    parts = [np.asarray(g).tobytes(), np.float64(lml).tobytes()]
    if W is not None:
        parts.append(np.float64(W.double().sum().item()).tobytes())
        parts.append(np.float64((W * W).sum().item()).tobytes())
    print("checksum", hashlib.sha256(b"".join(parts)).hexdigest()[:16])
