"""Wall time of full fits (ARD recipe: 1 + 6 L-BFGS-B runs) at the reference's sizes."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import multifidelity_datafusion_gps_b200 as pkg  # noqa: E402

rs = np.random.RandomState(10)
rs.uniform(size=(100, 2))
X5 = rs.uniform(size=(5, 2))
X30 = np.vstack([X5, np.random.RandomState(3).uniform(size=(25, 2))])
for name, X in (("N_h=5", X5), ("N_h=30", X30)):
    for rep in range(3):
        np.random.seed(0)
        m = pkg.GPDF(2, 0.001, 2, bench.hf_2d, bench.lf_2d)
        t0 = time.perf_counter()
        m.fit(X)
        dt = time.perf_counter() - t0
        print("%s rep %d: fit %.2f ms, %d evaluations (%.1f us each), lml %.12f" %
              (name, rep, 1e3 * dt, m.hf_model.n_evals, 1e6 * dt / m.hf_model.n_evals, m.hf_model.log_likelihood()))
