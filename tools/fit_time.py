"""Wall time of a full fit (7 L-BFGS-B runs) and of single LML+gradient evaluations at tiny N on the GPU."""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import multifidelity_datafusion_gps_b200 as pkg

A2 = [2.2 * np.pi, np.pi]                       # reference tests/test_mfgp_adapt_2d.py:9-19


class util:
    @staticmethod
    def hf_2d(x):
        x = np.atleast_2d(x)
        return (np.sin(x[:, 0] * A2[0]) * np.sin(x[:, 1] * A2[1]))[:, None]

    @staticmethod
    def lf_2d(x):
        x = np.atleast_2d(x)
        return util.hf_2d(x) - 1.2 * (np.sin(x[:, 0] * np.pi * 0.1) + np.sin(x[:, 1] * np.pi * 0.1))[:, None]

rs = np.random.RandomState(10)
X = rs.uniform(size=(8, 2))
m = pkg.NARGP(2, util.hf_2d, util.lf_2d)
np.random.seed(1); t = time.perf_counter(); m.fit(X); dt = time.perf_counter() - t
print("GPU fit: %.3f s, %d evals, %.1f us/eval" % (dt, m.hf_model.n_evals, 1e6 * dt / m.hf_model.n_evals))
import cProfile, pstats
np.random.seed(1)
pr = cProfile.Profile(); pr.enable(); m.fit(X); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)

# steady-state evaluation rate at tiny N (wall clock, after warm-up)
for n in (8, 30, 100, 128):
    Xn = rs.uniform(size=(n, 2))
    mm = pkg.NARGP(2, util.hf_2d, util.lf_2d)
    mm.fit(Xn, theta=np.array([1.0, 0.8, 1.0, 0.5, 0.1, 0.3, 1e-3]))
    th = mm.hf_model.param_array
    for _ in range(20):
        mm.hf_model.lml_and_grad(th)
    t = time.perf_counter()
    for _ in range(200):
        mm.hf_model.lml_and_grad(th)
    print("N=%d: %.1f us per LML+grad evaluation (wall)" % (n, 1e6 * (time.perf_counter() - t) / 200))
