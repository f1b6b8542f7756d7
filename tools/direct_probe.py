"""One adaptation step under the reference's default maximiser (DIRECT, 20 000 sequential single-point predicts) at
N_h = 30: resident point service vs launch-per-question latency path vs the NumPy oracle."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import multifidelity_datafusion_gps_b200 as pkg  # noqa: E402

rs = np.random.RandomState(10)
rs.uniform(size=(100, 2))
X30 = np.vstack([rs.uniform(size=(5, 2)), np.random.RandomState(3).uniform(size=(25, 2))])
theta = np.array([1.2, 0.9, 1e-3])
md = pkg.GPDF(2, 0.001, 2, bench.hf_2d, bench.lf_2d)
md.fit(X30, theta=theta)
md.get_input_with_highest_uncertainty(md)                       # warm
t0 = time.perf_counter(); x_s, f_s = md.get_input_with_highest_uncertainty(md); t_s = time.perf_counter() - t0
h = None
orig = md.point_service_start
md.point_service_start = lambda *a, **k: False
t0 = time.perf_counter(); x_l, f_l = md.get_input_with_highest_uncertainty(md); t_l = time.perf_counter() - t0
md.point_service_start = orig
print("DIRECT step, N_h=30: service %.3f s, launch-per-question %.3f s, same point %s, fopt %.6e" %
      (t_s, t_l, bool(np.array_equal(x_s, x_l) and f_s == f_l), f_s))
# raw round trips
assert md.point_service_start()
x = np.array([0.3, 0.7])
for _ in range(200):
    md.point_service_eval(x)
t0 = time.perf_counter()
for _ in range(5000):
    md.point_service_eval(x)
t_e = (time.perf_counter() - t0) / 5000
from multifidelity_datafusion_gps_b200 import _ffi
hh = md._svc_handle
fn, hp, px, po = md._svc_call
md._svc_x[:2] = x
md._svc_x[2:] = np.asarray(bench.lf_2d(x[None, :] + md._svc_offs_tau)).ravel()
t0 = time.perf_counter()
for _ in range(5000):
    fn(hp, px, po)
t_c = (time.perf_counter() - t0) / 5000
md.point_service_stop()
t0 = time.perf_counter()
for _ in range(2000):
    md.predict_point(x[None])
t_p = (time.perf_counter() - t0) / 2000
print("per question: service eval %.1f us (C round trip alone %.1f us), predict_point %.1f us" %
      (1e6 * t_e, 1e6 * t_c, 1e6 * t_p))
