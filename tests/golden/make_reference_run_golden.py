"""Generate tests/golden/reference_runs.npz by EXECUTING the reference's own classes.

The reference (pure Python) owns the orchestration of the hot path -- augmentation
(src/MFDataFusion.py:177-208), kernel composition (src/abstractMFGP.py:62-80), the ARD fit recipe
(:131-137), predict incl. add_noise (src/MFDataFusion.py:141-156), the adaptation loop
(src/abstractMFGP.py:317-359), the DIRECT maximizer wrapper (src/adaptation_maximizers/
scipydirect_wrapper.py:16-31) -- on top of GPy, which cannot be installed here.  This script imports
/root/reference/src UNMODIFIED with tests/golden/fake_gpy.py registered as `GPy` (GP arithmetic +
paramz protocol from oracle/gpy_oracle.py; scipydirect -> scipy.optimize.direct) and records what the
reference's code computes on seeded scenarios.  The vectors pin
  * oracle/mfgp_oracle.py's restatement of that orchestration (tests/test_oracle.py), and
  * the CUDA classes at the hyper-parameters the reference run arrived at (tests/test_gpu_models.py).
GPy's internals stay restated.  Run in the build container only:
    python tests/golden/make_reference_run_golden.py
"""
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.golden import fake_gpy  # noqa: E402

fake_gpy.install()
sys.path.insert(0, "/root/reference")
import src.models as ref_models  # noqa: E402  (the reference, unmodified)

from tests import util  # noqa: E402


THETA_C = np.array([1.0, 0.8, 1.0, 0.5, 0.1, 0.3, 1e-3])
THETA_R = np.array([1.2, 0.9, 1e-3])


def scenario(name, model, X_hf, X_test, seed, out, adapt_steps=0):
    np.random.seed(seed)
    with redirect_stdout(io.StringIO()):
        model.fit(X_hf)
        if adapt_steps:
            model.adapt(adapt_steps)
    mean, var = model.predict(X_test)
    out[name + "/seed"] = np.array(seed)
    out[name + "/X_hf"] = X_hf
    out[name + "/X_test"] = X_test
    out[name + "/hf_X_final"] = model.hf_X
    out[name + "/hf_Y_final"] = model.hf_Y
    out[name + "/aug_X"] = model.hf_model.X
    out[name + "/theta"] = model.hf_model._o.theta.copy()
    out[name + "/lml"] = np.array(model.hf_model.log_likelihood())
    out[name + "/mean"] = mean
    out[name + "/var"] = var
    out[name + "/adapt_steps"] = np.array(getattr(model, "adapt_steps", 0))
    # the same trained object at fixed, well-conditioned hyper-parameters (the optimiser's own end point
    # has noise -> 0 on these tiny noiseless sets, cond(K_y) ~ 1e12): the reference's predict path
    # (augmentation, add_noise, GP.predict) for the CUDA parity test at 1e-8 / 1e-6
    theta_fixed = THETA_C if model.hf_model._o.kind == 1 else THETA_R
    model.hf_model._o.theta[:] = theta_fixed
    model.hf_model._o._post = None
    mean_f, var_f = model.predict(X_test)
    out[name + "/theta_fixed"] = theta_fixed
    out[name + "/lml_fixed"] = np.array(model.hf_model.log_likelihood())
    out[name + "/mean_fixed"] = mean_f
    out[name + "/var_fixed"] = var_f
    if model.data_driven_lf_approach:
        out[name + "/lf_theta"] = model.lf_model._o.theta.copy()
    print("%-28s N_h=%d D=%d lml=%.6f theta=%s" % (name, model.hf_X.shape[0], model.hf_model.X.shape[1],
                                                  out[name + "/lml"], np.round(out[name + "/theta"], 4)))


if __name__ == "__main__":
    out = {}
    rs = np.random.RandomState(10)                       # tests/utils.py:11,30-35
    X_lf2, X_hf2, X_test2 = rs.uniform(size=(60, 2)), rs.uniform(size=(8, 2)), rs.uniform(size=(50, 2))
    # config 1: NARGP 1-D, callable low fidelity (src/data/exampleCurves1D.py:10-13)
    hf_X1 = np.linspace(0, 1, 10)[:, None]
    Xt1 = np.linspace(0, 1, 101)[:, None]
    scenario("nargp_1d", ref_models.NARGP(1, util.f_high_1d, util.f_low_1d), hf_X1, Xt1, 1, out)
    # config 2: GPDF / GPDFC 2-D with delays (tests/test_mfgp_adapt_2d.py:27)
    scenario("gpdf_2d", ref_models.GPDF(2, 0.001, 2, util.hf_2d, util.lf_2d), X_hf2, X_test2, 2, out)
    scenario("gpdfc_2d", ref_models.GPDFC(2, 0.001, 2, util.hf_2d, util.lf_2d), X_hf2, X_test2, 3, out)
    # add_noise: re-inference at noise 1e-6 inside predict (src/MFDataFusion.py:154-155)
    scenario("gpdf_2d_add_noise", ref_models.GPDF(2, 0.001, 2, util.hf_2d, util.lf_2d, add_noise=True),
             X_hf2, X_test2, 4, out)
    # data-driven low fidelity: LF GP trained in the constructor (src/abstractMFGP.py:96-104)
    np.random.seed(5)
    m = ref_models.NARGP(2, util.hf_2d, None, lf_X=X_lf2, lf_Y=util.lf_2d(X_lf2))
    scenario("nargp_2d_data_driven", m, X_hf2, X_test2, 5, out)
    # adaptation loop with the reference's DIRECT wrapper (2 steps, refit after each)
    scenario("nargp_2d_adapt", ref_models.NARGP(2, util.hf_2d, util.lf_2d), X_hf2[:5], X_test2, 6, out,
             adapt_steps=2)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_runs.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%d arrays" % len(out))
