"""Recursive NARGP over more than two fidelity levels.

The reference stops at two levels (src/MFDataFusion.py:56-73: one low-fidelity source), but its NARGP model
is the two-level case of the recursion of the paper it cites (README.md:13, Perdikaris et al. 2017,
eq. 2.9-2.10): level 1 is a plain GP on (X_1, Y_1); level t > 1 is the reference's NARGP -- composite
kernel on [x, mu_{t-1}(x)] -- whose low-fidelity function is the posterior mean of level t-1.  This class
chains the package's own ``NARGP`` objects that way; every level's mean is evaluated on the GPU and handed
to the next level through the ``device_predict`` hook of ``MultifidelityDataFusion._augment_device``.
"""
import ctypes

import numpy as np
import torch

from .. import _ffi, gp
from ._presets import NARGP


class _LevelMean:
    """f_low of level t+1: posterior mean of level t (host callable, with the device hook)."""

    def __init__(self, model):
        self.model = model

    def __call__(self, X):
        return self.model.predict(np.atleast_2d(X))[0]

    def device_predict(self, dX):
        return self.model._predict_device(dX)[0]


class _Targets:
    """f_exact of a data-driven level: the level's own targets at its own training inputs."""

    def __init__(self, X, Y):
        self.X, self.Y = np.asarray(X, dtype=np.float64), np.asarray(Y, dtype=np.float64).reshape(-1, 1)

    def __call__(self, X):
        assert np.array_equal(np.asarray(X), self.X), "a data-driven level is fitted on its own inputs"
        return self.Y


class MultiLevelNARGP:
    def __init__(self, input_dim, level_X, level_Y, add_noise=False):
        assert len(level_X) == len(level_Y) >= 2
        self.input_dim = input_dim
        self.level_X = [np.asarray(X, dtype=np.float64) for X in level_X]
        self.level_Y = [np.asarray(Y, dtype=np.float64).reshape(-1, 1) for Y in level_Y]
        self.add_noise = add_noise
        self.models = []

    def fit(self, thetas=None, lf_theta=None):
        """Fit the levels bottom-up.  ``thetas`` (optional): one hyper-parameter vector per level >= 2
        (skips the optimiser, for parity tests); ``lf_theta``: the level-1 GP's (variance, lengthscale, noise)."""
        self.models = []
        for t in range(1, len(self.level_X)):
            f_exact = _Targets(self.level_X[t], self.level_Y[t])
            if t == 1:
                m = NARGP(self.input_dim, f_exact, None, name="level2", lf_X=self.level_X[0],
                          lf_Y=self.level_Y[0], add_noise=self.add_noise)
                if lf_theta is not None:
                    m.lf_model._set_params(np.asarray(lf_theta, dtype=np.float64))
            else:
                m = NARGP(self.input_dim, f_exact, _LevelMean(self.models[-1]), name="level%d" % (t + 1),
                          add_noise=self.add_noise)
            m.fit(self.level_X[t], theta=None if thetas is None else thetas[t - 1])
            self.models.append(m)
        return self

    def predict(self, X_test):
        """(mean (M,1), variance (M,1)) of the top level; lower levels enter through their means."""
        return self.models[-1].predict(X_test)

    def predict_level(self, t, X_test):
        """Prediction of fidelity level t (2 .. L)."""
        return self.models[t - 2].predict(X_test)

    # -- Monte-Carlo propagation through every level -------------------------------------------------
    def predict_mc_device(self, dX, n_samples=100, d_eps=None, seed=0, m0=0, d_weights=None,
                          include_lower_noise=True):
        """Sample-based propagation through ALL levels (Perdikaris et al. 2017, section 2(c)): per sample,
        z_1 ~ N(mu_1(x), v_1(x)) at level 1, then (mu_t, v_t) = level t at [x, z_{t-1}] and
        z_t ~ N(mu_t, v_t) up the chain; mean = mean_s mu_L, var = mean_s v_L + var_s mu_L.
        dX (M, d) CUDA -> (mean (M,), var (M,), weighted sum or None).  d_eps: optional (L-1, M, S) standard
        normals; otherwise Philox (one key per level, counter = global point index * S + sample, so shards
        with m0 reproduce the full batch).  With two levels this is NARGP.predict_mc."""
        top = self.models[-1]
        h = _ffi.get_handle(top.device)
        gps = [self.models[0].lf_model] + [m.hf_model for m in self.models]
        for m in self.models:
            m._apply_add_noise()
        structs = [g.level_struct() for g in gps]
        L = len(structs)
        ptrs = (ctypes.c_void_p * L)(*[ctypes.addressof(st) for st in structs])
        M, S = int(dX.shape[0]), int(n_samples)
        mean = torch.empty(M, dtype=torch.float64, device=dX.device)
        var = torch.empty(M, dtype=torch.float64, device=dX.device)
        wsum = ctypes.c_double(0.0) if d_weights is not None else None
        ws = gp.workspace(top.device, h.lib.mfgp_predict_mc_ws_bytes(gps[0].N, max(g.N for g in gps[1:]),
                                                                     self.input_dim, M, S))
        h.check(h.lib.mfgp_predict_mc_chain(
            h.h, ctypes.cast(ptrs, ctypes.c_void_p), L, dX.data_ptr(), M, S,
            d_eps.data_ptr() if d_eps is not None else None, int(seed), int(m0), int(include_lower_noise), 1,
            d_weights.data_ptr() if d_weights is not None else None, mean.data_ptr(), var.data_ptr(),
            ctypes.byref(wsum) if wsum is not None else None, ws.data_ptr(), ws.numel() * 8))
        return mean, var, (wsum.value if wsum is not None else None)

    def predict_mc(self, X_test, n_samples=100, eps=None, seed=0, weights=None, m0=0, include_lower_noise=True):
        """NumPy front end of predict_mc_device: (mean (M,1), var (M,1)); eps: optional (L-1, M, S)."""
        X_test = np.ascontiguousarray(X_test, dtype=np.float64)
        assert X_test.ndim == 2 and X_test.shape[1] == self.input_dim
        dev = self.models[-1].device
        d_eps = None
        if eps is not None:
            eps = np.asarray(eps, dtype=np.float64).reshape(len(self.models), X_test.shape[0], n_samples)
            d_eps = gp.to_device(eps, dev)
        d_w = gp.to_device(np.asarray(weights).ravel(), dev) if weights is not None else None
        mean, var, wsum = self.predict_mc_device(gp.to_device(X_test, dev), n_samples, d_eps, seed, m0, d_w,
                                                 include_lower_noise)
        self.last_pce_mean = wsum
        return mean.cpu().numpy()[:, None], var.cpu().numpy()[:, None]
