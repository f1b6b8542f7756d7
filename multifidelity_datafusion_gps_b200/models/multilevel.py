"""Recursive NARGP over more than two fidelity levels.

The reference stops at two levels (src/MFDataFusion.py:56-73: one low-fidelity source), but its NARGP model
is the two-level case of the recursion of the paper it cites (README.md:13, Perdikaris et al. 2017,
eq. 2.9-2.10): level 1 is a plain GP on (X_1, Y_1); level t > 1 is the reference's NARGP -- composite
kernel on [x, mu_{t-1}(x)] -- whose low-fidelity function is the posterior mean of level t-1.  This class
chains the package's own ``NARGP`` objects that way; every level's mean is evaluated on the GPU and handed
to the next level through the ``device_predict`` hook of ``MultifidelityDataFusion._augment_device``.
"""
import numpy as np

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
