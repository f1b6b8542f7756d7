"""Shared case generators for the parity tests (seeded; the same inputs feed oracle and CUDA)."""
import numpy as np

from oracle import gpy_oracle as go

# -- the reference's own test functions, restated -------------------------------------------------
A2 = [2.2 * np.pi, np.pi]                       # tests/test_mfgp_adapt_2d.py:9


def hf_2d(x):                                    # tests/test_mfgp_adapt_2d.py:12-14
    x = np.atleast_2d(x)
    return (np.sin(x[:, 0] * A2[0]) * np.sin(x[:, 1] * A2[1]))[:, None]


def lf_2d(x):                                    # tests/test_mfgp_adapt_2d.py:17-19
    x = np.atleast_2d(x)
    return hf_2d(x) - 1.2 * (np.sin(x[:, 0] * np.pi * 0.1) + np.sin(x[:, 1] * np.pi * 0.1))[:, None]


def hf_4d(x):                                    # tests/test_mfgp_adapt_4d.py:13-15
    x = np.atleast_2d(x)
    return (np.prod(np.sin(x[:, :4] * np.pi), axis=1) + 5.0)[:, None]


def lf_4d(x):                                    # tests/test_mfgp_adapt_4d.py:18-21
    x = np.atleast_2d(x)
    return hf_4d(x) - 0.25 * (np.sin(x[:, 0] * np.pi * 0.1) + np.sin(x[:, 1] * np.pi * 0.05)
                              + np.sin(x[:, 2] * 0.15 * np.pi) + np.sin(x[:, 3] * 0.2 * np.pi))[:, None]


def f_low_1d(t):                                 # src/data/exampleCurves1D.py:11
    return np.sin(8 * np.pi * t)


def f_high_1d(t):                                # src/data/exampleCurves1D.py:12
    return np.sin(8 * np.pi * t) ** 2


def random_case(seed, N, d, E, kind, noise=1e-2):
    """Random inputs in [0,1]^(d+E) with smooth targets and moderate hyper-parameters."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(size=(N, d + E))
    Y = (np.sin(3.0 * X.sum(axis=1)) + 0.1 * rng.standard_normal(N))[:, None]
    if kind == go.KIND_COMPOSITE:
        theta = np.array([1.3, 0.7, 0.9, 0.5, 0.2, 0.4, noise])
    else:
        theta = np.array([1.1, 0.6, noise])
    return X, Y, theta


def rel_err(a, b, scale=None):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    s = np.max(np.abs(b)) if scale is None else scale
    return float(np.max(np.abs(a - b)) / max(s, 1e-300))
