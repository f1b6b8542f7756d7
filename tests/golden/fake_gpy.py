"""TEST INFRASTRUCTURE -- a GPy look-alike backed by the CPU oracle, so that the REFERENCE's own
orchestration code (/root/reference/src: MFDataFusion, abstractMFGP, models, iterators, maximizers)
can be executed in the build container where GPy 1.9.9 / paramz / DIRECT / scipydirect / matplotlib
are not installable.  Only the GP arithmetic and the paramz protocol come from oracle/gpy_oracle.py;
augmentation, kernel composition, the ARD recipe, predict (incl. add_noise) and the adaptation loop
are the reference's code, run unmodified.  Used by make_reference_run_golden.py only.

API surface = exactly what the reference touches:
  GPy.kern.RBF(input_dim, active_dims=None), kernel * kernel, kernel + kernel, kernel.to_dict()
  GPy.models.GPRegression(X, Y, kernel=None, initialize=True): [".*Gaussian_noise"] get/set, .fix(),
  .unfix(), .constrain_positive(), .optimize(max_iters=), .optimize_restarts(n, optimizer=, max_iters=,
  verbose=), .predict(X), .Y, .likelihood.variance (settable)
"""
import re
import sys
import types

import numpy as np

from oracle import gpy_oracle as go


class _Kern:
    def __mul__(self, other):
        return _Prod([self, other])

    def __add__(self, other):
        return _Add([self, other])


class RBF(_Kern):
    def __init__(self, input_dim, variance=1.0, lengthscale=None, ARD=False, active_dims=None, name="rbf"):
        assert not ARD, "the reference never sets ARD=True (SURVEY.md fact 3)"
        self.input_dim = int(input_dim)
        self.active_dims = np.arange(input_dim) if active_dims is None else np.asarray(active_dims)
        self.params = np.array([variance, 1.0 if lengthscale is None else lengthscale], dtype=np.float64)

    def leaves(self):
        return [self]

    def to_dict(self):
        return {"class": "GPy.kern.RBF", "variance": [self.params[0]], "lengthscale": [self.params[1]]}


class _Prod(_Kern):
    def __init__(self, parts):
        self.parts = parts

    def leaves(self):
        return [l for p in self.parts for l in p.leaves()]

    def to_dict(self):
        return {"class": "GPy.kern.Prod", "parts": {i: p.to_dict() for i, p in enumerate(self.parts)}}


class _Add(_Kern):
    def __init__(self, parts):
        self.parts = parts

    def leaves(self):
        return [l for p in self.parts for l in p.leaves()]

    def to_dict(self):
        return {"class": "GPy.kern.Add", "parts": {i: p.to_dict() for i, p in enumerate(self.parts)}}


def _classify(kernel, D):
    """-> (oracle kind, d, leaves in oracle parameter order)."""
    if isinstance(kernel, RBF):
        assert len(kernel.active_dims) == D
        return go.KIND_RBF, D, [kernel]
    # RBF(aug) * RBF(std) + RBF(std)    (src/abstractMFGP.py:76-80)
    assert isinstance(kernel, _Add) and isinstance(kernel.parts[0], _Prod) and isinstance(kernel.parts[1], RBF)
    k1, k2 = kernel.parts[0].parts
    k3 = kernel.parts[1]
    d = len(k2.active_dims)
    assert np.array_equal(k2.active_dims, np.arange(d)) and np.array_equal(k3.active_dims, np.arange(d))
    assert np.array_equal(k1.active_dims, np.arange(d, D))
    return go.KIND_COMPOSITE, d, [k1, k2, k3]


class _Likelihood:
    def __init__(self, model):
        self._m = model

    @property
    def variance(self):
        return self._m._o.theta[-1]

    @variance.setter
    def variance(self, v):
        self._m._o.theta[-1] = float(v)
        self._m._o._post = None


class _ParamHandle:
    def __init__(self, model, idx):
        self._m, self._idx = model, idx

    def fix(self):
        self._m._o.fixed[self._idx] = True

    def unfix(self):
        self._m._o.fixed[self._idx] = False

    def constrain_positive(self):
        pass   # every parameter already lives under the Logexp transform


class GPRegression:
    def __init__(self, X, Y, kernel=None, initialize=True, **kw):
        X, Y = np.asarray(X, dtype=np.float64), np.asarray(Y, dtype=np.float64)
        if kernel is None:
            kernel = RBF(X.shape[1])               # GPRegression's default kernel
        self.kern = kernel
        kind, d, self._leaves = _classify(kernel, X.shape[1])
        theta = np.concatenate([l.params for l in self._leaves] + [[1.0]])   # Gaussian_noise default 1.0
        self._o = go.OracleGPRegression(X, Y, kind, d=d, theta=theta)
        self.X, self.Y = self._o.X, self._o.Y
        self.likelihood = _Likelihood(self)
        self.names = []
        for i, _ in enumerate(self._leaves):
            self.names += ["k%d.variance" % i, "k%d.lengthscale" % i]
        self.names += ["Gaussian_noise.variance"]

    def _match(self, pattern):
        rx = re.compile(pattern)
        return np.array([bool(rx.match(n)) for n in self.names])

    def __getitem__(self, pattern):
        return _ParamHandle(self, self._match(pattern))

    def __setitem__(self, pattern, value):
        self._o.theta[self._match(pattern)] = float(np.ravel(value)[0])
        self._o._post = None

    def _write_back(self):     # GPy links the kernel object's parameters into the model: fits warm-start
        for i, l in enumerate(self._leaves):
            l.params[:] = self._o.theta[2 * i:2 * i + 2]

    def optimize(self, optimizer=None, max_iters=1000, messages=False, **kw):
        self._o.optimize(max_iters=max_iters)
        self._write_back()

    def optimize_restarts(self, num_restarts=10, robust=False, verbose=True, parallel=False,
                          num_processes=None, **kw):
        self._o.optimize_restarts(num_restarts, max_iters=kw.get("max_iters", 1000))
        self._write_back()

    def predict(self, Xnew, full_cov=False, **kw):
        return self._o.predict(np.asarray(Xnew, dtype=np.float64))

    def log_likelihood(self):
        return self._o.log_likelihood()


class _DirectResult:
    def __init__(self, x, fun):
        self.x, self.fun = x, fun


def direct_minimize(func, bounds, **kw):
    """scipydirect.minimize(func, bounds) stand-in: SciPy's DIRECT with a fixed evaluation budget."""
    import scipy.optimize as sopt
    res = sopt.direct(lambda x: float(np.ravel(func(np.asarray(x)))[0]), list(bounds), maxfun=150, maxiter=1000)
    return _DirectResult(np.asarray(res.x), float(res.fun))


def install():
    """Register the stand-ins under the names the reference imports."""
    gpy = types.ModuleType("GPy")
    gpy.kern = types.ModuleType("GPy.kern")
    gpy.kern.RBF = RBF
    gpy.models = types.ModuleType("GPy.models")
    gpy.models.GPRegression = GPRegression
    sys.modules.update({"GPy": gpy, "GPy.kern": gpy.kern, "GPy.models": gpy.models})
    plt = types.ModuleType("matplotlib.pyplot")
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    direct = types.ModuleType("DIRECT")
    direct.solve = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("DIRECT.solve is not on the tested path"))
    sys.modules["DIRECT"] = direct
    sd = types.ModuleType("scipydirect")
    sd.minimize = direct_minimize
    sys.modules["scipydirect"] = sd
