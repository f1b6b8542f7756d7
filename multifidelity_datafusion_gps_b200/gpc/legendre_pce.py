"""Pseudo-spectral PCE for independent uniform inputs, projected on the GPU (kernel K9).

Stands where the reference has ``ChaospyWrapper`` (src/gpc/chaospy_wrapper.py:9-37): the same
protocol (``calculate_coefficients``, ``get_mean``, ``get_var``, ``update_order``,
``update_function``) for the distribution the reference uses, ``cp.J(cp.Uniform(0, 1), ...)``
(tests/test_mfgp_adapt_4d.py:40), with chaospy's conventions:

* ``quadrature_order=q`` -> tensor Gauss-Legendre rule with q + 1 nodes per dimension
  (``cp.generate_quadrature(q, dist, rule="gaussian")``),
* ``polynomial_order=p`` -> all multi-indices of total degree <= p (``cp.generate_expansion``),
* coefficients by discrete projection (``cp.fit_quadrature``); with the orthonormal Legendre basis
  mean = c_0 and variance = sum_{k>0} c_k^2 (what ``cp.E`` / ``cp.Var`` of the fitted expansion return).

``function`` maps (Q, d) -> (Q,) or (Q, 1) like the reference's ``temp_f = lambda x: model.predict(x)[0]``.
If it is a model of this package (or ``model.predict`` / a callable with attribute ``device_predict``)
the evaluations stay on the device: nodes -> K6/K7 -> K9 without a host round trip.
"""
import ctypes
import itertools

import numpy as np
import torch

from .. import _ffi, gp
from .gpc_abstract import AbstractGPC


def total_degree_multi_index(dim, order):
    """All multi-indices with |k|_1 <= order, graded (by total degree), then reverse-lexicographic;
    the all-zero index comes first.  (P, dim) int32."""
    out = []
    for total in range(order + 1):
        level = [k for k in itertools.product(range(total + 1), repeat=dim) if sum(k) == total]
        out.extend(sorted(level, reverse=True))
    return np.asarray(out, dtype=np.int32).reshape(-1, dim)


def gauss_legendre_tensor_grid(n_per_dim, lower, upper):
    """Tensor Gauss-Legendre rule on the box: nodes (n^d, d), weights (n^d,) summing to 1."""
    lower, upper = np.asarray(lower, dtype=np.float64), np.asarray(upper, dtype=np.float64)
    t, w = np.polynomial.legendre.leggauss(n_per_dim)
    d = lower.shape[0]
    axes = [lower[i] + (t + 1.0) * 0.5 * (upper[i] - lower[i]) for i in range(d)]
    nodes = np.stack([g.ravel() for g in np.meshgrid(*axes, indexing="ij")], axis=1)
    wts = np.ones(1)
    for _ in range(d):
        wts = np.multiply.outer(wts, 0.5 * w).ravel() if wts.size > 1 else 0.5 * w
    return np.ascontiguousarray(nodes), np.ascontiguousarray(wts)


class LegendrePCE(AbstractGPC):

    def __init__(self, function, lower_bound, upper_bound, polynomial_order=8, quadrature_order=8, device=None):
        self.lower_bound = np.asarray(lower_bound, dtype=np.float64).ravel()
        self.upper_bound = np.asarray(upper_bound, dtype=np.float64).ravel()
        assert self.lower_bound.shape == self.upper_bound.shape and np.all(self.upper_bound > self.lower_bound)
        self.dim = self.lower_bound.shape[0]
        # the device of the wrapped model (one process per GPU: rank r's model lives on cuda:LOCAL_RANK),
        # else torch's current device; nodes and weights must sit where the model's handle runs
        owner = getattr(function, "__self__", function)
        if device is None:
            device = getattr(owner, "device", None)
        self.device = gp.current_device() if device is None else int(device)
        self.coefficients = None
        self._set_order(polynomial_order, quadrature_order)
        super().__init__(function)

    def _set_order(self, polynomial_order, quadrature_order):
        self.polynomial_order, self.quadrature_order = int(polynomial_order), int(quadrature_order)
        self.quad_points, self.quad_weights = gauss_legendre_tensor_grid(
            self.quadrature_order + 1, self.lower_bound, self.upper_bound)
        self.multi_index = total_degree_multi_index(self.dim, self.polynomial_order)
        self._d_nodes = gp.to_device(self.quad_points, self.device)
        self._d_wts = gp.to_device(self.quad_weights, self.device)

    def update_order(self, new_order):
        self._set_order(new_order, new_order)

    def _evaluate_on_device(self):
        fn = self.function
        owner = getattr(fn, "__self__", None)
        owner_dev = getattr(owner if owner is not None else fn, "device", None)
        assert owner_dev is None or int(owner_dev) == self.device, \
            "quadrature nodes live on cuda:%d but the model on cuda:%s" % (self.device, owner_dev)
        if hasattr(fn, "device_predict"):
            return fn.device_predict(self._d_nodes)
        for obj in (fn, owner):
            if obj is not None and hasattr(obj, "_predict_device") and getattr(fn, "__name__", "predict") == "predict":
                return obj._predict_device(self._d_nodes)[0]
        vals = fn(self.quad_points)
        if isinstance(vals, tuple):                      # model.predict returns (mean, variance)
            vals = vals[0]
        vals = np.asarray(vals, dtype=np.float64)
        assert vals.size == self.quad_points.shape[0], "function must return one value per node"
        return gp.to_device(np.ascontiguousarray(vals.reshape(-1)), self.device)

    def project(self, d_values):
        """Coefficients (P,) of device values (Q,) at the quadrature nodes."""
        h = _ffi.get_handle(self.device)
        Q, P = self.quad_points.shape[0], self.multi_index.shape[0]
        assert d_values.numel() == Q
        d_values = d_values.reshape(-1).contiguous()
        ws_bytes = h.lib.mfgp_pce_ws_bytes(self.dim, P)
        ws = gp.workspace(self.device, ws_bytes)
        d_coeff = torch.empty(P, dtype=torch.float64, device=d_values.device)
        coeff = np.empty(P)
        mi = np.ascontiguousarray(self.multi_index, dtype=np.int32)
        h.check(h.lib.mfgp_pce_project(
            h.h, self._d_nodes.data_ptr(), self.lower_bound.ctypes.data_as(ctypes.c_void_p),
            self.upper_bound.ctypes.data_as(ctypes.c_void_p), self.dim, self._d_wts.data_ptr(),
            d_values.data_ptr(), Q, mi.ctypes.data_as(ctypes.c_void_p), P, self.polynomial_order,
            d_coeff.data_ptr(), coeff.ctypes.data_as(ctypes.c_void_p), ws.data_ptr(), ws.numel() * 8))
        return coeff

    def calculate_coefficients(self):
        self.coefficients = self.project(self._evaluate_on_device())

    def get_mean(self):
        return float(self.coefficients[0])

    def get_var(self):
        return float(np.sum(self.coefficients[1:] ** 2))
