"""Plug-in protocol of the gPC layer.

A gPC object wraps a callable ``function`` (nodes (Q, d) -> values) and answers four questions about its
polynomial-chaos surrogate; ``MFGP_GPC`` drives any object that implements them.  The method names are
the reference's (src/gpc/gpc_abstract.py) so that its wrappers and ours are interchangeable:

``calculate_coefficients()``  evaluate ``function`` on the quadrature nodes and project
``get_mean()``, ``get_var()`` statistics of the surrogate
``update_order(new_order)``   change polynomial and quadrature order (coefficients become stale)
``update_function(f)``        swap the callable and recompute
``get_mean_var()``            both statistics as a tuple
"""
from abc import ABC, abstractmethod


class AbstractGPC(ABC):
    """Base class: stores the callable, leaves the four surrogate operations to the subclass."""

    function = None

    def __init__(self, function):
        self.function = function

    def update_function(self, function):
        """Point the surrogate at another callable and re-project immediately."""
        self.function = function
        self.calculate_coefficients()

    def get_mean_var(self):
        return (self.get_mean(), self.get_var())

    @abstractmethod
    def calculate_coefficients(self):
        """Evaluate ``self.function`` at the quadrature nodes and compute the expansion coefficients."""

    @abstractmethod
    def get_mean(self):
        """Mean of the surrogate under the input distribution."""

    @abstractmethod
    def get_var(self):
        """Variance of the surrogate under the input distribution."""

    @abstractmethod
    def update_order(self, new_order):
        """Use ``new_order`` for both the polynomial and the quadrature order."""
