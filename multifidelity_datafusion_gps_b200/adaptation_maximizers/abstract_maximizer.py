"""Plug-in point of the acquisition step (reference src/adaptation_maximizers/abstract_maximizer.py:5-28)."""
from abc import ABCMeta, abstractmethod

import numpy as np


class AbstractMaximizer(metaclass=ABCMeta):
    """Wrapper for the uncertainty maximisation of the adaptation process."""

    @abstractmethod
    def __init__(self):
        super().__init__()

    @abstractmethod
    def maximize(self, model_predict: callable, lower_bound: np.ndarray, upper_bound: np.ndarray):
        """Return ``(x, fopt)``: the input in [lower_bound, upper_bound] with the largest predictive
        variance and the NEGATED variance there (the reference minimises ``-variance``,
        src/adaptation_maximizers/scipydirect_wrapper.py:22-31)."""
