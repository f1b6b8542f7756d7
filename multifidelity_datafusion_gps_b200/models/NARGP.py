"""NARGP preset (reference src/models/NARGP.py:15-21): no delays (num_derivatives=0, tau=0),
composite kernel k1(f_low(x), f_low(x')) k2(x, x') + k3(x, x')."""
import numpy as np

from ..MFDataFusion import MultifidelityDataFusion


class NARGP(MultifidelityDataFusion):
    def __init__(self, input_dim: int, f_exact: callable, f_low: callable, name: str = 'NARGP',
                 lower_bound: np.ndarray = None, upper_bound: np.ndarray = None, lf_X: np.ndarray = None,
                 lf_Y: np.ndarray = None, lf_hf_adapt_ratio: int = 1, eps: float = 1e-8,
                 add_noise: bool = False, adapt_maximizer=None):
        super().__init__(name=name, input_dim=input_dim, num_derivatives=0, tau=0, f_exact=f_exact,
                         lower_bound=lower_bound, upper_bound=upper_bound, f_low=f_low, lf_X=lf_X,
                         lf_Y=lf_Y, lf_hf_adapt_ratio=lf_hf_adapt_ratio, use_composite_kernel=True,
                         eps=eps, add_noise=add_noise, adapt_maximizer=adapt_maximizer)
