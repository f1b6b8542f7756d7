"""GPDFC preset (reference src/models/GPDFC.py:16-22): delays plus the composite NARGP kernel.
``lengthscale_hyperparams`` returns what the reference's plot helper reads (:25-34), without drawing."""
import numpy as np

from ..MFDataFusion import MultifidelityDataFusion


class GPDFC(MultifidelityDataFusion):
    def __init__(self, input_dim: int, tau: float, num_derivatives: int, f_exact: callable, f_low: callable,
                 name: str = 'GPDFC', lower_bound: np.ndarray = None, upper_bound: np.ndarray = None,
                 lf_X: np.ndarray = None, lf_Y: np.ndarray = None, lf_hf_adapt_ratio: int = 1,
                 eps: float = 1e-8, add_noise: bool = False, adapt_maximizer=None):
        super().__init__(name=name, input_dim=input_dim, num_derivatives=num_derivatives, tau=tau,
                         f_exact=f_exact, lower_bound=lower_bound, upper_bound=upper_bound, f_low=f_low,
                         lf_X=lf_X, lf_Y=lf_Y, lf_hf_adapt_ratio=lf_hf_adapt_ratio,
                         use_composite_kernel=True, eps=eps, add_noise=add_noise,
                         adapt_maximizer=adapt_maximizer)

    def lengthscale_hyperparams(self):
        kern = self.kernel.to_dict()
        l1 = kern["parts"][1]["lengthscale"][0]
        l2 = kern["parts"][0]["parts"][0]["lengthscale"][0]
        l3 = kern["parts"][0]["parts"][1]["lengthscale"][0]
        return l1, l2, l3
