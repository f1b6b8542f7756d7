"""The reference's three model presets as thin subclasses of ``MultifidelityDataFusion``.

Each preset only pins a few constructor arguments of the base class; everything else is passed through
under the reference's parameter names, so existing call sites keep working:

=========  ==========================================  =====================================================
preset     reference                                   pins
=========  ==========================================  =====================================================
``NARGP``  src/models/NARGP.py:15-21                   no delays (``num_derivatives=0``, ``tau=0``), composite kernel
``GPDF``   src/models/GPDF.py:15-21                    delays as given, one RBF kernel over all augmented columns
``GPDFC``  src/models/GPDFC.py:16-22                   delays as given, composite kernel
=========  ==========================================  =====================================================
"""
from ..MFDataFusion import MultifidelityDataFusion


def _forward(self, pinned, name, input_dim, f_exact, f_low, lower_bound, upper_bound, lf_X, lf_Y,
             lf_hf_adapt_ratio, eps, add_noise, adapt_maximizer):
    """One place for the pass-through to the base constructor (keyword for keyword)."""
    MultifidelityDataFusion.__init__(
        self, name=name, input_dim=input_dim, f_exact=f_exact, f_low=f_low,
        lower_bound=lower_bound, upper_bound=upper_bound, lf_X=lf_X, lf_Y=lf_Y,
        lf_hf_adapt_ratio=lf_hf_adapt_ratio, eps=eps, add_noise=add_noise,
        adapt_maximizer=adapt_maximizer, **pinned)


class NARGP(MultifidelityDataFusion):
    """k1(f_low(x), f_low(x')) k2(x, x') + k3(x, x') on [x, f_low(x)]: no delayed evaluations."""

    def __init__(self, input_dim, f_exact, f_low, name='NARGP', lower_bound=None, upper_bound=None,
                 lf_X=None, lf_Y=None, lf_hf_adapt_ratio=1, eps=1e-8, add_noise=False, adapt_maximizer=None):
        _forward(self, dict(num_derivatives=0, tau=0, use_composite_kernel=True), name, input_dim, f_exact,
                 f_low, lower_bound, upper_bound, lf_X, lf_Y, lf_hf_adapt_ratio, eps, add_noise, adapt_maximizer)


class GPDF(MultifidelityDataFusion):
    """Delay-augmented inputs [x, f_low(x), f_low(x - tau e_1), ...] under one RBF kernel."""

    def __init__(self, input_dim, tau, num_derivatives, f_exact, f_low, name='GPDF', lower_bound=None,
                 upper_bound=None, lf_X=None, lf_Y=None, lf_hf_adapt_ratio=1, eps=1e-8, add_noise=False,
                 adapt_maximizer=None):
        _forward(self, dict(num_derivatives=num_derivatives, tau=tau, use_composite_kernel=False), name,
                 input_dim, f_exact, f_low, lower_bound, upper_bound, lf_X, lf_Y, lf_hf_adapt_ratio, eps,
                 add_noise, adapt_maximizer)


class GPDFC(MultifidelityDataFusion):
    """Delay-augmented inputs under the composite NARGP kernel."""

    def __init__(self, input_dim, tau, num_derivatives, f_exact, f_low, name='GPDFC', lower_bound=None,
                 upper_bound=None, lf_X=None, lf_Y=None, lf_hf_adapt_ratio=1, eps=1e-8, add_noise=False,
                 adapt_maximizer=None):
        _forward(self, dict(num_derivatives=num_derivatives, tau=tau, use_composite_kernel=True), name,
                 input_dim, f_exact, f_low, lower_bound, upper_bound, lf_X, lf_Y, lf_hf_adapt_ratio, eps,
                 add_noise, adapt_maximizer)

    def lengthscale_hyperparams(self):
        """(l of k3, l of k1, l of k2): the numbers the reference's bar-chart helper reads from
        ``kernel.to_dict()`` (src/models/GPDFC.py:25-34), returned instead of drawn."""
        parts = self.kernel.to_dict()["parts"]
        product = parts[0]["parts"]
        return (parts[1]["lengthscale"][0], product[0]["lengthscale"][0], product[1]["lengthscale"][0])
