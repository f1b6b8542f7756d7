"""Adaptive multi-fidelity gPC loop.

Alternates rounds of model adaptation with re-projections of the model mean and keeps the histories the
reference's driver keeps under the same attribute names (src/gpc/mfgp_gpc.py): ``mean_history``,
``var_history``, ``cost_history`` (cumulative number of high-fidelity evaluations) and, when test data are
given, ``mse_history``.  One round = ``adapt_per_steps`` (5) acquisitions; ``num_adapts`` rounds per call.
"""


class MFGP_GPC(object):

    adapt_per_steps = 5

    def __init__(self, mfgp_obj, gpc_obj, num_adapts, init_cost, X_test=None, Y_test=None):
        self.mfgp_obj = mfgp_obj
        self.gpc_obj = gpc_obj
        self.num_adapts = num_adapts
        self.X_test, self.Y_test = X_test, Y_test
        self.calculate_mse = X_test is not None and Y_test is not None
        self.mean_history, self.var_history, self.cost_history = [], [], []
        if self.calculate_mse:
            self.mse_history = []
        self.gpc_obj.calculate_coefficients()
        self._record(init_cost)

    def _record(self, cost):
        """Append the current surrogate statistics, the cumulative cost and (optionally) the test MSE."""
        mean, var = self.gpc_obj.get_mean_var()
        self.mean_history.append(mean)
        self.var_history.append(var)
        self.cost_history.append(cost)
        if self.calculate_mse:
            self.mse_history.append(self.mfgp_obj.get_mse(self.X_test, self.Y_test))

    def adapt(self):
        for _ in range(self.num_adapts):
            self.mfgp_obj.adapt(self.adapt_per_steps)
            # the bound method, not a lambda around it: a gPC object of this package then keeps the
            # evaluations on the device
            self.gpc_obj.update_function(self.mfgp_obj.predict)
            self._record(self.cost_history[-1] + self.mfgp_obj.adapt_steps)
