from ._base import _FlowBase


class RadialFlow(_FlowBase):
    """x = y + alpha*beta*(y - gamma) / (alpha + |y - gamma|_1); drop-in for the reference's
    estimators/normalizing_flows/RadialFlow.py:6-84 (params [alpha_raw, beta_raw, gamma(d)])."""

    flow_type = "radial"

    def __init__(self, t, n_dims, name="RadialFlow"):
        super().__init__(t, n_dims, name)

    @staticmethod
    def get_param_size(n_dims):
        return 1 + 1 + n_dims
