from ._base import _FlowBase


class PlanarFlow(_FlowBase):
    """x = y + u_hat * tanh(w^T y + b); drop-in for the reference's
    estimators/normalizing_flows/PlanarFlow.py:7-80 (params [u(d), w_raw(d), b])."""

    flow_type = "planar"

    def __init__(self, t, n_dims, name="Inverted_Planar_Flow"):
        super().__init__(t, n_dims, name)

    @staticmethod
    def get_param_size(n_dims):
        return n_dims + n_dims + 1
