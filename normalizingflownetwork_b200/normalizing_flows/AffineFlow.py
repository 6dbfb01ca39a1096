from ._base import _FlowBase


class AffineFlow(_FlowBase):
    """x = (1 + scale_raw) * y + shift; drop-in for the reference's
    estimators/normalizing_flows/AffineFlow.py:4-17 (params [shift(d), scale_raw(d)])."""

    flow_type = "affine"

    def __init__(self, t, n_dims, name="AffineFlow"):
        super().__init__(t, n_dims, name)

    @staticmethod
    def get_param_size(n_dims):
        return 2 * n_dims
