"""Shared body of the three bijector facades (see PlanarFlow.py / RadialFlow.py / AffineFlow.py)."""
import torch

from .. import functional as F


class _FlowBase:
    """Per-sample-parameterised bijector whose arithmetic runs in libnfn_b200.so.

    Mirrors the tfp.bijectors.Bijector surface the reference's callers and tests use
    (/root/reference/tests/test_flows.py:19-41): ``forward(z)``,
    ``_forward_log_det_jacobian(z)``, ``forward_min_event_ndims``.
    """

    flow_type = None
    forward_min_event_ndims = 1
    inverse_min_event_ndims = 1

    def __init__(self, t, n_dims, name=None):
        if not torch.is_tensor(t):
            raise TypeError("t must be a CUDA torch.Tensor (the parameter rows emitted by the network)")
        assert t.shape[-1] == self.get_param_size(n_dims)
        self._t = t
        self.n_dims = n_dims
        self.name = name or type(self).__name__

    def _z(self, z):
        if not torch.is_tensor(z):
            z = torch.as_tensor(z, dtype=torch.float32)
        return z.to(device=self._t.device, dtype=torch.float32)

    def forward(self, z):
        return F.flow_forward(self.flow_type, self._t, self._z(z), self.n_dims, want_fldj=False)[0]

    def _forward(self, z):
        return self.forward(z)

    def _forward_log_det_jacobian(self, z):
        return F.flow_forward(self.flow_type, self._t, self._z(z), self.n_dims, want_z=False)[1]

    def forward_log_det_jacobian(self, z, event_ndims=1):
        assert event_ndims == 1
        return self._forward_log_det_jacobian(z)
