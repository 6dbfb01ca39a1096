"""Registry of the conditional bijectors of the flow chain.

Each class wraps one parameter slice ``t[..., begin:begin + size]`` of the network output and evaluates
``forward(z)`` / ``_forward_log_det_jacobian(z)`` through ``nfn_flow_forward`` (single-bijector kernel,
csrc/nfn_generic.cu); inside a chain the same arithmetic runs fused in ``chain_kernel``.  ``flow_type`` on
the class is the name used in ``flow_types`` lists and in ``nfn_chain_desc.flow_type`` (0 planar, 1 radial,
2 affine).  The registry keeps the key set of the reference's ``estimators/normalizing_flows/__init__.py:5``.
"""
from .AffineFlow import AffineFlow
from .PlanarFlow import PlanarFlow
from .RadialFlow import RadialFlow

FLOWS = {"planar": PlanarFlow, "radial": RadialFlow, "affine": AffineFlow}
