"""normalizingflownetwork_b200 -- B200 (sm_100a) flow-chain / mixture-head log-likelihood.

Drop-in for the hot path of siboehm/NormalizingFlowNetwork behind the reference's own
names: ``normalizing_flows.FLOWS`` bijectors, ``DistributionLayers`` (InverseNormalizingFlowLayer,
GaussianMixtureLayer, GaussianKernelsLayer, MeanFieldLayer) and the estimators built on
them.  All density arithmetic runs in hand-written CUDA kernels inside
``libnfn_b200.so`` (C ABI: ``include/nfn_b200.h``); PyTorch carries the tensors.
There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from . import functional  # noqa: F401
from .normalizing_flows import FLOWS, AffineFlow, PlanarFlow, RadialFlow  # noqa: F401
from .DistributionLayers import (  # noqa: F401
    GaussianKernelsLayer,
    GaussianMixtureLayer,
    InverseNormalizingFlowLayer,
    MeanFieldLayer,
)

__version__ = "0.1.0"
