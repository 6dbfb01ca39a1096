"""The estimator facade (reference API: ``fit / log_pdf / pdf / score``) on top of the fused CUDA heads.

The ``MaximumLikelihoodNNEstimator`` and ``BayesianNNEstimator`` submodules are the two training regimes (point weights with
Adam on the mean NLL; mean-field weight posteriors with the KL term, S weight draws folded into the batch);
``models`` holds the six concrete estimators, one per (regime, head) pair: flow chain, Gaussian mixture,
kernel mixture.  ``ESTIMATORS`` uses the registry names of the reference's ``estimators/__init__.py:8-15``.
"""
from .models import (
    BayesKernelMixtureNetwork,
    BayesMixtureDensityNetwork,
    BayesNormalizingFlowNetwork,
    KernelMixtureNetwork,
    MixtureDensityNetwork,
    NormalizingFlowNetwork,
)

ESTIMATORS = {
    "NFN": NormalizingFlowNetwork,
    "MDN": MixtureDensityNetwork,
    "KMN": KernelMixtureNetwork,
    "bayesian_NFN": BayesNormalizingFlowNetwork,
    "bayesian_MDN": BayesMixtureDensityNetwork,
    "bayesian_KMN": BayesKernelMixtureNetwork,
}
