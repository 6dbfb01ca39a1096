from .models import (  # noqa: F401
    BayesKernelMixtureNetwork,
    BayesMixtureDensityNetwork,
    BayesNormalizingFlowNetwork,
    KernelMixtureNetwork,
    MixtureDensityNetwork,
    NormalizingFlowNetwork,
)

# same keys as the reference's estimators/__init__.py:8-15
ESTIMATORS = {
    "bayesian_NFN": BayesNormalizingFlowNetwork,
    "bayesian_KMN": BayesKernelMixtureNetwork,
    "bayesian_MDN": BayesMixtureDensityNetwork,
    "NFN": NormalizingFlowNetwork,
    "KMN": KernelMixtureNetwork,
    "MDN": MixtureDensityNetwork,
}
