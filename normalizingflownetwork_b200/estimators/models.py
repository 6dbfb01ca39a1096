"""The six concrete estimators and their sklearn-grid factories (reference
estimators/{NormalizingFlowNetwork,MixtureDensityNetwork,KernelMixtureNetwork,
BayesNormalizingFlowNetwork,BayesMixtureDensityNetwork,BayesKernelMixtureNetwork}.py)."""
import numpy as np

from ..DistributionLayers import GaussianKernelsLayer, GaussianMixtureLayer, InverseNormalizingFlowLayer
from .BayesianNNEstimator import BayesianNNEstimator
from .MaximumLikelihoodNNEstimator import MaximumLikelihoodNNEstimator


def _radial_chain(n_dims, n_flows, trainable_base_dist):
    # the NFN estimators only ever build radial chains (NormalizingFlowNetwork.py:16-18)
    return InverseNormalizingFlowLayer(flow_types=["radial"] * n_flows, n_dims=n_dims,
                                       trainable_base_dist=trainable_base_dist)


class _KmnFitMixin:
    def fit(self, x, y, batch_size=None, epochs=None, verbose=1, **kwargs):
        y = np.asarray(y)
        y_mean = np.mean(y, axis=0, dtype=np.float32)
        y_std = np.std(y, axis=0, dtype=np.float32)
        self.dist_layer.set_center_points((y - y_mean) / y_std)
        return super().fit(x=x, y=y, batch_size=batch_size, epochs=epochs, verbose=verbose, **kwargs)


class NormalizingFlowNetwork(MaximumLikelihoodNNEstimator):
    def __init__(self, n_dims, n_flows=10, trainable_base_dist=True, **kwargs):
        super().__init__(_radial_chain(n_dims, n_flows, trainable_base_dist), **kwargs)

    @staticmethod
    def build_function(n_dims=1, n_flows=3, hidden_sizes=(16, 16), trainable_base_dist=True,
                       noise_reg=("fixed_rate", 0.0), learning_rate=3e-3, activation="tanh"):
        return NormalizingFlowNetwork(n_dims=n_dims, n_flows=n_flows, hidden_sizes=hidden_sizes,
                                      trainable_base_dist=trainable_base_dist, noise_reg=noise_reg,
                                      learning_rate=learning_rate, activation=activation)


class MixtureDensityNetwork(MaximumLikelihoodNNEstimator):
    def __init__(self, n_dims, n_centers, **kwargs):
        super().__init__(GaussianMixtureLayer(n_centers=n_centers, n_dims=n_dims), **kwargs)

    @staticmethod
    def build_function(n_dims=1, n_centers=5, hidden_sizes=(16, 16), noise_reg=("fixed_rate", 0.0),
                       learning_rate=2e-3, activation="relu"):
        return MixtureDensityNetwork(n_dims=n_dims, n_centers=n_centers, hidden_sizes=hidden_sizes,
                                     noise_reg=noise_reg, learning_rate=learning_rate, activation=activation)


class KernelMixtureNetwork(_KmnFitMixin, MaximumLikelihoodNNEstimator):
    def __init__(self, n_dims, n_centers=50, **kwargs):
        super().__init__(GaussianKernelsLayer(n_centers=n_centers, n_dims=n_dims, trainable_scale=True), **kwargs)

    @staticmethod
    def build_function(n_dims=1, n_centers=30, hidden_sizes=(16, 16), noise_reg=("fixed_rate", 0.0),
                       learning_rate=2e-3, activation="relu"):
        return KernelMixtureNetwork(n_dims=n_dims, n_centers=n_centers, hidden_sizes=hidden_sizes,
                                    noise_reg=noise_reg, learning_rate=learning_rate, activation=activation)


_BAYES_DEFAULTS = dict(kl_use_exact=True, hidden_sizes=(10,), activation="tanh", noise_reg=("fixed_rate", 0.0),
                       learning_rate=2e-2, trainable_prior=False, map_mode=False, prior_scale=1.0)


class BayesNormalizingFlowNetwork(BayesianNNEstimator):
    def __init__(self, n_dims, kl_weight_scale, n_flows=2, trainable_base_dist=True, **kwargs):
        super().__init__(_radial_chain(n_dims, n_flows, trainable_base_dist), kl_weight_scale, **kwargs)

    @staticmethod
    def build_function(n_dims, kl_weight_scale, n_flows=2, trainable_base_dist=True, **kwargs):
        kw = dict(_BAYES_DEFAULTS)
        kw.update(kwargs)
        return BayesNormalizingFlowNetwork(n_dims=n_dims, kl_weight_scale=kl_weight_scale, n_flows=n_flows,
                                           trainable_base_dist=trainable_base_dist, **kw)


class BayesMixtureDensityNetwork(BayesianNNEstimator):
    def __init__(self, n_dims, kl_weight_scale, n_centers=5, **kwargs):
        super().__init__(GaussianMixtureLayer(n_centers=n_centers, n_dims=n_dims), kl_weight_scale, **kwargs)

    @staticmethod
    def build_function(n_dims, kl_weight_scale, n_centers=5, **kwargs):
        kw = dict(_BAYES_DEFAULTS)
        kw.update(kwargs)
        return BayesMixtureDensityNetwork(n_dims=n_dims, kl_weight_scale=kl_weight_scale, n_centers=n_centers, **kw)


class BayesKernelMixtureNetwork(_KmnFitMixin, BayesianNNEstimator):
    def __init__(self, n_dims, kl_weight_scale, n_centers=50, **kwargs):
        super().__init__(GaussianKernelsLayer(n_centers=n_centers, n_dims=n_dims, trainable_scale=True),
                         kl_weight_scale, **kwargs)

    @staticmethod
    def build_function(n_dims, kl_weight_scale, n_centers=50, **kwargs):
        kw = dict(_BAYES_DEFAULTS)
        kw.update(kwargs)
        return BayesKernelMixtureNetwork(n_dims=n_dims, kl_weight_scale=kl_weight_scale, n_centers=n_centers, **kw)
