"""Maximum-likelihood estimator: MLP -> distribution layer, Adam on the mean NLL
(reference estimators/MaximumLikelihoodNNEstimator.py)."""
import torch

from .. import functional as F
from .BaseEstimator import BaseEstimator, _GaussianNoise, _Normalise

ACTIVATIONS = {"relu": torch.nn.ReLU, "tanh": torch.nn.Tanh, "linear": torch.nn.Identity,
               "sigmoid": torch.nn.Sigmoid, "elu": torch.nn.ELU}


class _Dense(torch.nn.Module):
    """Keras Dense: glorot-uniform kernel, zero bias, activation."""

    def __init__(self, units, activation, seed=0):
        super().__init__()
        self.linear = torch.nn.LazyLinear(units)
        self.act = ACTIVATIONS[activation]()
        self._activation = activation
        self._seed = seed
        self.fused = True   # one kernel each way for the layer (csrc/nfn_mlp.cu) when its shape allows
        self.xnorm = None   # (x_mean, x_std) handed over by _Normalise for THIS call: normalise on load

    def forward(self, x):
        # Glorot-initialise only a weight that does not exist yet: after load_state_dict() the lazy layer has
        # been materialised WITH the checkpoint's values, which must survive the first call
        if isinstance(self.linear.weight, torch.nn.parameter.UninitializedParameter):
            self.linear(x)  # materialises the lazy weight
            with torch.no_grad():
                # own generator: the initial weights depend on (random_seed, layer), not on
                # whatever consumed the global RNG between construction and the first call
                gen = torch.Generator(device="cpu").manual_seed(self._seed)
                w = torch.empty(self.linear.weight.shape)
                torch.nn.init.xavier_uniform_(w, generator=gen)
                self.linear.weight.copy_(w)
                self.linear.bias.zero_()
        lin = self.linear
        xnorm, self.xnorm = self.xnorm, None
        if (self.fused and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2
                and F.dense_act_supported(lin.in_features, lin.out_features, self._activation)):
            if xnorm is not None:
                return F.dense_act(x.contiguous(), lin.weight, lin.bias, self._activation, xnorm[0], xnorm[1])
            return F.dense_act(x.contiguous(), lin.weight, lin.bias, self._activation)
        if xnorm is not None:   # (cannot happen: can_fuse_xnorm said yes) -- normalise here rather than skip it
            x = (x - xnorm[0]) / (xnorm[1] + 1e-8)
        return self.act(lin(x))

    def can_fuse_xnorm(self, x):
        """True when this layer's kernels can read raw x and normalise it on load (first layer, both ways)."""
        lin = self.linear
        if isinstance(lin.weight, torch.nn.parameter.UninitializedParameter):
            return False
        return (self.fused and x.dtype == torch.float32
                and F.dense_act_xnorm_supported(lin.in_features, lin.out_features, self._activation))


class MaximumLikelihoodNNEstimator(BaseEstimator):
    def __init__(self, dist_layer, hidden_sizes=(16, 16), noise_reg=("fixed_rate", 0.0), learning_rate=3e-3,
                 activation="relu", random_seed=22, device=None):
        assert len(noise_reg) == 2
        torch.manual_seed(random_seed)
        torch.nn.Module.__init__(self)  # so helper modules can hold a reference to self
        self.random_seed = random_seed
        layers = self._get_dense_layers(hidden_sizes=hidden_sizes, output_size=dist_layer.get_total_param_size(),
                                        activation=activation)
        super().__init__(layers, dist_layer, noise_fn_type=noise_reg[0], noise_scale_factor=noise_reg[1],
                         random_seed=random_seed, device=device)
        self.learning_rate = learning_rate

    def _get_dense_layers(self, hidden_sizes, output_size, activation):
        assert type(hidden_sizes) == tuple or type(hidden_sizes) == list
        normalization = [_Normalise(self)]
        noise_reg = [_GaussianNoise(self, "x_noise_std")]
        seed = 1000 * getattr(self, "random_seed", 22)
        hidden = [_Dense(size, activation, seed + i) for i, size in enumerate(hidden_sizes)]
        output = [_Dense(output_size, "linear", seed + len(hidden_sizes))]
        if hidden:   # the first hidden layer can take the input normalisation as a prologue of its own kernel
            normalization[0].fused_into = hidden[0]
        return normalization + noise_reg + hidden + output

    def _ensure_optimizer(self):
        if self.optimizer is None:
            # Keras Adam defaults (epsilon 1e-7)
            self.optimizer = self._make_adam()

    def fit(self, x, y, batch_size=None, epochs=None, verbose=1, **kwargs):
        import numpy as np

        self._assign_data_normalization(np.asarray(x), np.asarray(y))
        with torch.no_grad():  # materialise lazy layers before the optimiser sees the parameters
            self.params_from_x(np.asarray(x)[:2])
        self._ensure_optimizer()
        return super().fit(x, y, batch_size=batch_size, epochs=epochs, verbose=verbose, **kwargs)
