"""Stand-in ``tensorflow`` / ``tensorflow_probability`` modules so that the reference's OWN
source files can be executed in this image.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  TensorFlow and TFP are absent and not
installable here, but the reference's hot path is ~250 lines of Python living under
``/root/reference/estimators`` that touch TF through ~20 elementwise ops and TFP through seven
glue classes.  This module provides exactly that surface on top of torch-CPU tensors
(float64 by default) and registers it in ``sys.modules``; ``load_reference()`` then imports
``estimators/normalizing_flows/*.py`` and ``estimators/DistributionLayers.py`` UNMODIFIED from
``/root/reference`` and returns them.  What this pins: every formula, constant, slice offset
and ordering decision the reference itself makes (parameter constraints, the L1 radius, the
tail-first parameter layout, the nested GradientTape derivative, the base distribution, the
mixture layouts).  What it does not pin: TFP's own glue, which is restated below from its
documented semantics -- each class cites the reference call site it serves.

Because the tensors are torch tensors, gradients of the reference's log_prob with respect to
its inputs come from torch autograd THROUGH the reference's code (``tf.GradientTape`` is mapped
onto ``torch.autograd.grad(create_graph=True)``).

Beyond the flow / layer surface the module also stands in for the Keras pieces the estimator classes use
(``Sequential``, ``Dense``, ``Lambda``, ``GaussianNoise``, ``DenseVariational``, a minimal mini-batch Adam
``fit`` / ``evaluate``), which is enough for all 23 tests of the reference's own suite that it does not mark
slow to pass unmodified (``oracle/ref_pytest_plugin.py``).  The golden fixtures never use the training loop
or the stand-in RNG: they set weights and draws explicitly.

``oracle/make_reference_run.py`` uses this to freeze ``tests/golden/reference_run.json``; it can
only run where ``/root/reference`` exists (this container), the fixture travels.
"""
import importlib
import importlib.machinery
import math
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = "/root/reference"
_DTYPE = torch.float64
_GEN = torch.Generator().manual_seed(22)  # stands in for TF's global RNG (tf.random.set_seed reseeds it)


def _randn(shape):
    return torch.randn(tuple(shape), generator=_GEN, dtype=torch.float64).to(_DTYPE)


class _EagerTensor(torch.Tensor):
    """torch tensor that, like a TF EagerTensor, combines with NumPy arrays on either side
    (BaseEstimator.py:85 computes ``ndarray - tensor * ndarray``)."""

    def numpy(self):
        # an EagerTensor's .numpy() never fails because some earlier tape watched an ancestor
        return torch.Tensor.numpy(self.detach().as_subclass(torch.Tensor))

    @property
    def shape(self):
        # tf.TensorShape compares equal to a list (tests/test_ml_estimator.py:28: ``.shape == [10]``)
        return _Shape(self.size())


def _binary(name):
    base = getattr(torch.Tensor, name)

    def op(self, other):
        if isinstance(other, (np.ndarray, np.generic, _Variable, list, tuple)):  # TF converts nested lists too
            other = _t(other)
        return base(self, other)

    op.__name__ = name
    return op


for _name in ("__add__", "__radd__", "__sub__", "__rsub__", "__mul__", "__rmul__", "__truediv__", "__rtruediv__"):
    setattr(_EagerTensor, _name, _binary(_name))


def _t(x):
    """tf.convert_to_tensor for the shim's single working dtype."""
    if isinstance(x, _Variable):
        return x.value()
    if isinstance(x, torch.Tensor):
        x = x if x.dtype == _DTYPE else x.to(_DTYPE)
    else:
        x = torch.as_tensor(np.asarray(x, dtype=np.float64), dtype=_DTYPE)
    return x if isinstance(x, _EagerTensor) else x.as_subclass(_EagerTensor)


def _softplus(x):
    # log(1 + exp(x)) without torch's x > 20 shortcut (exact to the last bit in float64)
    x = _t(x)
    return torch.logaddexp(x, torch.zeros_like(x))


def _reduce_sum(x, axis=None, keepdims=False):
    x = _t(x)
    if axis is None:
        return torch.sum(x)
    return torch.sum(x, dim=axis, keepdim=keepdims)


class _GradientTape:
    """RadialFlow.py:64-67: ``g.watch(r); h = ...; g.gradient(h, r)``.  tape.gradient of a
    non-scalar target is the gradient of its sum (cotangent of ones)."""

    def __init__(self, persistent=False):
        self._mode = None

    def __enter__(self):
        # a tape records whatever the caller's torch grad mode is (the harness scores under no_grad)
        self._mode = torch.enable_grad()
        self._mode.__enter__()
        return self

    def __exit__(self, *exc):
        self._mode.__exit__(*exc)
        return False

    def watch(self, x):
        if not x.requires_grad:
            x.requires_grad_(True)

    def gradient(self, target, source):
        keep = torch.is_grad_enabled()
        with torch.enable_grad():
            g = torch.autograd.grad(target, source, grad_outputs=torch.ones_like(target),
                                    create_graph=True)[0]
        return g if keep else g.detach()


class _Variable:
    """tf.Variable as GaussianKernelsLayer uses it (DistributionLayers.py:101-105, :169-170)."""

    def __init__(self, initial_value=None, dtype=None, trainable=True, **kw):
        self._v = _t(initial_value).clone()
        self.trainable = trainable

    def assign(self, value):
        value = _t(value)
        assert tuple(value.shape) == tuple(self._v.shape), (value.shape, self._v.shape)
        self._v = value.clone()
        return self

    def value(self):
        return self._v

    @property
    def shape(self):
        return self._v.shape

    def numpy(self):
        return _t(self._v).numpy()

    def __mul__(self, other):
        return self._v * _t(other)

    __rmul__ = __mul__

    def __getitem__(self, i):
        return _t(self._v)[i]


# ------------------------------------------------------------------ tfp.bijectors
class _Bijector:
    """tfp.bijectors.Bijector base as PlanarFlow.py:21 / RadialFlow.py:21 construct it.  With
    event_ndims equal to the declared minimum (1) the public methods are the private ones."""

    def __init__(self, validate_args=False, name=None, forward_min_event_ndims=None,
                 inverse_min_event_ndims=None, **kw):
        if forward_min_event_ndims is None:
            forward_min_event_ndims = inverse_min_event_ndims
        if inverse_min_event_ndims is None:
            inverse_min_event_ndims = forward_min_event_ndims
        self.forward_min_event_ndims = forward_min_event_ndims
        self.inverse_min_event_ndims = inverse_min_event_ndims
        self.name = name

    def forward(self, x):
        return self._forward(_t(x))

    def forward_log_det_jacobian(self, x, event_ndims=None):
        assert event_ndims in (None, self.forward_min_event_ndims)
        return self._forward_log_det_jacobian(_t(x))


class _Affine(_Bijector):
    """tfp.bijectors.Affine(shift, scale_diag) (AffineFlow.py:7-9): y = scale_diag * x + shift,
    fldj = sum_i log|scale_diag_i| (LinearOperatorDiag.log_abs_determinant), event_ndims 1."""

    def __init__(self, shift=None, scale_diag=None, name="affine", **kw):
        super().__init__(name=name, forward_min_event_ndims=1)
        self.shift = _t(shift)
        self._scale_diag = _t(scale_diag)

    def _forward(self, x):
        return self._scale_diag * x + self.shift

    def _forward_log_det_jacobian(self, x):
        return torch.sum(torch.log(torch.abs(self._scale_diag)), -1)


class _Chain(_Bijector):
    """tfp.bijectors.Chain (DistributionLayers.py:278): Chain([b0, b1, ...]).forward applies
    the LAST bijector first; each log-det is evaluated at that bijector's own input."""

    def __init__(self, bijectors=None, name="chain", **kw):
        super().__init__(name=name, forward_min_event_ndims=1)
        self.bijectors = list(bijectors or [])

    def _forward(self, x):
        for b in reversed(self.bijectors):
            x = b.forward(x)
        return x

    def _forward_log_det_jacobian(self, x):
        ld = torch.zeros((), dtype=_DTYPE)
        for b in reversed(self.bijectors):
            ld = ld + b.forward_log_det_jacobian(x, event_ndims=1)
            x = b.forward(x)
        return ld


class _Invert(_Bijector):
    """tfp.bijectors.Invert (DistributionLayers.py:250): swaps forward and inverse."""

    def __init__(self, bijector, name=None, **kw):
        super().__init__(name=name, forward_min_event_ndims=bijector.inverse_min_event_ndims,
                         inverse_min_event_ndims=bijector.forward_min_event_ndims)
        self.bijector = bijector

    def inverse(self, y):
        return self.bijector.forward(y)

    def inverse_log_det_jacobian(self, y, event_ndims=None):
        return self.bijector.forward_log_det_jacobian(y, event_ndims=event_ndims)


# ------------------------------------------------------------------ tfp.distributions
class _Shape(tuple):
    """tf.TensorShape: ``TensorShape([d]) == d`` is True (``as_shape(d)``), which
    BaseEstimator.py:81 relies on (``output.event_shape == y.shape[-1]``)."""

    def __eq__(self, other):
        if isinstance(other, int):
            other = (other,)
        return tuple(self) == tuple(other)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = tuple.__hash__


_EPS_QUEUE = []  # standard-normal draws handed to ``sample()`` in order (filled by the harness)


def push_draws(eps_list):
    """Queue the standard-normal draws the next ``sample()`` calls consume (one array per call):
    TF's RNG stream cannot be reproduced, so the harness owns the randomness and records it."""
    _EPS_QUEUE.extend(_t(e) for e in eps_list)


class _Distribution:
    # DistributionLayers.py:36: ``tfd.Distribution.mean if map_mode else tfd.Distribution.sample`` is the
    # layer's convert_to_tensor_fn, i.e. these are called unbound with the distribution as ``self``
    def mean(self):
        return self._mean()

    def sample(self, sample_shape=None):
        if sample_shape is None:
            return self._sample()
        n = int(sample_shape)
        return torch.stack([self._sample() for _ in range(n)], 0)

    def prob(self, x):
        return torch.exp(self.log_prob(x))


class _MultivariateNormalDiag(_Distribution):
    """tfd.MultivariateNormalDiag(loc, scale_diag | scale_identity_multiplier)
    (DistributionLayers.py:125, :200, :283, :292).  log_prob(x) =
    -1/2 sum(((x-loc)/s)^2) - sum(log|s|) - d/2 log(2 pi); an identity multiplier of shape [M]
    is one isotropic scale per batch member (batch_shape [M], event_shape [d])."""

    def __init__(self, loc=None, scale_diag=None, scale_identity_multiplier=None, name=None, **kw):
        self.loc = _t(loc)
        if scale_diag is not None:
            self.scale = _t(scale_diag)
        elif scale_identity_multiplier is not None:
            self.scale = _t(scale_identity_multiplier)[..., None]
        else:
            self.scale = torch.ones((), dtype=_DTYPE)

    @property
    def event_shape(self):
        return _Shape(self.loc.shape[-1:])

    @property
    def batch_shape(self):
        return _Shape(torch.broadcast_shapes(self.loc.shape, self.scale.shape)[:-1])

    def _sample(self):
        shape = tuple(torch.broadcast_shapes(self.loc.shape, self.scale.shape))
        return _t(self.loc + self.scale * _randn(shape))

    def log_prob(self, x):
        e = (_t(x) - self.loc) / self.scale
        d = e.shape[-1]
        log_s = torch.log(torch.abs(self.scale)) + torch.zeros_like(e)
        return -0.5 * torch.sum(e * e, -1) - torch.sum(log_s, -1) - 0.5 * d * math.log(2.0 * math.pi)


class _Categorical(_Distribution):
    def __init__(self, logits=None, probs=None, **kw):
        self.logits = _t(logits) if logits is not None else torch.log(_t(probs))


class _Mixture(_Distribution):
    """tfd.Mixture(cat, components).log_prob (DistributionLayers.py:198-211):
    logsumexp_k(log_softmax(logits)_k + components[k].log_prob(x))."""

    def __init__(self, cat=None, components=None, **kw):
        self.cat, self.components = cat, list(components)
        # TFP validates the static shapes and raises ValueError (tests/test_distribution_layers.py:51,64)
        if self.cat.logits.shape[-1] != len(self.components):
            raise ValueError("cat.num_classes != len(components): %d vs %d"
                             % (self.cat.logits.shape[-1], len(self.components)))
        if len({tuple(c.event_shape) for c in self.components}) != 1:
            raise ValueError("components must all have the same event shape")

    @property
    def batch_shape(self):
        return _Shape(self.cat.logits.shape[:-1])

    def _sample(self):
        """One draw per batch member from the stand-in RNG (the reference's tests use it to make data)."""
        logits = self.cat.logits.detach().as_subclass(torch.Tensor)
        u = torch.rand(tuple(logits.shape[:-1]) + (1,), generator=_GEN, dtype=torch.float64).to(logits.dtype)
        idx = (torch.cumsum(torch.softmax(logits, -1), -1) < u).sum(-1).clamp(max=logits.shape[-1] - 1)
        draws = torch.stack([c._sample().detach().as_subclass(torch.Tensor) for c in self.components], -2)
        idx = idx.reshape(idx.shape + (1, 1)).expand(idx.shape + (1, draws.shape[-1]))
        return _t(torch.gather(draws, -2, idx).squeeze(-2))

    @property
    def event_shape(self):
        return self.components[0].event_shape

    def log_prob(self, x):
        lp = torch.stack([c.log_prob(x) for c in self.components], -1)
        return torch.logsumexp(lp + torch.log_softmax(self.cat.logits, -1), -1)


class _MixtureSameFamily(_Distribution):
    """tfd.MixtureSameFamily.log_prob (DistributionLayers.py:124-131): the event gets a
    component axis, ``x[..., None, :]``, and is scored by the batched component distribution."""

    def __init__(self, mixture_distribution=None, components_distribution=None, **kw):
        self.mixture_distribution = mixture_distribution
        self.components_distribution = components_distribution

    @property
    def batch_shape(self):
        return _Shape(self.mixture_distribution.logits.shape[:-1])

    def _sample(self):
        """One draw per batch member (torch RNG; only its shape is ever looked at,
        tests/test_distribution_layers.py:118,123)."""
        comp = self.components_distribution
        idx = torch.distributions.Categorical(logits=self.mixture_distribution.logits.detach()).sample()
        loc = (comp.loc + torch.zeros_like(comp.loc / comp.scale)).detach()
        scale = (comp.scale + torch.zeros_like(loc)).detach().abs()
        pick = idx[..., None, None].expand(*idx.shape, 1, loc.shape[-1])
        return _t((loc + scale * torch.randn_like(loc)).gather(-2, pick).squeeze(-2))

    @property
    def event_shape(self):
        return self.components_distribution.event_shape

    def log_prob(self, x):
        lp = self.components_distribution.log_prob(_t(x)[..., None, :])
        return torch.logsumexp(lp + torch.log_softmax(self.mixture_distribution.logits, -1), -1)


class _TransformedDistribution(_Distribution):
    """tfd.TransformedDistribution.log_prob (DistributionLayers.py:246):
    base.log_prob(bijector.inverse(y)) + bijector.inverse_log_det_jacobian(y, event_ndims=1)."""

    def __init__(self, distribution=None, bijector=None, **kw):
        self.distribution, self.bijector = distribution, bijector

    @property
    def event_shape(self):
        return self.distribution.event_shape

    @property
    def batch_shape(self):
        return self.distribution.batch_shape

    def log_prob(self, y):
        y = _t(y)
        x = self.bijector.inverse(y)
        ildj = self.bijector.inverse_log_det_jacobian(y, event_ndims=1)
        return self.distribution.log_prob(x) + ildj


class _Normal(_Distribution):
    def __init__(self, loc=None, scale=None, **kw):
        self.loc, self.scale = _t(loc), _t(scale)

    @property
    def batch_shape(self):
        return _Shape(torch.broadcast_shapes(self.loc.shape, self.scale.shape))

    def _mean(self):
        return self.loc + torch.zeros_like(self.scale)

    def _sample(self):
        shape = tuple(torch.broadcast_shapes(self.loc.shape, self.scale.shape))
        if _EPS_QUEUE:  # draws owned (and recorded) by a harness
            eps = _EPS_QUEUE.pop(0)
            assert tuple(eps.shape) == shape
        else:
            eps = _randn(shape)
        return self.loc + self.scale * eps

    def log_prob(self, x):
        e = (_t(x) - self.loc) / self.scale
        return -0.5 * e * e - torch.log(self.scale) - 0.5 * math.log(2.0 * math.pi)


class _Independent(_Distribution):
    def __init__(self, distribution=None, reinterpreted_batch_ndims=1, **kw):
        self.distribution, self.n = distribution, reinterpreted_batch_ndims

    def log_prob(self, x):
        lp = self.distribution.log_prob(x)
        return torch.sum(lp, dim=tuple(range(-self.n, 0)))

    @property
    def batch_shape(self):
        b = self.distribution.batch_shape
        return _Shape(b[: len(b) - self.n])

    @property
    def event_shape(self):
        b = self.distribution.batch_shape
        return _Shape(b[len(b) - self.n:])

    def _mean(self):
        return self.distribution._mean()

    def _sample(self):
        return self.distribution._sample()


def _kl_divergence(q, r):
    """tfd.kl_divergence for Independent(Normal) pairs (the only pair on this path): the sum over
    the reinterpreted dims of  log(s_r / s_q) + (s_q^2 + (m_q - m_r)^2) / (2 s_r^2) - 1/2."""
    assert isinstance(q, _Independent) and isinstance(r, _Independent) and q.n == r.n
    a, b = q.distribution, r.distribution
    kl = torch.log(b.scale / a.scale) + (a.scale ** 2 + (a.loc - b.loc) ** 2) / (2.0 * b.scale ** 2) - 0.5
    return torch.sum(kl, dim=tuple(range(-q.n, 0)))


# ------------------------------------------------------------------ tfp.layers / tf.keras
class _DistributionLambda:
    """tfp.layers.DistributionLambda: calling the layer builds the distribution from ``t``."""

    def __init__(self, make_distribution_fn=None, convert_to_tensor_fn=None, dtype=None, **kw):
        self._make_distribution_fn = make_distribution_fn
        self._convert_to_tensor_fn = convert_to_tensor_fn

    def __call__(self, t):
        d = self._make_distribution_fn(_t(t))
        d._convert_to_tensor_fn = self._convert_to_tensor_fn  # what tf.convert_to_tensor(d) would apply
        return d


class _DenseVariational:
    """tfp.layers.DenseVariational (BayesianNNEstimator.py:128-148), restated from TFP's documented
    behaviour: q = posterior(x), r = prior(x); the layer adds ``kl_weight * KL`` as a scalar loss (exact
    KL, or the one-sample estimate log q(w) - log r(w)); w = convert_to_tensor(q) is split into
    kernel = reshape(w[:in * units], [in, units]) and bias = w[in * units:]; out = act(x @ kernel + bias)."""

    def __init__(self, units, make_posterior_fn=None, make_prior_fn=None, kl_weight=None, kl_use_exact=False,
                 activation=None, use_bias=True, **kw):
        assert use_bias
        self.units, self.activation = units, activation
        self._make_posterior_fn, self._make_prior_fn = make_posterior_fn, make_prior_fn
        self.kl_weight, self.kl_use_exact = kl_weight, kl_use_exact
        self._posterior = self._prior = None
        self.losses = []

    def build(self, in_features):
        self.in_features = in_features
        self._posterior = self._make_posterior_fn(in_features * self.units, self.units, None)
        self._prior = self._make_prior_fn(in_features * self.units, self.units, None)

    def __call__(self, x):
        x = _t(x)
        if self._posterior is None:
            self.build(x.shape[-1])
        q, r = self._posterior(x), self._prior(x)
        w = q._convert_to_tensor_fn(q)
        if self.kl_use_exact:
            kl = _kl_divergence(q, r)
        else:
            kl = q.log_prob(w) - r.log_prob(w)
        if self.kl_weight is not None:
            kl = self.kl_weight * kl
        self.losses = [torch.sum(kl)]
        nk = self.in_features * self.units
        kernel, bias = w[..., :nk].reshape(self.in_features, self.units), w[..., nk:]
        return _Dense._ACT[self.activation](x @ kernel + bias)


class _VariableLayer:
    """tfp.layers.VariableLayer (DistributionLayers.py:80-85): ignores its input."""

    def __init__(self, shape=None, dtype=None, initializer="zeros", trainable=True, **kw):
        # "normal" (BayesianNNEstimator.py:102) is Keras RandomNormal(stddev=0.05); the fixture harness
        # overwrites ``variable`` with its own seeded values
        assert initializer in ("zeros", "normal")
        n = shape if isinstance(shape, int) else int(np.prod(shape))
        self.variable = torch.zeros(n, dtype=_DTYPE) if initializer == "zeros" else 0.05 * _randn((n,))
        self.trainable = trainable

    def __call__(self, _x):
        return self.variable


class _Lambda:
    def __init__(self, function, **kw):
        self.function = function

    def __call__(self, x):
        return self.function(x)


class _GaussianNoise:
    """tf.keras.layers.GaussianNoise (BaseEstimator.py:68, MaximumLikelihoodNNEstimator.py:41):
    identity unless training."""

    def __init__(self, stddev, **kw):
        self.stddev = stddev

    def __call__(self, x, training=False):
        std = float(_t(self.stddev))
        if training and std != 0.0:  # from the stand-in RNG: only "is there noise" is ever asserted
            x = _t(x)
            return x + std * _randn(x.shape)
        return x


class _Dense:
    """tf.keras.layers.Dense(units, activation) (MaximumLikelihoodNNEstimator.py:42-43):
    ``activation(x @ kernel + bias)``, kernel [in, units].  The fixture harness sets the weights
    (``set_weights``); otherwise they are built on first call like Keras does."""

    _ACT = {"linear": lambda v: v, None: lambda v: v, "tanh": torch.tanh, "relu": torch.relu,
            "sigmoid": torch.sigmoid}

    def __init__(self, units, activation=None, **kw):
        self.units, self.activation = units, activation
        self.kernel = self.bias = None

    def set_weights(self, weights):
        kernel, bias = weights
        self.kernel, self.bias = _t(kernel), _t(bias)
        assert self.kernel.shape[1] == self.units and tuple(self.bias.shape) == (self.units,)

    def get_weights(self):
        return [self.kernel, self.bias]

    def __call__(self, x):
        x = _t(x)
        if self.kernel is None:  # Keras defaults: glorot-uniform kernel, zero bias
            fan_in = x.shape[-1]
            limit = math.sqrt(6.0 / (fan_in + self.units))
            self.kernel = ((torch.rand(fan_in, self.units, generator=_GEN, dtype=torch.float64) * 2 - 1) * limit).to(_DTYPE)
            self.bias = torch.zeros(self.units, dtype=_DTYPE)
        return self._ACT[self.activation](x @ self.kernel + self.bias)


class _Sequential:
    """tf.keras.Sequential as BaseEstimator subclasses it (BaseEstimator.py:8, :18, :61-69):
    layer list, ``add``, ``call(x, training)``, ``__call__`` = inference call; ``compile`` only
    records its arguments and ``fit`` (the Keras training loop) is not emulated."""

    def __init__(self, layers=None, **kw):
        self.layers = list(layers or [])

    def add(self, layer):
        self.layers.append(layer)

    def call(self, x, training=False):
        x = _t(x)
        for layer in self.layers:
            x = layer(x, training=training) if isinstance(layer, _GaussianNoise) else layer(x)
        return x

    def __call__(self, x, training=False):
        return self.call(x, training=training)

    @property
    def losses(self):
        """Keras ``model.losses``: the scalars layers added during the last call (added to the compiled loss)."""
        return [l for layer in self.layers for l in getattr(layer, "losses", [])]

    def compile(self, optimizer=None, loss=None, **kw):
        self.optimizer, self.loss = optimizer, loss

    # -- a minimal Keras training loop, enough for the reference's own (non-slow) estimator tests:
    #    mini-batches of 32, shuffled; loss = mean(loss_fn(y, model(x, training=True))) + sum(model.losses);
    #    Adam with Keras' epsilon.  Not used for any fixture (those set weights explicitly).
    def _total_loss(self, x, y, training):
        per_sample = self.loss(_t(y), self.call(x, training=training))
        return per_sample.mean() + sum(self.losses, torch.zeros((), dtype=_DTYPE))

    def fit(self, x=None, y=None, batch_size=None, epochs=None, verbose=0, callbacks=None, shuffle=True, **kw):
        x, y = _t(x).as_subclass(torch.Tensor), _t(y).as_subclass(torch.Tensor)
        n, bs = x.shape[0], int(batch_size or 32)
        with torch.no_grad():
            self.call(x[:2], training=False)  # builds the lazily created weights
        slots = _trainable_slots(self, set())
        params = []
        for holder, attr in slots:
            v = getattr(holder, attr).detach().as_subclass(torch.Tensor).clone().requires_grad_(True)
            setattr(holder, attr, v)
            params.append(v)
        opt = torch.optim.Adam(params, lr=float(getattr(self.optimizer, "learning_rate", 1e-3)), eps=1e-7)
        self.history = []
        for _ in range(int(epochs or 1)):
            order = torch.randperm(n, generator=_GEN) if shuffle else torch.arange(n)
            total = 0.0
            for lo in range(0, n, bs):
                idx = order[lo:lo + bs]
                opt.zero_grad(set_to_none=True)
                loss = self._total_loss(x[idx], y[idx], training=True)
                loss.backward()
                opt.step()
                total += float(loss.detach()) * len(idx)
            self.history.append(total / n)
        for (holder, attr), v in zip(slots, params):
            setattr(holder, attr, v.detach())
        return self

    def evaluate(self, x=None, y=None, **kw):
        with torch.no_grad():
            return float(self._total_loss(x, y, training=False))


class _Adam:
    def __init__(self, learning_rate=0.001, **kw):
        self.learning_rate = learning_rate


def _trainable_slots(obj, seen):
    """(holder, attribute) pairs of the trainable tensors reachable from a layer / model: Dense kernels and
    biases, trainable VariableLayers (DenseVariational posteriors / priors, the KMN scale model)."""
    if id(obj) in seen:
        return []
    seen.add(id(obj))
    out = []
    if isinstance(obj, _Dense):
        out += [(obj, "kernel"), (obj, "bias")]
    elif isinstance(obj, _VariableLayer):
        if obj.trainable:
            out.append((obj, "variable"))
    if isinstance(obj, (list, tuple)):
        for o in obj:
            out += _trainable_slots(o, seen)
    elif isinstance(obj, (_Sequential, _DenseVariational, _DistributionLambda, _Dense, _Lambda)):
        for v in vars(obj).values():
            if isinstance(v, (list, tuple, _Sequential, _DenseVariational, _DistributionLambda, _Dense, _VariableLayer)):
                out += _trainable_slots(v, seen)
    return out


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__shim__ = True
    # a spec, so that ``importlib.util.find_spec("tensorflow")`` probes (torch._dynamo does one) do not trip
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    return m


def install(dtype=torch.float64):
    """Register the stand-in modules.  Refuses to shadow a real TensorFlow."""
    global _DTYPE
    for name in ("tensorflow", "tensorflow_probability"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__shim__", False):
            raise RuntimeError(f"a real {name} is loaded; run the real reference instead")
    _DTYPE = dtype

    tf_math = _module(
        "tensorflow.math",
        reduce_sum=_reduce_sum,
        softplus=_softplus,
        tanh=lambda x: torch.tanh(_t(x)),
        sin=lambda x: torch.sin(_t(x)),
        log=lambda x: torch.log(_t(x)),
        abs=lambda x: torch.abs(_t(x)),
        expm1=lambda x: torch.expm1(_t(x)),
        reduce_logsumexp=lambda x, axis=None: torch.logsumexp(_t(x), dim=axis),
    )
    tf_nn = _module("tensorflow.nn", softplus=_softplus)
    keras = _module(
        "tensorflow.keras",
        Sequential=_Sequential,
        models=_module("tensorflow.keras.models", Sequential=_Sequential),
        layers=_module("tensorflow.keras.layers", Lambda=_Lambda, GaussianNoise=_GaussianNoise, Dense=_Dense),
        optimizers=_module("tensorflow.keras.optimizers", Adam=_Adam),
        callbacks=_module("tensorflow.keras.callbacks", TerminateOnNaN=type("TerminateOnNaN", (), {})),
        backend=_module("tensorflow.keras.backend", clear_session=lambda: None),
    )
    tf2 = _module("tensorflow.python.tf2", enabled=lambda: True)
    tf_python = _module("tensorflow.python", tf2=tf2)
    tf = _module(
        "tensorflow",
        math=tf_math,
        nn=tf_nn,
        keras=keras,
        python=tf_python,
        float32="float32",
        float64="float64",
        abs=tf_math.abs,
        reduce_sum=_reduce_sum,
        reduce_prod=lambda x, axis=None: torch.prod(_t(x)) if axis is None else torch.prod(_t(x), dim=axis),
        random=_module("tensorflow.random", set_seed=lambda seed: _GEN.manual_seed(int(seed))),
        squeeze=lambda x, axis=None: torch.squeeze(_t(x)) if axis is None else torch.squeeze(_t(x), dim=axis),
        expand_dims=lambda x, axis: torch.unsqueeze(_t(x), axis),
        zeros_like=lambda x: torch.zeros_like(_t(x)),
        ones_like=lambda x: torch.ones_like(_t(x)),
        zeros=lambda shape, dtype=None: _t(torch.zeros(shape, dtype=_DTYPE)),
        ones=lambda shape, dtype=None: _t(torch.ones(shape, dtype=_DTYPE)),
        constant=lambda v, dtype=None: _t(v),
        concat=lambda values, axis=0: torch.cat([_t(v) for v in values], dim=axis),
        convert_to_tensor=lambda v, dtype=None: _t(v),
        GradientTape=_GradientTape,
        Variable=_Variable,
    )
    tf.__path__ = []  # a package, so that ``from tensorflow.python import tf2`` resolves
    tf_python.__path__ = []

    bijectors = _module("tensorflow_probability.bijectors", Bijector=_Bijector, Affine=_Affine,
                        Chain=_Chain, Invert=_Invert)
    distributions = _module(
        "tensorflow_probability.distributions",
        Distribution=_Distribution,
        MultivariateNormalDiag=_MultivariateNormalDiag,
        Categorical=_Categorical,
        Mixture=_Mixture,
        MixtureSameFamily=_MixtureSameFamily,
        TransformedDistribution=_TransformedDistribution,
        Normal=_Normal,
        Independent=_Independent,
    )
    layers = _module("tensorflow_probability.layers", DistributionLambda=_DistributionLambda,
                     VariableLayer=_VariableLayer, DenseVariational=_DenseVariational)
    tfp = _module("tensorflow_probability", bijectors=bijectors, distributions=distributions,
                  layers=layers)
    tfp.__path__ = []

    for m in (tf, tf_math, tf_nn, keras, keras.models, keras.layers, keras.optimizers, keras.callbacks,
              keras.backend, tf.random, tf_python, tf2, tfp, bijectors, distributions, layers):
        sys.modules[m.__name__] = m
    return tf, tfp


def uninstall():
    """Drop the stand-in modules, the reference modules loaded through them and any queued draws."""
    del _EPS_QUEUE[:]
    for name in list(sys.modules):
        top = name.split(".")[0]
        if top in ("tensorflow", "tensorflow_probability", "estimators", "evaluation") and \
                getattr(sys.modules.get(top), "__shim__", False):
            if name != top:
                del sys.modules[name]
    for top in ("tensorflow", "tensorflow_probability", "estimators", "evaluation"):
        if getattr(sys.modules.get(top), "__shim__", False):
            del sys.modules[top]


def load_reference(root=REFERENCE_ROOT, dtype=torch.float64):
    """Import the reference's flow and distribution-layer modules, unmodified, from ``root``.

    ``estimators/__init__.py`` is NOT executed (it pulls in the Keras estimator classes, which
    are outside the path): an empty package object with the reference's ``__path__`` stands in
    for it, so ``estimators.normalizing_flows`` and ``estimators.DistributionLayers`` are the
    reference's own files.  Returns ``(FLOWS, DistributionLayers module)``.
    """
    import os

    if not os.path.isdir(os.path.join(root, "estimators")):
        raise FileNotFoundError(f"{root}/estimators not found (the reference does not travel)")
    install(dtype)
    pkg = types.ModuleType("estimators")
    pkg.__path__ = [os.path.join(root, "estimators")]
    pkg.__shim__ = True
    pkg.__spec__ = importlib.machinery.ModuleSpec("estimators", None, is_package=True)
    pkg.__spec__.submodule_search_locations = pkg.__path__
    for name in [n for n in sys.modules if n == "estimators" or n.startswith("estimators.")]:
        del sys.modules[name]
    sys.modules["estimators"] = pkg
    keep, sys.dont_write_bytecode = sys.dont_write_bytecode, True  # never write into the reference tree
    try:
        flows = importlib.import_module("estimators.normalizing_flows")
        layers = importlib.import_module("estimators.DistributionLayers")
    finally:
        sys.dont_write_bytecode = keep
    for mod in (flows, layers):
        assert mod.__file__.startswith(root), mod.__file__
    return flows.FLOWS, layers


def load_reference_module(name, root=REFERENCE_ROOT):
    """Import one more of the reference's modules (e.g. ``estimators.NormalizingFlowNetwork``)
    after ``load_reference()``; same guarantees (unmodified file under ``root``, no bytecode)."""
    assert getattr(sys.modules.get("estimators"), "__shim__", False), "call load_reference() first"
    keep, sys.dont_write_bytecode = sys.dont_write_bytecode, True
    try:
        mod = importlib.import_module(name)
    finally:
        sys.dont_write_bytecode = keep
    assert mod.__file__.startswith(root), mod.__file__
    return mod


def load_reference_package(root=REFERENCE_ROOT, dtype=torch.float64):
    """``load_reference()`` plus the body of the reference's ``estimators/__init__.py`` executed in the package
    object (so ``from estimators import NormalizingFlowNetwork, ...`` works as in the reference's tests) and an
    ``evaluation`` package object for ``evaluation.scorers``."""
    import os

    flows, layers = load_reference(root, dtype)
    pkg = sys.modules["estimators"]
    init = os.path.join(root, "estimators", "__init__.py")
    keep, sys.dont_write_bytecode = sys.dont_write_bytecode, True
    try:
        with open(init) as f:
            exec(compile(f.read(), init, "exec"), pkg.__dict__)
        ev = types.ModuleType("evaluation")
        ev.__path__ = [os.path.join(root, "evaluation")]
        ev.__shim__ = True
        ev.__spec__ = importlib.machinery.ModuleSpec("evaluation", None, is_package=True)
        ev.__spec__.submodule_search_locations = ev.__path__
        sys.modules["evaluation"] = ev
        importlib.import_module("evaluation.scorers")
    finally:
        sys.dont_write_bytecode = keep
    return pkg
