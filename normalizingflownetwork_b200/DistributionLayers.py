"""Distribution layers: drop-in for the reference's estimators/DistributionLayers.py.

Same class names, constructor arguments, static helpers, assertion behaviour and
parameter-row layouts as the reference (paths below are into /root/reference):
  InverseNormalizingFlowLayer  estimators/DistributionLayers.py:215-294
  GaussianMixtureLayer         estimators/DistributionLayers.py:174-212
  GaussianKernelsLayer         estimators/DistributionLayers.py:74-171
  MeanFieldLayer               estimators/DistributionLayers.py:17-71
Calling a layer on the network output ``t`` returns a light distribution object whose
``log_prob(y)`` / ``prob(y)`` run the fused sm_100a kernels of libnfn_b200.so (there is no
CPU path for them); ``event_shape`` / ``batch_shape`` behave like TFP's for the checks the
reference's tests make.
"""
import math

import numpy as np
import torch

from . import functional as F
from .normalizing_flows import FLOWS


class _Shape(list):
    """Compares equal to lists/tuples/TensorShape-like and, like tf.TensorShape, to a bare
    int for rank-1 shapes (the reference relies on that in BaseEstimator.py:83)."""

    def __eq__(self, other):
        if isinstance(other, int):
            return len(self) == 1 and self[0] == other
        return list(self) == list(other)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None


def _to_tensor_like(y, ref):
    if not torch.is_tensor(y):
        y = torch.as_tensor(np.asarray(y, dtype=np.float32))
    y = y.to(device=ref.device, dtype=torch.float32)
    if y.dim() == 1:
        y = y.unsqueeze(0)  # TFP broadcasts a bare event [d] against the batch
    return y


class _Chain:
    """What tfp.bijectors.Chain exposes to the reference's tests
    (tests/test_distribution_layers.py:185-200)."""

    inverse_min_event_ndims = 1
    forward_min_event_ndims = 1

    def __init__(self, bijectors):
        self.bijectors = list(bijectors)

    def forward(self, z):
        for bij in reversed(self.bijectors):
            z = bij.forward(z)
        return z

    def forward_log_det_jacobian(self, z, event_ndims=1):
        total = 0.0
        for bij in reversed(self.bijectors):
            total = total + bij._forward_log_det_jacobian(z)
            z = bij.forward(z)
        return total


class FlowChainDistribution:
    """TransformedDistribution(base, Invert(Chain(flows))) parameterised by ``t[B, P]``."""

    def __init__(self, t, flow_types, n_dims, trainable_base_dist):
        self.t = t
        self.flow_types = tuple(flow_types)
        self.n_dims = n_dims
        self.trainable_base_dist = trainable_base_dist

    @property
    def event_shape(self):
        return _Shape([self.n_dims])

    @property
    def batch_shape(self):
        return _Shape(self.t.shape[:-1])

    @property
    def bijector(self):
        t_flows = self.t[..., 2 * self.n_dims:] if self.trainable_base_dist else self.t
        return InverseNormalizingFlowLayer._get_bijector(t_flows, self.flow_types, self.n_dims)

    def log_prob(self, y):
        return F.chain_log_prob(self.t, _to_tensor_like(y, self.t), self.flow_types, self.n_dims,
                                self.trainable_base_dist)

    def prob(self, y):
        return torch.exp(self.log_prob(y))

    def log_prob_x(self, y, xform):
        """Forward-only log_prob with the estimator's y pipeline (normalisation, Jacobian shift, optional exp)
        fused into the kernel: ``y`` is the RAW event, ``xform`` a ``functional.make_xform`` result."""
        return F.chain_forward(self.t.detach(), _to_tensor_like(y, self.t), self.flow_types, self.n_dims,
                               self.trainable_base_dist, xform=xform)

    def log_prob_grid(self, y_grid, xform=None):
        """[n_y, B]: every event of ``y_grid[n_y, d]`` against every batch row, parameters read once
        (what the reference's plot_model does with one ``dist.prob(y[i])`` call per grid line)."""
        return F.chain_forward_grid(self.t, _to_tensor_like(y_grid, self.t), self.flow_types, self.n_dims,
                                    self.trainable_base_dist, xform=xform)

    def prob_grid(self, y_grid):
        return torch.exp(self.log_prob_grid(y_grid))


class FusedDenseFlowChainDistribution(FlowChainDistribution):
    """The same density with the emitting Dense(P) layer folded into the kernel: holds the last hidden
    activation h[B, H] and the layer's (W[H, P], bias[P]) instead of t[B, P]; ``log_prob`` runs the fused
    dense+chain kernel (t is formed in shared memory and never written to HBM).  ``.t`` materialises the
    parameter rows on demand for callers that want them."""

    def __init__(self, h, W, bias, flow_types, n_dims, trainable_base_dist):
        self.h, self.W, self.bias = h, W, bias
        self.flow_types = tuple(flow_types)
        self.n_dims = n_dims
        self.trainable_base_dist = trainable_base_dist
        self._t = None

    @property
    def t(self):
        if self._t is None:
            self._t = torch.addmm(self.bias, self.h, self.W)
        return self._t

    @property
    def batch_shape(self):
        return _Shape(self.h.shape[:-1])

    def log_prob(self, y):
        y = _to_tensor_like(y, self.h)
        if torch.is_grad_enabled() and (self.h.requires_grad or self.W.requires_grad):
            return super().log_prob(y)  # autograd path goes through t
        if y.shape[0] not in (self.h.shape[0], 1):
            return super().log_prob(y)
        return F.dense_chain_forward(self.h, self.W, self.bias, y, self.flow_types, self.n_dims,
                                     self.trainable_base_dist)

    def log_prob_x(self, y, xform):
        y = _to_tensor_like(y, self.h)
        if y.shape[0] not in (self.h.shape[0], 1):
            return super().log_prob_x(y, xform)
        return F.dense_chain_forward(self.h.detach(), self.W.detach(), self.bias.detach(), y, self.flow_types,
                                     self.n_dims, self.trainable_base_dist, xform=xform)


class InverseNormalizingFlowLayer(torch.nn.Module):
    """Turns the network output into an inverted-flow density (reference :215-294).

    ``flow_types`` are applied in order from the data side: y passes through
    ``flow_types[0]`` first; parameters are sliced over the REVERSED list (reference
    :267-278), i.e. the last flow owns the first columns after the optional 2*n_dims
    base-distribution block.  This layer does not work for scalars.
    """

    def __init__(self, flow_types, n_dims, trainable_base_dist=False):
        super().__init__()
        assert all([flow_type in FLOWS for flow_type in flow_types])
        self._flow_types = tuple(flow_types)
        self._trainable_base_dist = trainable_base_dist
        self._n_dims = n_dims
        self._make = self._get_distribution_fn(n_dims, self._flow_types, trainable_base_dist)

    def forward(self, t):
        return self._make(t)

    @staticmethod
    def _get_distribution_fn(n_dims, flow_types, trainable_base_dist):
        flow_types = tuple(flow_types)
        width = sum(FLOWS[f].get_param_size(n_dims) for f in flow_types) + (
            2 * n_dims if trainable_base_dist else 0
        )

        def make(t):
            assert t.shape[-1] == width
            return FlowChainDistribution(t, flow_types, n_dims, trainable_base_dist)

        return make

    def get_total_param_size(self):
        num_flow_params = sum(FLOWS[f].get_param_size(self._n_dims) for f in self._flow_types)
        base_dist_params = 2 * self._n_dims if self._trainable_base_dist else 0
        return num_flow_params + base_dist_params

    @staticmethod
    def _get_bijector(t, flow_types, n_dims):
        flow_types = list(reversed(flow_types))
        param_sizes = [FLOWS[f].get_param_size(n_dims) for f in flow_types]
        assert sum(param_sizes) == t.shape[-1]
        begins = np.concatenate([[0], np.cumsum(param_sizes)[:-1]]).astype(int) if param_sizes else []
        return _Chain(
            FLOWS[f](t[..., int(b): int(b) + s].contiguous(), n_dims)
            for b, s, f in zip(begins, param_sizes, flow_types)
        )

    @staticmethod
    def _get_base_dist(t, n_dims, trainable):
        if trainable:
            loc = t[..., 0:n_dims]
            scale = 1e-3 + torch.nn.functional.softplus(
                math.log(math.expm1(1.0)) + 0.1 * t[..., n_dims: 2 * n_dims]
            )
        else:
            loc = torch.zeros_like(t[..., 0:n_dims])
            scale = torch.ones_like(t[..., 0:n_dims])
        return torch.distributions.Independent(torch.distributions.Normal(loc, scale), 1)


# ----------------------------------------------------------------------------- MDN
class GaussianMixtureDistribution:
    def __init__(self, t, n_centers, n_dims):
        self.t = t
        self.n_centers = n_centers
        self.n_dims = n_dims

    @property
    def event_shape(self):
        return _Shape([self.n_dims])

    @property
    def batch_shape(self):
        return _Shape(self.t.shape[:-1])

    def log_prob(self, y):
        return F.mdn_log_prob(self.t, _to_tensor_like(y, self.t), self.n_centers, self.n_dims)

    def log_prob_x(self, y, xform):
        """Forward-only, raw ``y``, the estimator's y pipeline fused into the kernel (see FlowChainDistribution)."""
        return F.mdn_forward(self.t.detach(), _to_tensor_like(y, self.t), self.n_centers, self.n_dims, xform=xform)

    def prob(self, y):
        return torch.exp(self.log_prob(y))


class FusedDenseGaussianMixtureDistribution(GaussianMixtureDistribution):
    """The same mixture with the emitting Dense(P) layer folded into the kernel (see
    FusedDenseFlowChainDistribution): holds h[B, H] and the layer's (W[H, P], bias[P]) instead of t[B, P];
    ``.t`` materialises the parameter rows on demand."""

    def __init__(self, h, W, bias, n_centers, n_dims):
        self.h, self.W, self.bias = h, W, bias
        self.n_centers = n_centers
        self.n_dims = n_dims
        self._t = None

    @property
    def t(self):
        if self._t is None:
            self._t = torch.addmm(self.bias, self.h, self.W)
        return self._t

    @property
    def batch_shape(self):
        return _Shape(self.h.shape[:-1])

    def log_prob(self, y):
        y = _to_tensor_like(y, self.h)
        if torch.is_grad_enabled() and (self.h.requires_grad or self.W.requires_grad):
            return super().log_prob(y)  # autograd path goes through t
        if y.shape[0] not in (self.h.shape[0], 1):
            return super().log_prob(y)
        return F.dense_mdn_forward(self.h, self.W, self.bias, y, self.n_centers, self.n_dims)

    def log_prob_x(self, y, xform):
        y = _to_tensor_like(y, self.h)
        if y.shape[0] not in (self.h.shape[0], 1):
            return super().log_prob_x(y, xform)
        return F.dense_mdn_forward(self.h.detach(), self.W.detach(), self.bias.detach(), y, self.n_centers,
                                   self.n_dims, xform=xform)


class GaussianMixtureLayer(torch.nn.Module):
    """Mixture of ``n_centers`` diagonal Gaussians with per-sample locs, softplus scales and
    mixture logits (reference :174-212)."""

    def __init__(self, n_centers, n_dims):
        super().__init__()
        self._n_centers = n_centers
        self._n_dims = n_dims
        self._make = self._get_distribution_fn(n_centers, n_dims)

    def forward(self, t):
        return self._make(t)

    def get_total_param_size(self):
        return self._n_centers + 2 * self._n_dims * self._n_centers

    @staticmethod
    def _get_distribution_fn(n_centers, n_dims):
        width = 2 * n_centers * n_dims + n_centers

        def make(t):
            if t.shape[-1] != width:  # TFP raises ValueError here (tests/test_distribution_layers.py:51,64)
                raise ValueError("GaussianMixtureLayer expects %d parameter columns, got %d" % (width, t.shape[-1]))
            return GaussianMixtureDistribution(t, n_centers, n_dims)

        return make


# ----------------------------------------------------------------------------- KMN
class GaussianKernelsDistribution:
    def __init__(self, t, locs, scales):
        self.t = t
        self.locs = locs      # [M, d]
        self.scales = scales  # [M]

    @property
    def event_shape(self):
        return _Shape([self.locs.shape[-1]])

    @property
    def batch_shape(self):
        return _Shape(self.t.shape[:-1])

    def log_prob(self, y):
        return F.kmn_log_prob(self.t, _to_tensor_like(y, self.t), self.locs.to(self.t.device),
                              self.scales.to(self.t.device))

    def log_prob_x(self, y, xform):
        """Forward-only, raw ``y``, the estimator's y pipeline fused into the kernel (see FlowChainDistribution)."""
        return F.kmn_forward(self.t.detach(), _to_tensor_like(y, self.t), self.locs.to(self.t.device),
                             self.scales.detach().to(self.t.device), xform=xform)

    def prob(self, y):
        return torch.exp(self.log_prob(y))

    def sample(self):
        """One draw per batch row (plain torch; sampling is not on the hot path)."""
        idx = torch.distributions.Categorical(logits=self.t).sample()
        locs = self.locs.to(self.t.device)[idx]
        scales = self.scales.to(self.t.device).abs()[idx]
        return locs + scales.unsqueeze(-1) * torch.randn_like(locs)


class FusedDenseGaussianKernelsDistribution(GaussianKernelsDistribution):
    """The kernel mixture with the emitting Dense(P) layer folded into the kernel (see
    FusedDenseFlowChainDistribution): holds h[B, H] and the layer's (W[H, M], bias[M]); ``.t`` (the logits)
    materialises on demand."""

    def __init__(self, h, W, bias, locs, scales):
        self.h, self.W, self.bias = h, W, bias
        self.locs = locs
        self.scales = scales
        self._t = None

    @property
    def t(self):
        if self._t is None:
            self._t = torch.addmm(self.bias, self.h, self.W)
        return self._t

    @property
    def batch_shape(self):
        return _Shape(self.h.shape[:-1])

    def log_prob(self, y):
        y = _to_tensor_like(y, self.h)
        if torch.is_grad_enabled() and (self.h.requires_grad or self.W.requires_grad or self.scales.requires_grad):
            return super().log_prob(y)  # autograd path goes through t
        if y.shape[0] not in (self.h.shape[0], 1):
            return super().log_prob(y)
        return F.dense_kmn_forward(self.h, self.W, self.bias, y, self.locs.to(self.h.device),
                                   self.scales.detach().to(self.h.device))

    def log_prob_x(self, y, xform):
        y = _to_tensor_like(y, self.h)
        if y.shape[0] not in (self.h.shape[0], 1):
            return super().log_prob_x(y, xform)
        return F.dense_kmn_forward(self.h.detach(), self.W.detach(), self.bias.detach(), y, self.locs.to(self.h.device),
                                   self.scales.detach().to(self.h.device), xform=xform)


class GaussianKernelsLayer(torch.nn.Module):
    """Kernel mixture: logits over fixed centres x trainable per-scale bandwidths
    (reference :74-171).  Bandwidth of scale group i is
    ``softplus(v_i) + log(expm1(init_i))`` exactly as in the reference (:88-95); it can be
    negative and only its magnitude enters the density."""

    def __init__(self, n_centers, n_dims, trainable_scale=True, init_scales=(0.3, 0.7)):
        super().__init__()
        self.n_centers = n_centers
        self.n_scales = len(init_scales)
        self.n_dims = n_dims
        self.init_scales = tuple(float(s) for s in init_scales)
        self.scale_vars = torch.nn.Parameter(torch.zeros(self.n_scales), requires_grad=trainable_scale)
        self.register_buffer("locs", torch.zeros(self.n_scales * n_centers, n_dims))
        self.register_buffer("_offsets", torch.tensor([math.log(math.expm1(s)) for s in self.init_scales]))

    def scale_model(self, _unused=0.0):
        s = torch.nn.functional.softplus(self.scale_vars) + self._offsets
        return s.repeat_interleave(self.n_centers)

    def get_total_param_size(self):
        return self.n_centers * self.n_scales

    def forward(self, t):
        return self._get_distribution_fn()(t)

    def _get_distribution_fn(self):
        def dist(t):
            assert t.shape[-1] == self.n_centers * self.n_scales
            return GaussianKernelsDistribution(t, self.locs, self.scale_model())

        return dist

    def set_center_points(self, y):
        """Host-side centre selection, once per fit (reference :135-171): the farthest
        points by cosine spread at the edges plus k-means centres, tiled over the scales."""
        from sklearn.cluster import KMeans
        from sklearn.metrics.pairwise import cosine_distances

        y = np.asarray(y)
        ndim_y = y.shape[1]
        n_edge_points = min(2 * ndim_y, self.n_centers // 2)
        farthest_idx = np.argsort(np.linalg.norm(y - y.mean(axis=0), axis=1))[-2 * n_edge_points:]
        y_far = y[farthest_idx]
        dists = cosine_distances(y_far)
        selected = [0]
        for _ in range(1, n_edge_points):
            nearest_selected = np.min(dists[:, selected], axis=1)
            selected.append(int(np.argsort(nearest_selected)[-1]))
        centers_at_edges = y_far[selected]
        y_rest = np.delete(y, farthest_idx[selected], axis=0)
        k = self.n_centers - n_edge_points
        km = KMeans(n_clusters=k, n_init=10, random_state=22).fit(y_rest)
        centers = np.concatenate([centers_at_edges, km.cluster_centers_], axis=0)
        tiled = np.concatenate([centers] * self.n_scales, axis=0).astype(np.float32)
        self.locs.copy_(torch.from_numpy(tiled).to(self.locs.device))


# ----------------------------------------------------------------------------- mean field
class MeanFieldLayer(torch.nn.Module):
    """n_dims independent normals parameterised by the input (reference :17-71); used for the
    weight posteriors/priors of the Bayesian estimators.  Plain torch: weight-space tensors
    are O(#weights), not O(batch), and are not on the hot path."""

    def __init__(self, n_dims, scale=None, map_mode=False, dtype=None):
        super().__init__()
        self.n_dims = n_dims
        self.scale = 1.0 if map_mode else scale
        self.map_mode = map_mode
        self._make = self._get_distribution_fn(self.n_dims, self.scale)

    def forward(self, t):
        return self._make(t)

    @staticmethod
    def _get_distribution_fn(n_dims, scale=None):
        if scale is None:

            def dist_fn(t):
                assert t.shape[-1] == 2 * n_dims
                sd = 1e-3 + torch.nn.functional.softplus(
                    math.log(math.expm1(1.0)) + 0.05 * t[..., n_dims: 2 * n_dims]
                )
                return _IndependentNormal(t[..., 0:n_dims], sd)

        else:
            assert scale > 0.0

            def dist_fn(t):
                assert t.shape[-1] == n_dims
                return _IndependentNormal(t[..., 0:n_dims], torch.full_like(t[..., 0:n_dims], float(scale)))

        return dist_fn

    def get_total_param_size(self):
        return 2 * self.n_dims if self.scale is None else self.n_dims


class _IndependentNormal(torch.distributions.Independent):
    def __init__(self, loc, scale):
        # no argument validation: it reduces every parameter tensor to a bool and reads it back on the host -- eight
        # device-to-host synchronisations per Bayesian training step (measured), and it cannot be graph-captured;
        # the scale is 1e-3 + softplus(.) or a positive constant by construction
        super().__init__(torch.distributions.Normal(loc, scale, validate_args=False), 1, validate_args=False)

    @property
    def event_shape(self):
        return _Shape(super().event_shape)

    @property
    def batch_shape(self):
        return _Shape(super().batch_shape)
