"""Literal torch-CPU restatement of the reference's flow-chain / mixture log-likelihood.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED (no TF/TFP here).

One tensor op per reference TF op, in the reference's order, so that
  * in float64 it is the parity oracle (gradients from torch autograd), and
  * in float32 with all host threads it is the timed "TF-equivalent CPU" baseline
    (``bench.py --impl reference`` / ``cpu_baseline``), since TensorFlow itself cannot
    run in this image.

Reference files followed (paths into /root/reference):
  estimators/normalizing_flows/PlanarFlow.py:20-80
  estimators/normalizing_flows/RadialFlow.py:20-84
  estimators/normalizing_flows/AffineFlow.py:4-10
  estimators/DistributionLayers.py:74-133 (KMN head), :174-212 (MDN), :215-294 (NF layer)
  estimators/BaseEstimator.py:55-86 (normalisation and the sum-log-y_std Jacobian term)
TFP glue semantics (Chain / Invert / TransformedDistribution / MultivariateNormalDiag /
Mixture / MixtureSameFamily / Affine) restated from SURVEY.md App. A.1.
"""
import math

import torch
import torch.nn.functional as F

FLOW_NAMES = ("planar", "radial", "affine")
LOG_EXPM1_1 = math.log(math.expm1(1.0))  # tf.math.log(tf.math.expm1(1.0))
HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


def flow_param_size(name, n_dims):
    """PlanarFlow.py:35-41, RadialFlow.py:36-42, AffineFlow.py:11-17."""
    if name == "planar":
        return n_dims + n_dims + 1
    if name == "radial":
        return 1 + 1 + n_dims
    if name == "affine":
        return 2 * n_dims
    raise KeyError(name)


def chain_param_size(flow_types, n_dims, trainable_base_dist):
    """DistributionLayers.py:257-265."""
    n = sum(flow_param_size(f, n_dims) for f in flow_types)
    return n + (2 * n_dims if trainable_base_dist else 0)


# --------------------------------------------------------------------------- bijectors
class PlanarOracle:
    """PlanarFlow.py:20-80, op for op."""

    def __init__(self, t, n_dims):
        assert t.shape[-1] == 2 * n_dims + 1
        u = t[..., 0:n_dims]
        w = t[..., n_dims : 2 * n_dims] + 1
        b = t[..., 2 * n_dims : 2 * n_dims + 1]
        # _u_circ (PlanarFlow.py:43-53)
        wtu = torch.sum(w * u, 1, keepdim=True)
        m_wtu = -1.0 + F.softplus(wtu) + 1e-5
        norm_w_squared = torch.sum(w ** 2, 1, keepdim=True) + 1e-9
        self.u = u + (m_wtu - wtu) * (w / norm_w_squared)
        self.w = w
        self.b = b

    def _wzb(self, z):
        return torch.sum(self.w * z, 1, keepdim=True) + self.b

    def forward(self, z):
        return z + self.u * torch.tanh(self._wzb(z))

    def fldj(self, z):
        psi = (1.0 - torch.tanh(self._wzb(z)) ** 2) * self.w
        det_grad = 1.0 + torch.sum(self.u * psi, 1)
        return torch.log(torch.abs(det_grad))


class RadialOracle:
    """RadialFlow.py:20-84, op for op (L1 radius, alpha*beta coupling, no abs on det)."""

    def __init__(self, t, n_dims):
        assert t.shape[-1] == n_dims + 2
        alpha = t[..., 0:1]
        beta = t[..., 1:2]
        self.gamma = t[..., 2 : n_dims + 2]
        self.alpha = F.softplus(0.3 * alpha - 2.0)
        self.beta = F.softplus(0.1 * beta + LOG_EXPM1_1) - 1.0
        self.n_dims = n_dims

    def _r(self, z):
        return torch.sum(torch.abs(z - self.gamma), 1, keepdim=True)

    def forward(self, z):
        r = self._r(z)
        h = 1.0 / (self.alpha + r)
        return z + (self.alpha * self.beta * h) * (z - self.gamma)

    def fldj(self, z):
        r = self._r(z)
        h = 1.0 / (self.alpha + r)
        # the reference takes dh/dr with an inner GradientTape (RadialFlow.py:62-65);
        # the tape result is itself differentiable, and equals -1/(alpha+r)^2.
        der_h = -1.0 / (self.alpha + r) ** 2
        ab = self.alpha * self.beta
        det = (1.0 + ab * h) ** (self.n_dims - 1) * (1.0 + ab * h + ab * der_h * r)
        return torch.log(det.squeeze(-1))


class AffineOracle:
    """AffineFlow.py:4-10 -> tfp.bijectors.Affine(shift, scale_diag=1+t[d:2d])."""

    def __init__(self, t, n_dims):
        assert t.shape[-1] == 2 * n_dims
        self.shift = t[..., 0:n_dims]
        self.scale = 1.0 + t[..., n_dims : 2 * n_dims]

    def forward(self, z):
        return self.scale * z + self.shift

    def fldj(self, z):
        ld = torch.sum(torch.log(torch.abs(self.scale)), -1)
        # Affine's fldj has the batch shape of its parameters; broadcast with z's rows
        return ld + torch.zeros(z.shape[0], dtype=z.dtype)


ORACLE_FLOWS = {"planar": PlanarOracle, "radial": RadialOracle, "affine": AffineOracle}


def build_bijectors(t_flows, flow_types, n_dims):
    """DistributionLayers.py:267-278: slices are assigned over the REVERSED flow list.

    Returns the bijector list in Chain order (bijectors[0] is the LAST entry of
    ``flow_types`` and owns the first columns).
    """
    rev = list(reversed(flow_types))
    sizes = [flow_param_size(f, n_dims) for f in rev]
    assert sum(sizes) == t_flows.shape[-1]
    begins = [sum(sizes[0:i]) for i in range(len(sizes))]
    return [ORACLE_FLOWS[f](t_flows[..., b : b + s], n_dims) for b, s, f in zip(begins, sizes, rev)]


def chain_log_prob(t, y, flow_types, n_dims, trainable_base_dist):
    """TransformedDistribution(base, Invert(Chain(bijectors))).log_prob(y).

    DistributionLayers.py:245-255, :280-294 + TFP semantics:
    Chain.forward applies the LAST bijector first, i.e. data passes through
    ``flow_types`` in the given order; the log-dets are summed at each flow's own input.
    ``y`` may be ``[1, d]`` (broadcast against the batch) or ``[B, d]``.
    """
    assert t.shape[-1] == chain_param_size(flow_types, n_dims, trainable_base_dist)
    t_flows = t[..., 2 * n_dims :] if trainable_base_dist else t
    bijectors = build_bijectors(t_flows, flow_types, n_dims)
    z = y
    ldj = torch.zeros((), dtype=t.dtype)
    for bij in reversed(bijectors):
        ldj = ldj + bij.fldj(z)
        z = bij.forward(z)
    if trainable_base_dist:
        loc = t[..., 0:n_dims]
        scale = 1e-3 + F.softplus(LOG_EXPM1_1 + 0.1 * t[..., n_dims : 2 * n_dims])
    else:
        loc = torch.zeros_like(t[..., 0:n_dims])
        scale = torch.ones_like(t[..., 0:n_dims])
    e = (z - loc) / scale
    base = -0.5 * torch.sum(e * e, -1) - torch.sum(torch.log(scale), -1) - n_dims * HALF_LOG_2PI
    return base + ldj


def mdn_log_prob(t, y, n_centers, n_dims):
    """GaussianMixtureLayer (DistributionLayers.py:196-212) + tfd.Mixture.log_prob."""
    assert t.shape[-1] == 2 * n_centers * n_dims + n_centers
    comps = []
    for loc_start in range(0, 2 * n_centers * n_dims, 2 * n_dims):
        loc = t[..., loc_start : loc_start + n_dims]
        scale = F.softplus(0.05 * t[..., loc_start + n_dims : loc_start + 2 * n_dims] + LOG_EXPM1_1)
        e = (y - loc) / scale
        comps.append(
            -0.5 * torch.sum(e * e, -1) - torch.sum(torch.log(scale), -1) - n_dims * HALF_LOG_2PI
        )
    logits = t[..., 2 * n_centers * n_dims : 2 * n_centers * n_dims + n_centers]
    lp = torch.stack(comps, -1) + torch.log_softmax(logits, -1)
    return torch.logsumexp(lp, -1)


def kmn_scales(scale_vars, n_centers, init_scales):
    """GaussianKernelsLayer.scale_model (DistributionLayers.py:79-99): one bandwidth per
    scale group, ``softplus(v_i) + log(expm1(init_i))`` (may be negative, App. B.7)."""
    parts = [
        torch.zeros(n_centers, dtype=scale_vars.dtype)
        + F.softplus(scale_vars[i])
        + math.log(math.expm1(init_scales[i]))
        for i in range(len(init_scales))
    ]
    return torch.cat(parts, 0)


def kmn_log_prob(t, y, locs, scales):
    """GaussianKernelsLayer._get_distribution_fn (DistributionLayers.py:118-133) +
    tfd.MixtureSameFamily.log_prob with MultivariateNormalDiag(scale_identity_multiplier).

    ``locs`` [M, d] fixed centres, ``scales`` [M] isotropic bandwidths (sign ignored by
    the density: TFP uses |scale| in the log-det and scale^2 in the quadratic form).
    """
    M, d = locs.shape
    assert t.shape[-1] == M
    e = (y[..., None, :] - locs[None, :, :]) / scales[None, :, None]
    comp = -0.5 * torch.sum(e * e, -1) - d * torch.log(torch.abs(scales))[None, :] - d * HALF_LOG_2PI
    lp = comp + torch.log_softmax(t, -1)
    return torch.logsumexp(lp, -1)


# --------------------------------------------------------------------------- estimator-level
def normalise_y(y, y_mean, y_std):
    """BaseEstimator.py:61-66 (no epsilon on y_std)."""
    return (y - torch.ones_like(y) * y_mean) / y_std


def nll(logp, y_std):
    """BaseEstimator.py:55-59: -log_prob + sum(log y_std), per sample."""
    return -logp + torch.sum(torch.log(y_std))


# --------------------------------------------------------------------------- helpers
def with_grad(fn, t, y, *args, upstream=None, want_dy=False):
    """Evaluate ``fn(t, y, *args)`` in the dtype of ``t`` and return (logp, dt[, dy])
    where the cotangent of logp is ``upstream`` (default: ones)."""
    t = t.detach().clone().requires_grad_(True)
    y = y.detach().clone().requires_grad_(want_dy)
    logp = fn(t, y, *args)
    g = torch.ones_like(logp) if upstream is None else upstream
    grads = torch.autograd.grad(logp, [t, y] if want_dy else [t], grad_outputs=g)
    if want_dy:
        return logp.detach(), grads[0], grads[1]
    return logp.detach(), grads[0]
