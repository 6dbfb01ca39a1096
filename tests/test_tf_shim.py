"""CPU: the TFP glue restated in ``oracle/tf_shim.py`` against an independent implementation of the same
semantics, ``torch.distributions`` (itself modelled on TFP).  This is the part of the parity chain that is
NOT pinned by running the reference's code: seven small closed forms.  Agreement with a
second library narrows what "unpinned at the TFP boundary" can hide to both libraries sharing a misreading.
"""
import math

import numpy as np
import pytest
import torch
import torch.distributions as td

from oracle import tf_shim

F64 = torch.float64


@pytest.fixture()
def tfp():
    tf, tfp_mod = tf_shim.install(F64)
    yield tf, tfp_mod
    tf_shim.uninstall()


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(22 + seed)
    return scale * torch.randn(*shape, generator=g, dtype=F64)


def test_mvn_diag_matches_torch(tfp):
    _, m = tfp
    loc, scale, x = rnd(7, 3), rnd(7, 3, seed=1).abs() + 0.1, rnd(7, 3, seed=2)
    ours = m.distributions.MultivariateNormalDiag(loc=loc, scale_diag=scale)
    ref = td.Independent(td.Normal(loc, scale), 1)
    assert torch.allclose(ours.log_prob(x), ref.log_prob(x), rtol=1e-13, atol=1e-13)
    assert ours.event_shape == [3] and ours.batch_shape == [7]
    # one event broadcast against the batch, as the estimators' pdf(x, y) callers do
    assert torch.allclose(ours.log_prob(x[:1]), ref.log_prob(x[:1]), rtol=1e-13, atol=1e-13)
    # scale_identity_multiplier of shape [M]: batch of M isotropic normals; a NEGATIVE multiplier enters as |s|
    # (LinearOperator log_abs_determinant; GaussianKernelsLayer produces them)
    locs, s = rnd(5, 4, 2, seed=3), torch.tensor([0.3, -0.7, 1.1, -0.2], dtype=F64)
    ours = m.distributions.MultivariateNormalDiag(loc=locs, scale_identity_multiplier=s)
    ref = td.Independent(td.Normal(locs, s.abs()[:, None].expand(4, 2)), 1)
    y = rnd(5, 1, 2, seed=4)
    assert torch.allclose(ours.log_prob(y), ref.log_prob(y), rtol=1e-13, atol=1e-13)
    assert ours.batch_shape == [5, 4]


def test_mixture_and_mixture_same_family_match_torch(tfp):
    _, m = tfp
    B, K, d = 6, 4, 3
    locs, scales = rnd(B, K, d), rnd(B, K, d, seed=1).abs() + 0.2
    logits, x = rnd(B, K, seed=2, scale=2.0), rnd(B, d, seed=3)
    ref = td.MixtureSameFamily(td.Categorical(logits=logits), td.Independent(td.Normal(locs, scales), 1))
    comps = [m.distributions.MultivariateNormalDiag(loc=locs[:, k], scale_diag=scales[:, k]) for k in range(K)]
    mix = m.distributions.Mixture(cat=m.distributions.Categorical(logits=logits), components=comps)
    assert torch.allclose(mix.log_prob(x), ref.log_prob(x), rtol=1e-13, atol=1e-13)
    assert mix.event_shape == [d] and mix.batch_shape == [B]
    msf = m.distributions.MixtureSameFamily(
        mixture_distribution=m.distributions.Categorical(logits=logits),
        components_distribution=m.distributions.MultivariateNormalDiag(loc=locs, scale_diag=scales))
    assert torch.allclose(msf.log_prob(x), ref.log_prob(x), rtol=1e-13, atol=1e-13)
    assert tuple(msf.sample().shape) == (B, d)
    with pytest.raises(ValueError):  # TFP validates the static shapes (reference tests/test_distribution_layers.py:51)
        m.distributions.Mixture(cat=m.distributions.Categorical(logits=logits[:, :3]), components=comps)


def test_affine_chain_invert_transformed_distribution_match_torch(tfp):
    _, m = tfp
    B, d = 5, 3
    shifts = [rnd(B, d, seed=i) for i in range(3)]
    scales = [rnd(B, d, seed=10 + i) + 1.5 * torch.sign(rnd(B, d, seed=20 + i)) for i in range(3)]  # both signs
    ours = [m.bijectors.Affine(shift=sh, scale_diag=sc) for sh, sc in zip(shifts, scales)]
    theirs = [td.AffineTransform(sh, sc, event_dim=1) for sh, sc in zip(shifts, scales)]
    x = rnd(B, d, seed=30)
    for a, b in zip(ours, theirs):
        assert torch.allclose(a.forward(x), b(x), rtol=1e-13, atol=1e-13)
        assert torch.allclose(a.forward_log_det_jacobian(x, event_ndims=1), b.log_abs_det_jacobian(x, b(x)),
                              rtol=1e-13, atol=1e-13)
    # Chain([b0, b1, b2]).forward applies b2 first
    chain = m.bijectors.Chain(ours)
    composed = td.ComposeTransform([theirs[2], theirs[1], theirs[0]])
    assert torch.allclose(chain.forward(x), composed(x), rtol=1e-13, atol=1e-13)
    assert torch.allclose(chain.forward_log_det_jacobian(x, event_ndims=1),
                          composed.log_abs_det_jacobian(x, composed(x)), rtol=1e-12, atol=1e-12)
    assert [type(b) for b in chain.bijectors] == [type(b) for b in ours]
    # TransformedDistribution(base, Invert(chain)).log_prob(y) = base.log_prob(chain(y)) + fldj_chain(y)
    loc, scale = rnd(B, d, seed=40), rnd(B, d, seed=41).abs() + 0.3
    dist = m.distributions.TransformedDistribution(
        distribution=m.distributions.MultivariateNormalDiag(loc=loc, scale_diag=scale),
        bijector=m.bijectors.Invert(chain))
    ref = td.TransformedDistribution(td.Independent(td.Normal(loc, scale), 1), [composed.inv])
    y = rnd(B, d, seed=42)
    assert torch.allclose(dist.log_prob(y), ref.log_prob(y), rtol=1e-12, atol=1e-12)
    assert dist.event_shape == [d] and dist.batch_shape == [B]
    assert m.bijectors.Invert(chain).inverse_min_event_ndims == 1


def test_independent_normal_and_exact_kl_match_torch(tfp):
    _, m = tfp
    n = 11
    lq, sq = rnd(n), rnd(n, seed=1).abs() + 0.05
    lr, sr = rnd(n, seed=2), torch.full((n,), 0.7, dtype=F64)
    q = m.distributions.Independent(m.distributions.Normal(loc=lq, scale=sq), reinterpreted_batch_ndims=1)
    r = m.distributions.Independent(m.distributions.Normal(loc=lr, scale=0.7), reinterpreted_batch_ndims=1)
    tq, tr = td.Independent(td.Normal(lq, sq), 1), td.Independent(td.Normal(lr, sr), 1)
    w = rnd(n, seed=3)
    assert torch.allclose(q.log_prob(w), tq.log_prob(w), rtol=1e-13, atol=1e-13)
    assert q.event_shape == [n] and q.batch_shape == []
    assert torch.allclose(tf_shim._kl_divergence(q, r), td.kl_divergence(tq, tr), rtol=1e-13, atol=1e-13)


def test_softplus_tape_and_dense_variational_semantics(tfp):
    tf, m = tfp
    x = torch.tensor([-800.0, -30.0, -1.0, 0.0, 1.0, 30.0, 800.0], dtype=F64)
    want = np.logaddexp(0.0, x.numpy())  # no large-x shortcut
    assert np.allclose(tf.nn.softplus(x).numpy(), want, rtol=1e-15, atol=0.0)
    assert float(tf.math.log(tf.math.expm1(1.0))) == pytest.approx(math.log(math.e - 1.0), rel=1e-15)
    # GradientTape.gradient of a non-scalar target is the gradient of its sum
    r = torch.tensor([[0.5], [2.0]], dtype=F64)
    with tf.GradientTape() as g:
        g.watch(r)
        h = 1.0 / (0.3 + r)
    assert torch.allclose(g.gradient(h, r), -1.0 / (0.3 + r) ** 2, rtol=1e-14, atol=0.0)
    # DenseVariational: kernel = first in*units entries reshaped [in, units], bias = the rest
    size = 3 * 2 + 2
    post = lambda k, b, dtype=None: tf.keras.Sequential([
        m.layers.VariableLayer(k + b, initializer="zeros"),
        m.layers.DistributionLambda(
            make_distribution_fn=lambda t: m.distributions.Independent(
                m.distributions.Normal(loc=t, scale=1.0), reinterpreted_batch_ndims=1),
            convert_to_tensor_fn=m.distributions.Distribution.mean)])
    layer = m.layers.DenseVariational(2, make_posterior_fn=post, make_prior_fn=post, kl_weight=0.5,
                                      kl_use_exact=True, activation="linear")
    layer.build(3)
    w = torch.arange(1.0, size + 1.0, dtype=F64)
    layer._posterior.layers[0].variable = w
    xin = rnd(4, 3)
    out = layer(xin)
    assert torch.allclose(out, xin @ w[:6].reshape(3, 2) + w[6:], rtol=1e-14, atol=1e-14)
    assert float(layer.losses[0]) == pytest.approx(0.5 * 0.5 * float((w ** 2).sum()), rel=1e-13)  # KL(N(w,1) || N(0,1))
