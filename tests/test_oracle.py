"""CPU: the oracle against the frozen golden vectors and against itself (three independent
restatements), plus the structural invariants the reference's own tests pin.

PARITY UNPINNED at the TF/TFP boundary (see oracle/__init__.py): the anchors are the
float64 literal restatement + autograd, the closed-form NumPy oracle, 50-digit mpmath on the
reference's `tf.ones` test inputs, and the values derived independently in SURVEY.md A.7.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import analytic_np as an
from oracle import flow_oracle as fo
from oracle import known_answers_mp as mpo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# SURVEY.md Appendix A.7 (derived by the surveyor with a separate scratch implementation)
SURVEY_SINGLE = {
    ("planar", 1, 0.0): (0.42913470175048546, 0.387494634929230),
    ("planar", 1, 1.0): (1.5606825126078827, 0.011057057118784),
    ("planar", 4, 0.0): (0.6664277689016205, 1.371171955032434),
    ("planar", 4, 1.0): (1.8750431491506423, 0.000000426460371),
    ("radial", 1, 0.0): (-0.009247761883502171, 0.001327824857924),
    ("radial", 1, 1.0): (1.0, 0.062377589623311),
    ("radial", 4, 0.0): (-0.002591161603426877, 0.007867740104109),
    ("radial", 4, 1.0): (1.0, 0.249510358493246),
    ("affine", 1, 0.0): (1.0, 0.693147180559945),
    ("affine", 1, 1.0): (3.0, 0.693147180559945),
    ("affine", 4, 0.0): (1.0, 2.772588722239781),
    ("affine", 4, 1.0): (3.0, 2.772588722239781),
}
SURVEY_LAYER = [
    (("radial", "planar"), 1, False, 1.0, 0.0, -0.607325135824565),
    (("radial", "planar"), 1, False, 1.0, 0.5, -1.370997076823677),
    (("radial", "planar"), 1, False, 0.0, 0.0, -1.285437026497321),
    (("radial", "planar"), 1, False, 0.0, 0.5, -1.259262823128820),
    (("radial", "planar"), 2, True, 1.0, 0.0, -1.291524729500954),
    (("radial", "planar"), 2, True, 1.0, 0.5, -1.976888594254271),
    (("radial", "planar"), 2, True, 0.0, 0.0, -2.206374560589498),
    (("planar", "radial", "affine"), 1, False, 1.0, 0.0, -1.530683715706439),
    (("planar", "radial", "affine"), 1, False, 1.0, 0.5, -4.885873955865832),
]


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def test_known_answers_match_survey_and_mpmath():
    ka = load("known_answers.json")
    seen = 0
    for c in ka["single_flow"]:
        key = (c["flow"], c["n_dims"], c["z"])
        fwd, fldj = SURVEY_SINGLE[key]
        assert c["forward"][0] == pytest.approx(fwd, rel=1e-12, abs=1e-15)
        assert c["fldj"] == pytest.approx(fldj, rel=1e-8, abs=1e-15)
        th = [mpo.mp.mpf(1)] * mpo.psize(c["flow"], c["n_dims"])
        z2, fl = mpo.flow_step(c["flow"], th, [mpo.mp.mpf(c["z"])] * c["n_dims"])
        assert float(fl) == pytest.approx(c["fldj"], rel=1e-14, abs=1e-18)
        assert [float(v) for v in z2] == pytest.approx(c["forward"], rel=1e-14)
        seen += 1
    assert seen == 12
    for ft, d, tb, t, y, lp in SURVEY_LAYER:
        P = fo.chain_param_size(ft, d, tb)
        got = fo.chain_log_prob(torch.full((1, P), t, dtype=torch.float64),
                                torch.full((1, d), y, dtype=torch.float64), ft, d, tb)
        assert float(got[0]) == pytest.approx(lp, rel=1e-13)
    for c in ka["layer"]:
        P = fo.chain_param_size(c["flow_types"], c["n_dims"], c["trainable_base_dist"])
        got = an.chain_forward_backward(np.full((1, P), c["t"]), np.full((1, c["n_dims"]), c["y"]),
                                        c["flow_types"], c["n_dims"], c["trainable_base_dist"], need_grad=False)
        assert got[0] == pytest.approx(c["log_prob"], rel=1e-13)
    for c in ka["mdn"]:
        K, d = c["n_centers"], c["n_dims"]
        got = an.mdn_forward_backward(np.full((1, 2 * K * d + K), c["t"]), np.full((1, d), c["y"]), K, d,
                                      need_grad=False)
        assert got[0] == pytest.approx(c["log_prob"], rel=1e-13)


def test_golden_chain_vectors_reproduce():
    """Both oracles reproduce the frozen vectors (guards against silent oracle drift)."""
    for c in load("chain_vectors.json"):
        ft, d, tb = c["flow_types"], c["n_dims"], c["trainable_base_dist"]
        t, y, up = np.array(c["t"], np.float32), np.array(c["y"], np.float32), np.array(c["upstream"], np.float32)
        lp, dt, dy = an.chain_forward_backward(t, y, ft, d, tb, upstream=up)
        np.testing.assert_allclose(lp, c["log_prob"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(dt, np.array(c["dt"]).reshape(dt.shape), rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(dy, c["dy"], rtol=1e-9, atol=1e-11)
        t64, y64 = torch.tensor(t, dtype=torch.float64), torch.tensor(y, dtype=torch.float64)
        lp2, dt2 = fo.with_grad(fo.chain_log_prob, t64, y64, ft, d, tb, upstream=torch.tensor(up, dtype=torch.float64))
        np.testing.assert_allclose(lp2.numpy(), c["log_prob"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(dt2.numpy(), np.array(c["dt"]).reshape(dt.shape), rtol=1e-9, atol=1e-11)
        lb = fo.chain_log_prob(t64, y64[3:4], ft, d, tb)
        np.testing.assert_allclose(lb.numpy(), c["log_prob_y_row3_broadcast"], rtol=1e-12, atol=1e-12)


def test_golden_mixture_vectors_reproduce():
    mv = load("mixture_vectors.json")
    for c in mv["mdn"]:
        t, y, up = np.array(c["t"], np.float32), np.array(c["y"], np.float32), np.array(c["upstream"], np.float32)
        lp, dt, dy = an.mdn_forward_backward(t, y, c["n_centers"], c["n_dims"], upstream=up)
        np.testing.assert_allclose(lp, c["log_prob"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(dt, c["dt"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(dy, c["dy"], rtol=1e-9, atol=1e-12)
    for c in mv["kmn"]:
        t, y, up = np.array(c["t"], np.float32), np.array(c["y"], np.float32), np.array(c["upstream"], np.float32)
        lp, dt, ds, dy = an.kmn_forward_backward(t, y, np.array(c["locs"], np.float32), np.array(c["scales"]),
                                                 upstream=up)
        np.testing.assert_allclose(lp, c["log_prob"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(dt, c["dt"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(ds, c["dscales"], rtol=1e-9, atol=1e-12)
        # negative bandwidth of the 0.3 scale group is part of the fixture
        assert min(c["scales"]) < 0 < max(c["scales"])
        # chain rule through scale_model: d/d scale_vars = sum over the group of dscales * sigmoid(v)
        nc = c["n_centers"]
        sv = np.array(c["scale_vars"])
        grp = np.array([ds[i * nc:(i + 1) * nc].sum() for i in range(len(sv))]) * an.sigmoid(sv)
        np.testing.assert_allclose(grp, c["dscale_vars"], rtol=1e-8, atol=1e-12)


def test_parameter_order_is_reversed():
    """DistributionLayers.py:267-278 / tests/test_distribution_layers.py:185-193: the LAST flow
    of flow_types owns the first columns; bijectors[0] is affine for (planar, radial, affine)."""
    ft, d = ("planar", "radial", "affine"), 1
    bij = fo.build_bijectors(torch.zeros(10, 8, dtype=torch.float64), ft, d)
    assert [type(b).__name__ for b in bij] == ["AffineOracle", "RadialOracle", "PlanarOracle"]
    offs, P = an.layout(ft, d, False)
    assert P == 8 and offs == [5, 2, 0]
    offs, P = an.layout(ft, 3, True)
    assert P == 24 and offs == [6 + 6 + 5, 6 + 6, 6]
    # perturbing the first two columns (affine block) changes the output like an affine flow does
    t = torch.zeros(1, 8, dtype=torch.float64)
    y = torch.tensor([[0.3]], dtype=torch.float64)
    base = fo.chain_log_prob(t, y, ft, d, False)
    t2 = t.clone()
    t2[0, 1] = 1.0  # affine scale_raw: z -> 2 z, fldj += log 2
    lp2 = fo.chain_log_prob(t2, y, ft, d, False)
    # identity radial at t=0, so z_after_affine doubles: check through the base density
    zs = y
    for b in reversed(fo.build_bijectors(t, ft, d)):
        zs = b.forward(zs)
    expect = float(base) + 0.5 * float(zs[0, 0]) ** 2 - 0.5 * (2 * float(zs[0, 0])) ** 2 + np.log(2.0)
    assert float(lp2) == pytest.approx(expect, rel=1e-12)


def test_row_independence_and_broadcast():
    rng = np.random.default_rng(1)
    ft, d, tb = ["planar", "radial", "affine"], 3, True
    P = fo.chain_param_size(ft, d, tb)
    t = rng.normal(0, 1, (17, P))
    y = rng.normal(0, 1, (17, d))
    lp = an.chain_forward_backward(t, y, ft, d, tb, need_grad=False)
    perm = rng.permutation(17)
    lp_p = an.chain_forward_backward(t[perm], y[perm], ft, d, tb, need_grad=False)
    np.testing.assert_array_equal(lp[perm], lp_p)
    lb = an.chain_forward_backward(t, y[:1], ft, d, tb, need_grad=False)
    lfull = an.chain_forward_backward(t, np.repeat(y[:1], 17, 0), ft, d, tb, need_grad=False)
    np.testing.assert_array_equal(lb, lfull)


def test_quirks_app_b():
    # B.4: t = 0 is the identity for radial and affine, NOT for planar
    y = np.zeros((1, 1))
    std_normal = -0.5 * np.log(2 * np.pi)
    assert an.chain_forward_backward(np.zeros((1, 3)), y, ["radial"], 1, False, need_grad=False)[0] == pytest.approx(std_normal)
    assert an.chain_forward_backward(np.zeros((1, 2)), y, ["affine"], 1, False, need_grad=False)[0] == pytest.approx(std_normal)
    planar0 = an.chain_forward_backward(np.zeros((1, 3)), y, ["planar"], 1, False, need_grad=False)[0]
    assert abs(planar0 - std_normal) > 0.1
    # B.1: L1 radius -- for d=2 the radial flow with gamma=0 depends on |z1|+|z2|
    t = np.array([[0.5, 2.0, 0.0, 0.0]])
    a = an.chain_forward_backward(t, np.array([[0.3, 0.4]]), ["radial"], 2, False, need_grad=False)
    b = an.chain_forward_backward(t, np.array([[0.7, 0.0]]), ["radial"], 2, False, need_grad=False)
    # same L1 radius (0.7) -> same log-det; base term differs, so compare log-dets
    ra = fo.RadialOracle(torch.tensor(t), 2).fldj(torch.tensor([[0.25, 0.5]], dtype=torch.float64))
    rb = fo.RadialOracle(torch.tensor(t), 2).fldj(torch.tensor([[0.75, 0.0]], dtype=torch.float64))
    assert float(ra) == pytest.approx(float(rb), rel=1e-14)
    # ... and it is NOT the L2 radius: same L2 norm, different L1 norm -> different log-det
    rc = fo.RadialOracle(torch.tensor(t), 2).fldj(torch.tensor([[0.3, 0.4]], dtype=torch.float64))
    rd = fo.RadialOracle(torch.tensor(t), 2).fldj(torch.tensor([[0.5, 0.0]], dtype=torch.float64))
    assert abs(float(rc) - float(rd)) > 1e-3
    assert a[0] != b[0]
    # sign(0) = 0 sub-gradient: y exactly on the radial centre gives a finite gradient
    t = np.array([[0.2, 0.7, 0.25]])
    _, dt, dy = an.chain_forward_backward(t, np.array([[0.25]]), ["radial"], 1, False)
    tt = torch.tensor(t, requires_grad=True)
    yy = torch.tensor([[0.25]], dtype=torch.float64, requires_grad=True)
    lp = fo.chain_log_prob(tt, yy, ["radial"], 1, False)
    g_t, g_y = torch.autograd.grad(lp.sum(), [tt, yy])
    np.testing.assert_allclose(dt, g_t.numpy(), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(dy, g_y.numpy(), rtol=1e-12, atol=1e-14)


def test_fp32_literal_restatement_tracks_fp64():
    """The fp32 run of the literal restatement (the timed CPU baseline) stays close to fp64."""
    rng = np.random.default_rng(22)
    ft, d, tb = ["radial"] * 3, 1, True
    t = rng.normal(0, 0.5, (4096, 11)).astype(np.float32)
    y = rng.normal(0, 1, (4096, 1)).astype(np.float32)
    lp32 = fo.chain_log_prob(torch.tensor(t), torch.tensor(y), ft, d, tb).numpy()
    lp64 = fo.chain_log_prob(torch.tensor(t, dtype=torch.float64), torch.tensor(y, dtype=torch.float64), ft, d, tb).numpy()
    assert np.max(np.abs(lp32 - lp64) / np.maximum(1, np.abs(lp64))) < 1e-5


def test_estimator_level_terms():
    """BaseEstimator.py:55-66,86: nll = -log_prob(normalised y) + sum(log y_std)."""
    y = torch.tensor([[1.0, 4.0]], dtype=torch.float64)
    y_mean = torch.tensor([0.5, 1.0], dtype=torch.float64)
    y_std = torch.tensor([2.0, 3.0], dtype=torch.float64)
    yn = fo.normalise_y(y, y_mean, y_std)
    np.testing.assert_allclose(yn.numpy(), [[0.25, 1.0]])
    lp = torch.tensor([-1.5], dtype=torch.float64)
    assert float(fo.nll(lp, y_std)[0]) == pytest.approx(1.5 + np.log(6.0))


# ----------------------------------------------------------------------------- first-principles pin
def _np_logp(t_row, y, ft, d, tb):
    out = an.chain_forward_backward(np.repeat(t_row, y.shape[0], 0), y, ft, d, tb, need_grad=False)
    return out[0] if isinstance(out, tuple) else out


@pytest.mark.parametrize("ft", [["radial"] * 3, ["radial"] * 5, ["planar"] * 4, ["planar", "radial", "affine"],
                                ["affine", "radial", "planar", "radial"]], ids=lambda f: "".join(x[0] for x in f))
@pytest.mark.parametrize("tb", [True, False])
def test_oracle_density_is_normalised_1d(ft, tb):
    """Independent of TFP, the shim and torch.distributions: whatever `TransformedDistribution(base,
    Invert(Chain(flows)))` means, it is a probability density in y, so exp(log_prob) must integrate to 1 for ANY
    parameter row.  A misread direction of the chain or sign of a log-det-Jacobian fails this at once (the reading
    restated in SURVEY.md App. A.1 and in both oracles is: log p(y) = base.log_prob(f(y)) + sum fldj_k, y passing
    through flow_types[0] first).  1-D events, trapezoid rule on a grid that covers the mass."""
    y = np.linspace(-60.0, 60.0, 240001)[:, None]
    for seed in range(3):
        P = an_param_size(ft, 1, tb)
        t = np.random.default_rng(seed).normal(0.0, 0.5, (1, P))
        lp = _np_logp(t, y, ft, 1, tb)
        assert abs(np.trapezoid(np.exp(lp), y[:, 0]) - 1.0) <= 1e-6, (ft, tb, seed)
        # the literal (torch float64) restatement gives the same density on a coarser slice
        ys = y[::400]
        lp_t = fo.chain_log_prob(torch.tensor(np.repeat(t, ys.shape[0], 0)), torch.tensor(ys), ft, 1, tb).numpy()
        assert np.max(np.abs(lp_t - lp[::400])) <= 1e-10


def an_param_size(ft, d, tb):
    return (2 * d if tb else 0) + sum(an.SIZES[f](d) for f in ft)


@pytest.mark.parametrize("ft,seeds", [(["planar", "radial", "affine"] * 3 + ["planar"], (0, 2)),   # BASELINE config 2
                                      (["radial", "planar"] * 2, (0, 1)), (["radial"] * 3, (0,))],
                         ids=["cfg2", "rprp", "rrr"])
def test_oracle_density_is_normalised_2d(ft, seeds):
    """The same in two dimensions, where the reference's radial flow uses an L1 radius and a tape-derived
    log-det (RadialFlow.py:51-70): the restated closed form still is the log-det of the map it applies."""
    g = np.linspace(-30.0, 30.0, 1201)
    Y = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
    for seed in seeds:
        t = np.random.default_rng(seed).normal(0.0, 0.3, (1, an_param_size(ft, 2, True)))
        p = np.exp(_np_logp(t, Y, ft, 2, True)).reshape(g.size, g.size)
        total = np.trapezoid(np.trapezoid(p, g, axis=1), g)
        assert abs(total - 1.0) <= 1e-4, (ft, seed, total)


def test_mixture_heads_are_normalised_1d():
    """MDN (Mixture of MVNDiag, DistributionLayers.py:196-212) and KMN (MixtureSameFamily over fixed centres,
    :118-133; negative bandwidths are legal) as restated: densities in y."""
    y = np.linspace(-60.0, 60.0, 120001)[:, None]
    rng = np.random.default_rng(4)
    K = 5
    t = rng.normal(0.0, 0.7, (1, 3 * K))
    lp = an.mdn_forward_backward(np.repeat(t, y.shape[0], 0), y, K, 1, need_grad=False)
    lp = lp[0] if isinstance(lp, tuple) else lp
    assert abs(np.trapezoid(np.exp(lp), y[:, 0]) - 1.0) <= 1e-6
    M = 12
    locs, scales = rng.normal(0.0, 2.0, (M, 1)), rng.normal(0.0, 0.6, M)   # some bandwidths negative
    assert (scales < 0).any()
    lp = an.kmn_forward_backward(np.repeat(rng.normal(0.0, 1.0, (1, M)), y.shape[0], 0), y, locs, scales, need_grad=False)
    lp = lp[0] if isinstance(lp, tuple) else lp
    assert abs(np.trapezoid(np.exp(lp), y[:, 0]) - 1.0) <= 1e-6
