"""CPU: the oracle against outputs of the reference's OWN code.

``tests/golden/reference_run.json`` was produced by ``python -m oracle.make_reference_run``: the
reference's ``estimators/normalizing_flows/*.py`` and ``estimators/DistributionLayers.py``, imported
unmodified from /root/reference and executed on torch-CPU float64 stand-ins for the TF ops and the
TFP glue classes (``oracle/tf_shim.py``), on the inputs of the other golden fixtures.  That pins the
reference's in-repo formulas, constants, slicing and ordering; TFP's own glue arithmetic and float32
rounding stay restated (parity "unpinned" at the TFP boundary).

Where /root/reference is present (this container, not the GPU box) the reference code is run again
and must reproduce the committed fixture.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import analytic_np as an
from oracle import flow_oracle as fo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HAVE_REFERENCE = os.path.isdir("/root/reference/estimators")


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def t64(a):
    return torch.tensor(np.asarray(a, dtype=np.float64), dtype=torch.float64)


def paired(ref_cases, cases):
    assert [(r["name"], r.get("sigma")) for r in ref_cases] == [(c["name"], c.get("sigma")) for c in cases]
    return zip(ref_cases, cases)


def test_fixture_provenance_is_stated():
    prov = load("reference_run.json")["provenance"]
    assert "unmodified" in prov["what"] and "oracle/tf_shim.py" in prov["what"]
    assert "TFP" in prov["does_not_pin"]


def test_single_flows_and_known_layers_match_reference_code():
    ref, ka = load("reference_run.json"), load("known_answers.json")
    assert len(ref["single_flow"]) == len(ka["single_flow"]) == 12
    for r, c in zip(ref["single_flow"], ka["single_flow"]):
        assert (r["flow"], r["n_dims"], r["z"]) == (c["flow"], c["n_dims"], c["z"])
        np.testing.assert_allclose(r["forward"], c["forward"], rtol=1e-12, atol=1e-14)
        assert r["fldj"] == pytest.approx(c["fldj"], rel=1e-12, abs=1e-14)
    known = ka["layer"] + ka["mdn"]
    assert len(ref["known_layers"]) == len(known)
    for r, c in zip(ref["known_layers"], known):
        assert r["log_prob"] == pytest.approx(c["log_prob"], rel=1e-12, abs=1e-14)
        if "flow_types" in r:  # get_total_param_size of the reference's layer (DistributionLayers.py:257-265)
            assert r["param_size"] == fo.chain_param_size(r["flow_types"], r["n_dims"], r["trainable_base_dist"])
        else:
            assert r["param_size"] == 2 * r["mdn_n_centers"] * r["n_dims"] + r["mdn_n_centers"]


def test_chain_oracles_match_reference_code():
    ref, cv = load("reference_run.json")["chains"], load("chain_vectors.json")
    n = 0
    for r, c in paired(ref, cv):
        ft, d, tb = c["flow_types"], c["n_dims"], c["trainable_base_dist"]
        # the reference builds its bijector list over the REVERSED flow list (DistributionLayers.py:269)
        assert r["bijector_order"] == [{"planar": "PlanarFlow", "radial": "RadialFlow", "affine": "AffineFlow"}[f]
                                       for f in reversed(ft)]
        t, y, up = t64(c["t"]), t64(c["y"]), t64(c["upstream"])
        logp, dt, dy = fo.with_grad(fo.chain_log_prob, t, y, ft, d, tb, upstream=up, want_dy=True)
        np.testing.assert_allclose(logp.numpy(), r["log_prob"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(dt.numpy(), np.asarray(r["dt"]).reshape(dt.shape), rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(dy.numpy(), r["dy"], rtol=1e-9, atol=1e-11)
        la, dta, dya = an.chain_forward_backward(np.asarray(c["t"], dtype=np.float32).reshape(len(c["y"]), -1),
                                                 np.asarray(c["y"], dtype=np.float32), ft, d, tb,
                                                 upstream=np.asarray(c["upstream"], dtype=np.float32))
        np.testing.assert_allclose(la, r["log_prob"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(dta, np.asarray(r["dt"]).reshape(dta.shape), rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(dya, r["dy"], rtol=1e-9, atol=1e-11)
        lb = fo.chain_log_prob(t, y[3:4], ft, d, tb)
        np.testing.assert_allclose(lb.numpy(), r["log_prob_y_row3_broadcast"], rtol=1e-12, atol=1e-12)
        n += 1
    assert n == 24


def test_mixture_oracles_match_reference_code():
    ref, mv = load("reference_run.json")["mixtures"], load("mixture_vectors.json")
    for r, c in paired(ref["mdn"], mv["mdn"]):
        logp, dt, dy = fo.with_grad(fo.mdn_log_prob, t64(c["t"]), t64(c["y"]), c["n_centers"], c["n_dims"],
                                    upstream=t64(c["upstream"]), want_dy=True)
        np.testing.assert_allclose(logp.numpy(), r["log_prob"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(dt.numpy(), r["dt"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(dy.numpy(), r["dy"], rtol=1e-9, atol=1e-12)
    for r, c in paired(ref["kmn"], mv["kmn"]):
        sv = t64(c["scale_vars"]).requires_grad_(True)
        scales = fo.kmn_scales(sv, c["n_centers"], c["init_scales"])
        # the reference's scale_model really returns negative bandwidths for init 0.3
        np.testing.assert_allclose(scales.detach().numpy(), r["scales"], rtol=1e-12, atol=1e-14)
        assert min(r["scales"]) < 0.0
        t, y = t64(c["t"]).requires_grad_(True), t64(c["y"]).requires_grad_(True)
        logp = fo.kmn_log_prob(t, y, t64(c["locs"]), scales)
        dt, dy, dsv = torch.autograd.grad(logp, [t, y, sv], grad_outputs=t64(c["upstream"]))
        np.testing.assert_allclose(logp.detach().numpy(), r["log_prob"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(dt.numpy(), r["dt"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(dy.numpy(), r["dy"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(dsv.numpy(), r["dscale_vars"], rtol=1e-9, atol=1e-12)


@pytest.mark.skipif(not HAVE_REFERENCE, reason="/root/reference does not travel to the GPU box")
def test_rerunning_the_reference_code_reproduces_the_fixture():
    from oracle import make_reference_run as mrr

    import sys

    assert mrr.diff(load("reference_run.json"), mrr.compute()) <= 1e-13
    assert "tensorflow" not in sys.modules and "estimators" not in sys.modules  # the stand-ins are gone again


@pytest.mark.skipif(not HAVE_REFERENCE, reason="/root/reference does not travel to the GPU box")
def test_reference_error_behaviour_under_the_shim():
    """The width assertions the facade mirrors (PlanarFlow.py:22, RadialFlow.py:23, AffineFlow.py:6,
    DistributionLayers.py:272) fire in the reference's own code."""
    from oracle import tf_shim

    FLOWS, DL = tf_shim.load_reference()
    try:
        assert DL.__file__ == "/root/reference/estimators/DistributionLayers.py"
        for name in ("planar", "radial", "affine"):
            P = FLOWS[name].get_param_size(2)
            with pytest.raises(AssertionError):
                FLOWS[name](torch.ones((3, P + 1), dtype=torch.float64), 2)
        layer = DL.InverseNormalizingFlowLayer(("planar", "radial"), 2, trainable_base_dist=True)
        with pytest.raises(AssertionError):
            layer(torch.ones((3, layer.get_total_param_size() - 1), dtype=torch.float64)).log_prob(
                torch.zeros((3, 2), dtype=torch.float64))
    finally:
        tf_shim.uninstall()


# ----------------------------------------------------------------------------- estimator level
def test_estimator_fixture_is_self_consistent():
    """reference_estimator_run.json (the reference's own BaseEstimator / MaximumLikelihoodNNEstimator /
    NFN / MDN / KMN classes, oracle/make_reference_estimator_run.py): score == -loss == mean log_pdf
    (reference tests/test_evaluation.py:29), pdf == exp(log_pdf), and the oracle reproduces log_pdf from
    the recorded network output."""
    fx = load("reference_estimator_run.json")
    assert "unmodified" in fx["provenance"]["what"]
    assert [c["cls"] for c in fx["cases"]] == ["NormalizingFlowNetwork", "NormalizingFlowNetwork",
                                               "MixtureDensityNetwork", "KernelMixtureNetwork"]
    for c in fx["cases"]:
        lp = np.asarray(c["log_pdf"])
        assert c["score"] == pytest.approx(-c["loss"], rel=1e-13)
        assert c["score"] == pytest.approx(lp.mean(), rel=1e-12)
        np.testing.assert_allclose(np.exp(lp), c["pdf"], rtol=1e-10)
        b, t = c["build"], t64(c["t"])
        y_mean, y_std = t64(c["stats"]["y_mean"]), t64(c["stats"]["y_std"])
        y_circ = fo.normalise_y(t64(c["y"]), y_mean, y_std)
        if c["cls"] == "NormalizingFlowNetwork":
            o = fo.chain_log_prob(t, y_circ, ["radial"] * b["n_flows"], b["n_dims"], b.get("trainable_base_dist", True))
        elif c["cls"] == "MixtureDensityNetwork":
            o = fo.mdn_log_prob(t, y_circ, b["n_centers"], b["n_dims"])
        else:
            o = fo.kmn_log_prob(t, y_circ, t64(c["locs"]), t64(c["scales"]))
        np.testing.assert_allclose(-fo.nll(o, y_std).numpy(), lp, rtol=1e-11, atol=1e-11)
    # rule-of-thumb noise level (BaseEstimator.py:36-39): 0.2 * (256 + 1) ** (-1 / (4 + 2))
    assert fx["cases"][3]["noise_std"][0] == pytest.approx(0.2 * 257 ** (-1 / 6), rel=1e-6)


def test_kmn_centre_selection_matches_reference_code():
    """GaussianKernelsLayer.set_center_points (host-side, once per fit) against the centres the reference's
    own set_center_points chose (DistributionLayers.py:135-171) with the same KMeans seeding."""
    from normalizingflownetwork_b200.DistributionLayers import GaussianKernelsLayer

    c = load("reference_estimator_run.json")["cases"][3]
    y = np.asarray(c["y"], np.float32)
    layer = GaussianKernelsLayer(c["build"]["n_centers"], c["build"]["n_dims"], trainable_scale=True)
    layer.set_center_points((y - np.mean(y, axis=0, dtype=np.float32)) / np.std(y, axis=0, dtype=np.float32))
    np.testing.assert_allclose(layer.locs.numpy(), np.asarray(c["locs"], np.float32), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(layer.scale_model().detach().numpy(), c["scales"], rtol=1e-6, atol=1e-7)


@pytest.mark.skipif(not HAVE_REFERENCE, reason="/root/reference does not travel to the GPU box")
def test_rerunning_the_reference_estimators_reproduces_the_fixture():
    from oracle import make_reference_estimator_run as mer
    from oracle.make_reference_run import diff

    assert diff(load("reference_estimator_run.json"), mer.compute()) <= 1e-12


def test_bayesian_fixture_matches_oracle_draw_loop():
    """The reference's BayesianNNEstimator.score (50 weight draws -> logsumexp - log 50 -> mean,
    BayesianNNEstimator.py:65-76) and MAP-mode log_pdf, re-derived with the oracle from the recorded
    posterior parameters and draws."""
    fx = load("reference_estimator_run.json")
    assert [c["name"] for c in fx["bayes_cases"]] == ["bayes_nfn_map", "bayes_nfn_50_draws", "bayes_mdn_50_draws",
                                                      "bayes_kmn_map"]
    # with the six classes of "cases" + "bayes_cases" every entry of the reference's ESTIMATORS dict has been run
    assert {c["cls"] for c in fx["cases"]} | {c["cls"] for c in fx["bayes_cases"]} == {
        "NormalizingFlowNetwork", "MixtureDensityNetwork", "KernelMixtureNetwork", "BayesNormalizingFlowNetwork",
        "BayesMixtureDensityNetwork", "BayesKernelMixtureNetwork"}
    for c in fx["bayes_cases"]:
        b = c["build"]
        x, y = t64(c["x"]), t64(c["y"])
        st = {k: t64(v) for k, v in c["stats"].items()}
        h0 = (x - st["x_mean"]) / t64(np.asarray(c["stats"]["x_std"], np.float32) + 1e-8)
        y_circ = fo.normalise_y(y, st["y_mean"], st["y_std"])
        S = c["posterior_draws"]
        assert S == (1 if b["map_mode"] else 50)
        scores = []
        for s in range(S):
            h = h0
            for li, v in enumerate(c["posterior_params"]):
                v = t64(v)
                size = v.shape[0] if b["map_mode"] else v.shape[0] // 2
                fan_in = h.shape[1]
                units = size // (fan_in + 1)
                if b["map_mode"]:
                    w = v
                else:  # MeanFieldLayer (DistributionLayers.py:42-56)
                    scale = 1e-3 + torch.nn.functional.softplus(fo.LOG_EXPM1_1 + 0.05 * v[size:])
                    w = v[:size] + scale * t64(c["eps"][li][s])
                h = h @ w[: fan_in * units].reshape(fan_in, units) + w[fan_in * units:]
                if li < len(c["posterior_params"]) - 1:
                    h = torch.tanh(h)
            if c["cls"] == "BayesNormalizingFlowNetwork":
                lp = fo.chain_log_prob(h, y_circ, ["radial"] * b["n_flows"], b["n_dims"], True)
            elif c["cls"] == "BayesMixtureDensityNetwork":
                lp = fo.mdn_log_prob(h, y_circ, b["n_centers"], b["n_dims"])
            else:
                lp = fo.kmn_log_prob(h, y_circ, t64(c["locs"]), t64(c["scales"]))
            scores.append(-fo.nll(lp, st["y_std"]))
        scores = torch.stack(scores)
        got = (torch.logsumexp(scores, 0) - np.log(S)).mean()
        assert float(got) == pytest.approx(c["score"], rel=1e-11)
        if b["map_mode"]:
            np.testing.assert_allclose(scores[0].numpy(), c["log_pdf"], rtol=1e-11, atol=1e-11)
            assert c["loss"] == pytest.approx(-c["score"] + sum(c["kl"]), rel=1e-12)


def test_reference_code_in_float32_sits_within_the_parity_bars():
    """reference_run_f32.json: the reference's own code in its working precision (float32) against its
    float64 run.  The GPU parity bars (1e-5 log-prob, 1e-4 gradients) are an order of magnitude above what
    the reference itself loses to rounding, so they do not hide formula differences
    (profiles/r01_accuracy_vs_reference_fp32.md has the CUDA columns)."""
    ref, f32 = load("reference_run.json"), load("reference_run_f32.json")
    worst_lp = worst_dt = 0.0
    for (r, h) in list(zip(ref["chains"], f32["chains"])) + list(zip(ref["mixtures"]["mdn"], f32["mdn"])):
        assert (r["name"], r["sigma"]) == (h["name"], h["sigma"])
        for key in ("log_prob", "dt"):
            a = np.asarray(h[key], dtype=np.float64)
            b = np.asarray(r[key], dtype=np.float64).reshape(a.shape)
            e = float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))
            if key == "log_prob":
                worst_lp = max(worst_lp, e)
            else:
                worst_dt = max(worst_dt, e)
    assert 1e-8 < worst_lp < 2e-6 and 1e-8 < worst_dt < 2e-5, (worst_lp, worst_dt)


@pytest.mark.skipif(not HAVE_REFERENCE, reason="/root/reference does not travel to the GPU box")
def test_reference_own_unit_tests_pass_on_the_stand_ins():
    """The reference's own test-suite, unmodified, against its own code on the TF / Keras / TFP stand-ins: all 23
    tests it does not mark slow, in all six files (flows, distribution layers, ML and Bayesian estimators incl. a
    minimal Keras fit loop, noise regularisation, scorers).  They pin behaviour, not values (shapes, parameter
    counts, bijector order, exception types, row independence, score == -evaluate, noise off at test time,
    MAP mode deterministic): evidence that the restated glue behaves as the reference expects."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""), PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run(
        [sys.executable, "-m", "pytest", "-p", "oracle.ref_pytest_plugin", "-p", "no:cacheprovider", "--rootdir", "/tmp",
         "-c", os.devnull, "-q", "-W", "ignore", "-m", "not slow", "/root/reference/tests"],
        cwd="/tmp", env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "23 passed, 6 deselected" in r.stdout, r.stdout[-500:]


@pytest.mark.skipif(not HAVE_REFERENCE, reason="/root/reference does not travel to the GPU box")
def test_oracles_equal_reference_code_on_random_chains_and_mixtures():
    """Beyond the frozen cases: 60 seeded random chain configurations (0-12 flows of random type, d = 1..6, both
    base-distribution modes, sigma 0.5 / 1 / 2, per-row and broadcast events) and 20 random MDN / KMN heads, the
    reference's own layer code (stand-ins, float64) against both oracles, values and gradients."""
    from oracle import tf_shim

    rng = np.random.default_rng(2222)
    FLOWS, DL = tf_shim.load_reference()
    try:
        names = sorted(FLOWS)
        for case in range(60):
            K, d, tb = int(rng.integers(0, 13)), int(rng.integers(1, 7)), bool(rng.integers(0, 2))
            ft = [names[i] for i in rng.integers(0, 3, size=K)]
            layer = DL.InverseNormalizingFlowLayer(ft, d, trainable_base_dist=tb)
            P = layer.get_total_param_size()
            assert P == fo.chain_param_size(ft, d, tb)
            if P == 0:
                continue
            B, sigma = int(rng.integers(1, 9)), float(rng.choice([0.5, 1.0, 2.0]))
            t = t64(rng.normal(0.0, sigma, size=(B, P))).requires_grad_(True)
            y = t64(rng.normal(0.0, 1.0, size=(B if case % 3 else 1, d))).requires_grad_(True)
            up = t64(rng.normal(size=(B,)))
            lp = layer(t).log_prob(y)
            dt, dy = torch.autograd.grad(lp, [t, y], grad_outputs=up)
            olp, odt, ody = fo.with_grad(fo.chain_log_prob, t.detach(), y.detach(), ft, d, tb, upstream=up, want_dy=True)
            tag = "case %d %s d=%d tb=%s" % (case, ft, d, tb)
            np.testing.assert_allclose(olp.numpy(), lp.detach().numpy(), rtol=1e-11, atol=1e-11, err_msg=tag)
            np.testing.assert_allclose(odt.numpy(), dt.numpy(), rtol=1e-8, atol=1e-10, err_msg=tag)
            np.testing.assert_allclose(ody.numpy(), dy.numpy(), rtol=1e-8, atol=1e-10, err_msg=tag)
            if y.shape[0] == B:
                alp, adt, ady = an.chain_forward_backward(t.detach().numpy(), y.detach().numpy(), ft, d, tb,
                                                          upstream=up.numpy())
                np.testing.assert_allclose(alp, lp.detach().numpy(), rtol=1e-11, atol=1e-11, err_msg=tag)
                np.testing.assert_allclose(adt, dt.numpy(), rtol=1e-8, atol=1e-10, err_msg=tag)
                np.testing.assert_allclose(ady, dy.numpy(), rtol=1e-8, atol=1e-10, err_msg=tag)
        for case in range(20):
            K, d, B = int(rng.integers(1, 9)), int(rng.integers(1, 5)), int(rng.integers(1, 7))
            y = t64(rng.normal(size=(B, d)))
            up = t64(rng.normal(size=(B,)))
            if case % 2:
                layer = DL.GaussianMixtureLayer(K, d)
                t = t64(rng.normal(0.0, 2.0, size=(B, layer.get_total_param_size()))).requires_grad_(True)
                lp = layer(t).log_prob(y)
                olp, odt = fo.with_grad(fo.mdn_log_prob, t.detach(), y, K, d, upstream=up)
            else:
                init = (0.3, 0.7, 1.5)[: int(rng.integers(1, 4))]
                layer = DL.GaussianKernelsLayer(K, d, trainable_scale=True, init_scales=init)
                sv = t64(rng.normal(0.0, 0.5, size=(len(init),)))
                layer.scale_model.layers[0].variable = sv
                locs = np.tile(rng.normal(size=(K, d)), (len(init), 1))
                layer.locs.assign(locs)
                layer.locs = layer.locs.value()[None]
                t = t64(rng.normal(0.0, 2.0, size=(B, K * len(init)))).requires_grad_(True)
                lp = layer(t).log_prob(y)
                scales = fo.kmn_scales(sv, K, init)
                olp, odt = fo.with_grad(lambda tt, yy: fo.kmn_log_prob(tt, yy, t64(locs), scales), t.detach(), y,
                                        upstream=up)
            (dt,) = torch.autograd.grad(lp, [t], grad_outputs=up)
            np.testing.assert_allclose(olp.numpy(), lp.detach().numpy(), rtol=1e-11, atol=1e-11)
            np.testing.assert_allclose(odt.numpy(), dt.numpy(), rtol=1e-8, atol=1e-11)
    finally:
        tf_shim.uninstall()
