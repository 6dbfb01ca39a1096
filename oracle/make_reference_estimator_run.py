"""Run the reference's OWN estimator classes (normalisation, NLL closure, pdf / log_pdf / score)
with fixed weights and freeze what they return as ``tests/golden/reference_estimator_run.json``.

TEST INFRASTRUCTURE ONLY.  Runs only where ``/root/reference`` exists (this container):

    python -m oracle.make_reference_estimator_run            # rewrite the fixture
    python -m oracle.make_reference_estimator_run --check    # recompute and diff

``estimators/{BaseEstimator,MaximumLikelihoodNNEstimator,NormalizingFlowNetwork,MixtureDensityNetwork,
KernelMixtureNetwork}.py`` are imported unmodified through ``oracle/tf_shim.py`` (torch-CPU float64
stand-ins for TF / Keras / TFP; nothing is trained here: weights and draws are set explicitly).  Per case the script builds the
reference estimator via its ``build_function``, calls the reference's ``_assign_data_normalization`` /
``_assign_noise_regularisation``, sets seeded Dense weights, and records
  log_pdf(x, y), pdf(x, y), score(x, y)                       (BaseEstimator.py:43-47, :71-86)
  mean of _get_neg_log_likelihood()(y, model.call(x))          (BaseEstimator.py:55-59, the Keras loss)
  its autograd gradients w.r.t. every Dense kernel / bias      (what one train step back-propagates)
and for the kernel mixture the centres chosen by the reference's ``set_center_points``
(DistributionLayers.py:135-171; ``KMeans(n_jobs=-2)`` no longer exists in scikit-learn, so the KMeans
the reference module sees is wrapped to drop ``n_jobs`` and to fix ``n_init=10, random_state=22``).
Each case is also evaluated with the oracle (``oracle/flow_oracle.py``) and must agree to 1e-11.
"""
import argparse
import importlib.util
import json
import os
import sys

import numpy as np
import torch

from oracle import flow_oracle as fo
from oracle import tf_shim
from oracle.make_reference_run import GOLDEN, diff, lst

OUT = os.path.join(GOLDEN, "reference_estimator_run.json")
F64 = torch.float64

CASES = [
    dict(name="nfn_d1_radial3_tanh", cls="NormalizingFlowNetwork", data="cosine", n=256,
         build=dict(n_dims=1, n_flows=3, hidden_sizes=(16, 16), activation="tanh")),
    dict(name="nfn_d2_radial4_relu_nobase", cls="NormalizingFlowNetwork", data="gauss2", n=192,
         build=dict(n_dims=2, n_flows=4, hidden_sizes=(8,), trainable_base_dist=False, activation="relu")),
    dict(name="mdn_d2_k5_relu", cls="MixtureDensityNetwork", data="gauss2", n=192,
         build=dict(n_dims=2, n_centers=5, hidden_sizes=(16, 16), activation="relu")),
    dict(name="kmn_d1_c10_tanh_rule_of_thumb", cls="KernelMixtureNetwork", data="cosine", n=256,
         build=dict(n_dims=1, n_centers=10, hidden_sizes=(16,), activation="tanh",
                    noise_reg=("rule_of_thumb", 0.2))),
]


def reference_cosine(n):
    spec = importlib.util.spec_from_file_location(
        "_ref_dummy_data_gen", os.path.join(tf_shim.REFERENCE_ROOT, "simulation", "dummy_data_gen.py"))
    mod = importlib.util.module_from_spec(spec)
    keep, sys.dont_write_bytecode = sys.dont_write_bytecode, True
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.dont_write_bytecode = keep
    return mod.gen_cosine_noise_data(n, noise_std=0.3, heterosced_noise=0.5)  # demo.py:14


def reference_scorers():
    """evaluation/scorers.py (NumPy / SciPy only), loaded by path so evaluation/__init__.py stays out."""
    import contextlib
    import io

    spec = importlib.util.spec_from_file_location(
        "_ref_scorers", os.path.join(tf_shim.REFERENCE_ROOT, "evaluation", "scorers.py"))
    mod = importlib.util.module_from_spec(spec)
    keep, sys.dont_write_bytecode = sys.dont_write_bytecode, True
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.dont_write_bytecode = keep

    def quiet(fn):  # bayesian_log_likelihood_score prints two shapes per draw (scorers.py:24-26)
        def call(model, x, y):
            with contextlib.redirect_stdout(io.StringIO()), torch.no_grad():
                return float(fn(mod.DummySklearWrapper(model), x, y))
        return call

    return quiet(mod.mle_log_likelihood_score), quiet(mod.bayesian_log_likelihood_score)


def reference_plot_model(model, x, y_range, y_num):
    """evaluation/visualization/flow_plotting.py:33-53 (plot_model), loaded by path with a recording
    ``matplotlib.pyplot`` stand-in: returns the density heat-map the reference hands to ``imshow``
    (one ``dist.prob(y_i)`` call per grid line, divided by ``sum(y_std)``)."""
    import types

    seen = {}
    plt = types.ModuleType("matplotlib.pyplot")
    plt.imshow = lambda img, **kw: seen.setdefault("heatmap", np.array(img, dtype=np.float64))
    for name in ("xlabel", "ylabel", "xticks", "yticks", "colorbar", "plot"):
        setattr(plt, name, lambda *a, **k: None)
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    keep = {k: sys.modules.get(k) for k in ("matplotlib", "matplotlib.pyplot")}
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt})
    keep_bc, sys.dont_write_bytecode = sys.dont_write_bytecode, True
    try:
        spec = importlib.util.spec_from_file_location(
            "_ref_flow_plotting", os.path.join(tf_shim.REFERENCE_ROOT, "evaluation", "visualization", "flow_plotting.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        with torch.no_grad():
            mod.plot_model(x, model, y_range, y_num=y_num)
    finally:
        sys.dont_write_bytecode = keep_bc
        for k, v in keep.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return seen["heatmap"]


def make_data(kind, n, rng):
    if kind == "cosine":
        return reference_cosine(n)
    # pdf / log_pdf assert x.shape == y.shape (BaseEstimator.py:72, :79): x is 2-D as well
    x = rng.normal(0.0, 1.5, size=(n, 2)).astype(np.float32)
    y = (np.stack([np.sin(x[:, 0]) + 0.3 * x[:, 1], 0.5 * x[:, 0] * x[:, 1]], 1)
         + rng.normal(0.0, 0.4, size=(n, 2))).astype(np.float32) * np.float32(2.5) + np.float32(1.0)
    return x, y


def oracle_log_prob(case, model, t, y_circ, DL):
    b = case["build"]
    if case["cls"] == "NormalizingFlowNetwork":
        return fo.chain_log_prob(t, y_circ, ["radial"] * b["n_flows"], b["n_dims"], b.get("trainable_base_dist", True))
    if case["cls"] == "MixtureDensityNetwork":
        return fo.mdn_log_prob(t, y_circ, b["n_centers"], b["n_dims"])
    layer = model.dist_layer
    scales = fo.kmn_scales(torch.zeros(layer.n_scales, dtype=F64), layer.n_centers, (0.3, 0.7))
    return fo.kmn_log_prob(t, y_circ, layer.locs[0].detach(), scales)


def run_case(case, mods, DL, rng):
    from sklearn.cluster import KMeans

    x, y = make_data(case["data"], case["n"], rng)
    model = getattr(mods[case["cls"]], case["cls"]).build_function(**case["build"])
    rec = {"name": case["name"], "cls": case["cls"], "build": {k: (list(v) if isinstance(v, tuple) else v)
                                                             for k, v in case["build"].items()},
           "x": x.tolist(), "y": y.tolist()}
    if case["cls"] == "KernelMixtureNetwork":
        # the first lines of KernelMixtureNetwork.fit (KernelMixtureNetwork.py:37-40), then BaseEstimator.fit's
        DL.KMeans = lambda n_clusters, n_jobs=None: KMeans(n_clusters=n_clusters, n_init=10, random_state=22)
        y_mean = np.mean(y, axis=0, dtype=np.float32)
        y_std = np.std(y, axis=0, dtype=np.float32)
        model.dist_layer.set_center_points((y - y_mean) / y_std)
        rec["locs"] = lst(model.dist_layer.locs[0])
        rec["scales"] = lst(model.dist_layer.scale_model(0.0))
    # what BaseEstimator.fit does before handing over to Keras (BaseEstimator.py:19-21)
    model._assign_data_normalization(x, y)
    model._assign_noise_regularisation(n_dims=x.shape[1] + y.shape[1], n_datapoints=x.shape[0])
    rec["noise_std"] = [float(model.x_noise_std.value()), float(model.y_noise_std.value())]
    rec["stats"] = {k: np.asarray(getattr(model, k), dtype=np.float64).tolist()
                    for k in ("x_mean", "x_std", "y_mean", "y_std")}
    tf_layers = sys.modules["tensorflow"].keras.layers
    dense = [l for l in model.layers if isinstance(l, tf_layers.Dense)]
    widths = [x.shape[1]] + [l.units for l in dense]
    weights = []
    for l, fan_in in zip(dense, widths):
        k = torch.tensor(rng.normal(0.0, 0.7 / np.sqrt(fan_in), size=(fan_in, l.units)).astype(np.float32),
                         dtype=F64, requires_grad=True)
        b = torch.tensor(rng.normal(0.0, 0.2, size=(l.units,)).astype(np.float32), dtype=F64, requires_grad=True)
        l.set_weights([k, b])
        weights += [k, b]
    assert dense[-1].units == model.layers[-1].get_total_param_size()
    rec["weights"] = [lst(w) for w in weights]  # kernel [in, units], bias, per Dense layer in order

    with torch.no_grad():
        log_pdf, pdf, score = model.log_pdf(x, y), model.pdf(x, y), model.score(x, y)
    nll = model._get_neg_log_likelihood()  # the loss handed to compile() (MaximumLikelihoodNNEstimator.py:33-35)
    per_sample = nll(y, model.call(x, training=False))
    loss = per_sample.mean()  # Keras reduces the per-sample loss with a mean
    grads = torch.autograd.grad(loss, weights)
    assert abs(float(score) + float(per_sample.detach().mean())) <= 1e-13 * max(1.0, abs(float(score)))

    # the oracle on the same weights
    # x_std + 1e-8 happens in NumPy float32 in the reference (MaximumLikelihoodNNEstimator.py:40): a no-op
    # for x_std ~ 1, which is also what the product's float32 torch expression evaluates to
    h = (torch.tensor(x, dtype=F64) - torch.tensor(np.asarray(model.x_mean), dtype=F64)) / \
        torch.tensor(np.asarray(model.x_std) + 1e-8, dtype=F64)
    ws = [w.detach() for w in weights]
    act = {"tanh": torch.tanh, "relu": torch.relu}[case["build"]["activation"]]
    for i in range(0, len(ws), 2):
        h = h @ ws[i] + ws[i + 1]
        if i + 2 < len(ws):
            h = act(h)
    y_std = torch.tensor(np.asarray(model.y_std), dtype=F64)
    y_circ = fo.normalise_y(torch.tensor(y, dtype=F64), torch.tensor(np.asarray(model.y_mean), dtype=F64), y_std)
    o_lp = oracle_log_prob(case, model, h, y_circ, DL)
    o_log_pdf = o_lp - torch.sum(torch.log(y_std))
    assert torch.allclose(o_log_pdf, log_pdf, rtol=1e-11, atol=1e-11), case["name"]
    assert torch.allclose(fo.nll(o_lp, y_std), per_sample.detach(), rtol=1e-11, atol=1e-11), case["name"]
    assert torch.allclose(torch.exp(o_log_pdf), pdf, rtol=1e-10, atol=1e-300), case["name"]

    scorer = reference_scorers()[0](model, x, y)  # evaluation/scorers.py:30-34
    assert abs(scorer - float(score)) <= 1e-12 * max(1.0, abs(scorer))
    if case["data"] == "cosine" and case["cls"] == "NormalizingFlowNetwork":
        # the density grid of plot_model (x must be increasing: the cosine x is a linspace)
        rec["plot_y_range"], rec["plot_y_num"] = [-6.0, 6.0], 16
        heat = reference_plot_model(model, x, rec["plot_y_range"], rec["plot_y_num"])
        assert heat.shape == (16, len(x))
        # row i is the density of y_i = linspace(hi, lo)[i] at every x: cross-check one line against pdf()
        y_line = np.full_like(y, np.linspace(6.0, -6.0, 16)[5])
        with torch.no_grad():
            assert np.allclose(heat[5], model.pdf(x, y_line).numpy(), rtol=1e-12, atol=1e-300)
        rec["plot_heatmap"] = heat.tolist()
    rec.update(log_pdf=lst(log_pdf), pdf=lst(pdf), score=float(score), loss=float(loss.detach()),
               grads=[lst(g) for g in grads], t=lst(h), scorer_mle=scorer)
    return rec


BAYES_CASES = [
    # MAP mode: the posterior mean is used, one "draw" (BayesianNNEstimator.py:71), fully deterministic
    dict(name="bayes_nfn_map", n=128,
         build=dict(n_dims=1, kl_weight_scale=1.0 / 128, n_flows=2, hidden_sizes=(10,), activation="tanh",
                    map_mode=True, prior_scale=1.0)),
    # 50 posterior draws, logsumexp over them (BayesianNNEstimator.py:65-76); the draws are the harness's
    dict(name="bayes_nfn_50_draws", n=128,
         build=dict(n_dims=1, kl_weight_scale=1.0 / 128, n_flows=2, hidden_sizes=(10,), activation="tanh",
                    map_mode=False, prior_scale=0.7)),
    dict(name="bayes_mdn_50_draws", cls="BayesMixtureDensityNetwork", data="gauss2", n=96,
         build=dict(n_dims=2, kl_weight_scale=1.0 / 96, n_centers=3, hidden_sizes=(10,), activation="tanh",
                    map_mode=False, prior_scale=1.0)),
    dict(name="bayes_kmn_map", cls="BayesKernelMixtureNetwork", data="cosine", n=128,
         build=dict(n_dims=1, kl_weight_scale=1.0 / 128, n_centers=6, hidden_sizes=(10,), activation="tanh",
                    map_mode=True, prior_scale=1.0)),
]


def run_bayes_case(case, mods, DL, rng):
    from sklearn.cluster import KMeans

    x, y = make_data(case.get("data", "cosine"), case["n"], rng)
    b = case["build"]
    cls = case.get("cls", "BayesNormalizingFlowNetwork")
    model = getattr(mods[cls], cls).build_function(**b)
    kmn = {}
    if cls == "BayesKernelMixtureNetwork":  # BayesKernelMixtureNetwork.py:47-50
        DL.KMeans = lambda n_clusters, n_jobs=None: KMeans(n_clusters=n_clusters, n_init=10, random_state=22)
        model.dist_layer.set_center_points((y - np.mean(y, axis=0, dtype=np.float32)) / np.std(y, axis=0, dtype=np.float32))
        kmn = {"locs": lst(model.dist_layer.locs[0]), "scales": lst(model.dist_layer.scale_model(0.0))}
    model._assign_data_normalization(x, y)
    model._assign_noise_regularisation(n_dims=x.shape[1] + y.shape[1], n_datapoints=x.shape[0])
    tfp_layers = sys.modules["tensorflow_probability"].layers
    dv = [l for l in model.layers if isinstance(l, tfp_layers.DenseVariational)]
    widths = [x.shape[1]] + [l.units for l in dv]
    post = []
    for l, fan_in in zip(dv, widths):
        l.build(fan_in)
        size = fan_in * l.units + l.units
        n_var = size if b["map_mode"] else 2 * size
        assert tuple(l._posterior.layers[0].variable.shape) == (n_var,)
        v = torch.tensor(rng.normal(0.0, 0.5, size=(n_var,)).astype(np.float32), dtype=F64, requires_grad=True)
        l._posterior.layers[0].variable = v  # the VariableLayer weight (BayesianNNEstimator.py:100-105)
        post.append(v)
    draws = 1 if b["map_mode"] else 50
    rec = {"name": case["name"], "cls": cls, **kmn,
           "build": {k: (list(v) if isinstance(v, tuple) else v) for k, v in b.items()},
           "x": x.tolist(), "y": y.tolist(), "posterior_params": [lst(v) for v in post], "posterior_draws": draws,
           "stats": {k: np.asarray(getattr(model, k), dtype=np.float64).tolist()
                     for k in ("x_mean", "x_std", "y_mean", "y_std")}}
    if not b["map_mode"]:
        # draw s consumes one standard-normal vector per DenseVariational layer, in layer order
        eps = [rng.normal(size=(draws, v.shape[0] // 2)).astype(np.float32) for v in post]
        tf_shim.push_draws([eps[l][s] for s in range(draws) for l in range(len(dv))])
        rec["eps"] = [e.tolist() for e in eps]
    with torch.no_grad():
        rec["score"] = float(model.score(x, y))  # the reference's own draw loop + logsumexp - log S
    if not b["map_mode"]:
        tf_shim.push_draws([eps[l][s] for s in range(draws) for l in range(len(dv))])  # same draws again
    rec["scorer_bayes"] = reference_scorers()[1](model, x, y)  # evaluation/scorers.py:13-27 (scipy logsumexp)
    assert abs(rec["scorer_bayes"] - rec["score"]) <= 1e-12 * max(1.0, abs(rec["score"]))
    if b["map_mode"]:
        with torch.no_grad():
            rec["log_pdf"], rec["pdf"] = lst(model.log_pdf(x, y)), lst(model.pdf(x, y))
    else:
        tf_shim.push_draws([np.zeros(v.shape[0] // 2) for v in post])  # eps = 0: the KL below does not depend on it
    nll = model._get_neg_log_likelihood()
    per_sample = nll(y, model.call(x, training=False))
    kls = model.losses  # what DenseVariational added through add_loss; Keras adds them to the compiled loss
    assert len(kls) == len(dv)
    rec["kl"] = [float(k.detach()) for k in kls]
    if b["map_mode"]:
        loss = per_sample.mean() + sum(kls)
        rec["loss"] = float(loss.detach())
        rec["grads"] = [lst(g) for g in torch.autograd.grad(loss, post)]
        assert abs(rec["score"] + float(per_sample.detach().mean())) <= 1e-12
    # independent restatement of the exact Normal-Normal KL (DistributionLayers.py:42-68 parameterisation)
    for v, k, l in zip(post, rec["kl"], dv):
        v = v.detach()
        size = v.shape[0] if b["map_mode"] else v.shape[0] // 2
        loc = v[:size]
        sq = torch.ones(size, dtype=F64) if b["map_mode"] else \
            1e-3 + torch.nn.functional.softplus(fo.LOG_EXPM1_1 + 0.05 * v[size:])
        sr = b["prior_scale"]
        want = b["kl_weight_scale"] * float(torch.sum(torch.log(sr / sq) + (sq ** 2 + loc ** 2) / (2 * sr ** 2) - 0.5))
        assert abs(want - k) <= 1e-12 * max(1.0, abs(want)), (case["name"], want, k)
    assert not tf_shim._EPS_QUEUE, "queued draws left over"
    return rec


def compute():
    FLOWS, DL = tf_shim.load_reference()
    try:
        mods = {c: tf_shim.load_reference_module("estimators." + c)
                for c in ("NormalizingFlowNetwork", "MixtureDensityNetwork", "KernelMixtureNetwork",
                          "BayesNormalizingFlowNetwork", "BayesMixtureDensityNetwork", "BayesKernelMixtureNetwork")}
        rng = np.random.default_rng(22)
        return {
            "provenance": {
                "what": "outputs of the reference's own estimators/{BaseEstimator,MaximumLikelihoodNNEstimator,"
                        "NormalizingFlowNetwork,MixtureDensityNetwork,KernelMixtureNetwork}.py, imported unmodified "
                        "from /root/reference and executed on torch-CPU float64 stand-ins for TF / Keras / TFP "
                        "(oracle/tf_shim.py), with seeded Dense weights",
                "generator": "python -m oracle.make_reference_estimator_run",
                "weights_layout": "per Dense layer: kernel [in, units] then bias [units] (Keras layout)",
                "does_not_pin": "TFP / Keras own arithmetic (restated in the shim), float32 rounding, the Keras "
                                "training loop and optimizer; KMeans seeding differs from the reference's unseeded call",
            },
            "cases": [run_case(c, mods, DL, rng) for c in CASES],
            "bayes_cases": [run_bayes_case(c, mods, DL, rng) for c in BAYES_CASES],
        }
    finally:
        tf_shim.uninstall()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    got = compute()
    if args.check:
        with open(OUT) as f:
            want = json.load(f)
        m = diff(want, got)
        print(f"max |fixture - rerun| = {m:.3e}")
        sys.exit(0 if m <= 1e-12 else 1)
    with open(OUT, "w") as f:
        json.dump(got, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", [c["name"] for c in got["cases"]])
    for c in got["cases"]:
        print("  %-34s score %.6f loss %.6f" % (c["name"], c["score"], c["loss"]))
    for c in got["bayes_cases"]:
        print("  %-34s score %.6f kl %s" % (c["name"], c["score"], c["kl"]))


if __name__ == "__main__":
    main()
