"""GPU: the estimator facade (reference API: fit / log_pdf / pdf / score, Bayesian variant
with S draws folded into the batch) on top of the fused kernels."""
import numpy as np
import pytest
import torch

from oracle import flow_oracle as fo

pytestmark = pytest.mark.gpu


def _cosine(n):
    from normalizingflownetwork_b200.simulation import gen_cosine_noise_data

    return gen_cosine_noise_data(n, noise_std=0.3, heterosced_noise=0.5)


def _oracle_log_pdf(model, x, y):
    """float64 oracle on the model's own network output: isolates the head + normalisation."""
    with torch.no_grad():
        t = model.params_from_x(x).double().cpu()
    y64 = (torch.tensor(y, dtype=torch.float64) - model.y_mean.double().cpu()) / model.y_std.double().cpu()
    layer = model.dist_layer
    lp = fo.chain_log_prob(t, y64, list(layer._flow_types), layer._n_dims, layer._trainable_base_dist)
    return (lp - torch.sum(torch.log(model.y_std.double().cpu()))).numpy()


def test_dense_layer_generation(cuda_device):
    from normalizingflownetwork_b200.estimators import BayesNormalizingFlowNetwork, NormalizingFlowNetwork

    layers = NormalizingFlowNetwork(1)._get_dense_layers(hidden_sizes=(2, 2, 2), output_size=2, activation="linear")
    assert len(layers) == 6
    layers = BayesNormalizingFlowNetwork(1, 1.0)._get_dense_layers(hidden_sizes=(2, 2, 2), output_size=2,
                                                                 posterior=None, prior=None)
    assert len(layers) == 6


@pytest.mark.parametrize("name,kwargs", [("NFN", dict(n_flows=3)), ("MDN", dict(n_centers=5)), ("KMN", dict(n_centers=5))])
@pytest.mark.parametrize("d", [1, 3])
def test_ml_model_output_dims(cuda_device, name, kwargs, d):
    from normalizingflownetwork_b200.estimators import ESTIMATORS

    x_train = np.linspace([[-1]] * d, [[1]] * d, 10).reshape((10, d))
    y_train = np.linspace([[-1]] * d, [[1]] * d, 10).reshape((10, d)) + 0.01 * np.random.default_rng(0).normal(size=(10, d))
    model = ESTIMATORS[name](d, **kwargs)
    model.fit(x_train, y_train, epochs=1, verbose=0)
    output = model(x_train)
    assert output.event_shape == [d]
    assert output.batch_shape == [10]
    lp = output.log_prob([[0.0] * d])
    assert tuple(lp.shape) == (10,) and torch.isfinite(lp).all()
    assert tuple(model.pdf(x_train, y_train).shape) == (10,)


def test_nfn_log_pdf_matches_oracle_and_training_improves(cuda_device):
    from normalizingflownetwork_b200.estimators import NormalizingFlowNetwork

    x, y = _cosine(2048)
    model = NormalizingFlowNetwork.build_function(n_dims=1, n_flows=3, hidden_sizes=(16, 16), activation="tanh")
    model.fit(x, y, batch_size=256, epochs=2, verbose=0)
    first = model.history[0]
    model.fit(x, y, batch_size=256, epochs=25, verbose=0)
    assert model.history[-1] < first - 0.1, model.history
    got = model.log_pdf(x, y).cpu().numpy()
    ref = _oracle_log_pdf(model, x, y)
    assert np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref))) <= 1e-5
    np.testing.assert_allclose(model.pdf(x, y).cpu().numpy(), np.exp(ref), rtol=2e-5, atol=1e-7)
    # score == -evaluate (reference tests/test_evaluation.py:29)
    assert model.score(x, y) == pytest.approx(-model.evaluate(x, y), rel=1e-5)
    assert model.score(x, y) == pytest.approx(float(np.mean(ref)), rel=1e-5)
    # noise is off at test time: pdf is deterministic (reference tests/test_noise_reg.py)
    assert torch.equal(model.pdf(x, y), model.pdf(x, y))


def test_train_step_gradients_match_float64_autograd(cuda_device):
    """One fused train step vs float64 torch autograd through the oracle on a CPU copy of the MLP."""
    from normalizingflownetwork_b200.estimators import NormalizingFlowNetwork

    x, y = _cosine(512)
    model = NormalizingFlowNetwork(1, n_flows=3, hidden_sizes=(16, 16), activation="tanh", learning_rate=0.0)
    model.fit(x, y, batch_size=512, epochs=1, verbose=0, shuffle=False)  # lr = 0: weights unchanged
    lin = [m.linear for m in model.net if hasattr(m, "linear")]
    W = [(l.weight.detach().double().cpu().requires_grad_(True), l.bias.detach().double().cpu().requires_grad_(True))
         for l in lin]
    xm, xs = model.x_mean.double().cpu(), model.x_std.double().cpu()
    h = (torch.tensor(x, dtype=torch.float64) - xm) / (xs + 1e-8)
    for i, (w, b) in enumerate(W):
        h = h @ w.T + b
        if i < len(W) - 1:
            h = torch.tanh(h)
    y64 = (torch.tensor(y, dtype=torch.float64) - model.y_mean.double().cpu()) / model.y_std.double().cpu()
    nll = -fo.chain_log_prob(h, y64, ["radial"] * 3, 1, True).mean()
    nll.backward()
    model.optimizer.zero_grad(set_to_none=True)
    loss = model.train_step(model._to_dev(x), model._to_dev(y))
    assert float(loss) == pytest.approx(float(nll) + float(torch.sum(torch.log(model.y_std.double().cpu()))), rel=1e-5)
    for l, (w, b) in zip(lin, W):
        gw, gb = l.weight.grad.double().cpu(), l.bias.grad.double().cpu()
        scale = max(1e-3, float(w.grad.abs().max()))
        assert float((gw - w.grad).abs().max()) <= 1e-4 * scale
        assert float((gb - b.grad).abs().max()) <= 1e-4 * max(1e-3, float(b.grad.abs().max()))


def test_mdn_and_kmn_train(cuda_device):
    from normalizingflownetwork_b200.estimators import KernelMixtureNetwork, MixtureDensityNetwork

    x, y = _cosine(1024)
    for model in (MixtureDensityNetwork(1, n_centers=5, activation="tanh"),
                  KernelMixtureNetwork(1, n_centers=10, activation="tanh")):
        model.fit(x, y, batch_size=128, epochs=12, verbose=0)
        assert np.isfinite(model.history).all()
        assert model.history[-1] < model.history[0]
        assert np.isfinite(model.score(x, y))
    assert model.dist_layer.scale_vars.grad is not None  # KMN bandwidths are trained through dscales


def test_bayesian_nfn(cuda_device):
    from normalizingflownetwork_b200.estimators import BayesNormalizingFlowNetwork
    from normalizingflownetwork_b200.evaluation.scorers import DummySklearWrapper, bayesian_log_likelihood_score

    x, y = _cosine(512)
    for n_flows, tb in [(3, False), (0, True), (10, True)]:
        m = BayesNormalizingFlowNetwork(1, kl_weight_scale=1.0 / x.shape[0], n_flows=n_flows, hidden_sizes=(16, 16),
                                        trainable_base_dist=tb)
        m.fit(x, y, batch_size=64, epochs=2, verbose=0)
        out = m(x[:10])
        assert out.event_shape == [1] and out.batch_shape == [10]
        assert tuple(out.log_prob([[0.0]]).shape) == (10,)
        assert tuple(m.pdf(x[:10], y[:10]).shape) == (10,)
        assert np.isfinite(m.history).all()
    # S draws folded into the batch == logsumexp over the draws - log S
    S, B = 8, 100
    m.train(False)
    with torch.no_grad():
        m._wgen = torch.Generator(device=cuda_device).manual_seed(5)
        t = m.params_from_x_draws(x[:B], S)
        assert tuple(t.shape) == (S * B, m.dist_layer.get_total_param_size())
        yy = m._y_input(y[:B], training=False)
        logp = (m.dist_layer(t).log_prob(yy.repeat(S, 1)) - m._log_ystd_sum()).view(S, B)
        ref = torch.logsumexp(logp.double(), 0) - np.log(S)
        m._wgen = torch.Generator(device=cuda_device).manual_seed(5)
        got = m.log_posterior_predictive(x[:B], y[:B], posterior_draws=S)
    np.testing.assert_allclose(got.cpu().numpy(), ref.cpu().numpy(), rtol=1e-5, atol=1e-5)
    # Bayes: stochastic between calls; MAP: deterministic (reference tests/test_bayesian_estimator.py:83-135)
    assert m.score(x, y, posterior_draws=4) != m.score(x, y, posterior_draws=4)
    mm = BayesNormalizingFlowNetwork(1, kl_weight_scale=1.0 / x.shape[0], n_flows=2, map_mode=True)
    mm.fit(x, y, batch_size=64, epochs=1, verbose=0)
    assert mm.score(x, y) == mm.score(x, y)
    assert np.isfinite(bayesian_log_likelihood_score(DummySklearWrapper(m), x, y))


def test_cuda_graph_train_step_matches_eager(cuda_device):
    """fit(cuda_graph=True) replays one captured graph per mini-batch and learns like the eager loop."""
    from normalizingflownetwork_b200.estimators import NormalizingFlowNetwork

    x, y = _cosine(2048)
    eager = NormalizingFlowNetwork.build_function(n_dims=1, n_flows=3, hidden_sizes=(16, 16), activation="tanh")
    graphed = NormalizingFlowNetwork.build_function(n_dims=1, n_flows=3, hidden_sizes=(16, 16), activation="tanh")
    eager.fit(x, y, batch_size=512, epochs=12, verbose=0)
    graphed.fit(x, y, batch_size=512, epochs=12, verbose=0, cuda_graph=True)
    assert np.isfinite(graphed.history).all()
    assert graphed.history[-1] < graphed.history[0] - 0.05
    # same seed, same data order, noise off: the two loops follow the same trajectory
    np.testing.assert_allclose(graphed.history, eager.history, rtol=2e-3, atol=2e-3)


def test_bayesian_training_with_folded_draws(cuda_device):
    """BASELINE config 4: S Monte-Carlo weight draws folded into the batch of one training step."""
    from normalizingflownetwork_b200.estimators import BayesNormalizingFlowNetwork

    x, y = _cosine(1024)
    m = BayesNormalizingFlowNetwork(1, kl_weight_scale=1.0 / x.shape[0], n_flows=5, hidden_sizes=(16, 16),
                                    n_train_draws=8)
    m.fit(x, y, batch_size=256, epochs=15, verbose=0)
    assert np.isfinite(m.history).all()
    assert m.history[-1] < m.history[0]
    # the S-draw step averages the NLL over draws: its loss is close to the 1-draw loss in expectation
    m1 = BayesNormalizingFlowNetwork(1, kl_weight_scale=1.0 / x.shape[0], n_flows=5, hidden_sizes=(16, 16))
    m1.fit(x, y, batch_size=256, epochs=15, verbose=0)
    assert abs(m.history[-1] - m1.history[-1]) < 0.2 * abs(m1.history[-1])


def test_pdf_grid_matches_per_line_pdf(cuda_device):
    """plot_model (evaluation/visualization/flow_plotting.py:33-53) scores one grid line per call;
    pdf_grid does the whole heat-map in one launch and must agree line by line."""
    from normalizingflownetwork_b200.estimators import NormalizingFlowNetwork

    x, y = _cosine(512)
    model = NormalizingFlowNetwork.build_function(n_dims=1, n_flows=3, hidden_sizes=(16, 16), activation="tanh")
    model.fit(x, y, batch_size=128, epochs=2, verbose=0)
    xs = np.linspace(-3, 3, 37, dtype=np.float32).reshape(-1, 1)
    ys = np.linspace(-4, 4, 21, dtype=np.float32).reshape(-1, 1)
    grid = model.pdf_grid(xs, ys).cpu().numpy()
    assert grid.shape == (21, 37)
    for j in range(21):
        line = model.pdf(xs, np.full_like(xs, ys[j, 0])).cpu().numpy()
        np.testing.assert_allclose(grid[j], line, rtol=2e-5, atol=1e-7)


def test_cuda_graph_log_pdf_matches_eager(cuda_device):
    """Graph-captured scoring (config 1 is launch-latency-bound) returns exactly the eager result and
    follows in-place weight updates."""
    from normalizingflownetwork_b200.estimators import NormalizingFlowNetwork

    x, y = _cosine(2048)
    model = NormalizingFlowNetwork.build_function(n_dims=1, n_flows=3, hidden_sizes=(16, 16), activation="tanh")
    model.fit(x, y, batch_size=512, epochs=2, verbose=0)
    xd, yd = model._to_dev(x), model._to_dev(y)
    model.capture_log_pdf(2048, 1, 1)
    assert torch.equal(model.log_pdf_graphed(xd, yd), model.log_pdf(xd, yd))
    model.fit(x, y, batch_size=512, epochs=1, verbose=0)   # in-place updates: the graph sees the new weights
    assert torch.equal(model.log_pdf_graphed(xd, yd), model.log_pdf(xd, yd))
    perm = torch.randperm(2048, device=xd.device)
    assert torch.equal(model.log_pdf_graphed(xd[perm], yd[perm]), model.log_pdf(xd[perm], yd[perm]))


# ----------------------------------------------------------------------------- the reference's own estimator code
def _reference_estimator_cases():
    import json
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_estimator_run.json")
    with open(path) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("idx", range(4))
def test_estimators_match_reference_code_run(cuda_device, idx):
    """log_pdf / pdf / score and the train-step gradients against what the reference's OWN
    BaseEstimator / MaximumLikelihoodNNEstimator / {NFN, MDN, KMN} classes returned for the same data and
    weights (tests/golden/reference_estimator_run.json, oracle/make_reference_estimator_run.py)."""
    from normalizingflownetwork_b200 import estimators as E

    c = _reference_estimator_cases()[idx]
    x, y = np.asarray(c["x"], np.float32), np.asarray(c["y"], np.float32)
    build = dict(c["build"])
    for k in ("hidden_sizes", "noise_reg"):
        if k in build:
            build[k] = tuple(build[k])
    model = getattr(E, c["cls"]).build_function(learning_rate=0.0, **build)
    try:
        # lr = 0: fit only assigns the normalisation / noise / centres and materialises the lazy weights
        model.fit(x, y, batch_size=len(x), epochs=1, verbose=0, shuffle=False)
        for k, v in c["stats"].items():
            np.testing.assert_allclose(getattr(model, k).cpu().numpy(), np.asarray(v, np.float32), rtol=1e-6, atol=1e-7)
        assert float(model.x_noise_std) == pytest.approx(c["noise_std"][0], rel=1e-6, abs=1e-12)
        assert float(model.y_noise_std) == pytest.approx(c["noise_std"][1], rel=1e-6, abs=1e-12)
        lin = [m.linear for m in model.net if hasattr(m, "linear")]
        assert 2 * len(lin) == len(c["weights"])
        with torch.no_grad():
            for i, l in enumerate(lin):
                kernel = torch.tensor(c["weights"][2 * i], dtype=torch.float32)  # Keras layout [in, units]
                l.weight.copy_(kernel.t().to(l.weight.device))
                l.bias.copy_(torch.tensor(c["weights"][2 * i + 1], dtype=torch.float32).to(l.bias.device))
            if "locs" in c:
                locs = torch.tensor(c["locs"], dtype=torch.float32, device=model.dist_layer.locs.device)
                # the product's own centre selection picked the same centres as the reference's code
                np.testing.assert_allclose(model.dist_layer.locs.cpu().numpy(), locs.cpu().numpy(), rtol=1e-4, atol=1e-5)
                model.dist_layer.locs.copy_(locs)
                np.testing.assert_allclose(model.dist_layer.scale_model().detach().cpu().numpy(), c["scales"],
                                           rtol=1e-6, atol=1e-7)

        ref_lp = np.asarray(c["log_pdf"])
        got = model.log_pdf(x, y).cpu().numpy()
        assert np.max(np.abs(got - ref_lp) / np.maximum(1.0, np.abs(ref_lp))) <= 2e-5, c["name"]
        np.testing.assert_allclose(model.pdf(x, y).cpu().numpy(), c["pdf"], rtol=1e-4, atol=1e-30)
        assert model.score(x, y) == pytest.approx(c["score"], rel=2e-5)
        assert model.evaluate(x, y) == pytest.approx(c["loss"], rel=2e-5)
        from normalizingflownetwork_b200.evaluation import scorers  # reference evaluation/scorers.py:30-34

        assert scorers.mle_log_likelihood_score(scorers.DummySklearWrapper(model), x, y) == \
            pytest.approx(c["scorer_mle"], rel=2e-5)

        if "plot_heatmap" in c:
            # the density grid the reference's plot_model loop hands to imshow (flow_plotting.py:33-53), one launch here
            lo, hi = c["plot_y_range"]
            grid = model.pdf_grid(x, np.linspace(hi, lo, num=c["plot_y_num"]))
            heat = np.asarray(c["plot_heatmap"])
            assert tuple(grid.shape) == heat.shape
            np.testing.assert_allclose(grid.cpu().numpy(), heat, rtol=1e-4, atol=1e-7)

        model._set_noise(0.0)  # training-time noise draws are RNG specific; the reference run had none either
        loss = model.train_step(model._to_dev(x), model._to_dev(y))
        assert float(loss) == pytest.approx(c["loss"], rel=2e-5)
        for i, l in enumerate(lin):
            gk = np.asarray(c["grads"][2 * i]).T  # -> torch layout [units, in]
            gb = np.asarray(c["grads"][2 * i + 1])
            for got_g, ref_g, what in ((l.weight.grad, gk, "kernel"), (l.bias.grad, gb, "bias")):
                scale = max(1e-3, float(np.abs(ref_g).max()))
                err = float(np.abs(got_g.double().cpu().numpy() - ref_g).max())
                assert err <= 2e-4 * scale, "%s layer %d %s grad err %.3e (scale %.3e)" % (c["name"], i, what, err, scale)
    finally:
        model._set_noise(0.0)


@pytest.mark.parametrize("idx", range(4))
def test_bayesian_estimator_matches_reference_code_run(cuda_device, idx, monkeypatch):
    """BayesNormalizingFlowNetwork against the reference's OWN BayesianNNEstimator code
    (tests/golden/reference_estimator_run.json "bayes_cases"): MAP mode (deterministic: log_pdf, pdf, score,
    loss incl. the exact KL terms, posterior-parameter gradients) and the 50-draw posterior-predictive score
    with the SAME weight draws (the fixture records them; torch.randn is patched to hand them out)."""
    import json
    import os

    from normalizingflownetwork_b200 import estimators as E
    from normalizingflownetwork_b200.estimators.BayesianNNEstimator import DenseVariational

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_estimator_run.json")
    with open(path) as f:
        c = json.load(f)["bayes_cases"][idx]
    x, y = np.asarray(c["x"], np.float32), np.asarray(c["y"], np.float32)
    build = dict(c["build"])
    build["hidden_sizes"] = tuple(build["hidden_sizes"])
    model = getattr(E, c["cls"]).build_function(learning_rate=0.0, **build)
    model.fit(x, y, batch_size=len(x), epochs=1, verbose=0, shuffle=False)  # lr = 0: set-up only
    if "locs" in c:
        locs = torch.tensor(c["locs"], dtype=torch.float32, device=model.dist_layer.locs.device)
        np.testing.assert_allclose(model.dist_layer.locs.cpu().numpy(), locs.cpu().numpy(), rtol=1e-4, atol=1e-5)
        with torch.no_grad():
            model.dist_layer.locs.copy_(locs)
    layers = [l for l in model.net if isinstance(l, DenseVariational)]
    assert len(layers) == len(c["posterior_params"])
    with torch.no_grad():
        for l, v in zip(layers, c["posterior_params"]):
            assert l.posterior_params.numel() == len(v)
            l.posterior_params.copy_(torch.tensor(v, dtype=torch.float32).to(l.posterior_params.device))

    if build["map_mode"]:
        ref_lp = np.asarray(c["log_pdf"])
        got = model.log_pdf(x, y).cpu().numpy()
        assert np.max(np.abs(got - ref_lp) / np.maximum(1.0, np.abs(ref_lp))) <= 2e-5
        np.testing.assert_allclose(model.pdf(x, y).cpu().numpy(), c["pdf"], rtol=1e-4, atol=1e-30)
        assert model.score(x, y) == pytest.approx(c["score"], rel=2e-5)
        assert model.evaluate(x, y) == pytest.approx(c["loss"], rel=2e-5)
        from normalizingflownetwork_b200.evaluation import scorers  # reference evaluation/scorers.py:13-27

        assert scorers.bayesian_log_likelihood_score(scorers.DummySklearWrapper(model), x, y) == \
            pytest.approx(c["scorer_bayes"], rel=2e-5)
        for l, kl in zip(layers, c["kl"]):
            assert float(l.last_kl) == pytest.approx(kl, rel=1e-5)
        loss = model.train_step(model._to_dev(x), model._to_dev(y))
        assert float(loss) == pytest.approx(c["loss"], rel=2e-5)
        for l, g in zip(layers, c["grads"]):
            g = np.asarray(g)
            err = float(np.abs(l.posterior_params.grad.double().cpu().numpy() - g).max())
            assert err <= 2e-4 * max(1e-3, float(np.abs(g).max())), err
    else:
        queue = [torch.tensor(e, dtype=torch.float32) for e in c["eps"]]  # per layer: [draws, size]

        def handed_out(*shape, device=None, generator=None, **kw):
            shape = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
            e = queue.pop(0)
            assert tuple(e.shape) == shape, (e.shape, shape)
            return e.to(device)

        monkeypatch.setattr(torch, "randn", handed_out)
        try:
            score = model.score(x, y)
        finally:
            monkeypatch.undo()
        assert not queue
        assert score == pytest.approx(c["score"], rel=2e-5)
        for l, kl in zip(layers, c["kl"]):
            assert float(l.last_kl) == pytest.approx(kl, rel=1e-5)


def test_y_noise_reg_like_the_reference(cuda_device):
    """reference tests/test_noise_reg.py:55-75: the y input model is deterministic in evaluation and noisy in
    training; pdf is deterministic after a fit with noise (tests/test_noise_reg.py:31-34)."""
    from normalizingflownetwork_b200.estimators import NormalizingFlowNetwork

    x_train = np.linspace([[-1]] * 3, [[1]] * 3, 10, dtype=np.float32).reshape((10, 3))
    y_train = np.linspace([[-1]] * 3, [[1]] * 3, 10, dtype=np.float32).reshape((10, 3))
    noise = NormalizingFlowNetwork(3, n_flows=3, hidden_sizes=(16, 16), trainable_base_dist=True,
                                   noise_reg=("fixed_rate", 1.0))
    try:
        noise.fit(x_train, y_train, epochs=10, verbose=0)
        assert float(noise.y_noise_std) == 1.0 and float(noise.x_noise_std) == 1.0
        input_model = noise._get_input_model()
        y1 = input_model(y_train, training=False).cpu().numpy()
        y2 = input_model(y_train, training=False).cpu().numpy()
        assert np.all(y1 == y2)
        np.testing.assert_allclose(y1, (y_train - noise.y_mean.cpu().numpy()) / noise.y_std.cpu().numpy(), rtol=1e-6)
        y1 = input_model(y_train, training=True).cpu().numpy()
        y2 = input_model(y_train, training=True).cpu().numpy()
        assert not np.all(y1 == y2)
        out1 = noise.pdf(x_train, y_train).cpu().numpy()
        out2 = noise.pdf(x_train, y_train).cpu().numpy()
        assert np.all(out1 == out2) and np.all(np.isfinite(out1))
    finally:
        noise._set_noise(0.0)


@pytest.mark.parametrize("name,kwargs", [("NFN", dict(n_flows=3)), ("MDN", dict(n_centers=4)), ("KMN", dict(n_centers=5)),
                                         ("bayesian_NFN", dict(kl_weight_scale=1e-3, n_flows=2, map_mode=True)),
                                         ("bayesian_MDN", dict(kl_weight_scale=1e-3, n_centers=3, map_mode=True))])
def test_checkpoint_round_trip(cuda_device, name, kwargs):
    """save -> fresh estimator -> strict load -> identical log_pdf (the first forward after the load must not
    re-initialise the lazily created weights; the [dim]-shaped normalisation statistics and the variational
    layers' parameters must be restored)."""
    from normalizingflownetwork_b200.estimators import ESTIMATORS

    d = 2
    rng = np.random.default_rng(3)
    x = rng.normal(size=(300, d)).astype(np.float32)
    y = (np.sin(x) + 0.1 * rng.normal(size=(300, d))).astype(np.float32) * np.array([1.0, 3.0], np.float32)
    model = ESTIMATORS[name](d, **kwargs)
    model.fit(x, y, batch_size=100, epochs=2, verbose=0)
    want = model.log_pdf(x, y).clone()
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    fresh = ESTIMATORS[name](d, **kwargs)
    if name == "KMN":  # the kernel centres are data, not state: chosen by fit / set_center_points
        fresh.dist_layer.set_center_points(((y - y.mean(0)) / y.std(0)).astype(np.float32))
    res = fresh.load_state_dict(state, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert tuple(fresh.y_std.shape) == (d,)
    # (the very first call of a restored model still sees a lazy output layer and takes the unfused head: same
    # weights, different rounding)
    got = fresh.log_pdf(x, y)
    assert torch.allclose(got, want, rtol=0, atol=2e-5), float((got - want).abs().max())
    again = fresh.log_pdf(x, y)   # the second call must not have been perturbed by a late initialisation
    assert torch.allclose(again, want, rtol=0, atol=2e-5), float((again - want).abs().max())
