"""GPU: S posterior weight draws folded into the batch (reference BayesianNNEstimator.py:65-76, BASELINE config 4) on
the fused kernels: per-draw first layer, per-draw emitting layer + flow chain, and the estimator's training step."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N,NP,act", [(1, 10, 16, "tanh"), (3, 16, 16, "relu"), (8, 20, 32, "sigmoid"), (2, 5, 8, "elu")])
@pytest.mark.parametrize("S,B", [(1, 77), (4, 1000), (32, 513)])
def test_first_layer_with_folded_draws(cuda_device, nfn_lib, K, N, NP, act, S, B):
    from normalizingflownetwork_b200 import functional as F

    assert F.dense_act_draws_supported(K, N, NP, act)
    g = torch.Generator(device=cuda_device).manual_seed(100 * S + B + K)
    x = torch.randn((B, K), generator=g, device=cuda_device) * 2.0 + 0.5
    w = torch.randn((S, K * N + N), generator=g, device=cuda_device) * 0.7
    mean, std = x.mean(0), x.std(0)
    out = F.dense_act_forward_draws(x, w, N, act, NP, x_mean=mean, x_std=std)
    xn = ((x - mean) / (std + 1e-8)).double()
    wd = w.double().requires_grad_(True)
    pre = torch.baddbmm(wd[:, K * N:].unsqueeze(1), xn.unsqueeze(0).expand(S, -1, -1), wd[:, :K * N].view(S, K, N))
    fn = {"tanh": torch.tanh, "relu": torch.relu, "sigmoid": torch.sigmoid, "elu": torch.nn.functional.elu}[act]
    ref = fn(pre)
    assert out.shape == (S * B, NP)
    assert torch.allclose(out.view(S, B, NP)[:, :, :N].double(), ref, rtol=1e-5, atol=1e-5)
    assert float(out.view(S, B, NP)[:, :, N:].abs().max()) == 0.0 if NP > N else True
    up = torch.randn((S * B, NP), generator=g, device=cuda_device)
    dw = F.dense_act_backward_draws(x, out, up, S, N, act, x_mean=mean, x_std=std)
    ref.backward(up.view(S, B, NP)[:, :, :N].double())
    scale = max(1.0, float(wd.grad.abs().max()))
    assert float((dw.double() - wd.grad).abs().max()) <= 2e-4 * scale
    # without the fused normalisation
    out2 = F.dense_act_forward_draws(x, w, N, act, NP)
    pre2 = torch.baddbmm(w[:, K * N:].unsqueeze(1), x.unsqueeze(0).expand(S, -1, -1), w[:, :K * N].view(S, K, N))
    assert torch.allclose(out2.view(S, B, NP)[:, :, :N], fn(pre2), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("S,Bd", [(1, 300), (5, 128), (7, 1000), (32, 4099)])
def test_dense_chain_with_folded_draws_equals_one_launch_per_draw(cuda_device, nfn_lib, S, Bd):
    """Per-draw weights in ONE launch == S launches of the same kernel, one per draw (and y is read per sample)."""
    from normalizingflownetwork_b200 import functional as F

    ft, d, tb, H = ["radial"] * 5, 1, True, 16
    P = F.chain_param_size(ft, d, tb)
    g = torch.Generator(device=cuda_device).manual_seed(7 * S + Bd)
    h = torch.tanh(torch.randn((S * Bd, H), generator=g, device=cuda_device))
    W = torch.randn((S, H, P), generator=g, device=cuda_device) * 0.2
    b = torch.randn((S, P), generator=g, device=cuda_device) * 0.1
    y = torch.randn((Bd, d), generator=g, device=cuda_device)
    up = torch.randn(S * Bd, generator=g, device=cuda_device)
    ls = torch.zeros(1, dtype=torch.float64, device=cuda_device)
    lp, dh, dW, db = F.dense_chain_forward_backward_draws(h, W, b, y, ft, d, tb, g_logp=up, g_scale=0.25, logp_sum=ls)
    F.set_option("dense_mma", "sync")
    try:
        for s in range(S):
            sl = slice(s * Bd, (s + 1) * Bd)
            lp_s, dh_s, dW_s, db_s = F.dense_chain_forward_backward(h[sl], W[s], b[s], y, ft, d, tb, g_logp=up[sl],
                                                                    g_scale=0.25)
            assert torch.allclose(lp[sl], lp_s, rtol=1e-6, atol=1e-6)
            assert torch.allclose(dh[sl], dh_s, rtol=1e-5, atol=1e-6)
            assert torch.allclose(dW[s], dW_s, rtol=1e-4, atol=1e-5 * max(1.0, float(dW_s.abs().max())))
            assert torch.allclose(db[s], db_s, rtol=1e-4, atol=1e-5 * max(1.0, float(db_s.abs().max())))
    finally:
        F.set_option("dense_mma", "auto")
    assert abs(ls.item() - lp.double().sum().item()) <= 1e-6 + 1e-9 * float(lp.double().abs().sum())
    lp_f = F.dense_chain_forward_draws(h, W, b, y, ft, d, tb)
    assert torch.allclose(lp_f, lp, rtol=1e-6, atol=1e-6)
    # one y row for everybody
    lp_b = F.dense_chain_forward_draws(h, W, b, y[:1], ft, d, tb)
    lp_r = F.dense_chain_forward_draws(h, W, b, y[:1].expand(Bd, -1).contiguous(), ft, d, tb)
    assert torch.equal(lp_b, lp_r)


@pytest.mark.parametrize("noise", [0.0, 0.2])
@pytest.mark.parametrize("head", ["nfn", "mdn"])
def test_bayesian_train_step_on_the_folded_draw_kernels(cuda_device, nfn_lib, noise, head):
    """Bayesian NFN / MDN with S = 8 folded draws: the fused step (3 kernels, no t / dt / repeated y) leaves the
    same loss and the same gradients as the composed one (batched GEMMs + the streaming head + autograd)."""
    from normalizingflownetwork_b200.estimators import BayesMixtureDensityNetwork, BayesNormalizingFlowNetwork

    rng = np.random.default_rng(5)
    x = rng.uniform(-3, 3, (2000, 1)).astype(np.float32)
    y = (np.cos(x) + 0.3 * rng.normal(0, 1, (2000, 1))).astype(np.float32)
    out = []
    for fuse in (True, "autograd", False):   # one library call / per-kernel calls under autograd / composed torch path
        if head == "mdn":
            m = BayesMixtureDensityNetwork(1, kl_weight_scale=1.0 / 2000, n_centers=5, hidden_sizes=(10,),
                                           activation="tanh", n_train_draws=8, learning_rate=1e-2,
                                           noise_reg=("fixed_rate", noise))
        else:
            m = BayesNormalizingFlowNetwork(1, kl_weight_scale=1.0 / 2000, n_flows=5, hidden_sizes=(10,),
                                            activation="tanh", n_train_draws=8, learning_rate=1e-2,
                                            noise_reg=("fixed_rate", noise))
        m.fuse_draws = bool(fuse)
        if fuse == "autograd":
            m._one_call_ok = lambda plan: False
        m._assign_data_normalization(x, y)
        with torch.no_grad():
            m.params_from_x(x[:2])
        m.optimizer = torch.optim.Adam(m.parameters(), lr=m.learning_rate, eps=1e-7)
        m._wgen = None
        assert (m._fused_draws_plan() is not None) == bool(fuse) or noise > 0.0
        loss = m.train_step(m._to_dev(x), m._to_dev(y))
        out.append((float(loss), [p.grad.detach().clone() for p in m.parameters() if p.grad is not None]))
    l1, g1 = out[-1]
    for l0, g0 in out[:-1]:
        assert np.isfinite(l0) and abs(l0 - l1) <= 2e-5 * max(1.0, abs(l1))
        assert len(g0) == len(g1) and len(g0) >= 2
        for a, b in zip(g0, g1):
            assert float((a - b).abs().max()) <= 2e-4 * max(1e-3, float(b.abs().max()))


def test_bayesian_train_step_replays_as_a_cuda_graph(cuda_device, nfn_lib):
    """The S-draw training step captured once and replayed: fresh weight draws on every replay (the private generator is
    registered with the graph), finite decreasing loss, parameters move."""
    from normalizingflownetwork_b200.estimators import BayesNormalizingFlowNetwork

    rng = np.random.default_rng(8)
    x = rng.uniform(-3, 3, (4096, 1)).astype(np.float32)
    y = (np.cos(x) + 0.3 * rng.normal(0, 1, (4096, 1))).astype(np.float32)
    m = BayesNormalizingFlowNetwork(1, kl_weight_scale=1.0 / 4096, n_flows=5, hidden_sizes=(10,), activation="tanh",
                                    n_train_draws=8, learning_rate=1e-2)
    m._assign_data_normalization(x, y)
    with torch.no_grad():
        m.params_from_x(x[:2])
    m.optimizer = torch.optim.Adam(m.parameters(), lr=m.learning_rate, eps=1e-7)
    assert m._fused_draws_plan() is not None
    xd, yd = m._to_dev(x), m._to_dev(y)
    m.capture_train_step(4096, 1, 1)
    before = [p.detach().clone() for p in m.parameters()]
    losses = [float(m.train_step_graphed(xd, yd)) for _ in range(60)]
    assert all(np.isfinite(losses))
    assert len(set(round(v, 6) for v in losses[:5])) > 1          # the replays are not one frozen draw
    assert np.mean(losses[-10:]) < np.mean(losses[:10])
    assert any(float((p - q).abs().max()) > 0 for p, q in zip(m.parameters(), before))


@pytest.mark.parametrize("trainable_prior", [False, True])
def test_variational_sample_matches_torch_distributions(cuda_device, nfn_lib, trainable_prior):
    """w = loc + sigma * eps and the exact Normal-Normal KL in one kernel each way against the composed torch version
    (MeanFieldLayer + torch.distributions.kl_divergence), values and gradients."""
    from normalizingflownetwork_b200 import functional as F
    from normalizingflownetwork_b200.DistributionLayers import MeanFieldLayer

    n, S, ps = 187, 32, 0.7
    g = torch.Generator(device=cuda_device).manual_seed(4)
    params = (torch.randn(2 * n, generator=g, device=cuda_device) * 2.0).requires_grad_(True)
    prior = (torch.randn(n, generator=g, device=cuda_device) * 0.3).requires_grad_(trainable_prior)
    eps = torch.randn((S, n), generator=g, device=cuda_device)
    up = torch.randn((S, n), generator=g, device=cuda_device)
    w, kl = F.variational_sample(params, prior, eps, ps)
    (w * up).sum().add(0.37 * kl).backward()
    got = (w.detach().clone(), float(kl), params.grad.clone(), prior.grad.clone() if trainable_prior else None)
    p2 = params.detach().double().requires_grad_(True)
    pr2 = prior.detach().double().requires_grad_(trainable_prior)
    q, r = MeanFieldLayer(n, scale=None)(p2), MeanFieldLayer(n, scale=ps)(pr2)
    w_ref = q.base_dist.loc + q.base_dist.scale * eps.double()
    kl_ref = torch.distributions.kl_divergence(q, r)
    (w_ref * up.double()).sum().add(0.37 * kl_ref).backward()
    assert torch.allclose(got[0].double(), w_ref, rtol=1e-5, atol=1e-6)
    assert got[1] == pytest.approx(float(kl_ref), rel=1e-5)
    assert float((got[2].double() - p2.grad).abs().max()) <= 1e-4 * float(p2.grad.abs().max())
    if trainable_prior:
        assert float((got[3].double() - pr2.grad).abs().max()) <= 1e-4 * max(1e-6, float(pr2.grad.abs().max()))
