"""GPU: the hidden layers of the conditioning network as one kernel each way (csrc/nfn_mlp.cu) against
float64 torch, and the estimators with / without them."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("act", ["tanh", "relu", "linear", "sigmoid", "elu"])
@pytest.mark.parametrize("K,N,B", [(1, 16, 1000), (2, 32, 5000), (4, 8, 333), (16, 16, 4097), (3, 8, 129), (16, 32, 777), (64, 64, 300), (64, 32, 300), (10, 16, 1)])
def test_dense_act_matches_float64(cuda_device, nfn_lib, act, K, N, B):
    from normalizingflownetwork_b200 import functional as F

    assert F.dense_act_supported(K, N, act)
    g = torch.Generator(device=cuda_device).manual_seed(100 * K + N + B)
    x = torch.randn((B, K), generator=g, device=cuda_device)
    w = torch.randn((N, K), generator=g, device=cuda_device) / max(1.0, K ** 0.5)
    b = torch.randn(N, generator=g, device=cuda_device) * 0.1
    up = torch.randn((B, N), generator=g, device=cuda_device)
    ref_act = {"tanh": torch.tanh, "relu": torch.relu, "linear": lambda v: v, "sigmoid": torch.sigmoid,
               "elu": torch.nn.functional.elu}[act]
    x64 = x.double().requires_grad_(True)
    w64 = w.double().requires_grad_(True)
    b64 = b.double().requires_grad_(True)
    ref = ref_act(x64 @ w64.T + b64)
    ref.backward(up.double())

    xg = x.clone().requires_grad_(True)
    wg = w.clone().requires_grad_(True)
    bg = b.clone().requires_grad_(True)
    out = F.dense_act(xg, wg, bg, act)
    out.backward(up)
    tol = dict(rtol=2e-5, atol=2e-5)
    assert torch.allclose(out.double(), ref, **tol)
    assert torch.allclose(xg.grad.double(), x64.grad, **tol)
    scale = max(1.0, float(w64.grad.abs().max()))
    assert float((wg.grad.double() - w64.grad).abs().max()) <= 2e-5 * scale * max(1.0, B ** 0.5 / 8)
    assert float((bg.grad.double() - b64.grad).abs().max()) <= 2e-5 * max(1.0, float(b64.grad.abs().max())) * max(1.0, B ** 0.5 / 8)
    # forward-only call without autograd, and the first layer of a network (no input gradient wanted)
    with torch.no_grad():
        assert torch.equal(F.dense_act(x, w, b, act), out.detach())
    wg2 = w.clone().requires_grad_(True)
    bg2 = b.clone().requires_grad_(True)
    F.dense_act(x, wg2, bg2, act).backward(up)
    assert torch.allclose(bg2.grad, bg.grad, rtol=1e-4, atol=1e-4 * max(1.0, float(b64.grad.abs().max())))
    assert torch.allclose(wg2.grad, wg.grad, rtol=1e-4, atol=1e-4 * scale)   # atomics: summation order varies


def test_unsupported_layer_shapes_are_reported(cuda_device, nfn_lib):
    from normalizingflownetwork_b200 import functional as F

    assert not F.dense_act_supported(16, 10, "tanh")      # units not in {8, 16, 32, 64}
    assert not F.dense_act_supported(65, 16, "tanh")      # too many inputs
    assert not F.dense_act_supported(16, 16, "softplus")  # activation the kernels do not know


def test_estimator_same_with_and_without_fused_hidden_layers(cuda_device):
    """NormalizingFlowNetwork with its hidden layers on the fused kernels gives the same log_pdf and the same
    first optimiser step as with torch's Linear + activation."""
    from normalizingflownetwork_b200.estimators import NormalizingFlowNetwork
    from normalizingflownetwork_b200.simulation import gen_cosine_noise_data

    x, y = gen_cosine_noise_data(4096, noise_std=0.3, heterosced_noise=0.5)
    models = []
    for fused in (True, False):
        m = NormalizingFlowNetwork.build_function(n_dims=1, n_flows=3, hidden_sizes=(16, 16), activation="tanh")
        for layer in m.net:
            if hasattr(layer, "fused"):
                layer.fused = fused
        m.fit(x, y, batch_size=1024, epochs=3, verbose=0)
        models.append(m)
    a, b = (m.log_pdf(x, y).cpu().numpy() for m in models)
    assert np.max(np.abs(a - b)) <= 5e-4 * max(1.0, np.max(np.abs(b)))   # 12 Adam steps apart at fp32 rounding
    assert abs(models[0].history[-1] - models[1].history[-1]) <= 1e-3
