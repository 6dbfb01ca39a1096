"""GPU: the emitting Dense(P) layer fused into the flow kernel against the
float64 oracle composed with a float64 matmul."""
import os

import numpy as np
import pytest
import torch

from oracle import analytic_np as an

pytestmark = pytest.mark.gpu

CASES = [
    # (flow_types, d, trainable_base, H) -- first three have ahead-of-time instances, the rest are JIT-specialised
    (["planar", "radial", "affine"] * 3 + ["planar"], 2, True, 16),
    (["radial"] * 3, 1, True, 16),
    (["radial"] * 5, 1, True, 16),
    (["radial", "planar"], 3, True, 32),       # P = 18: not a multiple of 8 (padded mma tiles)
    (["affine", "radial"], 2, False, 64),      # P = 8, no base parameters
    (["planar"], 1, True, 48),                 # P = 5 (odd row stride)
]


def _oracle(h, W, b, y, ft, d, tb, up):
    t = h.astype(np.float64) @ W.astype(np.float64) + b.astype(np.float64)
    lp, dt, _ = an.chain_forward_backward(t, y, ft, d, tb, upstream=up)
    return lp, dt @ W.astype(np.float64).T, h.astype(np.float64).T @ dt, dt.sum(0)


def rel(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref)))


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("B", [1, 100, 128 * 5 + 77, 20_000])
def test_dense_chain_vs_oracle(cuda_device, nfn_lib, case, B):
    from normalizingflownetwork_b200 import functional as F

    ft, d, tb, H = CASES[case]
    P = an.layout(ft, d, tb)[1]
    rng = np.random.default_rng(1000 * case + B)
    h = np.tanh(rng.normal(0, 1.0, (B, H))).astype(np.float32)
    W = (rng.normal(0, 0.5, (H, P)) / np.sqrt(H)).astype(np.float32)
    b = rng.normal(0, 0.2, (P,)).astype(np.float32)
    y = rng.normal(0, 1.0, (B, d)).astype(np.float32)
    up = rng.normal(0, 1.0, (B,)).astype(np.float32)
    ref_lp, ref_dh, ref_dW, ref_db = _oracle(h, W, b, y, ft, d, tb, up * 0.5)
    dev = lambda x: torch.tensor(x, device=cuda_device)
    lsum = torch.zeros(1, dtype=torch.float64, device=cuda_device)
    lp, dh, dW, db = F.dense_chain_forward_backward(dev(h), dev(W), dev(b), dev(y), ft, d, tb, g_logp=dev(up),
                                                    g_scale=0.5, logp_sum=lsum)
    assert rel(lp.cpu().numpy(), ref_lp) <= 1e-5
    assert rel(dh.cpu().numpy(), ref_dh) <= 1e-4
    # dW / db are sums over B rows: compare against their own scale
    scale_w = max(1.0, np.abs(ref_dW).max())
    assert np.abs(dW.cpu().numpy() - ref_dW).max() <= 1e-4 * scale_w
    assert np.abs(db.cpu().numpy() - ref_db).max() <= 1e-4 * max(1.0, np.abs(ref_db).max())
    assert abs(lsum.item() - ref_lp.sum()) <= 1e-5 * np.abs(ref_lp).sum() + 1e-6
    lp_f = F.dense_chain_forward(dev(h), dev(W), dev(b), dev(y), ft, d, tb)
    assert rel(lp_f.cpu().numpy(), ref_lp) <= 1e-5
    # y broadcast
    ref_b = an.chain_forward_backward(h.astype(np.float64) @ W.astype(np.float64) + b, y[:1], ft, d, tb, need_grad=False)
    assert rel(F.dense_chain_forward(dev(h), dev(W), dev(b), dev(y[:1]), ft, d, tb).cpu().numpy(), ref_b) <= 1e-5


def test_dense_chain_equals_unfused_composition(cuda_device, nfn_lib):
    """Same answer as torch matmul + the plain chain kernel, and gradients accumulate (+=)."""
    from normalizingflownetwork_b200 import functional as F

    ft, d, tb, H = CASES[0]
    P, B = 48, 50_000
    g = torch.Generator(device=cuda_device).manual_seed(3)
    h = torch.tanh(torch.randn((B, H), generator=g, device=cuda_device))
    W = torch.randn((H, P), generator=g, device=cuda_device) * 0.2
    b = torch.randn(P, generator=g, device=cuda_device) * 0.1
    y = torch.randn((B, d), generator=g, device=cuda_device)
    t = (h.double() @ W.double() + b.double()).float()
    lp_u, dt_u, _ = F.chain_forward_backward(t, y, ft, d, tb, g_scale=-1.0 / B)
    lp, dh, dW, db = F.dense_chain_forward_backward(h, W, b, y, ft, d, tb, g_scale=-1.0 / B)
    assert torch.allclose(lp, lp_u, rtol=1e-5, atol=1e-5)
    assert torch.allclose(dh * B, (dt_u.double() @ W.double().T).float() * B, rtol=1e-3, atol=1e-3)
    assert torch.allclose(dW, (h.double().T @ dt_u.double()).float(), rtol=1e-3, atol=1e-5)
    assert torch.allclose(db, dt_u.double().sum(0).float(), rtol=1e-3, atol=1e-5)
    dW2, db2 = dW.clone(), db.clone()
    F.dense_chain_forward_backward(h, W, b, y, ft, d, tb, g_scale=-1.0 / B, dW=dW2, dbias=db2)
    assert torch.allclose(dW2, 2 * dW, rtol=1e-4, atol=1e-6) and torch.allclose(db2, 2 * db, rtol=1e-4, atol=1e-6)


def test_dense_chain_unsupported_width_is_reported(cuda_device, nfn_lib):
    from normalizingflownetwork_b200 import _lib
    from normalizingflownetwork_b200 import functional as F

    with pytest.raises(_lib.NfnError) as ei:
        F.dense_chain_forward(torch.zeros((4, 10), device=cuda_device), torch.zeros((10, 11), device=cuda_device),
                              torch.zeros(11, device=cuda_device), torch.zeros((4, 1), device=cuda_device),
                              ["radial"] * 3, 1, True)
    assert ei.value.code == -6
