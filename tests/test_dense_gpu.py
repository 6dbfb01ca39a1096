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
    (["radial", "planar"] * 8, 4, True, 16),   # BASELINE config 3 chain: P = 128 (three GEMM-3 passes, 512 TMEM columns)
]


def _oracle(h, W, b, y, ft, d, tb, up):
    t = h.astype(np.float64) @ W.astype(np.float64) + b.astype(np.float64)
    lp, dt, _ = an.chain_forward_backward(t, y, ft, d, tb, upstream=up)
    # conditioning of each row w.r.t. its parameters: rows where |dlogp/dt| is huge (a planar flow
    # with w = 1 + w_raw ~ 0) amplify ANY fp32 rounding of t -- an exact fp32 matmul already misses
    # 1e-5 there -- so the log-prob bar is checked on the well-conditioned rows (the vast majority)
    unit = an.chain_forward_backward(t, y, ft, d, tb)[1]
    well = np.abs(unit).max(1) < 50.0
    return lp, dt @ W.astype(np.float64).T, h.astype(np.float64).T @ dt, dt.sum(0), well


@pytest.fixture(params=["auto", "sync", "tc5"])
def dense_impl(request):
    """auto: the library's own choice (tcgen05 / TMEM kernel for P >= 32, warp-level mma.sync below);
    sync / tc5: one implementation forced for every chain that has both (ahead-of-time instances)."""
    from normalizingflownetwork_b200 import functional as F

    F.set_option("dense_mma", request.param)
    yield request.param
    F.set_option("dense_mma", "auto")


def rel(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref)))


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("B", [1, 100, 128 * 5 + 77, 20_000])
def test_dense_chain_vs_oracle(cuda_device, nfn_lib, dense_impl, case, B):
    from normalizingflownetwork_b200 import functional as F

    ft, d, tb, H = CASES[case]
    P = an.layout(ft, d, tb)[1]
    rng = np.random.default_rng(1000 * case + B)
    h = np.tanh(rng.normal(0, 1.0, (B, H))).astype(np.float32)
    W = (rng.normal(0, 0.5, (H, P)) / np.sqrt(H)).astype(np.float32)
    b = rng.normal(0, 0.2, (P,)).astype(np.float32)
    y = rng.normal(0, 1.0, (B, d)).astype(np.float32)
    up = rng.normal(0, 1.0, (B,)).astype(np.float32)
    ref_lp, ref_dh, ref_dW, ref_db, well = _oracle(h, W, b, y, ft, d, tb, up * 0.5)
    assert well.mean() > 0.8
    dev = lambda x: torch.tensor(x, device=cuda_device)
    lsum = torch.zeros(1, dtype=torch.float64, device=cuda_device)
    lp, dh, dW, db = F.dense_chain_forward_backward(dev(h), dev(W), dev(b), dev(y), ft, d, tb, g_logp=dev(up),
                                                    g_scale=0.5, logp_sum=lsum)
    assert rel(lp.cpu().numpy()[well], ref_lp[well]) <= 1e-5
    assert rel(lp.cpu().numpy(), ref_lp) <= 1e-3
    assert rel(dh.cpu().numpy()[well], ref_dh[well]) <= 1e-4
    # dW / db are sums over ALL rows, ill-conditioned ones included (their dt carries the amplified
    # rounding of t): compare against the sums' own scale; the tight GEMM check is the composition test
    scale_w = max(1.0, np.abs(ref_dW).max())
    assert np.abs(dW.cpu().numpy() - ref_dW).max() <= 2e-3 * scale_w
    assert np.abs(db.cpu().numpy() - ref_db).max() <= 2e-3 * max(1.0, np.abs(ref_db).max())
    assert abs(lsum.item() - lp.double().sum().item()) <= 1e-9 * np.abs(ref_lp).sum() + 1e-6  # fp64 accumulator
    lp_f = F.dense_chain_forward(dev(h), dev(W), dev(b), dev(y), ft, d, tb)
    assert rel(lp_f.cpu().numpy()[well], ref_lp[well]) <= 1e-5
    # y broadcast
    tb64 = h.astype(np.float64) @ W.astype(np.float64) + b
    ref_b = an.chain_forward_backward(tb64, y[:1], ft, d, tb, need_grad=False)
    well_b = np.abs(an.chain_forward_backward(tb64, y[:1], ft, d, tb)[1]).max(1) < 50.0
    got_b = F.dense_chain_forward(dev(h), dev(W), dev(b), dev(y[:1]), ft, d, tb).cpu().numpy()
    assert rel(got_b[well_b], ref_b[well_b]) <= 1e-5


def test_dense_chain_equals_unfused_composition(cuda_device, nfn_lib, dense_impl):
    """Same answer as torch matmul + the plain chain kernel, and gradients accumulate (+=)."""
    from normalizingflownetwork_b200 import functional as F

    ft, d, tb, H = CASES[0]
    P, B = 48, 50_000
    g = torch.Generator(device=cuda_device).manual_seed(3)
    h = torch.tanh(torch.randn((B, H), generator=g, device=cuda_device))
    # small parameter scale: no ill-conditioned rows, so the sums over the batch are comparable too
    W = torch.randn((H, P), generator=g, device=cuda_device) * 0.06
    b = torch.randn(P, generator=g, device=cuda_device) * 0.05
    y = torch.randn((B, d), generator=g, device=cuda_device)
    t = (h.double() @ W.double() + b.double()).float()
    lp_u, dt_u, _ = F.chain_forward_backward(t, y, ft, d, tb, g_scale=-1.0 / B)
    lp, dh, dW, db = F.dense_chain_forward_backward(h, W, b, y, ft, d, tb, g_scale=-1.0 / B)
    # rows whose gradient is huge amplify the last-bit difference between the two t's; compare the rest
    well = (dt_u.abs().amax(1) * B) < 50.0
    assert well.float().mean() > 0.8
    assert torch.allclose(lp[well], lp_u[well], rtol=1e-5, atol=1e-5)
    dh_ref = (dt_u.double() @ W.double().T).float()
    assert torch.allclose(dh[well] * B, dh_ref[well] * B, rtol=1e-3, atol=1e-3)
    dW_ref, db_ref = (h.double().T @ dt_u.double()).float(), dt_u.double().sum(0).float()
    assert (dW - dW_ref).abs().max() <= 2e-3 * dW_ref.abs().max()
    assert (db - db_ref).abs().max() <= 2e-3 * db_ref.abs().max()
    dW2, db2 = dW.clone(), db.clone()
    F.dense_chain_forward_backward(h, W, b, y, ft, d, tb, g_scale=-1.0 / B, dW=dW2, dbias=db2)
    assert torch.allclose(dW2, 2 * dW, rtol=1e-4, atol=1e-6) and torch.allclose(db2, 2 * db, rtol=1e-4, atol=1e-6)


def test_dense_chain_many_tiles_per_cta(cuda_device, nfn_lib, dense_impl):
    """Every CTA runs several tiles (the tcgen05 kernel pipelines GEMM 1 of tile i+1 ahead of GEMM 2/3 of
    tile i and drains dh / dW one tile late): ragged 150,001 rows against the float64 oracle."""
    from normalizingflownetwork_b200 import functional as F

    ft, d, tb, H = CASES[0]
    P, B = 48, 150_001
    rng = np.random.default_rng(77)
    h = np.tanh(rng.normal(0, 1.0, (B, H))).astype(np.float32)
    W = (rng.normal(0, 0.3, (H, P)) / np.sqrt(H)).astype(np.float32)
    b = rng.normal(0, 0.1, (P,)).astype(np.float32)
    y = rng.normal(0, 1.0, (B, d)).astype(np.float32)
    ref_lp, ref_dh, ref_dW, ref_db, well = _oracle(h, W, b, y, ft, d, tb, -1.0 / B)
    dev = lambda x: torch.tensor(x, device=cuda_device)
    lp, dh, dW, db = F.dense_chain_forward_backward(dev(h), dev(W), dev(b), dev(y), ft, d, tb, g_scale=-1.0 / B)
    assert rel(lp.cpu().numpy()[well], ref_lp[well]) <= 1e-5
    assert rel(dh.cpu().numpy()[well] * B, ref_dh[well] * B) <= 1e-4
    assert np.abs(dW.cpu().numpy() - ref_dW).max() <= 2e-3 * max(1.0, np.abs(ref_dW).max())
    assert np.abs(db.cpu().numpy() - ref_db).max() <= 2e-3 * max(1.0, np.abs(ref_db).max())
    lp_f = F.dense_chain_forward(dev(h), dev(W), dev(b), dev(y), ft, d, tb)
    assert torch.equal(lp_f, lp)


def test_dense_chain_unsupported_width_is_reported(cuda_device, nfn_lib):
    from normalizingflownetwork_b200 import _lib
    from normalizingflownetwork_b200 import functional as F

    with pytest.raises(_lib.NfnError) as ei:
        F.dense_chain_forward(torch.zeros((4, 10), device=cuda_device), torch.zeros((10, 11), device=cuda_device),
                              torch.zeros(11, device=cuda_device), torch.zeros((4, 1), device=cuda_device),
                              ["radial"] * 3, 1, True)
    assert ei.value.code == -6


# ----------------------------------------------------------------------------- fused Dense(P) + MDN head
MDN_CASES = [
    # (n_centers, d, H) -- the first two have ahead-of-time instances, the rest are runtime-specialised
    (20, 2, 16),   # BASELINE config 5: P = 100
    (5, 1, 16),    # the reference's default MixtureDensityNetwork: P = 15 (odd row stride)
    (7, 3, 32),    # P = 49
    (4, 4, 64),    # P = 36 (padded row stride)
    (3, 2, 48),    # P = 15
]


def _mdn_oracle(h, W, b, y, K, d, up):
    t = h.astype(np.float64) @ W.astype(np.float64) + b.astype(np.float64)
    lp, dt, _ = an.mdn_forward_backward(t, y, K, d, upstream=up)
    return lp, dt @ W.astype(np.float64).T, h.astype(np.float64).T @ dt, dt.sum(0)


@pytest.mark.parametrize("case", range(len(MDN_CASES)))
@pytest.mark.parametrize("B", [1, 100, 128 * 5 + 77, 20_000])
def test_dense_mdn_vs_oracle(cuda_device, nfn_lib, case, B):
    """Dense(P) + Gaussian mixture head in one kernel against the float64 oracle composed with float64 matmuls
    (reference MaximumLikelihoodNNEstimator.py:43 + DistributionLayers.py:196-212)."""
    from normalizingflownetwork_b200 import functional as F

    K, d, H = MDN_CASES[case]
    P = K * (2 * d + 1)
    assert F.dense_mdn_supported(H, K, d)
    rng = np.random.default_rng(5000 * case + B)
    h = np.tanh(rng.normal(0, 1.0, (B, H))).astype(np.float32)
    W = (rng.normal(0, 1.0, (H, P)) / np.sqrt(H)).astype(np.float32)
    b = rng.normal(0, 0.5, (P,)).astype(np.float32)
    y = rng.normal(0, 1.0, (B, d)).astype(np.float32)
    up = rng.normal(0, 1.0, (B,)).astype(np.float32)
    ref_lp, ref_dh, ref_dW, ref_db = _mdn_oracle(h, W, b, y, K, d, up * 0.5)
    dev = lambda x: torch.tensor(x, device=cuda_device)
    lsum = torch.zeros(1, dtype=torch.float64, device=cuda_device)
    lp, dh, dW, db = F.dense_mdn_forward_backward(dev(h), dev(W), dev(b), dev(y), K, d, g_logp=dev(up), g_scale=0.5,
                                                  logp_sum=lsum)
    assert rel(lp.cpu().numpy(), ref_lp) <= 1e-5
    assert rel(dh.cpu().numpy(), ref_dh) <= 1e-4
    assert np.abs(dW.cpu().numpy() - ref_dW).max() <= 1e-4 * max(1.0, np.abs(ref_dW).max())
    assert np.abs(db.cpu().numpy() - ref_db).max() <= 1e-4 * max(1.0, np.abs(ref_db).max())
    assert abs(lsum.item() - lp.double().sum().item()) <= 1e-9 * np.abs(ref_lp).sum() + 1e-6
    lp_f = F.dense_mdn_forward(dev(h), dev(W), dev(b), dev(y), K, d)
    assert rel(lp_f.cpu().numpy(), ref_lp) <= 1e-5
    # y broadcast
    t64 = h.astype(np.float64) @ W.astype(np.float64) + b
    ref_b = an.mdn_forward_backward(t64, y[:1], K, d, need_grad=False)
    got_b = F.dense_mdn_forward(dev(h), dev(W), dev(b), dev(y[:1]), K, d).cpu().numpy()
    assert rel(got_b, ref_b) <= 1e-5


def test_dense_mdn_equals_unfused_composition(cuda_device, nfn_lib):
    """Same answer as float64 matmul + the streaming MDN kernel (identical row arithmetic: nfn_mixture_row.cuh),
    gradients accumulate (+=), the y pipeline of the estimators rides along, ragged multi-tile batch."""
    from normalizingflownetwork_b200 import functional as F

    K, d, H = 20, 2, 16
    P, B = K * (2 * d + 1), 150_001
    g = torch.Generator(device=cuda_device).manual_seed(9)
    h = torch.tanh(torch.randn((B, H), generator=g, device=cuda_device))
    W = torch.randn((H, P), generator=g, device=cuda_device) * 0.25
    b = torch.randn(P, generator=g, device=cuda_device) * 0.3
    y = torch.randn((B, d), generator=g, device=cuda_device) * 2.0 + 1.0
    xf = F.make_xform(d, mean=[1.0, 0.8], std=[2.0, 1.7], logp_shift=-float(np.log(2.0) + np.log(1.7)))
    t = (h.double() @ W.double() + b.double()).float()
    lp_u, dt_u, _ = F.mdn_forward_backward(t, y, K, d, g_scale=-1.0 / B, xform=xf)
    lp, dh, dW, db = F.dense_mdn_forward_backward(h, W, b, y, K, d, g_scale=-1.0 / B, xform=xf)
    assert torch.allclose(lp, lp_u, rtol=1e-5, atol=1e-5)
    dh_ref = (dt_u.double() @ W.double().T).float()
    assert torch.allclose(dh * B, dh_ref * B, rtol=1e-4, atol=1e-4)
    dW_ref, db_ref = (h.double().T @ dt_u.double()).float(), dt_u.double().sum(0).float()
    assert (dW - dW_ref).abs().max() <= 1e-4 * dW_ref.abs().max()
    assert (db - db_ref).abs().max() <= 1e-4 * db_ref.abs().max()
    dW2, db2 = dW.clone(), db.clone()
    F.dense_mdn_forward_backward(h, W, b, y, K, d, g_scale=-1.0 / B, dW=dW2, dbias=db2, xform=xf)
    assert torch.allclose(dW2, 2 * dW, rtol=1e-4, atol=1e-7) and torch.allclose(db2, 2 * db, rtol=1e-4, atol=1e-7)
    assert torch.allclose(F.dense_mdn_forward(h, W, b, y, K, d, xform=xf), lp, rtol=1e-6, atol=1e-6)


def test_dense_mdn_unsupported_shapes_are_reported(cuda_device, nfn_lib):
    from normalizingflownetwork_b200 import _lib
    from normalizingflownetwork_b200 import functional as F

    assert not F.dense_mdn_supported(16, 200, 2) and not F.dense_mdn_supported(10, 5, 1)
    with pytest.raises(_lib.NfnError) as ei:   # hidden width 10
        F.dense_mdn_forward(torch.zeros((4, 10), device=cuda_device), torch.zeros((10, 15), device=cuda_device),
                            torch.zeros(15, device=cuda_device), torch.zeros((4, 1), device=cuda_device), 5, 1)
    assert ei.value.code == -6


def test_mdn_estimator_uses_the_fused_layer(cuda_device, nfn_lib):
    """MixtureDensityNetwork: log_pdf and the training step's gradients through the fused Dense(P)+MDN kernel equal the
    unfused path (torch layer + streaming head + autograd) on the same weights."""
    from normalizingflownetwork_b200.estimators import MixtureDensityNetwork

    rng = np.random.default_rng(3)
    x = rng.normal(0, 1, (4096, 1)).astype(np.float32)
    y = (np.sin(x) + 0.3 * rng.normal(0, 1, (4096, 1))).astype(np.float32)
    torch.manual_seed(11)
    m = MixtureDensityNetwork.build_function(n_dims=1, n_centers=5, hidden_sizes=(16, 16), activation="tanh")
    m.fit(x, y, batch_size=1024, epochs=2, verbose=0)
    assert m._fusable_last_layer() is not None
    lp_f = torch.as_tensor(m.log_pdf(x, y)).cpu()
    m.fuse_last_layer = False
    assert m._fusable_last_layer() is None
    lp_u = torch.as_tensor(m.log_pdf(x, y)).cpu()
    assert torch.allclose(lp_f, lp_u, rtol=1e-5, atol=1e-5)
    # gradients of one step (lr = 0: the weights stay put)
    m.optimizer = torch.optim.SGD(m.parameters(), lr=0.0)
    xd, yd = m._to_dev(x), m._to_dev(y)
    grads = []
    for fuse in (True, False):
        m.fuse_last_layer = fuse
        loss = float(m.train_step(xd, yd))
        grads.append((loss, [p.grad.detach().clone() for p in m.parameters() if p.grad is not None]))
    (l0, g0), (l1, g1) = grads
    assert abs(l0 - l1) <= 1e-5 * max(1.0, abs(l1))
    assert len(g0) == len(g1) and len(g0) >= 6
    for a_, b_ in zip(g0, g1):
        assert float((a_ - b_).abs().max()) <= 2e-4 * max(1e-3, float(b_.abs().max()))


# ----------------------------------------------------------------------------- fused Dense(P) + KMN head
KMN_CASES = [
    # (kernels M, d, H) -- the first has an ahead-of-time instance, the rest are runtime-specialised
    (100, 1, 16),   # the reference's default KernelMixtureNetwork: 50 centres x 2 bandwidths
    (60, 1, 16),    # its build_function default: 30 x 2
    (24, 3, 32),
    (7, 2, 48),     # odd width: scalar logit access
]


@pytest.mark.parametrize("case", range(len(KMN_CASES)))
@pytest.mark.parametrize("B", [1, 100, 128 * 5 + 77, 20_000])
def test_dense_kmn_vs_oracle(cuda_device, nfn_lib, case, B):
    """Dense(P) + kernel-mixture head in one kernel against the float64 oracle composed with float64 matmuls
    (reference MaximumLikelihoodNNEstimator.py:43 + DistributionLayers.py:118-133), negative bandwidths included."""
    from normalizingflownetwork_b200 import functional as F

    M, d, H = KMN_CASES[case]
    assert F.dense_kmn_supported(H, M, d)
    rng = np.random.default_rng(7000 * case + B)
    h = np.tanh(rng.normal(0, 1.0, (B, H))).astype(np.float32)
    W = (rng.normal(0, 1.0, (H, M)) / np.sqrt(H)).astype(np.float32)
    b = rng.normal(0, 0.5, (M,)).astype(np.float32)
    y = rng.normal(0, 1.0, (B, d)).astype(np.float32)
    locs = rng.normal(0, 1.0, (M, d)).astype(np.float32)
    scales = (rng.uniform(0.3, 0.9, (M,)) * rng.choice([-1.0, 1.0], (M,))).astype(np.float32)
    up = rng.normal(0, 1.0, (B,)).astype(np.float32)
    t = h.astype(np.float64) @ W.astype(np.float64) + b.astype(np.float64)
    ref_lp, dt, ref_dsc, _ = an.kmn_forward_backward(t, y, locs, scales, upstream=up * 0.5)
    ref_dh, ref_dW, ref_db = dt @ W.astype(np.float64).T, h.astype(np.float64).T @ dt, dt.sum(0)
    dev = lambda x: torch.tensor(x, device=cuda_device)
    lsum = torch.zeros(1, dtype=torch.float64, device=cuda_device)
    lp, dh, dW, db, dsc = F.dense_kmn_forward_backward(dev(h), dev(W), dev(b), dev(y), dev(locs), dev(scales),
                                                      g_logp=dev(up), g_scale=0.5, logp_sum=lsum)
    assert rel(lp.cpu().numpy(), ref_lp) <= 1e-5
    assert rel(dh.cpu().numpy(), ref_dh) <= 1e-4
    assert np.abs(dW.cpu().numpy() - ref_dW).max() <= 1e-4 * max(1.0, np.abs(ref_dW).max())
    assert np.abs(db.cpu().numpy() - ref_db).max() <= 1e-4 * max(1.0, np.abs(ref_db).max())
    assert np.abs(dsc.cpu().numpy() - ref_dsc).max() <= 1e-4 * max(1.0, np.abs(ref_dsc).max()) * max(1.0, B ** 0.5 / 8)
    assert abs(lsum.item() - lp.double().sum().item()) <= 1e-9 * np.abs(ref_lp).sum() + 1e-6
    lp_f = F.dense_kmn_forward(dev(h), dev(W), dev(b), dev(y), dev(locs), dev(scales))
    assert rel(lp_f.cpu().numpy(), ref_lp) <= 1e-5
    got_b = F.dense_kmn_forward(dev(h), dev(W), dev(b), dev(y[:1]), dev(locs), dev(scales)).cpu().numpy()
    assert rel(got_b, an.kmn_forward_backward(t, y[:1], locs, scales, need_grad=False)) <= 1e-5
    # fixed bandwidths: no gradient buffer
    out = F.dense_kmn_forward_backward(dev(h), dev(W), dev(b), dev(y), dev(locs), dev(scales), g_logp=dev(up), g_scale=0.5,
                                       want_dscales=False)
    assert out[4] is None and torch.allclose(out[0], lp)


def test_dense_kmn_equals_unfused_composition(cuda_device, nfn_lib):
    """Same answer as float64 matmul + the streaming KMN kernel (identical row arithmetic: nfn_mixture_row.cuh) on a
    ragged multi-tile batch, with the estimators' y pipeline."""
    from normalizingflownetwork_b200 import functional as F

    M, d, H, B = 100, 1, 16, 150_001
    g = torch.Generator(device=cuda_device).manual_seed(19)
    h = torch.tanh(torch.randn((B, H), generator=g, device=cuda_device))
    W = torch.randn((H, M), generator=g, device=cuda_device) * 0.25
    b = torch.randn(M, generator=g, device=cuda_device) * 0.3
    y = torch.randn((B, d), generator=g, device=cuda_device) * 2.0 + 1.0
    locs = torch.randn((M, d), generator=g, device=cuda_device)
    scales = torch.rand(M, generator=g, device=cuda_device) * 0.5 + 0.3
    xf = F.make_xform(d, mean=[1.0], std=[2.0], logp_shift=-float(np.log(2.0)))
    t = (h.double() @ W.double() + b.double()).float()
    lp_u, dt_u, _, dsc_u = F.kmn_forward_backward(t, y, locs, scales, g_scale=-1.0 / B, xform=xf)
    lp, dh, dW, db, dsc = F.dense_kmn_forward_backward(h, W, b, y, locs, scales, g_scale=-1.0 / B, xform=xf)
    assert torch.allclose(lp, lp_u, rtol=1e-5, atol=1e-5)
    assert torch.allclose(dh * B, (dt_u.double() @ W.double().T).float() * B, rtol=1e-4, atol=1e-4)
    dW_ref, db_ref = (h.double().T @ dt_u.double()).float(), dt_u.double().sum(0).float()
    assert (dW - dW_ref).abs().max() <= 1e-4 * dW_ref.abs().max()
    assert (db - db_ref).abs().max() <= 1e-4 * db_ref.abs().max()
    assert (dsc - dsc_u).abs().max() <= 1e-4 * max(1e-6, float(dsc_u.abs().max()))


def test_kmn_estimator_uses_the_fused_layer(cuda_device, nfn_lib):
    """KernelMixtureNetwork: log_pdf and the training step's gradients (bandwidths included) through the fused
    Dense(P)+KMN kernel equal the unfused path on the same weights."""
    from normalizingflownetwork_b200.estimators import KernelMixtureNetwork

    rng = np.random.default_rng(13)
    x = rng.normal(0, 1, (4096, 1)).astype(np.float32)
    y = (np.sin(x) + 0.3 * rng.normal(0, 1, (4096, 1))).astype(np.float32)
    torch.manual_seed(5)
    m = KernelMixtureNetwork.build_function(n_dims=1, n_centers=30, hidden_sizes=(16, 16), activation="tanh")
    m.fit(x, y, batch_size=1024, epochs=2, verbose=0)
    assert m._fusable_last_layer() is not None
    lp_f = torch.as_tensor(m.log_pdf(x, y)).cpu()
    m.fuse_last_layer = False
    lp_u = torch.as_tensor(m.log_pdf(x, y)).cpu()
    assert torch.allclose(lp_f, lp_u, rtol=1e-5, atol=1e-5)
    m.optimizer = torch.optim.SGD(m.parameters(), lr=0.0)
    xd, yd = m._to_dev(x), m._to_dev(y)
    grads = []
    for fuse in (True, False):
        m.fuse_last_layer = fuse
        loss = float(m.train_step(xd, yd))
        grads.append((loss, {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}))
    (l0, g0), (l1, g1) = grads
    assert abs(l0 - l1) <= 1e-5 * max(1.0, abs(l1))
    assert set(g0) == set(g1) and len(g0) >= 7        # three layers' weights and biases + the bandwidths
    for n in g0:
        assert float((g0[n] - g1[n]).abs().max()) <= 2e-4 * max(1e-3, float(g1[n].abs().max())), n
