"""GPU: the estimators' y pipeline fused into the heads (reference
estimators/BaseEstimator.py:55-86): normalisation on load, in-kernel Philox noise, the -sum log y_std Jacobian
on the way out, and exp() for pdf -- against the unfused composition of the same kernels, and the fused input
normalisation of the first Dense layer (MaximumLikelihoodNNEstimator.py:40)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG2 = (["planar", "radial", "affine"] * 3 + ["planar"], 2, True)
CFG4 = (["radial"] * 5, 1, True)


# ----------------------------------------------------------------------------- numpy restatement of the noise
def philox4x32_10(c, k):
    """c: uint32 [..., 4], k: uint32 [..., 2] -> uint32 [..., 4] (Salmon et al., the constants curand uses)."""
    c = c.astype(np.uint64).copy()
    k = k.astype(np.uint64).copy()
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    W0, W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[..., 0], M1 * c[..., 2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        n0 = hi1 ^ c[..., 1] ^ k[..., 0]
        n2 = hi0 ^ c[..., 3] ^ k[..., 1]
        c = np.stack([n0, lo1, n2, lo0], axis=-1)
        k = np.stack([(k[..., 0] + W0) & mask, (k[..., 1] + W1) & mask], axis=-1)
    return c.astype(np.uint32)


def row_normals(B, d, seed, offset):
    """What csrc/nfn_chain_kernel.cuh::row_normals draws for rows 0..B-1 (float32 Box-Muller on Philox words)."""
    out = np.zeros((B, d), np.float32)
    r = np.arange(B, dtype=np.uint64)
    for b in range((d + 3) // 4):
        c = np.stack([r & np.uint64(0xFFFFFFFF), r >> np.uint64(32),
                      np.full(B, offset & 0xFFFFFFFF, np.uint64),
                      np.full(B, ((offset >> 32) ^ (b << 28)) & 0xFFFFFFFF, np.uint64)], axis=-1)
        k = np.stack([np.full(B, seed & 0xFFFFFFFF, np.uint64), np.full(B, seed >> 32, np.uint64)], axis=-1)
        w = philox4x32_10(c, k).astype(np.float32)
        for h in range(2):
            u1 = (w[:, 2 * h] * np.float32(2.3283064365386963e-10) + np.float32(1.1641532182693481e-10)).astype(np.float32)
            u2 = (w[:, 2 * h + 1] * np.float32(2.3283064365386963e-10)).astype(np.float32)
            rad = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
            ang = (np.float32(2.0) * u2).astype(np.float64) * np.pi
            if 4 * b + 2 * h < d:
                out[:, 4 * b + 2 * h] = rad * np.cos(ang).astype(np.float32)
            if 4 * b + 2 * h + 1 < d:
                out[:, 4 * b + 2 * h + 1] = rad * np.sin(ang).astype(np.float32)
    return out


def test_philox_known_answer():
    # Random123 known-answer test: counter = key = 0 and the all-ones / pi-digits vectors
    z = philox4x32_10(np.zeros((1, 4), np.uint32), np.zeros((1, 2), np.uint32))[0]
    assert [hex(int(v)) for v in z] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    o = philox4x32_10(np.full((1, 4), 0xFFFFFFFF, np.uint32), np.full((1, 2), 0xFFFFFFFF, np.uint32))[0]
    assert [hex(int(v)) for v in o] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]


def _data(B, P, d, device, seed=3):
    g = torch.Generator(device=device).manual_seed(seed)
    t = torch.randn((B, P), generator=g, device=device) * 0.5
    y_raw = torch.randn((B, d), generator=g, device=device) * torch.tensor([2.0, 0.5, 1.5, 3.0][:d], device=device) + 1.25
    return t, y_raw


@pytest.mark.parametrize("io", ["cpasync", "tma", "generic"])
@pytest.mark.parametrize("chain", [CFG2, CFG4, (["radial", "planar"] * 8, 4, True)])
def test_chain_xform_equals_unfused_composition(cuda_device, nfn_lib, io, chain):
    from normalizingflownetwork_b200 import functional as F

    ft, d, tb = chain
    P = F.chain_param_size(ft, d, tb)
    B = 4099
    t, y_raw = _data(B, P, d, cuda_device)
    mean = [0.3, -1.0, 0.7, 0.1][:d]
    std = [1.7, 0.6, 2.2, 0.9][:d]
    shift = -float(np.sum(np.log(np.asarray(std, np.float32))))
    y_n = (y_raw - torch.tensor(mean, device=cuda_device)) / torch.tensor(std, device=cuda_device)
    F.set_option("force_generic", io == "generic")
    F.set_option("chain_io", "tma" if io == "tma" else "cpasync")
    try:
        ref = F.chain_forward(t, y_n, ft, d, tb)
        xf = F.make_xform(d, mean, std, logp_shift=shift)
        got = F.chain_forward(t, y_raw, ft, d, tb, xform=xf)
        assert torch.allclose(got, ref + shift, rtol=0, atol=2e-6 * float(ref.abs().max()))
        dens = F.chain_forward(t, y_raw, ft, d, tb, xform=F.make_xform(d, mean, std, logp_shift=shift, exp_out=True))
        assert torch.allclose(dens, torch.exp(ref + shift), rtol=2e-5, atol=1e-30)
        # fused forward + reverse sweep: same gradients, logp_sum carries the shift
        ls0 = torch.zeros(1, dtype=torch.float64, device=cuda_device)
        ls1 = torch.zeros(1, dtype=torch.float64, device=cuda_device)
        lp0, dt0, dy0 = F.chain_forward_backward(t, y_n, ft, d, tb, g_scale=-1.0 / B, want_dy=True, logp_sum=ls0)
        lp1, dt1, dy1 = F.chain_forward_backward(t, y_raw, ft, d, tb, g_scale=-1.0 / B, want_dy=True, logp_sum=ls1, xform=xf)
        assert torch.allclose(dt1, dt0, rtol=1e-5, atol=1e-9) and torch.allclose(dy1, dy0, rtol=1e-5, atol=1e-9)
        assert float(ls1) == pytest.approx(float(ls0) + B * shift, rel=1e-6)
        # density grid (plot_model): every parameter row against every normalised grid event
        if io != "generic":
            yg = y_raw[:7].contiguous()
            gr = F.chain_forward_grid(t[:300].contiguous(), yg, ft, d, tb,
                                      xform=F.make_xform(d, mean, std, logp_shift=shift, exp_out=True))
            gr_ref = torch.exp(F.chain_forward_grid(t[:300].contiguous(), y_n[:7].contiguous(), ft, d, tb) + shift)
            assert torch.allclose(gr, gr_ref, rtol=2e-5, atol=1e-30)
    finally:
        F.set_option("force_generic", 0)
        F.set_option("chain_io", "auto")


@pytest.mark.parametrize("chain", [CFG2, CFG4])
def test_in_kernel_noise_is_the_documented_philox_stream(cuda_device, nfn_lib, chain):
    """y' = (y - mean) / std + noise_std * n(row, offset): n restated in numpy (Philox4x32-10 + Box-Muller)."""
    from normalizingflownetwork_b200 import functional as F

    ft, d, tb = chain
    P = F.chain_param_size(ft, d, tb)
    B = 1000
    t, y_raw = _data(B, P, d, cuda_device, seed=8)
    mean, std = [0.2, -0.4][:d], [1.3, 0.8][:d]
    seed, offset, sigma = 22, 5, 0.35
    ctr = torch.tensor([3], dtype=torch.int64, device=cuda_device)   # device-side part of the offset
    n = row_normals(B, d, seed, offset + 3)
    y_n = ((y_raw.cpu().numpy() - np.asarray(mean, np.float32)) / np.asarray(std, np.float32) + np.float32(sigma) * n)
    ref = F.chain_forward(t, torch.tensor(y_n.astype(np.float32), device=cuda_device), ft, d, tb)
    xf = F.make_xform(d, mean, std, noise_std=sigma, seed=seed, offset=offset, offset_dev=ctr)
    for io in ("cpasync", "tma"):
        F.set_option("chain_io", io)
        got = F.chain_forward(t, y_raw, ft, d, tb, xform=xf)
        # the noise itself agrees to float32 rounding of log / sincospi; log-probs to the usual bar
        assert torch.allclose(got, ref, rtol=0, atol=2e-4), float((got - ref).abs().max())
        lp, dt, _ = F.chain_forward_backward(t, y_raw, ft, d, tb, xform=xf)
        assert torch.allclose(lp, got, rtol=0, atol=1e-6)
    F.set_option("chain_io", "auto")
    # statistics of a large draw, and a different offset gives a different stream
    big = row_normals(200_000, 2, 7, 0)
    assert abs(big.mean()) < 0.01 and abs(big.std() - 1.0) < 0.01 and abs(np.corrcoef(big.T)[0, 1]) < 0.01
    assert not np.allclose(row_normals(100, 2, 7, 1), big[:100])


def test_mixture_heads_xform(cuda_device, nfn_lib):
    from normalizingflownetwork_b200 import functional as F

    B, d, K = 3001, 2, 20
    g = torch.Generator(device=cuda_device).manual_seed(11)
    mean, std = [0.5, -2.0], [1.5, 0.7]
    shift = -float(np.sum(np.log(np.asarray(std, np.float32))))
    y_raw = torch.randn((B, d), generator=g, device=cuda_device) * 2 + 0.5
    y_n = (y_raw - torch.tensor(mean, device=cuda_device)) / torch.tensor(std, device=cuda_device)
    xf = F.make_xform(d, mean, std, logp_shift=shift)
    t = torch.randn((B, 2 * K * d + K), generator=g, device=cuda_device) * 0.5
    ref = F.mdn_forward(t, y_n, K, d)
    assert torch.allclose(F.mdn_forward(t, y_raw, K, d, xform=xf), ref + shift, rtol=0, atol=1e-5)
    lp0, dt0, _ = F.mdn_forward_backward(t, y_n, K, d, g_scale=-1.0 / B)
    lp1, dt1, _ = F.mdn_forward_backward(t, y_raw, K, d, g_scale=-1.0 / B, xform=xf)
    assert torch.allclose(lp1, lp0 + shift, rtol=0, atol=1e-5) and torch.allclose(dt1, dt0, rtol=1e-5, atol=1e-10)
    locs = torch.randn((12, d), generator=g, device=cuda_device)
    scales = torch.tensor([0.3] * 6 + [-0.7] * 6, device=cuda_device)
    tk = torch.randn((B, 12), generator=g, device=cuda_device)
    refk = F.kmn_forward(tk, y_n, locs, scales)
    assert torch.allclose(F.kmn_forward(tk, y_raw, locs, scales, xform=xf), refk + shift, rtol=0, atol=1e-5)
    dens = F.kmn_forward(tk, y_raw, locs, scales, xform=F.make_xform(d, mean, std, logp_shift=shift, exp_out=True))
    assert torch.allclose(dens, torch.exp(refk + shift), rtol=2e-5, atol=1e-30)


@pytest.mark.parametrize("impl", ["sync", "tc5"])
def test_dense_chain_xform(cuda_device, nfn_lib, impl):
    from normalizingflownetwork_b200 import functional as F

    ft, d, tb = CFG2
    P, H, B = 48, 16, 70_001
    g = torch.Generator(device=cuda_device).manual_seed(12)
    h = torch.tanh(torch.randn((B, H), generator=g, device=cuda_device))
    W = torch.randn((H, P), generator=g, device=cuda_device) * 0.1
    b = torch.randn(P, generator=g, device=cuda_device) * 0.1
    mean, std = [0.5, -2.0], [1.5, 0.7]
    shift = -float(np.sum(np.log(np.asarray(std, np.float32))))
    y_raw = torch.randn((B, d), generator=g, device=cuda_device) * 2 + 0.5
    y_n = (y_raw - torch.tensor(mean, device=cuda_device)) / torch.tensor(std, device=cuda_device)
    xf = F.make_xform(d, mean, std, logp_shift=shift)
    F.set_option("dense_mma", impl)
    try:
        ref = F.dense_chain_forward(h, W, b, y_n, ft, d, tb)
        got = F.dense_chain_forward(h, W, b, y_raw, ft, d, tb, xform=xf)
        assert torch.allclose(got, ref + shift, rtol=0, atol=1e-5)
        lp0, dh0, dW0, db0 = F.dense_chain_forward_backward(h, W, b, y_n, ft, d, tb, g_scale=-1.0 / B)
        lp1, dh1, dW1, db1 = F.dense_chain_forward_backward(h, W, b, y_raw, ft, d, tb, g_scale=-1.0 / B, xform=xf)
        assert torch.allclose(dh1, dh0, rtol=1e-5, atol=1e-10)
        assert torch.allclose(dW1, dW0, rtol=1e-4, atol=1e-6) and torch.allclose(db1, db0, rtol=1e-4, atol=1e-6)
    finally:
        F.set_option("dense_mma", "auto")


def test_first_layer_normalises_on_load(cuda_device, nfn_lib):
    from normalizingflownetwork_b200 import functional as F

    g = torch.Generator(device=cuda_device).manual_seed(13)
    for K, N, act in ((1, 16, "tanh"), (3, 16, "relu"), (4, 32, "tanh")):
        B = 5003
        x = torch.randn((B, K), generator=g, device=cuda_device) * 3 + 1
        w = torch.randn((N, K), generator=g, device=cuda_device)
        b = torch.randn(N, generator=g, device=cuda_device)
        xm = torch.randn(K, generator=g, device=cuda_device)
        xs = torch.rand(K, generator=g, device=cuda_device) + 0.5
        assert F.dense_act_xnorm_supported(K, N, act)
        xn = (x - xm) / (xs + 1e-8)
        ref = F.dense_act_forward(xn, w, b, act)
        got = F.dense_act_forward(x, w, b, act, xm, xs)
        assert torch.equal(got, ref)
        dout = torch.randn((B, N), generator=g, device=cuda_device)
        _, dW0, db0 = F.dense_act_backward(xn, ref, dout, w, act, need_dx=False)
        _, dW1, db1 = F.dense_act_backward(x, ref, dout, w, act, need_dx=False, x_mean=xm, x_std=xs)
        # (sums of ~5000 float32 products in atomics' arrival order: compare at the scale of the largest entry)
        assert float((dW1 - dW0).abs().max()) <= 1e-5 * float(dW0.abs().max())
        assert float((db1 - db0).abs().max()) <= 1e-5 * max(1.0, float(db0.abs().max()))


def _count_kernels(fn):
    from torch.profiler import ProfilerActivity, profile

    fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    names = [e.name for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower()
             and "memset" not in e.name.lower()]
    return names


def test_log_pdf_is_three_launches(cuda_device, nfn_lib):
    """config 1 (2048 rows, MLP (16, 16) tanh, 3 radial flows): x normalisation rides in the first layer's
    kernel, y normalisation and the Jacobian in the head's -- log_pdf / pdf are 3 kernels, eagerly."""
    from normalizingflownetwork_b200.estimators import NormalizingFlowNetwork
    from normalizingflownetwork_b200.simulation import gen_cosine_noise_data

    x, y = gen_cosine_noise_data(2048, noise_std=0.3, heterosced_noise=0.5)
    model = NormalizingFlowNetwork.build_function(n_dims=1, n_flows=3, hidden_sizes=(16, 16), activation="tanh")
    model.fit(x, y, batch_size=512, epochs=2, verbose=0)
    xd, yd = model._to_dev(x), model._to_dev(y)
    names = _count_kernels(lambda: model.log_pdf(xd, yd))
    assert len(names) <= 3, names
    names = _count_kernels(lambda: model.pdf(xd, yd))
    assert len(names) <= 3, names
    # and the fused pipeline equals the composed one
    with torch.no_grad():
        t = model.params_from_x(xd)
        y_n = (yd - model.y_mean) / model.y_std
        ref = model.dist_layer(t).log_prob(y_n) - torch.sum(torch.log(model.y_std))
    assert torch.allclose(model.log_pdf(xd, yd), ref, rtol=0, atol=2e-5)
    assert torch.allclose(model.pdf(xd, yd), torch.exp(ref), rtol=3e-5, atol=1e-30)
