"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle.

Bars (BASELINE.json north_star; fp32 kernels vs float64 oracle on identical fp32 inputs):
  log-prob   |delta| <= 1e-5 * max(1, |logp|)
  gradients  |delta| <= 1e-4 * max(1, |g|)
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import analytic_np as an

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LOGP_RTOL = 1e-5
GRAD_RTOL = 1e-4

CONFIG_CHAINS = {
    "cfg1": (["radial"] * 3, 1, True),
    "cfg2": (["planar", "radial", "affine"] * 3 + ["planar"], 2, True),
    "cfg3": (["radial", "planar"] * 8, 4, True),
    "cfg4": (["radial"] * 5, 1, True),
}


def rel_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return np.abs(got - ref) / np.maximum(1.0, np.abs(ref))


def assert_logp(got, ref, tol=LOGP_RTOL, what=""):
    e = rel_err(got, ref)
    assert np.all(np.isfinite(np.asarray(got))), what
    assert e.max() <= tol, "%s logp rel err %.3e at %d" % (what, e.max(), int(e.argmax()))


def assert_grad(got, ref, tol=GRAD_RTOL, what=""):
    e = rel_err(got, ref)
    assert np.all(np.isfinite(np.asarray(got))), what
    assert e.max() <= tol, "%s grad rel err %.3e at %s" % (what, e.max(), np.unravel_index(e.argmax(), e.shape))


@pytest.fixture(params=["fast", "accurate"])
def math_mode(request, nfn_lib):
    from normalizingflownetwork_b200 import functional as F

    F.set_math_mode(request.param == "accurate")
    yield request.param
    F.set_math_mode(False)


@pytest.fixture(params=["specialized", "specialized-tma", "jit", "jit-tma", "generic"])
def kernel_path(request, nfn_lib):
    """specialized: ahead-of-time instance where one exists (else runtime-specialised);
    jit: every chain through the NVRTC runtime specialiser; generic: runtime-chain kernel.
    Both specialised paths exist in two generations: cp.async CTA tiles and bulk-copy / TMA warp tiles."""
    from normalizingflownetwork_b200 import functional as F

    F.set_option("force_generic", request.param == "generic")
    F.set_option("force_jit", request.param.startswith("jit"))
    F.set_option("chain_io", "tma" if request.param.endswith("-tma") else "cpasync")
    yield request.param
    F.set_option("force_generic", 0)
    F.set_option("force_jit", 0)
    F.set_option("chain_io", "auto")


def dev(x, device):
    return torch.tensor(np.asarray(x, dtype=np.float32), device=device)


# ----------------------------------------------------------------------------- golden fixtures
def test_known_answers_single_flow(cuda_device, nfn_lib):
    from normalizingflownetwork_b200 import FLOWS

    ka = json.load(open(os.path.join(GOLDEN, "known_answers.json")))
    for c in ka["single_flow"]:
        d = c["n_dims"]
        cls = FLOWS[c["flow"]]
        flow = cls(torch.ones((10, cls.get_param_size(d)), device=cuda_device), d)
        z = torch.full((10, d), c["z"], device=cuda_device)
        fwd = flow.forward(z).cpu().numpy()
        fldj = flow._forward_log_det_jacobian(z).cpu().numpy()
        assert fwd.shape == (10, d) and fldj.shape == (10,)
        np.testing.assert_allclose(fwd, np.tile(c["forward"], (10, 1)), rtol=2e-6, atol=2e-7)
        # the accurate path keeps tiny log-dets (4.26e-7 for planar d=4, t=1, z=1) relative-accurate
        np.testing.assert_allclose(fldj, c["fldj"], rtol=2e-3, atol=3e-7)
        # [1, d] z broadcasts against the batch (reference tests/test_flows.py:24-29)
        assert flow.forward(z[:1]).shape == (10, d)
        assert flow._forward_log_det_jacobian(z[:1]).shape == (10,)


def test_known_answers_layer(cuda_device, nfn_lib, math_mode, kernel_path):
    from normalizingflownetwork_b200 import InverseNormalizingFlowLayer

    ka = json.load(open(os.path.join(GOLDEN, "known_answers.json")))
    for c in ka["layer"]:
        layer = InverseNormalizingFlowLayer(c["flow_types"], c["n_dims"], c["trainable_base_dist"])
        P = layer.get_total_param_size()
        dist = layer(torch.full((3, P), c["t"], device=cuda_device))
        lp = dist.log_prob(torch.full((3, c["n_dims"]), c["y"], device=cuda_device)).cpu().numpy()
        assert_logp(lp, np.full(3, c["log_prob"]), what=str(c))


def test_golden_chain_vectors(cuda_device, nfn_lib, math_mode, kernel_path):
    from normalizingflownetwork_b200 import functional as F

    cases = json.load(open(os.path.join(GOLDEN, "chain_vectors.json")))
    for c in cases:
        ft, d, tb = c["flow_types"], c["n_dims"], c["trainable_base_dist"]
        t, y, up = dev(c["t"], cuda_device), dev(c["y"], cuda_device), dev(c["upstream"], cuda_device)
        if t.shape[1] == 0:
            continue
        tag = "%s sigma=%s %s/%s" % (c["name"], c["sigma"], math_mode, kernel_path)
        # at sigma=1 the planar determinant can get small: the bar is stated for sigma=0.5
        ltol = LOGP_RTOL if c["sigma"] <= 0.5 else 5e-5
        gtol = GRAD_RTOL if c["sigma"] <= 0.5 else 1e-3
        lp = F.chain_forward(t, y, ft, d, tb)
        assert_logp(lp.cpu().numpy(), c["log_prob"], ltol, tag)
        lp2, dt, dy = F.chain_forward_backward(t, y, ft, d, tb, g_logp=up, want_dy=True)
        assert torch.equal(lp, lp2), tag
        assert_grad(dt.cpu().numpy(), c["dt"], gtol, tag + " dt")
        assert_grad(dy.cpu().numpy(), c["dy"], gtol, tag + " dy")
        lb = F.chain_forward(t, y[3:4], ft, d, tb)
        assert_logp(lb.cpu().numpy(), c["log_prob_y_row3_broadcast"], ltol, tag + " bcast")


def test_cuda_vs_reference_code_run(cuda_device, nfn_lib, math_mode):
    """The CUDA path against what the reference's OWN flow / layer code returned for the same inputs
    (tests/golden/reference_run.json, produced by oracle/make_reference_run.py; see tests/test_reference_run.py)."""
    from normalizingflownetwork_b200 import functional as F

    ref = json.load(open(os.path.join(GOLDEN, "reference_run.json")))
    cases = json.load(open(os.path.join(GOLDEN, "chain_vectors.json")))
    mv = json.load(open(os.path.join(GOLDEN, "mixture_vectors.json")))
    n = 0
    for r, c in zip(ref["chains"], cases):
        assert (r["name"], r["sigma"]) == (c["name"], c["sigma"])
        ft, d, tb = c["flow_types"], c["n_dims"], c["trainable_base_dist"]
        t, y, up = dev(c["t"], cuda_device), dev(c["y"], cuda_device), dev(c["upstream"], cuda_device)
        if t.shape[1] == 0:
            continue
        tag = "reference run %s sigma=%s %s" % (c["name"], c["sigma"], math_mode)
        ltol = LOGP_RTOL if c["sigma"] <= 0.5 else 5e-5
        gtol = GRAD_RTOL if c["sigma"] <= 0.5 else 1e-3
        lp, dt, dy = F.chain_forward_backward(t, y, ft, d, tb, g_logp=up, want_dy=True)
        assert_logp(lp.cpu().numpy(), r["log_prob"], ltol, tag)
        assert_grad(dt.cpu().numpy(), np.asarray(r["dt"]).reshape(tuple(dt.shape)), gtol, tag + " dt")
        assert_grad(dy.cpu().numpy(), r["dy"], gtol, tag + " dy")
        lb = F.chain_forward(t, y[3:4], ft, d, tb)
        assert_logp(lb.cpu().numpy(), r["log_prob_y_row3_broadcast"], ltol, tag + " bcast")
        n += 1
    assert n >= 22
    for r, c in zip(ref["mixtures"]["mdn"], mv["mdn"]):
        assert (r["name"], r["sigma"]) == (c["name"], c["sigma"])
        t, y, up = dev(c["t"], cuda_device), dev(c["y"], cuda_device), dev(c["upstream"], cuda_device)
        lp, dt, dy = F.mdn_forward_backward(t, y, c["n_centers"], c["n_dims"], g_logp=up, want_dy=True)
        assert_logp(lp.cpu().numpy(), r["log_prob"], what=r["name"])
        assert_grad(dt.cpu().numpy(), r["dt"], what=r["name"] + " dt")
        assert_grad(dy.cpu().numpy(), r["dy"], what=r["name"] + " dy")
    for r, c in zip(ref["mixtures"]["kmn"], mv["kmn"]):
        assert r["name"] == c["name"]
        t, y, up = dev(c["t"], cuda_device), dev(c["y"], cuda_device), dev(c["upstream"], cuda_device)
        # the bandwidths the reference's scale_model returned (negative for init 0.3)
        locs, scales = dev(c["locs"], cuda_device), dev(r["scales"], cuda_device)
        lp, dt, dy, _ = F.kmn_forward_backward(t, y, locs, scales, g_logp=up, want_dy=True)
        assert_logp(lp.cpu().numpy(), r["log_prob"], what=r["name"])
        assert_grad(dt.cpu().numpy(), r["dt"], what=r["name"] + " dt")
        assert_grad(dy.cpu().numpy(), r["dy"], what=r["name"] + " dy")



# ----------------------------------------------------------------------------- seeded parity
@pytest.mark.parametrize("cfg", sorted(CONFIG_CHAINS))
@pytest.mark.parametrize("B", [1, 127, 129, 4096 + 37])
def test_chain_vs_oracle_ragged(cuda_device, nfn_lib, math_mode, kernel_path, cfg, B):
    from normalizingflownetwork_b200 import functional as F

    ft, d, tb = CONFIG_CHAINS[cfg]
    rng = np.random.default_rng(22 + B)
    P = an.layout(ft, d, tb)[1]
    t = rng.normal(0, 0.5, (B, P)).astype(np.float32)
    y = rng.normal(0, 1.0, (B, d)).astype(np.float32)
    up = rng.normal(0, 1.0, (B,)).astype(np.float32)
    ref_lp, ref_dt, ref_dy = an.chain_forward_backward(t, y, ft, d, tb, upstream=up * 0.5)
    lp, dt, dy = F.chain_forward_backward(dev(t, cuda_device), dev(y, cuda_device), ft, d, tb,
                                          g_logp=dev(up, cuda_device), g_scale=0.5, want_dy=True)
    tag = "%s B=%d %s/%s" % (cfg, B, math_mode, kernel_path)
    assert_logp(lp.cpu().numpy(), ref_lp, what=tag)
    assert_grad(dt.cpu().numpy(), ref_dt, what=tag + " dt")
    assert_grad(dy.cpu().numpy(), ref_dy, what=tag + " dy")
    assert_logp(F.chain_forward(dev(t, cuda_device), dev(y, cuda_device), ft, d, tb).cpu().numpy(), ref_lp,
                what=tag + " fwd")


@pytest.mark.parametrize("cfg", sorted(CONFIG_CHAINS))
@pytest.mark.parametrize("B,n_y", [(1, 1), (129, 7), (1000, 33)])
def test_chain_grid_vs_oracle(cuda_device, nfn_lib, math_mode, kernel_path, cfg, B, n_y):
    """Outer-product density grid (plot_model, evaluation/visualization/flow_plotting.py:33-53):
    logp[j, b] must equal the broadcast-y log_prob of event j, which the oracle computes row by row."""
    from normalizingflownetwork_b200 import functional as F

    ft, d, tb = CONFIG_CHAINS[cfg]
    rng = np.random.default_rng(1000 + B + n_y)
    P = an.layout(ft, d, tb)[1]
    t = rng.normal(0, 0.5, (B, P)).astype(np.float32)
    yg = rng.normal(0, 1.5, (n_y, d)).astype(np.float32)
    got = F.chain_forward_grid(dev(t, cuda_device), dev(yg, cuda_device), ft, d, tb).cpu().numpy()
    assert got.shape == (n_y, B)
    tag = "%s grid B=%d n_y=%d %s/%s" % (cfg, B, n_y, math_mode, kernel_path)
    for j in range(n_y):
        ref = an.chain_forward_backward(t, yg[j:j + 1], ft, d, tb, need_grad=False)
        assert_logp(got[j], ref, what=tag + " event %d" % j)
    # the grid entry and the broadcast-y entry run the same arithmetic
    one = F.chain_forward(dev(t, cuda_device), dev(yg[:1], cuda_device), ft, d, tb).cpu().numpy()
    np.testing.assert_array_equal(one, got[0])


@pytest.mark.parametrize("cfg", sorted(CONFIG_CHAINS))
def test_chain_vs_oracle_2p16(cuda_device, nfn_lib, cfg):
    """2^16 rows, t ~ N(0, 0.5^2), y ~ N(0, 1), seed 22 -- the BASELINE.md parity protocol."""
    from normalizingflownetwork_b200 import functional as F

    ft, d, tb = CONFIG_CHAINS[cfg]
    B = 1 << 16
    rng = np.random.default_rng(22)
    P = an.layout(ft, d, tb)[1]
    t = rng.normal(0, 0.5, (B, P)).astype(np.float32)
    y = rng.normal(0, 1.0, (B, d)).astype(np.float32)
    ref_lp, ref_dt, _ = an.chain_forward_backward(t, y, ft, d, tb, upstream=-1.0)
    lsum = torch.zeros(1, dtype=torch.float64, device=cuda_device)
    col = torch.zeros(P, dtype=torch.float64, device=cuda_device)
    lp, dt, _ = F.chain_forward_backward(dev(t, cuda_device), dev(y, cuda_device), ft, d, tb, g_scale=-1.0,
                                         logp_sum=lsum, dt_colsum=col)
    assert_logp(lp.cpu().numpy(), ref_lp, what=cfg)
    assert_grad(dt.cpu().numpy(), ref_dt, what=cfg)
    assert abs(lsum.item() - ref_lp.sum()) <= 1e-5 * np.abs(ref_lp).sum()
    np.testing.assert_allclose(col.cpu().numpy(), ref_dt.sum(0), rtol=1e-3, atol=1e-2)


def test_empty_and_zero_flow_chains(cuda_device, nfn_lib, kernel_path):
    from normalizingflownetwork_b200 import functional as F

    # B = 0
    lp = F.chain_forward(torch.zeros((0, 11), device=cuda_device), torch.zeros((0, 1), device=cuda_device),
                         ["radial"] * 3, 1, True)
    assert lp.shape == (0,)
    # n_flows = 0 with trainable base (reference tests/test_bayesian_estimator.py:45-61)
    rng = np.random.default_rng(3)
    t = rng.normal(0, 1, (300, 4)).astype(np.float32)
    y = rng.normal(0, 1, (300, 2)).astype(np.float32)
    ref_lp, ref_dt, ref_dy = an.chain_forward_backward(t, y, [], 2, True)
    lp, dt, dy = F.chain_forward_backward(dev(t, cuda_device), dev(y, cuda_device), [], 2, True, want_dy=True)
    assert_logp(lp.cpu().numpy(), ref_lp)
    assert_grad(dt.cpu().numpy(), ref_dt)
    assert_grad(dy.cpu().numpy(), ref_dy)
    # n_flows = 0 without base parameters: P = 0, standard normal
    y1 = rng.normal(0, 1, (257, 3)).astype(np.float32)
    lp = F.chain_forward(torch.zeros((257, 0), device=cuda_device), dev(y1, cuda_device), [], 3, False)
    ref = -0.5 * (y1.astype(np.float64) ** 2).sum(1) - 1.5 * np.log(2 * np.pi)
    assert_logp(lp.cpu().numpy(), ref)


def test_long_generic_chain_max_dims(cuda_device, nfn_lib):
    """K = 64 flows, d = 8: the descriptor's maximum, served by the generic kernel."""
    from normalizingflownetwork_b200 import functional as F

    ft = (["planar", "radial", "affine", "radial"] * 16)[:64]
    d, tb = 8, True
    assert not F.chain_is_specialized(ft, d, tb)
    rng = np.random.default_rng(5)
    P = an.layout(ft, d, tb)[1]
    t = rng.normal(0, 0.3, (513, P)).astype(np.float32)
    y = rng.normal(0, 1.0, (513, d)).astype(np.float32)
    ref_lp, ref_dt, ref_dy = an.chain_forward_backward(t, y, ft, d, tb)
    lp, dt, dy = F.chain_forward_backward(dev(t, cuda_device), dev(y, cuda_device), ft, d, tb, want_dy=True)
    assert_logp(lp.cpu().numpy(), ref_lp, tol=5e-5)
    assert_grad(dt.cpu().numpy(), ref_dt, tol=1e-3)
    assert_grad(dy.cpu().numpy(), ref_dy, tol=1e-3)


def test_runtime_specialiser_serves_unlisted_chains(cuda_device, nfn_lib):
    """A chain without an ahead-of-time instance is compiled at first use (NVRTC) and cached."""
    from normalizingflownetwork_b200 import functional as F

    ft, d, tb = ["affine", "planar", "radial", "planar", "affine", "radial"], 3, True
    assert not F.chain_is_specialized(ft, d, tb)
    rng = np.random.default_rng(11)
    P = an.layout(ft, d, tb)[1]
    B = 10_000 + 3
    t = rng.normal(0, 0.5, (B, P)).astype(np.float32)
    y = rng.normal(0, 1.0, (B, d)).astype(np.float32)
    ref_lp, ref_dt, ref_dy = an.chain_forward_backward(t, y, ft, d, tb, upstream=-1.0)
    before = nfn_lib.nfn_jit_cache_size()
    lp, dt, dy = F.chain_forward_backward(dev(t, cuda_device), dev(y, cuda_device), ft, d, tb, g_scale=-1.0,
                                          want_dy=True)
    assert nfn_lib.nfn_jit_cache_size() == before + 1, "the chain should have been JIT-specialised, not run generically"
    assert_logp(lp.cpu().numpy(), ref_lp)
    assert_grad(dt.cpu().numpy(), ref_dt)
    assert_grad(dy.cpu().numpy(), ref_dy)
    F.chain_forward(dev(t, cuda_device), dev(y, cuda_device), ft, d, tb)
    assert nfn_lib.nfn_jit_cache_size() == before + 1  # second call hits the cache
    # NFN_B200_JIT=0 sends the same chain to the generic kernel; results agree within tolerance
    F.set_option("jit", 0)
    try:
        lp_g, dt_g, _ = F.chain_forward_backward(dev(t, cuda_device), dev(y, cuda_device), ft, d, tb, g_scale=-1.0)
    finally:
        F.set_option("jit", 1)
    assert_logp(lp_g.cpu().numpy(), ref_lp)
    assert torch.allclose(lp, lp_g, rtol=1e-5, atol=1e-5)
    assert torch.allclose(dt, dt_g, rtol=1e-3, atol=1e-4)


def test_row_independence_bit_exact(cuda_device, nfn_lib):
    """Permuting rows permutes outputs bit-for-bit (reference tests/test_flows.py:31-41)."""
    from normalizingflownetwork_b200 import functional as F

    ft, d, tb = CONFIG_CHAINS["cfg2"]
    g = torch.Generator(device="cpu").manual_seed(22)
    B = 5000
    t = (torch.randn((B, 48), generator=g) * 0.5).to(cuda_device)
    y = torch.randn((B, d), generator=g).to(cuda_device)
    perm = torch.randperm(B, generator=g).to(cuda_device)
    lp, dt, _ = F.chain_forward_backward(t, y, ft, d, tb)
    lp2, dt2, _ = F.chain_forward_backward(t[perm].contiguous(), y[perm].contiguous(), ft, d, tb)
    assert torch.equal(lp[perm], lp2)
    assert torch.equal(dt[perm], dt2)


def test_autograd_function_matches_fused(cuda_device, nfn_lib):
    from normalizingflownetwork_b200 import InverseNormalizingFlowLayer

    ft, d, tb = CONFIG_CHAINS["cfg2"]
    rng = np.random.default_rng(9)
    t = dev(rng.normal(0, 0.5, (1000, 48)), cuda_device).requires_grad_(True)
    y = dev(rng.normal(0, 1.0, (1000, d)), cuda_device).requires_grad_(True)
    dist = InverseNormalizingFlowLayer(ft, d, tb)(t)
    loss = -dist.log_prob(y).mean()
    loss.backward()
    ref_lp, ref_dt, ref_dy = an.chain_forward_backward(t.detach().cpu().numpy(), y.detach().cpu().numpy(), ft, d,
                                                       tb, upstream=-1.0 / 1000)
    assert abs(loss.item() + ref_lp.mean()) < 1e-5 * max(1.0, abs(ref_lp.mean()))
    assert_grad(t.grad.cpu().numpy() * 1000, ref_dt * 1000)
    assert_grad(y.grad.cpu().numpy() * 1000, ref_dy * 1000)


def test_host_pipeline_matches_device(cuda_device, nfn_lib):
    import ctypes

    from normalizingflownetwork_b200 import _lib
    from normalizingflownetwork_b200 import functional as F

    ft, d, tb = CONFIG_CHAINS["cfg2"]
    B, P = 300_001, 48  # > 3 chunks of the pipeline, ragged tail
    g = torch.Generator(device="cpu").manual_seed(22)
    t = (torch.randn((B, P), generator=g) * 0.5).pin_memory()
    y = torch.randn((B, d), generator=g).pin_memory()
    logp = torch.empty(B).pin_memory()
    dt = torch.empty((B, P)).pin_memory()
    col = torch.empty(P, dtype=torch.float64)
    lsum = ctypes.c_double(0.0)
    desc = _lib.make_desc(ft, d, tb)
    with torch.cuda.device(cuda_device):
        _lib.check(nfn_lib.nfn_chain_forward_backward_host(
            ctypes.byref(desc), _lib.ptr(t), _lib.ptr(y), B, None, ctypes.c_float(-1.0 / B), _lib.ptr(logp),
            _lib.ptr(dt), ctypes.byref(lsum), _lib.ptr(col), B))
    lp_d, dt_d, _ = F.chain_forward_backward(t.to(cuda_device), y.to(cuda_device), ft, d, tb, g_scale=-1.0 / B)
    assert torch.equal(logp, lp_d.cpu())
    assert torch.equal(dt, dt_d.cpu())
    assert abs(lsum.value - lp_d.double().sum().item()) < 1e-6 * B
    np.testing.assert_allclose(col.numpy(), dt_d.sum(0).cpu().numpy(), rtol=1e-3, atol=1e-5)
    logp2 = torch.empty(B).pin_memory()
    with torch.cuda.device(cuda_device):
        _lib.check(nfn_lib.nfn_chain_forward_host(ctypes.byref(desc), _lib.ptr(t), _lib.ptr(y), B,
                                                  _lib.ptr(logp2), B))
        _lib.check(nfn_lib.nfn_host_release())
    assert torch.equal(logp2, lp_d.cpu())


def test_full_size_properties_cfg2(cuda_device, nfn_lib):
    """BASELINE config 2 at full size (2^20 rows): size-independent properties."""
    from normalizingflownetwork_b200 import functional as F

    ft, d, tb = CONFIG_CHAINS["cfg2"]
    B, P = 1 << 20, 48
    g = torch.Generator(device=cuda_device).manual_seed(22)
    t = torch.randn((B, P), generator=g, device=cuda_device) * 0.5
    y = torch.randn((B, d), generator=g, device=cuda_device)
    lsum = torch.zeros(1, dtype=torch.float64, device=cuda_device)
    col = torch.zeros(P, dtype=torch.float64, device=cuda_device)
    lp, dt, _ = F.chain_forward_backward(t, y, ft, d, tb, g_scale=-1.0 / B, logp_sum=lsum, dt_colsum=col)
    assert torch.isfinite(lp).all() and torch.isfinite(dt).all()
    assert abs(lsum.item() - lp.double().sum().item()) < 1e-7 * B
    np.testing.assert_allclose(col.cpu().numpy(), dt.double().sum(0).cpu().numpy(), rtol=2e-3, atol=1e-6)
    # linearity in the cotangent (exact up to denormal intermediates: sech^2 can underflow)
    _, dt2, _ = F.chain_forward_backward(t, y, ft, d, tb, g_scale=-2.0 / B)
    assert torch.allclose(dt2, dt * 2, rtol=1e-5, atol=1e-12)
    # forward-only kernel and fused kernel agree bit for bit on logp
    assert torch.equal(F.chain_forward(t, y, ft, d, tb), lp)
    # a random sample of rows against the float64 oracle
    idx = torch.randint(0, B, (4096,), generator=g, device=cuda_device)
    ref_lp, ref_dt, _ = an.chain_forward_backward(t[idx].cpu().numpy(), y[idx].cpu().numpy(), ft, d, tb,
                                                  upstream=-1.0)
    assert_logp(lp[idx].cpu().numpy(), ref_lp)
    assert_grad((dt[idx] * B).cpu().numpy(), ref_dt)


def test_full_size_properties_cfg3_forward(cuda_device, nfn_lib):
    """BASELINE config 3 at its per-GPU size (2^23 events x 128 parameters, 4.3 GB, forward only):
    finite, a row shard scored alone equals its slice of the whole bit for bit (what lets ranks shard rows with
    no exchange), one event broadcast against a shard, and a random sample of rows against the float64 oracle."""
    from normalizingflownetwork_b200 import functional as F

    ft, d, tb = CONFIG_CHAINS["cfg3"]
    B, P = 1 << 23, 128
    g = torch.Generator(device=cuda_device).manual_seed(22)
    t = torch.randn((B, P), generator=g, device=cuda_device)
    t.mul_(0.5)
    y = torch.randn((B, d), generator=g, device=cuda_device)
    lp = F.chain_forward(t, y, ft, d, tb)
    assert lp.shape == (B,) and torch.isfinite(lp).all()
    lo, hi = (1 << 22) + 4096, (1 << 22) + 4096 + 100_003  # ragged shard length
    assert torch.equal(F.chain_forward(t[lo:hi], y[lo:hi], ft, d, tb), lp[lo:hi])
    lb = F.chain_forward(t[lo:lo + 4096], y[lo + 7:lo + 8], ft, d, tb)
    idx = torch.randint(0, B, (4096,), generator=g, device=cuda_device)
    ref = an.chain_forward_backward(t[idx].cpu().numpy(), y[idx].cpu().numpy(), ft, d, tb, need_grad=False)
    assert_logp(lp[idx].cpu().numpy(), ref)
    ref_b = an.chain_forward_backward(t[lo:lo + 4096].cpu().numpy(), y[lo + 7:lo + 8].cpu().numpy(), ft, d, tb,
                                      need_grad=False)
    assert_logp(lb.cpu().numpy(), ref_b)


def test_full_size_properties_cfg4(cuda_device, nfn_lib):
    """BASELINE config 4 at its per-GPU size (2^20 rows = S * B folded draws, 5 radial flows, d = 1, fwd+bwd)."""
    from normalizingflownetwork_b200 import functional as F

    ft, d, tb = CONFIG_CHAINS["cfg4"]
    B, P = 1 << 20, 17
    g = torch.Generator(device=cuda_device).manual_seed(22)
    t = torch.randn((B, P), generator=g, device=cuda_device) * 0.5
    y = torch.randn((B, d), generator=g, device=cuda_device)
    lsum = torch.zeros(1, dtype=torch.float64, device=cuda_device)
    col = torch.zeros(P, dtype=torch.float64, device=cuda_device)
    lp, dt, _ = F.chain_forward_backward(t, y, ft, d, tb, g_scale=-1.0 / B, logp_sum=lsum, dt_colsum=col)
    assert torch.isfinite(lp).all() and torch.isfinite(dt).all()
    assert abs(lsum.item() - lp.double().sum().item()) < 1e-7 * B
    np.testing.assert_allclose(col.cpu().numpy(), dt.double().sum(0).cpu().numpy(), rtol=2e-3, atol=1e-6)
    assert torch.allclose(F.chain_forward(t, y, ft, d, tb), lp, rtol=1e-6, atol=1e-6)
    # the [S, B] -> [B] posterior-predictive epilogue over the folded draws (S = 32 draws of 2^15 rows)
    S = 32
    lme = F.logmeanexp_draws(lp.view(S, B // S))
    want = torch.logsumexp(lp.view(S, B // S).double(), 0) - np.log(S)
    assert torch.allclose(lme.double(), want, rtol=1e-5, atol=1e-5)
    idx = torch.randint(0, B, (4096,), generator=g, device=cuda_device)
    ref_lp, ref_dt, _ = an.chain_forward_backward(t[idx].cpu().numpy(), y[idx].cpu().numpy(), ft, d, tb,
                                                  upstream=-1.0)
    assert_logp(lp[idx].cpu().numpy(), ref_lp)
    assert_grad((dt[idx] * B).cpu().numpy(), ref_dt)


def test_full_size_properties_cfg5_mdn(cuda_device, nfn_lib):
    """BASELINE config 5 at full size (MDN, 20 components, d = 2, P = 100, 2^22 rows, fwd+bwd)."""
    from normalizingflownetwork_b200 import functional as F

    K, d = 20, 2
    B, P = 1 << 22, 100
    g = torch.Generator(device=cuda_device).manual_seed(22)
    t = torch.randn((B, P), generator=g, device=cuda_device) * 0.5
    y = torch.randn((B, d), generator=g, device=cuda_device)
    lsum = torch.zeros(1, dtype=torch.float64, device=cuda_device)
    lp, dt, _ = F.mdn_forward_backward(t, y, K, d, g_scale=-1.0 / B, logp_sum=lsum)
    assert torch.isfinite(lp).all() and torch.isfinite(dt).all()
    assert abs(lsum.item() - lp.double().sum().item()) < 1e-7 * B
    # the mixture-weight gradients of a row sum to zero (softmax): a checksum over the whole batch
    assert float(dt[:, 2 * K * d:].double().sum(1).abs().max()) * B < 1e-4
    lo, hi = (1 << 21) + 1024, (1 << 21) + 1024 + 50_001
    assert torch.allclose(F.mdn_forward(t[lo:hi], y[lo:hi], K, d), lp[lo:hi], rtol=1e-6, atol=1e-6)
    idx = torch.randint(0, B, (4096,), generator=g, device=cuda_device)
    ref_lp, ref_dt, _ = an.mdn_forward_backward(t[idx].cpu().numpy(), y[idx].cpu().numpy(), K, d, upstream=-1.0)
    assert_logp(lp[idx].cpu().numpy(), ref_lp)
    assert_grad((dt[idx] * B).cpu().numpy(), ref_dt)


# ----------------------------------------------------------------------------- mixture heads
def test_golden_mixture_vectors(cuda_device, nfn_lib, math_mode):
    from normalizingflownetwork_b200 import functional as F

    mv = json.load(open(os.path.join(GOLDEN, "mixture_vectors.json")))
    for c in mv["mdn"]:
        K, d = c["n_centers"], c["n_dims"]
        t, y, up = dev(c["t"], cuda_device), dev(c["y"], cuda_device), dev(c["upstream"], cuda_device)
        tag = "%s sigma=%s %s" % (c["name"], c["sigma"], math_mode)
        assert_logp(F.mdn_forward(t, y, K, d).cpu().numpy(), c["log_prob"], what=tag)
        lp, dt, dy = F.mdn_forward_backward(t, y, K, d, g_logp=up, want_dy=True)
        assert_logp(lp.cpu().numpy(), c["log_prob"], what=tag)
        assert_grad(dt.cpu().numpy(), c["dt"], what=tag + " dt")
        assert_grad(dy.cpu().numpy(), c["dy"], what=tag + " dy")
    for c in mv["kmn"]:
        t, y, up = dev(c["t"], cuda_device), dev(c["y"], cuda_device), dev(c["upstream"], cuda_device)
        locs, scales = dev(c["locs"], cuda_device), dev(c["scales"], cuda_device)
        tag = "%s %s" % (c["name"], math_mode)
        assert_logp(F.kmn_forward(t, y, locs, scales).cpu().numpy(), c["log_prob"], what=tag)
        lp, dt, dy, dsc = F.kmn_forward_backward(t, y, locs, scales, g_logp=up, want_dy=True)
        assert_logp(lp.cpu().numpy(), c["log_prob"], what=tag)
        assert_grad(dt.cpu().numpy(), c["dt"], what=tag + " dt")
        assert_grad(dy.cpu().numpy(), c["dy"], what=tag + " dy")
        assert_grad(dsc.cpu().numpy(), c["dscales"], what=tag + " dscales")


@pytest.mark.parametrize("K,d,B", [(20, 2, 1 << 15), (3, 1, 1000), (5, 5, 4097), (7, 3, 129), (1, 1, 5)])
def test_mdn_vs_oracle(cuda_device, nfn_lib, math_mode, K, d, B):
    from normalizingflownetwork_b200 import functional as F

    rng = np.random.default_rng(22)
    P = 2 * K * d + K
    t = rng.normal(0, 0.5, (B, P)).astype(np.float32)
    y = rng.normal(0, 1.0, (B, d)).astype(np.float32)
    ref_lp, ref_dt, ref_dy = an.mdn_forward_backward(t, y, K, d, upstream=-1.0)
    lsum = torch.zeros(1, dtype=torch.float64, device=cuda_device)
    col = torch.zeros(P, dtype=torch.float64, device=cuda_device)
    lp, dt, dy = F.mdn_forward_backward(dev(t, cuda_device), dev(y, cuda_device), K, d, g_scale=-1.0, want_dy=True,
                                        logp_sum=lsum, dt_colsum=col)
    assert_logp(lp.cpu().numpy(), ref_lp)
    assert_grad(dt.cpu().numpy(), ref_dt)
    assert_grad(dy.cpu().numpy(), ref_dy)
    assert abs(lsum.item() - ref_lp.sum()) <= 1e-5 * np.abs(ref_lp).sum()
    np.testing.assert_allclose(col.cpu().numpy(), ref_dt.sum(0), rtol=1e-3, atol=1e-2)
    assert_logp(F.mdn_forward(dev(t, cuda_device), dev(y[:1], cuda_device), K, d).cpu().numpy(),
                an.mdn_forward_backward(t, y[:1], K, d, need_grad=False))


@pytest.mark.parametrize("M,d,B", [(20, 1, 5000), (60, 2, 1 << 14), (100, 3, 257)])
def test_kmn_vs_oracle(cuda_device, nfn_lib, math_mode, M, d, B):
    from normalizingflownetwork_b200 import functional as F

    rng = np.random.default_rng(22)
    t = rng.normal(0, 1.0, (B, M)).astype(np.float32)
    y = rng.normal(0, 1.0, (B, d)).astype(np.float32)
    locs = rng.normal(0, 1.0, (M, d)).astype(np.float32)
    scales = np.where(np.arange(M) < M // 2, -0.357, 0.45).astype(np.float32)  # negative bandwidth, App. B.7
    up = rng.normal(0, 1.0, (B,)).astype(np.float32)
    ref_lp, ref_dt, ref_ds, ref_dy = an.kmn_forward_backward(t, y, locs, scales, upstream=up)
    lp, dt, dy, dsc = F.kmn_forward_backward(dev(t, cuda_device), dev(y, cuda_device), dev(locs, cuda_device),
                                             dev(scales, cuda_device), g_logp=dev(up, cuda_device), want_dy=True)
    assert_logp(lp.cpu().numpy(), ref_lp)
    assert_grad(dt.cpu().numpy(), ref_dt)
    assert_grad(dy.cpu().numpy(), ref_dy)
    np.testing.assert_allclose(dsc.cpu().numpy(), ref_ds, rtol=2e-3, atol=2e-3 * np.abs(ref_ds).max())


def test_logmeanexp_draws(cuda_device, nfn_lib):
    from normalizingflownetwork_b200 import functional as F

    rng = np.random.default_rng(22)
    x = rng.normal(-3, 2, (50, 10_001)).astype(np.float32)
    ref = np.log(np.mean(np.exp(x.astype(np.float64)), axis=0))
    out = F.logmeanexp_draws(dev(x, cuda_device)).cpu().numpy()
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-5)


def test_errors_are_loud(cuda_device, nfn_lib):
    from normalizingflownetwork_b200 import InverseNormalizingFlowLayer
    from normalizingflownetwork_b200 import functional as F

    with pytest.raises(AssertionError):
        F.chain_forward(torch.zeros((4, 12), device=cuda_device), torch.zeros((4, 1), device=cuda_device),
                        ["radial"] * 3, 1, True)
    with pytest.raises(RuntimeError):  # CPU tensors are rejected: no fallback
        InverseNormalizingFlowLayer(["radial"], 1, False)(torch.zeros((4, 3))).log_prob(torch.zeros((4, 1)))


def test_outputs_stay_inside_their_buffers(cuda_device, nfn_lib, kernel_path):
    """compute-sanitizer is closed on this pool, so out-of-bounds writes are hunted with canaries:
    every output lives inside a larger sentinel-filled allocation that must come back untouched
    (ragged row counts exercise the partial last tile of every kernel path)."""
    import ctypes

    from normalizingflownetwork_b200 import _lib
    from normalizingflownetwork_b200 import functional as F

    SENT = 12345.0
    PAD = 1024  # floats on each side

    def guarded(n):
        buf = torch.full((n + 2 * PAD,), SENT, device=cuda_device)
        return buf, buf[PAD: PAD + n]

    def intact(buf, n):
        return bool((buf[:PAD] == SENT).all()) and bool((buf[PAD + n:] == SENT).all())

    g = torch.Generator(device=cuda_device).manual_seed(4)
    for ft, d, tb in [CONFIG_CHAINS["cfg2"], CONFIG_CHAINS["cfg4"], (["affine", "planar"], 3, False)]:
        desc = _lib.make_desc(ft, d, tb)
        P = nfn_lib.nfn_chain_param_size(ctypes.byref(desc))
        for B in (1, 127, 128, 129, 1000):
            t = torch.randn((B, P), generator=g, device=cuda_device) * 0.5
            y = torch.randn((B, d), generator=g, device=cuda_device)
            b_lp, lp = guarded(B)
            b_dt, dt = guarded(B * P)
            b_dy, dy = guarded(B * d)
            _lib.check(nfn_lib.nfn_chain_forward_backward(
                ctypes.byref(desc), _lib.ptr(t), _lib.ptr(y), B, None, ctypes.c_float(1.0), _lib.ptr(lp),
                _lib.ptr(dt), _lib.ptr(dy), None, None, B, _lib.current_stream(cuda_device)))
            torch.cuda.synchronize()
            assert intact(b_lp, B) and intact(b_dt, B * P) and intact(b_dy, B * d), (ft, B)
            assert torch.isfinite(lp).all() and torch.isfinite(dt).all()
            b_lp2, lp2 = guarded(B)
            _lib.check(nfn_lib.nfn_chain_forward(ctypes.byref(desc), _lib.ptr(t), _lib.ptr(y), B, _lib.ptr(lp2), B,
                                                 _lib.current_stream(cuda_device)))
            torch.cuda.synchronize()
            assert intact(b_lp2, B) and torch.equal(lp, lp2)
    # fused dense kernel and MDN head (not path-dependent, run once)
    if kernel_path == "specialized":
        ft, d, tb = CONFIG_CHAINS["cfg2"]
        desc = _lib.make_desc(ft, d, tb)
        H, P = 16, 48
        for impl, B in [(i, b) for i in ("sync", "tc5") for b in (1, 33, 129, 1000, 128 * 300 + 5)]:
            F.set_option("dense_mma", impl)   # both GEMM implementations (warp-level mma.sync; tcgen05 / TMEM)
            h = torch.tanh(torch.randn((B, H), generator=g, device=cuda_device))
            W = torch.randn((H, P), generator=g, device=cuda_device) * 0.1
            bias = torch.zeros(P, device=cuda_device)
            y = torch.randn((B, d), generator=g, device=cuda_device)
            b_lp, lp = guarded(B)
            b_dh, dh = guarded(B * H)
            b_dw, dW = guarded(H * P)
            b_db, db = guarded(P)
            dW.zero_()
            db.zero_()
            _lib.check(nfn_lib.nfn_dense_chain_forward_backward(
                ctypes.byref(desc), H, _lib.ptr(h), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(y), B, None,
                ctypes.c_float(1.0), _lib.ptr(lp), _lib.ptr(dh), _lib.ptr(dW), _lib.ptr(db), None, B,
                _lib.current_stream(cuda_device)))
            torch.cuda.synchronize()
            assert intact(b_lp, B) and intact(b_dh, B * H) and intact(b_dw, H * P) and intact(b_db, P), (impl, B)
            assert torch.isfinite(lp).all() and torch.isfinite(dh).all() and torch.isfinite(dW).all(), (impl, B)
        F.set_option("dense_mma", "auto")
        K, dm = 20, 2
        Pm = 2 * K * dm + K
        for B in (1, 129, 1000):
            t = torch.randn((B, Pm), generator=g, device=cuda_device) * 0.5
            y = torch.randn((B, dm), generator=g, device=cuda_device)
            b_lp, lp = guarded(B)
            b_dt, dt = guarded(B * Pm)
            _lib.check(nfn_lib.nfn_mdn_forward_backward(K, dm, _lib.ptr(t), _lib.ptr(y), B, None, ctypes.c_float(1.0),
                                                        _lib.ptr(lp), _lib.ptr(dt), None, None, None, B,
                                                        _lib.current_stream(cuda_device)))
            torch.cuda.synchronize()
            assert intact(b_lp, B) and intact(b_dt, B * Pm), B


@pytest.mark.parametrize("cfg", ["cfg2", "cfg3"])
def test_accuracy_tail_at_2e18_rows(cuda_device, nfn_lib, cfg):
    """How many of 2^18 synthetic rows miss the bars (1e-5 log-prob, 1e-4 gradient), for t ~ N(0, 0.5^2) -- the
    distribution the bars are stated for -- and for the harder sigma = 1 that SURVEY §7 warns
    about.  Measured (profiles/r02_sigma1_tail.md): sigma = 0.5: no row beyond 1e-5 in log-prob, <= 0.001 % of rows
    beyond 1e-4 in the gradient; sigma = 1: 0.002 % (cfg2) / 0.024 % (cfg3) beyond 1e-5, nothing beyond 1e-3 in
    log-prob.  The tail is a property of float32 STATE, not of one sub-expression: a near-singular planar or
    radial step amplifies the rounding already carried by z, so the accurate math mode shortens it only a little.
    This test pins those fractions (with head-room) so that a kernel change cannot silently fatten the tail."""
    from normalizingflownetwork_b200 import functional as F

    ft, d, tb = CONFIG_CHAINS[cfg]
    P = F.chain_param_size(ft, d, tb)
    B = 1 << 18
    limits = {0.5: (0.0, 0.0, 0.004, 0.0), 1.0: (0.06, 0.002, 0.8, 0.05)}   # % of rows: lp>1e-5, lp>1e-4, dt>1e-4, dt>1e-3
    for sigma, (l5, l4, g4, g3) in limits.items():
        rng = np.random.default_rng(22)
        t = rng.normal(0, sigma, (B, P)).astype(np.float32)
        y = rng.normal(0, 1.0, (B, d)).astype(np.float32)
        ref_lp, ref_dt, _ = an.chain_forward_backward(t, y, ft, d, tb, upstream=1.0)
        lp, dt, _ = F.chain_forward_backward(dev(t, cuda_device), dev(y, cuda_device), ft, d, tb)
        e_lp = rel_err(lp.cpu().numpy(), ref_lp)
        e_dt = rel_err(dt.cpu().numpy(), ref_dt).max(1)
        frac = [100.0 * float(np.mean(e > thr)) for e, thr in ((e_lp, 1e-5), (e_lp, 1e-4), (e_dt, 1e-4), (e_dt, 1e-3))]
        assert frac[0] <= l5 and frac[1] <= l4 and frac[2] <= g4 and frac[3] <= g3, (cfg, sigma, frac)
