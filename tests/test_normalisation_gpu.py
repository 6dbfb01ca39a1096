"""GPU: a first-principles check that needs no oracle at all -- what the CUDA kernels return is a probability
density in y.  exp(log_prob) from the flow-chain head integrates to 1 for arbitrary parameter rows (1-D events,
trapezoid rule over the density-grid entry point), also through the fused y pipeline in DATA units
(`pdf` = exp(log_prob((y - mean) / std) - sum log std), reference estimators/BaseEstimator.py:71-75), and so do the
MDN / KMN mixture heads.  A wrong chain direction, log-det sign or Jacobian shift fails this at once.
The CPU twin for the oracles is tests/test_oracle.py::test_oracle_density_is_normalised_*."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N_Y = 240001
L = 60.0


def _integral(p, y):
    """Trapezoid rule in float64: p [n_y, B] on the grid y [n_y] -> [B]."""
    p = p.double()
    h = (y[1:] - y[:-1]).double().unsqueeze(1)
    return (0.5 * (p[1:] + p[:-1]) * h).sum(0)


@pytest.mark.parametrize("ft", [["radial"] * 3, ["radial"] * 5, ["planar", "radial", "affine"], ["planar"] * 4],
                         ids=lambda f: "".join(x[0] for x in f))
def test_chain_density_integrates_to_one(cuda_device, nfn_lib, ft):
    from normalizingflownetwork_b200 import functional as F

    d, tb, B = 1, True, 5
    P = F.chain_param_size(ft, d, tb)
    rng = np.random.default_rng(len(ft))
    t = torch.tensor(rng.normal(0.0, 0.5, (B, P)).astype(np.float32), device=cuda_device)
    y = torch.linspace(-L, L, N_Y, device=cuda_device, dtype=torch.float64)
    yg = y.float().unsqueeze(1).contiguous()
    lp = F.chain_forward_grid(t, yg, ft, d, tb)                      # [n_y, B]
    assert lp.shape == (N_Y, B)
    total = _integral(torch.exp(lp.double()), yg[:, 0])
    assert float((total - 1.0).abs().max()) <= 1e-4, total.tolist()   # float32 log-probs; a misread sign or direction is off by O(1)
    # data units: the kernel normalises y on load, shifts by -log std and exponentiates (the estimator's pdf)
    mean, std = [1.7], [3.0]
    xf = F.make_xform(d, mean, std, logp_shift=-math.log(std[0]), exp_out=True)
    pdf = F.chain_forward_grid(t, yg, ft, d, tb, xform=xf)
    total = _integral(pdf, yg[:, 0])
    assert float((total - 1.0).abs().max()) <= 1e-4, total.tolist()


def test_mixture_heads_integrate_to_one(cuda_device, nfn_lib):
    from normalizingflownetwork_b200 import functional as F

    rng = np.random.default_rng(4)
    n = 120001
    y = torch.linspace(-L, L, n, device=cuda_device, dtype=torch.float64)
    yg = y.float().unsqueeze(1).contiguous()
    K = 5
    t = torch.tensor(rng.normal(0.0, 0.7, (1, 3 * K)).astype(np.float32), device=cuda_device)
    lp = F.mdn_forward(t.expand(n, 3 * K).contiguous(), yg, K, 1)
    assert abs(float(_integral(torch.exp(lp.double()).unsqueeze(1), yg[:, 0])) - 1.0) <= 1e-4
    M = 12
    locs = torch.tensor(rng.normal(0.0, 2.0, (M, 1)).astype(np.float32), device=cuda_device)
    scales = torch.tensor(rng.normal(0.0, 0.6, M).astype(np.float32), device=cuda_device)   # some negative: legal
    assert bool((scales < 0).any())
    logits = torch.tensor(rng.normal(0.0, 1.0, (1, M)).astype(np.float32), device=cuda_device)
    lp = F.kmn_forward(logits.expand(n, M).contiguous(), yg, locs, scales)
    assert abs(float(_integral(torch.exp(lp.double()).unsqueeze(1), yg[:, 0])) - 1.0) <= 1e-4


def test_config2_chain_density_integrates_to_one_2d(cuda_device, nfn_lib):
    """BASELINE config 2's chain (10 flows, 2-D events): the headline kernel's density over a 1201 x 1201 grid,
    one launch of the density-grid entry point per parameter row pair (same rows as the CPU twin in test_oracle.py)."""
    from normalizingflownetwork_b200 import functional as F

    ft, d, tb = ["planar", "radial", "affine"] * 3 + ["planar"], 2, True
    P = F.chain_param_size(ft, d, tb)
    t = np.concatenate([np.random.default_rng(seed).normal(0.0, 0.3, (1, P)) for seed in (0, 2)]).astype(np.float32)
    g = torch.linspace(-30.0, 30.0, 1201, device=cuda_device, dtype=torch.float64).float()
    Y = torch.stack(torch.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2).contiguous()
    lp = F.chain_forward_grid(torch.tensor(t, device=cuda_device), Y, ft, d, tb)      # [1201 * 1201, 2]
    p = torch.exp(lp.double()).reshape(g.numel(), g.numel(), 2)
    h = (g[1:] - g[:-1]).double()
    inner = (0.5 * (p[:, 1:] + p[:, :-1]) * h[None, :, None]).sum(1)                   # over the second coordinate
    total = (0.5 * (inner[1:] + inner[:-1]) * h[:, None]).sum(0)
    assert float((total - 1.0).abs().max()) <= 1e-3, total.tolist()
