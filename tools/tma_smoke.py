"""Hang / correctness guard for the bulk-copy (TMA) warp-tile chain kernels: every case is run through both
kernel generations (cp.async CTA tiles, the round-1 kernels, and the warp-tile kernels) and compared.
Run under `timeout` on the GPU box before the test-suite is pointed at the new kernels."""
import sys

import torch

sys.path.insert(0, ".")
from normalizingflownetwork_b200 import functional as F  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(5)
CHAINS = [
    (["radial"] * 5, 1, True),                                   # P = 17: linear layout, ragged last tile path
    (["planar", "radial", "affine"] * 3 + ["planar"], 2, True),  # P = 48: 64B swizzle
    (["radial", "planar"] * 8, 4, True),                         # P = 128: 128B swizzle
    (["radial"] * 3, 1, True),                                   # P = 11
    (["radial"] * 10, 2, True),                                  # P = 44: linear, V = 4
    (["affine", "planar"], 3, False),                            # P = 13 (runtime-specialised)
    (["radial", "planar"], 2, True),                             # P = 13
    (["planar"] * 4, 1, False),                                  # P = 12
    (["radial"] * 2 + ["affine"], 2, True),                      # P = 16 (JIT): 64B swizzle
    (["affine"] * 2, 2, True),                                   # P = 12
    (["affine"] * 5, 2, True),                                   # P = 24 (JIT): 32B swizzle
]
worst = 0.0
for ft, d, tb in CHAINS:
    P = F.chain_param_size(ft, d, tb)
    for B in (1, 31, 32, 33, 127, 1000, 4096, 100_003):
        t = torch.randn((B, P), generator=g, device=dev) * 0.5
        y = torch.randn((B, d), generator=g, device=dev)
        up = torch.randn(B, generator=g, device=dev)
        res = {}
        for io in ("cpasync", "tma"):
            F.set_option("chain_io", io)
            lp = F.chain_forward(t, y, ft, d, tb)
            col = torch.zeros(P, dtype=torch.float64, device=dev)
            ls = torch.zeros(1, dtype=torch.float64, device=dev)
            lp2, dt, dy = F.chain_forward_backward(t, y, ft, d, tb, g_logp=up, g_scale=-1.0 / B, want_dy=True,
                                                   logp_sum=ls, dt_colsum=col)
            lpb = F.chain_forward(t, y[:1], ft, d, tb)
            torch.cuda.synchronize()
            res[io] = (lp, lp2, dt, dy, col, ls, lpb)
        a, b = res["cpasync"], res["tma"]
        for i, name in enumerate(("logp", "logp(bwd)", "dt", "dy", "colsum", "logp_sum", "logp(bcast)")):
            x, z = a[i].double(), b[i].double()
            err = float((x - z).abs().max() / max(1.0, float(x.abs().max())))
            tol = 1e-5 if name == "colsum" else 2e-6
            worst = max(worst, err)
            if not (err <= tol) or not bool(torch.isfinite(z).all()):
                print("MISMATCH", ft, d, tb, "B=%d" % B, name, err)
                sys.exit(1)
        # the sums must match the tensors they summarise
        assert abs(float(b[5]) - float(b[1].double().sum())) <= 1e-6 * max(1.0, abs(float(b[5])))
    print("ok  P=%3d d=%d K=%2d" % (P, d, len(ft)), flush=True)
F.set_option("chain_io", "auto")
print("tma_smoke ok, worst relative difference between the two generations: %.2e" % worst)
