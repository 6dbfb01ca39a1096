"""sigma = 1 accuracy tail (SURVEY.md §7 hard part 1): 2^18 rows of t ~ N(0, 1), y ~ N(0, 1)
for the cfg2 and cfg3 chains; fraction of rows whose log-prob / gradient error against the float64 oracle exceeds
1e-5 / 1e-4 (relative to max(1, |ref|)), fast and accurate math.  Prints a markdown table."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from normalizingflownetwork_b200 import functional as F  # noqa: E402
from oracle import analytic_np as an  # noqa: E402

dev = torch.device("cuda:0")
CHAINS = {"cfg2": (["planar", "radial", "affine"] * 3 + ["planar"], 2, True), "cfg3": (["radial", "planar"] * 8, 4, True)}
B = 1 << 18
print("| chain | sigma | math | logp: max rel err | rows > 1e-5 | rows > 1e-4 | dt: max rel err | rows > 1e-4 | rows > 1e-3 |")
print("|---|---|---|---|---|---|---|---|---|")
for name, (ft, d, tb) in CHAINS.items():
    P = F.chain_param_size(ft, d, tb)
    for sigma in (0.5, 1.0):
        rng = np.random.default_rng(22)
        t = rng.normal(0, sigma, (B, P)).astype(np.float32)
        y = rng.normal(0, 1.0, (B, d)).astype(np.float32)
        ref_lp, ref_dt, _ = an.chain_forward_backward(t, y, ft, d, tb, upstream=1.0)
        td, yd = torch.tensor(t, device=dev), torch.tensor(y, device=dev)
        for mode in ("fast", "accurate"):
            F.set_math_mode(mode == "accurate")
            lp, dt, _ = F.chain_forward_backward(td, yd, ft, d, tb)
            lp, dt = lp.cpu().numpy().astype(np.float64), dt.cpu().numpy().astype(np.float64)
            e_lp = np.abs(lp - ref_lp) / np.maximum(1.0, np.abs(ref_lp))
            e_dt = (np.abs(dt - ref_dt) / np.maximum(1.0, np.abs(ref_dt))).max(1)
            fin = np.isfinite(e_lp) & np.isfinite(e_dt)
            print("| %s | %.1f | %s | %.2e | %.4f %% | %.4f %% | %.2e | %.4f %% | %.4f %% |" % (
                name, sigma, mode, np.nanmax(e_lp[fin]), 100 * np.mean(e_lp > 1e-5), 100 * np.mean(e_lp > 1e-4),
                np.nanmax(e_dt[fin]), 100 * np.mean(e_dt > 1e-4), 100 * np.mean(e_dt > 1e-3)), flush=True)
        F.set_math_mode(False)
