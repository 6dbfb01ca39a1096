"""Accuracy of the CUDA path next to the reference's own float32 evaluation (runs on the GPU box).

For every golden chain / MDN case: max relative error (|delta| / max(1, |ref|)) against the float64 run of the
reference's own code (tests/golden/reference_run.json) of
  * the reference's code executed in float32, its working precision (tests/golden/reference_run_f32.json),
  * the CUDA kernels in fast and in accurate math mode.
Writes a markdown table to gpurun_out/accuracy_report.md.

    gpurun -- 'python tools/accuracy_report.py'
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64).reshape(a.shape)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))


def main():
    from normalizingflownetwork_b200 import functional as F

    dev = torch.device("cuda:0")
    to = lambda a: torch.tensor(np.asarray(a, dtype=np.float32), device=dev)
    ref, f32 = load("reference_run.json"), load("reference_run_f32.json")
    cases, mv = load("chain_vectors.json"), load("mixture_vectors.json")
    rows = []
    for kind, refs, f32s, ins in (("chain", ref["chains"], f32["chains"], cases),
                                  ("mdn", ref["mixtures"]["mdn"], f32["mdn"], mv["mdn"])):
        for r, h, c in zip(refs, f32s, ins):
            assert (r["name"], r["sigma"]) == (h["name"], h["sigma"]) == (c["name"], c["sigma"])
            t, y, up = to(c["t"]), to(c["y"]), to(c["upstream"])
            row = [kind, c["name"], c["sigma"], rel(h["log_prob"], r["log_prob"]), rel(h["dt"], r["dt"])]
            for accurate in (False, True):
                F.set_math_mode(accurate)
                if kind == "chain":
                    lp, dt, _ = F.chain_forward_backward(t, y, c["flow_types"], c["n_dims"], c["trainable_base_dist"],
                                                         g_logp=up, want_dy=True)
                else:
                    lp, dt, _ = F.mdn_forward_backward(t, y, c["n_centers"], c["n_dims"], g_logp=up, want_dy=True)
                row += [rel(lp.cpu().numpy(), r["log_prob"]), rel(dt.cpu().numpy(), r["dt"])]
            F.set_math_mode(False)
            rows.append(row)
    lines = [
        "# Accuracy against the float64 run of the reference's own code (tools/accuracy_report.py, B200)",
        "",
        "max over the 12 rows of each golden case of |delta| / max(1, |ref|); `ref fp32` = the reference's own flow / layer",
        "code executed in float32 (its working precision) on the CPU stand-ins; bars: log-prob 1e-5, gradients 1e-4",
        "(sigma 0.5), 5e-5 / 1e-3 (sigma 1.0).",
        "",
        "| head | case | sigma | ref fp32 logp | ref fp32 dt | CUDA fast logp | CUDA fast dt | CUDA accurate logp | CUDA accurate dt |",
        "|---|---|---|---|---|---|---|---|---|",
    ]
    for r in rows:
        lines.append("| %s | %s | %s | " % tuple(r[:3]) + " | ".join("%.1e" % v for v in r[3:]) + " |")
    worst = np.max(np.asarray([r[3:] for r in rows], dtype=np.float64), axis=0)
    lines.append("| | **worst** | | " + " | ".join("**%.1e**" % v for v in worst) + " |")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "accuracy_report.md"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[-4:]))


if __name__ == "__main__":
    main()
