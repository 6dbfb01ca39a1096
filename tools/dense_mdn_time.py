"""Time the fused Dense(P)+MDN kernel at BASELINE config 5's shape (K = 20, d = 2, P = 100, H = 16, 2^22 rows) next
to what it replaces: torch's GEMMs (t = h W + b; dh = dt W^T; dW = h^T dt; db = sum dt) around the streaming MDN kernel.

    python tools/dense_mdn_time.py [--steps 20] [--rows 4194304]
"""
import argparse
import os
import sys

sys.path.insert(0, os.getcwd())
import torch  # noqa: E402

from normalizingflownetwork_b200 import functional as F  # noqa: E402


def timed(fn, steps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--rows", type=int, default=1 << 22)
    ap.add_argument("--centers", type=int, default=20)
    ap.add_argument("--dims", type=int, default=2)
    ap.add_argument("--hidden", type=int, default=16)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, K, d, H = a.rows, a.centers, a.dims, a.hidden
    P = K * (2 * d + 1)
    g = torch.Generator(device=dev).manual_seed(5)
    h = torch.tanh(torch.randn((B, H), generator=g, device=dev))
    W = torch.randn((H, P), generator=g, device=dev) * 0.3
    b = torch.randn(P, generator=g, device=dev) * 0.1
    y = torch.randn((B, d), generator=g, device=dev)
    dW, db = torch.zeros((H, P), device=dev), torch.zeros(P, device=dev)

    fused = timed(lambda: F.dense_mdn_forward_backward(h, W, b, y, K, d, g_scale=-1.0 / B, dW=dW, dbias=db), a.steps)
    fused_f = timed(lambda: F.dense_mdn_forward(h, W, b, y, K, d), a.steps)
    t = torch.addmm(b, h, W)
    head = timed(lambda: F.mdn_forward_backward(t, y, K, d, g_scale=-1.0 / B), a.steps)
    head_f = timed(lambda: F.mdn_forward(t, y, K, d), a.steps)

    def unfused():
        tt = torch.addmm(b, h, W)
        _, dt, _ = F.mdn_forward_backward(tt, y, K, d, g_scale=-1.0 / B)
        dh = dt @ W.t()
        dWu = h.t() @ dt
        dbu = dt.sum(0)
        return dh, dWu, dbu

    unf = timed(unfused, max(3, a.steps // 2))
    unf_f = timed(lambda: F.mdn_forward(torch.addmm(b, h, W), y, K, d), max(3, a.steps // 2))
    peak = 6550.4
    print("MDN K=%d d=%d P=%d H=%d rows=%d" % (K, d, P, H, B))
    print("  fwd+bwd: fused %8.1f us | streaming head alone %8.1f us | unfused layer + head (torch GEMMs) %8.1f us  -> %.2fx"
          % (fused, head, unf, unf / fused))
    print("  forward: fused %8.1f us | streaming head alone %8.1f us | unfused layer + head (torch GEMM)  %8.1f us  -> %.2fx"
          % (fused_f, head_f, unf_f, unf_f / fused_f))
    print("  bytes/row fused %d (%.2f of %.0f GB/s) vs streaming head %d" % (
        4 * (2 * H + d + 1), 4 * (2 * H + d + 1) * B / (fused * 1e-6) / 1e9 / peak, peak, 4 * (2 * P + d + 1)))


if __name__ == "__main__":
    main()
