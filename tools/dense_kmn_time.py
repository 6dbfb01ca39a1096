"""Time the fused Dense(P)+KMN kernel at the reference's default KernelMixtureNetwork shape (50 centres x 2 bandwidths
= 100 kernels, d = 1, H = 16, 2^22 rows) next to torch's GEMMs around the streaming KMN kernel.

    python tools/dense_kmn_time.py [--steps 20] [--rows 4194304]
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
    ap.add_argument("--kernels", type=int, default=100)
    ap.add_argument("--dims", type=int, default=1)
    ap.add_argument("--hidden", type=int, default=16)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, M, d, H = a.rows, a.kernels, a.dims, a.hidden
    g = torch.Generator(device=dev).manual_seed(5)
    h = torch.tanh(torch.randn((B, H), generator=g, device=dev))
    W = torch.randn((H, M), generator=g, device=dev) * 0.3
    b = torch.randn(M, generator=g, device=dev) * 0.1
    y = torch.randn((B, d), generator=g, device=dev)
    locs = torch.randn((M, d), generator=g, device=dev)
    scales = torch.rand(M, generator=g, device=dev) * 0.5 + 0.3
    dW, db = torch.zeros((H, M), device=dev), torch.zeros(M, device=dev)
    fused = timed(lambda: F.dense_kmn_forward_backward(h, W, b, y, locs, scales, g_scale=-1.0 / B, dW=dW, dbias=db), a.steps)
    fused_f = timed(lambda: F.dense_kmn_forward(h, W, b, y, locs, scales), a.steps)
    t = torch.addmm(b, h, W)
    head = timed(lambda: F.kmn_forward_backward(t, y, locs, scales, g_scale=-1.0 / B), a.steps)

    def unfused():
        tt = torch.addmm(b, h, W)
        _, dt, _, dsc = F.kmn_forward_backward(tt, y, locs, scales, g_scale=-1.0 / B)
        return dt @ W.t(), h.t() @ dt, dt.sum(0), dsc

    unf = timed(unfused, max(3, a.steps // 2))
    unf_f = timed(lambda: F.kmn_forward(torch.addmm(b, h, W), y, locs, scales), max(3, a.steps // 2))
    print("KMN M=%d d=%d H=%d rows=%d" % (M, d, H, B))
    print("  fwd+bwd: fused %8.1f us | streaming head alone %8.1f us | unfused layer + head (torch GEMMs) %8.1f us  -> %.2fx"
          % (fused, head, unf, unf / fused))
    print("  forward: fused %8.1f us | unfused layer + head (torch GEMM) %8.1f us  -> %.2fx" % (fused_f, unf_f, unf_f / fused_f))
    print("  bytes/row fused %d vs streaming head %d" % (4 * (2 * H + d + 1), 4 * (2 * M + d + 1)))


if __name__ == "__main__":
    main()
