#!/usr/bin/env python
"""Density-grid scoring (plot_model, reference evaluation/visualization/flow_plotting.py:33-53):
one nfn_chain_forward_grid launch vs one broadcast-y forward launch per grid line.

    python tools/grid_probe.py [--rows 65536] [--ny 256] [--config cfg2]

Per grid line the looped version re-reads the whole parameter tensor (4 P bytes per row); the grid
kernel stages each parameter tile once and only writes 4 bytes per (row, event).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from normalizingflownetwork_b200 import functional as F  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1 << 16)
    ap.add_argument("--ny", type=int, default=256)
    ap.add_argument("--config", default="cfg2")
    args = ap.parse_args()
    ft, d, tb, _, _ = bench.CONFIGS[args.config]
    P = bench.param_size(ft, d, tb)
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(22)
    t = torch.randn((args.rows, P), generator=g, device=dev) * 0.5
    yg = torch.randn((args.ny, d), generator=g, device=dev)

    def grid():
        return F.chain_forward_grid(t, yg, ft, d, tb)

    def looped():
        return torch.stack([F.chain_forward(t, yg[j:j + 1], ft, d, tb) for j in range(args.ny)])

    def timed(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    a, b = grid(), looped()
    print("bitwise equal:", bool(torch.equal(a, b)))
    ms_g, ms_l = timed(grid), timed(looped)
    pairs = args.rows * args.ny
    print("%s  rows %d x %d events (P = %d)" % (args.config, args.rows, args.ny, P))
    print("  grid kernel      %8.3f ms  %.3e pairs/s  (output %.1f GB/s)" % (ms_g, pairs / ms_g * 1e3, 4 * pairs / ms_g / 1e6))
    print("  one launch/line  %8.3f ms  %.3e pairs/s" % (ms_l, pairs / ms_l * 1e3))


if __name__ == "__main__":
    main()
