#!/usr/bin/env python
"""Experiment: host-buffer end-to-end pass, chunked copy-engine pipeline vs zero-copy.

    python tools/zero_copy_probe.py [--config cfg2] [--steps 10]

(a) nfn_chain_forward_backward_host: H2D -> kernel -> D2H chunks on 3 streams (copy engines);
(b) the DEVICE entry point handed the pinned host pointers themselves (UVA): the kernel's
    cp.async reads of t and its 128-bit stores of dt cross PCIe directly, reads and writes
    overlap from the first tile and nothing is staged in HBM.
Prints rows/s and effective PCIe GB/s per direction for both.
"""
import argparse
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from normalizingflownetwork_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--rows", type=int, default=0)
    args = ap.parse_args()
    lib = _lib.load()
    ft, d, tb, B, bwd = bench.CONFIGS[args.config]
    if args.rows:
        B = args.rows
    P = bench.param_size(ft, d, tb)
    desc = _lib.make_desc(ft, d, tb)
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(22)
    h_t = (torch.randn((B, P), generator=g) * 0.5).pin_memory()
    h_y = torch.randn((B, d), generator=g).pin_memory()
    h_logp = torch.empty(B).pin_memory()
    h_dt = torch.empty((B, P)).pin_memory()
    h_logp2 = torch.empty(B).pin_memory()
    h_dt2 = torch.empty((B, P)).pin_memory()
    h_sum = ctypes.c_double(0.0)
    d_sum = torch.zeros(1, dtype=torch.float64, device=dev)
    gs = ctypes.c_float(-1.0 / B)
    stream = _lib.current_stream(dev)

    def pipe():
        _lib.check(lib.nfn_chain_forward_backward_host(ctypes.byref(desc), _lib.ptr(h_t), _lib.ptr(h_y), B, None, gs,
                                                       _lib.ptr(h_logp), _lib.ptr(h_dt), ctypes.byref(h_sum), None, B))

    def zero_copy():
        _lib.check(lib.nfn_chain_forward_backward(ctypes.byref(desc), _lib.ptr(h_t), _lib.ptr(h_y), B, None, gs,
                                                  _lib.ptr(h_logp2), _lib.ptr(h_dt2), None,
                                                  ctypes.c_void_p(d_sum.data_ptr()), None, B, stream))
        torch.cuda.synchronize()

    def timed(fn, label):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        torch.cuda.synchronize()
        el = (time.perf_counter() - t0) / args.steps
        h2d = 4 * B * (P + d)
        d2h = 4 * B * (1 + P)
        print("%-28s %8.3f ms/step  %.3e rows/s  H2D %.1f GB/s  D2H %.1f GB/s" % (
            label, el * 1e3, B / el, h2d / el / 1e9, d2h / el / 1e9), flush=True)
        return el

    for mb in os.environ.get("PROBE_CHUNK_MB", "16").split(","):
        os.environ["NFN_B200_HOST_CHUNK_MB"] = mb
        lib.nfn_host_release()
        timed(pipe, "copy-engine pipeline %sMiB" % mb)
    timed(zero_copy, "zero-copy (UVA pointers)")
    print("bitwise equal logp:", bool(torch.equal(h_logp, h_logp2)), " dt:", bool(torch.equal(h_dt, h_dt2)))


if __name__ == "__main__":
    main()
