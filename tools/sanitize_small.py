"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from normalizingflownetwork_b200 import functional as F
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
def rnd(*s, sc=0.5): return torch.randn(s, generator=g, device=dev) * sc
for ft, d, tb in [(["planar", "radial", "affine"] * 3 + ["planar"], 2, True), (["radial"] * 5, 1, True),
                  (["radial", "planar"] * 8, 4, True), (["affine", "planar"], 3, False)]:
    P = F.chain_param_size(ft, d, tb)
    for B in (1, 129, 700):
        t, y = rnd(B, P), rnd(B, d, sc=1.0)
        F.chain_forward(t, y, ft, d, tb)
        col = torch.zeros(P, dtype=torch.float64, device=dev); ls = torch.zeros(1, dtype=torch.float64, device=dev)
        F.chain_forward_backward(t, y, ft, d, tb, g_scale=-1.0 / B, want_dy=True, logp_sum=ls, dt_colsum=col)
    F.set_option("force_generic", 1)
    F.chain_forward_backward(rnd(300, P), rnd(300, d, sc=1.0), ft, d, tb)
    F.set_option("force_generic", 0)
    for io in ("cpasync", "tma"):   # both kernel generations
        F.set_option("chain_io", io)
        F.chain_forward(rnd(700, P), rnd(700, d, sc=1.0), ft, d, tb)
        F.chain_forward_backward(rnd(700, P), rnd(700, d, sc=1.0), ft, d, tb, want_dy=True)
    F.set_option("chain_io", "auto")
for K, d in [(20, 2), (3, 1), (5, 5)]:
    P = 2 * K * d + K
    F.mdn_forward_backward(rnd(333, P), rnd(333, d, sc=1.0), K, d, want_dy=True)
    F.mdn_forward(rnd(129, P), rnd(1, d, sc=1.0), K, d)
for M, d in [(20, 1), (30, 2)]:
    F.kmn_forward_backward(rnd(400, M, sc=1.0), rnd(400, d, sc=1.0), rnd(M, d, sc=1.0),
                           torch.full((M,), 0.4, device=dev), want_dy=True)
for ft, d, tb, H in [(["planar", "radial", "affine"] * 3 + ["planar"], 2, True, 16), (["radial", "planar"], 3, True, 32)]:
    P = F.chain_param_size(ft, d, tb)
    for B in (1, 200, 641):
        F.dense_chain_forward_backward(torch.tanh(rnd(B, H, sc=1.0)), rnd(H, P, sc=0.1), rnd(P, sc=0.1), rnd(B, d, sc=1.0),
                                       ft, d, tb, g_scale=-1.0 / B)
        F.dense_chain_forward(torch.tanh(rnd(B, H, sc=1.0)), rnd(H, P, sc=0.1), rnd(P, sc=0.1), rnd(B, d, sc=1.0), ft, d, tb)
F.logmeanexp_draws(rnd(8, 1000))
from normalizingflownetwork_b200 import parallel
comm = parallel.PeerComm(49, dev)
F.chain_forward_backward_peer(rnd(1000, 48), rnd(1000, 2, sc=1.0), ["planar", "radial", "affine"] * 3 + ["planar"], 2, True,
                              comm, want_colsum=True)
torch.cuda.synchronize(); comm.close()
print("sanitize_small: done")
