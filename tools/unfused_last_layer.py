"""Cost of the emitting Dense(P) layer around the flow kernel (cfg2 shape): h[B,16] @ W[16,48] + b -> t,
fused flow fwd+bwd -> dt, then dh = dt @ W^T, dW = h^T dt, db = sum dt  (torch / cuBLAS for the GEMMs)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from normalizingflownetwork_b200 import functional as F

dev = torch.device("cuda:0")
B, H, P, d = 1 << 20, 16, 48, 2
ft = ["planar", "radial", "affine"] * 3 + ["planar"]
g = torch.Generator(device=dev).manual_seed(22)
h = torch.tanh(torch.randn((B, H), generator=g, device=dev))
W = torch.randn((H, P), generator=g, device=dev) * 0.3
b = torch.zeros(P, device=dev)
y = torch.randn((B, d), generator=g, device=dev)
t = torch.empty((B, P), device=dev); dt = torch.empty((B, P), device=dev); logp = torch.empty(B, device=dev)
dh = torch.empty((B, H), device=dev); dW = torch.empty((H, P), device=dev)

def timed(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n

def fwd_gemm(): torch.addmm(b, h, W, out=t)
def flow(): F.chain_forward_backward(t, y, ft, d, True, g_scale=-1.0 / B, out_logp=logp, out_dt=dt)
def dgrad(): torch.mm(dt, W.t(), out=dh)
def wgrad(): torch.mm(h.t(), dt, out=dW)
def bgrad(): return dt.sum(0)
def step(): fwd_gemm(); flow(); dgrad(); wgrad(); bgrad()
dWf = torch.zeros((H, P), device=dev); dbf = torch.zeros(P, device=dev)
def fused(): F.dense_chain_forward_backward(h, W, b, y, ft, d, True, g_scale=-1.0 / B, dW=dWf, dbias=dbf)
def fused_fwd(): F.dense_chain_forward(h, W, b, y, ft, d, True)
if len(sys.argv) > 1 and sys.argv[1] == "fused-only":
    print("%-26s %8.1f us" % ("FUSED layer+flow fwd+bwd", timed(fused, 10)))
    sys.exit(0)
for name, fn in [("FUSED layer+flow fwd+bwd", fused), ("FUSED layer+flow fwd", fused_fwd), ("t = h W + b", fwd_gemm), ("flow fwd+bwd", flow), ("dh = dt W^T", dgrad), ("dW = h^T dt", wgrad),
                 ("db = sum dt", bgrad), ("whole unfused step", step)]:
    print("%-26s %8.1f us" % (name, timed(fn)))
