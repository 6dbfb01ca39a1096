"""Does the fused kernel's time depend on WHERE dt sits relative to t?  (read stream and write stream of the same
launch meeting in the same DRAM banks / channels).  cfg2 chain, 2^20 rows; t fixed, dt placed at controlled byte
offsets inside one large allocation; 60 launches per placement."""
import sys

import torch

sys.path.insert(0, ".")
from normalizingflownetwork_b200 import functional as F  # noqa: E402

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
ft, d, tb = {"cfg2": (["planar", "radial", "affine"] * 3 + ["planar"], 2, True), "cfg4": (["radial"] * 5, 1, True),
             "r10d2": (["radial"] * 10, 2, True)}[name]
P = F.chain_param_size(ft, d, tb)
B = 1 << 20
g = torch.Generator(device=dev).manual_seed(22)
pool = torch.empty((3 * B * P + (64 << 20),), device=dev)
t = pool[: B * P].view(B, P)
t.copy_(torch.randn((B, P), generator=g, device=dev) * 0.5)
y = torch.randn((B, d), generator=g, device=dev)
logp = torch.empty(B, device=dev)
col = torch.zeros(P, dtype=torch.float64, device=dev)
ls = torch.zeros(1, dtype=torch.float64, device=dev)
print("t at 0x%x  (B*P*4 = %d bytes)" % (t.data_ptr(), B * P * 4))
for off in (0, 256, 1024, 4096, 8192, 65536, 1 << 20, (1 << 20) + 4096, 2 << 20, (2 << 20) + 65536, 3 << 20, 5 << 20,
            7 << 20, 8 << 20, 16 << 20, (16 << 20) + (1 << 19), 24 << 20, 32 << 20):
    base = B * P + off // 4
    dt = pool[base: base + B * P].view(B, P)
    def run():
        F.chain_forward_backward(t, y, ft, d, tb, g_scale=-1.0 / B, logp_sum=ls, dt_colsum=col, out_logp=logp, out_dt=dt)
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(60):
        run()
    e1.record()
    torch.cuda.synchronize()
    print("dt offset %9d B (dt - t = %d mod 2^21 = %7d): %.2f us" % (off, dt.data_ptr() - t.data_ptr(),
          (dt.data_ptr() - t.data_ptr()) % (1 << 21), e0.elapsed_time(e1) * 1e3 / 60), flush=True)
