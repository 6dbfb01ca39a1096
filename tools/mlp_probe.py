import sys, os
sys.path.insert(0, os.getcwd())
import torch
from normalizingflownetwork_b200 import functional as F
dev = torch.device("cuda:0")
B, K, N = 1 << 20, 16, 16
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn((B, K), generator=g, device=dev); w = torch.randn((N, K), generator=g, device=dev) * 0.3
b = torch.zeros(N, device=dev); up = torch.randn((B, N), generator=g, device=dev)
out = F.dense_act_forward(x, w, b, "tanh")
for _ in range(3):
    F.dense_act_backward(x, out, up, w, "tanh", need_dx=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): F.dense_act_backward(x, out, up, w, "tanh", need_dx=True)
e1.record(); torch.cuda.synchronize()
print("bwd K=16 N=16: %.1f us" % (e0.elapsed_time(e1) * 100))
x1 = torch.randn((B, 1), generator=g, device=dev); w1 = torch.randn((N, 1), generator=g, device=dev)
o1 = F.dense_act_forward(x1, w1, b, "tanh")
for _ in range(3): F.dense_act_backward(x1, o1, up, w1, "tanh", need_dx=False)
torch.cuda.synchronize(); e0.record()
for _ in range(10): F.dense_act_backward(x1, o1, up, w1, "tanh", need_dx=False)
e1.record(); torch.cuda.synchronize()
print("bwd K=1 N=16 (no dx): %.1f us" % (e0.elapsed_time(e1) * 100))
