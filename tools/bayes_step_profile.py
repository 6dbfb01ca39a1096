"""Where the cfg4 Bayesian training step goes: kernel time by name (torch.profiler) next to the wall time per step."""
import os
import sys

sys.path.insert(0, os.getcwd())
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from normalizingflownetwork_b200.estimators import BayesNormalizingFlowNetwork  # noqa: E402

S, Bl = 32, 1 << 15
g = torch.Generator(device="cpu").manual_seed(22)
x = torch.rand((Bl, 1), generator=g) * 6.0 - 3.0
y = torch.cos(x) + 0.3 * torch.randn((Bl, 1), generator=g)
model = BayesNormalizingFlowNetwork(1, kl_weight_scale=1.0 / Bl, n_flows=5, hidden_sizes=(10,), activation="tanh",
                                    n_train_draws=S, learning_rate=2e-2)
model._assign_data_normalization(x.numpy(), y.numpy())
with torch.no_grad():
    model.params_from_x(x[:2].numpy())
model.optimizer = model._make_adam()
xd, yd = model._to_dev(x), model._to_dev(y)
for _ in range(5):
    model.train_step(xd, yd)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    model.train_step(xd, yd)
e1.record()
torch.cuda.synchronize()
print("eager step: %.1f us" % (e0.elapsed_time(e1) * 100))
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5):
        model.train_step(xd, yd)
    torch.cuda.synchronize()
ka = prof.key_averages()
rows = sorted(((k.device_time_total / 5.0, k.count / 5.0, k.key) for k in ka if k.device_time_total > 0), reverse=True)
tot = sum(r[0] for r in rows if not r[2].startswith("aten::") and not r[2].startswith("autograd::"))
print("device time per step by kernel (us), kernels only: total %.1f" % tot)
n = 0
for t, c, k in rows:
    if k.startswith("aten::") or k.startswith("autograd::") or "Backward" in k or k.startswith("Optimizer"):
        continue
    print("  %8.1f us  x%5.1f  %s" % (t, c, k[:110]))
    n += 1
    if n >= 25:
        break
