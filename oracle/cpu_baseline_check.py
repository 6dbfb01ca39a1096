"""Is the timed CPU port (`oracle/flow_oracle.py`, bench.py's cpu_baseline / --impl reference) a fair stand-in
for the reference's Python?  Where /root/reference exists, run the reference's OWN layer code on the float32
torch-CPU stand-ins (oracle/tf_shim.py) and the port on the same config-2 inputs, forward + autograd backward,
and print both timings and the number of aten ops each dispatches.  TEST INFRASTRUCTURE ONLY (CPU).

    python -m oracle.cpu_baseline_check [rows]
"""
import os
import sys
import time

import torch
from torch.utils._python_dispatch import TorchDispatchMode

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import flow_oracle as fo  # noqa: E402
from oracle import tf_shim  # noqa: E402


class CountOps(TorchDispatchMode):
    def __init__(self):
        super().__init__()
        self.n = 0

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        self.n += 1
        return func(*args, **(kwargs or {}))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    ft, d, tb = ["planar", "radial", "affine"] * 3 + ["planar"], 2, True
    P = fo.chain_param_size(ft, d, tb)
    g = torch.Generator().manual_seed(22)
    t = (0.5 * torch.randn(B, P, generator=g)).requires_grad_(True)
    y = torch.randn(B, d, generator=g)
    _, DL = tf_shim.load_reference(dtype=torch.float32)
    layer = DL.InverseNormalizingFlowLayer(ft, d, trainable_base_dist=tb)

    def ref_step():
        loss = -layer(t).log_prob(y).mean()
        return torch.autograd.grad(loss, t)[0]

    def port_step():
        loss = -fo.chain_log_prob(t, y, ft, d, tb).mean()
        return torch.autograd.grad(loss, t)[0]

    out = {}
    for name, step in (("reference code on stand-ins", ref_step), ("port (oracle/flow_oracle.py)", port_step)):
        with CountOps() as c:
            g0 = step()
        for _ in range(2):
            step()
        t0 = time.perf_counter()
        n = 5
        for _ in range(n):
            step()
        dt = (time.perf_counter() - t0) / n
        out[name] = g0
        print("%-32s %8.1f ms/step  %.3e samples/s  %5d aten ops (fwd+bwd)" % (name, dt * 1e3, B / dt, c.n))
    a, b = out.values()
    print("max |dt_ref - dt_port| = %.2e  (threads: %d)" % (float((a - b).abs().max()), torch.get_num_threads()))
    tf_shim.uninstall()


if __name__ == "__main__":
    main()
