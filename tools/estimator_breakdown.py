#!/usr/bin/env python
"""Where an estimator-level step spends its time at a large batch: NormalizingFlowNetwork (MLP (16,16) tanh,
cfg2-like chain is not an estimator default, so: 10 radial flows, 1-D y) at B = 2^20 rows.
Prints the CUDA-event time of log_pdf and of one train step, and the per-kernel breakdown (torch profiler)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from normalizingflownetwork_b200.estimators import NormalizingFlowNetwork

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
rng = np.random.default_rng(0)
x = rng.normal(size=(B, 1)).astype(np.float32)
y = (np.cos(x) + 0.3 * rng.normal(size=(B, 1))).astype(np.float32)
model = NormalizingFlowNetwork.build_function(n_dims=1, n_flows=10, hidden_sizes=(16, 16), activation="tanh")
model.fit(x[:4096], y[:4096], batch_size=1024, epochs=1, verbose=0)
xd, yd = model._to_dev(x), model._to_dev(y)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


print("rows %d" % B)
print("log_pdf      %8.1f us" % timed(lambda: model.log_pdf(xd, yd)))
print("train_step   %8.1f us" % timed(lambda: model.train_step(xd, yd)))
if len(sys.argv) > 2 and sys.argv[2] == "graph":
    model.capture_train_step(B, 1, 1)
    print("train_step (CUDA graph) %8.1f us" % timed(lambda: model.train_step_graphed(xd, yd)))
    model.capture_log_pdf(B, 1, 1)
    print("log_pdf    (CUDA graph) %8.1f us" % timed(lambda: model.log_pdf_graphed(xd, yd)))
    sys.exit(0)
for name, fn in (("log_pdf", lambda: model.log_pdf(xd, yd)), ("train_step", lambda: model.train_step(xd, yd))):
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
    print("---- %s: CUDA kernels, 5 calls" % name)
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
