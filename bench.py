#!/usr/bin/env python
"""Headline benchmark: flow-chain log-prob forward+backward samples/s (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2]

Workload (BASELINE.json configs[1], the one the metric is quoted on): 10-flow chain
(planar, radial, affine) x 3 + planar over 2-D y with a trainable base distribution,
P = 48 parameters per sample, batch 2^20 rows PER GPU (weak scaling), fused forward +
reverse sweep with the mean-NLL cotangent -1/B.  One "step" = one pass of the hot path
over one batch of synthetic fp32 input.

  value     device-resident: inputs already in HBM, CUDA-event timed, max over ranks
  e2e       the same pass through the HOST-buffer C-ABI call
            (nfn_chain_forward_backward_host): pinned host t/y -> H2D -> kernel -> D2H of
            logp/dt + the loss scalar, copies inside the timed region
  roofline  algorithmic bytes 4*(2P+d+1) per row / measured kernel time vs the measured
            HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline / --impl reference
            the op-for-op float32 torch-CPU restatement of the reference's TF graph
            (oracle/flow_oracle.py; TensorFlow itself cannot run in this image) with
            autograd backward on all host cores, on a bounded sample of the same workload.

N > 1 (torchrun): every rank runs its own 2^20-row shard; the step additionally performs
the one exchange a data-parallel training step needs from this path, a packed all-reduce
of [dt column sums (bias gradient of the emitting layer) | sum logp] over NCCL.
"""
import argparse
import ctypes
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (flow_types, n_dims, trainable_base, rows per GPU, fwd+bwd?)
    "cfg2": (["planar", "radial", "affine"] * 3 + ["planar"], 2, True, 1 << 20, True),
    "cfg3": (["radial", "planar"] * 8, 4, True, 1 << 23, False),
    "cfg4": (["radial"] * 5, 1, True, 1 << 20, True),
    "cfg1": (["radial"] * 3, 1, True, 2048, True),
    # MDN head: flow_types slot holds ("mdn", n_centers)
    "cfg5": (("mdn", 20), 2, None, 1 << 22, True),
    # not BASELINE configs: the chains NormalizingFlowNetwork builds by default (n_flows = 10, all radial), tuning only
    "r10d1": (["radial"] * 10, 1, True, 1 << 20, True),
    "r10d2": (["radial"] * 10, 2, True, 1 << 20, True),
}
WORKLOAD_NAMES = {
    "cfg5": "MDN 20-component Gaussian mixture head, 2-D y, P=100, batch 2^22 per GPU, logsumexp log-lik fwd+bwd",
    "cfg2": "NFN 10 flows (planar,radial,affine)x3+planar, 2-D y, P=48, batch 2^20 per GPU, log-prob fwd+bwd",
    "cfg3": "NFN 16 flows (radial,planar)x8, 4-D y, P=128, 2^23 (x,y) pairs per GPU, density-grid log-prob fwd",
    "cfg4": "Bayesian NFN 5 radial flows, 1-D y, P=17, S*B = 2^20 rows per GPU, log-prob fwd+bwd",
    "cfg1": "NFN 3 radial flows, 1-D y, P=11, batch 2048, log-prob fwd+bwd",
    "r10d1": "NFN 10 radial flows, 1-D y, P=32, batch 2^20 per GPU, log-prob fwd+bwd (tuning)",
    "r10d2": "NFN 10 radial flows, 2-D y, P=44, batch 2^20 per GPU, log-prob fwd+bwd (tuning)",
}
METRIC = "flow log-prob fwd+bwd samples/sec"
UNIT = "samples/s"


def is_mdn(flow_types):
    return isinstance(flow_types, tuple) and len(flow_types) == 2 and flow_types[0] == "mdn"


def param_size(flow_types, d, tb):
    if is_mdn(flow_types):
        return 2 * flow_types[1] * d + flow_types[1]
    sz = {"planar": 2 * d + 1, "radial": d + 2, "affine": 2 * d}
    return sum(sz[f] for f in flow_types) + (2 * d if tb else 0)


def arm_config(cfg, world, K, B, P, d, bwd, specialized, want_col, packed, use_peer, peer_blocking, extra_warmup):
    """`config` of the JSON line: the workload and how the GPU arm runs it.  Both arms print THIS dict for the same
    command line (the reference arm times a bounded sample of the same workload on the host cores and says which in
    `cpu_baseline.sample`), so that the two lines of a round compare like with like."""
    bytes_per_row = 4 * ((2 * P if bwd else P) + d + 1)
    exchange = ""
    if packed:
        exchange = "; [dt column sums (P) | sum logp] summed over ranks every step, " + (
            ("fused into the kernel over NVLink peer memory, split-phase (one CTA of step i+1's grid pushes step i's "
             "totals, collects the peers' and writes the sums; the last step's by a flush inside the timed region)"
             if not peer_blocking else
             "fused into the kernel's last CTA over NVLink peer memory (push + wait)") if use_peer else
            "one NCCL all-reduce")
    return {
        "workload": WORKLOAD_NAMES[cfg], "rows_per_gpu": B, "param_width": P, "n_dims": d,
        "fwd_bwd": bwd, "specialized_kernel": specialized,
        "math": "accurate" if os.environ.get("NFN_B200_MATH") == "accurate" else "fast",
        "dt_column_sums_in_kernel": want_col,
        "l2": ("inputs+outputs per step (%d MB) exceed the 126 MB L2; no flush needed" if bytes_per_row * B > (126 << 20) else
               "inputs+outputs per step (%d MB) FIT the 126 MB L2 and are not flushed: a debug / latency-bound row count, "
               "not a roofline claim") % (bytes_per_row * B // (1 << 20)),
        "parallelism": "dp%d (rows sharded, no data-path collective%s)" % (world, exchange),
        "t_sigma": 0.5, "seed": 22, "extra_untimed_warmup_steps": extra_warmup,
        "timing": "one CUDA-event pair around the %d steps, no per-launch probes" % K + (
            "; after barrier + synchronize the ranks' streams are lined up by one tiny all-reduce queued directly "
            "before the start event" if world > 1 else ""),
    }


def static_arm_config(args):
    """The same dict without a GPU, the library or the package (the reference arm): every entry follows from the
    command line; the exchange is the default the GPU arm selects for this world size."""
    cfg = args.config
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ft, d, tb, rows, bwd = CONFIGS[cfg]
    mdn = is_mdn(ft)
    bwd = bool(bwd and not args.fwd_only)
    # every chain in CONFIGS has an ahead-of-time kernel instance (build.py:SPECIALIZED_CHAINS; pinned by
    # tests/test_bench_contract.py) -- stated here so that this arm imports nothing of the product
    specialized = True
    packed = bool((world > 1 or args.force_peer) and bwd)
    use_peer = bool(packed and args.exchange == "peer" and not mdn)
    return arm_config(cfg, world, args.steps, int(args.rows or rows), param_size(ft, d, tb), d, bwd, specialized,
                      bool(bwd and not args.no_colsum and not mdn), packed, use_peer, args.peer_blocking,
                      max(0, 600 - args.warmup) if world > 1 else 0)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        except Exception:
            pass
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def load_traffic(cfg):
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu capture, or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(cfg)
        except Exception:
            return None
    return None


class _NearGpuCpus:
    """Context: bind the calling thread to the CPUs NVML reports as local to GPU `gpu` (NUMA placement of
    pinned host buffers), restore the previous affinity on exit.  Best effort: any failure leaves the
    affinity untouched and says so in `note`."""

    def __init__(self, gpu):
        self.gpu, self.old, self.note = gpu, None, "unchanged"

    def __enter__(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            idx = self.gpu
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    idx = self.gpu
            old = os.sched_getaffinity(0)
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(idx))
            new = os.sched_getaffinity(0)
            self.old = old
            self.note = "%d of %d cpus (NVML ideal affinity of the GPU)" % (len(new), len(old))
        except Exception as e:  # no NVML, cpuset without the ideal CPUs, ...
            self.note = "unchanged (%s)" % type(e).__name__
        return self

    def __exit__(self, *exc):
        if self.old is not None:
            try:
                os.sched_setaffinity(0, self.old)
            except Exception:
                pass
        return False


class ClockSampler:
    """SM clock + throttle reasons polled through NVML in a background thread DURING the
    timed region (the region is a few milliseconds, too short for `nvidia-smi -lms`)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []  # (perf_counter, sm_mhz, reasons bitmask)
        self.stop_flag = False
        self.thread = None
        self.h = None
        self.max_mhz = None

    def start(self):
        try:
            import threading

            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            idx = self.gpu
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    idx = self.gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))

            def loop():
                while not self.stop_flag:
                    try:
                        mhz = pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
                        rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                        self.samples.append((time.perf_counter(), float(mhz), int(rs)))
                    except Exception:
                        pass
                    time.sleep(0.0005)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None

    def stop(self, t0=None, t1=None):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        if self.thread is None:
            return out
        self.stop_flag = True
        self.thread.join(timeout=2)
        nv = self.nv
        inside = [s for s in self.samples if t0 is not None and t0 <= s[0] <= t1]
        window = "timed region"
        if len(inside) < 3:
            inside = self.samples
            window = "warm-up + timed region (timed region too short for >= 3 samples)"
        if inside:
            names = {
                "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            }
            mask = 0
            for s in inside:
                mask |= s[2]
            out = {"sm_mhz": statistics.median(s[1] for s in inside), "sm_max_mhz": self.max_mhz,
                   "reasons": sorted(k for k, v in names.items() if mask & v), "samples": len(inside),
                   "window": window}
        try:
            nv.nvmlShutdown()
        except Exception:
            pass
        return out


# --------------------------------------------------------------------------- CPU baseline
def cpu_pass_fn(cfg, rows, seed=22):
    """Returns a callable doing one fwd(+bwd) pass of the TF-equivalent fp32 CPU restatement."""
    import torch

    from oracle import flow_oracle as fo

    ft, d, tb, _, bwd = CONFIGS[cfg]
    P = param_size(ft, d, tb)
    g = torch.Generator().manual_seed(seed)
    t = (torch.randn((rows, P), generator=g) * 0.5)
    y = torch.randn((rows, d), generator=g)

    def logp_fn(tt):
        if is_mdn(ft):
            return fo.mdn_log_prob(tt, y, ft[1], d)
        return fo.chain_log_prob(tt, y, ft, d, tb)

    def one_pass():
        if bwd:
            tt = t.detach().requires_grad_(True)
            nll = -logp_fn(tt).mean()
            nll.backward()
            return float(nll.detach())
        with torch.no_grad():
            return float(logp_fn(t).mean())

    return one_pass


def cpu_baseline(cfg, budget_s=12.0, rows=1 << 16):
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fn = cpu_pass_fn(cfg, rows)
    fn()  # warm-up (allocator, thread pool)
    n, t0 = 0, time.perf_counter()
    while True:
        fn()
        n += 1
        el = time.perf_counter() - t0
        if el >= budget_s or n >= 200:
            break
    return {
        "value": rows * n / el, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": "%d passes over %d rows of the same workload (fp32 torch-CPU op-for-op restatement of the "
                  "reference's TF graph + autograd; TensorFlow not installable here)" % (n, rows),
    }


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (port; see cpu_baseline)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch

    cfg = args.config
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rows = 1 << 16
    fn = cpu_pass_fn(cfg, rows)
    for _ in range(max(1, args.warmup)):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    el = time.perf_counter() - t0
    value = rows * args.steps / el
    ft, d, tb, _, bwd = CONFIGS[cfg]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the GPU arm's `config` for the same command line (what this run samples); the sample itself is below
        "config": static_arm_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d rows per step (a bounded sample of the workload in `config`, host cores only); "
                                   "fp32 torch-CPU op-for-op restatement of the reference's TF "
                                   "graph + autograd (TensorFlow/TFP not installable in this image)" % rows},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------- our arm
class HeadWorkload:
    """One BASELINE config as device-resident synthetic tensors + the C-ABI launch of its head.

    step() = one launch of the head kernel over this rank's rows (+ the one exchange a
    data-parallel training step needs from the path when world > 1: the sum over ranks of
    [dt column sums (P) | sum logp], fused into the kernel's last CTA over NVLink peer memory,
    or one NCCL all-reduce with --exchange nccl).  `sets` > 1 rotates that many independent
    tensor sets so that a per-step footprint near the 126 MB L2 cannot be served from it."""

    def __init__(self, cfg, args, device, rank, world, lib, n_steps, rows=None, fwd_only=False, colsum=True,
                 sets=1):
        import torch

        from normalizingflownetwork_b200 import _lib, parallel

        self.torch, self._lib, self.lib, self.parallel = torch, _lib, lib, parallel
        self.cfg, self.device, self.rank, self.world = cfg, device, rank, world
        ft, d, tb, B, bwd = CONFIGS[cfg]
        self.ft, self.d, self.tb = ft, d, tb
        self.bwd = bool(bwd and not fwd_only)
        self.B = int(rows or B)
        self.P = param_size(ft, d, tb)
        self.mdn = is_mdn(ft)
        self.desc = None if self.mdn else _lib.make_desc(ft, d, tb)
        self.specialized = True if self.mdn else bool(lib.nfn_chain_is_specialized(ctypes.byref(self.desc)))
        B, P = self.B, self.P
        gen = torch.Generator(device=device).manual_seed(22 + rank)
        self.sets = []
        for _ in range(max(1, sets)):
            t = torch.randn((B, P), generator=gen, device=device) * 0.5
            y = torch.randn((B, d), generator=gen, device=device)
            logp = torch.empty(B, device=device)
            dt = torch.empty((B, P), device=device) if self.bwd else None
            self.sets.append((t, y, logp, dt))
        # fp64 accumulators [dt column sums (P) | sum logp], one pre-zeroed row per step (keeps memsets off
        # the critical path); a data-parallel step sums its row over ranks with ONE exchange
        self.acc_ring = torch.zeros((n_steps + 8, P + 1), dtype=torch.float64, device=device)
        # (the mixture kernels leave the column sums to a second pass over dt: not part of their step here)
        self.want_col = bool(self.bwd and colsum and not self.mdn)
        self.packed = bool((world > 1 or args.force_peer) and self.bwd)
        self.use_peer = bool(self.packed and args.exchange == "peer" and not self.mdn)
        self.comm = None
        if self.use_peer:
            # cudaIpc peer mapping can be unavailable (e.g. ranks in different containers): every rank must
            # agree, so the outcome is all-reduced and the NCCL exchange is the fallback
            try:
                self.comm = parallel.PeerComm(P + 1, device)
                # split-phase: a launch sends the PREVIOUS step's totals from its head and collects their sums at its
                # tail (NVLink latency and rank skew hide behind a kernel of tile work); the last step's by a flush
                self.comm.set_deferred(not args.peer_blocking)
                ok = 1
            except Exception as exc:  # noqa: BLE001
                ok = 0
                if rank == 0:
                    print("peer exchange unavailable (%s): falling back to NCCL" % exc, file=sys.stderr)
            flag = torch.tensor([ok], device=device)
            if world > 1:
                torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
            if int(flag.item()) == 0:
                if self.comm is not None:
                    self.comm.close()
                self.comm, self.use_peer = None, False
        self.step_no = 0
        self._align = torch.zeros(1, device=device)
        self.g_scale = -1.0 / (B * world)
        self.stream = _lib.current_stream(device)
        self.bytes_per_row = 4 * ((2 * P if self.bwd else P) + d + 1)

    # ------------------------------------------------------------------ launches
    def kernel(self):
        _lib, lib, P, B, d = self._lib, self.lib, self.P, self.B, self.d
        t, y, logp, dt = self.sets[self.step_no % len(self.sets)]
        row = self.acc_ring.data_ptr() + self.step_no * (P + 1) * 8
        lsum = ctypes.c_void_p(row + 8 * P)
        col = ctypes.c_void_p(row) if self.want_col else None
        if self.mdn and not self.bwd:
            _lib.check(lib.nfn_mdn_forward(self.ft[1], d, _lib.ptr(t), _lib.ptr(y), B, _lib.ptr(logp), B, self.stream))
        elif self.use_peer:  # ONE launch: fused fwd+bwd + all-reduce of [dt column sums | sum logp] over NVLink
            _lib.check(lib.nfn_chain_forward_backward_peer(
                ctypes.byref(self.desc), _lib.ptr(t), _lib.ptr(y), B, None, ctypes.c_float(self.g_scale),
                _lib.ptr(logp), _lib.ptr(dt), None, 1 if self.want_col else 0, self.comm.comm, ctypes.c_void_p(row), B,
                self.stream))
        elif self.mdn:
            _lib.check(lib.nfn_mdn_forward_backward(
                self.ft[1], d, _lib.ptr(t), _lib.ptr(y), B, None, ctypes.c_float(self.g_scale), _lib.ptr(logp),
                _lib.ptr(dt), None, lsum, col, B, self.stream))
        elif self.bwd:
            _lib.check(lib.nfn_chain_forward_backward(
                ctypes.byref(self.desc), _lib.ptr(t), _lib.ptr(y), B, None, ctypes.c_float(self.g_scale),
                _lib.ptr(logp), _lib.ptr(dt), None, lsum, col, B, self.stream))
        else:
            _lib.check(lib.nfn_chain_forward(ctypes.byref(self.desc), _lib.ptr(t), _lib.ptr(y), B, _lib.ptr(logp), B,
                                             self.stream))

    def exchange(self):
        if self.packed and not self.use_peer:
            self.torch.distributed.all_reduce(self.acc_ring[self.step_no])
        self.step_no += 1

    def finish(self):
        """Inside the timed region, after the last step: completes the pending split-phase exchange."""
        if self.comm is not None:
            self.comm.flush()

    def step(self):
        self.kernel()
        self.exchange()

    @property
    def step_is_one_kernel(self):
        return not (self.packed and not self.use_peer)

    # ------------------------------------------------------------------ timing
    def time_steps(self, K, with_exchange=True):
        """K back-to-back steps under ONE CUDA-event pair on the launch stream (no per-launch probes: an event
        pair around every launch opens gaps and defeats the programmatic-dependent-launch overlap), bracketed
        by barrier + synchronize; returns the max over ranks in ms."""
        torch, parallel = self.torch, self.parallel
        torch.cuda.synchronize()
        parallel.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if self.world > 1:
            # the host threads of the ranks leave the barrier tens of microseconds apart, which a 20-step region
            # (1.4 ms) would bill to the ranks that wait for the late one's first exchange: line the GPUs up in
            # STREAM order -- a tiny all-reduce ends at (nearly) the same instant on every GPU, the start event
            # and the K steps are queued right behind it without another host synchronisation
            torch.distributed.all_reduce(self._align)
        e0.record()
        if with_exchange:
            for _ in range(K):
                self.kernel()
                self.exchange()
        else:
            for _ in range(K):
                self.kernel()
                self.step_no += 1
        em = torch.cuda.Event(enable_timing=True)
        em.record()
        self.finish()
        e1.record()
        torch.cuda.synchronize()
        parallel.barrier()
        torch.cuda.synchronize()
        self.last_flush_ms = em.elapsed_time(e1)   # the split-phase flush (inside the timed region), for the record
        return parallel.max_over_ranks(e0.elapsed_time(e1), self.device)

    # ------------------------------------------------------------------ the result of the last step, checked
    def exchange_check(self):
        """Compares the accumulator row of the LAST executed step -- after the cross-rank exchange when
        world > 1 -- with a float64 torch reduction of the same step's outputs summed over ranks by a plain
        NCCL all-reduce.  This is what proves that the fused peer all-reduce delivered the cross-rank sum
        (and that the in-kernel column sums are in the timed region)."""
        torch = self.torch
        if not self.bwd:
            return None
        self.finish()
        if self.comm is not None:
            self.comm.status()   # raises if any exchange timed out
        last = self.step_no - 1
        t, y, logp, dt = self.sets[last % len(self.sets)]
        ref = torch.zeros(self.P + 1, dtype=torch.float64, device=self.device)
        if self.want_col:
            ref[: self.P] = dt.double().sum(0)
        ref[self.P] = logp.double().sum()
        if self.world > 1:
            torch.distributed.all_reduce(ref)
        # every executed step that ran on the same tensor set as the last one must hold the same sums: the last
        # row is completed by the flush, the earlier ones by the head of the launch that followed them
        rows = [r for r in range(last, max(-1, last - 4 * len(self.sets)), -len(self.sets)) if r >= 0]
        err_col, err_lp, nan, got = 0.0, 0.0, False, None
        # a synthetic row can be singular (det -> 0 at some row counts): its column is non-finite in the reference
        # as well; such columns must be non-finite in the kernel's sums too and are left out of the error norm
        fin = torch.isfinite(ref)
        col_scale = float(ref[: self.P][fin[: self.P]].abs().max()) if (self.want_col and bool(fin[: self.P].any())) else 0.0
        for r in rows:
            got = self.acc_ring[r].clone()
            if not self.want_col:
                got[: self.P] = 0.0
            else:
                dcol = (got[: self.P] - ref[: self.P])[fin[: self.P]]
                if dcol.numel():
                    err_col = max(err_col, float(dcol.abs().max()) / max(col_scale, 1e-30))
            if bool(fin[self.P]):
                err_lp = max(err_lp, abs(float(got[self.P] - ref[self.P])) / max(abs(float(ref[self.P])), 1e-30))
            nan = nan or bool((~torch.isfinite(got[fin])).any().item()) or bool(torch.isfinite(got[~fin]).any().item())
        out = {"max_rel_err": max(err_col, err_lp), "colsum_rel_err": err_col, "logp_sum_rel_err": err_lp, "nan": nan,
               "steps_checked": len(rows), "payload_values": self.P + 1,
               "non_finite_reference_columns": int((~fin).sum().item()),
               "payload_nonzero": int((got != 0).sum().item()), "payload_abs_max": float(got.abs().max()),
               "ranks_summed": self.world,
               "how": ("fused peer all-reduce inside the chain kernel" if self.use_peer else
                       "NCCL all-reduce of the accumulator row" if self.packed else "single rank: in-kernel fp64 accumulators")
                      + " vs float64 torch sums of dt / logp" + (" + NCCL all-reduce" if self.world > 1 else "")}
        # column sums: fp32 partial sums per warp / CTA, then fp64 atomics -> 1e-4 of the largest column; sum logp: fp64
        out["ok"] = bool((not nan) and err_col < 1e-4 and err_lp < 1e-6 and
                         (out["payload_nonzero"] > (self.P // 2 if self.want_col else 0)))
        return out

    def close(self):
        if self.comm is not None:
            self.comm.close()
            self.comm = None
        self.sets = []


def cfg4_bayes_train_step(device, rank, world, steps=10, warmup=4):
    """BASELINE config 4 as the estimator-level training step it names: Bayesian NFN (5 radial flows, 1-D y, one
    hidden layer of 10 tanh units, the reference's defaults), S = 32 Monte-Carlo weight draws folded into the batch,
    2^15 samples per GPU -> S * B = 2^20 folded rows per GPU (2^18 samples on 8 GPUs), data-parallel: every rank
    runs the variational MLP (batched GEMMs), ONE fused head launch over its 2^20 folded rows, the MLP backward,
    then ONE float32 all-reduce of the flat gradient buffer and Adam.  Reports the step time (max over ranks) and
    where it goes: the head kernel alone and the all-reduce alone, each timed separately with CUDA events."""
    import torch

    from normalizingflownetwork_b200 import parallel
    from normalizingflownetwork_b200 import functional as F
    from normalizingflownetwork_b200.estimators import BayesNormalizingFlowNetwork

    S, Bl = 32, 1 << 15
    g = torch.Generator(device="cpu").manual_seed(22 + rank)
    x = torch.rand((Bl, 1), generator=g) * 6.0 - 3.0
    y = torch.cos(x) + 0.3 * torch.randn((Bl, 1), generator=g)
    model = BayesNormalizingFlowNetwork(1, kl_weight_scale=1.0 / (Bl * world), n_flows=5, hidden_sizes=(10,),
                                        activation="tanh", n_train_draws=S, learning_rate=2e-2)
    model._assign_data_normalization(x.numpy(), y.numpy())
    with torch.no_grad():
        model.params_from_x(x[:2].numpy())
    model.optimizer = model._make_adam()
    xd, yd = model._to_dev(x), model._to_dev(y)
    Bg = Bl * world

    def timed(fn, n):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        parallel.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        parallel.barrier()
        return parallel.max_over_ranks(e0.elapsed_time(e1), device) / n   # ms

    losses = []
    step_ms = timed(lambda: losses.append(model.train_step(xd, yd, global_batch=Bg)), steps)
    loss = float(losses[-1])
    # breakdown: the head kernel on t[S * Bl, 17] alone, and the all-reduce of the flat gradient buffer alone
    layer = model.dist_layer
    with torch.no_grad():
        t = model.params_from_x_draws(xd, S)
    yy = yd.repeat(S, 1)
    ls = torch.zeros(1, dtype=torch.float64, device=device)
    xf = model._xform(1, training=True)
    head_ms = timed(lambda: F.chain_forward_backward(t, yy, layer._flow_types, 1, True, g_scale=-1.0 / (S * Bg),
                                                     logp_sum=ls, xform=xf), steps)
    red_ms = None
    n_grad = sum(p.numel() for p in model.parameters() if p.requires_grad)
    if world > 1:
        red = model._grad_reducer()
        red_ms = timed(red.reduce, steps)
    # one GPU: the same step captured once into a CUDA graph and replayed (the eager step is ~120 small launches).
    # Not at N > 1: a captured step containing the NCCL all-reduce never completed its replay on 2 GPUs (measured).
    graph_ms = graph_err = None
    if world == 1:
        try:
            model.capture_train_step(Bl, 1, 1)
            graph_ms = timed(lambda: model.train_step_graphed(xd, yd), steps)
            if not np.isfinite(float(model._graph_loss)):
                graph_err = "non-finite loss from the captured step"
        except Exception as exc:  # noqa: BLE001 -- an optional extra must not take the headline down
            graph_ms, graph_err = None, str(exc)[:200]
    peak, _ = load_peaks()
    fused = model._fused_draws_plan() is not None
    return {"workload": "Bayesian NFN 5 radial flows, 1-D y, hidden (10,) tanh, S=32 draws folded, %d samples per GPU "
                        "(S*B = 2^20 folded rows per GPU), Adam training step, data-parallel" % Bl,
            "global_samples_per_step": Bg, "folded_rows_per_gpu": S * Bl, "ms_per_step": step_ms,
            "samples_per_s": Bg / (step_ms * 1e-3), "folded_rows_per_s": S * Bg / (step_ms * 1e-3),
            "loss": loss, "loss_finite": bool(np.isfinite(loss)),
            "breakdown_ms": {"head_kernel_fwd_bwd": head_ms, "flat_grad_allreduce": red_ms,
                             "everything_else_and_launch_gaps": step_ms - head_ms - (red_ms or 0.0)},
            "head_roofline_frac": 4 * (2 * 17 + 1 + 1) * S * Bl / (head_ms * 1e-3) / 1e9 / peak,
            "allreduce": ("one float32 NCCL all-reduce of %d values (flat gradient buffer + sum logp)" % (n_grad + 2))
                         if world > 1 else None, "eager": True,
            "folded_draw_kernels": fused, "how": (
                "3 kernels for the network: first variational layer over S*B folded rows, emitting layer + flow chain + "
                "both backward GEMMs with per-draw weights (t / dt / repeated y never in HBM), first layer's per-draw "
                "weight gradient; weight draws, exact KL and Adam in torch" if fused else
                "batched GEMMs + streaming head + autograd"),
            "cuda_graph_ms_per_step": graph_ms, "cuda_graph_error": graph_err,
            "samples_per_s_cuda_graph": (Bg / (graph_ms * 1e-3)) if graph_ms else None}


def measure_other_config(cfg, args, device, rank, world, lib, steps=20, warmup=3):
    """A few steps of another BASELINE config, same launch path and timing rules as the headline."""
    import torch

    peak, _ = load_peaks()
    # cfg4's per-step footprint (151 MB) is only just above the 126 MB L2: rotate two tensor sets
    sets = 2 if cfg == "cfg4" else 1
    wl = HeadWorkload(cfg, args, device, rank, world, lib, steps + warmup + 32, sets=sets)
    try:
        for _ in range(warmup + (20 if world > 1 else 0)):
            wl.step()
        wl.finish()   # (first launch of the flush kernel outside the timed region)
        if world > 1:
            torch.distributed.all_reduce(wl._align)
        wl.step()
        ms = wl.time_steps(steps) / steps
        check = wl.exchange_check()
        out = {"workload": WORKLOAD_NAMES[cfg], "rows_per_gpu": wl.B, "fwd_bwd": wl.bwd, "steps": steps,
               "ms_per_step": ms, "samples_per_s": wl.B * world / (ms * 1e-3),
               "roofline_frac": wl.bytes_per_row * wl.B / (ms * 1e-3) / 1e9 / peak,
               "bytes_per_row": wl.bytes_per_row, "tensor_sets_rotated": sets}
        if check is not None:
            out["exchange_check"] = {k: check[k] for k in ("ok", "max_rel_err", "nan", "payload_nonzero")}
        return out
    finally:
        wl.close()
        del wl
        torch.cuda.empty_cache()


def run_ours(args):
    import torch

    from normalizingflownetwork_b200 import _lib, parallel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: normalizingflownetwork_b200 has no CPU fallback")
    rank, world, local_rank = parallel.init_process_group()
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    lib = _lib.load()
    cfg = args.config
    K, W = args.steps, args.warmup
    # N > 1: the per-step exchange couples the ranks, so cold NVLink links / peer mappings and start-up skew
    # would be billed to the first timed steps; top the warm-up up to 30 untimed steps (reported in config)
    # (NVLink links idle into a low-power state and a fresh process group is cold: ~45 ms of untimed steps.  The
    # count is FIXED: every rank must issue the same number of exchanges, a time-based loop would not)
    extra_warmup = max(0, 600 - W) if world > 1 else 0
    wl = HeadWorkload(cfg, args, device, rank, world, lib, 3 * K + W + extra_warmup + 64,
                      rows=args.rows or None,
                      fwd_only=args.fwd_only, colsum=not args.no_colsum)
    B, P, d, bwd, mdn = wl.B, wl.P, wl.d, wl.bwd, wl.mdn
    t, y, logp, dt = wl.sets[0]

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(W + extra_warmup):
        wl.step()
    # everything the timed region launches must have been launched once before it (CUDA loads kernels lazily: the
    # first launch of the split-phase flush kernel alone costs ~50 us): flush + stream alignment, then two more steps
    wl.finish()
    if world > 1:
        torch.distributed.all_reduce(wl._align)
    wl.step()
    wl.step()
    lib.nfn_launch_count_reset()
    wall0 = time.perf_counter()
    total_ms = wl.time_steps(K)
    launches = int(lib.nfn_launch_count_reset())
    ms_per_step = total_ms / K
    value = B * world * K / (total_ms * 1e-3)
    check = wl.exchange_check()
    # average launch duration of the dominant kernel: the timed region itself when a step is exactly one
    # launch of it, otherwise a second pass of K back-to-back launches without the NCCL exchange
    if wl.step_is_one_kernel:
        kern_ms, kern_how = ms_per_step, "timed region / steps (a step is exactly one launch of this kernel)"
    else:
        kern_ms = wl.time_steps(K, with_exchange=False) / K
        kern_how = "separate pass: %d back-to-back launches under one event pair (the step adds an NCCL all-reduce)" % K
    wall1 = time.perf_counter()
    replay_note = None
    if rank == 0 and sum(1 for smp in sampler.samples if wall0 <= smp[0] <= wall1) < 3 and world == 1:
        # the timed region was shorter than three NVML polls: replay the identical launches (untimed) for
        # ~40 ms so that the clocks are read under the same load
        t_rep = time.perf_counter()
        while time.perf_counter() - t_rep < 0.04:
            for _ in range(20):
                wl.kernel()
            torch.cuda.synchronize()
        wall1 = time.perf_counter()
        replay_note = "timed region + 40 ms untimed replay of the same launches (timed region shorter than 3 NVML polls)"
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    if clocks is not None and replay_note and clocks.get("window") == "timed region":
        clocks["window"] = replay_note

    if args.no_e2e:   # tuning runs: device-resident timing only
        if rank == 0:
            peak, _ = load_peaks()
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                              "ms_per_step": ms_per_step, "tuning_only": True, "exchange_check": check,
                              "flush_ms": getattr(wl, "last_flush_ms", None),
                              "config": {"workload": WORKLOAD_NAMES[cfg], "rows_per_gpu": B,
                                         "dt_column_sums_in_kernel": wl.want_col},
                              "roofline": {"kernel_ms": kern_ms, "frac": wl.bytes_per_row * B / (kern_ms * 1e-3) / 1e9 / peak}}),
                  flush=True)
        wl.close()
        return 0
    # ---- end to end through the host-buffer C-ABI call (pinned host buffers)
    # the pinned buffers are allocated (first-touched) and the calls issued from the CPUs NVML reports as
    # local to this GPU, so that on a multi-socket host the buffers do not sit behind the inter-socket link
    # (the pool's VMs expose one NUMA node, where this is a no-op: "16 of 16 cpus").  The affinity is restored
    # before the CPU-baseline leg.
    near = _NearGpuCpus(local_rank)
    near.__enter__()
    h_t = torch.empty((B, P), dtype=torch.float32).pin_memory()
    h_y = torch.empty((B, d), dtype=torch.float32).pin_memory()
    h_t.copy_(t)
    h_y.copy_(y)
    h_logp = torch.empty(B, dtype=torch.float32).pin_memory()
    h_dt = torch.empty((B, P), dtype=torch.float32).pin_memory() if bwd else None
    h_sum = ctypes.c_double(0.0)
    ft, desc, g_scale = wl.ft, wl.desc, wl.g_scale

    def e2e_step():
        if mdn and not bwd:
            raise SystemExit("--fwd-only has no host entry point for the MDN head")
        if mdn:
            _lib.check(lib.nfn_mdn_forward_backward_host(
                ft[1], d, _lib.ptr(h_t), _lib.ptr(h_y), B, None, ctypes.c_float(g_scale), _lib.ptr(h_logp),
                _lib.ptr(h_dt), ctypes.byref(h_sum), B))
        elif bwd:
            _lib.check(lib.nfn_chain_forward_backward_host(
                ctypes.byref(desc), _lib.ptr(h_t), _lib.ptr(h_y), B, None, ctypes.c_float(g_scale),
                _lib.ptr(h_logp), _lib.ptr(h_dt), ctypes.byref(h_sum), None, B))
        else:
            _lib.check(lib.nfn_chain_forward_host(ctypes.byref(desc), _lib.ptr(h_t), _lib.ptr(h_y), B,
                                                  _lib.ptr(h_logp), B))

    e2e_steps = max(3, min(K, 10))
    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize()
    parallel.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = parallel.max_over_ranks(time.perf_counter() - t0, device)
    parallel.barrier()
    e2e_value = B * world * e2e_steps / e2e_s
    near.__exit__(None, None, None)
    h2d = 4 * B * (P + d)
    d2h = 4 * B * (1 + (P if bwd else 0)) + 8
    lib.nfn_host_release()
    # the host path must reproduce the device path bit for bit
    same = bool(torch.equal(h_logp, logp.cpu()))
    specialized, use_peer, packed, want_col = wl.specialized, wl.use_peer, wl.packed, wl.want_col
    bytes_per_row = wl.bytes_per_row
    wl.close()
    del wl, t, y, logp, dt, h_t, h_dt
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs, a few steps each, on the same line (every rank takes part: weak scaling)
    others = {}
    if not args.rows and not args.fwd_only and not args.no_other_configs and cfg == "cfg2":
        for oc in ("cfg3", "cfg4", "cfg5"):
            try:
                others[oc] = measure_other_config(oc, args, device, rank, world, lib)
            except Exception as exc:  # noqa: BLE001 -- an extra must not take the headline down
                others[oc] = {"error": str(exc)[:200]}
        try:
            others["cfg4-train"] = cfg4_bayes_train_step(device, rank, world)
        except Exception as exc:  # noqa: BLE001
            others["cfg4-train"] = {"error": str(exc)[:300]}
        if world == 1:
            try:
                others["cfg1-pipeline"] = cfg1_pipeline(args.steps, 3, cpu=False)
            except Exception as exc:  # noqa: BLE001
                others["cfg1-pipeline"] = {"error": str(exc)[:200]}

    if rank != 0:
        return 0 if (check is None or check["ok"]) else 3
    peak, peak_src = load_peaks()
    achieved = bytes_per_row * B / (kern_ms * 1e-3) / 1e9
    cpu = cpu_baseline(cfg) if (world == 1 and not args.no_cpu_baseline) else None
    fused = fused_dense_step(cfg, device) if (world == 1 and not args.rows and not args.fwd_only) else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": arm_config(cfg, world, K, B, P, d, bwd, specialized, want_col, packed, use_peer, args.peer_blocking,
                             extra_warmup),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": ("nfn_mdn_forward_backward_host" if mdn else
                        "nfn_chain_forward_backward_host" if bwd else "nfn_chain_forward_host"),
                "host_equals_device_bitwise": same, "host_cpus": near.note},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": load_traffic(cfg), "peak_source": peak_src,
                     "kernel_ms": kern_ms, "kernel_ms_how": kern_how,
                     "algorithmic_bytes_per_launch": bytes_per_row * B},
        "clocks": clocks,
    }
    if check is not None:
        line["exchange_check"] = check
    if others:
        line["other_configs"] = others
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if fused is not None:
        line["fused_dense_step"] = fused
        line["fused_dense_mdn_step"] = fused_dense_mdn_step(device)
    print(json.dumps(line), flush=True)
    if check is not None and not check["ok"]:
        print("exchange_check FAILED: %r" % (check,), file=sys.stderr)
        return 3
    return 0


def fused_dense_step(cfg, device, steps=20):
    """SURVEY §8(f) rank 1, reported next to the headline (not part of it): the emitting Dense(16 -> P)
    layer + the flow chain, forward + backward, in ONE kernel (h[B,16] in, logp / dh / dW / db out; the
    parameter tensor t never exists in HBM).  Same rows and chain as the headline workload."""
    import torch

    from normalizingflownetwork_b200 import functional as F

    ft, d, tb, B, bwd = CONFIGS[cfg]
    if is_mdn(ft) or not bwd:
        return None
    P, H = param_size(ft, d, tb), 16
    g = torch.Generator(device=device).manual_seed(22)
    h = torch.tanh(torch.randn((B, H), generator=g, device=device))
    W = torch.randn((H, P), generator=g, device=device) * 0.3
    b = torch.zeros(P, device=device)
    y = torch.randn((B, d), generator=g, device=device)
    dW, db = torch.zeros((H, P), device=device), torch.zeros(P, device=device)
    try:
        def step():
            F.dense_chain_forward_backward(h, W, b, y, ft, d, tb, g_scale=-1.0 / B, dW=dW, dbias=db)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / steps
    except Exception as exc:  # noqa: BLE001 -- an optional extra must not take the headline down
        return {"error": str(exc)[:200]}
    return {"what": "Dense(16->P) layer + flow chain fused, fwd+bwd, one kernel (tcgen05 / TMEM GEMMs, h tiles staged by a separate warp group); t never in HBM",
            "rows": B, "hidden": H, "us_per_step": us, "samples_per_s": B / (us * 1e-6),
            "bytes_per_row": 4 * (2 * H + d + 1), "hbm_frac": 4 * (2 * H + d + 1) * B / (us * 1e-6) / 1e9 / load_peaks()[0]}


def fused_dense_mdn_step(device, steps=10):
    """The emitting Dense(16 -> 100) layer + the 20-component MDN head of BASELINE config 5, forward + backward, in
    ONE kernel (h[B,16] in, logp / dh / dW / db out), next to what it replaces at the estimator level: torch's
    GEMMs around the streaming head kernel (t = h W + b, dh = dt W^T, dW = h^T dt, db = sum dt).  Reported next to
    the headline, not part of it; reference MaximumLikelihoodNNEstimator.py:43 + DistributionLayers.py:196-212."""
    import torch

    from normalizingflownetwork_b200 import functional as F

    (_, K), d, _, B, _ = CONFIGS["cfg5"]
    H, P = 16, K * (2 * d + 1)
    g = torch.Generator(device=device).manual_seed(55)
    h = torch.tanh(torch.randn((B, H), generator=g, device=device))
    W = torch.randn((H, P), generator=g, device=device) * 0.3
    b = torch.randn(P, generator=g, device=device) * 0.1
    y = torch.randn((B, d), generator=g, device=device)
    dW, db = torch.zeros((H, P), device=device), torch.zeros(P, device=device)

    def timed(fn, n):
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

    def unfused():
        t = torch.addmm(b, h, W)
        _, dt, _ = F.mdn_forward_backward(t, y, K, d, g_scale=-1.0 / B)
        return dt @ W.t(), h.t() @ dt, dt.sum(0)

    try:
        us = timed(lambda: F.dense_mdn_forward_backward(h, W, b, y, K, d, g_scale=-1.0 / B, dW=dW, dbias=db), steps)
        us_unfused = timed(unfused, max(3, steps // 2))
    except Exception as exc:  # noqa: BLE001 -- an optional extra must not take the headline down
        return {"error": str(exc)[:200]}
    return {"what": "Dense(16->100) layer + 20-component MDN head fused, fwd+bwd, one kernel (3xTF32 mma.sync GEMMs, the "
                    "streaming head's row arithmetic); t / dt never in HBM",
            "rows": B, "hidden": H, "us_per_step": us, "samples_per_s": B / (us * 1e-6),
            "bytes_per_row": 4 * (2 * H + d + 1), "hbm_frac": 4 * (2 * H + d + 1) * B / (us * 1e-6) / 1e9 / load_peaks()[0],
            "unfused_layer_plus_head_us": us_unfused, "speedup_vs_unfused": us_unfused / us,
            "note": "issue-bound (8 warps per SM next to a 51 KB parameter tile), not memory-bound"}


def cfg1_pipeline(K, W, cpu=True):
    """BASELINE config 1 as the real pipeline: NormalizingFlowNetwork (3 radial flows, 1-D y,
    MLP (16,16) tanh) on gen_cosine_noise_data(2048): latency of log_pdf and of one Adam step
    (eager launches vs one CUDA-graph replay), next to the same pipeline on the host cores
    (torch-CPU MLP + the fp32 op-for-op restatement of the reference's TF graph)."""
    import torch

    from normalizingflownetwork_b200.estimators import NormalizingFlowNetwork
    from normalizingflownetwork_b200.simulation import gen_cosine_noise_data

    B = 2048
    x, y = gen_cosine_noise_data(B, noise_std=0.3, heterosced_noise=0.5)
    model = NormalizingFlowNetwork.build_function(n_dims=1, n_flows=3, hidden_sizes=(16, 16), activation="tanh")
    model.fit(x, y, batch_size=B, epochs=3, verbose=0)
    xd, yd = model._to_dev(x), model._to_dev(y)

    def timed(fn, n):
        for _ in range(W):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / n  # us

    us_logpdf = timed(lambda: model.log_pdf(xd, yd), K)
    us_step = timed(lambda: model.train_step(xd, yd), K)
    model.capture_train_step(B, 1, 1)
    us_step_graph = timed(lambda: model.train_step_graphed(xd, yd), K)
    model.capture_log_pdf(B, 1, 1)
    us_logpdf_graph = timed(lambda: model.log_pdf_graphed(xd, yd), K)
    # host-side pipeline: numpy in, numpy out, copies included
    t0 = time.perf_counter()
    for _ in range(K):
        model.log_pdf(x, y).cpu()
    us_logpdf_e2e = (time.perf_counter() - t0) * 1e6 / K
    t0 = time.perf_counter()
    for _ in range(K):
        model.log_pdf_graphed(x, y).cpu()
    us_logpdf_e2e_graph = (time.perf_counter() - t0) * 1e6 / K

    if not cpu:
        return _cfg1_line(B, K, W, us_logpdf, us_logpdf_graph, us_logpdf_e2e, us_logpdf_e2e_graph, us_step, us_step_graph, None)
    # CPU port of the same pipeline
    from oracle import flow_oracle as fo

    torch.set_num_threads(os.cpu_count() or 1)
    lin = [m.linear for m in model.net if hasattr(m, "linear")]
    Ws = [(l.weight.detach().cpu().clone().requires_grad_(True), l.bias.detach().cpu().clone().requires_grad_(True))
          for l in lin]
    params = [p for wb in Ws for p in wb]
    opt = torch.optim.Adam(params, lr=3e-3, eps=1e-7)
    xm, xs = model.x_mean.cpu(), model.x_std.cpu()
    ym, ys = model.y_mean.cpu(), model.y_std.cpu()
    xc, yc = torch.tensor(x), torch.tensor(y)

    def cpu_logp():
        h = (xc - xm) / (xs + 1e-8)
        for i, (w, b) in enumerate(Ws):
            h = h @ w.T + b
            if i < len(Ws) - 1:
                h = torch.tanh(h)
        return fo.chain_log_prob(h, (yc - ym) / ys, ["radial"] * 3, 1, True) - torch.sum(torch.log(ys))

    def cpu_step():
        opt.zero_grad()
        (-cpu_logp().mean()).backward()
        opt.step()

    def cpu_timed(fn, n):
        fn()
        t = time.perf_counter()
        for _ in range(n):
            fn()
        return (time.perf_counter() - t) * 1e6 / n

    with torch.no_grad():
        us_cpu_logpdf = cpu_timed(cpu_logp, 50)
    us_cpu_step = cpu_timed(cpu_step, 50)
    return _cfg1_line(B, K, W, us_logpdf, us_logpdf_graph, us_logpdf_e2e, us_logpdf_e2e_graph, us_step, us_step_graph,
                      {"kind": "port", "cores": os.cpu_count(), "log_pdf_us": us_cpu_logpdf,
                       "fit_step_us": us_cpu_step, "unit": "us",
                       "sample": "same 2048-row batch; torch-CPU MLP + fp32 restatement of the reference's TF graph"})


def _cfg1_line(B, K, W, us_logpdf, us_logpdf_graph, us_logpdf_e2e, us_logpdf_e2e_graph, us_step, us_step_graph, cpu):
    line = {
        "metric": "NFN config-1 pipeline latency (log_pdf, Adam fit step), batch 2048", "unit": "us",
        "higher_is_better": False, "n_gpus": 1, "steps": K, "warmup": W, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "NormalizingFlowNetwork 3 radial flows, 1-D y, MLP (16,16) tanh, "
                               "gen_cosine_noise_data(2048, 0.3, 0.5), batch 2048"},
        "value": us_step_graph, "ms_per_step": us_step_graph * 1e-3, "vs_baseline": None,
        "log_pdf_us": us_logpdf, "log_pdf_cuda_graph_us": us_logpdf_graph, "log_pdf_host_in_host_out_us": us_logpdf_e2e,
        "log_pdf_host_in_host_out_cuda_graph_us": us_logpdf_e2e_graph,
        "fit_step_eager_us": us_step, "fit_step_cuda_graph_us": us_step_graph,
        "samples_per_s_fit_graph": B / (us_step_graph * 1e-6), "samples_per_s_log_pdf": B / (us_logpdf * 1e-6),
        "note": "launch-latency-bound (197 KB of head traffic): no roofline claim",
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    return line


def run_cfg1_pipeline(args):
    print(json.dumps(cfg1_pipeline(args.steps, args.warmup)), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS) + ["cfg1-pipeline"])
    ap.add_argument("--rows", type=int, default=0, help="override rows per GPU (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="device-resident timing only (tuning sweeps)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: how the step sums its fp64 accumulators over ranks -- 'peer' = fused into the "
                         "kernel's last CTA over NVLink peer memory (default), 'nccl' = a separate all-reduce")
    ap.add_argument("--peer-blocking", action="store_true",
                    help="N > 1, peer exchange: wait for the peers in the last CTA of every launch (round-1 behaviour) "
                         "instead of the split-phase exchange")
    ap.add_argument("--force-peer", action="store_true",
                    help="N = 1: run the step through the peer-exchange entry point (a world of one) to measure the "
                         "protocol's local cost")
    ap.add_argument("--no-colsum", action="store_true",
                    help="leave the in-kernel dt column sums (bias gradient of the emitting layer) out (tuning)")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the other BASELINE configs that are reported next to the cfg2 headline")
    ap.add_argument("--fwd-only", action="store_true", help="time the forward-only kernel of the config (tuning)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.config == "cfg1-pipeline":
        return run_cfg1_pipeline(args)
    if args.impl == "reference":
        return run_reference(args)
    rc = run_ours(args)
    try:  # leave the process group cleanly (NCCL warns on stderr otherwise)
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        pass
    return rc


if __name__ == "__main__":
    sys.exit(main())
