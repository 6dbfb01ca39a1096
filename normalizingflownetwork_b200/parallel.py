"""Data-parallel plumbing for the hot path: one process per GPU, rows sharded, no
data-path collective.  The only exchange a training step needs is one small all-reduce of
[bias-gradient column sums of dt | sum of log-probs | MLP gradients]; the
density-grid scoring path needs none.  Works with the ``nccl`` backend on GPUs and with
``gloo`` on CPU tensors (used by the world_size-2 tests of this host-side logic).
"""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(
        os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for world size 1)."""
    rank, world, local_rank = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kwargs["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local_rank


def shard_rows(n_rows, rank, world):
    """Contiguous row block [lo, hi) of rank ``rank``; block sizes differ by at most one and
    every block boundary is the same on every rank."""
    base, rem = divmod(int(n_rows), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


class PackedAllReduce:
    """One flat float64 buffer = [segments...]; a single all-reduce(sum) per step.

    ``pack`` copies the given tensors (any float dtype) into the buffer, ``reduce`` sums it
    across ranks in place, ``unpack`` returns float64 views in the original shapes.  Messages
    are a few KB, i.e. latency-bound: one collective, no bucketing.
    """

    def __init__(self, shapes, device):
        self.shapes = [tuple(s) for s in shapes]
        self.sizes = [int(torch.Size(s).numel()) for s in self.shapes]
        self.buf = torch.zeros(sum(self.sizes), dtype=torch.float64, device=device)

    def pack(self, tensors):
        assert len(tensors) == len(self.sizes)
        off = 0
        for t, n in zip(tensors, self.sizes):
            self.buf[off: off + n].copy_(t.reshape(-1))
            off += n
        return self.buf

    def reduce(self):
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.buf, op=dist.ReduceOp.SUM)
        return self.buf

    def unpack(self):
        out, off = [], 0
        for s, n in zip(self.shapes, self.sizes):
            out.append(self.buf[off: off + n].view(s))
            off += n
        return out


class FlatGradReducer:
    """All parameter gradients of a model as views of ONE preallocated float32 buffer, plus a few float64 scalars
    (the step's sum of log-probs) carried as float32 (hi, lo) pairs: a data-parallel step is ``zero()``, the
    usual backward (autograd accumulates into the views in place), ``reduce()`` = one all-reduce(sum) of the flat
    buffer -- no concatenation, no dtype round trip, no per-parameter collectives.

    Every rank must call ``reduce()`` once per step, including a rank whose shard of the mini-batch is empty
    (it contributes zeros): the collective sequence is identical on all ranks by construction.
    Works on CUDA tensors over NCCL and on CPU tensors over gloo (tests/test_dist_gloo.py)."""

    def __init__(self, params, n_scalars=1):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters"
        dev = self.params[0].device
        self.sizes = [p.numel() for p in self.params]
        self.n_scalars = int(n_scalars)
        self.flat = torch.zeros(sum(self.sizes) + 2 * self.n_scalars, dtype=torch.float32, device=dev)
        off = 0
        self.views = []
        for p, n in zip(self.params, self.sizes):
            assert p.dtype == torch.float32 and p.device == dev
            self.views.append(self.flat[off: off + n].view(p.shape))
            off += n
        self._tail = self.flat[off:]
        self.attach()

    def attach(self):
        """(Re)install the views as the parameters' .grad (optimizers' zero_grad(set_to_none=True) drops them)."""
        for p, v in zip(self.params, self.views):
            p.grad = v

    def zero(self):
        self.flat.zero_()
        self.attach()

    def put_scalars(self, values64):
        """values64: float64 tensor [n_scalars]; stored as float32 (hi, lo) so that the sum keeps ~48 bits."""
        hi = values64.to(torch.float32)
        lo = (values64 - hi.to(torch.float64)).to(torch.float32)
        self._tail[0::2].copy_(hi)
        self._tail[1::2].copy_(lo)

    def get_scalars(self):
        return self._tail[0::2].to(torch.float64) + self._tail[1::2].to(torch.float64)

    def reduce(self):
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        return self.flat


class PeerComm:
    """Peer-memory communicator for the fused in-kernel all-reduce (include/nfn_b200.h,
    csrc/nfn_peer.cu).  One per process / GPU.  Each rank allocates one IPC-exportable region,
    the 64-byte cudaIpc handles are exchanged with one all_gather over the default process
    group (NCCL), and every rank maps its peers over NVLink.  Without an initialised process
    group the communicator is a world of one (useful on a single GPU: same code path).
    """

    def __init__(self, n_values, device):
        import ctypes

        from . import _lib

        self._lib = _lib
        self.lib = _lib.load()
        self.n_values = int(n_values)
        self.device = torch.device(device)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.mapped = []
        self.comm = None
        self.region = None
        err = None
        with torch.cuda.device(self.device):
            region = ctypes.c_void_p()
            handle = ctypes.create_string_buffer(64)
            try:
                _lib.check(self.lib.nfn_peer_alloc(self.world, self.n_values, ctypes.byref(region), handle))
                self.region = region
            except Exception as exc:  # noqa: BLE001 -- every rank must still reach the collectives below
                err = exc
            handles = [bytes(handle.raw)]
            if self.world > 1:
                mine = torch.tensor(list(handle.raw), dtype=torch.uint8, device=self.device)
                gathered = [torch.empty_like(mine) for _ in range(self.world)]
                dist.all_gather(gathered, mine)
                handles = [bytes(g.cpu().tolist()) for g in gathered]
            if err is None:
                try:
                    regions = (ctypes.c_void_p * self.world)()
                    for r in range(self.world):
                        if r == self.rank:
                            regions[r] = region.value
                        else:
                            m = ctypes.c_void_p()
                            _lib.check(self.lib.nfn_peer_open(handles[r], ctypes.byref(m)))
                            self.mapped.append(m)
                            regions[r] = m.value
                    comm = ctypes.c_void_p()
                    _lib.check(self.lib.nfn_peer_comm_create(self.world, self.rank, self.n_values, regions,
                                                             ctypes.byref(comm)))
                    self.comm = comm
                except Exception as exc:  # noqa: BLE001
                    err = exc
            if self.world > 1:
                # agreement + barrier in one: every rank has mapped every region, or nobody uses the comm
                ok = torch.tensor([0 if err is not None else 1], device=self.device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if int(ok.item()) == 0 and err is None:
                    err = RuntimeError("peer-memory set-up failed on another rank")
        if err is not None:
            self._release()
            raise RuntimeError("PeerComm set-up failed: %s" % err)

    def allreduce(self, values, out=None):
        """Sum device float64 ``values[n_values]`` over all ranks (one tiny kernel, no NCCL)."""
        if out is None:
            out = torch.empty(self.n_values, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            self._lib.check(self.lib.nfn_peer_allreduce(self.comm, self._lib.ptr(values), self._lib.ptr(out),
                                                        self._lib.current_stream(self.device)))
        return out

    def set_deferred(self, deferred=True):
        """Split-phase mode: a launch pushes its totals, the NEXT launch on this communicator (or flush())
        collects them into the earlier call's ``reduced`` tensor (include/nfn_b200.h)."""
        self._lib.check(self.lib.nfn_peer_set_deferred(self.comm, 1 if deferred else 0))

    def flush(self):
        """Complete the pending exchange of split-phase mode (one tiny kernel on the current stream)."""
        with torch.cuda.device(self.device):
            self._lib.check(self.lib.nfn_peer_flush(self.comm, self._lib.current_stream(self.device)))

    def status(self):
        """Synchronises; raises NfnError(NFN_ERR_PEER_TIMEOUT) if any exchange timed out on a peer."""
        with torch.cuda.device(self.device):
            self._lib.check(self.lib.nfn_peer_status(self.comm))

    def _release(self):
        with torch.cuda.device(self.device):
            if self.comm is not None:
                self.lib.nfn_peer_comm_destroy(self.comm)
            for m in self.mapped:
                self.lib.nfn_peer_close(m)
            if self.region is not None:
                self.lib.nfn_peer_free(self.region)
        self.comm, self.region, self.mapped = None, None, []

    def close(self):
        if getattr(self, "comm", None) is None:
            return
        err = None
        with torch.cuda.device(self.device):
            try:
                self.flush()
                self.status()
            except Exception as exc:  # noqa: BLE001 -- still unmap; every rank must reach the barrier
                err = exc
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier()  # nobody unmaps while a peer may still push
        self._release()
        if err is not None:
            raise err


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def max_over_ranks(value, device):
    """Max of a python float over ranks (timings are reported as the slowest rank)."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
