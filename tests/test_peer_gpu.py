"""GPU: the fused in-kernel all-reduce over peer memory.

world = 1 runs on any single GPU (same kernel epilogue, the rank pushes to itself); the
world = 2 case spawns one process per GPU and is skipped when fewer than 2 GPUs are visible.
"""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG2 = (["planar", "radial", "affine"] * 3 + ["planar"], 2, True)


def _check_rank(rank, world, device, steps=6):
    from normalizingflownetwork_b200 import functional as F
    from normalizingflownetwork_b200 import parallel

    ft, d, tb = CFG2
    P = 48
    comm = parallel.PeerComm(P + 1, device)
    assert comm.world == world and comm.rank == rank
    g = torch.Generator(device=device).manual_seed(100 + rank)
    for step in range(steps):
        B = 50_000 + 1000 * rank + step  # ragged and different per rank
        t = torch.randn((B, P), generator=g, device=device) * 0.5
        y = torch.randn((B, d), generator=g, device=device)
        lp, dt, red = F.chain_forward_backward_peer(t, y, ft, d, tb, comm, g_scale=-1.0 / B, want_colsum=True)
        # reference: local sums, then a plain all-reduce
        local = torch.cat([dt.double().sum(0), lp.double().sum().reshape(1)])
        if world > 1:
            torch.distributed.all_reduce(local)
        assert torch.isfinite(red).all()
        # column sums are accumulated in fp32 inside a CTA (fp64 across CTAs and ranks)
        assert torch.allclose(red, local, rtol=1e-4, atol=1e-5), (step, (red - local).abs().max())
        # want_colsum = 0: only the logp slot is filled
        _, _, red2 = F.chain_forward_backward_peer(t, y, ft, d, tb, comm, g_scale=-1.0 / B)
        assert float(red2[:P].abs().max()) == 0.0
        assert torch.allclose(red2[P], local[P], rtol=1e-9)
    # stand-alone exchange (used by heads without the fused epilogue)
    v = torch.arange(P + 1, dtype=torch.float64, device=device) + rank
    out = comm.allreduce(v)
    expect = world * torch.arange(P + 1, dtype=torch.float64, device=device) + sum(range(world))
    assert torch.equal(out, expect)
    # a chain served by the generic kernel gets the same exchange from a second tiny launch
    F.set_option("force_generic", 1)
    try:
        t = torch.randn((3000, P), generator=g, device=device) * 0.5
        y = torch.randn((3000, d), generator=g, device=device)
        lp, dt, red = F.chain_forward_backward_peer(t, y, ft, d, tb, comm, want_colsum=True)
    finally:
        F.set_option("force_generic", 0)
    local = torch.cat([dt.double().sum(0), lp.double().sum().reshape(1)])
    if world > 1:
        torch.distributed.all_reduce(local)
    assert torch.allclose(red, local, rtol=1e-4, atol=1e-5)
    # split-phase mode: a launch only pushes; the next launch on the communicator (or flush) collects into the
    # EARLIER call's `reduced` tensor -- both kernel generations
    comm.set_deferred(True)
    for io in ("cpasync", "tma"):
        F.set_option("chain_io", io)
        pend = []
        for step in range(5):
            B = 20_000 + 777 * rank + 31 * step
            t = torch.randn((B, P), generator=g, device=device) * 0.5
            y = torch.randn((B, d), generator=g, device=device)
            red = torch.full((P + 1,), float("nan"), dtype=torch.float64, device=device)
            lp, dt, red = F.chain_forward_backward_peer(t, y, ft, d, tb, comm, g_scale=-1.0 / B, want_colsum=True,
                                                        reduced=red)
            local = torch.cat([dt.double().sum(0), lp.double().sum().reshape(1)])
            if world > 1:
                torch.distributed.all_reduce(local)
            pend.append((red, local))
            if step >= 1:  # the previous call's sums are complete once THIS launch has run
                torch.cuda.synchronize()
                r0, l0 = pend[step - 1]
                assert torch.allclose(r0, l0, rtol=1e-4, atol=1e-5), (io, step, (r0 - l0).abs().max())
        comm.flush()
        torch.cuda.synchronize()
        assert torch.allclose(pend[-1][0], pend[-1][1], rtol=1e-4, atol=1e-5), io
    F.set_option("chain_io", "auto")
    comm.set_deferred(False)
    comm.status()   # no exchange timed out
    comm.close()


def test_peer_allreduce_world1(cuda_device, nfn_lib):
    _check_rank(0, 1, cuda_device)


def _worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from normalizingflownetwork_b200 import parallel

    parallel.init_process_group("nccl")
    _check_rank(rank, world, torch.device("cuda", rank))
    torch.distributed.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs on one box")
def test_peer_allreduce_world2(nfn_lib):
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, 2, port)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
