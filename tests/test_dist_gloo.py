"""CPU, world_size 2 over gloo: the host-side data-parallel plumbing (row sharding, the one
packed all-reduce of [dt column sums | sum logp], max-over-ranks timing)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from normalizingflownetwork_b200 import parallel
from oracle import analytic_np as an


def test_shard_rows_partitions_exactly():
    for n in (0, 1, 7, 8, 1 << 20, (1 << 26) + 3):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_rows(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = parallel.init_process_group("gloo")
    assert (r, w) == (rank, world)
    ft, d, tb = ["radial"] * 3, 1, True
    rng = np.random.default_rng(22)  # every rank draws the same global batch, then takes its shard
    B, P = 1001, 11
    t = rng.normal(0, 0.5, (B, P)).astype(np.float32)
    y = rng.normal(0, 1.0, (B, d)).astype(np.float32)
    lo, hi = parallel.shard_rows(B, rank, world)
    # the oracle stands in for the device kernel here: this test covers the exchange, not the math
    lp, dt, _ = an.chain_forward_backward(t[lo:hi], y[lo:hi], ft, d, tb, upstream=-1.0 / B)
    packed = parallel.PackedAllReduce([(P,), (1,)], torch.device("cpu"))
    packed.pack([torch.tensor(dt.sum(0)), torch.tensor([lp.sum()])])
    packed.reduce()
    col, lsum = packed.unpack()
    slowest = parallel.max_over_ranks(float(rank + 1), torch.device("cpu"))
    parallel.barrier()
    if rank == 0:
        out.put((col.numpy().copy(), float(lsum[0]), slowest))
    dist.destroy_process_group()


def test_packed_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    col, lsum, slowest = out.get()
    rng = np.random.default_rng(22)
    t = rng.normal(0, 0.5, (1001, 11)).astype(np.float32)
    y = rng.normal(0, 1.0, (1001, 1)).astype(np.float32)
    lp, dt, _ = an.chain_forward_backward(t, y, ["radial"] * 3, 1, True, upstream=-1.0 / 1001)
    np.testing.assert_allclose(col, dt.sum(0), rtol=1e-10, atol=1e-12)
    assert lsum == pytest.approx(lp.sum(), rel=1e-12)
    assert slowest == 2.0


# ----------------------------------------------------------------------------- flat gradient reducer
def _tiny_model(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(3, 8), torch.nn.Tanh(), torch.nn.Linear(8, 2))


def _reducer_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    parallel.init_process_group("gloo")
    model = _tiny_model(5)
    red = parallel.FlatGradReducer(model.parameters(), n_scalars=1)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(9)
    x = torch.randn((9, 3), generator=g)
    y = torch.randn((9, 2), generator=g)
    logs = []
    # three "mini-batches": 4 rows, 4 rows and a tail of ONE row -- rank 1's shard of the tail is empty, and it
    # must still take part in that step's all-reduce (a skipped rank pairs its next collective with
    # the peers' current one)
    for lo_b, hi_b in ((0, 4), (4, 8), (8, 9)):
        gb = hi_b - lo_b
        a, b = parallel.shard_rows(gb, rank, world)
        xb, yb = x[lo_b + a: lo_b + b], y[lo_b + a: lo_b + b]
        red.zero()
        local = torch.zeros(1, dtype=torch.float64)
        if xb.shape[0] > 0:
            sq = ((model(xb) - yb) ** 2).sum()
            (sq / gb).backward()
            local += sq.detach().double()
        red.put_scalars(local)
        red.reduce()
        logs.append(float(red.get_scalars()[0]))
        opt.step()
    if rank == 0:
        out.put((logs, [p.detach().clone().numpy() for p in model.parameters()]))
    else:
        out.put(("rank1", [p.detach().clone().numpy() for p in model.parameters()]))
    parallel.barrier()
    dist.destroy_process_group()


def test_flat_grad_reducer_world2_gloo_with_an_empty_shard():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_reducer_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = [out.get(), out.get()]
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    logs = [g for g in got if g[0] != "rank1"][0]
    other = [g for g in got if g[0] == "rank1"][0]
    # single-process reference: the same three steps on the whole mini-batches
    model = _tiny_model(5)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(9)
    x = torch.randn((9, 3), generator=g)
    y = torch.randn((9, 2), generator=g)
    ref_logs = []
    for lo_b, hi_b in ((0, 4), (4, 8), (8, 9)):
        opt.zero_grad()
        sq = ((model(x[lo_b:hi_b]) - y[lo_b:hi_b]) ** 2).sum()
        (sq / (hi_b - lo_b)).backward()
        ref_logs.append(float(sq))
        opt.step()
    np.testing.assert_allclose(logs[0], ref_logs, rtol=1e-6)
    for a, b, c in zip(logs[1], other[1], model.parameters()):
        np.testing.assert_array_equal(a, b)                      # replicas stay bit-identical
        np.testing.assert_allclose(a, c.detach().numpy(), rtol=1e-5, atol=1e-6)   # and follow the 1-process run
