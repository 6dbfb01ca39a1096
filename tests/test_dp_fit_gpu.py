"""GPU, 2 ranks: the estimator-level data-parallel step (rows of the mini-batch sharded over ranks, fused
Dense(P)+chain kernel on every rank, ONE float32 all-reduce of the flat gradient buffer) equals the single-GPU step
on the whole mini-batch; a tail batch smaller than the world (an EMPTY shard on one rank) neither hangs nor
desynchronises the replicas.  Needs 2 GPUs on the box (skipped otherwise; `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(name):
    from normalizingflownetwork_b200.estimators import ESTIMATORS

    if name == "NFN":
        return ESTIMATORS["NFN"](1, n_flows=3, hidden_sizes=(16, 16), activation="tanh")
    if name == "MDN":
        return ESTIMATORS["MDN"](1, n_centers=5, hidden_sizes=(16, 16), activation="tanh")
    return ESTIMATORS["bayesian_NFN"](1, kl_weight_scale=1e-3, n_flows=3, hidden_sizes=(16,), activation="tanh",
                                      n_train_draws=4)


def _data(n):
    from normalizingflownetwork_b200.simulation import gen_cosine_noise_data

    return gen_cosine_noise_data(n, noise_std=0.3, heterosced_noise=0.5)


def _one_step(model, x, y, lo, hi, gb):
    """Normalisation statistics from the WHOLE data set on every rank, lazy layers materialised, then one step."""
    model._assign_data_normalization(x, y)
    model._set_noise(0.0)
    with torch.no_grad():
        model.params_from_x(x[:2])
    if model.optimizer is None:
        model.optimizer = torch.optim.SGD(model.parameters(), lr=0.0)   # gradients are what is compared
    xd, yd = model._to_dev(x), model._to_dev(y)
    loss = model.train_step(xd[lo:hi], yd[lo:hi], global_batch=gb)
    grads = [p.grad.detach().clone().cpu() for p in model.parameters() if p.requires_grad]
    return float(loss), grads


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from normalizingflownetwork_b200 import parallel

    torch.cuda.set_device(rank)
    parallel.init_process_group("nccl")
    x, y = _data(1001)
    res = {}
    for name in ("NFN", "MDN", "bayesian_NFN"):
        model = _make(name)
        lo, hi = parallel.shard_rows(1001, rank, world)
        res[name] = _one_step(model, x, y, lo, hi, 1001)
    # a fit whose tail batch has ONE row: rank 1's shard of it is empty
    xs, ys = _data(1025)
    m = _make("NFN")
    m.fit(xs, ys, batch_size=512, epochs=3, verbose=0, shuffle=False)
    w = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
    w0 = w.clone()
    torch.distributed.broadcast(w0, 0)
    res["tail_fit"] = (list(m.history), bool(torch.equal(w, w0)))
    if rank == 0:
        q.put(res)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs on one box")
def test_dp_step_equals_single_gpu_step(nfn_lib):
    import torch.multiprocessing as mp

    # single-GPU reference first (no process group in this process)
    x, y = _data(1001)
    ref = {name: _one_step(_make(name), x, y, 0, 1001, 1001) for name in ("NFN", "MDN", "bayesian_NFN")}
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    for name in ("NFN", "MDN", "bayesian_NFN"):
        loss, grads = got[name]
        rloss, rgrads = ref[name]
        assert loss == pytest.approx(rloss, rel=2e-5), name
        for g, r in zip(grads, rgrads):
            scale = max(1.0, float(r.abs().max()))
            assert float((g - r).abs().max()) <= 2e-4 * scale, name
    hist, identical = got["tail_fit"]
    assert identical and all(np.isfinite(hist)) and len(hist) == 3
