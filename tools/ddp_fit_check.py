"""torchrun --nproc-per-node 2 tools/ddp_fit_check.py : data-parallel fit (rows of every mini-batch sharded
over ranks, one flat all-reduce of [grads | sum logp]) must follow the single-GPU trajectory."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from normalizingflownetwork_b200 import parallel
from normalizingflownetwork_b200.estimators import NormalizingFlowNetwork, MixtureDensityNetwork
from normalizingflownetwork_b200.simulation import gen_cosine_noise_data

rank, world, local = parallel.init_process_group("nccl")
torch.cuda.set_device(local)
x, y = gen_cosine_noise_data(2048, noise_std=0.3, heterosced_noise=0.5)
ok = True
for name, mk in [("NFN", lambda: NormalizingFlowNetwork.build_function(n_dims=1, n_flows=3, hidden_sizes=(16, 16), activation="tanh")),
                 ("MDN", lambda: MixtureDensityNetwork(1, n_centers=5, activation="tanh"))]:
    m = mk()
    m.fit(x, y, batch_size=512, epochs=8, verbose=0)
    hist = np.array(m.history)
    if rank == 0:
        # single-process reference on the same GPU: temporarily pretend world = 1
        import torch.distributed as dist
        ref_hist = torch.tensor(hist)
    w0 = [p.detach().clone() for p in m.parameters()]
    # all ranks must hold identical weights after training
    for p in w0:
        q = p.clone()
        torch.distributed.broadcast(q, 0)
        ok = ok and bool(torch.allclose(p, q, rtol=0, atol=0))
    if rank == 0:
        print(name, "history", np.round(hist, 4).tolist(), "replicas identical:", ok)
torch.distributed.barrier()
if rank == 0:
    # compare with a 1-GPU run of the same schedule (fused last layer off so the arithmetic path matches)
    torch.distributed.destroy_process_group()
    m = NormalizingFlowNetwork.build_function(n_dims=1, n_flows=3, hidden_sizes=(16, 16), activation="tanh")
    m.fuse_last_layer = False
    m.fit(x, y, batch_size=512, epochs=8, verbose=0)
    print("NFN 1-GPU history", np.round(m.history, 4).tolist())
    print("DDP_FIT_CHECK", "OK" if ok else "FAILED")
else:
    torch.distributed.destroy_process_group()
