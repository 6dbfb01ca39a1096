"""Time the fused Dense(P)+chain kernels (tcgen05 / TMEM by default) on the ahead-of-time chains, B = 2^20, H = 16,
and check them against the unfused composition (float64 matmul -> chain kernel -> float64 matmuls).

    python tools/dense_time.py [--steps 50] [--rows 1048576] [--hidden 16]
Select the implementation with NFN_B200_DENSE_MMA=tc5|sync; A/B libraries with NFN_B200_LIB=<path>.
"""
import argparse
import os
import sys

sys.path.insert(0, os.getcwd())
import torch  # noqa: E402

from normalizingflownetwork_b200 import functional as F  # noqa: E402

CHAINS = {
    "cfg2": (["planar", "radial", "affine"] * 3 + ["planar"], 2),
    "r10d1": (["radial"] * 10, 1),
    "cfg4": (["radial"] * 5, 1),
    "cfg1": (["radial"] * 3, 1),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--rows", type=int, default=1 << 20)
    ap.add_argument("--hidden", type=int, default=16)
    ap.add_argument("--chains", default="cfg2,r10d1,cfg4,cfg1")
    ap.add_argument("--no-check", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, H = a.rows, a.hidden
    for name in a.chains.split(","):
        ft, d = CHAINS[name]
        P = F.chain_param_size(ft, d, True)
        g = torch.Generator(device=dev).manual_seed(22)
        h = torch.tanh(torch.randn((B, H), generator=g, device=dev))
        W = torch.randn((H, P), generator=g, device=dev) * 0.3
        b = torch.randn(P, generator=g, device=dev) * 0.1
        y = torch.randn((B, d), generator=g, device=dev)
        dW, db = torch.zeros((H, P), device=dev), torch.zeros(P, device=dev)

        def step_bwd():
            return F.dense_chain_forward_backward(h, W, b, y, ft, d, True, g_scale=-1.0 / B, dW=dW, dbias=db)

        def step_fwd():
            return F.dense_chain_forward(h, W, b, y, ft, d, True)

        out = {}
        for tag, fn in (("fwd+bwd", step_bwd), ("fwd", step_fwd)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            out[tag] = e0.elapsed_time(e1) * 1e3 / a.steps
        msg = "%-6s P=%3d H=%d rows=%d  fwd+bwd %7.1f us   fwd %7.1f us" % (name, P, H, B, out["fwd+bwd"], out["fwd"])
        if not a.no_check:
            n = min(B, 1 << 16)
            hs, ys = h[:n].contiguous(), y[:n].contiguous()
            dW.zero_(); db.zero_()
            res = F.dense_chain_forward_backward(hs, W, b, ys, ft, d, True, g_scale=-1.0 / n, dW=dW, dbias=db)
            logp, dh = res[0], res[1]
            t64 = (hs.double() @ W.double() + b.double())
            t = t64.float().contiguous()
            r2 = F.chain_forward_backward(t, ys, ft, d, True, g_scale=-1.0 / n)
            lp_ref, dt = r2[0], r2[1]
            dh_ref = (dt.double() @ W.double().t())
            dW_ref = hs.double().t() @ dt.double()
            db_ref = dt.double().sum(0)

            def rel(x, r):
                return float((x.double() - r.double()).abs().max() / r.double().abs().max().clamp_min(1e-30))
            msg += "   | logp abs err %.2e  dh rel %.2e  dW rel %.2e  db rel %.2e" % (
                float((logp.double() - lp_ref.double()).abs().max()), rel(dh, dh_ref), rel(dW, dW_ref), rel(db, db_ref))
        print(msg, flush=True)


if __name__ == "__main__":
    main()
