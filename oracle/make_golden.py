"""Generate the frozen golden fixtures under ``tests/golden/``.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (no TF/TFP in this image): the vectors are
produced by the float64 literal restatement (``flow_oracle.py``) and are written only
if the independent closed-form NumPy oracle (``analytic_np.py``) and, for the
``tf.ones`` known-answer cases, the 50-digit mpmath evaluation agree with it.

    python -m oracle.make_golden        # rewrites tests/golden/*.json deterministically
"""
import json
import os

import numpy as np
import torch

from oracle import analytic_np as an
from oracle import flow_oracle as fo
from oracle import known_answers_mp as mpo

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")

# (name, flow_types, n_dims, trainable_base_dist) -- the BASELINE.json canonical chains
#  plus the chains the reference's tests build.
CHAINS = [
    ("cfg1_radial3_d1", ["radial"] * 3, 1, True),
    ("cfg2_mixed10_d2", ["planar", "radial", "affine"] * 3 + ["planar"], 2, True),
    ("cfg3_rp8_d4", ["radial", "planar"] * 8, 4, True),
    ("cfg4_radial5_d1", ["radial"] * 5, 1, True),
    ("test_pra_d1_nobase", ["planar", "radial", "affine"], 1, False),
    ("test_pra_d3_base", ["planar", "radial", "affine"], 3, True),
    ("test_rp_d1_nobase", ["radial", "planar"], 1, False),
    ("test_rp_d2_base", ["radial", "planar"], 2, True),
    ("k0_d2_base", [], 2, True),
    ("k0_d1_nobase_planar1", ["planar"], 1, False),
    ("affine_only_d5", ["affine", "affine"], 5, False),
    ("planar_d8_base", ["planar", "planar", "radial"], 8, True),
]
MDN_CASES = [("cfg5_mdn_k20_d2", 20, 2), ("mdn_k3_d1", 3, 1), ("mdn_k5_d5", 5, 5)]
KMN_CASES = [("kmn_m20_d1", 10, 1, (0.3, 0.7)), ("kmn_m20_d2", 10, 2, (0.3, 0.7))]


def f32(a):
    return np.asarray(a, dtype=np.float32)


def tolist(a):
    return np.asarray(a, dtype=np.float64).tolist()


def known_answers():
    out = {"single_flow": [], "layer": [], "mdn": []}
    for name in ("planar", "radial", "affine"):
        for d in (1, 4):
            for zval in (0.0, 1.0):
                th = [1.0] * mpo.psize(name, d)
                z2, fl = mpo.flow_step(name, [mpo.mp.mpf(x) for x in th], [mpo.mp.mpf(zval)] * d)
                # cross-check with the torch literal restatement in float64
                bij = fo.ORACLE_FLOWS[name](torch.ones(1, len(th), dtype=torch.float64), d)
                zt = torch.full((1, d), zval, dtype=torch.float64)
                assert abs(float(bij.fldj(zt)[0]) - float(fl)) < 1e-13
                assert np.allclose(bij.forward(zt)[0].numpy(), [float(v) for v in z2], atol=1e-13)
                out["single_flow"].append(
                    {"flow": name, "n_dims": d, "t": 1.0, "z": zval,
                     "forward": [float(v) for v in z2], "fldj": float(fl)}
                )
    layer_cases = [
        (["radial", "planar"], 1, False, [(1.0, 0.0), (1.0, 0.5), (0.0, 0.0), (0.0, 0.5)]),
        (["radial", "planar"], 2, True, [(1.0, 0.0), (1.0, 0.5), (0.0, 0.0)]),
        (["planar", "radial", "affine"], 1, False, [(1.0, 0.0), (1.0, 0.5)]),
        (["planar", "radial", "affine"], 3, True, [(1.0, 0.0), (0.0, 0.25)]),
        (["radial"] * 3, 1, True, [(1.0, 0.0), (0.0, 0.0), (-1.0, 2.0)]),
        ([], 2, True, [(0.0, 0.0), (1.0, 0.5)]),
    ]
    for ft, d, tb, pts in layer_cases:
        P = fo.chain_param_size(ft, d, tb)
        for tval, yval in pts:
            lp = mpo.chain_log_prob([tval] * P, [yval] * d, ft, d, tb)
            lt = fo.chain_log_prob(
                torch.full((1, P), tval, dtype=torch.float64),
                torch.full((1, d), yval, dtype=torch.float64), ft, d, tb)
            la = an.chain_forward_backward(np.full((1, P), tval), np.full((1, d), yval), ft, d, tb,
                                           need_grad=False)
            assert abs(float(lt[0]) - float(lp)) < 1e-12 and abs(la[0] - float(lp)) < 1e-12
            out["layer"].append({"flow_types": ft, "n_dims": d, "trainable_base_dist": tb,
                                 "t": tval, "y": yval, "log_prob": float(lp)})
    for K, d, tval, yval in [(5, 1, 1.0, 0.0), (3, 1, 0.0, 0.5), (5, 3, 1.0, 0.5), (20, 2, 0.5, -0.25)]:
        P = 2 * K * d + K
        lp = mpo.mdn_log_prob([tval] * P, [yval] * d, K, d)
        lt = fo.mdn_log_prob(torch.full((1, P), tval, dtype=torch.float64),
                             torch.full((1, d), yval, dtype=torch.float64), K, d)
        assert abs(float(lt[0]) - float(lp)) < 1e-12
        out["mdn"].append({"n_centers": K, "n_dims": d, "t": tval, "y": yval, "log_prob": float(lp)})
    return out


def chain_vectors(rng, B=12):
    out = []
    for name, ft, d, tb in CHAINS:
        P = fo.chain_param_size(ft, d, tb)
        for sigma in (0.5, 1.0):
            t = f32(rng.normal(0.0, sigma, size=(B, P)))
            y = f32(rng.normal(0.0, 1.0, size=(B, d)))
            t[0, :] = 0.0  # t=0 row: identity for radial/affine, NOT for planar (App. B.4)
            y[1, :] = 0.0
            if "radial" in ft:  # sign(0)=0 sub-gradient: put y exactly on a radial centre
                k_first_radial = ft.index("radial")
                if k_first_radial == 0:
                    offs, _ = an.layout(ft, d, tb)
                    y[2, :] = t[2, offs[0] + 2 : offs[0] + 2 + d]
            up = f32(rng.normal(0.0, 1.0, size=(B,)))
            t64, y64 = torch.tensor(t, dtype=torch.float64), torch.tensor(y, dtype=torch.float64)
            logp, dt, dy = fo.with_grad(fo.chain_log_prob, t64, y64, ft, d, tb,
                                        upstream=torch.tensor(up, dtype=torch.float64), want_dy=True)
            la, dta, dya = an.chain_forward_backward(t, y, ft, d, tb, upstream=up)
            assert np.allclose(la, logp.numpy(), rtol=1e-12, atol=1e-12), name
            assert np.allclose(dta, dt.numpy(), rtol=1e-9, atol=1e-11), name
            assert np.allclose(dya, dy.numpy(), rtol=1e-9, atol=1e-11), name
            # broadcast y [1,d]
            lb = fo.chain_log_prob(t64, y64[3:4], ft, d, tb)
            out.append({"name": name, "flow_types": ft, "n_dims": d, "trainable_base_dist": tb,
                        "sigma": sigma, "t": tolist(t), "y": tolist(y), "upstream": tolist(up),
                        "log_prob": tolist(logp), "dt": tolist(dt), "dy": tolist(dy),
                        "log_prob_y_row3_broadcast": tolist(lb)})
    return out


def mixture_vectors(rng, B=12):
    out = {"mdn": [], "kmn": []}
    for name, K, d in MDN_CASES:
        P = 2 * K * d + K
        for sigma in (0.5, 2.0):
            t = f32(rng.normal(0.0, sigma, size=(B, P)))
            y = f32(rng.normal(0.0, 1.0, size=(B, d)))
            t[0, :] = 0.0
            up = f32(rng.normal(0.0, 1.0, size=(B,)))
            t64, y64 = torch.tensor(t, dtype=torch.float64), torch.tensor(y, dtype=torch.float64)
            logp, dt, dy = fo.with_grad(fo.mdn_log_prob, t64, y64, K, d,
                                        upstream=torch.tensor(up, dtype=torch.float64), want_dy=True)
            la, dta, dya = an.mdn_forward_backward(t, y, K, d, upstream=up)
            assert np.allclose(la, logp.numpy(), rtol=1e-12, atol=1e-12)
            assert np.allclose(dta, dt.numpy(), rtol=1e-9, atol=1e-12)
            assert np.allclose(dya, dy.numpy(), rtol=1e-9, atol=1e-12)
            out["mdn"].append({"name": name, "n_centers": K, "n_dims": d, "sigma": sigma,
                               "t": tolist(t), "y": tolist(y), "upstream": tolist(up),
                               "log_prob": tolist(logp), "dt": tolist(dt), "dy": tolist(dy)})
    for name, nc, d, init in KMN_CASES:
        M = nc * len(init)
        t = f32(rng.normal(0.0, 1.0, size=(B, M)))
        y = f32(rng.normal(0.0, 1.0, size=(B, d)))
        locs = f32(np.tile(rng.normal(0.0, 1.0, size=(nc, d)), (len(init), 1)))
        svars = f32(rng.normal(0.0, 0.3, size=(len(init),)))
        up = f32(rng.normal(0.0, 1.0, size=(B,)))
        sv64 = torch.tensor(svars, dtype=torch.float64, requires_grad=True)
        scales = fo.kmn_scales(sv64, nc, init)  # negative for init=0.3 (App. B.7)
        t64 = torch.tensor(t, dtype=torch.float64, requires_grad=True)
        y64 = torch.tensor(y, dtype=torch.float64, requires_grad=True)
        logp = fo.kmn_log_prob(t64, y64, torch.tensor(locs, dtype=torch.float64), scales)
        dt, dy, dsv = torch.autograd.grad(logp, [t64, y64, sv64],
                                          grad_outputs=torch.tensor(up, dtype=torch.float64))
        la, dta, dsa, dya = an.kmn_forward_backward(t, y, locs, scales.detach().numpy(), upstream=up)
        assert np.allclose(la, logp.detach().numpy(), rtol=1e-12, atol=1e-12)
        assert np.allclose(dta, dt.numpy(), rtol=1e-9, atol=1e-12)
        assert np.allclose(dya, dy.numpy(), rtol=1e-9, atol=1e-12)
        out["kmn"].append({"name": name, "n_centers": nc, "n_dims": d, "init_scales": list(init),
                           "t": tolist(t), "y": tolist(y), "locs": tolist(locs),
                           "scale_vars": tolist(svars), "scales": tolist(scales.detach()),
                           "upstream": tolist(up), "log_prob": tolist(logp.detach()),
                           "dt": tolist(dt), "dy": tolist(dy), "dscales": tolist(dsa),
                           "dscale_vars": tolist(dsv)})
    return out


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    rng = np.random.default_rng(22)  # 22 is the reference's seed (BaseEstimator.py:12)
    files = {
        "known_answers.json": known_answers(),
        "chain_vectors.json": chain_vectors(rng),
        "mixture_vectors.json": mixture_vectors(rng),
    }
    for fn, obj in files.items():
        with open(os.path.join(GOLDEN, fn), "w") as f:
            json.dump(obj, f)
        print("wrote", fn, os.path.getsize(os.path.join(GOLDEN, fn)), "bytes")


if __name__ == "__main__":
    main()
