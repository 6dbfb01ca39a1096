"""Run the reference's OWN flow / distribution-layer code on the frozen golden inputs and
freeze what it returns as ``tests/golden/reference_run.json``.

TEST INFRASTRUCTURE ONLY.  Runs only where ``/root/reference`` exists (this container):

    python -m oracle.make_reference_run            # rewrite the fixture
    python -m oracle.make_reference_run --check    # recompute and diff against the fixture

The reference's ``estimators/normalizing_flows/{Planar,Radial,Affine}Flow.py`` and
``estimators/DistributionLayers.py`` are imported unmodified through ``oracle/tf_shim.py``
(torch-CPU float64 stand-ins for the TF ops and the TFP glue classes; TensorFlow itself is not
installable here).  For every case of ``tests/golden/{known_answers,chain_vectors,
mixture_vectors}.json`` the reference code is given the same ``t``, ``y`` and upstream cotangent;
its ``log_prob`` and the autograd gradients through it are written out, and the script refuses to
write if they differ from the oracle's frozen values by more than 1e-12 (values) / 1e-9 (gradients).
Reference entry points exercised:
  FLOWS[name](t, d).forward / ._forward_log_det_jacobian        (tests/test_flows.py:19-41)
  InverseNormalizingFlowLayer(flow_types, d, trainable)(t).log_prob(y), .get_total_param_size(),
  ._get_bijector(...).bijectors                                  (DistributionLayers.py:215-294)
  GaussianMixtureLayer(K, d)(t).log_prob(y)                      (DistributionLayers.py:174-212)
  GaussianKernelsLayer(nc, d, True, init)(t).log_prob(y), scale_model, locs (DistributionLayers.py:74-133)
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

from oracle import tf_shim

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
OUT = os.path.join(GOLDEN, "reference_run.json")
F64 = torch.float64


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def t64(a, grad=False):
    return torch.tensor(np.asarray(a, dtype=np.float64), dtype=F64, requires_grad=grad)


def lst(x):
    return x.detach().numpy().astype(np.float64).tolist()


def close(a, b, rtol, atol, what):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else a
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if not np.allclose(a, b, rtol=rtol, atol=atol):
        raise AssertionError(f"reference code and oracle disagree: {what}: max |diff| = {np.max(np.abs(a - b))}")


def run_single_flows(FLOWS, ka):
    out = []
    for c in ka["single_flow"]:
        d = c["n_dims"]
        P = FLOWS[c["flow"]].get_param_size(d)
        # tests/test_flows.py:19-29: t = tf.ones((batch, param_size)), z constant
        flow = FLOWS[c["flow"]](torch.full((3, P), c["t"], dtype=F64), d)
        z = torch.full((3, d), c["z"], dtype=F64)
        fwd, fldj = flow.forward(z), flow._forward_log_det_jacobian(z)
        assert flow.forward_min_event_ndims == 1
        assert tuple(fwd.shape) == (3, d) and tuple(fldj.shape) == (3,)
        close(fwd[0], c["forward"], 1e-12, 1e-14, f"single {c['flow']} forward")
        close(fldj[0], c["fldj"], 1e-12, 1e-14, f"single {c['flow']} fldj")
        out.append({"flow": c["flow"], "n_dims": d, "t": c["t"], "z": c["z"],
                    "forward": lst(fwd[0]), "fldj": float(fldj[0].detach())})
    return out


def run_layer_known(DL, ka):
    out = []
    for c in ka["layer"]:
        ft, d, tb = c["flow_types"], c["n_dims"], c["trainable_base_dist"]
        layer = DL.InverseNormalizingFlowLayer(ft, d, trainable_base_dist=tb)
        P = layer.get_total_param_size()
        lp = layer(torch.full((2, P), c["t"], dtype=F64)).log_prob(torch.full((2, d), c["y"], dtype=F64))
        close(lp[0], c["log_prob"], 1e-12, 1e-14, f"layer known answer {ft}")
        out.append({"flow_types": ft, "n_dims": d, "trainable_base_dist": tb, "t": c["t"], "y": c["y"],
                    "param_size": P, "log_prob": float(lp[0].detach())})
    for c in ka["mdn"]:
        K, d = c["n_centers"], c["n_dims"]
        layer = DL.GaussianMixtureLayer(K, d)
        P = layer.get_total_param_size()
        lp = layer(torch.full((2, P), c["t"], dtype=F64)).log_prob(torch.full((2, d), c["y"], dtype=F64))
        close(lp[0], c["log_prob"], 1e-12, 1e-14, f"mdn known answer K={K}")
        out.append({"mdn_n_centers": K, "n_dims": d, "t": c["t"], "y": c["y"], "param_size": P,
                    "log_prob": float(lp[0].detach())})
    return out


def run_chains(FLOWS, DL, cv):
    out = []
    for c in cv:
        ft, d, tb = c["flow_types"], c["n_dims"], c["trainable_base_dist"]
        layer = DL.InverseNormalizingFlowLayer(ft, d, trainable_base_dist=tb)
        t, y, up = t64(c["t"], True), t64(c["y"], True), t64(c["upstream"])
        assert layer.get_total_param_size() == t.shape[-1]
        logp = layer(t).log_prob(y)
        dt, dy = torch.autograd.grad(logp, [t, y], grad_outputs=up)
        lb = layer(t).log_prob(y[3:4])  # one event broadcast against the batch (BaseEstimator.py:71-86 callers)
        # DistributionLayers.py:267-278: the bijector list is in REVERSED flow_types order
        bij = DL.InverseNormalizingFlowLayer._get_bijector(
            t[..., 2 * d:] if tb else t, ft, d)
        order = [type(b).__name__ for b in bij.bijectors]
        want = [FLOWS[f].__name__ for f in reversed(ft)]
        assert order == want and bij.inverse_min_event_ndims == 1
        tag = f"{c['name']} sigma={c['sigma']}"
        close(logp.detach(), c["log_prob"], 1e-12, 1e-12, tag + " log_prob")
        close(dt, c["dt"], 1e-9, 1e-11, tag + " dt")
        close(dy, c["dy"], 1e-9, 1e-11, tag + " dy")
        close(lb.detach(), c["log_prob_y_row3_broadcast"], 1e-12, 1e-12, tag + " broadcast")
        out.append({"name": c["name"], "sigma": c["sigma"], "bijector_order": order,
                    "log_prob": lst(logp), "dt": lst(dt), "dy": lst(dy),
                    "log_prob_y_row3_broadcast": lst(lb)})
    return out


def run_mixtures(DL, mv):
    out = {"mdn": [], "kmn": []}
    for c in mv["mdn"]:
        K, d = c["n_centers"], c["n_dims"]
        layer = DL.GaussianMixtureLayer(K, d)
        t, y, up = t64(c["t"], True), t64(c["y"], True), t64(c["upstream"])
        assert layer.get_total_param_size() == t.shape[-1]
        logp = layer(t).log_prob(y)
        dt, dy = torch.autograd.grad(logp, [t, y], grad_outputs=up)
        tag = f"{c['name']} sigma={c['sigma']}"
        close(logp.detach(), c["log_prob"], 1e-12, 1e-12, tag + " log_prob")
        close(dt, c["dt"], 1e-9, 1e-12, tag + " dt")
        close(dy, c["dy"], 1e-9, 1e-12, tag + " dy")
        out["mdn"].append({"name": c["name"], "sigma": c["sigma"], "log_prob": lst(logp),
                           "dt": lst(dt), "dy": lst(dy)})
    for c in mv["kmn"]:
        nc, d, init = c["n_centers"], c["n_dims"], tuple(c["init_scales"])
        layer = DL.GaussianKernelsLayer(nc, d, trainable_scale=True, init_scales=init)
        assert layer.get_total_param_size() == nc * len(init)
        sv = t64(c["scale_vars"], True)
        layer.scale_model.layers[0].variable = sv  # the VariableLayer's weight (DistributionLayers.py:80-85)
        # what set_center_points does with the chosen centres (DistributionLayers.py:169-170)
        layer.locs.assign(np.float32(np.asarray(c["locs"])))
        layer.locs = sys.modules["tensorflow"].expand_dims(layer.locs, axis=0)
        t, y, up = t64(c["t"], True), t64(c["y"], True), t64(c["upstream"])
        logp = layer(t).log_prob(y)
        dt, dy, dsv = torch.autograd.grad(logp, [t, y, sv], grad_outputs=up)
        scales = layer.scale_model(0.0)
        close(scales.detach(), c["scales"], 1e-12, 1e-14, c["name"] + " scales")
        close(logp.detach(), c["log_prob"], 1e-12, 1e-12, c["name"] + " log_prob")
        close(dt, c["dt"], 1e-9, 1e-12, c["name"] + " dt")
        close(dy, c["dy"], 1e-9, 1e-12, c["name"] + " dy")
        close(dsv, c["dscale_vars"], 1e-9, 1e-12, c["name"] + " dscale_vars")
        out["kmn"].append({"name": c["name"], "scales": lst(scales), "log_prob": lst(logp),
                           "dt": lst(dt), "dy": lst(dy), "dscale_vars": lst(dsv)})
    return out


def compute():
    FLOWS, DL = tf_shim.load_reference()
    ka, cv, mv = load("known_answers.json"), load("chain_vectors.json"), load("mixture_vectors.json")
    try:
        return _run_all(FLOWS, DL, ka, cv, mv)
    finally:
        tf_shim.uninstall()


def _run_all(FLOWS, DL, ka, cv, mv):
    return {
        "provenance": {
            "what": "outputs of the reference's own estimators/normalizing_flows/*.py and "
                    "estimators/DistributionLayers.py, imported unmodified from /root/reference and executed "
                    "on torch-CPU float64 stand-ins for the TF ops and TFP glue classes (oracle/tf_shim.py)",
            "inputs": "t, y, upstream of tests/golden/{known_answers,chain_vectors,mixture_vectors}.json "
                      "(same order; matched by name and sigma)",
            "generator": "python -m oracle.make_reference_run",
            "pins": "the reference's in-repo formulas, constants, slicing and ordering",
            "does_not_pin": "TFP's own Chain/Invert/TransformedDistribution/MultivariateNormalDiag/Mixture/"
                            "MixtureSameFamily/Affine arithmetic (restated in the shim) and float32 rounding",
        },
        "single_flow": run_single_flows(FLOWS, ka),
        "known_layers": run_layer_known(DL, ka),
        "chains": run_chains(FLOWS, DL, cv),
        "mixtures": run_mixtures(DL, mv),
    }


def compute_f32():
    """The same reference code executed in float32 (the reference's working precision) on the chain and
    MDN inputs: how far the reference's OWN float32 evaluation sits from the float64 truth.  Context for the
    parity tolerances (tools/accuracy_report.py); float32 results may differ in the last bits across CPUs, so
    this file is not part of the rerun-equality check."""
    FLOWS, DL = tf_shim.load_reference(dtype=torch.float32)
    try:
        out = {"provenance": "reference flow / layer code on torch-CPU float32 stand-ins (oracle/tf_shim.py); "
                             "python -m oracle.make_reference_run --f32", "chains": [], "mdn": []}
        f32 = lambda a, grad=False: torch.tensor(np.asarray(a, dtype=np.float32), requires_grad=grad)
        for c in load("chain_vectors.json"):
            layer = DL.InverseNormalizingFlowLayer(c["flow_types"], c["n_dims"],
                                                   trainable_base_dist=c["trainable_base_dist"])
            t, y = f32(c["t"], True), f32(c["y"], True)
            logp = layer(t).log_prob(y)
            assert logp.dtype == torch.float32
            dt, dy = torch.autograd.grad(logp, [t, y], grad_outputs=f32(c["upstream"]))
            out["chains"].append({"name": c["name"], "sigma": c["sigma"], "log_prob": lst(logp),
                                  "dt": lst(dt), "dy": lst(dy)})
        for c in load("mixture_vectors.json")["mdn"]:
            layer = DL.GaussianMixtureLayer(c["n_centers"], c["n_dims"])
            t, y = f32(c["t"], True), f32(c["y"], True)
            logp = layer(t).log_prob(y)
            dt, dy = torch.autograd.grad(logp, [t, y], grad_outputs=f32(c["upstream"]))
            out["mdn"].append({"name": c["name"], "sigma": c["sigma"], "log_prob": lst(logp),
                               "dt": lst(dt), "dy": lst(dy)})
        return out
    finally:
        tf_shim.uninstall()
        tf_shim.install(torch.float64)  # leave the module-level dtype as it was
        tf_shim.uninstall()


def diff(a, b, path=""):
    """Largest absolute difference between two nested fixtures (structure must match)."""
    if isinstance(a, dict):
        assert a.keys() == b.keys(), path
        return max([diff(a[k], b[k], f"{path}/{k}") for k in a] or [0.0])
    if isinstance(a, list) and a and not isinstance(a[0], (int, float)):
        assert len(a) == len(b), path
        return max(diff(x, y, f"{path}[{i}]") for i, (x, y) in enumerate(zip(a, b)))
    if isinstance(a, (list, int, float)) and not isinstance(a, bool):
        return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)))) if np.size(a) else 0.0
    assert a == b, (path, a, b)
    return 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true", help="recompute and compare with the committed fixture")
    ap.add_argument("--f32", action="store_true", help="write reference_run_f32.json (reference code in float32)")
    args = ap.parse_args()
    if args.f32:
        out = os.path.join(GOLDEN, "reference_run_f32.json")
        with open(out, "w") as f:
            json.dump(compute_f32(), f)
        print("wrote", out, os.path.getsize(out), "bytes")
        return
    got = compute()
    if args.check:
        with open(OUT) as f:
            want = json.load(f)
        m = diff(want, got)
        print(f"max |fixture - rerun| = {m:.3e}")
        sys.exit(0 if m <= 1e-13 else 1)
    with open(OUT, "w") as f:
        json.dump(got, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;",
          len(got["single_flow"]), "single flows,", len(got["known_layers"]), "known layers,",
          len(got["chains"]), "chains,", len(got["mixtures"]["mdn"]), "mdn,", len(got["mixtures"]["kmn"]), "kmn")


if __name__ == "__main__":
    main()
