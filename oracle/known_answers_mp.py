"""50-digit mpmath scalar evaluation of the reference's own test inputs.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED (no TF/TFP here).

Third, fully independent oracle: pure-Python scalar loops in 50-digit arithmetic for the
inputs the reference's tests use (``tf.ones`` parameter rows, z in {0, 1}:
/root/reference/tests/test_flows.py:19-38, tests/test_distribution_layers.py:154-163).
Formulas follow PlanarFlow.py:43-80, RadialFlow.py:44-84, AffineFlow.py:4-10 and
DistributionLayers.py:245-294 literally.
"""
import mpmath as mp

mp.mp.dps = 50
C0 = mp.log(mp.expm1(1))


def sp(x):
    return mp.log1p(mp.exp(x))


def psize(name, d):
    return {"planar": 2 * d + 1, "radial": d + 2, "affine": 2 * d}[name]


def flow_step(name, th, z):
    """One flow on one sample: returns (z', fldj). ``th`` and ``z`` are lists of mpf."""
    d = len(z)
    if name == "planar":
        u = th[0:d]
        w = [x + 1 for x in th[d : 2 * d]]
        b = th[2 * d]
        wtu = mp.fsum(wi * ui for wi, ui in zip(w, u))
        m = -1 + sp(wtu) + mp.mpf("1e-5")
        n = mp.fsum(wi * wi for wi in w) + mp.mpf("1e-9")
        uh = [ui + (m - wtu) * (wi / n) for ui, wi in zip(u, w)]
        a = mp.fsum(wi * zi for wi, zi in zip(w, z)) + b
        tau = mp.tanh(a)
        psi = [(1 - tau ** 2) * wi for wi in w]
        det = 1 + mp.fsum(ui * pi for ui, pi in zip(uh, psi))
        return [zi + ui * tau for zi, ui in zip(z, uh)], mp.log(abs(det))
    if name == "radial":
        alpha = sp(mp.mpf("0.3") * th[0] - 2)
        beta = sp(mp.mpf("0.1") * th[1] + C0) - 1
        gam = th[2 : d + 2]
        r = mp.fsum(abs(zi - gi) for zi, gi in zip(z, gam))
        h = 1 / (alpha + r)
        der_h = -1 / (alpha + r) ** 2
        ab = alpha * beta
        det = (1 + ab * h) ** (d - 1) * (1 + ab * h + ab * der_h * r)
        return [zi + ab * h * (zi - gi) for zi, gi in zip(z, gam)], mp.log(det)
    if name == "affine":
        s = [1 + x for x in th[d : 2 * d]]
        return (
            [si * zi + sh for si, zi, sh in zip(s, z, th[0:d])],
            mp.fsum(mp.log(abs(si)) for si in s),
        )
    raise KeyError(name)


def chain_log_prob(trow, y, flow_types, d, trainable_base):
    """One sample through the whole layer (DistributionLayers.py:245-294)."""
    trow = [mp.mpf(x) for x in trow]
    z = [mp.mpf(x) for x in y]
    off = 2 * d if trainable_base else 0
    offs = {}
    for k in range(len(flow_types) - 1, -1, -1):  # last flow owns the first columns
        offs[k] = off
        off += psize(flow_types[k], d)
    assert off == len(trow)
    ld = mp.mpf(0)
    for k, name in enumerate(flow_types):
        z, f = flow_step(name, trow[offs[k] : offs[k] + psize(name, d)], z)
        ld += f
    if trainable_base:
        mu = trow[0:d]
        sig = [mp.mpf("1e-3") + sp(C0 + mp.mpf("0.1") * x) for x in trow[d : 2 * d]]
    else:
        mu = [mp.mpf(0)] * d
        sig = [mp.mpf(1)] * d
    e = [(zi - mi) / si for zi, mi, si in zip(z, mu, sig)]
    return (
        -mp.fsum(ei * ei for ei in e) / 2
        - mp.fsum(mp.log(si) for si in sig)
        - mp.mpf(d) / 2 * mp.log(2 * mp.pi)
        + ld
    )


def mdn_log_prob(trow, y, K, d):
    """One sample of the MDN head (DistributionLayers.py:196-212)."""
    trow = [mp.mpf(x) for x in trow]
    y = [mp.mpf(x) for x in y]
    logits = trow[2 * K * d : 2 * K * d + K]
    lse = mp.log(mp.fsum(mp.exp(l) for l in logits))
    terms = []
    for k in range(K):
        s0 = 2 * k * d
        mu = trow[s0 : s0 + d]
        sig = [sp(mp.mpf("0.05") * x + C0) for x in trow[s0 + d : s0 + 2 * d]]
        lp = (
            -mp.fsum(((yi - mi) / si) ** 2 for yi, mi, si in zip(y, mu, sig)) / 2
            - mp.fsum(mp.log(si) for si in sig)
            - mp.mpf(d) / 2 * mp.log(2 * mp.pi)
        )
        terms.append(logits[k] - lse + lp)
    return mp.log(mp.fsum(mp.exp(x) for x in terms))
