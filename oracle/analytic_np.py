"""Independent NumPy float64 closed-form forward + reverse sweep of the hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED (no TF/TFP here).

This is the *second* oracle: it does not share code or op order with
``flow_oracle.py`` (which is literal + autograd).  It evaluates the closed forms of
SURVEY.md App. A.2-A.4 (derived from PlanarFlow.py:20-80, RadialFlow.py:20-84,
AffineFlow.py:4-10, DistributionLayers.py:196-212/:245-294) with hand-derived
derivatives, so agreement of the two to ~1e-13 checks both the restatement and the
analytic backward that the CUDA kernels implement.
"""
import numpy as np

C0 = np.log(np.expm1(1.0))
HALF_LOG_2PI = 0.5 * np.log(2.0 * np.pi)
SIZES = {"planar": lambda d: 2 * d + 1, "radial": lambda d: d + 2, "affine": lambda d: 2 * d}


def softplus(x):
    return np.maximum(x, 0.0) + np.log1p(np.exp(-np.abs(x)))


def sigmoid(x):
    e = np.exp(-np.abs(x))
    return np.where(x >= 0, 1.0 / (1.0 + e), e / (1.0 + e))


def layout(flow_types, d, trainable_base):
    """Column offset of each flow's parameters, indexed in ``flow_types`` order.

    DistributionLayers.py:267-278 slices over the reversed list: the LAST flow owns the
    first columns after the optional 2d base-distribution block (:252, :283-288).
    """
    off = 2 * d if trainable_base else 0
    offs = [0] * len(flow_types)
    for k in range(len(flow_types) - 1, -1, -1):
        offs[k] = off
        off += SIZES[flow_types[k]](d)
    return offs, off


def chain_forward_backward(t, y, flow_types, d, trainable_base, upstream=None, need_grad=True):
    """Returns logp[B] (and dt[B,P], dy[B,d] scaled by ``upstream`` [B] or scalar)."""
    t = np.asarray(t, dtype=np.float64)
    B = t.shape[0]
    y = np.broadcast_to(np.asarray(y, dtype=np.float64), (B, d))
    offs, P = layout(flow_types, d, trainable_base)
    assert t.shape[1] == P, (t.shape, P)
    K = len(flow_types)
    zs = [y]
    ld = np.zeros(B)
    saved = []
    z = y
    for k in range(K):
        th = t[:, offs[k] : offs[k] + SIZES[flow_types[k]](d)]
        if flow_types[k] == "planar":
            u, w, b = th[:, :d], th[:, d : 2 * d] + 1.0, th[:, 2 * d]
            wtu = np.sum(w * u, 1)
            m = -1.0 + softplus(wtu) + 1e-5
            n = np.sum(w * w, 1) + 1e-9
            c = (m - wtu) / n
            uh = u + c[:, None] * w
            a = np.sum(w * z, 1) + b
            tau = np.tanh(a)
            s = 1.0 - tau * tau
            wuh = np.sum(w * uh, 1)
            D = 1.0 + s * wuh
            ld = ld + np.log(np.abs(D))
            saved.append((u, w, wtu, n, c, uh, tau, s, wuh, D))
            z = z + uh * tau[:, None]
        elif flow_types[k] == "radial":
            ar, br, gam = th[:, 0], th[:, 1], th[:, 2 : d + 2]
            alpha = softplus(0.3 * ar - 2.0)
            beta = softplus(0.1 * br + C0) - 1.0
            delta = z - gam
            r = np.sum(np.abs(delta), 1)
            h = 1.0 / (alpha + r)
            ab = alpha * beta
            T1 = 1.0 + ab * h
            T2 = 1.0 + ab * h - ab * h * h * r
            ld = ld + (d - 1) * np.log(T1) + np.log(T2)
            saved.append((ar, br, alpha, beta, delta, r, h, ab, T1, T2))
            z = z + (ab * h)[:, None] * delta
        else:
            shift, s = th[:, :d], 1.0 + th[:, d : 2 * d]
            ld = ld + np.sum(np.log(np.abs(s)), 1)
            saved.append((s,))
            z = s * z + shift
        zs.append(z)
    if trainable_base:
        mu = t[:, :d]
        sraw = t[:, d : 2 * d]
        sig = 1e-3 + softplus(C0 + 0.1 * sraw)
    else:
        mu = np.zeros((B, d))
        sig = np.ones((B, d))
    e = (z - mu) / sig
    logp = -0.5 * np.sum(e * e, 1) - np.sum(np.log(sig), 1) - d * HALF_LOG_2PI + ld
    if not need_grad:
        return logp
    dt = np.zeros_like(t)
    G = -e / sig
    if trainable_base:
        dt[:, :d] = e / sig
        dt[:, d : 2 * d] = (e * e - 1.0) / sig * 0.1 * sigmoid(C0 + 0.1 * sraw)
    for k in range(K - 1, -1, -1):
        zin = zs[k]
        o = offs[k]
        if flow_types[k] == "planar":
            u, w, wtu, n, c, uh, tau, s, wuh, D = saved[k]
            q = np.sum(uh * G, 1)
            g_a = s * q - 2.0 * tau * s * wuh / D
            g_uh = tau[:, None] * G + (s / D)[:, None] * w
            g_w = zin * g_a[:, None] + (s / D)[:, None] * uh
            p = np.sum(g_uh * w, 1)
            g_wtu = p * (sigmoid(wtu) - 1.0) / n
            g_n = -p * c / n
            dt[:, o : o + d] = g_uh + g_wtu[:, None] * w
            dt[:, o + d : o + 2 * d] = (
                g_w + c[:, None] * g_uh + g_wtu[:, None] * u + 2.0 * g_n[:, None] * w
            )
            dt[:, o + 2 * d] = g_a
            G = G + w * g_a[:, None]
        elif flow_types[k] == "radial":
            ar, br, alpha, beta, delta, r, h, ab, T1, T2 = saved[k]
            Dl = np.sum(delta * G, 1)
            g_h = ab * Dl + (d - 1) * ab / T1 + (ab - 2.0 * ab * h * r) / T2
            g_r = -g_h * h * h - ab * h * h / T2
            g_ab = h * Dl + (d - 1) * h / T1 + (h - h * h * r) / T2
            v = (ab * h)[:, None] * G + np.sign(delta) * g_r[:, None]
            dt[:, o] = (-g_h * h * h + beta * g_ab) * 0.3 * sigmoid(0.3 * ar - 2.0)
            dt[:, o + 1] = alpha * g_ab * 0.1 * sigmoid(0.1 * br + C0)
            dt[:, o + 2 : o + 2 + d] = -v
            G = G + v
        else:
            (s,) = saved[k]
            dt[:, o : o + d] = G
            dt[:, o + d : o + 2 * d] = zin * G + 1.0 / s
            G = s * G
    if upstream is not None:
        up = np.broadcast_to(np.asarray(upstream, dtype=np.float64), (B,))[:, None]
        dt = dt * up
        G = G * up
    return logp, dt, G


def mdn_forward_backward(t, y, n_centers, d, upstream=None, need_grad=True):
    """MDN head: DistributionLayers.py:196-212 + Mixture.log_prob; SURVEY.md App. A.4."""
    t = np.asarray(t, dtype=np.float64)
    B = t.shape[0]
    K = n_centers
    y = np.broadcast_to(np.asarray(y, dtype=np.float64), (B, d))
    assert t.shape[1] == 2 * K * d + K
    g = t[:, : 2 * K * d].reshape(B, K, 2, d)
    mu, sraw = g[:, :, 0, :], g[:, :, 1, :]
    sig = softplus(0.05 * sraw + C0)
    logits = t[:, 2 * K * d :]
    mx = logits.max(1, keepdims=True)
    lsm = logits - mx - np.log(np.sum(np.exp(logits - mx), 1, keepdims=True))
    e = (y[:, None, :] - mu) / sig
    lp = lsm - 0.5 * np.sum(e * e, 2) - np.sum(np.log(sig), 2) - d * HALF_LOG_2PI
    m2 = lp.max(1, keepdims=True)
    logp = (m2 + np.log(np.sum(np.exp(lp - m2), 1, keepdims=True)))[:, 0]
    if not need_grad:
        return logp
    rho = np.exp(lp - logp[:, None])
    dt = np.zeros_like(t)
    dg = dt[:, : 2 * K * d].reshape(B, K, 2, d)
    dg[:, :, 0, :] = rho[:, :, None] * e / sig
    dg[:, :, 1, :] = rho[:, :, None] * (e * e - 1.0) / sig * 0.05 * sigmoid(0.05 * sraw + C0)
    dt[:, 2 * K * d :] = rho - np.exp(lsm)
    dy = -np.sum(rho[:, :, None] * e / sig, 1)
    if upstream is not None:
        up = np.broadcast_to(np.asarray(upstream, dtype=np.float64), (B,))[:, None]
        dt = dt * up
        dy = dy * up
    return logp, dt, dy


def kmn_forward_backward(t, y, locs, scales, upstream=None, need_grad=True):
    """KMN head: DistributionLayers.py:118-133 + MixtureSameFamily.log_prob.

    Returns logp, dt (= d logp / d logits), dscales[M] (sum over the batch of
    upstream * d logp / d scale_m), dy.
    """
    t = np.asarray(t, dtype=np.float64)
    locs = np.asarray(locs, dtype=np.float64)
    scales = np.asarray(scales, dtype=np.float64)
    B, M = t.shape
    d = locs.shape[1]
    y = np.broadcast_to(np.asarray(y, dtype=np.float64), (B, d))
    mx = t.max(1, keepdims=True)
    lsm = t - mx - np.log(np.sum(np.exp(t - mx), 1, keepdims=True))
    e = (y[:, None, :] - locs[None]) / scales[None, :, None]
    q = np.sum(e * e, 2)
    lp = lsm - 0.5 * q - d * np.log(np.abs(scales))[None] - d * HALF_LOG_2PI
    m2 = lp.max(1, keepdims=True)
    logp = (m2 + np.log(np.sum(np.exp(lp - m2), 1, keepdims=True)))[:, 0]
    if not need_grad:
        return logp
    rho = np.exp(lp - logp[:, None])
    up = np.ones(B) if upstream is None else np.broadcast_to(np.asarray(upstream, np.float64), (B,))
    dt = (rho - np.exp(lsm)) * up[:, None]
    dscales = np.sum(up[:, None] * rho * (q - d) / scales[None], 0)
    dy = -np.sum(rho[:, :, None] * e / scales[None, :, None], 1) * up[:, None]
    return logp, dt, dscales, dy
