"""Tensor-level entry points over the C ABI (torch.Tensor in, torch.Tensor out).

Every function here enqueues hand-written sm_100a kernels from libnfn_b200.so on the
current CUDA stream; torch only owns the memory.  CPU tensors are rejected (no fallback).

Autograd: ``chain_log_prob`` / ``mdn_log_prob`` / ``kmn_log_prob`` are differentiable in
the parameter tensor ``t`` (and ``y``): the forward launches the forward-only kernel and
the backward launches the fused forward+reverse-sweep kernel with the incoming cotangent,
so no per-flow activation is ever stored.  Training loops that own the loss should call
``chain_forward_backward`` directly (one launch, cotangent -1/B folded in) and feed ``dt``
to ``t.backward(dt)``.
"""
import ctypes

import torch

from . import _lib


def _as_f32_cuda(x, name, device=None):
    if not torch.is_tensor(x):
        if device is None:
            raise RuntimeError("%s: cannot infer the CUDA device from a non-tensor input" % name)
        x = torch.as_tensor(x, dtype=torch.float32).to(device)
    _lib.require_cuda(x, name)
    if x.dtype != torch.float32:
        x = x.to(torch.float32)
    return x.contiguous()


def _aligned(x, nbytes=16):
    """Kernels need 16-byte aligned bases; torch allocations are, sliced views may not be."""
    return x if x.data_ptr() % nbytes == 0 else x.clone(memory_format=torch.contiguous_format)


def _prep_ty(t, y, n_dims, width, what):
    t = _as_f32_cuda(t, "t")
    if t.dim() != 2:
        raise ValueError("%s: parameter tensor must be rank 2 [B, P], got shape %s" % (what, tuple(t.shape)))
    y = _as_f32_cuda(y, "y", device=t.device)
    if y.dim() != 2 or y.shape[1] != n_dims:
        raise ValueError("%s: y must be [B, %d] or [1, %d], got %s" % (what, n_dims, n_dims, tuple(y.shape)))
    B = t.shape[0]
    if t.shape[1] != width:
        raise AssertionError("%s: expected %d parameter columns, got %d" % (what, width, t.shape[1]))
    if y.shape[0] not in (B, 1):
        if B == 1:  # [1, P] parameters against many y rows: materialise the parameter rows
            t = t.expand(y.shape[0], width).contiguous()
            B = y.shape[0]
        else:
            raise ValueError("%s: y has %d rows, parameters have %d" % (what, y.shape[0], B))
    return _aligned(t), _aligned(y), B


def _xf(xform):
    """NULL or a pointer to the caller's ``_lib.EventXform`` (the y pipeline fused into the head, see make_xform)."""
    return None if xform is None else ctypes.byref(xform)


def make_xform(n_dims, mean=None, std=None, logp_shift=0.0, noise_std=0.0, seed=0, offset=0, offset_dev=None,
               exp_out=False):
    """Build the event transform of a head launch (include/nfn_b200.h: nfn_event_xform).

    mean / std: host sequences of length n_dims -> y' = (y - mean) / std;  noise_std > 0 adds N(0, noise_std^2)
    noise drawn in the kernel (Philox keyed by ``seed``, counter (row, offset + *offset_dev));  logp_shift is
    added to every log-prob;  exp_out writes exp(logp + shift), i.e. the density."""
    xf = _lib.EventXform()
    flags = 0
    if mean is not None:
        assert len(mean) == n_dims and len(std) == n_dims
        for i in range(n_dims):
            xf.mean[i] = float(mean[i])
            xf.std[i] = float(std[i])
        flags |= _lib.XF_NORMALISE
    if noise_std and noise_std > 0.0:
        xf.noise_std = float(noise_std)
        xf.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        xf.offset = int(offset) & 0xFFFFFFFFFFFFFFFF
        xf.offset_dev = None if offset_dev is None else offset_dev.data_ptr()
        flags |= _lib.XF_NOISE
    if exp_out:
        flags |= _lib.XF_EXP
    xf.logp_shift = float(logp_shift)
    xf.flags = flags
    return xf


def _check_y(y, n_dims, B, what):
    """The C ABI cannot see tensor extents: a wrong-shaped y would be an out-of-bounds device read."""
    if y.dim() != 2 or y.shape[1] != n_dims or y.shape[0] not in (B, 1):
        raise ValueError("%s: y must be [%d, %d] or [1, %d], got %s" % (what, B, n_dims, n_dims, tuple(y.shape)))


def _prep_g(g_logp, B, dev, what):
    if g_logp is None:
        return None
    g_logp = _as_f32_cuda(g_logp, "g_logp", device=dev).reshape(-1)
    if g_logp.numel() != B:
        raise ValueError("%s: g_logp must have B=%d elements, got %d" % (what, B, g_logp.numel()))
    return g_logp


# ----------------------------------------------------------------------------- flow chain
def chain_param_size(flow_types, n_dims, trainable_base_dist):
    lib = _lib.load()
    return _lib.check(lib.nfn_chain_param_size(ctypes.byref(_lib.make_desc(flow_types, n_dims, trainable_base_dist))))


def chain_is_specialized(flow_types, n_dims, trainable_base_dist):
    lib = _lib.load()
    return bool(_lib.check(lib.nfn_chain_is_specialized(ctypes.byref(_lib.make_desc(flow_types, n_dims, trainable_base_dist)))))


def chain_forward(t, y, flow_types, n_dims, trainable_base_dist, xform=None):
    """log_prob[B] of the inverted flow chain (no autograd)."""
    lib = _lib.load()
    desc = _lib.make_desc(flow_types, n_dims, trainable_base_dist)
    P = _lib.check(lib.nfn_chain_param_size(ctypes.byref(desc)))
    t, y, B = _prep_ty(t, y, n_dims, P, "chain_forward")
    logp = torch.empty(B, dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        _lib.check(lib.nfn_chain_forward_x(ctypes.byref(desc), _lib.ptr(t), _lib.ptr(y), y.shape[0],
                                           _lib.ptr(logp), B, _xf(xform), _lib.current_stream(t.device)))
    return logp


def chain_forward_grid(t, y_grid, flow_types, n_dims, trainable_base_dist, xform=None):
    """log_prob of every parameter row against every event: returns [n_y, B] (event-major).
    The parameter tensor is read once, not once per event."""
    lib = _lib.load()
    desc = _lib.make_desc(flow_types, n_dims, trainable_base_dist)
    P = _lib.check(lib.nfn_chain_param_size(ctypes.byref(desc)))
    t = _aligned(_as_f32_cuda(t, "t"))
    if t.dim() != 2 or t.shape[1] != P:
        raise AssertionError("chain_forward_grid: expected [B, %d] parameters, got %s" % (P, tuple(t.shape)))
    y_grid = _aligned(_as_f32_cuda(y_grid, "y_grid", device=t.device))
    if y_grid.dim() != 2 or y_grid.shape[1] != n_dims:
        raise ValueError("chain_forward_grid: y_grid must be [n_y, %d]" % n_dims)
    B, ny = t.shape[0], y_grid.shape[0]
    logp = torch.empty((ny, B), dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        _lib.check(lib.nfn_chain_forward_grid_x(ctypes.byref(desc), _lib.ptr(t), _lib.ptr(y_grid), ny, _lib.ptr(logp),
                                                B, _xf(xform), _lib.current_stream(t.device)))
    return logp


def chain_forward_backward(t, y, flow_types, n_dims, trainable_base_dist, g_logp=None, g_scale=1.0,
                           want_dy=False, logp_sum=None, dt_colsum=None, out_logp=None, out_dt=None, xform=None):
    """Fused forward + reverse sweep.  Returns (logp[B], dt[B,P], dy[B,d] or None).

    dt = cot[:, None] * dlogp/dt with cot = g_scale * (g_logp if given else 1).
    ``logp_sum`` (float64 [1]) and ``dt_colsum`` (float64 [P]) are accumulated into when given.
    """
    lib = _lib.load()
    desc = _lib.make_desc(flow_types, n_dims, trainable_base_dist)
    P = _lib.check(lib.nfn_chain_param_size(ctypes.byref(desc)))
    t, y, B = _prep_ty(t, y, n_dims, P, "chain_forward_backward")
    dev = t.device
    logp = out_logp if out_logp is not None else torch.empty(B, dtype=torch.float32, device=dev)
    dt = out_dt if out_dt is not None else torch.empty((B, P), dtype=torch.float32, device=dev)
    dy = torch.empty((B, n_dims), dtype=torch.float32, device=dev) if want_dy else None
    if want_dy and y.shape[0] != B:
        raise ValueError("dy needs one y row per parameter row")
    if g_logp is not None:
        g_logp = _as_f32_cuda(g_logp, "g_logp", device=dev).reshape(-1)
        if g_logp.numel() != B:
            raise ValueError("g_logp must have B=%d elements" % B)
    with torch.cuda.device(dev):
        _lib.check(lib.nfn_chain_forward_backward_x(
            ctypes.byref(desc), _lib.ptr(t), _lib.ptr(y), y.shape[0], _lib.ptr(g_logp),
            ctypes.c_float(g_scale), _lib.ptr(logp), _lib.ptr(dt), _lib.ptr(dy), _lib.ptr(logp_sum),
            _lib.ptr(dt_colsum), B, _xf(xform), _lib.current_stream(dev)))
    return logp, dt, dy


def chain_forward_backward_peer(t, y, flow_types, n_dims, trainable_base_dist, comm, g_logp=None, g_scale=1.0,
                                want_colsum=False, reduced=None, out_logp=None, out_dt=None):
    """Fused forward + reverse sweep + in-kernel all-reduce over NVLink peer memory.

    ``comm`` is a ``parallel.PeerComm`` with ``n_values == P + 1``.  Returns
    (logp[B], dt[B,P], reduced[P+1] float64 = [sum over ALL ranks of dt column sums | of logp]).
    """
    lib = _lib.load()
    desc = _lib.make_desc(flow_types, n_dims, trainable_base_dist)
    P = _lib.check(lib.nfn_chain_param_size(ctypes.byref(desc)))
    t, y, B = _prep_ty(t, y, n_dims, P, "chain_forward_backward_peer")
    dev = t.device
    logp = out_logp if out_logp is not None else torch.empty(B, dtype=torch.float32, device=dev)
    dt = out_dt if out_dt is not None else torch.empty((B, P), dtype=torch.float32, device=dev)
    if reduced is None:
        reduced = torch.empty(P + 1, dtype=torch.float64, device=dev)
    g_logp = _prep_g(g_logp, B, dev, "chain_forward_backward_peer")
    with torch.cuda.device(dev):
        _lib.check(lib.nfn_chain_forward_backward_peer(
            ctypes.byref(desc), _lib.ptr(t), _lib.ptr(y), y.shape[0], _lib.ptr(g_logp), ctypes.c_float(g_scale),
            _lib.ptr(logp), _lib.ptr(dt), None, 1 if want_colsum else 0, comm.comm, _lib.ptr(reduced), B,
            _lib.current_stream(dev)))
    return logp, dt, reduced


# ----------------------------------------------------------------------------- hidden layers of the MLP
ACT_CODES = {"linear": 0, "tanh": 1, "relu": 2, "sigmoid": 3, "elu": 4}


def dense_act_supported(in_features, units, activation):
    """True when one fused kernel each way can serve ``Dense(units, activation)`` on ``in_features`` inputs."""
    if activation not in ACT_CODES:
        return False
    return bool(_lib.load().nfn_dense_act_supported(int(in_features), int(units), ACT_CODES[activation]))


def dense_act_forward(x, weight, bias, activation, x_mean=None, x_std=None):
    """act(x' @ weight.T + bias) in one kernel; weight is torch's [units, in_features].  With x_mean / x_std
    (device [in_features]) the layer reads x' = (x - x_mean) / (x_std + 1e-8): the estimators' input normalisation
    fused into their first layer."""
    lib = _lib.load()
    x = _aligned(_as_f32_cuda(x, "x"))
    weight = _as_f32_cuda(weight, "weight", device=x.device)
    bias = _as_f32_cuda(bias, "bias", device=x.device)
    B, K = x.shape
    N = weight.shape[0]
    assert tuple(weight.shape) == (N, K) and tuple(bias.shape) == (N,)
    if x_mean is not None:
        x_mean, x_std = _as_f32_cuda(x_mean, "x_mean", device=x.device), _as_f32_cuda(x_std, "x_std", device=x.device)
        assert x_mean.numel() == K and x_std.numel() == K
    out = torch.empty((B, N), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.nfn_dense_act_forward_x(_lib.ptr(x), _lib.ptr(x_mean), _lib.ptr(x_std), _lib.ptr(weight),
                                               _lib.ptr(bias), B, K, N, ACT_CODES[activation], _lib.ptr(out),
                                               _lib.current_stream(x.device)))
    return out


def dense_act_backward(x, out, dout, weight, activation, need_dx=True, x_mean=None, x_std=None):
    """Gradients of dense_act_forward given the layer's OUTPUT: returns (dx or None, dweight, dbias)."""
    lib = _lib.load()
    x = _aligned(_as_f32_cuda(x, "x"))
    dev = x.device
    out = _aligned(_as_f32_cuda(out, "out", device=dev))
    dout = _aligned(_as_f32_cuda(dout, "dout", device=dev))
    weight = _as_f32_cuda(weight, "weight", device=dev)
    B, K = x.shape
    N = weight.shape[0]
    dx = torch.empty((B, K), dtype=torch.float32, device=dev) if need_dx else None
    dW = torch.zeros((N, K), dtype=torch.float32, device=dev)
    db = torch.zeros(N, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.nfn_dense_act_backward_x(_lib.ptr(x), _lib.ptr(x_mean), _lib.ptr(x_std), _lib.ptr(out),
                                                _lib.ptr(dout), _lib.ptr(weight), B, K, N, ACT_CODES[activation],
                                                _lib.ptr(dx), _lib.ptr(dW), _lib.ptr(db), _lib.current_stream(dev)))
    return dx, dW, db


def dense_act_xnorm_supported(in_features, units, activation):
    """True when the FIRST layer of a network can take the input normalisation as a fused prologue (both ways)."""
    return dense_act_supported(in_features, units, activation) and in_features <= 4 and units <= 32


class _DenseAct(torch.autograd.Function):
    """Dense(units, activation) as one kernel each way (csrc/nfn_mlp.cu)."""

    @staticmethod
    def forward(ctx, x, weight, bias, activation, x_mean, x_std):
        out = dense_act_forward(x, weight, bias, activation, x_mean, x_std)
        ctx.save_for_backward(x, out, weight, x_mean, x_std)
        ctx.activation = activation
        return out

    @staticmethod
    def backward(ctx, dout):
        x, out, weight, x_mean, x_std = ctx.saved_tensors
        dx, dW, db = dense_act_backward(x, out, dout.contiguous(), weight, ctx.activation,
                                        need_dx=ctx.needs_input_grad[0], x_mean=x_mean, x_std=x_std)
        return dx, dW, db, None, None, None


def dense_act(x, weight, bias, activation, x_mean=None, x_std=None):
    """Differentiable fused hidden layer (forward kernel under no_grad, Function otherwise)."""
    return _DenseAct.apply(x, weight, bias, activation, x_mean, x_std)


# ----------------------------------------------------------------------------- fused Dense(P) + chain
def dense_chain_supported(hidden):
    return hidden in (16, 32, 48, 64)


def dense_chain_forward(h, W, bias, y, flow_types, n_dims, trainable_base_dist, xform=None):
    """log_prob[B] with the emitting layer fused: t = h @ W + bias never touches HBM.
    h [B, H], W [H, P] (Keras kernel layout = torch ``linear.weight.t()``), bias [P]."""
    lib = _lib.load()
    desc = _lib.make_desc(flow_types, n_dims, trainable_base_dist)
    P = _lib.check(lib.nfn_chain_param_size(ctypes.byref(desc)))
    h = _aligned(_as_f32_cuda(h, "h"))
    dev = h.device
    W = _as_f32_cuda(W, "W", device=dev)
    bias = _as_f32_cuda(bias, "bias", device=dev)
    y = _aligned(_as_f32_cuda(y, "y", device=dev))
    B, H = h.shape
    assert tuple(W.shape) == (H, P) and tuple(bias.shape) == (P,), "W must be [H, P] and bias [P]"
    _check_y(y, n_dims, B, "dense_chain_forward")
    logp = torch.empty(B, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.nfn_dense_chain_forward_x(ctypes.byref(desc), H, _lib.ptr(h), _lib.ptr(W), _lib.ptr(bias),
                                                 _lib.ptr(y), y.shape[0], _lib.ptr(logp), B, _xf(xform),
                                                 _lib.current_stream(dev)))
    return logp


def dense_chain_forward_backward(h, W, bias, y, flow_types, n_dims, trainable_base_dist, g_logp=None, g_scale=1.0,
                                 logp_sum=None, dW=None, dbias=None, xform=None):
    """Fused layer + chain, forward and reverse sweep.  Returns (logp[B], dh[B,H], dW[H,P], dbias[P]);
    dW / dbias are accumulated into when given (else fresh zero tensors)."""
    lib = _lib.load()
    desc = _lib.make_desc(flow_types, n_dims, trainable_base_dist)
    P = _lib.check(lib.nfn_chain_param_size(ctypes.byref(desc)))
    h = _aligned(_as_f32_cuda(h, "h"))
    dev = h.device
    W = _as_f32_cuda(W, "W", device=dev)
    bias = _as_f32_cuda(bias, "bias", device=dev)
    y = _aligned(_as_f32_cuda(y, "y", device=dev))
    B, H = h.shape
    assert tuple(W.shape) == (H, P) and tuple(bias.shape) == (P,), "W must be [H, P] and bias [P]"
    _check_y(y, n_dims, B, "dense_chain_forward_backward")
    logp = torch.empty(B, dtype=torch.float32, device=dev)
    dh = torch.empty((B, H), dtype=torch.float32, device=dev)
    if dW is None:
        dW = torch.zeros((H, P), dtype=torch.float32, device=dev)
    if dbias is None:
        dbias = torch.zeros(P, dtype=torch.float32, device=dev)
    g_logp = _prep_g(g_logp, B, dev, "dense_chain_forward_backward")
    with torch.cuda.device(dev):
        _lib.check(lib.nfn_dense_chain_forward_backward_x(
            ctypes.byref(desc), H, _lib.ptr(h), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(y), y.shape[0],
            _lib.ptr(g_logp), ctypes.c_float(g_scale), _lib.ptr(logp), _lib.ptr(dh), _lib.ptr(dW), _lib.ptr(dbias),
            _lib.ptr(logp_sum), B, _xf(xform), _lib.current_stream(dev)))
    return logp, dh, dW, dbias


# ----------------------------------------------------------------------------- mean-field weight posterior
class _VariationalSample(torch.autograd.Function):
    """(w [S, n], KL) of one DenseVariational layer from its flat posterior parameters: one kernel each way
    (csrc/nfn_variational.cu) instead of ~40 torch launches; reference DistributionLayers.py:17-71."""

    @staticmethod
    def forward(ctx, params, prior_loc, eps, prior_scale):
        lib = _lib.load()
        S, n = eps.shape
        dev = params.device
        p, pl, e = params.detach().contiguous(), prior_loc.detach().contiguous(), eps.contiguous()
        w = torch.empty((S, n), dtype=torch.float32, device=dev)
        kl = torch.zeros(1, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.nfn_variational_sample(_lib.ptr(p), _lib.ptr(pl), ctypes.c_float(prior_scale), _lib.ptr(e), n, S,
                                                  _lib.ptr(w), _lib.ptr(kl), _lib.current_stream(dev)))
        ctx.save_for_backward(p, pl, e)
        ctx.prior_scale = prior_scale
        ctx.need_prior = prior_loc.requires_grad
        return w, kl.to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, dw, dkl):
        lib = _lib.load()
        p, pl, e = ctx.saved_tensors
        S, n = e.shape
        dev = p.device
        dparams = torch.zeros_like(p)
        dprior = torch.zeros_like(pl) if ctx.need_prior else None
        dwc = dw.contiguous() if dw is not None else None
        g = dkl.to(torch.float32).reshape(1).contiguous() if dkl is not None else None
        with torch.cuda.device(dev):
            _lib.check(lib.nfn_variational_sample_backward(
                _lib.ptr(p), _lib.ptr(pl), ctypes.c_float(ctx.prior_scale), _lib.ptr(e), _lib.ptr(dwc), _lib.ptr(g), n, S,
                _lib.ptr(dparams), _lib.ptr(dprior), _lib.current_stream(dev)))
        return dparams, dprior, None, None


def variational_sample(posterior_params, prior_loc, eps, prior_scale):
    """w [S, n] = loc + (1e-3 + softplus(c0 + 0.05 raw)) * eps and the exact KL(q || N(prior_loc, prior_scale)) as a
    float32 scalar, differentiable with respect to ``posterior_params`` [2 n] (and ``prior_loc`` [n] if it is trained)."""
    return _VariationalSample.apply(posterior_params, prior_loc, eps, float(prior_scale))


def bayes_train_step(x, y, first, emitting, units, activation, hidden_width, n_draws, g_scale, ws, logp_sum,
                     flow_types=None, n_dims=1, trainable_base_dist=True, mdn_centers=0, x_mean=None, x_std=None,
                     xform=None):
    """Network part of one S-draw Bayesian training step in ONE library call (nfn_bayes_train_step): samples + KL of
    both variational layers, first layer over the folded rows, emitting layer + head + backward GEMMs with per-draw
    weights, first layer's weight gradient, gradients of the posterior parameters.  ``first`` / ``emitting``: dicts
    with posterior [2n], prior_loc [n], eps [S, n], dposterior [2n] (+=), dprior_loc [n] or None, prior_scale,
    kl_grad; ``ws``: workspace dict (w1, w2, dw1, dw2, h, dh, logp, kl [2] float64) reused across steps."""
    lib = _lib.load()
    dev = x.device
    desc = _lib.make_desc(flow_types if flow_types is not None else [], n_dims, trainable_base_dist)
    B, K = x.shape
    layers = []
    for i, (l, w, dw) in enumerate(((first, ws["w1"], ws["dw1"]), (emitting, ws["w2"], ws["dw2"]))):
        v = _lib.VariationalLayer()
        v.posterior, v.prior_loc, v.eps = l["posterior"].data_ptr(), l["prior_loc"].data_ptr(), l["eps"].data_ptr()
        v.w, v.dw, v.dposterior = w.data_ptr(), dw.data_ptr(), l["dposterior"].data_ptr()
        v.dprior_loc = l["dprior_loc"].data_ptr() if l.get("dprior_loc") is not None else None
        v.kl = ws["kl"].data_ptr() + 8 * i
        v.prior_scale, v.kl_grad, v.n = float(l["prior_scale"]), float(l["kl_grad"]), int(l["prior_loc"].numel())
        layers.append(v)
    with torch.cuda.device(dev):
        _lib.check(lib.nfn_bayes_train_step(
            ctypes.byref(desc), int(mdn_centers), int(n_draws), B, K, int(units), int(hidden_width), ACT_CODES[activation],
            _lib.ptr(x), _lib.ptr(x_mean), _lib.ptr(x_std), _lib.ptr(y), y.shape[0], ctypes.byref(layers[0]),
            ctypes.byref(layers[1]), ctypes.c_float(g_scale), _lib.ptr(ws["h"]), _lib.ptr(ws["dh"]), _lib.ptr(ws["logp"]),
            _lib.ptr(logp_sum), _xf(xform), _lib.current_stream(dev)))


# ----------------------------------------------------------------------------- folded posterior draws
def dense_act_draws_supported(in_features, units, out_width, activation):
    return (1 <= in_features <= 8 and 1 <= units <= 64 and out_width >= units and out_width % 8 == 0 and out_width <= 64
            and activation in ACT_CODES)


def dense_act_forward_draws(x, w, units, activation, out_width, x_mean=None, x_std=None):
    """First variational layer with S weight draws folded into the batch: x [B, in] (per sample), w [S, in*units + units]
    (tfp DenseVariational's flat [kernel | bias] sample per draw) -> out [S*B, out_width], draw-major, columns
    units .. out_width zero.  Reference: BayesianNNEstimator.py:65-76, :103-118."""
    lib = _lib.load()
    x = _as_f32_cuda(x, "x")
    dev = x.device
    w = _as_f32_cuda(w, "w", device=dev)
    B, K = x.shape
    S = w.shape[0]
    assert w.shape[1] == K * units + units, "w must be [S, in*units + units]"
    out = torch.empty((S * B, out_width), dtype=torch.float32, device=dev)
    xm = _as_f32_cuda(x_mean, "x_mean", device=dev) if x_mean is not None else None
    xs = _as_f32_cuda(x_std, "x_std", device=dev) if x_std is not None else None
    with torch.cuda.device(dev):
        _lib.check(lib.nfn_dense_act_forward_draws(_lib.ptr(x), _lib.ptr(xm), _lib.ptr(xs), _lib.ptr(w), S, B, K, units,
                                                   out_width, ACT_CODES[activation], _lib.ptr(out),
                                                   _lib.current_stream(dev)))
    return out


def dense_act_backward_draws(x, out, dout, n_draws, units, activation, x_mean=None, x_std=None):
    """Per-draw gradient of the folded first layer: dw [S, in*units + units] = [xn^T dpre_s | 1^T dpre_s]."""
    lib = _lib.load()
    x = _as_f32_cuda(x, "x")
    dev = x.device
    out = _aligned(_as_f32_cuda(out, "out", device=dev))
    dout = _aligned(_as_f32_cuda(dout, "dout", device=dev))
    B, K = x.shape
    assert out.shape == dout.shape and out.shape[0] == n_draws * B
    dw = torch.zeros((n_draws, K * units + units), dtype=torch.float32, device=dev)
    xm = _as_f32_cuda(x_mean, "x_mean", device=dev) if x_mean is not None else None
    xs = _as_f32_cuda(x_std, "x_std", device=dev) if x_std is not None else None
    with torch.cuda.device(dev):
        _lib.check(lib.nfn_dense_act_backward_draws(_lib.ptr(x), _lib.ptr(xm), _lib.ptr(xs), _lib.ptr(out), _lib.ptr(dout),
                                                    n_draws, B, K, units, out.shape[1], ACT_CODES[activation],
                                                    _lib.ptr(dw), _lib.current_stream(dev)))
    return dw


def dense_chain_forward_draws(h, W, bias, y, flow_types, n_dims, trainable_base_dist, xform=None):
    """Fused emitting layer + flow chain with per-draw weights: h [S*B, H] (draw-major), W [S, H, P], bias [S, P],
    y [B, d] per sample (or one row).  Returns logp [S*B]."""
    lib = _lib.load()
    desc = _lib.make_desc(flow_types, n_dims, trainable_base_dist)
    P = _lib.check(lib.nfn_chain_param_size(ctypes.byref(desc)))
    h = _aligned(_as_f32_cuda(h, "h"))
    dev = h.device
    W = _as_f32_cuda(W, "W", device=dev)
    bias = _as_f32_cuda(bias, "bias", device=dev)
    y = _aligned(_as_f32_cuda(y, "y", device=dev))
    S, H = W.shape[0], W.shape[1]
    assert tuple(W.shape) == (S, H, P) and tuple(bias.shape) == (S, P) and h.shape[1] == H and h.shape[0] % S == 0
    Bd = h.shape[0] // S
    if y.dim() != 2 or y.shape[1] != n_dims or y.shape[0] not in (Bd, 1):
        raise ValueError("y must be [rows_per_draw, n_dims] or one row")
    logp = torch.empty(S * Bd, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.nfn_dense_chain_forward_draws_x(ctypes.byref(desc), H, S, Bd, _lib.ptr(h), _lib.ptr(W), _lib.ptr(bias),
                                                       _lib.ptr(y), y.shape[0], _lib.ptr(logp), _xf(xform),
                                                       _lib.current_stream(dev)))
    return logp


def dense_chain_forward_backward_draws(h, W, bias, y, flow_types, n_dims, trainable_base_dist, g_logp=None, g_scale=1.0,
                                       logp_sum=None, xform=None):
    """Forward + reverse sweep of the above.  Returns (logp [S*B], dh [S*B, H], dW [S, H, P], dbias [S, P])."""
    lib = _lib.load()
    desc = _lib.make_desc(flow_types, n_dims, trainable_base_dist)
    P = _lib.check(lib.nfn_chain_param_size(ctypes.byref(desc)))
    h = _aligned(_as_f32_cuda(h, "h"))
    dev = h.device
    W = _as_f32_cuda(W, "W", device=dev)
    bias = _as_f32_cuda(bias, "bias", device=dev)
    y = _aligned(_as_f32_cuda(y, "y", device=dev))
    S, H = W.shape[0], W.shape[1]
    assert tuple(W.shape) == (S, H, P) and tuple(bias.shape) == (S, P) and h.shape[1] == H and h.shape[0] % S == 0
    Bd = h.shape[0] // S
    if y.dim() != 2 or y.shape[1] != n_dims or y.shape[0] not in (Bd, 1):
        raise ValueError("y must be [rows_per_draw, n_dims] or one row")
    logp = torch.empty(S * Bd, dtype=torch.float32, device=dev)
    dh = torch.empty((S * Bd, H), dtype=torch.float32, device=dev)
    dW = torch.zeros((S, H, P), dtype=torch.float32, device=dev)
    dbias = torch.zeros((S, P), dtype=torch.float32, device=dev)
    g_logp = _prep_g(g_logp, S * Bd, dev, "dense_chain_forward_backward_draws")
    with torch.cuda.device(dev):
        _lib.check(lib.nfn_dense_chain_forward_backward_draws_x(
            ctypes.byref(desc), H, S, Bd, _lib.ptr(h), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(y), y.shape[0],
            _lib.ptr(g_logp), ctypes.c_float(g_scale), _lib.ptr(logp), _lib.ptr(dh), _lib.ptr(dW), _lib.ptr(dbias),
            _lib.ptr(logp_sum), _xf(xform), _lib.current_stream(dev)))
    return logp, dh, dW, dbias


def dense_kmn_supported(hidden, n_components, n_dims):
    """Whether the fused Dense(P)+KMN kernel can serve this shape (see dense_mdn_supported; P = n_components plus the
    head's own shared-memory state: centres, two coefficients and four rows of bandwidth-gradient sums per kernel)."""
    if hidden not in (16, 32, 48, 64) or n_components < 1 or not 1 <= n_dims <= 8:
        return False
    P = n_components
    S = P + 4 if (P % 4 == 0 and (P // 4) % 2 == 0) else P
    P8 = (P + 7) // 8 * 8
    PW = P8
    while PW % 32 not in (8, 24):
        PW += 1
    floats = 128 * S + 2 * 128 * (hidden + 4) + hidden * PW + P8 + 4 * hidden * P8 + 4 * P8 + P * (n_dims + 2) + 4 * P
    return 4 * floats <= 220 * 1024


def dense_kmn_forward(h, W, bias, y, locs, scales, xform=None):
    """log_prob[B] of the kernel mixture with the emitting layer fused: the logits t = h @ W + bias never touch HBM.
    h [B, H], W [H, M], bias [M], locs [M, d], scales [M].  Reference: Dense(output_size) of
    MaximumLikelihoodNNEstimator.py:43 + DistributionLayers.py:118-133."""
    lib = _lib.load()
    locs = _as_f32_cuda(locs, "locs")
    M, d = locs.shape
    h = _aligned(_as_f32_cuda(h, "h", device=locs.device))
    dev = h.device
    scales = _as_f32_cuda(scales, "scales", device=dev).reshape(-1)
    W = _as_f32_cuda(W, "W", device=dev)
    bias = _as_f32_cuda(bias, "bias", device=dev)
    y = _aligned(_as_f32_cuda(y, "y", device=dev))
    B, H = h.shape
    assert tuple(W.shape) == (H, M) and tuple(bias.shape) == (M,) and scales.numel() == M, "W [H, M], bias [M], scales [M]"
    _check_y(y, d, B, "dense_kmn_forward")
    logp = torch.empty(B, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.nfn_dense_kmn_forward_x(M, d, H, _lib.ptr(h), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(y), y.shape[0],
                                               _lib.ptr(locs), _lib.ptr(scales), _lib.ptr(logp), B, _xf(xform),
                                               _lib.current_stream(dev)))
    return logp


def dense_kmn_forward_backward(h, W, bias, y, locs, scales, g_logp=None, g_scale=1.0, want_dscales=True, logp_sum=None,
                               dW=None, dbias=None, xform=None):
    """Fused layer + kernel-mixture head, forward and reverse sweep.  Returns (logp[B], dh[B,H], dW[H,M], dbias[M],
    dscales[M] or None); dW / dbias are accumulated into when given."""
    lib = _lib.load()
    locs = _as_f32_cuda(locs, "locs")
    M, d = locs.shape
    h = _aligned(_as_f32_cuda(h, "h", device=locs.device))
    dev = h.device
    scales = _as_f32_cuda(scales, "scales", device=dev).reshape(-1)
    W = _as_f32_cuda(W, "W", device=dev)
    bias = _as_f32_cuda(bias, "bias", device=dev)
    y = _aligned(_as_f32_cuda(y, "y", device=dev))
    B, H = h.shape
    assert tuple(W.shape) == (H, M) and tuple(bias.shape) == (M,) and scales.numel() == M, "W [H, M], bias [M], scales [M]"
    _check_y(y, d, B, "dense_kmn_forward_backward")
    logp = torch.empty(B, dtype=torch.float32, device=dev)
    dh = torch.empty((B, H), dtype=torch.float32, device=dev)
    if dW is None:
        dW = torch.zeros((H, M), dtype=torch.float32, device=dev)
    if dbias is None:
        dbias = torch.zeros(M, dtype=torch.float32, device=dev)
    dscales = torch.zeros(M, dtype=torch.float32, device=dev) if want_dscales else None
    g_logp = _prep_g(g_logp, B, dev, "dense_kmn_forward_backward")
    with torch.cuda.device(dev):
        _lib.check(lib.nfn_dense_kmn_forward_backward_x(
            M, d, H, _lib.ptr(h), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(y), y.shape[0], _lib.ptr(locs), _lib.ptr(scales),
            _lib.ptr(g_logp), ctypes.c_float(g_scale), _lib.ptr(logp), _lib.ptr(dh), _lib.ptr(dW), _lib.ptr(dbias),
            _lib.ptr(dscales), _lib.ptr(logp_sum), B, _xf(xform), _lib.current_stream(dev)))
    return logp, dh, dW, dbias, dscales


def _draws_operands(h, W, bias, y, P, n_dims, what):
    h = _aligned(_as_f32_cuda(h, "h"))
    dev = h.device
    W = _as_f32_cuda(W, "W", device=dev)
    bias = _as_f32_cuda(bias, "bias", device=dev)
    y = _aligned(_as_f32_cuda(y, "y", device=dev))
    S, H = W.shape[0], W.shape[1]
    assert tuple(W.shape) == (S, H, P) and tuple(bias.shape) == (S, P) and h.shape[1] == H and h.shape[0] % S == 0, what
    Bd = h.shape[0] // S
    if y.dim() != 2 or y.shape[1] != n_dims or y.shape[0] not in (Bd, 1):
        raise ValueError("%s: y must be [rows_per_draw, n_dims] or one row" % what)
    return h, W, bias, y, S, H, Bd, dev


def dense_mdn_forward_draws(h, W, bias, y, n_centers, n_dims, xform=None):
    """Fused emitting layer + MDN head with per-draw weights (see dense_chain_forward_draws).  Returns logp [S*B]."""
    lib = _lib.load()
    P = mdn_param_size(n_centers, n_dims)
    h, W, bias, y, S, H, Bd, dev = _draws_operands(h, W, bias, y, P, n_dims, "dense_mdn_forward_draws")
    logp = torch.empty(S * Bd, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.nfn_dense_mdn_forward_draws_x(n_centers, n_dims, H, S, Bd, _lib.ptr(h), _lib.ptr(W), _lib.ptr(bias),
                                                     _lib.ptr(y), y.shape[0], _lib.ptr(logp), _xf(xform),
                                                     _lib.current_stream(dev)))
    return logp


def dense_mdn_forward_backward_draws(h, W, bias, y, n_centers, n_dims, g_logp=None, g_scale=1.0, logp_sum=None,
                                     xform=None):
    """Forward + reverse sweep of the above.  Returns (logp [S*B], dh [S*B, H], dW [S, H, P], dbias [S, P])."""
    lib = _lib.load()
    P = mdn_param_size(n_centers, n_dims)
    h, W, bias, y, S, H, Bd, dev = _draws_operands(h, W, bias, y, P, n_dims, "dense_mdn_forward_backward_draws")
    logp = torch.empty(S * Bd, dtype=torch.float32, device=dev)
    dh = torch.empty((S * Bd, H), dtype=torch.float32, device=dev)
    dW = torch.zeros((S, H, P), dtype=torch.float32, device=dev)
    dbias = torch.zeros((S, P), dtype=torch.float32, device=dev)
    g_logp = _prep_g(g_logp, S * Bd, dev, "dense_mdn_forward_backward_draws")
    with torch.cuda.device(dev):
        _lib.check(lib.nfn_dense_mdn_forward_backward_draws_x(
            n_centers, n_dims, H, S, Bd, _lib.ptr(h), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(y), y.shape[0],
            _lib.ptr(g_logp), ctypes.c_float(g_scale), _lib.ptr(logp), _lib.ptr(dh), _lib.ptr(dW), _lib.ptr(dbias),
            _lib.ptr(logp_sum), _xf(xform), _lib.current_stream(dev)))
    return logp, dh, dW, dbias


class _ChainLogProb(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, y, flow_types, n_dims, trainable_base_dist):
        ctx.save_for_backward(t, y)
        ctx.cfg = (tuple(flow_types), n_dims, trainable_base_dist)
        return chain_forward(t, y, flow_types, n_dims, trainable_base_dist)

    @staticmethod
    def backward(ctx, g):
        t, y = ctx.saved_tensors
        flow_types, n_dims, tb = ctx.cfg
        need_dy = ctx.needs_input_grad[1]
        B = g.shape[0]
        if t.shape[0] == 1 and B != 1:
            t_rows = t.expand(B, t.shape[1]).contiguous()
        else:
            t_rows = t
        y_rows = y.expand(B, y.shape[1]).contiguous() if (need_dy and y.shape[0] != B) else y
        _, dt, dy = chain_forward_backward(t_rows, y_rows, flow_types, n_dims, tb, g_logp=g.contiguous(),
                                           want_dy=need_dy)
        if t.shape[0] == 1 and B != 1:
            dt = dt.sum(0, keepdim=True)
        if need_dy and y.shape[0] != B:
            dy = dy.sum(0, keepdim=True)
        return (dt if ctx.needs_input_grad[0] else None), (dy if need_dy else None), None, None, None


def chain_log_prob(t, y, flow_types, n_dims, trainable_base_dist):
    """Differentiable log_prob[B] (see module docstring)."""
    t = _as_f32_cuda(t, "t")
    y = _as_f32_cuda(y, "y", device=t.device)
    return _ChainLogProb.apply(t, y, tuple(flow_types), int(n_dims), bool(trainable_base_dist))


def flow_forward(flow_type, t, z, n_dims, want_z=True, want_fldj=True):
    """One bijector: returns (forward(z) [B,d] or None, fldj [B] or None)."""
    lib = _lib.load()
    t = _as_f32_cuda(t, "t")
    z = _as_f32_cuda(z, "z", device=t.device)
    if t.dim() != 2 or z.dim() != 2 or z.shape[1] != n_dims:
        raise ValueError("flow_forward: t must be [B, size], z must be [B, %d] or [1, %d]" % (n_dims, n_dims))
    B = t.shape[0]
    if z.shape[0] not in (B, 1):
        if B != 1:
            raise ValueError("flow_forward: z has %d rows, t has %d" % (z.shape[0], B))
        t = t.expand(z.shape[0], t.shape[1]).contiguous()
        B = z.shape[0]
    t, z = _aligned(t), _aligned(z)
    z_out = torch.empty((B, n_dims), dtype=torch.float32, device=t.device) if want_z else None
    fldj = torch.empty(B, dtype=torch.float32, device=t.device) if want_fldj else None
    with torch.cuda.device(t.device):
        _lib.check(lib.nfn_flow_forward(_lib.FLOW_CODES[flow_type], n_dims, _lib.ptr(t), _lib.ptr(z),
                                        z.shape[0], _lib.ptr(z_out), _lib.ptr(fldj), B,
                                        _lib.current_stream(t.device)))
    return z_out, fldj


# ----------------------------------------------------------------------------- MDN head
def mdn_param_size(n_centers, n_dims):
    return 2 * n_centers * n_dims + n_centers


def mdn_forward(t, y, n_centers, n_dims, xform=None):
    lib = _lib.load()
    t, y, B = _prep_ty(t, y, n_dims, mdn_param_size(n_centers, n_dims), "mdn_forward")
    logp = torch.empty(B, dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        _lib.check(lib.nfn_mdn_forward_x(n_centers, n_dims, _lib.ptr(t), _lib.ptr(y), y.shape[0],
                                         _lib.ptr(logp), B, _xf(xform), _lib.current_stream(t.device)))
    return logp


def mdn_forward_backward(t, y, n_centers, n_dims, g_logp=None, g_scale=1.0, want_dy=False,
                         logp_sum=None, dt_colsum=None, xform=None):
    lib = _lib.load()
    P = mdn_param_size(n_centers, n_dims)
    t, y, B = _prep_ty(t, y, n_dims, P, "mdn_forward_backward")
    dev = t.device
    logp = torch.empty(B, dtype=torch.float32, device=dev)
    dt = torch.empty((B, P), dtype=torch.float32, device=dev)
    dy = torch.empty((B, n_dims), dtype=torch.float32, device=dev) if want_dy else None
    if g_logp is not None:
        g_logp = _as_f32_cuda(g_logp, "g_logp", device=dev).reshape(-1)
    with torch.cuda.device(dev):
        _lib.check(lib.nfn_mdn_forward_backward_x(
            n_centers, n_dims, _lib.ptr(t), _lib.ptr(y), y.shape[0], _lib.ptr(g_logp),
            ctypes.c_float(g_scale), _lib.ptr(logp), _lib.ptr(dt), _lib.ptr(dy), _lib.ptr(logp_sum),
            _lib.ptr(dt_colsum), B, _xf(xform), _lib.current_stream(dev)))
    return logp, dt, dy


def dense_mdn_supported(hidden, n_centers, n_dims):
    """Whether the fused Dense(P)+MDN kernel can serve this shape: hidden width 16/32/48/64 and a 128-row tile of
    P = K (2 d + 1) columns (+ the layer's operands and accumulators) inside 220 KB of shared memory
    (``dense_smem_bytes`` in csrc/nfn_dense_chain.cuh)."""
    if hidden not in (16, 32, 48, 64) or n_centers < 1 or not 1 <= n_dims <= 8:
        return False
    P = mdn_param_size(n_centers, n_dims)
    S = P + 4 if (P % 4 == 0 and (P // 4) % 2 == 0) else P
    P8 = (P + 7) // 8 * 8
    PW = P8
    while PW % 32 not in (8, 24):
        PW += 1
    floats = 128 * S + 2 * 128 * (hidden + 4) + hidden * PW + P8 + 4 * hidden * P8 + 4 * P8
    return 4 * floats <= 220 * 1024


def dense_mdn_forward(h, W, bias, y, n_centers, n_dims, xform=None):
    """log_prob[B] of the K-component mixture with the emitting layer fused: t = h @ W + bias never touches HBM.
    h [B, H], W [H, P] (Keras kernel layout = torch ``linear.weight.t()``), bias [P], P = K (2 d + 1).
    Reference: Dense(output_size) of MaximumLikelihoodNNEstimator.py:43 + DistributionLayers.py:196-212."""
    lib = _lib.load()
    P = mdn_param_size(n_centers, n_dims)
    h = _aligned(_as_f32_cuda(h, "h"))
    dev = h.device
    W = _as_f32_cuda(W, "W", device=dev)
    bias = _as_f32_cuda(bias, "bias", device=dev)
    y = _aligned(_as_f32_cuda(y, "y", device=dev))
    B, H = h.shape
    assert tuple(W.shape) == (H, P) and tuple(bias.shape) == (P,), "W must be [H, P] and bias [P]"
    _check_y(y, n_dims, B, "dense_mdn_forward")
    logp = torch.empty(B, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.nfn_dense_mdn_forward_x(n_centers, n_dims, H, _lib.ptr(h), _lib.ptr(W), _lib.ptr(bias),
                                               _lib.ptr(y), y.shape[0], _lib.ptr(logp), B, _xf(xform),
                                               _lib.current_stream(dev)))
    return logp


def dense_mdn_forward_backward(h, W, bias, y, n_centers, n_dims, g_logp=None, g_scale=1.0, logp_sum=None, dW=None,
                               dbias=None, xform=None):
    """Fused layer + mixture head, forward and reverse sweep.  Returns (logp[B], dh[B,H], dW[H,P], dbias[P]);
    dW / dbias are accumulated into when given (else fresh zero tensors)."""
    lib = _lib.load()
    P = mdn_param_size(n_centers, n_dims)
    h = _aligned(_as_f32_cuda(h, "h"))
    dev = h.device
    W = _as_f32_cuda(W, "W", device=dev)
    bias = _as_f32_cuda(bias, "bias", device=dev)
    y = _aligned(_as_f32_cuda(y, "y", device=dev))
    B, H = h.shape
    assert tuple(W.shape) == (H, P) and tuple(bias.shape) == (P,), "W must be [H, P] and bias [P]"
    _check_y(y, n_dims, B, "dense_mdn_forward_backward")
    logp = torch.empty(B, dtype=torch.float32, device=dev)
    dh = torch.empty((B, H), dtype=torch.float32, device=dev)
    if dW is None:
        dW = torch.zeros((H, P), dtype=torch.float32, device=dev)
    if dbias is None:
        dbias = torch.zeros(P, dtype=torch.float32, device=dev)
    g_logp = _prep_g(g_logp, B, dev, "dense_mdn_forward_backward")
    with torch.cuda.device(dev):
        _lib.check(lib.nfn_dense_mdn_forward_backward_x(
            n_centers, n_dims, H, _lib.ptr(h), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(y), y.shape[0],
            _lib.ptr(g_logp), ctypes.c_float(g_scale), _lib.ptr(logp), _lib.ptr(dh), _lib.ptr(dW), _lib.ptr(dbias),
            _lib.ptr(logp_sum), B, _xf(xform), _lib.current_stream(dev)))
    return logp, dh, dW, dbias


class _MdnLogProb(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, y, n_centers, n_dims):
        ctx.save_for_backward(t, y)
        ctx.cfg = (n_centers, n_dims)
        return mdn_forward(t, y, n_centers, n_dims)

    @staticmethod
    def backward(ctx, g):
        t, y = ctx.saved_tensors
        K, d = ctx.cfg
        need_dy = ctx.needs_input_grad[1]
        B = g.shape[0]
        t_rows = t.expand(B, t.shape[1]).contiguous() if (t.shape[0] == 1 and B != 1) else t
        y_rows = y.expand(B, d).contiguous() if (need_dy and y.shape[0] != B) else y
        _, dt, dy = mdn_forward_backward(t_rows, y_rows, K, d, g_logp=g.contiguous(), want_dy=need_dy)
        if t.shape[0] == 1 and B != 1:
            dt = dt.sum(0, keepdim=True)
        if need_dy and y.shape[0] != B:
            dy = dy.sum(0, keepdim=True)
        return (dt if ctx.needs_input_grad[0] else None), (dy if need_dy else None), None, None


def mdn_log_prob(t, y, n_centers, n_dims):
    t = _as_f32_cuda(t, "t")
    y = _as_f32_cuda(y, "y", device=t.device)
    return _MdnLogProb.apply(t, y, int(n_centers), int(n_dims))


# ----------------------------------------------------------------------------- KMN head
def kmn_forward(t, y, locs, scales, xform=None):
    lib = _lib.load()
    locs = _as_f32_cuda(locs, "locs")
    M, d = locs.shape
    scales = _as_f32_cuda(scales, "scales", device=locs.device).reshape(-1)
    t, y, B = _prep_ty(t, y, d, M, "kmn_forward")
    logp = torch.empty(B, dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        _lib.check(lib.nfn_kmn_forward_x(M, d, _lib.ptr(t), _lib.ptr(y), y.shape[0], _lib.ptr(locs),
                                         _lib.ptr(scales), _lib.ptr(logp), B, _xf(xform),
                                         _lib.current_stream(t.device)))
    return logp


def kmn_forward_backward(t, y, locs, scales, g_logp=None, g_scale=1.0, want_dy=False, want_dscales=True,
                         logp_sum=None, xform=None):
    lib = _lib.load()
    locs = _as_f32_cuda(locs, "locs")
    M, d = locs.shape
    scales = _as_f32_cuda(scales, "scales", device=locs.device).reshape(-1)
    t, y, B = _prep_ty(t, y, d, M, "kmn_forward_backward")
    dev = t.device
    logp = torch.empty(B, dtype=torch.float32, device=dev)
    dt = torch.empty((B, M), dtype=torch.float32, device=dev)
    dy = torch.empty((B, d), dtype=torch.float32, device=dev) if want_dy else None
    dscales = torch.zeros(M, dtype=torch.float32, device=dev) if want_dscales else None
    if g_logp is not None:
        g_logp = _as_f32_cuda(g_logp, "g_logp", device=dev).reshape(-1)
    with torch.cuda.device(dev):
        _lib.check(lib.nfn_kmn_forward_backward_x(
            M, d, _lib.ptr(t), _lib.ptr(y), y.shape[0], _lib.ptr(locs), _lib.ptr(scales), _lib.ptr(g_logp),
            ctypes.c_float(g_scale), _lib.ptr(logp), _lib.ptr(dt), _lib.ptr(dy), _lib.ptr(dscales),
            _lib.ptr(logp_sum), B, _xf(xform), _lib.current_stream(dev)))
    return logp, dt, dy, dscales


class _KmnLogProb(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, y, locs, scales):
        ctx.save_for_backward(t, y, locs, scales)
        return kmn_forward(t, y, locs, scales)

    @staticmethod
    def backward(ctx, g):
        t, y, locs, scales = ctx.saved_tensors
        need_dy = ctx.needs_input_grad[1]
        B = g.shape[0]
        t_rows = t.expand(B, t.shape[1]).contiguous() if (t.shape[0] == 1 and B != 1) else t
        y_rows = y.expand(B, y.shape[1]).contiguous() if (need_dy and y.shape[0] != B) else y
        _, dt, dy, dsc = kmn_forward_backward(t_rows, y_rows, locs, scales, g_logp=g.contiguous(),
                                              want_dy=need_dy, want_dscales=ctx.needs_input_grad[3])
        if t.shape[0] == 1 and B != 1:
            dt = dt.sum(0, keepdim=True)
        if need_dy and y.shape[0] != B:
            dy = dy.sum(0, keepdim=True)
        if dsc is not None:
            dsc = dsc.reshape(scales.shape)
        return (dt if ctx.needs_input_grad[0] else None), (dy if need_dy else None), None, dsc


def kmn_log_prob(t, y, locs, scales):
    t = _as_f32_cuda(t, "t")
    y = _as_f32_cuda(y, "y", device=t.device)
    return _KmnLogProb.apply(t, y, _as_f32_cuda(locs, "locs", device=t.device),
                             _as_f32_cuda(scales, "scales", device=t.device))


# ----------------------------------------------------------------------------- epilogue
def logmeanexp_draws(logp_sb):
    """[S, B] -> [B]: logsumexp over posterior draws minus log S (BayesianNNEstimator.score)."""
    lib = _lib.load()
    x = _as_f32_cuda(logp_sb, "logp_sb")
    S, B = x.shape
    out = torch.empty(B, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.nfn_logmeanexp_draws(_lib.ptr(x), S, B, _lib.ptr(out), _lib.current_stream(x.device)))
    return out


def launch_count_reset():
    return int(_lib.load().nfn_launch_count_reset())


def set_math_mode(accurate):
    _lib.check(_lib.load().nfn_set_math_mode(1 if accurate else 0))


_OPTION_VALUES = {
    "chain_io": {"auto": -1, "cpasync": 0, "tma": 1},
    "dense_mma": {"auto": 0, "tc5": 1, "sync": 2},
}


def set_option(name, value):
    """Process-wide run-time switch of libnfn_b200 (include/nfn_b200.h: nfn_set_option).  ``value`` is an int,
    a bool, or for "chain_io" / "dense_mma" one of the symbolic names ("auto", "cpasync", "tma" / "tc5", "sync")."""
    if isinstance(value, str):
        value = _OPTION_VALUES[name][value]
    _lib.check(_lib.load().nfn_set_option(name.encode(), int(value)))


def get_option(name):
    return int(_lib.check(_lib.load().nfn_get_option(name.encode())))
