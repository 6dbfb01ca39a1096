"""GPU: out-of-bounds WRITES of the round-2 kernels, hunted with canaries (compute-sanitizer is closed on this pool).

Every output tensor the ``functional`` wrappers allocate while the guard is active lives inside a larger
sentinel-filled allocation; after the launches the bands on both sides must come back untouched.  The wrappers are
called exactly as the estimators call them (the launches under test are the product's own), at ragged row counts
that end in a partial tile.  ``tests/test_parity_gpu.py::test_outputs_stay_inside_their_buffers`` does the same for
the round-1 entry points through the raw C ABI."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

SENT = 12345.0
PAD = 1024   # elements on each side (4 KB of float32: keeps the interior 16-byte aligned)


class Guarded:
    """While active, ``torch.empty`` / ``torch.zeros`` of floating CUDA tensors return the interior of a padded,
    sentinel-filled allocation (inputs built with randn / full / tanh are not affected)."""

    def __init__(self):
        self.allocs = []

    def _alloc(self, orig, zero, args, kwargs):
        dtype = kwargs.get("dtype") or torch.float32
        device = kwargs.get("device")
        if device is None or torch.device(device).type != "cuda" or not dtype.is_floating_point or len(kwargs) > 2:
            return orig(*args, **kwargs)
        shape = args[0] if len(args) == 1 and isinstance(args[0], (tuple, list, torch.Size)) else args
        shape = tuple(int(s) for s in shape)
        n = math.prod(shape)
        buf = torch.full((n + 2 * PAD,), SENT, dtype=dtype, device=device)
        inner = buf[PAD: PAD + n]
        if zero:
            inner.zero_()
        self.allocs.append((buf, n, shape))
        return inner.view(shape)

    def __enter__(self):
        self._empty, self._zeros = torch.empty, torch.zeros
        torch.empty = lambda *a, **k: self._alloc(self._empty, False, a, k)
        torch.zeros = lambda *a, **k: self._alloc(self._zeros, True, a, k)
        return self

    def __exit__(self, *exc):
        torch.empty, torch.zeros = self._empty, self._zeros

    def check(self, what):
        """Bands intact for every allocation since the last check; returns a list of complaints."""
        torch.cuda.synchronize()
        bad = []
        for buf, n, shape in self.allocs:
            lo, hi = buf[:PAD], buf[PAD + n:]
            if not (bool((lo == SENT).all()) and bool((hi == SENT).all())):
                bad.append("%s: output of shape %s written outside its %d elements (%d below, %d above)" % (
                    what, shape, n, int((lo != SENT).sum()), int((hi != SENT).sum())))
        count = len(self.allocs)
        self.allocs = []
        return bad, count


def test_round2_kernels_write_only_inside_their_outputs(cuda_device, nfn_lib):
    from normalizingflownetwork_b200 import functional as F

    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(5)

    def rnd(*s, sc=1.0):
        return torch.randn(s, generator=g, device=dev) * sc

    problems, guarded_outputs = [], 0

    def case(what, fn):
        """Run one wrapper call under the guard; collect band violations and non-finite outputs."""
        nonlocal guarded_outputs
        with Guarded() as guard:
            try:
                outs = fn()
            except Exception as exc:  # noqa: BLE001 -- report every case, not only the first
                problems.append("%s: raised %s: %s" % (what, type(exc).__name__, exc))
                return
            bad, n = guard.check(what)
        problems.extend(bad)
        guarded_outputs += n
        if n == 0:
            problems.append("%s: no output went through the guard" % what)
        outs = outs if isinstance(outs, (tuple, list)) else (outs,)
        for i, o in enumerate(outs):
            if torch.is_tensor(o) and not bool(torch.isfinite(o).all()):
                problems.append("%s: output %d has non-finite values" % (what, i))

    H = 16
    for B in (1, 129, 1000, 128 * 37 + 5):
        # ---- emitting layer fused into the MDN head (ahead-of-time shapes: BASELINE config 5, the reference default)
        for K, d in ((20, 2), (5, 1)):
            P = F.mdn_param_size(K, d)
            assert F.dense_mdn_supported(H, K, d)
            args = (torch.tanh(rnd(B, H)), rnd(H, P, sc=0.1), rnd(P, sc=0.1), rnd(B, d))
            case("dense_mdn_forward_backward K=%d d=%d B=%d" % (K, d, B),
                 lambda a=args, K=K, d=d: F.dense_mdn_forward_backward(*a, K, d, g_scale=-1.0 / B))
            case("dense_mdn_forward K=%d d=%d B=%d" % (K, d, B), lambda a=args, K=K, d=d: F.dense_mdn_forward(*a, K, d))
        # ---- emitting layer fused into the KMN head (the reference's default: 50 centres x 2 bandwidths, 1-D y)
        M, d = 100, 1
        assert F.dense_kmn_supported(H, M, d)
        locs, scales = rnd(M, d), torch.full((M,), 0.4, device=dev)
        args = (torch.tanh(rnd(B, H)), rnd(H, M, sc=0.1), rnd(M, sc=0.1), rnd(B, d), locs, scales)
        case("dense_kmn_forward_backward B=%d" % B, lambda a=args: F.dense_kmn_forward_backward(*a, g_scale=-1.0 / B))
        case("dense_kmn_forward B=%d" % B, lambda a=args: F.dense_kmn_forward(*a))
        # ---- streaming KMN head (bandwidth gradient as a block reduce-scatter), 1-D and 2-D events
        for M, d in ((100, 1), (20, 2), (33, 3)):
            args = (rnd(B, M), rnd(B, d), rnd(M, d), torch.full((M,), 0.4, device=dev))
            case("kmn_forward_backward M=%d d=%d B=%d" % (M, d, B),
                 lambda a=args: F.kmn_forward_backward(*a, g_scale=-1.0 / B, want_dy=True))
        # ---- hidden layers: first layer with the x normalisation fused, 16 -> 16 (tensor-core backward), 16 -> 32
        for K, N, act, norm in ((1, 16, "tanh", True), (3, 16, "relu", True), (16, 16, "tanh", False), (16, 32, "tanh", False)):
            if not F.dense_act_supported(K, N, act):
                continue
            x, w, b = rnd(B, K), rnd(N, K, sc=0.3), rnd(N, sc=0.1)
            mean, std = (rnd(K, sc=0.1), torch.full((K,), 1.3, device=dev)) if norm else (None, None)
            out = F.dense_act_forward(x, w, b, act, x_mean=mean, x_std=std)
            case("dense_act_forward %d->%d %s B=%d" % (K, N, act, B),
                 lambda: F.dense_act_forward(x, w, b, act, x_mean=mean, x_std=std))
            up = rnd(B, N)
            case("dense_act_backward %d->%d %s B=%d" % (K, N, act, B),
                 lambda: F.dense_act_backward(x, out, up, w, act, need_dx=not norm, x_mean=mean, x_std=std))
        # ---- S posterior draws folded into the batch (BASELINE config 4's training step)
        S, K, N, NP = 4, 1, 10, 16
        assert F.dense_act_draws_supported(K, N, NP, "tanh")
        x, w = rnd(B, K), rnd(S, K * N + N, sc=0.7)
        hd = F.dense_act_forward_draws(x, w, N, "tanh", NP)
        case("dense_act_forward_draws B=%d" % B, lambda: F.dense_act_forward_draws(x, w, N, "tanh", NP))
        up = rnd(S * B, NP)
        case("dense_act_backward_draws B=%d" % B, lambda: F.dense_act_backward_draws(x, hd, up, S, N, "tanh"))
        ft, d, tb = ["radial"] * 5, 1, True
        P = F.chain_param_size(ft, d, tb)
        args = (hd, rnd(S, NP, P, sc=0.2), rnd(S, P, sc=0.1), rnd(B, d))
        case("dense_chain_forward_backward_draws B=%d" % B,
             lambda a=args: F.dense_chain_forward_backward_draws(*a, ft, d, tb, g_scale=-1.0 / (S * B)))
        case("dense_chain_forward_draws B=%d" % B, lambda a=args: F.dense_chain_forward_draws(*a, ft, d, tb))
        K5, d5 = 5, 1
        P5 = F.mdn_param_size(K5, d5)
        args = (hd, rnd(S, NP, P5, sc=0.2), rnd(S, P5, sc=0.1), rnd(B, d5))
        case("dense_mdn_forward_backward_draws B=%d" % B,
             lambda a=args: F.dense_mdn_forward_backward_draws(*a, K5, d5, g_scale=-1.0 / (S * B)))
        case("logmeanexp_draws B=%d" % B, lambda: F.logmeanexp_draws(rnd(S, B)))
    # ---- mean-field weight posterior: S samples + exact KL (odd sizes: vector tails)
    for n, S in ((37, 4), (186, 32), (1, 1)):
        params, prior, eps = rnd(2 * n, sc=0.3), rnd(n, sc=0.1), rnd(S, n)
        case("variational_sample n=%d S=%d" % (n, S), lambda: F.variational_sample(params, prior, eps, 1.0))
    assert not problems, "\n".join(problems)
    assert guarded_outputs >= 150
