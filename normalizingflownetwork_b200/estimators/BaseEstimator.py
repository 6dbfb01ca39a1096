"""BaseEstimator: the caller contract of the hot path (reference estimators/BaseEstimator.py).

Keeps ``fit / score / pdf / log_pdf`` and the data-normalisation / noise-regularisation
behaviour of the reference's Keras ``Sequential`` subclass, on a torch MLP (stock cuBLAS
GEMMs: the conditioning network is a few kFLOP per sample and not the hot path) with the
density head running in libnfn_b200.so.  One training step launches ONE fused
forward+reverse-sweep kernel for the head (cotangent -1/B folded in) and feeds ``dt`` to the
MLP backward; under torchrun the batch is sharded by rank and the flat gradient (+ loss) is
summed with one all-reduce.
"""
import math

import numpy as np
import torch
import torch.distributed as dist

from .. import functional as F
from ..DistributionLayers import (
    FusedDenseFlowChainDistribution,
    FusedDenseGaussianKernelsDistribution,
    FusedDenseGaussianMixtureDistribution,
    GaussianKernelsLayer,
    GaussianMixtureLayer,
    InverseNormalizingFlowLayer,
)


def default_device():
    if not torch.cuda.is_available():
        raise RuntimeError("normalizingflownetwork_b200 estimators need a CUDA device (no CPU path)")
    return torch.device("cuda", torch.cuda.current_device())


class _Normalise(torch.nn.Module):
    """Lambda x: (x - x_mean) / (x_std + 1e-8)  (MaximumLikelihoodNNEstimator.py:40).

    When the next Dense layer can take the normalisation as a fused prologue of its own kernel (nfn_mlp.cu:
    first layer, <= 4 inputs, no x noise this call) this module passes x through untouched and the layer
    reads the statistics itself (``fused_into`` is set by the estimator)."""

    def __init__(self, owner):
        super().__init__()
        self._owner = [owner]  # not a submodule
        self.fused_into = None  # the _Dense layer that normalises on load, or None

    def fused_now(self, x):
        o = self._owner[0]
        layer = self.fused_into
        return (layer is not None and x.is_cuda and x.dim() == 2 and not (o.training and o.x_noise_std > 0.0)
                and layer.can_fuse_xnorm(x))

    def forward(self, x):
        o = self._owner[0]
        if self.fused_now(x):
            self.fused_into.xnorm = (o.x_mean, o.x_std)   # consumed (and cleared) by the layer's next call
            return x
        return (x - o.x_mean) / (o.x_std + 1e-8)


class _GaussianNoise(torch.nn.Module):
    """tf.keras.layers.GaussianNoise: additive N(0, std) noise, active in training only."""

    def __init__(self, owner, which):
        super().__init__()
        self._owner = [owner]
        self._which = which

    def forward(self, x):
        std = getattr(self._owner[0], self._which)
        if self.training and std > 0.0:
            return x + std * torch.randn_like(x)
        return x


class BaseEstimator(torch.nn.Module):
    # class-level like the reference's tf.Variables (BaseEstimator.py:9-10); the Bayesian
    # subclass shadows them per instance
    x_noise_std = 0.0
    y_noise_std = 0.0

    def __init__(self, layers, dist_layer, noise_fn_type="fixed_rate", noise_scale_factor=0.0, random_seed=22,
                 device=None):
        super().__init__()
        self.device = torch.device(device) if device is not None else default_device()
        self.noise_fn_type = noise_fn_type
        self.noise_scale_factor = noise_scale_factor
        self.random_seed = random_seed
        self.net = torch.nn.Sequential(*layers)
        self.dist_layer = dist_layer
        for name in ("x_mean", "x_std", "y_mean", "y_std"):
            self.register_buffer(name, torch.zeros(1))
        self.x_std.fill_(1.0)
        self.y_std.fill_(1.0)
        self.optimizer = None
        self.history = []
        self.stop_training = False
        self.to(self.device)

    # ------------------------------------------------------------------ checkpoints
    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        # the normalisation statistics are registered as [1] placeholders and take the data's shape in fit():
        # give them the checkpoint's shape before the stock (shape-checking) copy
        for name in ("x_mean", "x_std", "y_mean", "y_std"):
            src = state_dict.get(prefix + name)
            cur = getattr(self, name)
            if src is not None and tuple(src.shape) != tuple(cur.shape):
                setattr(self, name, torch.zeros(tuple(src.shape), dtype=cur.dtype, device=cur.device))
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)
        self._invalidate_stats()

    # ------------------------------------------------------------------ forward
    def _to_dev(self, a):
        if torch.is_tensor(a):
            return a.to(device=self.device, dtype=torch.float32)
        return torch.as_tensor(np.asarray(a, dtype=np.float32), device=self.device)

    def params_from_x(self, x):
        """Network output t[B, P] for conditioning inputs x."""
        return self.net(self._to_dev(x))

    # ------------------------------------------------------------------ fused last layer
    fuse_last_layer = True

    def _fusable_last_layer(self):
        """The emitting Dense(P) layer if it can be folded into the head kernel:
        NF head with at least one parameter or MDN head, plain (non-variational) linear output layer whose
        input width is 16, 32, 48 or 64."""
        layer = self.dist_layer
        if not self.fuse_last_layer or not isinstance(layer, (InverseNormalizingFlowLayer, GaussianMixtureLayer,
                                                                  GaussianKernelsLayer)):
            return None
        last = self.net[-1]
        lin = getattr(last, "linear", None)
        if lin is None or not isinstance(lin, torch.nn.Linear) or isinstance(lin, torch.nn.LazyLinear):
            return None
        if not F.dense_chain_supported(lin.in_features) or layer.get_total_param_size() < 1:
            return None
        if isinstance(layer, GaussianMixtureLayer) and not F.dense_mdn_supported(lin.in_features, layer._n_centers,
                                                                                  layer._n_dims):
            return None
        if isinstance(layer, GaussianKernelsLayer) and not F.dense_kmn_supported(lin.in_features,
                                                                                  layer.get_total_param_size(),
                                                                                  layer.locs.shape[1]):
            return None
        return lin

    def _emitting_kernel(self, lin):
        """The emitting layer's weight in the Keras kernel layout [H, P] the fused head reads (= weight.t()),
        transposed once per weight VERSION, not once per call (a scoring call would otherwise spend a launch on it)."""
        # (the optimiser epoch is part of the key: the fused multi-tensor Adam and a replayed CUDA graph both update
        # the weights without touching torch's per-tensor version counter)
        key = (lin.weight.data_ptr(), lin.weight._version, getattr(self, "_weights_epoch", 0))
        if getattr(self, "_wt_key", None) != key:
            self._wt = lin.weight.detach().t().contiguous()
            self._wt_key = key
        return self._wt

    def hidden_from_x(self, x):
        h = self._to_dev(x)
        for layer in list(self.net)[:-1]:
            h = layer(h)
        return h

    def forward(self, x, training=False):
        was = self.training
        self.train(bool(training))
        try:
            with torch.set_grad_enabled(bool(training)):
                lin = self._fusable_last_layer()
                if lin is not None and not training:
                    layer = self.dist_layer
                    if isinstance(layer, GaussianMixtureLayer):
                        return FusedDenseGaussianMixtureDistribution(self.hidden_from_x(x), self._emitting_kernel(lin),
                                                                     lin.bias, layer._n_centers, layer._n_dims)
                    if isinstance(layer, GaussianKernelsLayer):
                        return FusedDenseGaussianKernelsDistribution(self.hidden_from_x(x), self._emitting_kernel(lin),
                                                                     lin.bias, layer.locs, layer.scale_model())
                    return FusedDenseFlowChainDistribution(self.hidden_from_x(x), self._emitting_kernel(lin),
                                                           lin.bias, layer._flow_types, layer._n_dims,
                                                           layer._trainable_base_dist)
                return self.dist_layer(self.params_from_x(x))
        finally:
            self.train(was)

    def call(self, x, training=False):
        return self.forward(x, training=training)

    # ------------------------------------------------------------------ data handling
    def _assign_data_normalization(self, x, y):
        x, y = np.asarray(x), np.asarray(y)
        for name, v in (("x_mean", np.mean(x, axis=0, dtype=np.float32)), ("x_std", np.std(x, axis=0, dtype=np.float32)),
                        ("y_mean", np.mean(y, axis=0, dtype=np.float32)), ("y_std", np.std(y, axis=0, dtype=np.float32))):
            new = torch.as_tensor(v, device=self.device)
            cur = getattr(self, name, None)
            # in place when possible: captured CUDA graphs (train step, log_pdf) hold these by address
            if torch.is_tensor(cur) and cur.shape == new.shape and cur.device == new.device and cur.dtype == new.dtype:
                cur.copy_(new)
            else:
                setattr(self, name, new)
        self._invalidate_stats()

    def _assign_noise_regularisation(self, n_dims, n_datapoints):
        assert self.noise_fn_type in ["rule_of_thumb", "fixed_rate"]
        if self.noise_fn_type == "rule_of_thumb":
            noise_std = self.noise_scale_factor * (n_datapoints + 1) ** (-1 / (4 + n_dims))
        else:
            noise_std = self.noise_scale_factor
        self._set_noise(float(noise_std))

    def _set_noise(self, std):
        type(self).x_noise_std = std
        type(self).y_noise_std = std

    def _y_input(self, y, training):
        """y normalisation + training-only noise (BaseEstimator.py:61-69)."""
        y = (self._to_dev(y) - self.y_mean) / self.y_std
        if training and self.y_noise_std > 0.0:
            y = y + self.y_noise_std * torch.randn_like(y)
        return y

    # ------------------------------------------------------------------ fused y pipeline
    def _invalidate_stats(self):
        self._stats_host = None
        self._stats_version = getattr(self, "_stats_version", 0) + 1

    def _y_stats_host(self):
        """Host copies of (y_mean, y_std, sum log y_std): kernel arguments by value, read back from the device once
        per change of the statistics (fit / load_state_dict), never on a scoring or training call."""
        st = getattr(self, "_stats_host", None)
        if st is None:
            mean = self.y_mean.detach().double().cpu().reshape(-1)
            std = self.y_std.detach().double().cpu().reshape(-1)
            # the Jacobian term in float32 like the reference's tf.reduce_sum(tf.math.log(y_std))
            shift = float(torch.sum(torch.log(self.y_std.detach().float())).cpu())
            st = (mean.tolist(), std.tolist(), shift)
            self._stats_host = st
        return st

    def _xform(self, n_dims, training=False, exp_out=False):
        """The event transform of one head launch: normalisation, training-only noise, -sum log y_std, exp."""
        mean, std, shift = self._y_stats_host()
        if len(mean) != n_dims:   # [1] placeholders before the first fit
            mean, std = [mean[0]] * n_dims, [std[0]] * n_dims
        noise = float(self.y_noise_std) if training else 0.0
        ctr = None
        if noise > 0.0:
            if getattr(self, "_noise_ctr", None) is None:
                self._noise_ctr = torch.zeros(1, dtype=torch.int64, device=self.device)
            self._noise_ctr += 1      # device-side step counter: a captured graph draws fresh noise per replay
            ctr = self._noise_ctr
        return F.make_xform(n_dims, mean, std, logp_shift=-shift, noise_std=noise, seed=self.random_seed,
                            offset_dev=ctr, exp_out=exp_out)

    def _get_input_model(self):
        """The reference's y input model as a callable ``(y, training=False)``: normalisation followed by
        noise that is active in training only (BaseEstimator.py:61-69, tests/test_noise_reg.py:55-75)."""
        return lambda y, training=False: self._y_input(y, training)

    def _log_ystd_sum(self):
        return torch.sum(torch.log(self.y_std))

    # ------------------------------------------------------------------ head: fused fwd + reverse sweep
    def _head_forward_backward(self, t, y, g_scale, logp_sum, xform=None):
        """Launches the fused kernel of the head; returns dt (and accumulates grads of the
        head's own trainable parameters, i.e. the KMN bandwidths).  With ``xform`` y is the RAW event and
        ``logp_sum`` receives sum(logp - sum log y_std)."""
        layer = self.dist_layer
        td = t.detach()
        if isinstance(layer, InverseNormalizingFlowLayer):
            _, dt, _ = F.chain_forward_backward(td, y, layer._flow_types, layer._n_dims, layer._trainable_base_dist,
                                                g_scale=g_scale, logp_sum=logp_sum, xform=xform)
        elif isinstance(layer, GaussianMixtureLayer):
            _, dt, _ = F.mdn_forward_backward(td, y, layer._n_centers, layer._n_dims, g_scale=g_scale,
                                              logp_sum=logp_sum, xform=xform)
        elif isinstance(layer, GaussianKernelsLayer):
            scales = layer.scale_model()
            _, dt, _, dsc = F.kmn_forward_backward(td, y, layer.locs, scales.detach(), g_scale=g_scale,
                                                   logp_sum=logp_sum, xform=xform)
            if scales.requires_grad:
                scales.backward(dsc)
        else:
            raise TypeError("unsupported distribution layer %r" % type(layer).__name__)
        return dt

    def _extra_loss(self):
        """Regulariser added to the mean NLL (KL term of the Bayesian estimators)."""
        return None

    def train_step(self, xb, yb, global_batch=None):
        """One optimiser step on a (local) mini-batch; returns the device scalar loss
        (mean NLL over the global batch + regulariser)."""
        self.train(True)
        B = xb.shape[0]
        Bg = global_batch or B
        world = dist.get_world_size() if dist.is_initialized() else 1
        reducer = self._grad_reducer() if world > 1 else None
        if reducer is not None:
            reducer.zero()     # gradients live in ONE flat buffer: every rank all-reduces it once per step
        else:
            self.optimizer.zero_grad(set_to_none=True)
        if B == 0:
            return self._empty_shard_step(reducer, Bg, world)
        # y stays RAW: normalisation, the training-only noise and the -sum log y_std Jacobian run inside the head
        # kernel (reference BaseEstimator.py:55-69), so logp_sum already is sum_b(log p_b - sum log y_std)
        y = self._to_dev(yb)
        xf = self._xform(y.shape[1], training=True)
        logp_sum = torch.zeros(1, dtype=torch.float64, device=self.device)
        lin = self._fusable_last_layer()
        if lin is not None and self._extra_loss() is None:
            # ONE kernel for [Dense(P) forward, flow chain forward + reverse sweep, Dense(P) backward]:
            # t / dt never touch HBM; the kernel returns dh, dW, dbias
            layer = self.dist_layer
            h = self.hidden_from_x(xb)
            if isinstance(layer, GaussianMixtureLayer):
                _, dh, dW, db = F.dense_mdn_forward_backward(
                    h.detach(), self._emitting_kernel(lin), lin.bias.detach(), y, layer._n_centers, layer._n_dims,
                    g_scale=-1.0 / Bg, logp_sum=logp_sum, xform=xf)
            elif isinstance(layer, GaussianKernelsLayer):
                scales = layer.scale_model()
                _, dh, dW, db, dsc = F.dense_kmn_forward_backward(
                    h.detach(), self._emitting_kernel(lin), lin.bias.detach(), y, layer.locs, scales.detach(),
                    g_scale=-1.0 / Bg, want_dscales=scales.requires_grad, logp_sum=logp_sum, xform=xf)
                if scales.requires_grad:
                    scales.backward(dsc)
            else:
                _, dh, dW, db = F.dense_chain_forward_backward(
                    h.detach(), self._emitting_kernel(lin), lin.bias.detach(), y, layer._flow_types,
                    layer._n_dims, layer._trainable_base_dist, g_scale=-1.0 / Bg, logp_sum=logp_sum, xform=xf)
            if reducer is not None:     # accumulate into the flat buffer's views
                lin.weight.grad.copy_(dW.t())
                lin.bias.grad.copy_(db)
            else:
                lin.weight.grad = dW.t().contiguous()
                lin.bias.grad = db
            if h.requires_grad:
                h.backward(dh)
        else:
            t = self.params_from_x(xb)
            dt = self._head_forward_backward(t, y, -1.0 / Bg, logp_sum, xform=xf)
            extra = self._extra_loss()
            if extra is not None:
                # the regulariser is replicated on every rank: weight 1/world so that the summed
                # gradient counts it once
                torch.autograd.backward([t, extra], [dt, torch.full_like(extra, 1.0 / world)])
            else:
                t.backward(dt)
        extra = self._extra_loss()
        if reducer is not None:
            logp_sum = self._reduce_step(reducer, logp_sum)
        self.optimizer.step()
        self._weights_epoch = getattr(self, "_weights_epoch", 0) + 1
        loss = -logp_sum.to(torch.float32) / Bg
        if extra is not None:
            loss = loss + extra.detach()
        return loss.reshape(())

    # ------------------------------------------------------------------ CUDA-graph train step
    def capture_train_step(self, batch_size, x_dim, y_dim, global_batch=None):
        """Capture one whole optimiser step (MLP forward, fused head kernel, MLP backward, Adam)
        into a CUDA graph for mini-batches of exactly ``batch_size`` rows.  Small batches are
        launch-latency-bound (config 1: 2048 rows = 197 KB of head traffic); replaying one graph
        removes ~30 separate launches per step.  Single-process only: a captured step that contains the
        data-parallel all-reduce of the flat gradient buffer did not complete on 2 GPUs (measured: the replay
        never returned), so data-parallel training stays eager."""
        assert self.optimizer is not None, "call fit() once (or _ensure_optimizer) before capturing"
        assert not (dist.is_initialized() and dist.get_world_size() > 1), "graph capture is single-GPU"
        for g in self.optimizer.param_groups:
            g["capturable"] = True
        # Adam's step counters must live on the device for a capturable optimiser
        for st in self.optimizer.state.values():
            if "step" in st and torch.is_tensor(st["step"]) and not st["step"].is_cuda:
                st["step"] = st["step"].to(self.device)
        self._gx = torch.zeros((batch_size, x_dim), device=self.device)
        self._gy = torch.zeros((batch_size, y_dim), device=self.device)
        saved_model = {k: v.detach().clone() for k, v in self.state_dict().items()}
        # optimiser state tensors are baked into the graph by address: remember their values and
        # put them back IN PLACE after the warm-up / capture steps (which run on zeros)
        saved_opt = {p: {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                     for p, st in self.optimizer.state.items()}
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):  # warm-up outside capture: kernel attributes, JIT, allocator pools
                self.train_step(self._gx, self._gy, global_batch=global_batch)
        torch.cuda.current_stream(self.device).wait_stream(side)
        self._graph = torch.cuda.CUDAGraph()
        for gen in self._graph_generators():   # private torch generators drawn from inside the step
            self._graph.register_generator_state(gen)
        with torch.cuda.graph(self._graph):
            self._graph_loss = self.train_step(self._gx, self._gy, global_batch=global_batch)
        # the warm-up and capture steps ran on zeros: restore the real state
        self.load_state_dict(saved_model)
        with torch.no_grad():
            for p, st in self.optimizer.state.items():
                old = saved_opt.get(p, {})
                for k, v in st.items():
                    if torch.is_tensor(v):
                        if k in old:
                            v.copy_(old[k])
                        else:
                            v.zero_()
        self._graph_batch = batch_size
        # the head kernels take the normalisation statistics BY VALUE: a captured step is tied to them
        self._graph_stats_version = getattr(self, "_stats_version", 0)
        return self

    def _make_adam(self):
        """Adam as the reference configures it (tf.compat.v2.optimizers.Adam, epsilon 1e-7); on the GPU the fused
        multi-tensor implementation: the whole update is one launch instead of ~10."""
        params = [p for p in self.parameters()]
        fused = bool(params) and all(p.is_cuda for p in params)
        return torch.optim.Adam(params, lr=self.learning_rate, eps=1e-7, fused=fused)

    def _graph_generators(self):
        """torch.Generator objects the training step draws from (they must be registered with a capturing graph)."""
        return []

    def train_step_graphed(self, xb, yb):
        """Replay the captured step on a new mini-batch (device tensors of the captured shape)."""
        self._gx.copy_(xb)
        self._gy.copy_(yb)
        self._graph.replay()
        self._weights_epoch = getattr(self, "_weights_epoch", 0) + 1
        return self._graph_loss

    def capture_log_pdf(self, batch_size, x_dim, y_dim):
        """Capture ``log_pdf`` for batches of exactly ``batch_size`` rows into a CUDA graph: small scoring
        batches (config 1: 2048 rows) are launch-latency-bound, one replay replaces ~10 launches and their
        Python dispatch.  The graph reads the CURRENT weights by address (in-place optimiser updates are
        seen; re-capture after load_state_dict of new tensors or a change of normalisation statistics)."""
        self._sx = torch.zeros((batch_size, x_dim), device=self.device)
        self._sy = torch.zeros((batch_size, y_dim), device=self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):  # warm-up outside capture: kernel attributes, JIT, allocator pools
                self.log_pdf(self._sx, self._sy)
        torch.cuda.current_stream(self.device).wait_stream(side)
        self._score_graph = torch.cuda.CUDAGraph()
        self._wt_key = None   # the emitting layer's transposed kernel is re-formed INSIDE the graph: replays see new weights
        with torch.cuda.graph(self._score_graph):
            self._score_out = self.log_pdf(self._sx, self._sy)
        self._score_stats_version = getattr(self, "_stats_version", 0)
        return self

    def log_pdf_graphed(self, x, y):
        """Replay the captured scoring graph on a new batch of the captured shape (device tensors, or host
        arrays which are copied in); the returned tensor is the graph's static output buffer (clone it to
        keep it across replays)."""
        if getattr(self, "_score_stats_version", None) != getattr(self, "_stats_version", 0):
            # new normalisation statistics since the capture (they are kernel arguments by value): capture again
            self.capture_log_pdf(self._sx.shape[0], self._sx.shape[1], self._sy.shape[1])
        self._sx.copy_(x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x, dtype=np.float32)))
        self._sy.copy_(y if torch.is_tensor(y) else torch.as_tensor(np.asarray(y, dtype=np.float32)))
        self._score_graph.replay()
        return self._score_out

    # ------------------------------------------------------------------ data-parallel step
    def _grad_reducer(self):
        """The flat gradient buffer of this model (parallel.FlatGradReducer), built at the first data-parallel step
        (all lazy layers exist by then) and rebuilt if the set of trainable parameters changes."""
        from ..parallel import FlatGradReducer

        params = [p for p in self.parameters() if p.requires_grad]
        red = getattr(self, "_reducer", None)
        if red is None or len(red.params) != len(params) or any(a is not b for a, b in zip(red.params, params)):
            red = FlatGradReducer(params, n_scalars=1)
            self._reducer = red
        return red

    def _reduce_step(self, reducer, logp_sum):
        """ONE float32 all-reduce of [all parameter gradients | sum logp (hi, lo)]; returns the global sum logp."""
        reducer.put_scalars(logp_sum.reshape(-1))
        reducer.reduce()
        return reducer.get_scalars().reshape(1)

    def _empty_shard_step(self, reducer, Bg, world, denom=None):
        """This rank's shard of the mini-batch is empty (tail batch with fewer rows than ranks).  It still takes
        part in the step: zero data gradients, its share of the replicated regulariser, the same all-reduce and
        optimiser step as every other rank -- skipping would pair its NEXT collective with the peers' current one."""
        logp_sum = torch.zeros(1, dtype=torch.float64, device=self.device)
        extra = self._refresh_extra_loss()
        if extra is not None:
            extra.backward(torch.full_like(extra, 1.0 / world))
        if reducer is not None:
            logp_sum = self._reduce_step(reducer, logp_sum)
        self.optimizer.step()
        self._weights_epoch = getattr(self, "_weights_epoch", 0) + 1
        loss = -logp_sum.to(torch.float32) / (denom or Bg)
        if extra is not None:
            loss = loss + extra.detach()
        return loss.reshape(())

    def _refresh_extra_loss(self):
        """The regulariser without a forward pass over data (subclasses with one override this)."""
        return None

    # ------------------------------------------------------------------ Keras-like API
    def fit(self, x, y, batch_size=None, epochs=None, verbose=1, shuffle=True, cuda_graph=False, **kwargs):
        x, y = np.asarray(x), np.asarray(y)
        self._assign_data_normalization(x, y)
        assert len(x.shape) == len(y.shape) == 2, "Please pass a matrix not a vector"
        self._assign_noise_regularisation(n_dims=x.shape[1] + y.shape[1], n_datapoints=x.shape[0])
        batch_size = batch_size or 32          # Keras default mini-batch
        epochs = epochs or 1
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
        xd, yd = self._to_dev(x), self._to_dev(y)
        n = xd.shape[0]
        gen = torch.Generator(device="cpu").manual_seed(self.random_seed)
        self.stop_training = False
        if cuda_graph and world == 1 and n >= batch_size:
            if (getattr(self, "_graph_batch", None) != batch_size
                    or getattr(self, "_graph_stats_version", None) != getattr(self, "_stats_version", 0)):
                self.capture_train_step(batch_size, xd.shape[1], yd.shape[1])
        else:
            cuda_graph = False
        for epoch in range(epochs):
            perm = torch.randperm(n, generator=gen).to(self.device) if shuffle else torch.arange(n, device=self.device)
            losses = []
            for lo in range(0, n, batch_size):
                idx = perm[lo: lo + batch_size]
                gb = idx.numel()
                if world > 1:
                    from ..parallel import shard_rows
                    a, b = shard_rows(gb, rank, world)
                    idx = idx[a:b]
                if idx.numel() == 0 and world == 1:
                    continue
                if cuda_graph and idx.numel() == batch_size:
                    losses.append(self.train_step_graphed(xd[idx], yd[idx]).clone())
                else:
                    losses.append(self.train_step(xd[idx], yd[idx], global_batch=gb))
            ep_loss = torch.stack(losses).mean().item()   # one host sync per epoch
            self.history.append(ep_loss)
            if verbose:
                print("Epoch %d/%d - loss: %.4f" % (epoch + 1, epochs, ep_loss))
            if not math.isfinite(ep_loss):                # tf.keras.callbacks.TerminateOnNaN
                self.stop_training = True
                break
        self.train(False)
        return self

    def _neg_log_likelihood(self, x, y, training=False):
        """Per-sample NLL incl. the normalisation Jacobian (BaseEstimator.py:55-59)."""
        dist_ = self.forward(x, training=training)
        if not training and hasattr(dist_, "log_prob_x"):   # one launch: y pipeline fused into the head
            yd = self._to_dev(y)
            return -dist_.log_prob_x(yd, self._xform(yd.shape[-1]))
        return -dist_.log_prob(self._y_input(y, training)) + self._log_ystd_sum()

    def _get_neg_log_likelihood(self):
        return lambda y, p_y: -p_y.log_prob(self._y_input(y, self.training)) + self._log_ystd_sum()

    def evaluate(self, x, y, **kwargs):
        with torch.no_grad():
            v = self._neg_log_likelihood(x, y).mean()
            extra = self._extra_loss()
            return float(v + (extra if extra is not None else 0.0))

    def score(self, x_data, y_data):
        with torch.no_grad():
            return float(-self._neg_log_likelihood(np.asarray(x_data, np.float32), np.asarray(y_data, np.float32)).mean())

    def pdf(self, x, y):
        assert tuple(np.shape(x)) == tuple(np.shape(y))
        with torch.no_grad():
            output = self.forward(x)
            yd = self._to_dev(y)
            if hasattr(output, "log_prob_x"):   # exp(log p - sum log y_std) written by the kernel itself
                return output.log_prob_x(yd, self._xform(yd.shape[-1], exp_out=True))
            y_circ = (yd - self.y_mean) / self.y_std
            return output.prob(y_circ) / torch.prod(self.y_std)

    def pdf_grid(self, x, y_values):
        """Density heat-map [len(y_values), len(x)] of p(y | x) in data units: the reference's plot_model
        loop (evaluation/visualization/flow_plotting.py:33-53) as ONE kernel launch for NF heads."""
        with torch.no_grad():
            output = self.forward(x)
            y_raw = self._to_dev(np.asarray(y_values, np.float32).reshape(-1, self.y_mean.numel()))
            if hasattr(output, "log_prob_grid"):
                return output.log_prob_grid(y_raw, xform=self._xform(y_raw.shape[-1], exp_out=True))
            y_circ = (y_raw - self.y_mean) / self.y_std
            lp = torch.stack([output.log_prob(y_circ[i:i + 1]) for i in range(y_circ.shape[0])])
            return torch.exp(lp - self._log_ystd_sum())

    def log_pdf(self, x, y):
        x = np.asarray(x, dtype=np.float32) if not torch.is_tensor(x) else x
        y = np.asarray(y, dtype=np.float32) if not torch.is_tensor(y) else y
        assert tuple(x.shape) == tuple(y.shape)
        with torch.no_grad():
            output = self.forward(x)
            assert output.event_shape == y.shape[-1]
            yd = self._to_dev(y)
            if hasattr(output, "log_prob_x"):   # normalisation and -sum log y_std inside the head kernel
                return output.log_prob_x(yd, self._xform(yd.shape[-1]))
            y_circ = (yd - self.y_mean) / self.y_std
            return output.log_prob(y_circ) - self._log_ystd_sum()
