"""Bayesian estimator: mean-field variational dense layers + the density head
(reference estimators/BayesianNNEstimator.py, tfp.layers.DenseVariational semantics).

Weight-space work (sampling, exact Normal-Normal KL) is O(#weights) and stays in torch.  The
effect on the hot path is the S-draw posterior predictive: the reference loops 50 sequential
forward passes (BayesianNNEstimator.py:65-76); here the S weight draws are folded into the
batch -- one batched GEMM per layer, ONE head launch over S*B rows, then the [S, B] -> [B]
logsumexp epilogue kernel (nfn_logmeanexp_draws).
"""
import math

import numpy as np
import torch

from .. import functional as F
from ..DistributionLayers import MeanFieldLayer
from .BaseEstimator import BaseEstimator, _GaussianNoise, _Normalise
from .MaximumLikelihoodNNEstimator import ACTIVATIONS


class DenseVariational(torch.nn.Module):
    """y = act(x @ kernel + bias) with (kernel, bias) ~ q = mean-field normal; adds
    kl_weight * KL(q || prior) to the loss (exact, or one-sample estimate)."""

    def __init__(self, units, kl_weight, kl_use_exact, activation, map_mode, trainable_prior, prior_scale,
                 device=None):
        super().__init__()
        self._device = device
        self.units = units
        self.kl_weight = kl_weight
        self.kl_use_exact = kl_use_exact
        self.map_mode = map_mode
        self.trainable_prior = trainable_prior
        self.prior_scale = float(prior_scale)
        self.act = ACTIVATIONS[activation]()
        self.in_features = None
        self.register_parameter("posterior_params", None)
        self.register_parameter("prior_loc", None)
        self.last_kl = None

    def _materialise(self, in_features, device):
        self.in_features = in_features
        size = in_features * self.units + self.units
        # Keras initializer "normal" = RandomNormal(stddev=0.05); prior loc starts at zero
        self.posterior_params = torch.nn.Parameter(
            0.05 * torch.randn(size if self.map_mode else 2 * size, device=device))
        self.prior_loc = torch.nn.Parameter(torch.zeros(size, device=device), requires_grad=self.trainable_prior)
        self._post = MeanFieldLayer(size, scale=None, map_mode=self.map_mode)
        self._prior = MeanFieldLayer(size, scale=self.prior_scale)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        # a fresh layer has no parameters yet (its input width is unknown until the first call): take the
        # width from the checkpoint so that its tensors are loaded instead of reported as unexpected
        loc = state_dict.get(prefix + "prior_loc")
        if self.in_features is None and loc is not None:
            self._materialise((loc.numel() - self.units) // self.units, self._device or loc.device)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)

    def _dists(self):
        return self._post(self.posterior_params), self._prior(self.prior_loc)

    def _kl(self, q, r, w):
        if self.kl_use_exact:
            return torch.distributions.kl_divergence(q, r)
        return q.log_prob(w) - r.log_prob(w)

    def forward(self, x, n_draws=None, generator=None):
        """x: [B, in] (one shared draw, like the reference) or [S, B, in] with n_draws=S."""
        if self.in_features is None:
            self._materialise(x.shape[-1], x.device)
        q, r = self._dists()
        loc, scale = q.base_dist.loc, q.base_dist.scale
        nk = self.in_features * self.units
        if n_draws is None:
            if self.map_mode:
                w = loc
            else:
                w = loc + scale * torch.randn(loc.shape, device=loc.device, generator=generator)
            self.last_kl = self.kl_weight * self._kl(q, r, w)
            out = x @ w[:nk].view(self.in_features, self.units) + w[nk:]
        else:
            if self.map_mode:
                w = loc.expand(n_draws, -1)
            else:
                w = loc + scale * torch.randn((n_draws,) + tuple(loc.shape), device=loc.device, generator=generator)
            # exact KL does not depend on the draws; the one-sample estimate is averaged over them
            self.last_kl = self.kl_weight * (self._kl(q, r, None) if self.kl_use_exact
                                             else (q.log_prob(w) - r.log_prob(w)).mean())
            out = torch.baddbmm(w[:, nk:].unsqueeze(1), x, w[:, :nk].view(n_draws, self.in_features, self.units))
        return self.act(out)

    def refresh_kl(self, n_draws=None, generator=None):
        """The layer's KL term without data (a data-parallel rank whose shard of the mini-batch is empty): draws
        the weights exactly as ``forward`` would, so that the replicated weight-noise generators of all ranks stay
        in step, and leaves ``last_kl`` behind."""
        if self.in_features is None:
            return None
        q, r = self._dists()
        loc, scale = q.base_dist.loc, q.base_dist.scale
        if self.map_mode:
            w = loc if n_draws is None else loc.expand(n_draws, -1)
        else:
            shape = tuple(loc.shape) if n_draws is None else (n_draws,) + tuple(loc.shape)
            w = loc + scale * torch.randn(shape, device=loc.device, generator=generator)
        if n_draws is None:
            self.last_kl = self.kl_weight * self._kl(q, r, w)
        else:
            self.last_kl = self.kl_weight * (self._kl(q, r, None) if self.kl_use_exact
                                             else (q.log_prob(w) - r.log_prob(w)).mean())
        return self.last_kl


class BayesianNNEstimator(BaseEstimator):
    def __init__(self, dist_layer, kl_weight_scale, kl_use_exact=True, hidden_sizes=(10,), activation="tanh",
                 learning_rate=3e-2, noise_reg=("fixed_rate", 0.0), trainable_prior=False, map_mode=False,
                 prior_scale=1.0, random_seed=22, device=None, n_train_draws=1):
        torch.manual_seed(random_seed)
        torch.nn.Module.__init__(self)
        self.map_mode = map_mode
        from .BaseEstimator import default_device

        self._bayes_cfg = dict(trainable_prior=trainable_prior, prior_scale=prior_scale, map_mode=map_mode,
                               device=torch.device(device) if device is not None else default_device())
        layers = self._get_dense_layers(hidden_sizes=hidden_sizes, output_size=dist_layer.get_total_param_size(),
                                        posterior=None, prior=None, kl_weight_scale=kl_weight_scale,
                                        kl_use_exact=kl_use_exact, activation=activation)
        super().__init__(layers, dist_layer, noise_fn_type=noise_reg[0], noise_scale_factor=noise_reg[1],
                         random_seed=random_seed, device=device)
        self.map_mode = map_mode
        self.learning_rate = learning_rate
        # S Monte-Carlo weight draws per training step, folded into the batch (BASELINE config 4);
        # 1 = the reference's behaviour (one draw shared by the mini-batch)
        self.n_train_draws = int(n_train_draws)
        # instance-level noise levels (the reference shadows the class variables, :38-39)
        self.x_noise_std = 0.0
        self.y_noise_std = 0.0
        # all ranks must draw the same weights in data-parallel training
        self._wgen = None

    def _set_noise(self, std):
        self.x_noise_std = std
        self.y_noise_std = std

    def _get_dense_layers(self, hidden_sizes, output_size, posterior=None, prior=None, kl_weight_scale=1.0,
                          kl_use_exact=True, activation="relu"):
        assert type(hidden_sizes) == tuple or type(hidden_sizes) == list
        assert kl_weight_scale <= 1.0
        cfg = getattr(self, "_bayes_cfg", dict(trainable_prior=False, prior_scale=1.0, map_mode=False))
        mk = lambda units, act: DenseVariational(units, kl_weight_scale, kl_use_exact, act, cfg["map_mode"],
                                                 cfg["trainable_prior"], cfg["prior_scale"], cfg.get("device"))
        normalization = [_Normalise(self)]
        noise_reg = [_GaussianNoise(self, "x_noise_std")]
        hidden = [mk(size, activation) for size in hidden_sizes]
        output = [mk(output_size, "linear")]
        return normalization + noise_reg + hidden + output

    # ------------------------------------------------------------------ forward with folded draws
    def _weight_generator(self):
        if self._wgen is None:
            self._wgen = torch.Generator(device=self.device).manual_seed(self.random_seed + 1)
        return self._wgen

    def _graph_generators(self):
        return [self._weight_generator()]

    def params_from_x(self, x):
        h = self._to_dev(x)
        for layer in self.net:
            h = layer(h, generator=self._weight_generator()) if isinstance(layer, DenseVariational) else layer(h)
        return h

    def params_from_x_draws(self, x, n_draws):
        """t[S*B, P] for S posterior weight draws folded into the batch (draw-major)."""
        h = self._to_dev(x)
        h = self.net[1](self.net[0](h))
        h = h.unsqueeze(0).expand(n_draws, -1, -1)
        for layer in self.net[2:]:
            h = layer(h, n_draws=n_draws, generator=self._weight_generator())
        return h.reshape(n_draws * h.shape[1], h.shape[2])

    # ------------------------------------------------------------------ folded draws, fused kernels
    fuse_draws = True

    def _fused_draws_plan(self):
        """(first layer, emitting layer, padded hidden width) when the S-draw step can run on the folded-draw kernels:
        one hidden DenseVariational layer of <= 64 units fed by <= 8 inputs, a flow-chain or MDN head, no x noise."""
        from ..DistributionLayers import GaussianMixtureLayer, InverseNormalizingFlowLayer

        if not self.fuse_draws or self.map_mode:
            return None
        if not isinstance(self.dist_layer, (InverseNormalizingFlowLayer, GaussianMixtureLayer)):
            return None
        dense = [l for l in self.net if isinstance(l, DenseVariational)]
        if len(dense) != 2 or dense[0].in_features is None or dense[1].in_features is None:
            return None
        if self.training and self.x_noise_std > 0.0:
            return None
        l1, l2 = dense
        hp = (l1.units + 15) // 16 * 16
        name = {v: k for k, v in ACTIVATIONS.items()}.get(type(l1.act))
        if name is None or not F.dense_act_draws_supported(l1.in_features, l1.units, hp, name):
            return None
        if not F.dense_chain_supported(hp) or self.dist_layer.get_total_param_size() < 1:
            return None
        if isinstance(self.dist_layer, GaussianMixtureLayer) and not F.dense_mdn_supported(
                hp, self.dist_layer._n_centers, self.dist_layer._n_dims):
            return None
        return l1, l2, hp, name

    def _head_draws(self, h, W2, b2, y, backward, **kw):
        """The emitting layer + density head over the folded rows with per-draw weights, whichever head this is."""
        from ..DistributionLayers import GaussianMixtureLayer

        layer = self.dist_layer
        if isinstance(layer, GaussianMixtureLayer):
            fn = F.dense_mdn_forward_backward_draws if backward else F.dense_mdn_forward_draws
            return fn(h, W2, b2, y, layer._n_centers, layer._n_dims, **kw)
        fn = F.dense_chain_forward_backward_draws if backward else F.dense_chain_forward_draws
        return fn(h, W2, b2, y, layer._flow_types, layer._n_dims, layer._trainable_base_dist, **kw)

    def _draw_weights(self, layer, S, with_kl=True):
        """w [S, size] = loc + scale * eps (the same draws, in the same order, as DenseVariational.forward), and the
        layer's KL term left in ``last_kl``."""
        if layer.kl_use_exact and layer.posterior_params.is_cuda and with_kl:
            # one kernel each way for the sample and the exact KL (csrc/nfn_variational.cu)
            n = layer.prior_loc.numel()
            eps = torch.randn((S, n), device=layer.posterior_params.device, generator=self._weight_generator())
            w, kl = F.variational_sample(layer.posterior_params, layer.prior_loc, eps, layer.prior_scale)
            layer.last_kl = layer.kl_weight * kl
            return w
        q, r = layer._dists()
        loc, scale = q.base_dist.loc, q.base_dist.scale
        w = loc + scale * torch.randn((S,) + tuple(loc.shape), device=loc.device, generator=self._weight_generator())
        if with_kl:
            layer.last_kl = layer.kl_weight * (layer._kl(q, r, None) if layer.kl_use_exact
                                               else (q.log_prob(w) - r.log_prob(w)).mean())
        return w

    def _emitting_operands(self, l2, w2, hp):
        """The emitting layer's per-draw sample as the fused head's operands: W [S, hp, P] (rows past the layer's
        inputs zero, matching the zero columns of the padded hidden rows) and bias [S, P]."""
        S = w2.shape[0]
        P = self.dist_layer.get_total_param_size()
        nk = l2.in_features * P
        W2 = torch.zeros((S, hp, P), dtype=torch.float32, device=self.device)
        W2[:, :l2.in_features, :] = w2[:, :nk].view(S, l2.in_features, P)
        return W2, w2[:, nk:].contiguous()

    def _fused_draws_log_prob(self, plan, x, y, S, xform):
        """log p(y_b | x_b, w_s) for S fresh weight draws, [S, B]: two kernels (first layer over the folded rows,
        emitting layer + flow chain), no parameter tensor in HBM."""
        l1, l2, hp, act = plan
        layer = self.dist_layer
        with torch.no_grad():
            w1, w2 = self._draw_weights(l1, S), self._draw_weights(l2, S)   # (also leaves the layers' KL terms behind, like a forward pass)
            h = F.dense_act_forward_draws(x, w1, l1.units, act, hp, x_mean=self.x_mean, x_std=self.x_std)
            W2, b2 = self._emitting_operands(l2, w2, hp)
            logp = self._head_draws(h, W2, b2, y, False, xform=xform)
        return logp.view(S, x.shape[0])

    def _fused_draws_forward_backward(self, plan, xb, y, S, g_scale, logp_sum, xform):
        """x [B, in], raw y [B, d] -> per-draw weight samples (autograd leaves' children) and their gradients.
        Three kernels: first layer over S*B folded rows, emitting layer + flow chain + both backward GEMMs, first
        layer's per-draw weight gradient.  Neither t nor dt nor a repeated y ever exists in HBM."""
        l1, l2, hp, act = plan
        layer = self.dist_layer
        P = layer.get_total_param_size()
        w1, w2 = self._draw_weights(l1, S), self._draw_weights(l2, S)
        x = self._to_dev(xb)
        with torch.no_grad():
            h = F.dense_act_forward_draws(x, w1, l1.units, act, hp, x_mean=self.x_mean, x_std=self.x_std)
            nk = l2.in_features * P
            W2, b2 = self._emitting_operands(l2, w2, hp)
            _, dh, dW2, db2 = self._head_draws(h, W2, b2, y, True, g_scale=g_scale, logp_sum=logp_sum, xform=xform)
            dw1 = F.dense_act_backward_draws(x, h, dh, S, l1.units, act, x_mean=self.x_mean, x_std=self.x_std)
            dw2 = torch.cat([dW2[:, :l2.in_features, :].reshape(S, nk), db2], dim=1)
        return (w1, w2), (dw1, dw2)

    def _one_call_step(self, plan, xb, y, S, g_scale, logp_sum, xform, world):
        """The network part of the step as ONE library call (nfn_bayes_train_step): no autograd graph, no per-kernel
        dispatch; gradients are accumulated straight into the posterior parameters' .grad (which must exist)."""
        from ..DistributionLayers import GaussianMixtureLayer

        l1, l2, hp, act = plan
        x = self._to_dev(xb)
        B = x.shape[0]
        n1, n2 = l1.prior_loc.numel(), l2.prior_loc.numel()
        ws = getattr(self, "_bayes_ws", None)
        if ws is None or ws["key"] != (S, B, hp, n1, n2):
            f32 = dict(dtype=torch.float32, device=self.device)
            ws = dict(key=(S, B, hp, n1, n2), w1=torch.empty((S, n1), **f32), w2=torch.empty((S, n2), **f32),
                      dw1=torch.empty((S, n1), **f32), dw2=torch.empty((S, n2), **f32), h=torch.empty((S * B, hp), **f32),
                      dh=torch.empty((S * B, hp), **f32), logp=torch.empty(S * B, **f32),
                      kl=torch.zeros(2, dtype=torch.float64, device=self.device))
            self._bayes_ws = ws
        ws["kl"].zero_()
        gen = self._weight_generator()
        ops = []
        for l, n in ((l1, n1), (l2, n2)):
            if l.posterior_params.grad is None:
                l.posterior_params.grad = torch.zeros_like(l.posterior_params)
            if l.prior_loc.requires_grad and l.prior_loc.grad is None:
                l.prior_loc.grad = torch.zeros_like(l.prior_loc)
            ops.append(dict(posterior=l.posterior_params.detach(), prior_loc=l.prior_loc.detach(),
                            eps=torch.randn((S, n), device=self.device, generator=gen),
                            dposterior=l.posterior_params.grad,
                            dprior_loc=l.prior_loc.grad if l.prior_loc.requires_grad else None,
                            prior_scale=l.prior_scale, kl_grad=l.kl_weight / world))
        layer = self.dist_layer
        mdn = isinstance(layer, GaussianMixtureLayer)
        f32c = lambda t: t.to(device=self.device, dtype=torch.float32).contiguous()
        x, y = f32c(x), f32c(y)
        F.bayes_train_step(x, y, ops[0], ops[1], l1.units, act, hp, S, g_scale, ws, logp_sum,
                           flow_types=None if mdn else layer._flow_types, n_dims=layer._n_dims,
                           trainable_base_dist=True if mdn else layer._trainable_base_dist,
                           mdn_centers=layer._n_centers if mdn else 0, x_mean=f32c(self.x_mean), x_std=f32c(self.x_std), xform=xform)
        kl = ws["kl"].to(torch.float32)
        l1.last_kl, l2.last_kl = l1.kl_weight * kl[0], l2.kl_weight * kl[1]

    def _one_call_ok(self, plan):
        l1, l2 = plan[0], plan[1]
        return (l1.kl_use_exact and l2.kl_use_exact and l1.posterior_params.is_cuda
                and all(p is l1.posterior_params or p is l2.posterior_params or p is l1.prior_loc or p is l2.prior_loc
                        for p in self.parameters() if p.requires_grad))

    def train_step(self, xb, yb, global_batch=None):
        if self.n_train_draws <= 1 or self.map_mode:
            return super().train_step(xb, yb, global_batch=global_batch)
        # S draws folded into the batch: one batched GEMM per layer, ONE head launch over S*B rows
        import torch.distributed as dist

        self.train(True)
        S = self.n_train_draws
        B = xb.shape[0]
        Bg = global_batch or B
        world = dist.get_world_size() if dist.is_initialized() else 1
        reducer = self._grad_reducer() if world > 1 else None
        plan = self._fused_draws_plan() if B > 0 else None
        one_call = plan is not None and self._one_call_ok(plan)
        if reducer is not None:
            reducer.zero()
        elif one_call:
            self.optimizer.zero_grad(set_to_none=False)   # the library accumulates into the existing buffers
        else:
            self.optimizer.zero_grad(set_to_none=True)
        if B == 0:
            return self._empty_shard_step(reducer, Bg, world, denom=S * Bg)
        logp_sum = torch.zeros(1, dtype=torch.float64, device=self.device)
        if one_call:
            y = self._to_dev(yb)
            self._one_call_step(plan, xb, y, S, -1.0 / (S * Bg), logp_sum, self._xform(y.shape[1], training=True), world)
            extra = self._extra_loss()
        elif plan is not None:
            y = self._to_dev(yb)
            ws, dws = self._fused_draws_forward_backward(plan, xb, y, S, -1.0 / (S * Bg), logp_sum,
                                                         self._xform(y.shape[1], training=True))
            extra = self._extra_loss()
            torch.autograd.backward(list(ws) + [extra], list(dws) + [torch.full_like(extra, 1.0 / world)])
        else:
            t = self.params_from_x_draws(xb, S)
            # raw y: normalisation, training noise (independent per folded row) and the Jacobian run in the head kernel
            y = self._to_dev(yb).repeat(S, 1)
            dt = self._head_forward_backward(t, y, -1.0 / (S * Bg), logp_sum,
                                             xform=self._xform(y.shape[1], training=True))
            extra = self._extra_loss()
            torch.autograd.backward([t, extra], [dt, torch.full_like(extra, 1.0 / world)])
        if reducer is not None:
            logp_sum = self._reduce_step(reducer, logp_sum)
        self.optimizer.step()
        self._weights_epoch = getattr(self, "_weights_epoch", 0) + 1
        loss = -logp_sum.to(torch.float32) / (S * Bg) + extra.detach()
        return loss.reshape(())

    def _refresh_extra_loss(self):
        S = self.n_train_draws if (self.n_train_draws > 1 and not self.map_mode) else None
        for layer in self.net:
            if isinstance(layer, DenseVariational):
                layer.refresh_kl(n_draws=S, generator=self._weight_generator())
        return self._extra_loss()

    def _extra_loss(self):
        kls = [l.last_kl for l in self.net if isinstance(l, DenseVariational) and l.last_kl is not None]
        return torch.stack(kls).sum() if kls else None

    def fit(self, x, y, batch_size=None, epochs=None, verbose=1, **kwargs):
        self._assign_data_normalization(np.asarray(x), np.asarray(y))
        with torch.no_grad():
            self.params_from_x(np.asarray(x)[:2])
        if self.optimizer is None:
            self.optimizer = self._make_adam()
        return super().fit(x, y, batch_size=batch_size, epochs=epochs, verbose=verbose, **kwargs)

    def log_posterior_predictive(self, x_data, y_data, posterior_draws=None, max_rows=1 << 22):
        """log (1/S) sum_s p(y | x, w_s) per sample, S draws folded into the batch."""
        S = posterior_draws or (1 if self.map_mode else 50)
        x = self._to_dev(np.asarray(x_data, np.float32) if not torch.is_tensor(x_data) else x_data)
        y = self._to_dev(np.asarray(y_data, np.float32) if not torch.is_tensor(y_data) else y_data)
        B = x.shape[0]
        chunk = max(1, max_rows // S)
        out = torch.empty(B, dtype=torch.float32, device=self.device)
        self.train(False)
        plan = self._fused_draws_plan() if S > 1 else None
        with torch.no_grad():
            for lo in range(0, B, chunk):
                hi = min(B, lo + chunk)
                if plan is not None:
                    logp = self._fused_draws_log_prob(plan, x[lo:hi].contiguous(), y[lo:hi].contiguous(), S,
                                                      self._xform(y.shape[1]))
                    out[lo:hi] = F.logmeanexp_draws(logp)
                    continue
                t = self.params_from_x_draws(x[lo:hi], S)
                yy = y[lo:hi].repeat(S, 1)
                logp = self.dist_layer(t).log_prob_x(yy, self._xform(yy.shape[1]))
                out[lo:hi] = F.logmeanexp_draws(logp.view(S, hi - lo))
        return out

    def score(self, x_data, y_data, posterior_draws=None):
        return float(self.log_posterior_predictive(x_data, y_data, posterior_draws).mean())
