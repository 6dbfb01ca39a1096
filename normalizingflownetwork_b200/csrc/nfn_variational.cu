// nfn_variational.cu -- weight-space arithmetic of the Bayesian estimators' mean-field layers (sm_100a).
//
// tfp.layers.DenseVariational with the reference's posterior / prior (estimators/DistributionLayers.py:17-71,
// BayesianNNEstimator.py:78-118): the posterior over a layer's flat [kernel | bias] vector is an independent normal
// with loc = params[0:n], scale = 1e-3 + softplus(c0 + 0.05 params[n:2n]); the prior is N(prior_loc, prior_scale).
// A training step needs S weight samples w_s = loc + scale * eps_s and the exact KL(q || prior).  This is
// O(#weights) -- a few hundred numbers -- but in torch it is ~40 elementwise / reduction launches per layer and
// step, forward and backward, which is what an eager S-draw training step spends its time on once the O(batch)
// work runs in three kernels (measured: 1.75 ms of launches around 0.3 ms of kernels).  Here it is one launch each way:
//
//   forward    w[s][i] = loc_i + sigma_i eps[s][i] ;   kl += sum_i log(sr / sigma_i) + (sigma_i^2 + (loc_i - mr_i)^2) / (2 sr^2) - 1/2
//   backward   dparams[i]     += sum_s dw[s][i] + gkl (loc_i - mr_i) / sr^2
//              dparams[n + i] += (sum_s dw[s][i] eps[s][i] + gkl (sigma_i / sr^2 - 1 / sigma_i)) * 0.05 sigmoid(c0 + 0.05 raw_i)
//              dprior_loc[i]  += -gkl (loc_i - mr_i) / sr^2                                  (trainable prior only)
//
// float32 per element (libm softplus / log, like torch), the KL sum in float64.
#include <cuda_runtime.h>

#include "nfn_common.h"

namespace nfn {
namespace {

constexpr int kVT = 256;

__device__ __forceinline__ float post_sigma(float raw, float* sigmoid_out) {
  const float a = fmaf(0.05f, raw, kC0);
  // softplus with torch's threshold-free, overflow-safe form: max(a, 0) + log1p(exp(-|a|))
  const float e = expf(-fabsf(a));
  const float sp = fmaxf(a, 0.0f) + log1pf(e);
  if (sigmoid_out) *sigmoid_out = (a >= 0.0f) ? 1.0f / (1.0f + e) : e / (1.0f + e);
  return 1e-3f + sp;
}

__global__ void __launch_bounds__(kVT) variational_fwd(const float* __restrict__ params, const float* __restrict__ prior_loc,
                                                     float prior_scale, const float* __restrict__ eps, int n, int S,
                                                     float* __restrict__ w, double* __restrict__ kl) {
  __shared__ double red[kVT / 32];
  double acc = 0.0;
  const float inv_sr2 = 1.0f / (prior_scale * prior_scale);
  for (int i = blockIdx.x * kVT + threadIdx.x; i < n; i += gridDim.x * kVT) {
    const float loc = __ldg(params + i);
    const float sigma = post_sigma(__ldg(params + n + i), nullptr);
    for (int s = 0; s < S; ++s) w[(long long)s * n + i] = fmaf(sigma, __ldg(eps + (long long)s * n + i), loc);
    const float d = loc - __ldg(prior_loc + i);
    acc += (double)(logf(prior_scale / sigma) + 0.5f * (sigma * sigma + d * d) * inv_sr2 - 0.5f);
  }
  if (kl) {
    const double sblk = block_sum<kVT>(acc, red);
    if (threadIdx.x == 0) atomicAdd(kl, sblk);
  }
}

__global__ void __launch_bounds__(kVT) variational_bwd(const float* __restrict__ params, const float* __restrict__ prior_loc,
                                                     float prior_scale, const float* __restrict__ eps,
                                                     const float* __restrict__ dw, const float* __restrict__ gkl, float gkl_value,
                                                     int n, int S, float* __restrict__ dparams,
                                                     float* __restrict__ dprior_loc) {
  const float g = gkl ? __ldg(gkl) : gkl_value;
  const float inv_sr2 = 1.0f / (prior_scale * prior_scale);
  for (int i = blockIdx.x * kVT + threadIdx.x; i < n; i += gridDim.x * kVT) {
    const float loc = __ldg(params + i);
    float sg;
    const float sigma = post_sigma(__ldg(params + n + i), &sg);
    float dl = 0.0f, ds = 0.0f;
    if (dw) {
      for (int s = 0; s < S; ++s) {
        const float v = __ldg(dw + (long long)s * n + i);
        dl += v;
        ds = fmaf(v, __ldg(eps + (long long)s * n + i), ds);
      }
    }
    const float d = loc - __ldg(prior_loc + i);
    dparams[i] += fmaf(g, d * inv_sr2, dl);
    dparams[n + i] += fmaf(g, sigma * inv_sr2 - 1.0f / sigma, ds) * 0.05f * sg;
    if (dprior_loc) dprior_loc[i] -= g * d * inv_sr2;
  }
}

}  // namespace

int launch_variational(bool bwd, const float* params, const float* prior_loc, float prior_scale, const float* eps,
                       const float* dw, const float* gkl, float gkl_value, int n, int S, float* w, double* kl, float* dparams,
                       float* dprior_loc, cudaStream_t st) {
  int blocks = (n + kVT - 1) / kVT;
  if (blocks > 64) blocks = 64;
  if (blocks < 1) blocks = 1;
  if (bwd) variational_bwd<<<blocks, kVT, 0, st>>>(params, prior_loc, prior_scale, eps, dw, gkl, gkl_value, n, S, dparams, dprior_loc);
  else variational_fwd<<<blocks, kVT, 0, st>>>(params, prior_loc, prior_scale, eps, n, S, w, kl);
  count_launch();
  return cuda_error(cudaGetLastError(), bwd ? "variational_bwd" : "variational_fwd");
}

}  // namespace nfn
