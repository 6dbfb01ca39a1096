// nfn_mixture_row.cuh -- per-row arithmetic of the MDN head (sm_100a), shared by the streaming kernel
// (nfn_mixture.cu: parameter rows staged from HBM) and the fused Dense(P)+MDN kernel (nfn_dense_chain.cuh:
// parameter rows formed in shared memory by the emitting layer's GEMM and never written to HBM).
//
//   tfd.Mixture of K diagonal Gaussians parameterised per sample (reference
//   estimators/DistributionLayers.py:196-212), row layout
//   [ (mu_k(d), sigma_raw_k(d))_{k<K} | logits(K) ], sigma = softplus(0.05 raw + c0).
//
// Forward is an online logsumexp (one EX2 per component); the reverse sweep reuses sigma (written over
// sigma_raw by the forward pass) so softplus is not recomputed.  All log-densities are carried in log2
// units (one EX2 / LG2 per use, no rescaling multiply).
#pragma once
#include "nfn_math.cuh"

namespace nfn {

#ifndef NFN_NEG_INF
#define NFN_NEG_INF (-__int_as_float(0x7f800000))
#endif

// per-thread load/store of N consecutive floats at a runtime offset whose alignment
// (in floats) is at least A (compile time)
template <int N, int A>
NFN_DEVI void ld_vec(const float* p, float (&v)[N]) {
  if constexpr (A >= 4 && N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N; i += 4) {
      const float4 x = *reinterpret_cast<const float4*>(p + i);
      v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
    }
  } else if constexpr (A >= 2 && N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      const float2 x = *reinterpret_cast<const float2*>(p + i);
      v[i] = x.x; v[i + 1] = x.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = p[i];
  }
}
template <int N, int A>
NFN_DEVI void st_vec(float* p, const float (&v)[N]) {
  if constexpr (A >= 4 && N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  } else if constexpr (A >= 2 && N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N; i += 2) *reinterpret_cast<float2*>(p + i) = make_float2(v[i], v[i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = v[i];
  }
}

// online logsumexp update with one exponential
template <class M>
NFN_DEVI void lse_push(float x, float& m, float& s) {
  const float e = M::exp(-fabsf(x - m));
  s = (x > m) ? fmaf(s, e, 1.0f) : s + e;
  m = fmaxf(m, x);
}
// the same in base 2 (x, m in log2 units): no multiply in front of the EX2
template <class M>
NFN_DEVI void lse2_push(float x, float& m, float& s) {
  const float e = M::ex2(-fabsf(x - m));
  s = (x > m) ? fmaf(s, e, 1.0f) : s + e;
  m = fmaxf(m, x);
}

// One MDN row: returns log p(y | row) (natural log); with BWD also overwrites the row with
// cot * d log p / d row and subtracts cot * d log p / d mu from dy (= the event gradient).
// V4: rows are 16-byte aligned (P % 4 == 0) so the (mu, sigma_raw) block of a component, 2*D floats
// at offset k*2*D, can be read with the widest aligned vectors.  LG: logits are read / their gradients
// written in groups of LG (4 when V4 and K % 4 == 0: a scalar LDS at row stride S = 4*odd is 4-way
// bank conflicted, a 128-bit one is not).
template <int D, bool V4, int LG, bool BWD, class M>
NFN_DEVI float mdn_row(float* row, const int K, const float (&y)[D], const float cot, float (&dy)[D]) {
  constexpr int A = V4 ? ((2 * D) % 4 == 0 ? 4 : ((2 * D) % 2 == 0 ? 2 : 1)) : 1;
  constexpr float kHalfLog2e = 0.5f * kLog2e;
  const int LO = 2 * K * D;  // logits offset
  // log-softmax normaliser of the logits (log2 units)
  float lm = NFN_NEG_INF, ls = 0.0f;
  for (int k0 = 0; k0 < K; k0 += LG) {
    float lg[LG];
    ld_vec<LG, LG>(row + LO + k0, lg);
#pragma unroll
    for (int j = 0; j < LG; ++j) lse2_push<M>(lg[j] * kLog2e, lm, ls);
  }
  const float lse2 = lm + M::lg2(ls);
  // components, online logsumexp
  float m = NFN_NEG_INF, s = 0.0f;
  for (int k0 = 0; k0 < K; k0 += LG) {
    float lg[LG];
    ld_vec<LG, LG>(row + LO + k0, lg);
#pragma unroll
    for (int j = 0; j < LG; ++j) {
      float* blk = row + (k0 + j) * 2 * D;
      float th[2 * D];
      ld_vec<2 * D, A>(blk, th);
      float quad = 0.0f, prod = 1.0f;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const float sig = M::softplus(fmaf(0.05f, th[D + i], kC0));
        const float e = M::div(y[i] - th[i], sig);
        quad = fmaf(e, e, quad);
        prod *= sig;
        if constexpr (BWD) th[D + i] = sig;
      }
      if constexpr (BWD) st_vec<2 * D, A>(blk, th);  // keep sigma for the reverse sweep
      const float lp2 = fmaf(lg[j], kLog2e, -kHalfLog2e * quad) - M::lg2(prod);
      lse2_push<M>(lp2, m, s);
    }
  }
  const float top2 = m + M::lg2(s);  // log2 sum_k exp(logit_k + log N_k + d/2 log 2pi)
  const float logp = (top2 - lse2) * kLn2 - (float)D * kHalfLog2Pi;
  if constexpr (BWD) {
    for (int k0 = 0; k0 < K; k0 += LG) {
      float lg[LG];
      ld_vec<LG, LG>(row + LO + k0, lg);
#pragma unroll
      for (int j = 0; j < LG; ++j) {
        float* blk = row + (k0 + j) * 2 * D;
        float th[2 * D];
        ld_vec<2 * D, A>(blk, th);                       // (mu, sigma)
        float e[D], rs[D];
        float quad = 0.0f, prod = 1.0f;
#pragma unroll
        for (int i = 0; i < D; ++i) {
          rs[i] = M::rcp(th[D + i]);
          e[i] = (y[i] - th[i]) * rs[i];
          quad = fmaf(e[i], e[i], quad);
          prod *= th[D + i];
        }
        const float l2 = lg[j] * kLog2e;
        const float lp2 = fmaf(-kHalfLog2e, quad, l2) - M::lg2(prod);
        const float crho = cot * M::ex2(lp2 - top2);     // cot * responsibility
        lg[j] = fmaf(-cot, M::ex2(l2 - lse2), crho);     // cot * (rho_k - softmax_k)
#pragma unroll
        for (int i = 0; i < D; ++i) {
          const float gm = crho * e[i] * rs[i];
          // d sigma / d raw = 0.05 * sigmoid(x), and sigmoid(x) = 1 - exp(-softplus(x))
          const float dsig = 0.05f * M::one_minus_exp_neg(th[D + i]);
          th[D + i] = crho * fmaf(e[i], e[i], -1.0f) * rs[i] * dsig;
          th[i] = gm;
          dy[i] -= gm;
        }
        st_vec<2 * D, A>(blk, th);
      }
      st_vec<LG, LG>(row + LO + k0, lg);
    }
  }
  return logp;
}

// One KMN row (tfd.MixtureSameFamily over K fixed centres with shared isotropic bandwidths, reference
// estimators/DistributionLayers.py:118-133): row = logits(K).  s_loc [K][D], s_coef [K] = -0.5 log2e / s^2,
// s_lnorm [K] = -D log2|s| are per-kernel constants in shared memory.  Called by ALL lanes of a warp (`valid` masks
// rows past the end): the bandwidth gradient is reduced over the warp's rows with shuffles and added to my_dsc [K]
// (this warp's partial sums, * 1/s applied when they are flushed; nullptr: not wanted).  Returns log p (natural log);
// with BWD overwrites the row with cot * d log p / d logits and accumulates the event gradient into dy.
template <int D, int LG, bool BWD, class M>
NFN_DEVI float kmn_row(float* row, const int K, const float (&y)[D], const float cot, const bool valid,
                       const float* s_loc, const float* s_coef, const float* s_lnorm, float* my_dsc, float (&dy)[D]) {
  float lse2 = 0.0f, top2 = 0.0f, logp = 0.0f;
  if (valid) {
    float lm = NFN_NEG_INF, ls = 0.0f, m = NFN_NEG_INF, s = 0.0f;
    for (int k0 = 0; k0 < K; k0 += LG) {
      float lg[LG];
      ld_vec<LG, LG>(row + k0, lg);
#pragma unroll
      for (int j = 0; j < LG; ++j) {
        const int k = k0 + j;
        const float l2 = lg[j] * kLog2e;
        lse2_push<M>(l2, lm, ls);
        float q = 0.0f;
#pragma unroll
        for (int i = 0; i < D; ++i) {
          const float dlt = y[i] - s_loc[k * D + i];
          q = fmaf(dlt, dlt, q);
        }
        lse2_push<M>(l2 + fmaf(s_coef[k], q, s_lnorm[k]), m, s);
      }
    }
    lse2 = lm + M::lg2(ls);
    top2 = m + M::lg2(s);
    logp = (top2 - lse2) * kLn2 - (float)D * kHalfLog2Pi;
  }
  if constexpr (BWD) {
    // Kernels in blocks of 32: every lane first forms its row's 32 bandwidth-gradient terms, then ONE butterfly
    // reduce-scatter (16 + 8 + 4 + 2 + 1 = 31 shuffles) leaves lane L with the warp's sum for kernel kb + L -- instead
    // of a 5-shuffle all-reduce per kernel (160 per block), which was most of this pass.  All lanes stay in the loops.
    const unsigned lane = threadIdx.x & 31u;
    for (int kb = 0; kb < K; kb += 32) {
      float wv[32];
#pragma unroll
      for (int gq = 0; gq < 32 / LG; ++gq) {
        const int k0 = kb + gq * LG;
        float lg[LG];
        const bool live = valid && k0 < K;     // (K % LG == 0: a group is inside or outside as a whole)
        if (live) {
          ld_vec<LG, LG>(row + k0, lg);
        } else {
#pragma unroll
          for (int j = 0; j < LG; ++j) lg[j] = 0.0f;
        }
#pragma unroll
        for (int j = 0; j < LG; ++j) {
          const int k = k0 + j;
          float wsc = 0.0f;
          if (live) {
            float q = 0.0f, dl[D];
#pragma unroll
            for (int i = 0; i < D; ++i) {
              dl[i] = y[i] - s_loc[k * D + i];
              q = fmaf(dl[i], dl[i], q);
            }
            const float l2 = lg[j] * kLog2e;
            const float crho = cot * M::ex2(l2 + fmaf(s_coef[k], q, s_lnorm[k]) - top2);
            lg[j] = fmaf(-cot, M::ex2(l2 - lse2), crho);
            const float c2 = -2.0f * kLn2 * s_coef[k];        // 1 / s^2
#pragma unroll
            for (int i = 0; i < D; ++i) dy[i] -= crho * dl[i] * c2;
            wsc = crho * fmaf(q, c2, -(float)D);               // * 1/s applied when the sums are flushed
          }
          wv[gq * LG + j] = wsc;
        }
        if (live) st_vec<LG, LG>(row + k0, lg);
      }
      if (my_dsc) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
          const bool up = (lane & off) != 0;    // this lane keeps the upper half of what it still holds
#pragma unroll
          for (int i = 0; i < off; ++i) {
            const float send = up ? wv[i] : wv[i + off];
            const float keep = up ? wv[i + off] : wv[i];
            wv[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        if (kb + (int)lane < K) my_dsc[kb + lane] += wv[0];
      }
    }
  }
  return logp;
}

}  // namespace nfn
