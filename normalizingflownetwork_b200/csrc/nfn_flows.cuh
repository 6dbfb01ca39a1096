// nfn_flows.cuh -- per-sample bijector arithmetic on register arrays (sm_100a).
//
// Each flow is a struct templated on the event dimension D and a math policy M with
//   fwd(th, z, ld)              : z <- f(z), ld.add(log|det J_f(z_in)|)
//   bwd(th, zin, G, cot, gth)   : reverse sweep.  G is d(cot*logp)/dz' on entry and
//                                 d(cot*logp)/dz on exit; gth receives d(cot*logp)/dth.
// Activations are recomputed in bwd from (th, zin); nothing but z_in is kept.
//
// The arithmetic restates the reference (paths into /root/reference):
//   planar : estimators/normalizing_flows/PlanarFlow.py:20-33 (slicing, w+1), :43-53
//            (_u_circ), :55-59 (_wzb), :68-72 (_forward), :74-80 (fldj)
//   radial : estimators/normalizing_flows/RadialFlow.py:20-34, :44-56, :58-70, :72-84
//            (L1 radius, alpha*beta coupling, no abs on the determinant)
//   affine : estimators/normalizing_flows/AffineFlow.py:4-10 (tfp Affine: shift, 1+scale)
//   base   : estimators/DistributionLayers.py:280-294 (+ MultivariateNormalDiag.log_prob)
// in the cancellation-free forms of SURVEY.md App. A.5; the reverse sweep is App. A.3.
#pragma once
#include "nfn_math.cuh"

namespace nfn {

enum : int { kPlanar = 0, kRadial = 1, kAffine = 2 };

__host__ __device__ constexpr int flow_param_size(int type, int d) {
  return type == kPlanar ? 2 * d + 1 : (type == kRadial ? d + 2 : 2 * d);
}

// ---------------------------------------------------------------- log-det accumulator
// Accurate: natural-log sum with log1p where the factor is 1+x.
// Fast    : log2 sum, one MUFU.LG2 per flow on that flow's combined determinant factor.
template <class M>
struct LogDetAcc {
  float v = 0.0f;
  NFN_DEVI void add_factor(float f) {  // f > 0 or |f| taken by caller
    if constexpr (M::kFast) v += M::lg2(f); else v += logf(f);
  }
  NFN_DEVI void add_one_plus(float x, int times) {  // factor (1+x)^times, accurate path only
    v += (float)times * log1pf(x);
  }
  NFN_DEVI float nat() const {
    if constexpr (M::kFast) return v * kLn2; else return v;
  }
};

// ---------------------------------------------------------------- planar
template <int D, class M>
struct PlanarFlow {
  static constexpr int N = 2 * D + 1;

  struct Pre {
    float w[D], uh[D];
    float b, c, rn, Q, wuh, sg_neg;
  };

  // constraint transform: u_hat = u + (m(wtu) - wtu) w / (|w|^2 + 1e-9)
  NFN_DEVI static void prepare(const float (&th)[N], Pre& p) {
    float wtu = 0.0f, n = 1e-9f;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      p.w[i] = th[D + i] + 1.0f;
      wtu = fmaf(p.w[i], th[i], wtu);
      n = fmaf(p.w[i], p.w[i], n);
    }
    p.b = th[2 * D];
    // softplus(+-wtu) and sigmoid(-wtu) from one e^{-|wtu|}
    const float e = M::exp(-fabsf(wtu));
    float L;
    if constexpr (M::kFast) L = M::lg2(1.0f + e) * kLn2; else L = log1pf(e);
    const float r = M::rcp(1.0f + e);
    p.sg_neg = (wtu >= 0.0f) ? e * r : r;               // 1 - sigmoid(wtu)
    const float sp_pos = fmaxf(wtu, 0.0f) + L;          // softplus(wtu)
    const float sp_neg = fmaxf(-wtu, 0.0f) + L;         // softplus(wtu) - wtu, no cancellation
    p.rn = M::rcp(n);
    p.c = (sp_neg - 0.99999f) * p.rn;                   // (m - wtu) / n,  m = -1 + sp + 1e-5
    // w . u_hat = m - c*1e-9 ;  Q = 1 + w . u_hat > 0
    p.wuh = (sp_pos - 0.99999f) - p.c * 1e-9f;
    p.Q = (sp_pos + 1e-5f) - p.c * 1e-9f;
#pragma unroll
    for (int i = 0; i < D; ++i) p.uh[i] = fmaf(p.c, p.w[i], th[i]);
  }

  NFN_DEVI static void fwd(const float (&th)[N], float (&z)[D], LogDetAcc<M>& ld) {
    Pre p;
    prepare(th, p);
    float a = p.b;
#pragma unroll
    for (int i = 0; i < D; ++i) a = fmaf(p.w[i], z[i], a);
    float tau, s;
    M::tanh_sech2(a, tau, s);
    const float det = fmaf(s, p.Q, tau * tau);           // 1 + s * (w . u_hat)
    if constexpr (M::kFast) {
      ld.add_factor(fabsf(det));
    } else {
      // |s*wuh| << 1 needs log1p (App. A.5); det = 1 + s*wuh exactly
      const float x = s * p.wuh;
      if (fabsf(x) < 0.25f) ld.add_one_plus(x, 1); else ld.add_factor(fabsf(det));
    }
#pragma unroll
    for (int i = 0; i < D; ++i) z[i] = fmaf(p.uh[i], tau, z[i]);
  }

  NFN_DEVI static void bwd(const float (&th)[N], const float (&zin)[D], float (&G)[D], float cot,
                           float (&gth)[N]) {
    Pre p;
    prepare(th, p);
    float a = p.b;
#pragma unroll
    for (int i = 0; i < D; ++i) a = fmaf(p.w[i], zin[i], a);
    float tau, s;
    M::tanh_sech2(a, tau, s);
    const float det = fmaf(s, p.Q, tau * tau);
    const float sD = s * M::div(cot, det);               // cot * s / D
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) q = fmaf(p.uh[i], G[i], q);
    const float g_a = fmaf(s, q, -2.0f * tau * p.wuh * sD);
    float guh[D];
    float pp = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      guh[i] = fmaf(tau, G[i], sD * p.w[i]);
      pp = fmaf(guh[i], p.w[i], pp);
    }
    const float g_wtu = -pp * p.sg_neg * p.rn;           // p (sigmoid(wtu) - 1) / n
    const float g_n2 = -2.0f * pp * p.c * p.rn;          // 2 * dL/dn
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const float gw = fmaf(zin[i], g_a, sD * p.uh[i]);
      gth[i] = fmaf(g_wtu, p.w[i], guh[i]);
      gth[D + i] = fmaf(g_n2, p.w[i], fmaf(g_wtu, th[i], fmaf(p.c, guh[i], gw)));
      G[i] = fmaf(p.w[i], g_a, G[i]);
    }
    gth[2 * D] = g_a;
  }
};

// ---------------------------------------------------------------- radial
template <int D, class M>
struct RadialFlow {
  static constexpr int N = D + 2;

  // fwd_save: as fwd, and replaces th[0], th[1] by the constrained alpha and beta + 1 so that
  // the reverse sweep (bwd_saved) needs no softplus: sigmoid(x) = 1 - exp(-softplus(x)).
  NFN_DEVI static void fwd_save(float (&th)[N], float (&z)[D], LogDetAcc<M>& ld) {
    const float alpha = M::softplus(fmaf(0.3f, th[0], -2.0f));
    const float spb = M::softplus(fmaf(0.1f, th[1], kC0));
    fwd_core(alpha, spb - 1.0f, th, z, ld);
    th[0] = alpha;
    th[1] = spb;
  }

  NFN_DEVI static void fwd(const float (&th)[N], float (&z)[D], LogDetAcc<M>& ld) {
    const float alpha = M::softplus(fmaf(0.3f, th[0], -2.0f));
    const float beta = M::softplus(fmaf(0.1f, th[1], kC0)) - 1.0f;
    fwd_core(alpha, beta, th, z, ld);
  }

  NFN_DEVI static void fwd_core(float alpha, float beta, const float (&th)[N], float (&z)[D],
                                LogDetAcc<M>& ld) {
    float delta[D];
    float r = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      delta[i] = z[i] - th[2 + i];
      r += fabsf(delta[i]);
    }
    const float h = M::rcp(alpha + r);
    const float abh = alpha * beta * h;
    const float ah = alpha * h;                          // 1 - h r
    const float x2 = abh * ah;                           // alpha^2 beta h^2
    if constexpr (M::kFast) {
      float f = 1.0f + x2;
      const float T1 = 1.0f + abh;
#pragma unroll
      for (int i = 0; i < D - 1; ++i) f *= T1;
      ld.add_factor(f);
    } else {
      ld.add_one_plus(abh, D - 1);
      ld.add_one_plus(x2, 1);
    }
#pragma unroll
    for (int i = 0; i < D; ++i) z[i] = fmaf(abh, delta[i], z[i]);
  }

  NFN_DEVI static void bwd(const float (&th)[N], const float (&zin)[D], float (&G)[D], float cot,
                           float (&gth)[N]) {
    float alpha, sga, spb, sgb;
    M::softplus_sigmoid(fmaf(0.3f, th[0], -2.0f), alpha, sga);
    M::softplus_sigmoid(fmaf(0.1f, th[1], kC0), spb, sgb);
    bwd_core(alpha, sga, spb, sgb, th, zin, G, cot, gth);
  }

  // th[0] = alpha, th[1] = beta + 1 as left by fwd_save
  NFN_DEVI static void bwd_saved(const float (&th)[N], const float (&zin)[D], float (&G)[D], float cot,
                                 float (&gth)[N]) {
    const float sga = M::one_minus_exp_neg(th[0]);
    const float sgb = M::one_minus_exp_neg(th[1]);
    bwd_core(th[0], sga, th[1], sgb, th, zin, G, cot, gth);
  }

  NFN_DEVI static void bwd_core(float alpha, float sga, float spb, float sgb, const float (&th)[N],
                                const float (&zin)[D], float (&G)[D], float cot, float (&gth)[N]) {
    const float beta = spb - 1.0f;
    float delta[D];
    float r = 0.0f, dG = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      delta[i] = zin[i] - th[2 + i];
      r += fabsf(delta[i]);
      dG = fmaf(delta[i], G[i], dG);
    }
    const float h = M::rcp(alpha + r);
    const float ab = alpha * beta;
    const float abh = ab * h;
    const float ah = alpha * h;
    // cot / T1 and cot / T2 from one reciprocal of T1 * T2 (both are > 0)
    const float T1 = 1.0f + abh, T2 = fmaf(abh, ah, 1.0f);
    const float r12 = cot * M::rcp(T1 * T2);
    const float rT1 = T2 * r12;
    const float rT2 = T1 * r12;
    const float k1 = (float)(D - 1) * rT1;
    // dL/dh, dL/dr, dL/d(alpha*beta) with T1 = 1+ab h, T2 = 1 + ab h (1 - h r)
    const float g_h = ab * (dG + k1 + (2.0f * ah - 1.0f) * rT2);
    const float h2 = h * h;
    const float g_r = -h2 * fmaf(ab, rT2, g_h);
    const float g_ab = h * (dG + k1 + ah * rT2);
    gth[0] = fmaf(beta, g_ab, -g_h * h2) * (0.3f * sga);
    gth[1] = alpha * g_ab * (0.1f * sgb);
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const float sgn = (delta[i] > 0.0f ? 1.0f : 0.0f) - (delta[i] < 0.0f ? 1.0f : 0.0f);
      const float v = fmaf(abh, G[i], sgn * g_r);        // sign(0) = 0, TF's abs gradient
      gth[2 + i] = -v;
      G[i] += v;
    }
  }
};

// ---------------------------------------------------------------- affine
template <int D, class M>
struct AffineFlow {
  static constexpr int N = 2 * D;

  NFN_DEVI static void fwd(const float (&th)[N], float (&z)[D], LogDetAcc<M>& ld) {
    float f = 1.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const float s = 1.0f + th[D + i];
      if constexpr (M::kFast) f *= fabsf(s); else ld.add_factor(fabsf(s));
      z[i] = fmaf(s, z[i], th[i]);
    }
    if constexpr (M::kFast) ld.add_factor(f);
  }

  NFN_DEVI static void bwd(const float (&th)[N], const float (&zin)[D], float (&G)[D], float cot,
                           float (&gth)[N]) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const float s = 1.0f + th[D + i];
      gth[i] = G[i];
      gth[D + i] = fmaf(zin[i], G[i], M::div(cot, s));
      G[i] *= s;
    }
  }
};

// ---------------------------------------------------------------- base distribution
// MultivariateNormalDiag(loc, 1e-3 + softplus(c0 + 0.1 raw)) or the standard normal.
template <int D, bool TRAINABLE, class M>
struct BaseDist {
  static constexpr int N = TRAINABLE ? 2 * D : 0;
  static constexpr int NA = N > 0 ? N : 1;

  // returns log N(z) (without the log-det term)
  NFN_DEVI static float log_prob(const float (&th)[NA], const float (&z)[D]) {
    float quad = 0.0f;
    if constexpr (TRAINABLE) {
      LogDetAcc<M> ls;
      float f = 1.0f;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const float sig = 1e-3f + M::softplus(fmaf(0.1f, th[D + i], kC0));
        const float e = M::div(z[i] - th[i], sig);
        quad = fmaf(e, e, quad);
        if constexpr (M::kFast) f *= sig; else ls.add_factor(sig);
      }
      if constexpr (M::kFast) ls.add_factor(f);
      return fmaf(-0.5f, quad, -ls.nat()) - (float)D * kHalfLog2Pi;
    } else {
#pragma unroll
      for (int i = 0; i < D; ++i) quad = fmaf(z[i], z[i], quad);
      return -0.5f * quad - (float)D * kHalfLog2Pi;
    }
  }

  // as log_prob, and replaces th[D+i] by sigma_i for bwd_saved
  NFN_DEVI static float log_prob_save(float (&th)[NA], const float (&z)[D]) {
    if constexpr (TRAINABLE) {
      float quad = 0.0f;
      LogDetAcc<M> ls;
      float f = 1.0f;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const float sig = 1e-3f + M::softplus(fmaf(0.1f, th[D + i], kC0));
        const float e = M::div(z[i] - th[i], sig);
        quad = fmaf(e, e, quad);
        if constexpr (M::kFast) f *= sig; else ls.add_factor(sig);
        th[D + i] = sig;
      }
      if constexpr (M::kFast) ls.add_factor(f);
      return fmaf(-0.5f, quad, -ls.nat()) - (float)D * kHalfLog2Pi;
    } else {
      return log_prob(th, z);
    }
  }

  // th[D+i] = sigma_i as left by log_prob_save; d sigma / d raw = 0.1 * (1 - exp(-(sigma - 1e-3)))
  NFN_DEVI static void bwd_saved(const float (&th)[NA], const float (&z)[D], float cot, float (&G)[D],
                                 float (&gth)[NA]) {
    if constexpr (TRAINABLE) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const float sg = M::one_minus_exp_neg(th[D + i] - 1e-3f);
        const float rs = M::rcp(th[D + i]);
        const float e = (z[i] - th[i]) * rs;
        const float ce = cot * e * rs;
        G[i] = -ce;
        gth[i] = ce;
        gth[D + i] = cot * fmaf(e, e, -1.0f) * rs * (0.1f * sg);
      }
    } else {
      bwd(th, z, cot, G, gth);
    }
  }

  // G <- d(cot*logp)/dz_K ; gth <- d(cot*logp)/d(mu, sigma_raw)
  NFN_DEVI static void bwd(const float (&th)[NA], const float (&z)[D], float cot, float (&G)[D],
                           float (&gth)[NA]) {
    if constexpr (TRAINABLE) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        float sp, sg;
        M::softplus_sigmoid(fmaf(0.1f, th[D + i], kC0), sp, sg);
        const float rs = M::rcp(1e-3f + sp);
        const float e = (z[i] - th[i]) * rs;
        const float ce = cot * e * rs;
        G[i] = -ce;
        gth[i] = ce;
        gth[D + i] = cot * fmaf(e, e, -1.0f) * rs * (0.1f * sg);
      }
    } else {
#pragma unroll
      for (int i = 0; i < D; ++i) G[i] = -cot * z[i];
    }
  }
};

}  // namespace nfn
