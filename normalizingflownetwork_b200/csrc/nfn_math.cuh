// nfn_math.cuh -- scalar fp32 math policies for the flow / mixture kernels (sm_100a).
//
// Two policies with the same interface:
//   MathAccurate : CUDA libm (expf/log1pf/logf/tanhf) and IEEE division.  Used by the
//                  single-bijector debug entry point and selectable for the chain.
//   MathFast     : one MUFU.EX2 / MUFU.LG2 / MUFU.RCP per transcendental, sharing e^{-|x|}
//                  between softplus and its derivative (SURVEY.md §7 "hard parts" 2).
//                  Never uses tanh.approx (2^-11).  Every form is cancellation-free
//                  (SURVEY.md App. A.5), so fp32 stays inside the 1e-5 log-prob bar.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#endif

namespace nfn {

#define NFN_DEVI __device__ __forceinline__

constexpr float kC0 = 0.541324854612918f;        // log(e - 1) = tf.math.log(tf.math.expm1(1.0))
constexpr float kHalfLog2Pi = 0.918938533204673f; // 0.5 * log(2 pi)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct MathAccurate {
  static constexpr bool kFast = false;
  NFN_DEVI static float rcp(float x) { return 1.0f / x; }
  NFN_DEVI static float div(float a, float b) { return a / b; }
  NFN_DEVI static float log(float x) { return logf(x); }
  NFN_DEVI static float exp(float x) { return expf(x); }
  NFN_DEVI static float ex2(float x) { return exp2f(x); }
  NFN_DEVI static float lg2(float x) { return log2f(x); }
  // softplus(x) and sigmoid(x) from one exponential
  NFN_DEVI static void softplus_sigmoid(float x, float& sp, float& sg) {
    const float e = expf(-fabsf(x));
    sp = fmaxf(x, 0.0f) + log1pf(e);
    const float r = 1.0f / (1.0f + e);
    sg = (x >= 0.0f) ? r : e * r;
  }
  NFN_DEVI static float softplus(float x) {
    return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x)));
  }
  // 1 - e^{-x}: sigmoid(v) when x = softplus(v)
  NFN_DEVI static float one_minus_exp_neg(float x) { return -expm1f(-x); }
  // tanh(a) and sech^2(a) = 1 - tanh^2(a), the latter with full relative accuracy
  NFN_DEVI static void tanh_sech2(float a, float& th, float& s2) {
    th = tanhf(a);
    const float e = expf(-2.0f * fabsf(a));
    const float r = 1.0f / (1.0f + e);
    s2 = 4.0f * e * r * r;
  }
};

struct MathFast {
  static constexpr bool kFast = true;
  NFN_DEVI static float ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
  }
  NFN_DEVI static float lg2(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
  }
  NFN_DEVI static float rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
  }
  NFN_DEVI static float div(float a, float b) { return a * rcp(b); }
  NFN_DEVI static float log(float x) { return lg2(x) * kLn2; }
  NFN_DEVI static float exp(float x) { return ex2(x * kLog2e); }
  NFN_DEVI static void softplus_sigmoid(float x, float& sp, float& sg) {
    const float e = ex2(-fabsf(x) * kLog2e);
    const float ope = 1.0f + e;
    sp = fmaf(lg2(ope), kLn2, fmaxf(x, 0.0f));
    const float r = rcp(ope);
    sg = (x >= 0.0f) ? r : e * r;
  }
  NFN_DEVI static float softplus(float x) {
    const float e = ex2(-fabsf(x) * kLog2e);
    return fmaf(lg2(1.0f + e), kLn2, fmaxf(x, 0.0f));
  }
  NFN_DEVI static float one_minus_exp_neg(float x) { return 1.0f - ex2(-kLog2e * x); }
  NFN_DEVI static void tanh_sech2(float a, float& th, float& s2) {
    const float e = ex2(-2.0f * kLog2e * fabsf(a));
    const float r = rcp(1.0f + e);
    const float m = (1.0f - e) * r;
    th = copysignf(m, a);
    s2 = 4.0f * e * r * r;
  }
};

}  // namespace nfn
