// nfn_mlp.cu -- the hidden layers of the conditioning network (sm_100a).
//
// SURVEY.md §8(f) rank 1 continued: once the emitting Dense(P) layer and the flow chain are one kernel
// (nfn_dense_tc5.cuh), an estimator-level step at a large batch is dominated by the 16..64-wide hidden
// layers around it -- `Dense(units, activation)` of MaximumLikelihoodNNEstimator.py:37-44 -- which cuBLAS
// serves with skinny SIMT GEMMs plus separate bias / activation / reduction kernels (measured at
// B = 2^20: ~900 us of a 960 us train step, next to a 73 us fused head kernel).  Here one layer is one
// kernel each way, row per thread, weights in shared memory (every weight read is a warp broadcast):
//
//   forward   out = act(x W^T + b)                                 reads 4K, writes 4N bytes per row
//   backward  dpre = dout * act'(out);  dx = dpre W;  dW += dpre^T x;  db += 1^T dpre
//
// fp32 throughout (the layers are tiny; exactness matters more than tensor-core throughput here).
// `weight` is torch's / Keras-transposed layout [N][K] (out_features x in_features).
#include <cuda_runtime.h>

#include <cstdint>

#include "nfn_common.h"

namespace nfn {
namespace {

enum Act { kLinear = 0, kTanh = 1, kRelu = 2, kSigmoid = 3, kElu = 4 };

template <int ACT>
__device__ __forceinline__ float act_fwd(float a) {
  if constexpr (ACT == kTanh) return tanhf(a);
  else if constexpr (ACT == kRelu) return a > 0.0f ? a : 0.0f;
  else if constexpr (ACT == kSigmoid) return 1.0f / (1.0f + expf(-a));
  else if constexpr (ACT == kElu) return a > 0.0f ? a : expm1f(a);
  else return a;
}
// derivative written in terms of the OUTPUT o = act(a), so the backward pass needs no pre-activations
template <int ACT>
__device__ __forceinline__ float act_bwd(float o) {
  if constexpr (ACT == kTanh) return 1.0f - o * o;
  else if constexpr (ACT == kRelu) return o > 0.0f ? 1.0f : 0.0f;
  else if constexpr (ACT == kSigmoid) return o * (1.0f - o);
  else if constexpr (ACT == kElu) return o > 0.0f ? 1.0f : o + 1.0f;
  else return 1.0f;
}

constexpr int kT = 128;  // rows per tile == threads per CTA

// Cooperative, fully coalesced copy of a tile of kT rows x C floats between global [B][C] and shared [kT][S]:
// consecutive threads move consecutive 16-byte chunks (row-per-thread global access at a 64-byte row stride
// touches 32 sectors per warp instruction and was the bottleneck of the first version of these kernels).
// Rows past B are zero-filled on the way in and skipped on the way out.
template <bool TO_SMEM>
__device__ __forceinline__ void tile_copy(float* s, int S, float* g, long long row0, long long B, int C) {
  if (C % 4 == 0) {
    const int C4 = C / 4;
    // (row, chunk) advance incrementally: no division in the loop
    int r = threadIdx.x / C4, c = threadIdx.x - r * C4;
    const int dr = kT / C4, dc = kT - dr * C4;
    for (int q = threadIdx.x; q < kT * C4; q += kT) {
      float4* sp = reinterpret_cast<float4*>(s + r * S + 4 * c);
      if (row0 + r < B) {
        float4* gp = reinterpret_cast<float4*>(g + (row0 + r) * C + 4 * c);
        // in: asynchronous 16-byte global -> shared copies, all of a thread's chunks in flight at once (a
        // register round trip serialised one DRAM latency per chunk: 54 % of the stall samples); out: plain store
        if (TO_SMEM) cp_async16(smem_u32(sp), gp); else *gp = *sp;
      } else if (TO_SMEM) {
        *sp = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      r += dr;
      c += dc;
      if (c >= C4) { c -= C4; ++r; }
    }
  } else {
    for (int e = threadIdx.x; e < kT * C; e += kT) {
      const int r = e / C, c = e - r * C;
      if (row0 + r < B) {
        if (TO_SMEM) cp_async4(smem_u32(s + r * S + c), g + (row0 + r) * C + c); else g[(row0 + r) * C + c] = s[r * S + c];
      } else if (TO_SMEM) {
        s[r * S + c] = 0.0f;
      }
    }
  }
}

__host__ __device__ constexpr int pad4(int n) { return (n + 3) / 4 * 4; }

// ---------------------------------------------------------------- forward
// KC > 0 fixes the input width at compile time (the common 16 -> 16 layer: a quarter of the instructions of the
// runtime-K version were address arithmetic)
template <int N, int ACT, int KC>
__global__ void __launch_bounds__(kT) dense_act_fwd(const float* __restrict__ x, const float* __restrict__ xmean,
                                                   const float* __restrict__ xstd, const float* __restrict__ weight,
                                                   const float* __restrict__ bias, float* __restrict__ out,
                                                   long long B, int K_rt) {
  const int K = KC > 0 ? KC : K_rt;
  extern __shared__ __align__(16) float smem[];
  const int SX = pad4(K) + 4;
  constexpr int SO = N + 4;
  float* sW = smem;                    // [K][N]  (transposed while loading)
  float* sb = sW + pad4(K * N);        // [N]
  float* sX = sb + N;                  // [kT][SX]
  float* sO = sX + kT * SX;            // [kT][SO]
  for (int i = threadIdx.x; i < K * N; i += kT) {
    const int n = i / K, k = i % K;    // coalesced read of weight[n][k]
    sW[k * N + n] = __ldg(weight + i);
  }
  for (int i = threadIdx.x; i < N; i += kT) sb[i] = __ldg(bias + i);
  const long long ntiles = (B + kT - 1) / kT;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * kT;
    tile_copy<true>(sX, SX, const_cast<float*>(x), row0, B, K);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();   // (also covers the weights on the first tile)
    float acc[N];
#pragma unroll
    for (int n = 0; n < N; ++n) acc[n] = sb[n];
    const float* xr = sX + threadIdx.x * SX;
    for (int k = 0; k < K; ++k) {
      float a = xr[k];
      // input normalisation of the estimators' first layer, fused: (x - x_mean) / (x_std + 1e-8)
      // (reference estimators/MaximumLikelihoodNNEstimator.py:40)
      if (xmean) a = __fdiv_rn(a - __ldg(xmean + k), __ldg(xstd + k) + 1e-8f);
      const float4* w = reinterpret_cast<const float4*>(sW + k * N);
#pragma unroll
      for (int c = 0; c < N / 4; ++c) {
        const float4 wv = w[c];
        acc[4 * c] = fmaf(a, wv.x, acc[4 * c]);
        acc[4 * c + 1] = fmaf(a, wv.y, acc[4 * c + 1]);
        acc[4 * c + 2] = fmaf(a, wv.z, acc[4 * c + 2]);
        acc[4 * c + 3] = fmaf(a, wv.w, acc[4 * c + 3]);
      }
    }
    float4* o = reinterpret_cast<float4*>(sO + threadIdx.x * SO);
#pragma unroll
    for (int c = 0; c < N / 4; ++c)
      o[c] = make_float4(act_fwd<ACT>(acc[4 * c]), act_fwd<ACT>(acc[4 * c + 1]), act_fwd<ACT>(acc[4 * c + 2]),
                         act_fwd<ACT>(acc[4 * c + 3]));
    __syncthreads();
    tile_copy<false>(sO, SO, out, row0, B, N);
    // the next iteration's sX fill cannot overtake this tile's reads of sX (they ended at the barrier above);
    // its sO writes come after its own first barrier, i.e. after every thread has left this copy
  }
}

// ---------------------------------------------------------------- backward
// Per tile of kT rows: x, dout and out arrive in shared memory with coalesced copies; every thread forms dpre
// for its row (written back over dout) and its dx row (written over out, stored with a coalesced copy).  The
// weight gradient dW[n][k] = sum_r dpre[r][n] x[r][k] is register-tiled: the K x N entries are cut into
// 4 x 4 blocks; a thread owns PER blocks (1 or 2) and one slice of the tile's rows, and per row reads one
// float4 of x and one float4 of dpre for 16 FMAs.  Accumulators live in registers for the whole kernel; the
// slices meet in shared memory at the end: one global atomic per entry per CTA.
template <int N, int ACT, int PER, int KC>
__global__ void __launch_bounds__(kT, 4) dense_act_bwd(const float* __restrict__ x, const float* __restrict__ out,
                                                      const float* __restrict__ dout, const float* __restrict__ weight,
                                                      float* __restrict__ dx, float* __restrict__ dW,
                                                      float* __restrict__ db, long long B, int K_rt) {
  const int K = KC > 0 ? KC : K_rt;
  extern __shared__ __align__(16) float smem[];
  const int K4 = (K + 3) / 4;
  constexpr int N4 = N / 4;
  const int SX = 4 * K4 + 4;                       // float4-aligned rows; the +4 skews the banks between rows
  constexpr int SD = N + 4;
  const int SO = (N > 4 * K4 ? N : 4 * K4) + 4;    // holds the out row, then the dx row
  float* sW = smem;                                // [K][N]
  float* sX = sW + pad4(K * N);                    // [kT][SX]   (columns K .. 4 K4 - 1 are zero)
  float* sD = sX + kT * SX;                        // [kT][SD]   dout, then dpre
  float* sO = sD + kT * SD;                        // [kT][SO]   out, then dx
  for (int i = threadIdx.x; i < K * N; i += kT) {
    const int n = i / K, k = i % K;
    sW[k * N + n] = __ldg(weight + i);
  }
  if (K % 4 != 0) {                                // the pad columns of x stay zero for the whole kernel
    for (int e = threadIdx.x; e < kT * SX; e += kT) sX[e] = 0.0f;
  }
  // ownership of the 4 x 4 blocks
  const int G = K4 * N4;                           // blocks
  const int Gt = (G + PER - 1) / PER;              // threads per row slice
  const int S = kT / Gt > 0 ? kT / Gt : 1;         // row slices
  const int R = (kT + S - 1) / S;                  // rows per slice
  const int slice = threadIdx.x / Gt, member = threadIdx.x % Gt;
  const bool active = slice < S;
  const int r_lo = slice * R, r_hi = (r_lo + R < kT) ? r_lo + R : kT;
  float gw[PER][16], gb[N];
#pragma unroll
  for (int j = 0; j < PER; ++j)
#pragma unroll
    for (int e = 0; e < 16; ++e) gw[j][e] = 0.0f;
#pragma unroll
  for (int n = 0; n < N; ++n) gb[n] = 0.0f;
  __syncthreads();
  const long long ntiles = (B + kT - 1) / kT;
  // (a second staging set with the next tile's copies in flight was tried: 148 vs 121 us -- the kernel is
  // bound by issue / shared-memory latency with 16 warps per SM, not by the copy latency)
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * kT;
    tile_copy<true>(sX, SX, const_cast<float*>(x), row0, B, K);
    tile_copy<true>(sD, SD, const_cast<float*>(dout), row0, B, N);
    tile_copy<true>(sO, SO, const_cast<float*>(out), row0, B, N);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    {
      float dpre[N];
      float4* myd = reinterpret_cast<float4*>(sD + threadIdx.x * SD);
      float4* myo = reinterpret_cast<float4*>(sO + threadIdx.x * SO);
#pragma unroll
      for (int c = 0; c < N4; ++c) {
        const float4 g = myd[c], o = myo[c];       // rows past B are zero: dpre = 0
        dpre[4 * c] = g.x * act_bwd<ACT>(o.x);
        dpre[4 * c + 1] = g.y * act_bwd<ACT>(o.y);
        dpre[4 * c + 2] = g.z * act_bwd<ACT>(o.z);
        dpre[4 * c + 3] = g.w * act_bwd<ACT>(o.w);
        myd[c] = make_float4(dpre[4 * c], dpre[4 * c + 1], dpre[4 * c + 2], dpre[4 * c + 3]);
      }
#pragma unroll
      for (int n = 0; n < N; ++n) gb[n] += dpre[n];
      if (dx != nullptr) {
        float* mydx = sO + threadIdx.x * SO;
        for (int k = 0; k < K; ++k) {
          const float4* w = reinterpret_cast<const float4*>(sW + k * N);
          float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
          for (int c = 0; c < N4; ++c) {
            const float4 wv = w[c];
            s0 = fmaf(dpre[4 * c], wv.x, s0);
            s1 = fmaf(dpre[4 * c + 1], wv.y, s1);
            s0 = fmaf(dpre[4 * c + 2], wv.z, s0);
            s1 = fmaf(dpre[4 * c + 3], wv.w, s1);
          }
          mydx[k] = s0 + s1;
        }
      }
    }
    __syncthreads();
    if (dx != nullptr) tile_copy<false>(sO, SO, dx, row0, B, K);
    if (active) {
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int blk = member * PER + j;
        if (blk < G) {
          const int k4 = blk / N4, n4 = blk % N4;
          const float* px = sX + 4 * k4;
          const float* pd = sD + 4 * n4;
#pragma unroll 4
          for (int rr = r_lo; rr < r_hi; ++rr) {
            const float4 xv = *reinterpret_cast<const float4*>(px + rr * SX);
            const float4 dv = *reinterpret_cast<const float4*>(pd + rr * SD);
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
            const float ds[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
              for (int b = 0; b < 4; ++b) gw[j][4 * a + b] = fmaf(xs[a], ds[b], gw[j][4 * a + b]);
          }
        }
      }
    }
    __syncthreads();
  }
  // the row slices' partial sums meet in shared memory (the staging tiles are free now), so that every
  // weight-gradient entry costs ONE global atomic per CTA
  float* sAcc = sX;                                // [S][G * 16] <= kT * PER * 16 floats, inside sX | sD | sO
  if (active) {
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int blk = member * PER + j;
      if (blk < G) {
#pragma unroll
        for (int e = 0; e < 16; ++e) sAcc[(slice * G + blk) * 16 + e] = gw[j][e];
      }
    }
  }
  __syncthreads();
  for (int o = threadIdx.x; o < G * 16; o += kT) {
    float v = 0.0f;
    for (int sl = 0; sl < S; ++sl) v += sAcc[sl * G * 16 + o];
    const int blk = o / 16, e = o % 16;
    const int k = 4 * (blk / N4) + e / 4, n = 4 * (blk % N4) + e % 4;
    if (k < K) atomicAdd(dW + n * K + k, v);
  }
  // bias gradient: warp shuffles, then the CTA's warps meet in shared memory: one atomic per column per CTA
  __syncthreads();   // sAcc has been read
  float* sRed = sX;  // [kT / 32][N]
#pragma unroll
  for (int n = 0; n < N; ++n) {
    float v = gb[n];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if ((threadIdx.x & 31) == 0) sRed[(threadIdx.x >> 5) * N + n] = v;
  }
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += kT) {
    float v = 0.0f;
#pragma unroll
    for (int wi = 0; wi < kT / 32; ++wi) v += sRed[wi * N + n];
    atomicAdd(db + n, v);
  }
}

// ---------------------------------------------------------------- backward on the tensor cores (hidden layers)
// The 16/32-wide hidden layers' backward is two small GEMMs per row tile -- dx = dpre W (contraction over the
// units) and dW += dpre^T x (contraction over the tile's ROWS) -- which the kernel above does with 2 x K x N scalar
// FMAs per row (46 M warp instructions at B = 2^20, 16 -> 16: 3x off its memory floor, issue-bound).  Here both run
// as warp-level mma.sync.m16n8k8 TF32 with the error-compensated split of nfn_dense_chain.cuh (3 products, fp32-grade):
// every warp owns 32 rows of the tile, the W fragments of dx are split once per kernel and live in registers, the
// dW accumulators stay in registers over ALL tiles of the CTA (N K / 32 per thread) and meet in shared memory once.
template <int N, int K, int ACT>
__global__ void __launch_bounds__(kT, 4) dense_act_bwd_mma(const float* __restrict__ x, const float* __restrict__ out,
                                                          const float* __restrict__ dout, const float* __restrict__ weight,
                                                          float* __restrict__ dx, float* __restrict__ dW,
                                                          float* __restrict__ db, long long B) {
  static_assert(N % 16 == 0 && K % 8 == 0 && N * K <= 512, "tile shape");
  constexpr int SX = K + 4, SD = N + 4, SO = (N > K ? N : K) + 4, KS = w_stride(K);
  constexpr int N4 = N / 4;
  constexpr int MT = N / 16;       // dW: m-tiles over the units
  constexpr int NK = K / 8;        // dW / dx: n-tiles over the inputs
  constexpr int KN = N / 8;        // dx: k-steps over the units
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;                // [N][KS]
  float* sX = sW + N * KS;         // [kT][SX]
  float* sD = sX + kT * SX;        // [kT][SD]   dout, then dpre
  float* sO = sD + kT * SD;        // [kT][SO]   out, then dx
  for (int i = threadIdx.x; i < N * K; i += kT) sW[(i / K) * KS + i % K] = __ldg(weight + i);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
  const int R0 = warp * 32;
  __syncthreads();
  // B fragments of dx = dpre W: (k = unit 8 ks + tig [+4], n = input 8 nt + g), split once
  unsigned wh[KN][NK][2], wl[KN][NK][2];
#pragma unroll
  for (int ks = 0; ks < KN; ++ks)
#pragma unroll
    for (int nt = 0; nt < NK; ++nt) {
      split_tf32(sW[(8 * ks + tig) * KS + 8 * nt + g], wh[ks][nt][0], wl[ks][nt][0]);
      split_tf32(sW[(8 * ks + tig + 4) * KS + 8 * nt + g], wh[ks][nt][1], wl[ks][nt][1]);
    }
  float cw[MT][NK][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NK; ++nt) cw[mt][nt][0] = cw[mt][nt][1] = cw[mt][nt][2] = cw[mt][nt][3] = 0.0f;
  float gb[N];
#pragma unroll
  for (int n = 0; n < N; ++n) gb[n] = 0.0f;

  const long long ntiles = (B + kT - 1) / kT;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * kT;
    tile_copy<true>(sX, SX, const_cast<float*>(x), row0, B, K);
    tile_copy<true>(sD, SD, const_cast<float*>(dout), row0, B, N);
    tile_copy<true>(sO, SO, const_cast<float*>(out), row0, B, N);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    {  // dpre for this thread's row, in place over dout (rows past B are zero: dpre = 0)
      float4* myd = reinterpret_cast<float4*>(sD + threadIdx.x * SD);
      const float4* myo = reinterpret_cast<const float4*>(sO + threadIdx.x * SO);
#pragma unroll
      for (int c = 0; c < N4; ++c) {
        const float4 d = myd[c], o = myo[c];
        const float4 p = make_float4(d.x * act_bwd<ACT>(o.x), d.y * act_bwd<ACT>(o.y), d.z * act_bwd<ACT>(o.z),
                                     d.w * act_bwd<ACT>(o.w));
        myd[c] = p;
        gb[4 * c] += p.x; gb[4 * c + 1] += p.y; gb[4 * c + 2] += p.z; gb[4 * c + 3] += p.w;
      }
    }
    __syncwarp();
    // ---- dx[32 rows of this warp][K] = dpre W, written over the (consumed) out rows
#pragma unroll
    for (int m2 = 0; m2 < 2; ++m2) {
      const int r0 = R0 + 16 * m2;
      unsigned ah[KN][4], al[KN][4];
#pragma unroll
      for (int ks = 0; ks < KN; ++ks) {
        split_tf32(sD[(r0 + g) * SD + 8 * ks + tig], ah[ks][0], al[ks][0]);
        split_tf32(sD[(r0 + g + 8) * SD + 8 * ks + tig], ah[ks][1], al[ks][1]);
        split_tf32(sD[(r0 + g) * SD + 8 * ks + tig + 4], ah[ks][2], al[ks][2]);
        split_tf32(sD[(r0 + g + 8) * SD + 8 * ks + tig + 4], ah[ks][3], al[ks][3]);
      }
#pragma unroll
      for (int nt = 0; nt < NK; ++nt) {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < KN; ++ks) mma_tf32(c, al[ks], wh[ks][nt]);
#pragma unroll
        for (int ks = 0; ks < KN; ++ks) mma_tf32(c, ah[ks], wl[ks][nt]);
#pragma unroll
        for (int ks = 0; ks < KN; ++ks) mma_tf32(c, ah[ks], wh[ks][nt]);
        *reinterpret_cast<float2*>(sO + (r0 + g) * SO + 8 * nt + 2 * tig) = make_float2(c[0], c[1]);
        *reinterpret_cast<float2*>(sO + (r0 + g + 8) * SO + 8 * nt + 2 * tig) = make_float2(c[2], c[3]);
      }
    }
    // ---- dW[N][K] += dpre^T x over this warp's 32 rows (4 k-steps of 8 rows)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      unsigned ah[MT][4], al[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const float* dp = sD + (R0 + 8 * ks + tig) * SD + 16 * mt + g;
        split_tf32(dp[0], ah[mt][0], al[mt][0]);
        split_tf32(dp[8], ah[mt][1], al[mt][1]);
        split_tf32(dp[4 * SD], ah[mt][2], al[mt][2]);
        split_tf32(dp[4 * SD + 8], ah[mt][3], al[mt][3]);
      }
#pragma unroll
      for (int nt = 0; nt < NK; ++nt) {
        unsigned bh[2], bl[2];
        split_tf32(sX[(R0 + 8 * ks + tig) * SX + 8 * nt + g], bh[0], bl[0]);
        split_tf32(sX[(R0 + 8 * ks + tig + 4) * SX + 8 * nt + g], bh[1], bl[1]);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma_tf32(cw[mt][nt], al[mt], bh);
          mma_tf32(cw[mt][nt], ah[mt], bl);
          mma_tf32(cw[mt][nt], ah[mt], bh);
        }
      }
    }
    __syncthreads();   // every warp's dx rows are in sO
    tile_copy<false>(sO, SO, dx, row0, B, K);
    __syncthreads();   // ... and stored before the next tile's copies land
  }
  // ---- the warps' dW fragments and bias sums meet in shared memory: one atomic per entry per CTA
  float* sAcc = sX;    // [kT / 32][N * K] <= 4 * 1024 floats, inside sX | sD | sO (kT * (SX + SD + SO) >= 128 * 60)
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NK; ++nt) {
      float* p0 = sAcc + warp * (N * K) + (16 * mt + g) * K + 8 * nt + 2 * tig;
      p0[0] = cw[mt][nt][0]; p0[1] = cw[mt][nt][1];
      p0[8 * K] = cw[mt][nt][2]; p0[8 * K + 1] = cw[mt][nt][3];
    }
  __syncthreads();
  for (int o = threadIdx.x; o < N * K; o += kT) {
    float v = 0.0f;
#pragma unroll
    for (int wi = 0; wi < kT / 32; ++wi) v += sAcc[wi * (N * K) + o];
    atomicAdd(dW + o, v);
  }
  __syncthreads();
  float* sRed = sX;    // [kT / 32][N]
#pragma unroll
  for (int n = 0; n < N; ++n) {
    float v = gb[n];
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sft);
    if (lane == 0) sRed[warp * N + n] = v;
  }
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += kT) {
    float v = 0.0f;
#pragma unroll
    for (int wi = 0; wi < kT / 32; ++wi) v += sRed[wi * N + n];
    atomicAdd(db + n, v);
  }
}

// ---------------------------------------------------------------- backward, first layer
// The first layer of the network has few inputs (the conditioning variable: K = 1 .. 4) and nobody wants its
// input gradient: what is left is a streaming reduction -- dW[n][k] = sum_r dpre[r][n] x[r][k], db[n] =
// sum_r dpre[r][n] -- which each thread accumulates in registers over its rows; warps meet with shuffles,
// one atomic per entry per warp at the end.  No shared memory, no barriers.
template <int N, int ACT, int KS>
__global__ void __launch_bounds__(kT) dense_act_bwd_first(const float* __restrict__ x, const float* __restrict__ xmean,
                                                         const float* __restrict__ xstd, const float* __restrict__ out,
                                                         const float* __restrict__ dout, float* __restrict__ dW,
                                                         float* __restrict__ db, long long B, int K) {
  float gw[KS][N], gb[N];
#pragma unroll
  for (int n = 0; n < N; ++n) {
    gb[n] = 0.0f;
#pragma unroll
    for (int k = 0; k < KS; ++k) gw[k][n] = 0.0f;
  }
  for (long long r = (long long)blockIdx.x * kT + threadIdx.x; r < B; r += (long long)gridDim.x * kT) {
    float xs[KS];
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      xs[k] = (k < K) ? __ldg(x + r * K + k) : 0.0f;
      if (xmean && k < K) xs[k] = __fdiv_rn(xs[k] - __ldg(xmean + k), __ldg(xstd + k) + 1e-8f);
    }
    const float4* po = reinterpret_cast<const float4*>(out + r * N);
    const float4* pg = reinterpret_cast<const float4*>(dout + r * N);
#pragma unroll
    for (int c = 0; c < N / 4; ++c) {
      const float4 o = __ldg(po + c), g = __ldg(pg + c);
      const float d[4] = {g.x * act_bwd<ACT>(o.x), g.y * act_bwd<ACT>(o.y), g.z * act_bwd<ACT>(o.z),
                          g.w * act_bwd<ACT>(o.w)};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        gb[4 * c + j] += d[j];
#pragma unroll
        for (int k = 0; k < KS; ++k) gw[k][4 * c + j] = fmaf(xs[k], d[j], gw[k][4 * c + j]);
      }
    }
  }
  // warps meet with shuffles, the CTA's warps in shared memory: ONE atomic per entry per CTA (atomics on the
  // same few addresses serialise in L2: per-warp atomics cost more than the whole streaming pass)
  __shared__ float sRed[kT / 32][(KS + 1) * N];
#pragma unroll
  for (int n = 0; n < N; ++n) {
    float v = gb[n];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if ((threadIdx.x & 31) == 0) sRed[threadIdx.x >> 5][n] = v;
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      float w = gw[k][n];
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) w += __shfl_xor_sync(0xffffffffu, w, s);
      if ((threadIdx.x & 31) == 0) sRed[threadIdx.x >> 5][(k + 1) * N + n] = w;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < (KS + 1) * N; e += kT) {
    float v = 0.0f;
#pragma unroll
    for (int wi = 0; wi < kT / 32; ++wi) v += sRed[wi][e];
    const int k = e / N - 1, n = e % N;
    if (k < 0) atomicAdd(db + n, v);
    else if (k < K) atomicAdd(dW + n * K + k, v);
  }
}

template <int N, int ACT>
cudaError_t launch_bwd_first(const float* x, const float* xmean, const float* xstd, const float* out, const float* dout,
                             float* dW, float* db, long long B, int K, cudaStream_t st) {
  const DeviceInfo& di = device_info();
  const long long ntiles = (B + kT - 1) / kT;
  long long grid = (long long)di.sm_count * 4;
  if (grid > ntiles) grid = ntiles;
  if (K == 1) dense_act_bwd_first<N, ACT, 1><<<(unsigned)grid, kT, 0, st>>>(x, xmean, xstd, out, dout, dW, db, B, K);
  else dense_act_bwd_first<N, ACT, 4><<<(unsigned)grid, kT, 0, st>>>(x, xmean, xstd, out, dout, dW, db, B, K);
  count_launch();
  return cudaGetLastError();
}

template <int N, int ACT>
cudaError_t launch_fwd(const float* x, const float* xmean, const float* xstd, const float* w, const float* b, float* out,
                       long long B, int K, cudaStream_t st) {
  const DeviceInfo& di = device_info();
  const size_t smem = (size_t)(pad4(K * N) + N + kT * (pad4(K) + 4) + kT * (N + 4)) * sizeof(float);
  auto kern = (K == 16) ? dense_act_fwd<N, ACT, 16> : dense_act_fwd<N, ACT, 0>;
  static thread_local int configured_for = -1;
  if (configured_for != di.device) {
    cudaError_t e = cudaFuncSetAttribute(dense_act_fwd<N, ACT, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(dense_act_fwd<N, ACT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    if (e != cudaSuccess) return e;
    configured_for = di.device;
  }
  const long long ntiles = (B + kT - 1) / kT;
  long long grid = (long long)di.sm_count * 6;
  if (grid > ntiles) grid = ntiles;
  kern<<<(unsigned)grid, kT, smem, st>>>(x, xmean, xstd, w, b, out, B, K);
  count_launch();
  return cudaGetLastError();
}

template <int N, int ACT, int PER>
cudaError_t launch_bwd(const float* x, const float* out, const float* dout, const float* w, float* dx, float* dW, float* db,
                       long long B, int K, cudaStream_t st) {
  const DeviceInfo& di = device_info();
  const int K4 = (K + 3) / 4;
  const int so = (N > 4 * K4 ? N : 4 * K4) + 4;
  const size_t smem = (size_t)(pad4(K * N) + kT * (4 * K4 + 4) + kT * (N + 4) + kT * so) * sizeof(float);
  auto kern = (K == 16) ? dense_act_bwd<N, ACT, PER, 16> : dense_act_bwd<N, ACT, PER, 0>;
  static thread_local int configured_for = -1;
  if (configured_for != di.device) {
    cudaError_t e = cudaFuncSetAttribute(dense_act_bwd<N, ACT, PER, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(dense_act_bwd<N, ACT, PER, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    if (e != cudaSuccess) return e;
    configured_for = di.device;
  }
  const long long ntiles = (B + kT - 1) / kT;
  long long grid = (long long)di.sm_count * 4;
  if (grid > ntiles) grid = ntiles;
  kern<<<(unsigned)grid, kT, smem, st>>>(x, out, dout, w, dx, dW, db, B, K);
  count_launch();
  return cudaGetLastError();
}

template <int N, int K, int ACT>
cudaError_t launch_bwd_mma(const float* x, const float* out, const float* dout, const float* w, float* dx, float* dW, float* db,
                           long long B, cudaStream_t st) {
  const DeviceInfo& di = device_info();
  constexpr int so = (N > K ? N : K) + 4;
  constexpr size_t smem = (size_t)(N * w_stride(K) + kT * (K + 4) + kT * (N + 4) + kT * so) * sizeof(float);
  auto kern = dense_act_bwd_mma<N, K, ACT>;
  static thread_local int configured_for = -1;
  if (configured_for != di.device) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured_for = di.device;
  }
  const long long ntiles = (B + kT - 1) / kT;
  long long grid = (long long)di.sm_count * 4;
  if (grid > ntiles) grid = ntiles;
  kern<<<(unsigned)grid, kT, smem, st>>>(x, out, dout, w, dx, dW, db, B);
  count_launch();
  return cudaGetLastError();
}

template <int N, int K>
cudaError_t dispatch_bwd_mma(int act, const float* x, const float* out, const float* dout, const float* w, float* dx,
                             float* dW, float* db, long long B, cudaStream_t st) {
  switch (act) {
    case kLinear: return launch_bwd_mma<N, K, kLinear>(x, out, dout, w, dx, dW, db, B, st);
    case kTanh: return launch_bwd_mma<N, K, kTanh>(x, out, dout, w, dx, dW, db, B, st);
    case kRelu: return launch_bwd_mma<N, K, kRelu>(x, out, dout, w, dx, dW, db, B, st);
    case kSigmoid: return launch_bwd_mma<N, K, kSigmoid>(x, out, dout, w, dx, dW, db, B, st);
    default: return launch_bwd_mma<N, K, kElu>(x, out, dout, w, dx, dW, db, B, st);
  }
}

template <int N>
cudaError_t dispatch_fwd(int act, const float* x, const float* xmean, const float* xstd, const float* w, const float* b,
                         float* out, long long B, int K, cudaStream_t st) {
  switch (act) {
    case kLinear: return launch_fwd<N, kLinear>(x, xmean, xstd, w, b, out, B, K, st);
    case kTanh: return launch_fwd<N, kTanh>(x, xmean, xstd, w, b, out, B, K, st);
    case kRelu: return launch_fwd<N, kRelu>(x, xmean, xstd, w, b, out, B, K, st);
    case kSigmoid: return launch_fwd<N, kSigmoid>(x, xmean, xstd, w, b, out, B, K, st);
    default: return launch_fwd<N, kElu>(x, xmean, xstd, w, b, out, B, K, st);
  }
}

template <int N, int PER>
cudaError_t dispatch_bwd(int act, const float* x, const float* xmean, const float* xstd, const float* out,
                         const float* dout, const float* w, float* dx, float* dW, float* db, long long B, int K,
                         cudaStream_t st) {
  // the fused input normalisation exists where it is used: first layer of the network (few inputs, no input gradient)
  if (xmean && !(N <= 32 && dx == nullptr && K <= 4)) return cudaErrorNotSupported;
  if constexpr (N <= 32) {
    if (dx == nullptr && K <= 4) {   // first layer of the network: streaming reduction, no staging
      switch (act) {
        case kLinear: return launch_bwd_first<N, kLinear>(x, xmean, xstd, out, dout, dW, db, B, K, st);
        case kTanh: return launch_bwd_first<N, kTanh>(x, xmean, xstd, out, dout, dW, db, B, K, st);
        case kRelu: return launch_bwd_first<N, kRelu>(x, xmean, xstd, out, dout, dW, db, B, K, st);
        case kSigmoid: return launch_bwd_first<N, kSigmoid>(x, xmean, xstd, out, dout, dW, db, B, K, st);
        default: return launch_bwd_first<N, kElu>(x, xmean, xstd, out, dout, dW, db, B, K, st);
      }
    }
  }
  switch (act) {
    case kLinear: return launch_bwd<N, kLinear, PER>(x, out, dout, w, dx, dW, db, B, K, st);
    case kTanh: return launch_bwd<N, kTanh, PER>(x, out, dout, w, dx, dW, db, B, K, st);
    case kRelu: return launch_bwd<N, kRelu, PER>(x, out, dout, w, dx, dW, db, B, K, st);
    case kSigmoid: return launch_bwd<N, kSigmoid, PER>(x, out, dout, w, dx, dW, db, B, K, st);
    default: return launch_bwd<N, kElu, PER>(x, out, dout, w, dx, dW, db, B, K, st);
  }
}

// ---------------------------------------------------------------- first layer with folded weight draws
// The Bayesian estimators fold S posterior weight draws into the batch (reference BayesianNNEstimator.py:65-76;
// BASELINE config 4): row s * Bd + b is sample b under draw s.  The first variational layer has few inputs
// (K <= 8) and per-draw weights w[s] = [kernel (K x N) | bias (N)] (tfp DenseVariational's flat layout), so one
// draw's weights are a few hundred bytes of shared memory and x is read once per draw from L2:
//   forward   out[s Bd + b][n] = act(xn_b . kernel_s[:, n] + bias_s[n])  for n < N, zero for N <= n < NP
//             (NP = the row width the fused Dense(P)+head kernel wants: a multiple of 16)
//   backward  dw[s] += [ xn^T dpre_s | 1^T dpre_s ],  dpre = dout * act'(out)      (no input gradient)
// blockIdx.y = draw; blockIdx.z (backward) = an 8-column slice of the units, whose (K + 1) x 8 sums stay in
// registers over the thread's rows and meet through warp shuffles and shared memory once, at the end.
constexpr int kDrawT = 256;
constexpr int kDrawMaxK = 8;

__device__ __forceinline__ void load_xn(const float* __restrict__ x, const float* __restrict__ xmean,
                                        const float* __restrict__ xstd, long long b, int K, float (&xn)[kDrawMaxK]) {
#pragma unroll
  for (int k = 0; k < kDrawMaxK; ++k) {
    float v = 0.0f;
    if (k < K) {
      v = __ldg(x + b * K + k);
      if (xmean) v = __fdiv_rn(v - __ldg(xmean + k), __ldg(xstd + k) + 1e-8f);
    }
    xn[k] = v;
  }
}

template <int ACT>
__global__ void __launch_bounds__(kDrawT) dense_act_draws_fwd(const float* __restrict__ x, const float* __restrict__ xmean,
                                                            const float* __restrict__ xstd, const float* __restrict__ w,
                                                            float* __restrict__ out, long long Bd, int K, int N, int NP) {
  extern __shared__ __align__(16) float smem[];   // [K * N + N]
  const int s = blockIdx.y;
  const float* ws = w + (long long)s * (K * N + N);
  for (int i = threadIdx.x; i < K * N + N; i += kDrawT) smem[i] = __ldg(ws + i);
  __syncthreads();
  const float* sb = smem + K * N;
  for (long long b = (long long)blockIdx.x * kDrawT + threadIdx.x; b < Bd; b += (long long)gridDim.x * kDrawT) {
    float xn[kDrawMaxK];
    load_xn(x, xmean, xstd, b, K, xn);
    float* o = out + ((long long)s * Bd + b) * NP;
    for (int n0 = 0; n0 < NP; n0 += 4) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + j;
        float acc = 0.0f;
        if (n < N) {
          acc = sb[n];
#pragma unroll
          for (int k = 0; k < kDrawMaxK; ++k)
            if (k < K) acc = fmaf(xn[k], smem[k * N + n], acc);
          acc = act_fwd<ACT>(acc);
        }
        v[j] = acc;
      }
      *reinterpret_cast<float4*>(o + n0) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

template <int ACT>
__global__ void __launch_bounds__(kDrawT, 2) dense_act_draws_bwd(const float* __restrict__ x, const float* __restrict__ xmean,
                                                            const float* __restrict__ xstd, const float* __restrict__ out,
                                                            const float* __restrict__ dout, float* __restrict__ dw,
                                                            long long Bd, int K, int N, int NP) {
  __shared__ float part[kDrawT / 32][(kDrawMaxK + 1) * 8];
  const int s = blockIdx.y, n0 = 8 * blockIdx.z;
  float acc[kDrawMaxK + 1][8];
#pragma unroll
  for (int k = 0; k <= kDrawMaxK; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = 0.0f;
  // two rows per iteration: eight independent 16-byte loads in flight per thread (the pass is latency-, not
  // bandwidth-bound with one CTA wave of ~14 rows per thread)
  const long long stride = (long long)gridDim.x * kDrawT;
  for (long long b0 = (long long)blockIdx.x * kDrawT + threadIdx.x; b0 < Bd; b0 += 2 * stride) {
    float4 o[2][2], d[2][2];
    float xn[2][kDrawMaxK];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long b = b0 + u * stride;
      if (b < Bd) {
        const long long row = ((long long)s * Bd + b) * NP + n0;
        o[u][0] = __ldg(reinterpret_cast<const float4*>(out + row));
        o[u][1] = __ldg(reinterpret_cast<const float4*>(out + row + 4));
        d[u][0] = __ldg(reinterpret_cast<const float4*>(dout + row));
        d[u][1] = __ldg(reinterpret_cast<const float4*>(dout + row + 4));
        load_xn(x, xmean, xstd, b, K, xn[u]);
      } else {
        o[u][0] = o[u][1] = d[u][0] = d[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);   // dpre = 0: contributes nothing
#pragma unroll
        for (int k = 0; k < kDrawMaxK; ++k) xn[u][k] = 0.0f;
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const float ov[8] = {o[u][0].x, o[u][0].y, o[u][0].z, o[u][0].w, o[u][1].x, o[u][1].y, o[u][1].z, o[u][1].w};
      const float dv[8] = {d[u][0].x, d[u][0].y, d[u][0].z, d[u][0].w, d[u][1].x, d[u][1].y, d[u][1].z, d[u][1].w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dpre = dv[j] * act_bwd<ACT>(ov[j]);
        acc[kDrawMaxK][j] += dpre;
#pragma unroll
        for (int k = 0; k < kDrawMaxK; ++k) acc[k][j] = fmaf(xn[u][k], dpre, acc[k][j]);
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k <= kDrawMaxK; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = acc[k][j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) part[warp][k * 8 + j] = v;
    }
  __syncthreads();
  float* dws = dw + (long long)s * (K * N + N);
  for (int e = threadIdx.x; e < (kDrawMaxK + 1) * 8; e += kDrawT) {
    const int k = e / 8, n = n0 + e % 8;
    if (n < N && (k < K || k == kDrawMaxK)) {
      float v = 0.0f;
#pragma unroll
      for (int wq = 0; wq < kDrawT / 32; ++wq) v += part[wq][e];
      atomicAdd(dws + (k == kDrawMaxK ? K * N + n : k * N + n), v);
    }
  }
}

template <int ACT>
cudaError_t launch_draws(bool bwd, const float* x, const float* xmean, const float* xstd, const float* w, const float* out_in,
                         const float* dout, float* out, float* dw, int S, long long Bd, int K, int N, int NP,
                         cudaStream_t st) {
  const DeviceInfo& di = device_info();
  const long long tiles = (Bd + kDrawT - 1) / kDrawT;
  const int slices = bwd ? (N + 7) / 8 : 1;
  // forward: ~8 CTAs per SM overall; backward: ONE resident wave (2 CTAs per SM at its register count), so that the
  // shuffle / shared-memory / atomic epilogue is paid once per ~14 rows of a thread instead of once per ~3
  // (rounded DOWN for the backward pass: 320 CTAs on 296 slots ran as two waves, the second nearly empty: 60 -> 42 us.
  // Staging the forward rows through shared memory for fully coalesced stores was tried: 29 -> 34 us, the two CTA
  // barriers per ~4 rows of a thread cost more than the half-filled store sectors)
  const long long per_sm = bwd ? 2 : 8;
  long long gx = bwd ? ((long long)di.sm_count * per_sm) / ((long long)S * slices)
                     : ((long long)di.sm_count * per_sm + (long long)S * slices - 1) / ((long long)S * slices);
  if (gx < 1) gx = 1;
  if (gx > tiles) gx = tiles;
  const dim3 grid((unsigned)gx, (unsigned)S, (unsigned)slices);
  if (bwd) dense_act_draws_bwd<ACT><<<grid, kDrawT, 0, st>>>(x, xmean, xstd, out_in, dout, dw, Bd, K, N, NP);
  else dense_act_draws_fwd<ACT><<<grid, kDrawT, (size_t)(K * N + N) * sizeof(float), st>>>(x, xmean, xstd, w, out, Bd, K, N, NP);
  count_launch();
  return cudaGetLastError();
}

}  // namespace

int mlp_layer_supported(int K, int N, int act) {
  return K >= 1 && K <= 64 && (N == 8 || N == 16 || N == 32 || N == 64) && act >= 0 && act <= 4;
}

int launch_dense_act_forward(const float* x, const float* xmean, const float* xstd, const float* w, const float* b,
                             float* out, long long B, int K, int N, int act, cudaStream_t st) {
  cudaError_t e;
  switch (N) {
    case 8: e = dispatch_fwd<8>(act, x, xmean, xstd, w, b, out, B, K, st); break;
    case 16: e = dispatch_fwd<16>(act, x, xmean, xstd, w, b, out, B, K, st); break;
    case 32: e = dispatch_fwd<32>(act, x, xmean, xstd, w, b, out, B, K, st); break;
    default: e = dispatch_fwd<64>(act, x, xmean, xstd, w, b, out, B, K, st); break;
  }
  return cuda_error(e, "dense_act_fwd");
}

int launch_dense_act_backward(const float* x, const float* xmean, const float* xstd, const float* out, const float* dout,
                              const float* w, float* dx, float* dW, float* db, long long B, int K, int N, int act,
                              cudaStream_t st) {
  cudaError_t e;
  // hidden layers of the usual widths (input gradient wanted): both GEMMs on the tensor cores (NFN_B200_MLP_MMA=0: off)
  // (32 x 32 keeps the scalar kernel: its 64 registers of split W fragments next to 32 accumulators spill)
  if (option(kOptMlpMma) && dx != nullptr && xmean == nullptr && (K == 16 || K == 32) && (N == 16 || N == 32) &&
      N * K <= 512) {
    if (N == 16) e = K == 16 ? dispatch_bwd_mma<16, 16>(act, x, out, dout, w, dx, dW, db, B, st)
                             : dispatch_bwd_mma<16, 32>(act, x, out, dout, w, dx, dW, db, B, st);
    else e = dispatch_bwd_mma<32, 16>(act, x, out, dout, w, dx, dW, db, B, st);
    return cuda_error(e, "dense_act_bwd_mma");
  }
  // 4 x 4 blocks of the weight gradient per thread: one while ceil(K/4) * N/4 <= 128, else two (K * N <= 4096)
  const int G = ((K + 3) / 4) * (N / 4);
  if (G <= kT) {
    switch (N) {
      case 8: e = dispatch_bwd<8, 1>(act, x, xmean, xstd, out, dout, w, dx, dW, db, B, K, st); break;
      case 16: e = dispatch_bwd<16, 1>(act, x, xmean, xstd, out, dout, w, dx, dW, db, B, K, st); break;
      case 32: e = dispatch_bwd<32, 1>(act, x, xmean, xstd, out, dout, w, dx, dW, db, B, K, st); break;
      default: e = dispatch_bwd<64, 1>(act, x, xmean, xstd, out, dout, w, dx, dW, db, B, K, st); break;
    }
  } else {
    switch (N) {
      case 8: e = dispatch_bwd<8, 2>(act, x, xmean, xstd, out, dout, w, dx, dW, db, B, K, st); break;
      case 16: e = dispatch_bwd<16, 2>(act, x, xmean, xstd, out, dout, w, dx, dW, db, B, K, st); break;
      case 32: e = dispatch_bwd<32, 2>(act, x, xmean, xstd, out, dout, w, dx, dW, db, B, K, st); break;
      default: e = dispatch_bwd<64, 2>(act, x, xmean, xstd, out, dout, w, dx, dW, db, B, K, st); break;
    }
  }
  return cuda_error(e, "dense_act_bwd");
}

int mlp_draws_supported(int K, int N, int NP, int act) {
  return K >= 1 && K <= kDrawMaxK && N >= 1 && N <= 64 && NP >= N && NP % 8 == 0 && NP <= 64 && act >= 0 && act <= 4;
}

int launch_dense_act_draws(bool bwd, const float* x, const float* xmean, const float* xstd, const float* w,
                           const float* out_in, const float* dout, float* out, float* dw, int S, long long Bd, int K, int N,
                           int NP, int act, cudaStream_t st) {
  cudaError_t e;
  switch (act) {
    case kLinear: e = launch_draws<kLinear>(bwd, x, xmean, xstd, w, out_in, dout, out, dw, S, Bd, K, N, NP, st); break;
    case kTanh: e = launch_draws<kTanh>(bwd, x, xmean, xstd, w, out_in, dout, out, dw, S, Bd, K, N, NP, st); break;
    case kRelu: e = launch_draws<kRelu>(bwd, x, xmean, xstd, w, out_in, dout, out, dw, S, Bd, K, N, NP, st); break;
    case kSigmoid: e = launch_draws<kSigmoid>(bwd, x, xmean, xstd, w, out_in, dout, out, dw, S, Bd, K, N, NP, st); break;
    default: e = launch_draws<kElu>(bwd, x, xmean, xstd, w, out_in, dout, out, dw, S, Bd, K, N, NP, st); break;
  }
  return cuda_error(e, bwd ? "dense_act_draws_bwd" : "dense_act_draws_fwd");
}

}  // namespace nfn
