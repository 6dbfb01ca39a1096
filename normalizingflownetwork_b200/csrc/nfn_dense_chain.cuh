// nfn_dense_chain.cuh -- the emitting Dense(P) layer fused into the flow-chain kernel (sm_100a).
//
// the parameter tensor t[B, P] is produced by the last Dense layer of
// the conditioning network (reference estimators/MaximumLikelihoodNNEstimator.py:43) from a
// 16-64 wide hidden activation h[B, H], and is the dominant HBM traffic of the hot path.  This
// kernel consumes h directly:
//
//   per tile of T rows:   t  = h W + b            (tensor cores, tile stays in shared memory)
//                          flows forward + reverse sweep per row (nfn_flows.cuh), dt in place
//                          dh = dt W^T,  dW += h^T dt,  db += 1^T dt     (tensor cores)
//
// so per row it reads 4(H + d) and writes 4(H + 1) bytes instead of 4(2P + d + 1), and the three
// skinny cuBLAS GEMMs around the flow kernel (measured 124 + 146 + 215 us + 75 us for the bias sum
// next to a 71 us flow kernel at B = 2^20, H = 16, P = 48) disappear.
//
// The GEMMs are tiny (K = 16..64): warp-level mma.sync.m16n8k8 TF32 with the 3xTF32 split
// (x = hi + lo, D += A_lo B_hi + A_hi B_lo + A_hi B_hi) keeps fp32-level accuracy, which the
// 1e-5 log-prob parity bar needs; plain TF32 (10-bit mantissa) would not pass it.
#pragma once
#include "nfn_chain_kernel.cuh"
#include "nfn_mixture_row.cuh"

namespace nfn {

struct DenseArgs {
  const float* h;     // [B, H] last hidden activation
  const float* W;     // [H, P] row-major (Keras kernel layout)
  const float* bias;  // [P]
  const float* y;
  const float* g_logp;
  float* logp;
  float* dh;          // [B, H]
  float* dW;          // [H, P]  += h^T dt
  float* dbias;       // [P]     += sum_b dt
  double* logp_sum;
  long long B;
  float g_scale;
  int y_broadcast;
  EventXform xf;
  // S posterior weight draws folded into the batch (reference BayesianNNEstimator.py:65-76; mma.sync body only):
  // rows are draw-major, row s * rows_per_draw + b is sample b under draw s; W [S][H][P], bias [S][P], dW [S][H][P],
  // dbias [S][P] are per draw, y [rows_per_draw][d] is per SAMPLE (not repeated); h, logp, dh, g_logp are folded.
  // draws <= 1: one weight set for all B rows.
  int draws;
  long long rows_per_draw;
  // Per-draw weights in tfp DenseVariational's FLAT sample layout (flat_rows > 0): draw s owns flat_stride floats at
  // W + s * flat_stride = [kernel (flat_rows x P) | bias (P)], flat_rows <= H (rows flat_rows .. H of the operand are
  // zero, matching the zero columns of a padded hidden row); dW has the same layout and receives the bias gradient
  // behind the kernel's.  `bias` / `dbias` are ignored then.
  int flat_rows;
  long long flat_stride;
  // KMN head only: fixed centres [M][d], shared bandwidths [M] (may be negative), their gradient [M] += (nullable)
  const float* locs;
  const float* scales;
  float* dscales;
};

// ---------------------------------------------------------------- tensor-core helpers
// A TF32 operand is the upper 19 bits of an fp32 pattern: the tensor core ignores the low 13
// mantissa bits, so the "high" part of a split needs no conversion at all (cvt.rna.tf32 is
// emulated with ~6 instructions on sm_100 and dominated the first version of this kernel).
// The residual x - trunc(x) is exact in fp32 and carries the next 11+ bits.
NFN_DEVI float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
NFN_DEVI void split_tf32(float x, unsigned& hi, unsigned& lo) {
  hi = __float_as_uint(x);
  lo = __float_as_uint(x - tf32_trunc(x));
}
// (not volatile: independent accumulators must be free to interleave -- a chain of dependent HMMAs
// costs its full latency per instruction)
NFN_DEVI void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// three-level split x = x0 + x1 + x2 (11 + 11 + remaining bits): the forward GEMM t = h W + b feeds the
// flows, whose log-prob can amplify a rounding of t by |dlogp/dt| ~ 10^2, so t gets full fp32 accuracy
NFN_DEVI void split3_tf32(float x, unsigned& x0, unsigned& x1, unsigned& x2) {
  x0 = __float_as_uint(x);
  const float r1 = x - tf32_trunc(x);
  x1 = __float_as_uint(r1);
  x2 = __float_as_uint(r1 - tf32_trunc(r1));
}
// D += A B with the six leading products of the three-level splits (smallest first)
NFN_DEVI void mma_6xtf32(float (&d)[4], const unsigned (&a0)[4], const unsigned (&a1)[4], const unsigned (&a2)[4],
                         const unsigned (&b0)[2], const unsigned (&b1)[2], const unsigned (&b2)[2]) {
  mma_tf32(d, a2, b0);
  mma_tf32(d, a0, b2);
  mma_tf32(d, a1, b1);
  mma_tf32(d, a1, b0);
  mma_tf32(d, a0, b1);
  mma_tf32(d, a0, b0);
}
// D += A B at fp32-level accuracy (small terms first)
NFN_DEVI void mma_3xtf32(float (&d)[4], const unsigned (&ah)[4], const unsigned (&al)[4], const unsigned (&bh)[2],
                         const unsigned (&bl)[2]) {
  mma_tf32(d, al, bh);
  mma_tf32(d, ah, bl);
  mma_tf32(d, ah, bh);
}

// smallest stride >= n with stride % 32 in {8, 24}: B-fragment reads (k = lane%4, n = lane/4) are
// then bank-conflict free
__host__ __device__ constexpr int w_stride(int n) {
  int s = n;
  while (!(s % 32 == 8 || s % 32 == 24)) ++s;
  return s;
}

// dynamic shared memory of the fused dense kernel (same arithmetic as DenseGeometry below)
__host__ __device__ constexpr unsigned dense_smem_bytes(int P, int H, int T, bool bwd, int head_extra_floats = 0) {
  const int S = row_stride(P), HS = H + 4, P8 = (P + 7) / 8 * 8, PW = w_stride(P8), NW = T / 32;
  return (unsigned)(4 * (T * S + 2 * T * HS + H * PW + P8 + (bwd ? NW * H * P8 + NW * P8 : 0) + head_extra_floats));
}
// shared-memory floats the KMN head keeps next to the tiles: centres, two per-kernel coefficients, per-warp
// bandwidth-gradient sums
__host__ __device__ constexpr int kmn_extra_floats(int M, int D, int T, bool bwd) {
  return M * (D + 2) + (bwd ? (T / 32) * M : 0);
}

template <int P, int H, int T>
struct DenseGeometry {
  static constexpr int S = row_stride(P);         // t / dt tile row stride
  static constexpr int HS = H + 4;                // h tile row stride: A-fragment reads conflict-free
  static constexpr int P8 = (P + 7) / 8 * 8;      // P padded to the mma N / K granule
  static constexpr int PW = w_stride(P8);         // W row stride in smem
  static constexpr int NW = T / 32;
  static constexpr int kT = T * S;                // floats
  static constexpr int kH = T * HS;               // per h buffer
  static constexpr int kW = H * PW;
  static constexpr int kAcc = NW * H * P8;        // per-warp dW partial sums
  static constexpr int kAccB = NW * P8;           // per-warp dbias partial sums
  __host__ __device__ static constexpr unsigned smem_bytes(bool bwd) {
    return (unsigned)(sizeof(float) * (kT + 2 * kH + kW + P8 + (bwd ? kAcc + kAccB : 0)));
  }
};

// ---------------------------------------------------------------- what runs on a parameter row
// A head turns row r of t (in shared memory, stride S) and the event y_r into log p(y_r | t_r) and, in place,
// cot * d log p / d t_r.  Two of them share the GEMM scaffolding below.
//   ChainHead<Spec>: the inverted flow chain (reference estimators/DistributionLayers.py:215-294)
// A head may keep per-kernel constants in shared memory behind the tiles (`extra_floats`, `stage`, `finish`) and may
// ask to be called by ALL lanes of a warp, rows past the end included (`kAllLanes`: warp-level reductions inside).
struct HeadNoState {
  static constexpr bool kAllLanes = false;
  __host__ __device__ static constexpr int extra_floats(int, bool) { return 0; }
  template <bool BWD> NFN_DEVI static void stage(const DenseArgs&, float*, int, int) {}
  template <bool BWD> NFN_DEVI static void finish(const DenseArgs&, float*, int, int) {}
};

template <class Spec>
struct ChainHead : HeadNoState {
  static constexpr int D = Spec::D;
  static constexpr int P = Spec::P();
  template <bool BWD, class M>
  NFN_DEVI static float run(float* row, float (&z)[D], float cot, bool, float*) {
    constexpr int V = row_vec(P);
    float zs[Spec::KA][D];
    LogDetAcc<M> ld;
    FwdSweep<Spec, M, V, BWD, 0>::run(row, z, zs, ld);
    using Base = BaseDist<D, Spec::BASE, M>;
    float bth[Base::NA];
    if constexpr (Spec::BASE) Span<0, 2 * D, V>::load(row, bth);
    const float lp = (BWD ? Base::log_prob_save(bth, z) : Base::log_prob(bth, z)) + ld.nat();
    if constexpr (BWD) {
      float Gz[D];
      float gb[Base::NA];
      Base::bwd_saved(bth, z, cot, Gz, gb);
      if constexpr (Spec::BASE) Span<0, 2 * D, V>::store(row, gb);
      BwdSweep<Spec, M, V, Spec::K - 1>::run(row, zs, Gz, cot);
    }
    return lp;
  }
};
//   MdnHead<K, D>: the K-component Gaussian mixture (reference estimators/DistributionLayers.py:196-212, fed by
//   the Dense(P) layer of MaximumLikelihoodNNEstimator.py:43); row arithmetic shared with nfn_mixture.cu
template <int K_, int D_>
struct MdnHead : HeadNoState {
  static constexpr int D = D_;
  static constexpr int K = K_;
  static constexpr int P = K_ * (2 * D_ + 1);
  template <bool BWD, class M>
  NFN_DEVI static float run(float* row, float (&z)[D], float cot, bool, float*) {
    constexpr bool V4 = (P % 4 == 0);           // row_stride keeps 16-byte row alignment then
    constexpr int LG = (V4 && K % 4 == 0) ? 4 : 1;
    float dy[D];
#pragma unroll
    for (int i = 0; i < D; ++i) dy[i] = 0.0f;
    return mdn_row<D, V4, LG, BWD, M>(row, K, z, cot, dy);
  }
};

//   KmnHead<M, D>: M fixed Gaussian kernels with shared bandwidths, the row is their logits (reference
//   estimators/DistributionLayers.py:74-133, fed by the Dense(P) layer of MaximumLikelihoodNNEstimator.py:43); row
//   arithmetic shared with nfn_mixture.cu.  Shared-memory state: centres, -0.5 log2e / s^2, -d log2|s|, and (BWD) one
//   row of bandwidth-gradient sums per warp.
template <int M_, int D_>
struct KmnHead {
  static constexpr int D = D_;
  static constexpr int P = M_;
  static constexpr bool kAllLanes = true;
  __host__ __device__ static constexpr int extra_floats(int T, bool bwd) { return kmn_extra_floats(M_, D_, T, bwd); }
  template <bool BWD>
  NFN_DEVI static void stage(const DenseArgs& a, float* ex, int tid, int T) {
    float* s_loc = ex;
    float* s_coef = s_loc + M_ * D_;
    float* s_lnorm = s_coef + M_;
    for (int i = tid; i < M_ * D_; i += T) s_loc[i] = __ldg(a.locs + i);
    for (int i = tid; i < M_; i += T) {
      const float sc = __ldg(a.scales + i);
      s_coef[i] = -0.5f * kLog2e / (sc * sc);
      s_lnorm[i] = -(float)D_ * log2f(fabsf(sc));
    }
    if constexpr (BWD) {
      float* s_dsc = s_lnorm + M_;
      for (int i = tid; i < (T / 32) * M_; i += T) s_dsc[i] = 0.0f;
    }
  }
  template <bool BWD, class Mth>
  NFN_DEVI static float run(float* row, float (&z)[D], float cot, bool valid, float* ex) {
    constexpr int LG = (M_ % 4 == 0) ? 4 : 1;   // row_stride keeps 16-byte row alignment then
    float* s_loc = ex;
    float* s_coef = s_loc + M_ * D_;
    float* s_lnorm = s_coef + M_;
    float* my_dsc = BWD ? s_lnorm + M_ + (threadIdx.x >> 5) * M_ : nullptr;
    float dy[D];
#pragma unroll
    for (int i = 0; i < D; ++i) dy[i] = 0.0f;
    return kmn_row<D, LG, BWD, Mth>(row, M_, z, cot, valid, s_loc, s_coef, s_lnorm, my_dsc, dy);
  }
  template <bool BWD>
  NFN_DEVI static void finish(const DenseArgs& a, float* ex, int tid, int T) {
    if constexpr (BWD) {
      if (a.dscales) {
        const float* s_dsc = ex + M_ * (D_ + 2);
        for (int i = tid; i < M_; i += T) {
          float v = 0.0f;
          for (int w = 0; w < T / 32; ++w) v += s_dsc[w * M_ + i];
          atomicAdd(a.dscales + i, v / __ldg(a.scales + i));
        }
      }
    }
  }
};

// ---------------------------------------------------------------- the fused body
template <class Head, int H, bool BWD, class M, int T>
NFN_DEVI void dense_head_body(const DenseArgs& a) {
  constexpr int D = Head::D;
  constexpr int P = Head::P;
  static_assert(P > 0, "the fused dense kernel needs at least one parameter column");
  static_assert(H % 16 == 0 && H >= 16 && H <= 64, "hidden width must be 16, 32, 48 or 64");
  using G = DenseGeometry<P, H, T>;
  constexpr int S = G::S, HS = G::HS, P8 = G::P8, PW = G::PW, NW = G::NW;
  constexpr int NT = P8 / 8;   // n-tiles over the parameter columns
  constexpr int KH = H / 8;    // k-steps over the hidden units

  extern __shared__ __align__(16) float smem[];
  __shared__ double red[T / 32];
  float* tT = smem;                    // [T][S]   t, then dt in place
  float* hT = tT + G::kT;              // [2][T][HS]
  float* sW = hT + 2 * G::kH;          // [H][PW]
  float* sB = sW + G::kW;              // [P8]
  float* sAcc = sB + P8;               // [NW][H][P8]   (BWD)
  float* sAccB = sAcc + G::kAcc;       // [NW][P8]      (BWD)
  float* sHead = sB + P8 + (BWD ? G::kAcc + G::kAccB : 0);   // the head's own constants / accumulators

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tig = lane & 3;
  // Tiles never straddle two draws: a draw is cut into its own tiles, and with folded draws every CTA takes a
  // CONTIGUOUS range of tiles, so that it changes weights at most a few times (one weight set: round-robin tiles)
  const bool batched = a.draws > 1;
  const long long Bd = batched ? a.rows_per_draw : a.B;       // rows that share one weight set
  const long long tpd = (Bd + T - 1) / T;                     // tiles per draw
  const long long ntiles = (batched ? (long long)a.draws : 1ll) * tpd;
  long long tile = batched ? ntiles * blockIdx.x / gridDim.x : blockIdx.x;
  const long long tile_end = batched ? ntiles * (blockIdx.x + 1) / gridDim.x : ntiles;
  const long long tstep = batched ? 1 : gridDim.x;
  // tile -> (draw, first row inside the draw, first folded row)
  auto draw_of = [&](long long t) -> long long { return batched ? t / tpd : 0; };
  auto rd0_of = [&](long long t) -> long long { return (t - draw_of(t) * tpd) * T; };
  auto fold0_of = [&](long long t) -> long long { return draw_of(t) * Bd + rd0_of(t); };

  // weights, bias (zero-padded to P8 columns) of one draw
  const int w_rows = a.flat_rows > 0 ? a.flat_rows : H;
  auto stage_weights = [&](long long s) {
    const float* Ws = a.flat_rows > 0 ? a.W + s * a.flat_stride : a.W + s * (long long)(H * P);
    const float* bs = a.flat_rows > 0 ? Ws + a.flat_rows * P : a.bias + s * (long long)P;
    for (int i = tid; i < H * PW; i += T) {
      const int k = i / PW, n = i % PW;
      sW[i] = (n < P && k < w_rows) ? __ldg(Ws + k * P + n) : 0.0f;
    }
    for (int i = tid; i < P8; i += T) sB[i] = (i < P) ? __ldg(bs + i) : 0.0f;
  };
  // per-warp partial sums of dW / db -> global (one atomic per entry per CTA and draw), then cleared
  auto flush_grads = [&](long long s) {
    float* dWs = a.flat_rows > 0 ? a.dW + s * a.flat_stride : a.dW + s * (long long)(H * P);
    float* dbs = a.flat_rows > 0 ? dWs + a.flat_rows * P : a.dbias + s * (long long)P;
    for (int i = tid; i < H * P8; i += T) {
      const int k = i / P8, n = i % P8;
      float v = 0.0f;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        v += sAcc[w * (H * P8) + i];
        sAcc[w * (H * P8) + i] = 0.0f;
      }
      if (n < P && k < w_rows) atomicAdd(dWs + k * P + n, v);
    }
    for (int n = tid; n < P8; n += T) {
      float v = 0.0f;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        v += sAccB[w * P8 + n];
        sAccB[w * P8 + n] = 0.0f;
      }
      if (n < P) atomicAdd(dbs + n, v);
    }
  };
  if constexpr (BWD) {
    for (int i = tid; i < G::kAcc + G::kAccB; i += T) sAcc[i] = 0.0f;
  }
  Head::template stage<BWD>(a, sHead, tid, T);   // (made visible by the barrier in front of the first tile)
  long long s_cur = -1;   // draw whose weights are staged

  // Every warp owns rows [32 warp, 32 warp + 32) of each tile end to end (h load, the three GEMMs,
  // the per-row flow chain, the dh store), so the tile loop needs only warp-level barriers.
  // async h rows of this warp: 16-byte chunks into rows of stride HS; rows past the draw's end are zeroed
  auto load_h = [&](int buf, long long t) {
    constexpr int CPR = H / 4;  // chunks per row
    float* dst = hT + buf * G::kH + (warp * 32) * HS;
    const long long rd = rd0_of(t) + warp * 32, row0 = fold0_of(t) + warp * 32;
    for (int q = lane; q < 32 * CPR; q += 32) {
      const int r = q / CPR, c = q % CPR;
      float* d = dst + r * HS + 4 * c;
      if (rd + r < Bd) cp_async16(smem_u32(d), a.h + (row0 + r) * H + 4 * c);
      else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };

  if (tile < tile_end) load_h(0, tile);
  cp_async_commit();

  float y_nxt[D];
  float g_nxt = 1.0f;
  // y is indexed by the SAMPLE (row inside the draw), the cotangent by the folded row
  auto fetch_event = [&](long long t) {
    const long long rd = rd0_of(t) + tid;
    if (t < tile_end && rd < Bd) {
      load_event<D>(a.y, a.y_broadcast ? 0 : rd, y_nxt);
      if constexpr (BWD) { if (a.g_logp) g_nxt = __ldg(a.g_logp + fold0_of(t) + tid); }
    }
  };
#pragma unroll
  for (int i = 0; i < D; ++i) y_nxt[i] = 0.0f;
  fetch_event(tile);
  double lsum = 0.0;
  int buf = 0;

  for (; tile < tile_end; tile += tstep) {
    if (draw_of(tile) != s_cur) {   // (CTA-uniform) first tile, or the next draw's weights
      __syncthreads();              // every warp is done with the old weights and has added its last partial sums
      if constexpr (BWD) { if (s_cur >= 0) flush_grads(s_cur); }
      s_cur = draw_of(tile);
      stage_weights(s_cur);
      __syncthreads();
    }
    const long long rd = rd0_of(tile) + tid;      // this thread's row inside the draw
    const long long r = fold0_of(tile) + tid;     // ... and in the folded batch
    float z[D];
#pragma unroll
    for (int i = 0; i < D; ++i) z[i] = y_nxt[i];
    if (a.xf.flags) xform_event<D>(a.xf, r, z);
    const float g_cur = g_nxt;
    {
      const long long nxt = tile + tstep;
      if (nxt < tile_end) load_h(buf ^ 1, nxt);   // the other buffer was drained at the end of the last iteration
      cp_async_commit();
      fetch_event(nxt);
      cp_async_wait<1>();
      __syncwarp();
    }
    float* hcur = hT + buf * G::kH;

    // ---- GEMM 1: t[32 rows of this warp][P8] = h W + b.  Both m-tiles and all n-tiles are in flight and
    // the products are issued term-major, so consecutive HMMAs never depend on each other, and every
    // W fragment is loaded and split once for both m-tiles.
    {
      unsigned a0[2][KH][4], a1[2][KH][4], a2[2][KH][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int r0 = warp * 32 + mt * 16;
#pragma unroll
        for (int ks = 0; ks < KH; ++ks) {
          split3_tf32(hcur[(r0 + g) * HS + 8 * ks + tig], a0[mt][ks][0], a1[mt][ks][0], a2[mt][ks][0]);
          split3_tf32(hcur[(r0 + g + 8) * HS + 8 * ks + tig], a0[mt][ks][1], a1[mt][ks][1], a2[mt][ks][1]);
          split3_tf32(hcur[(r0 + g) * HS + 8 * ks + tig + 4], a0[mt][ks][2], a1[mt][ks][2], a2[mt][ks][2]);
          split3_tf32(hcur[(r0 + g + 8) * HS + 8 * ks + tig + 4], a0[mt][ks][3], a1[mt][ks][3], a2[mt][ks][3]);
        }
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        float c[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int ks = 0; ks < KH; ++ks) {
          unsigned b0[2], b1[2], b2[2];
          split3_tf32(sW[(8 * ks + tig) * PW + 8 * nt + g], b0[0], b1[0], b2[0]);
          split3_tf32(sW[(8 * ks + tig + 4) * PW + 8 * nt + g], b0[1], b1[1], b2[1]);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) mma_tf32(c[mt], a2[mt][ks], b0);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) mma_tf32(c[mt], a0[mt][ks], b2);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) mma_tf32(c[mt], a1[mt][ks], b1);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) mma_tf32(c[mt], a1[mt][ks], b0);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) mma_tf32(c[mt], a0[mt][ks], b1);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) mma_tf32(c[mt], a0[mt][ks], b0);
        }
        const int c0 = 8 * nt + 2 * tig;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int r0 = warp * 32 + mt * 16;
          float* d0 = tT + (r0 + g) * S + c0;
          float* d1 = tT + (r0 + g + 8) * S + c0;
          if (c0 < P) { d0[0] = c[mt][0] + sB[c0]; d1[0] = c[mt][2] + sB[c0]; }
          if (c0 + 1 < P) { d0[1] = c[mt][1] + sB[c0 + 1]; d1[1] = c[mt][3] + sB[c0 + 1]; }
        }
      }
    }
    __syncwarp();

    // ---- per-row head (flow chain or mixture), dt written in place over t
    float* row = tT + tid * S;
    const bool valid = rd < Bd;
    if (valid || Head::kAllLanes) {
      const float lpn = Head::template run<BWD, M>(row, z, (BWD && valid) ? a.g_scale * g_cur : 0.0f, valid, sHead);
      if (valid) {
        const float lp = xform_out<M>(a.xf, lpn);
        a.logp[r] = lp;
        lsum += (double)lp;
      }
    }
    if constexpr (BWD) {
      if (!valid) {
#pragma unroll
        for (int j = 0; j < P; ++j) row[j] = 0.0f;   // rows past the end contribute nothing to dW / db
      }
    }

    if constexpr (BWD) {
      __syncwarp();
      // ---- GEMM 3: dW_warp[H][P8] += h_warp^T dt_warp  (K = this warp's 32 rows), db via a ones operand
      {
        const int R0 = warp * 32;
        float* acc = sAcc + warp * (H * P8);
        float* accb = sAccB + warp * P8;
        const unsigned one[4] = {0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u};
#pragma unroll
        for (int mt = 0; mt < H / 16; ++mt) {
          // A = h^T fragments of this warp's 32 rows, split once and reused by every n-tile
          unsigned ah[4][4], al[4][4];
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const float* hp = hcur + (R0 + 8 * ks + tig) * HS + 16 * mt + g;
            split_tf32(hp[0], ah[ks][0], al[ks][0]);
            split_tf32(hp[8], ah[ks][1], al[ks][1]);
            split_tf32(hp[4 * HS], ah[ks][2], al[ks][2]);
            split_tf32(hp[4 * HS + 8], ah[ks][3], al[ks][3]);
          }
          // n-tiles in groups of at most NG, so that the accumulators of a wide head (MDN: 13 n-tiles) stay in
          // registers; the A fragments above are reused by every group
          constexpr int NG = NT <= 7 ? NT : (NT + 1) / 2 <= 7 ? (NT + 1) / 2 : 7;
#pragma unroll
          for (int n0 = 0; n0 < NT; n0 += NG) {
            float c[NG][4], cb[NG][4];
#pragma unroll
            for (int nt = 0; nt < NG; ++nt) {
              c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.0f;
              cb[nt][0] = cb[nt][1] = cb[nt][2] = cb[nt][3] = 0.0f;
            }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              unsigned bh[NG][2], bl[NG][2];
#pragma unroll
              for (int nt = 0; nt < NG; ++nt) {
                if (n0 + nt < NT) {
                  const int col = 8 * (n0 + nt) + g;
                  split_tf32((col < P) ? tT[(R0 + 8 * ks + tig) * S + col] : 0.0f, bh[nt][0], bl[nt][0]);
                  split_tf32((col < P) ? tT[(R0 + 8 * ks + tig + 4) * S + col] : 0.0f, bh[nt][1], bl[nt][1]);
                }
              }
#pragma unroll
              for (int nt = 0; nt < NG; ++nt) if (n0 + nt < NT) mma_tf32(c[nt], al[ks], bh[nt]);
#pragma unroll
              for (int nt = 0; nt < NG; ++nt) if (n0 + nt < NT) mma_tf32(c[nt], ah[ks], bl[nt]);
#pragma unroll
              for (int nt = 0; nt < NG; ++nt) if (n0 + nt < NT) mma_tf32(c[nt], ah[ks], bh[nt]);
              if (mt == 0) {  // bias gradient: ones^T dt
#pragma unroll
                for (int nt = 0; nt < NG; ++nt) if (n0 + nt < NT) mma_tf32(cb[nt], one, bl[nt]);
#pragma unroll
                for (int nt = 0; nt < NG; ++nt) if (n0 + nt < NT) mma_tf32(cb[nt], one, bh[nt]);
              }
            }
#pragma unroll
            for (int nt = 0; nt < NG; ++nt) {
              if (n0 + nt < NT) {
                if (mt == 0 && g == 0) {
                  accb[8 * (n0 + nt) + 2 * tig] += cb[nt][0];
                  accb[8 * (n0 + nt) + 2 * tig + 1] += cb[nt][1];
                }
                float* p0 = acc + (16 * mt + g) * P8 + 8 * (n0 + nt) + 2 * tig;
                float* p1 = p0 + 8 * P8;
                p0[0] += c[nt][0]; p0[1] += c[nt][1];
                p1[0] += c[nt][2]; p1[1] += c[nt][3];
              }
            }
          }
        }
      }
      __syncwarp();
      // ---- GEMM 2: dh[32 rows][H] = dt W^T  -> into the (now consumed) h rows of this warp
      {
        constexpr int NH = H / 8;
        float c[2][NH][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < NH; ++nt) { c[mt][nt][0] = c[mt][nt][1] = c[mt][nt][2] = c[mt][nt][3] = 0.0f; }
#pragma unroll
        for (int ks = 0; ks < NT; ++ks) {
          const int k0 = 8 * ks + tig, k1 = k0 + 4;
          unsigned ah[2][4], al[2][4], bh[NH][2], bl[NH][2];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const int r0 = warp * 32 + mt * 16;
            split_tf32((k0 < P) ? tT[(r0 + g) * S + k0] : 0.0f, ah[mt][0], al[mt][0]);
            split_tf32((k0 < P) ? tT[(r0 + g + 8) * S + k0] : 0.0f, ah[mt][1], al[mt][1]);
            split_tf32((k1 < P) ? tT[(r0 + g) * S + k1] : 0.0f, ah[mt][2], al[mt][2]);
            split_tf32((k1 < P) ? tT[(r0 + g + 8) * S + k1] : 0.0f, ah[mt][3], al[mt][3]);
          }
#pragma unroll
          for (int nt = 0; nt < NH; ++nt) {
            split_tf32(sW[(8 * nt + g) * PW + k0], bh[nt][0], bl[nt][0]);
            split_tf32(sW[(8 * nt + g) * PW + k1], bh[nt][1], bl[nt][1]);
          }
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < NH; ++nt) mma_tf32(c[mt][nt], al[mt], bh[nt]);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < NH; ++nt) mma_tf32(c[mt][nt], ah[mt], bl[nt]);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < NH; ++nt) mma_tf32(c[mt][nt], ah[mt], bh[nt]);
        }
        __syncwarp();  // every lane of the warp is done reading this warp's h rows (GEMM 3)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int r0 = warp * 32 + mt * 16;
#pragma unroll
          for (int nt = 0; nt < NH; ++nt) {
            float* d0 = hcur + (r0 + g) * HS + 8 * nt + 2 * tig;
            float* d1 = hcur + (r0 + g + 8) * HS + 8 * nt + 2 * tig;
            *reinterpret_cast<float2*>(d0) = make_float2(c[mt][nt][0], c[mt][nt][1]);
            *reinterpret_cast<float2*>(d1) = make_float2(c[mt][nt][2], c[mt][nt][3]);
          }
        }
      }
      __syncwarp();
      // ---- dh rows of this warp -> global, coalesced 16-byte streaming stores
      {
        constexpr int CPR = H / 4;
        const long long row0 = fold0_of(tile) + warp * 32, rdw = rd0_of(tile) + warp * 32;
        const float* src = hcur + (warp * 32) * HS;
        for (int q = lane; q < 32 * CPR; q += 32) {
          const int rr = q / CPR, cc = q % CPR;
          if (rdw + rr < Bd)
            st_stream_f4(a.dh + (row0 + rr) * H + 4 * cc, *reinterpret_cast<const float4*>(src + rr * HS + 4 * cc));
        }
      }
    }
    __syncwarp();
    buf ^= 1;
  }
  cp_async_wait<0>();

  if (a.logp_sum) {
    const double sblk = block_sum<T>(lsum, red);
    if (tid == 0) atomicAdd(a.logp_sum, sblk);
  }
  if constexpr (BWD) {
    __syncthreads();
    if (s_cur >= 0) flush_grads(s_cur);
  }
  Head::template finish<BWD>(a, sHead, tid, T);   // (after the barrier above; forward-only heads have nothing to flush)
}

// the flow-chain instance under its historical name (ahead-of-time instances and the runtime specialiser use it)
template <class Spec, int H, bool BWD, class M, int T>
NFN_DEVI void dense_chain_body(const DenseArgs& a) {
  dense_head_body<ChainHead<Spec>, H, BWD, M, T>(a);
}

template <class Spec, int H, bool BWD, class M, int T, int MINB>
__global__ void __launch_bounds__(T, MINB) dense_chain_kernel(const DenseArgs a) {
  dense_chain_body<Spec, H, BWD, M, T>(a);
}

template <int K, int D, int H, bool BWD, class M, int T, int MINB>
__global__ void __launch_bounds__(T, MINB) dense_mdn_kernel(const DenseArgs a) {
  dense_head_body<MdnHead<K, D>, H, BWD, M, T>(a);
}

template <int MC, int D, int H, bool BWD, class M, int T, int MINB>
__global__ void __launch_bounds__(T, MINB) dense_kmn_kernel(const DenseArgs a) {
  dense_head_body<KmnHead<MC, D>, H, BWD, M, T>(a);
}

}  // namespace nfn
