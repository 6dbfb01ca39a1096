// nfn_mixture.cu -- fused Gaussian-mixture logsumexp heads (sm_100a).
//
//   MDN : tfd.Mixture of K diagonal Gaussians parameterised per sample
//         (reference estimators/DistributionLayers.py:196-212), row layout
//         [ (mu_k(d), sigma_raw_k(d))_{k<K} | logits(K) ], sigma = softplus(0.05 raw + c0).
//   KMN : tfd.MixtureSameFamily over M fixed centres with shared isotropic bandwidths
//         (estimators/DistributionLayers.py:118-133), row = logits(M).
//   logmeanexp over S posterior draws (estimators/BayesianNNEstimator.py:65-76).
//
// Same data movement as the flow chain: one thread per sample, T-row tiles staged with
// coalesced 16-byte cp.async into a double-buffered, bank-conflict-free padded smem tile,
// gradients written in place and streamed back with coalesced 16-byte stores.  The row
// width is a runtime value here (K is not a template parameter); D and the row vector
// width V are.  Forward is an online logsumexp (one EX2 per component); the reverse sweep
// reuses sigma (written over sigma_raw by the forward pass) so softplus is not recomputed.
#include <unordered_map>

#include "nfn_common.h"
#include "nfn_mixture_row.cuh"

namespace nfn {

constexpr int kMixT = 128;

// ------------------------------------------------------------------ runtime-width tile io
// Same chunk geometry as TileIO (nfn_chain_kernel.cuh) with the row width a runtime value.
struct RtTile {
  int P;      // floats per row in global
  int S;      // floats per row in smem (P, or P + 4 when P % 4 == 0 and P/4 is even)
  int P4;     // chunks per row (padded layout only)
  int dr, dc; // kMixT / P4, kMixT % P4: (row, chunk) step of a thread between its chunks
};

NFN_DEVI int rt_index(const RtTile& g, int e) {
  return g.S == g.P ? e : e + (e / g.P) * (g.S - g.P);
}

NFN_DEVI void rt_load_async(const RtTile& g, unsigned smem, const float* __restrict__ src0, long long row0,
                            long long B) {
  const float* src = src0 + row0 * g.P;
  const int chunks = (kMixT * g.P) / 4;
  if (B - row0 >= kMixT) {
    if (g.S == g.P) {
      for (int q = threadIdx.x; q < chunks; q += kMixT) cp_async16(smem + 16u * q, src + 4 * q);
    } else {
      int row = (int)threadIdx.x / g.P4, c = (int)threadIdx.x % g.P4;
      for (int q = threadIdx.x; q < chunks; q += kMixT) {
        cp_async16(smem + 4u * (row * g.S + 4 * c), src + 4 * q);
        row += g.dr;
        c += g.dc;
        if (c >= g.P4) { c -= g.P4; ++row; }
      }
    }
  } else {
    const int remain = (int)(B - row0) * g.P;
    for (int q = threadIdx.x; q < chunks; q += kMixT) {
      const int e = q * 4;
      if (e + 4 <= remain) {
        cp_async16(smem + 4u * rt_index(g, e), src + e);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (e + j < remain) cp_async4(smem + 4u * rt_index(g, e + j), src + e + j);
      }
    }
  }
}

NFN_DEVI void rt_store(const RtTile& g, unsigned smem, float* __restrict__ dst0, long long row0, long long B) {
  float* dst = dst0 + row0 * g.P;
  const int chunks = (kMixT * g.P) / 4;
  if (B - row0 >= kMixT) {
    if (g.S == g.P) {
      for (int q = threadIdx.x; q < chunks; q += kMixT) st_stream_f4(dst + 4 * q, lds_f4(smem + 16u * q));
    } else {
      int row = (int)threadIdx.x / g.P4, c = (int)threadIdx.x % g.P4;
      for (int q = threadIdx.x; q < chunks; q += kMixT) {
        st_stream_f4(dst + 4 * q, lds_f4(smem + 4u * (row * g.S + 4 * c)));
        row += g.dr;
        c += g.dc;
        if (c >= g.P4) { c -= g.P4; ++row; }
      }
    }
  } else {
    const int remain = (int)(B - row0) * g.P;
    for (int q = threadIdx.x; q < chunks; q += kMixT) {
      const int e = q * 4;
      if (e + 4 <= remain) {
        st_stream_f4(dst + e, lds_f4(smem + 4u * rt_index(g, e)));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (e + j < remain) {
            float v;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem + 4u * rt_index(g, e + j)));
            dst[e + j] = v;
          }
      }
    }
  }
}

// Tile pipeline shared by the two heads: nb == 2 double-buffers inside the CTA, nb == 1 keeps
// one buffer per CTA and relies on the other resident CTAs for overlap (wide rows).
struct RtPipe {
  unsigned base, buf_bytes;
  int nb;
  int slot = 0;
  NFN_DEVI void prologue(const RtTile& g, const float* t, long long tile, long long ntiles, long long B) {
    if (nb == 2) {
      if (tile < ntiles) rt_load_async(g, base, t, tile * kMixT, B);
      cp_async_commit();
    }
  }
  // makes the current tile resident; returns its smem byte address
  NFN_DEVI unsigned acquire(const RtTile& g, const float* t, long long tile, long long ntiles, long long B) {
    if (nb == 2) {
      const long long nxt = tile + gridDim.x;
      if (nxt < ntiles) rt_load_async(g, base + (unsigned)(slot ^ 1) * buf_bytes, t, nxt * kMixT, B);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      rt_load_async(g, base, t, tile * kMixT, B);
      cp_async_commit();
      cp_async_wait<0>();
    }
    __syncthreads();
    return base + (unsigned)slot * buf_bytes;
  }
  NFN_DEVI void advance() {
    if (nb == 2) slot ^= 1;
  }
};

// y (and the upstream cotangent) of the next tile are fetched one iteration ahead so their
// DRAM latency hides behind the current tile's arithmetic
template <int D, bool BWD>
struct EventPrefetch {
  float y_nxt[D];
  float g_nxt = 1.0f;
  NFN_DEVI void fetch(const MixArgs& a, long long r) {
    if (r < a.B) {
      load_event<D>(a.y, a.y_broadcast ? 0 : r, y_nxt);
      if constexpr (BWD) { if (a.g_logp) g_nxt = __ldg(a.g_logp + r); }
    }
  }
  NFN_DEVI void first(const MixArgs& a, long long tile, long long ntiles) {
#pragma unroll
    for (int i = 0; i < D; ++i) y_nxt[i] = 0.0f;
    if (tile < ntiles) fetch(a, tile * kMixT + threadIdx.x);
  }
  NFN_DEVI void rotate(const MixArgs& a, long long tile, float (&y)[D], float& g) {
#pragma unroll
    for (int i = 0; i < D; ++i) y[i] = y_nxt[i];
    if (a.xf.flags) xform_event<D>(a.xf, tile * kMixT + threadIdx.x, y);
    g = g_nxt;
    fetch(a, (tile + gridDim.x) * kMixT + threadIdx.x);
  }
};

// ------------------------------------------------------------------ MDN
// V4: rows are 16-byte aligned in smem (P % 4 == 0) so the (mu, sigma_raw) block of a
// component, 2*D floats at offset k*2*D, can be read with the widest aligned vectors.
// LG: logits are read / their gradients written in groups of LG (4 when V4 and K % 4 == 0:
// a scalar LDS at row stride S = 4*odd is 4-way bank conflicted, a 128-bit one is not).
// All log-densities are carried in log2 units (one EX2 / LG2 per use, no rescaling multiply).
template <int D, bool V4, int LG, bool BWD, class M>
__global__ void __launch_bounds__(kMixT, BWD ? 4 : 5) mdn_kernel(const MixArgs a, const RtTile g, const int nb) {
  extern __shared__ __align__(16) float smem[];
  __shared__ double red[kMixT / 32];
  const int K = a.K;
  const long long ntiles = (a.B + kMixT - 1) / kMixT;
  double lsum = 0.0;
  const int tile_floats = kMixT * g.S;
  RtPipe pipe;
  pipe.base = smem_u32(smem);
  pipe.buf_bytes = (unsigned)tile_floats * 4u;
  pipe.nb = nb;

  long long tile = blockIdx.x;
  pipe.prologue(g, a.t, tile, ntiles, a.B);
  EventPrefetch<D, BWD> pf;
  pf.first(a, tile, ntiles);

  for (; tile < ntiles; tile += gridDim.x) {
    float y[D];
    float g_cur;
    pf.rotate(a, tile, y, g_cur);
    const unsigned buf_addr = pipe.acquire(g, a.t, tile, ntiles, a.B);
    float* buf = smem + (size_t)pipe.slot * tile_floats;

    const long long r = tile * kMixT + threadIdx.x;
    if (r < a.B) {
      float* row = buf + threadIdx.x * g.S;
      float dy[D];
#pragma unroll
      for (int i = 0; i < D; ++i) dy[i] = 0.0f;
      const float logp = mdn_row<D, V4, LG, BWD, M>(row, K, y, BWD ? a.g_scale * g_cur : 0.0f, dy);
      const float lpo = xform_out<M>(a.xf, logp);
      a.logp[r] = lpo;
      lsum += (double)lpo;
      if constexpr (BWD) {
        if (a.dy) store_event<D>(a.dy, r, dy);
      }
    }
    __syncthreads();
    if constexpr (BWD) {
      rt_store(g, buf_addr, a.dt, tile * kMixT, a.B);
      __syncthreads();
    }
    pipe.advance();
  }
  cp_async_wait<0>();
  if (a.logp_sum) {
    const double sblk = block_sum<kMixT>(lsum, red);
    if (threadIdx.x == 0) atomicAdd(a.logp_sum, sblk);
  }
}

// ------------------------------------------------------------------ KMN
// smem: [2 buffers of T x S logits] [locs M x D] [coef M: -0.5 / s^2] [lognorm M: -D log|s|]
// LG: logits read / gradients written in groups of LG (4 when M % 4 == 0, see mdn_kernel).
template <int D, int LG, bool BWD, class M>
__global__ void __launch_bounds__(kMixT, BWD ? 4 : 5) kmn_kernel(const MixArgs a, const RtTile g, const int nb) {
  extern __shared__ __align__(16) float smem[];
  __shared__ double red[kMixT / 32];
  constexpr int NW = kMixT / 32;
  const int K = a.K;
  const int tile_floats = kMixT * g.S;
  float* s_loc = smem + nb * (size_t)tile_floats;
  float* s_coef = s_loc + K * D;   // -0.5 log2e / s^2  (log2 units)
  float* s_lnorm = s_coef + K;     // -D log2|s|
  float* s_dsc = s_lnorm + K;      // [NW][K] per-warp sums of d logp / d scale (BWD): no atomics
  for (int i = threadIdx.x; i < K * D; i += kMixT) s_loc[i] = __ldg(a.locs + i);
  for (int i = threadIdx.x; i < K; i += kMixT) {
    const float sc = __ldg(a.scales + i);
    s_coef[i] = -0.5f * kLog2e / (sc * sc);
    s_lnorm[i] = -(float)D * log2f(fabsf(sc));
  }
  if constexpr (BWD) {
    for (int i = threadIdx.x; i < NW * K; i += kMixT) s_dsc[i] = 0.0f;
  }
  const long long ntiles = (a.B + kMixT - 1) / kMixT;
  double lsum = 0.0;
  RtPipe pipe;
  pipe.base = smem_u32(smem);
  pipe.buf_bytes = (unsigned)tile_floats * 4u;
  pipe.nb = nb;

  long long tile = blockIdx.x;
  pipe.prologue(g, a.t, tile, ntiles, a.B);
  EventPrefetch<D, BWD> pf;
  pf.first(a, tile, ntiles);
  float* my_dsc = s_dsc + (threadIdx.x >> 5) * K;

  for (; tile < ntiles; tile += gridDim.x) {
    float y[D];
    float g_cur;
    pf.rotate(a, tile, y, g_cur);
    const unsigned buf_addr = pipe.acquire(g, a.t, tile, ntiles, a.B);
    float* buf = smem + (size_t)pipe.slot * tile_floats;

    const long long r = tile * kMixT + threadIdx.x;
    const bool valid = r < a.B;
    float* row = buf + threadIdx.x * g.S;
    float dy[D];
#pragma unroll
    for (int i = 0; i < D; ++i) dy[i] = 0.0f;
    const float logp = kmn_row<D, LG, BWD, M>(row, K, y, (BWD && valid) ? a.g_scale * g_cur : 0.0f, valid, s_loc, s_coef,
                                              s_lnorm, (BWD && a.dscales) ? my_dsc : nullptr, dy);
    if (valid) {
      const float lpo = xform_out<M>(a.xf, logp);
      a.logp[r] = lpo;
      lsum += (double)lpo;
      if constexpr (BWD) {
        if (a.dy) store_event<D>(a.dy, r, dy);
      }
    }
    __syncthreads();
    if constexpr (BWD) {
      rt_store(g, buf_addr, a.dt, tile * kMixT, a.B);
      __syncthreads();
    }
    pipe.advance();
  }
  cp_async_wait<0>();
  if constexpr (BWD) {
    if (a.dscales) {
      __syncthreads();
      for (int i = threadIdx.x; i < K; i += kMixT) {
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < NW; ++w) v += s_dsc[w * K + i];
        atomicAdd(a.dscales + i, v / __ldg(a.scales + i));
      }
    }
  }
  if (a.logp_sum) {
    const double sblk = block_sum<kMixT>(lsum, red);
    if (threadIdx.x == 0) atomicAdd(a.logp_sum, sblk);
  }
}

// ------------------------------------------------------------------ launchers
static RtTile make_tile(int P) {
  RtTile g;
  g.P = P;
  g.S = row_stride(P);
  g.P4 = (g.S != P) ? P / 4 : 1;
  g.dr = kMixT / g.P4;
  g.dc = kMixT % g.P4;
  return g;
}

// one buffer per CTA unless two buffers still leave >= 4 CTAs per SM
static int pick_nb(size_t tile_bytes, size_t extra) {
  return (4 * (2 * tile_bytes + extra + 1024) <= (size_t)227 * 1024) ? 2 : 1;
}

template <class Kern>
static int launch_tiled(Kern kern, const MixArgs& a, const RtTile& g, size_t smem, int nb, const char* name,
                        cudaStream_t st) {
  const DeviceInfo& di = device_info();
  if (smem + 2048 > (size_t)di.smem_optin)
    return set_error(NFN_ERR_UNSUPPORTED, "%s: row of %d floats needs %zu bytes of shared memory (> %d)", name,
                     g.P, smem, di.smem_optin - 2048);
  // once per (kernel, device): opt in to the device's full dynamic shared memory, so that no later launch with a
  // different component count has to touch the attribute again (a per-launch cudaFuncSetAttribute races between
  // host threads using different K); the occupancy answer is memoised per (kernel, device, bytes)
  struct Memo {
    int device = -1;
    std::unordered_map<size_t, int> occ;
  };
  static thread_local std::unordered_map<const void*, Memo> memo;
  Memo& m = memo[(const void*)kern];
  cudaError_t e;
  if (m.device != di.device) {
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return cuda_error(e, name);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, di.smem_optin - (int)fa.sharedSizeBytes);
    if (e != cudaSuccess) return cuda_error(e, name);
    m.device = di.device;
    m.occ.clear();
  }
  auto it = m.occ.find(smem);
  if (it == m.occ.end()) {
    int o = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kMixT, smem);
    if (e != cudaSuccess) return cuda_error(e, name);
    it = m.occ.emplace(smem, o < 1 ? 1 : o).first;
  }
  const int occ = it->second;
  const long long ntiles = (a.B + kMixT - 1) / kMixT;
  long long grid = (long long)di.sm_count * occ;
  if (grid > ntiles) grid = ntiles;
  kern<<<(unsigned)grid, kMixT, smem, st>>>(a, g, nb);
  count_launch();
  return cuda_error(cudaGetLastError(), name);
}

template <int D, bool V4, int LG>
static int launch_mdn_dvl(bool bwd, const MixArgs& a, const RtTile& g, size_t smem, int nb, cudaStream_t st) {
  if (math_mode() == 0) {
    return bwd ? launch_tiled(mdn_kernel<D, V4, LG, true, MathFast>, a, g, smem, nb, "mdn_kernel", st)
               : launch_tiled(mdn_kernel<D, V4, LG, false, MathFast>, a, g, smem, nb, "mdn_kernel", st);
  }
  return bwd ? launch_tiled(mdn_kernel<D, V4, LG, true, MathAccurate>, a, g, smem, nb, "mdn_kernel", st)
             : launch_tiled(mdn_kernel<D, V4, LG, false, MathAccurate>, a, g, smem, nb, "mdn_kernel", st);
}

template <int D>
static int launch_mdn_d(bool bwd, const MixArgs& a, cudaStream_t st) {
  const int P = 2 * a.K * D + a.K;
  const RtTile g = make_tile(P);
  const size_t tile = (size_t)kMixT * g.S * sizeof(float);
  const int nb = pick_nb(tile, 0);
  const size_t smem = tile * nb;
  int rc;
  if (P % 4 == 0 && a.K % 4 == 0) rc = launch_mdn_dvl<D, true, 4>(bwd, a, g, smem, nb, st);
  else if (P % 4 == 0) rc = launch_mdn_dvl<D, true, 1>(bwd, a, g, smem, nb, st);
  else rc = launch_mdn_dvl<D, false, 1>(bwd, a, g, smem, nb, st);
  if (rc == NFN_OK && bwd && a.dt_colsum) rc = launch_colsum(a.dt, a.B, P, a.dt_colsum, st);
  return rc;
}

int launch_mdn(int d, bool bwd, const MixArgs& a, cudaStream_t st) {
  switch (d) {
    case 1: return launch_mdn_d<1>(bwd, a, st);
    case 2: return launch_mdn_d<2>(bwd, a, st);
    case 3: return launch_mdn_d<3>(bwd, a, st);
    case 4: return launch_mdn_d<4>(bwd, a, st);
    case 5: return launch_mdn_d<5>(bwd, a, st);
    case 6: return launch_mdn_d<6>(bwd, a, st);
    case 7: return launch_mdn_d<7>(bwd, a, st);
    case 8: return launch_mdn_d<8>(bwd, a, st);
  }
  return set_error(NFN_ERR_DESC, "n_dims=%d", d);
}

template <int D, int LG>
static int launch_kmn_dl(bool bwd, const MixArgs& a, const RtTile& g, size_t smem, int nb, cudaStream_t st) {
  if (math_mode() == 0) {
    return bwd ? launch_tiled(kmn_kernel<D, LG, true, MathFast>, a, g, smem, nb, "kmn_kernel", st)
               : launch_tiled(kmn_kernel<D, LG, false, MathFast>, a, g, smem, nb, "kmn_kernel", st);
  }
  return bwd ? launch_tiled(kmn_kernel<D, LG, true, MathAccurate>, a, g, smem, nb, "kmn_kernel", st)
             : launch_tiled(kmn_kernel<D, LG, false, MathAccurate>, a, g, smem, nb, "kmn_kernel", st);
}

template <int D>
static int launch_kmn_d(bool bwd, const MixArgs& a, cudaStream_t st) {
  const RtTile g = make_tile(a.K);
  const size_t tile = (size_t)kMixT * g.S * sizeof(float);
  const size_t extra = (size_t)a.K * (D + 2 + kMixT / 32) * sizeof(float);
  const int nb = pick_nb(tile, extra);
  const size_t smem = tile * nb + extra;
  return (a.K % 4 == 0) ? launch_kmn_dl<D, 4>(bwd, a, g, smem, nb, st) : launch_kmn_dl<D, 1>(bwd, a, g, smem, nb, st);
}

int launch_kmn(int d, bool bwd, const MixArgs& a, cudaStream_t st) {
  switch (d) {
    case 1: return launch_kmn_d<1>(bwd, a, st);
    case 2: return launch_kmn_d<2>(bwd, a, st);
    case 3: return launch_kmn_d<3>(bwd, a, st);
    case 4: return launch_kmn_d<4>(bwd, a, st);
    case 5: return launch_kmn_d<5>(bwd, a, st);
    case 6: return launch_kmn_d<6>(bwd, a, st);
    case 7: return launch_kmn_d<7>(bwd, a, st);
    case 8: return launch_kmn_d<8>(bwd, a, st);
  }
  return set_error(NFN_ERR_DESC, "n_dims=%d", d);
}

// ------------------------------------------------------------------ posterior-draw epilogue
// out[b] = logsumexp_s in[s, b] - log S ; thread per b, coalesced across b for every s.
__global__ void __launch_bounds__(256) logmeanexp_kernel(const float* __restrict__ in, long long S, long long B,
                                                         float* __restrict__ out) {
  for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B;
       b += (long long)gridDim.x * blockDim.x) {
    float m = NFN_NEG_INF, s = 0.0f;
    for (long long k = 0; k < S; ++k) lse_push<MathAccurate>(__ldg(in + k * B + b), m, s);
    out[b] = m + logf(s) - logf((float)S);
  }
}

int launch_logmeanexp(const float* in, long long S, long long B, float* out, cudaStream_t st) {
  long long blocks = (B + 255) / 256;
  const long long cap = (long long)device_info().sm_count * 8;
  if (blocks > cap) blocks = cap;
  logmeanexp_kernel<<<(unsigned)blocks, 256, 0, st>>>(in, S, B, out);
  count_launch();
  return cuda_error(cudaGetLastError(), "logmeanexp_kernel");
}

}  // namespace nfn
