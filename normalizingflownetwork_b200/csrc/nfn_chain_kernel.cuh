// nfn_chain_kernel.cuh -- compile-time specialised fused flow-chain kernels (sm_100a).
//
// One thread = one (x, y) pair.  A CTA of T threads owns tiles of T consecutive rows of
// the parameter tensor t[B, P] == one contiguous T*P*4-byte span of HBM:
//
//   global t  --cp.async.cg 16 B, fully coalesced-->  smem tile [T][S]   (double buffered)
//   each thread reads its own row with 128/64/32-bit LDS at compile-time offsets,
//   runs all K flows in registers (z history kept for the reverse sweep), writes
//   d logp / d theta back IN PLACE over the parameters it just consumed,
//   smem tile  --LDS.128 + STG.128 (streaming), fully coalesced-->  global dt
//
// Row stride S (floats) is chosen so the per-thread vector reads are bank-conflict free:
// with V = min(4, largest power of two dividing P) the reads are V-wide and need S/V odd;
// S = P except when V == 4 and P/4 is even, where each row is padded by one 16-byte chunk
// (cp.async places chunks individually, so padding costs nothing).
//
// Replaces, per row: TransformedDistribution.log_prob over Invert(Chain(flows))
// (reference estimators/DistributionLayers.py:245-294) and its tape gradient
// (estimators/BaseEstimator.py:55-59 under Keras fit).
#pragma once
#include "nfn_flows.cuh"

namespace nfn {

// ---------------------------------------------------------------- compile-time chain
template <int D_, bool BASE_, int... F>
struct ChainSpec {
  static constexpr int D = D_;
  static constexpr bool BASE = BASE_;
  static constexpr int K = sizeof...(F);
  static constexpr int KA = K > 0 ? K : 1;
  static constexpr int base_size = BASE ? 2 * D : 0;
  __host__ __device__ static constexpr int type(int k) {
    const int arr[KA + 1] = {F..., -1};
    return arr[k];
  }
  __host__ __device__ static constexpr int size(int k) { return flow_param_size(type(k), D); }
  // the LAST flow owns the first columns after the base block (DistributionLayers.py:267-278)
  __host__ __device__ static constexpr int offset(int k) {
    int off = base_size;
    for (int j = K - 1; j > k; --j) off += flow_param_size(type(j), D);
    return off;
  }
  __host__ __device__ static constexpr int P() {
    int p = base_size;
    for (int j = 0; j < K; ++j) p += flow_param_size(type(j), D);
    return p;
  }
};

__host__ __device__ constexpr int row_vec(int P) { return (P % 4 == 0) ? 4 : ((P % 2 == 0) ? 2 : 1); }
__host__ __device__ constexpr int row_stride(int P) {
  return (P % 4 == 0 && (P / 4) % 2 == 0) ? P + 4 : P;
}

// Launch geometry of a specialised chain kernel: T rows per tile, NB tile buffers per CTA,
// MINB resident CTAs per SM promised to ptxas (sets the register budget).  Defaults come
// from A/B sweeps on B200 (profiles/tuning_r01.md): the fused forward+backward kernel wants
// registers more than warps (2 CTAs x 4 warps with a double-buffered tile, up to 255
// registers), the forward kernel wants one buffer per CTA and up to 4 CTAs per SM.
struct ChainGeometry {
  int T, NB, MINB;
  unsigned smem_bytes;
  int rows = 0;       // rows per tile when it differs from the thread count (0: == T)
  int max_ctas = 0;   // extra cap on resident CTAs per SM (TMEM columns); 0: none
  int force_ctas = 0; // resident CTAs per SM to use instead of the occupancy API's answer; 0: ask the API
};
__host__ __device__ constexpr ChainGeometry chain_geometry(int P, bool bwd) {
  const int T = 128;
  const int S = row_stride(P > 0 ? P : 4);
  const unsigned tile = P > 0 ? (unsigned)(T * S * 4) : 0u;
  // tile buffers that fit one SM (227 KB usable, ~1 KB reserved per CTA)
  const int bufs = tile ? (int)((227u * 1024u - 6u * 1024u) / tile) : 64;
  const int nb = (bwd && bufs >= 4) ? 2 : 1;
  const int by_smem = bufs / nb;
  const int want = bwd ? 2 : 4;
  const int minb = by_smem < 1 ? 1 : (by_smem < want ? by_smem : want);
  return ChainGeometry{T, nb, minb, tile * (unsigned)nb, 0, 0, 0};
}

// In-kernel all-reduce of the fp64 accumulators over NVLink peer memory (nfn_peer.cu).
// Every rank owns one IPC-shared region of 8-byte words laid out as
//   words[2][world][n_values][2]        (2 = step parity; 2 words per fp64 value)
// Low-latency protocol (flag travels WITH the data, so no fences and no separate flag write):
// each word carries 32 bits of the value and the 32-bit step tag; one 8-byte store is atomic.
// The last CTA of a launch pushes this rank's totals into [par][rank] of EVERY peer, then polls
// its own region until every peer's words carry this step's tag and sums them in rank order
// (deterministic).  world == 0 disables the exchange.
constexpr int kMaxPeers = 8;
struct PeerArgs {
  double* base[kMaxPeers];     // peer regions as mapped into this process (own region included)
  double* acc;                 // local accumulators [n_values], self-resetting
  double* out;                 // reduced result [n_values], local
  unsigned* ticket;            // CTA arrival counter, self-resetting
  unsigned long long step;     // launch sequence number of this communicator
  int world, rank, n_values;
  int pad_;
};

struct ChainArgs {
  const float* t;
  const float* y;
  const float* g_logp;
  float* logp;
  float* dt;
  float* dy;
  double* logp_sum;
  double* dt_colsum;
  long long B;
  float g_scale;
  int y_broadcast;
  // forward only: grid_ny > 0 scores the outer product of the B parameter rows with grid_ny events
  // y[grid_ny, d]; logp is [grid_ny, B] (event-major).  The parameter tile is staged once and
  // reused for every event (reference evaluation/visualization/flow_plotting.py:33-53).
  int grid_ny;
  int pad_;
  PeerArgs peer;
};

// ---------------------------------------------------------------- smem span load/store
template <int OFF, int N, int V>
struct Span {
  NFN_DEVI static void load(const float* row, float* out) {
    if constexpr (N <= 0) {
      return;
    } else if constexpr (V >= 4 && OFF % 4 == 0 && N >= 4) {
      const float4 v = *reinterpret_cast<const float4*>(row + OFF);
      out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
      Span<OFF + 4, N - 4, V>::load(row, out + 4);
    } else if constexpr (V >= 2 && OFF % 2 == 0 && N >= 2) {
      const float2 v = *reinterpret_cast<const float2*>(row + OFF);
      out[0] = v.x; out[1] = v.y;
      Span<OFF + 2, N - 2, V>::load(row, out + 2);
    } else {
      out[0] = row[OFF];
      Span<OFF + 1, N - 1, V>::load(row, out + 1);
    }
  }
  NFN_DEVI static void store(float* row, const float* in) {
    if constexpr (N <= 0) {
      return;
    } else if constexpr (V >= 4 && OFF % 4 == 0 && N >= 4) {
      *reinterpret_cast<float4*>(row + OFF) = make_float4(in[0], in[1], in[2], in[3]);
      Span<OFF + 4, N - 4, V>::store(row, in + 4);
    } else if constexpr (V >= 2 && OFF % 2 == 0 && N >= 2) {
      *reinterpret_cast<float2*>(row + OFF) = make_float2(in[0], in[1]);
      Span<OFF + 2, N - 2, V>::store(row, in + 2);
    } else {
      row[OFF] = in[0];
      Span<OFF + 1, N - 1, V>::store(row, in + 1);
    }
  }
};

// ---------------------------------------------------------------- async copy helpers
NFN_DEVI unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
NFN_DEVI void cp_async16(unsigned smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
NFN_DEVI void cp_async4(unsigned smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
NFN_DEVI void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
NFN_DEVI void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

NFN_DEVI float4 lds_f4(unsigned smem_src) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_src));
  return v;
}
NFN_DEVI void st_stream_f4(float* gdst, const float4& v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gdst), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// Tile geometry shared by the chain and mixture kernels: tile of T rows x P floats in
// global (contiguous), T x S floats in smem.  16-byte chunk q of the tile (q = e / 4 for
// flat element e) sits at byte q*16 in global and, in smem, at
//   dense  (S == P): byte q*16
//   padded (S == P+4, P % 4 == 0): row = q / (P/4), c = q % (P/4) -> byte (row*S + 4c)*4.
// Thread tid handles chunks tid, tid+T, ...: (row, c) advance incrementally, no division
// in the loop.  Full tiles (every tile but the last) take a path without bounds checks.
template <int P, int T>
struct TileIO {
  static constexpr int V = row_vec(P);
  static constexpr int S = row_stride(P);
  static constexpr int kChunks = (T * P) / 4;  // T % 4 == 0 -> exact
  static constexpr bool kPadded = (S != P);
  static constexpr int P4 = P / 4;             // chunks per row (padded layout only)
  static constexpr int kIters = (kChunks + T - 1) / T;
  static constexpr bool kExact = (kChunks % T == 0);
  static_assert(T % 4 == 0, "tile must be a whole number of 16-byte chunks");

  // smem float index of flat tile element e (slow path only)
  NFN_DEVI static int smem_index(int e) {
    if constexpr (!kPadded) return e; else return e + (e / P) * (S - P);
  }

  struct Cursor {  // smem byte offset of this thread's current chunk
    int row, c;
    NFN_DEVI void init() {
      if constexpr (kPadded) { row = (int)threadIdx.x / P4; c = (int)threadIdx.x % P4; }
    }
    NFN_DEVI unsigned offset(int i) const {
      if constexpr (kPadded) return (unsigned)((row * S + 4 * c) * 4);
      else return (unsigned)(((int)threadIdx.x + i * T) * 16);
    }
    NFN_DEVI void next() {
      if constexpr (kPadded) {
        row += T / P4;
        c += T % P4;
        if (c >= P4) { c -= P4; ++row; }
      }
    }
  };

  // async global -> smem for the tile starting at row0 (valid rows: min(T, B - row0))
  NFN_DEVI static void load_async(unsigned smem, const float* __restrict__ g, long long row0,
                                  long long B) {
    const float* src = g + row0 * P;
    if (B - row0 >= T) {
      Cursor cur;
      cur.init();
      const float* s = src + 4 * (int)threadIdx.x;
#pragma unroll
      for (int i = 0; i < kIters; ++i) {
        if (kExact || (int)threadIdx.x + i * T < kChunks) cp_async16(smem + cur.offset(i), s + i * (4 * T));
        cur.next();
      }
    } else {
      const int remain = (int)(B - row0) * P;  // floats available from src (< T*P)
      for (int q = threadIdx.x; q < kChunks; q += T) {
        const int e = q * 4;
        if (e + 4 <= remain) {
          cp_async16(smem + 4u * smem_index(e), src + e);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (e + j < remain) cp_async4(smem + 4u * smem_index(e + j), src + e + j);
        }
      }
    }
  }

  // smem -> global, coalesced 16-byte streaming stores
  NFN_DEVI static void store(unsigned smem, float* __restrict__ g, long long row0, long long B) {
    float* dst = g + row0 * P;
    if (B - row0 >= T) {
      Cursor cur;
      cur.init();
      float* d = dst + 4 * (int)threadIdx.x;
#pragma unroll
      for (int i = 0; i < kIters; ++i) {
        if (kExact || (int)threadIdx.x + i * T < kChunks) st_stream_f4(d + i * (4 * T), lds_f4(smem + cur.offset(i)));
        cur.next();
      }
    } else {
      const int remain = (int)(B - row0) * P;
      for (int q = threadIdx.x; q < kChunks; q += T) {
        const int e = q * 4;
        if (e + 4 <= remain) {
          st_stream_f4(dst + e, lds_f4(smem + 4u * smem_index(e)));
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (e + j < remain) {
              float v;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem + 4u * smem_index(e + j)));
              dst[e + j] = v;
            }
        }
      }
    }
  }
};

// ---------------------------------------------------------------- static chain sweeps
// SAVE (fused fwd+bwd kernels): radial flows write their constrained (alpha, beta+1) back
// over (alpha_raw, beta_raw) in the smem row so the reverse sweep skips the softplus.
template <class Spec, class M, int V, bool SAVE, int K0>
struct FwdSweep {
  // flows K0 .. K-1
  NFN_DEVI static void run(float* row, float (&z)[Spec::D], float (&zs)[Spec::KA][Spec::D],
                           LogDetAcc<M>& ld) {
    if constexpr (K0 < Spec::K) {
      constexpr int D = Spec::D;
      constexpr int type = Spec::type(K0);
      constexpr int N = flow_param_size(type, D);
      float th[N];
      Span<Spec::offset(K0), N, V>::load(row, th);
#pragma unroll
      for (int i = 0; i < D; ++i) zs[K0][i] = z[i];
      if constexpr (type == kPlanar) {
        PlanarFlow<D, M>::fwd(th, z, ld);
      } else if constexpr (type == kRadial) {
        if constexpr (SAVE) {
          RadialFlow<D, M>::fwd_save(th, z, ld);
          Span<Spec::offset(K0), 2, V>::store(row, th);
        } else {
          RadialFlow<D, M>::fwd(th, z, ld);
        }
      } else {
        AffineFlow<D, M>::fwd(th, z, ld);
      }
      FwdSweep<Spec, M, V, SAVE, K0 + 1>::run(row, z, zs, ld);
    }
  }
};

template <class Spec, class M, int V, int K0>
struct BwdSweep {
  // flows K0 .. 0 (descending)
  NFN_DEVI static void run(float* row, const float (&zs)[Spec::KA][Spec::D], float (&G)[Spec::D],
                           float cot) {
    if constexpr (K0 >= 0) {
      constexpr int D = Spec::D;
      constexpr int type = Spec::type(K0);
      constexpr int N = flow_param_size(type, D);
      float th[N], gth[N];
      Span<Spec::offset(K0), N, V>::load(row, th);
      if constexpr (type == kPlanar) PlanarFlow<D, M>::bwd(th, zs[K0], G, cot, gth);
      else if constexpr (type == kRadial) RadialFlow<D, M>::bwd_saved(th, zs[K0], G, cot, gth);
      else AffineFlow<D, M>::bwd(th, zs[K0], G, cot, gth);
      Span<Spec::offset(K0), N, V>::store(row, gth);
      BwdSweep<Spec, M, V, K0 - 1>::run(row, zs, G, cot);
    }
  }
};

template <int D>
NFN_DEVI void load_event(const float* __restrict__ y, long long r, float (&z)[D]) {
  if constexpr (D == 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(y) + r);
    z[0] = v.x; z[1] = v.y; z[2] = v.z; z[3] = v.w;
  } else if constexpr (D == 2) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(y) + r);
    z[0] = v.x; z[1] = v.y;
  } else {
#pragma unroll
    for (int i = 0; i < D; ++i) z[i] = __ldg(y + r * D + i);
  }
}

template <int D>
NFN_DEVI void store_event(float* __restrict__ y, long long r, const float (&z)[D]) {
#pragma unroll
  for (int i = 0; i < D; ++i) y[r * D + i] = z[i];
}

// block-wide sum of a double; result valid in thread 0
template <int T>
NFN_DEVI double block_sum(double v, double* scratch /* >= T/32 doubles */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < T / 32; ++i) s += scratch[i];
  }
  return s;
}

// ---------------------------------------------------------------- column sums of a dt tile
// s_col[j] += sum over the tile's valid rows of tile[row][j], without atomics (float
// atomicAdd on shared memory is a CAS spin loop, ruinous under contention).  All T threads
// take part: unit u -> (row chunk c, column group g of V columns) reduces its rows with V-wide
// LDS into s_part[c][.]; after a barrier thread j owns column j and folds the NCH partials.
// Must be called by every thread of the CTA (contains a barrier).
template <int P, int T, int V>
struct ColSum {
  static constexpr int G = P / V;                        // column groups (V divides P)
  static constexpr int NCH = (T / G) > 0 ? (T / G) : 1;  // row chunks reduced in parallel
  static constexpr int R = (T + NCH - 1) / NCH;          // rows per chunk
  static constexpr int kScratch = NCH > 1 ? NCH * P : 1; // <= T * V floats

  template <int S>
  NFN_DEVI static void add_tile(const float* tile, int rows, float* s_col, float* s_part) {
    for (int u = threadIdx.x; u < G * NCH; u += T) {
      const int g = u % G, c = u / G;
      const int r0 = c * R;
      int r1 = r0 + R;
      if (r1 > rows) r1 = rows;
      float acc[V];
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v] = 0.0f;
      const float* p = tile + g * V;
#pragma unroll 4
      for (int r = r0; r < r1; ++r) {
        if constexpr (V == 4) {
          const float4 x = *reinterpret_cast<const float4*>(p + r * S);
          acc[0] += x.x; acc[1] += x.y; acc[2] += x.z; acc[3] += x.w;
        } else if constexpr (V == 2) {
          const float2 x = *reinterpret_cast<const float2*>(p + r * S);
          acc[0] += x.x; acc[1] += x.y;
        } else {
          acc[0] += p[r * S];
        }
      }
#pragma unroll
      for (int v = 0; v < V; ++v) {
        if constexpr (NCH > 1) s_part[c * P + g * V + v] = acc[v];
        else s_col[g * V + v] += acc[v];                 // one unit per column group: no sharing
      }
    }
    if constexpr (NCH > 1) {
      __syncthreads();
      for (int j = threadIdx.x; j < P; j += T) {
        float s = 0.0f;
#pragma unroll
        for (int c = 0; c < NCH; ++c) s += s_part[c * P + j];
        s_col[j] += s;
      }
    }
  }
};

// ---------------------------------------------------------------- fused peer all-reduce
NFN_DEVI unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
NFN_DEVI void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Called by every thread of every CTA after the CTA's accumulator atomics.  The last CTA to
// arrive runs the exchange; a peer that never shows up is abandoned after ~1 s (NaN result)
// so that a failed rank cannot hang this GPU.
template <int T>
NFN_DEVI void peer_allreduce(const PeerArgs& p) {
  __shared__ int s_last;
  __threadfence();  // this CTA's atomics into p.acc are visible device-wide
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(p.ticket, 1u);
    s_last = (t == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int W = p.world, NV = p.n_values, par = (int)(p.step & 1ull);
  const unsigned long long tag = ((p.step + 1ull) & 0xffffffffull) << 32;  // never 0 in the first 2^32 steps
  for (int j = threadIdx.x; j < NV; j += T) {
    // 1. take (and reset) the local total, push (value half | tag) words to every peer
    const unsigned long long bits = atomicExch(reinterpret_cast<unsigned long long*>(p.acc + j), 0ull);
    const unsigned long long w0 = (bits & 0xffffffffull) | tag, w1 = (bits >> 32) | tag;
    const size_t off = ((size_t)(par * W + p.rank) * NV + j) * 2;
    for (int q = 0; q < W; ++q) {
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(p.base[q]) + off;
      st_volatile_u64(dst, w0);
      st_volatile_u64(dst + 1, w1);
    }
    // 2. poll this rank's region for every peer's words of this step; 3. sum in rank order
    const unsigned long long* mine =
        reinterpret_cast<const unsigned long long*>(p.base[p.rank]) + ((size_t)par * W * NV + j) * 2;
    double sum = 0.0;
    bool timeout = false;
    const long long t0 = clock64();
    for (int q = 0; q < W; ++q) {
      const unsigned long long* src = mine + (size_t)q * NV * 2;
      unsigned long long a0, a1;
      while (true) {
        a0 = ld_volatile_u64(src);
        a1 = ld_volatile_u64(src + 1);
        if ((a0 & 0xffffffff00000000ull) == tag && (a1 & 0xffffffff00000000ull) == tag) break;
        if (clock64() - t0 > 2000000000ll) {  // ~1 s at 2 GHz
          timeout = true;
          break;
        }
      }
      sum += __longlong_as_double((long long)((a0 & 0xffffffffull) | (a1 << 32)));
    }
    p.out[j] = timeout ? __longlong_as_double(0x7ff8000000000000ll) : sum;
  }
  if (threadIdx.x == 0) *p.ticket = 0u;
}

// ---------------------------------------------------------------- the kernel
template <class Spec, bool BWD, class M, int T, int NB>
NFN_DEVI void chain_body(const ChainArgs& a) {
  constexpr int D = Spec::D;
  constexpr int P = Spec::P();
  using IO = TileIO<(P > 0 ? P : 4), T>;
  constexpr int V = IO::V;
  constexpr int S = IO::S;
  static_assert(NB >= 1 && NB <= 4, "1..4 tile buffers");

  extern __shared__ __align__(16) float smem[];
  __shared__ double red[T / 32];

  const long long ntiles = (a.B + T - 1) / T;
  // per-CTA column sums of dt (bias gradient of the emitting layer), kept in smem across tiles
  using CS = ColSum<(P > 0 ? P : 4), T, V>;
  __shared__ float s_col[BWD && P > 0 ? P : 1];
  __shared__ float s_part[BWD && P > 0 ? CS::kScratch : 1];
  if constexpr (BWD && P > 0) {
    if (a.dt_colsum) {
      for (int j = threadIdx.x; j < P; j += T) s_col[j] = 0.0f;
    }
  }
  double lsum = 0.0;

  // Programmatic dependent launch (no-ops unless the launch asked for it, see launch_chain): the next
  // launch in the stream may be scheduled onto SMs as this grid's CTAs retire, and this grid touches global
  // memory only after everything before it in the stream has completed and flushed.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  const unsigned smem_base = smem_u32(smem);
  constexpr unsigned kBufBytes = (unsigned)(T * S * sizeof(float));
  long long tile = blockIdx.x;
  // NB-stage pipeline: tiles it+1 .. it+NB-1 are in flight while tile it is computed.
  // NB == 1 keeps one buffer per CTA and relies on the other resident CTAs for overlap
  // (wide rows: more CTAs per SM beat a second buffer).
  if constexpr (P > 0 && NB > 1) {
#pragma unroll
    for (int s = 0; s < NB - 1; ++s) {
      const long long tl = tile + (long long)s * gridDim.x;
      if (tl < ntiles) IO::load_async(smem_base + (unsigned)s * kBufBytes, a.t, tl * T, a.B);
      cp_async_commit();
    }
  }

  // y (and the upstream cotangent) of the NEXT tile are fetched into registers one iteration
  // ahead, so their DRAM latency hides behind the current tile's arithmetic
  float y_nxt[D];
  float g_nxt = 1.0f;
  {
    const long long r0 = tile * T + threadIdx.x;
#pragma unroll
    for (int i = 0; i < D; ++i) y_nxt[i] = 0.0f;
    if (tile < ntiles && r0 < a.B) {
      load_event<D>(a.y, a.y_broadcast ? 0 : r0, y_nxt);
      if constexpr (BWD) { if (a.g_logp) g_nxt = __ldg(a.g_logp + r0); }
    }
  }

  int slot = 0;  // buffer holding the current tile
  for (; tile < ntiles; tile += gridDim.x) {
    float* buf = smem + (size_t)slot * (T * S);
    float z[D];
#pragma unroll
    for (int i = 0; i < D; ++i) z[i] = y_nxt[i];
    const float g_cur = g_nxt;
    {
      const long long rn = (tile + gridDim.x) * T + threadIdx.x;
      if (rn < a.B) {
        load_event<D>(a.y, a.y_broadcast ? 0 : rn, y_nxt);
        if constexpr (BWD) { if (a.g_logp) g_nxt = __ldg(a.g_logp + rn); }
      }
    }
    if constexpr (P > 0) {
      if constexpr (NB > 1) {
        // refill the buffer of the previous tile: every thread left it at the barrier
        // that ended the previous iteration
        const long long nxt = tile + (long long)(NB - 1) * gridDim.x;
        const int ps = (slot == 0) ? NB - 1 : slot - 1;
        if (nxt < ntiles) IO::load_async(smem_base + (unsigned)ps * kBufBytes, a.t, nxt * T, a.B);
        cp_async_commit();
        cp_async_wait<NB - 1>();
      } else {
        IO::load_async(smem_base, a.t, tile * T, a.B);
        cp_async_commit();
        cp_async_wait<0>();
      }
      __syncthreads();
    }

    const long long r = tile * T + threadIdx.x;
    if constexpr (!BWD) {
      if (a.grid_ny > 0) {
        if (r < a.B) {
          float* row = buf + threadIdx.x * S;
          using Base = BaseDist<D, Spec::BASE, M>;
          float bth[Base::NA];
          if constexpr (Spec::BASE) Span<0, 2 * D, V>::load(row, bth);
          for (int j = 0; j < a.grid_ny; ++j) {
            float zg[D];
            load_event<D>(a.y, j, zg);  // same address for the whole warp: one broadcast load
            float zs[Spec::KA][D];
            LogDetAcc<M> ld;
            FwdSweep<Spec, M, V, false, 0>::run(row, zg, zs, ld);
            a.logp[(long long)j * a.B + r] = Base::log_prob(bth, zg) + ld.nat();
          }
        }
        if constexpr (P > 0) __syncthreads();
        slot = (slot + 1 == NB) ? 0 : slot + 1;
        continue;
      }
    }
    if (r < a.B) {
      float* row = buf + threadIdx.x * S;
      float zs[Spec::KA][D];
      LogDetAcc<M> ld;
      FwdSweep<Spec, M, V, BWD, 0>::run(row, z, zs, ld);
      using Base = BaseDist<D, Spec::BASE, M>;
      float bth[Base::NA];
      if constexpr (Spec::BASE) Span<0, 2 * D, V>::load(row, bth);
      // fused kernel: sigma stays in registers (bth) for the reverse sweep
      const float lp = (BWD ? Base::log_prob_save(bth, z) : Base::log_prob(bth, z)) + ld.nat();
      a.logp[r] = lp;
      lsum += (double)lp;
      if constexpr (BWD) {
        const float cot = a.g_scale * g_cur;
        float G[D];
        float gb[Base::NA];
        Base::bwd_saved(bth, z, cot, G, gb);
        if constexpr (Spec::BASE) Span<0, 2 * D, V>::store(row, gb);
        BwdSweep<Spec, M, V, Spec::K - 1>::run(row, zs, G, cot);
        if (a.dy) store_event<D>(a.dy, r, G);
      }
    }

    if constexpr (BWD && P > 0) {
      __syncthreads();
      IO::store(smem_base + (unsigned)slot * kBufBytes, a.dt, tile * T, a.B);
      if (a.dt_colsum) {
        const long long rem = a.B - tile * T;
        CS::template add_tile<S>(buf, rem < T ? (int)rem : T, s_col, s_part);
      }
      __syncthreads();
    } else if constexpr (P > 0) {
      __syncthreads();
    }
    slot = (slot + 1 == NB) ? 0 : slot + 1;
  }
  if constexpr (P > 0) cp_async_wait<0>();

  if (a.logp_sum) {
    const double s = block_sum<T>(lsum, red);
    if (threadIdx.x == 0) atomicAdd(a.logp_sum, s);
  }
  if constexpr (BWD && P > 0) {
    if (a.dt_colsum) {
      __syncthreads();
      for (int j = threadIdx.x; j < P; j += T) atomicAdd(a.dt_colsum + j, (double)s_col[j]);
    }
  }
  if (a.peer.world > 0) peer_allreduce<T>(a.peer);
}

template <class Spec, bool BWD, class M, int T, int NB, int MINB>
__global__ void __launch_bounds__(T, MINB) chain_kernel(const ChainArgs a) {
  chain_body<Spec, BWD, M, T, NB>(a);
}

}  // namespace nfn
