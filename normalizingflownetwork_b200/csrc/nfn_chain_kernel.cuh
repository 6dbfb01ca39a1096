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

// ---------------------------------------------------------------- event transform
// The estimators' y pipeline, folded into the heads as a prologue / epilogue so that it costs no extra pass over
// y or logp (reference estimators/BaseEstimator.py):
//   kXfNormalise  y' = (y - y_mean) / y_std            (:61-69; IEEE division, like the reference's op)
//   kXfNoise      y' += noise_std * N(0, 1)             (:66-68, training only) -- Philox4x32-10 keyed by `seed`,
//                 counter = (global row, `offset`): reproducible, independent of the launch geometry
//   logp_shift    added to every log-prob: -sum log y_std, the normalisation Jacobian (:55-59, :85-86)
//   kXfExp        the output is exp(logp + logp_shift) = prob / prod y_std  (pdf, :71-75)
// flags == 0 and logp_shift == 0 is the identity (what the plain entry points pass).  dy, where requested, is the
// gradient with respect to the TRANSFORMED event y'.
enum : int { kXfNormalise = 1, kXfNoise = 2, kXfExp = 4 };
struct EventXform {
  float mean[8];
  float std[8];
  float noise_std;
  float logp_shift;
  unsigned long long seed;
  unsigned long long offset;
  const unsigned long long* offset_dev;   // nullable: added to `offset` (a device-side step counter)
  int flags;
  int pad_;
};

NFN_DEVI void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1,
                            unsigned (&out)[4]) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// D standard normals for global row r (Box-Muller on Philox words; block b of 4 normals uses counter word 3 = b)
template <int D>
NFN_DEVI void row_normals(const EventXform& xf, long long r, float (&n)[D]) {
  const unsigned long long off = xf.offset + (xf.offset_dev ? __ldg(xf.offset_dev) : 0ull);
#pragma unroll
  for (int b = 0; b < (D + 3) / 4; ++b) {
    unsigned w[4];
    philox4x32_10((unsigned)r, (unsigned)((unsigned long long)r >> 32), (unsigned)off,
                  (unsigned)(off >> 32) ^ ((unsigned)b << 28), (unsigned)xf.seed, (unsigned)(xf.seed >> 32), w);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float u1 = fmaf((float)w[2 * h], 2.3283064365386963e-10f, 1.1641532182693481e-10f);  // (0, 1]
      const float u2 = (float)w[2 * h + 1] * 2.3283064365386963e-10f;                             // [0, 1]
      const float rad = sqrtf(-2.0f * logf(u1));
      float sn, cs;
      sincospif(2.0f * u2, &sn, &cs);
      if (4 * b + 2 * h < D) n[4 * b + 2 * h] = rad * cs;
      if (4 * b + 2 * h + 1 < D) n[4 * b + 2 * h + 1] = rad * sn;
    }
  }
}

// y -> y' in place; r is the GLOBAL row of the event (the noise counter), `noise` = false for grid events
template <int D>
NFN_DEVI void xform_event(const EventXform& xf, long long r, float (&z)[D], bool noise = true) {
  if (xf.flags & kXfNormalise) {
#pragma unroll
    for (int i = 0; i < D; ++i) z[i] = __fdiv_rn(z[i] - xf.mean[i], xf.std[i]);
  }
  if (noise && (xf.flags & kXfNoise)) {
    float n[D];
    row_normals<D>(xf, r, n);
#pragma unroll
    for (int i = 0; i < D; ++i) z[i] = fmaf(xf.noise_std, n[i], z[i]);
  }
}

// the value written to logp[]: log-prob (+ shift), or the density itself
template <class M>
NFN_DEVI float xform_out(const EventXform& xf, float lp) {
  const float v = lp + xf.logp_shift;
  if (xf.flags & kXfExp) {
    if constexpr (M::kFast) return M::exp(v); else return expf(v);
  }
  return v;
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
  double* out;                 // reduced result [n_values] of THIS launch's exchange, local
  unsigned* ticket;            // CTA arrival counter, self-resetting
  unsigned long long step;     // sequence number of this launch's exchange
  int world, rank, n_values;
  // Split-phase mode (deferred != 0): a launch leaves its totals in the local accumulators `acc` (two sets,
  // alternating by step parity) and does NOTHING else at its tail.  The NEXT launch on the communicator gives up
  // one CTA of its persistent grid to the exchange: that CTA pushes the previous totals to the peers, collects the
  // cross-rank sums (`pending_step`, into `pending_out`) and exits, all while the other CTAs stream tiles.  No
  // ticket, no fence, no round trip on any kernel's critical path; the price is 1 / grid of the tile throughput.
  int deferred;
  double* pending_out;         // nullptr: nothing pending
  double* acc_prev;            // the accumulators the previous launch filled
  unsigned long long pending_step;
  unsigned* status;            // sticky error word: != 0 once a peer failed to arrive within the time-out
  long long timeout_cycles;
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
  EventXform xf;
};

// ---------------------------------------------------------------- smem span load/store
// A parameter row is addressed either through a plain pointer (row-major tile with row stride S,
// the cp.async kernels and the fused Dense kernels) or through an SRow (the warp-tile kernels
// below, where the bulk-copy engine lays the tile down unpadded, 128-byte-swizzled when the row
// width needs it).  Span<OFF, N, V> moves N consecutive floats starting at column OFF with the
// widest accesses the alignment allows; all offsets are compile-time.
NFN_DEVI char* dyn_smem() {
  extern __shared__ __align__(16) char nfn_dyn_smem_[];
  return nfn_dyn_smem_;
}

// CW 16-byte chunks per box row (1 = linear layout), BOXB bytes between consecutive boxes of a tile:
// column col lives at x[(col/4) % CW] + ((col/4) / CW) * BOXB + (col % 4) * 4 bytes past the dynamic
// shared-memory base; x[c] = (this thread's row base) + ((c ^ swizzle key of the row) << 4).
template <int CW, int BOXB>
struct SRow {
  unsigned x[CW];
  template <int COL>
  NFN_DEVI char* at() const {
    constexpr int k = COL / 4;
    return dyn_smem() + (x[k % CW] + (unsigned)((k / CW) * BOXB + (COL % 4) * 4));
  }
};

template <int OFF, class T> NFN_DEVI const T* row_ptr(const float* row) { return reinterpret_cast<const T*>(row + OFF); }
template <int OFF, class T> NFN_DEVI T* row_ptr(float* row) { return reinterpret_cast<T*>(row + OFF); }
template <int OFF, class T, int CW, int BOXB>
NFN_DEVI T* row_ptr(SRow<CW, BOXB> row) { return reinterpret_cast<T*>(row.template at<OFF>()); }

template <int OFF, int N, int V>
struct Span {
  template <class R>
  NFN_DEVI static void load(R row, float* out) {
    if constexpr (N <= 0) {
      return;
    } else if constexpr (V >= 4 && OFF % 4 == 0 && N >= 4) {
      const float4 v = *row_ptr<OFF, const float4>(row);
      out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
      Span<OFF + 4, N - 4, V>::load(row, out + 4);
    } else if constexpr (V >= 2 && OFF % 2 == 0 && N >= 2) {
      const float2 v = *row_ptr<OFF, const float2>(row);
      out[0] = v.x; out[1] = v.y;
      Span<OFF + 2, N - 2, V>::load(row, out + 2);
    } else {
      out[0] = *row_ptr<OFF, const float>(row);
      Span<OFF + 1, N - 1, V>::load(row, out + 1);
    }
  }
  template <class R>
  NFN_DEVI static void store(R row, const float* in) {
    if constexpr (N <= 0) {
      return;
    } else if constexpr (V >= 4 && OFF % 4 == 0 && N >= 4) {
      *row_ptr<OFF, float4>(row) = make_float4(in[0], in[1], in[2], in[3]);
      Span<OFF + 4, N - 4, V>::store(row, in + 4);
    } else if constexpr (V >= 2 && OFF % 2 == 0 && N >= 2) {
      *row_ptr<OFF, float2>(row) = make_float2(in[0], in[1]);
      Span<OFF + 2, N - 2, V>::store(row, in + 2);
    } else {
      *row_ptr<OFF, float>(row) = in[0];
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
  template <class R>
  NFN_DEVI static void run(R row, float (&z)[Spec::D], float (&zs)[Spec::KA][Spec::D],
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
  template <class R>
  NFN_DEVI static void run(R row, const float (&zs)[Spec::KA][Spec::D], float (&G)[Spec::D],
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

// Collect: sum, in rank order (deterministic), the words every peer pushed into THIS rank's region for exchange
// `step` and write the totals to `out`.  Must be called by ALL threads of one CTA (barriers inside).  Work item
// (value j, peer q) -> one thread: every word pair is requested at once (one L2 latency in the steady state, where
// everything has long arrived) instead of peer after peer; the per-value sums then run over shared memory.
// A peer that does not arrive within timeout_cycles is abandoned: NaN in `out` AND the sticky status word is
// set, which the host reads back (nfn_peer_status) -- a time-out is an error, never a silent NaN.
constexpr int kPeerChunk = 64;   // values per pass (shared scratch: kPeerChunk x kMaxPeers doubles)
NFN_DEVI void peer_collect(const PeerArgs& p, unsigned long long step, double* out, int tid, int nthreads) {
  __shared__ double s_val[kPeerChunk * kMaxPeers];
  __shared__ int s_timeout;
  const int W = p.world, NV = p.n_values, par = (int)(step & 1ull);
  const unsigned long long tag = ((step + 1ull) & 0xffffffffull) << 32;  // never 0 in the first 2^32 steps
  const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(p.base[p.rank]) + (size_t)par * W * NV * 2;
  const long long t0 = clock64();
  if (tid == 0) s_timeout = 0;
  for (int j0 = 0; j0 < NV; j0 += kPeerChunk) {
    const int nj = NV - j0 < kPeerChunk ? NV - j0 : kPeerChunk;
    __syncthreads();
    for (int i = tid; i < nj * W; i += nthreads) {
      const int j = j0 + i / W, q = i % W;
      const unsigned long long* src = mine + ((size_t)q * NV + j) * 2;
      unsigned long long a0 = ld_volatile_u64(src), a1 = ld_volatile_u64(src + 1);
      while ((a0 & 0xffffffff00000000ull) != tag || (a1 & 0xffffffff00000000ull) != tag) {
        if (clock64() - t0 > p.timeout_cycles) {
          s_timeout = 1;
          break;
        }
        a0 = ld_volatile_u64(src);
        a1 = ld_volatile_u64(src + 1);
      }
      s_val[i] = __longlong_as_double((long long)((a0 & 0xffffffffull) | (a1 << 32)));
    }
    __syncthreads();
    const bool timeout = s_timeout != 0;
    for (int jj = tid; jj < nj; jj += nthreads) {
      double sum = 0.0;
      for (int q = 0; q < W; ++q) sum += s_val[jj * W + q];
      out[j0 + jj] = timeout ? __longlong_as_double(0x7ff8000000000000ll) : sum;
    }
    if (timeout && tid == 0) atomicOr(p.status, 1u);
  }
}

// Push: take (and reset) the local totals `acc` of exchange `step` and write (value half | step tag) words into
// slot [parity][rank] of EVERY peer's region (own region included).
NFN_DEVI void peer_push(const PeerArgs& p, double* acc, unsigned long long step, int tid, int nthreads) {
  const int W = p.world, NV = p.n_values, par = (int)(step & 1ull);
  const unsigned long long tag = ((step + 1ull) & 0xffffffffull) << 32;
  for (int j = tid; j < NV; j += nthreads) {
    const unsigned long long bits = atomicExch(reinterpret_cast<unsigned long long*>(acc + j), 0ull);
    const unsigned long long w0 = (bits & 0xffffffffull) | tag, w1 = (bits >> 32) | tag;
    const size_t off = ((size_t)(par * W + p.rank) * NV + j) * 2;
    for (int q = 0; q < W; ++q) {
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(p.base[q]) + off;
      st_volatile_u64(dst, w0);
      st_volatile_u64(dst + 1, w1);
    }
  }
}

// Split-phase exchange of the PREVIOUS launch's totals: push them to every peer, collect every peer's, write the
// sums.  Whole CTA (barriers inside); runs after griddepcontrol.wait, i.e. the previous launch has completed and
// its accumulators are final.  Slot reuse is safe: exchange s+2 is pushed (by launch s+3) only after the same
// rank's launch s+2 has collected s+1 from every peer, and a peer pushes s+1 only after its launch s+1 -- its
// collect of s included -- has completed.
NFN_DEVI void peer_exchange_prev(const PeerArgs& p, int tid, int nthreads) {
  if (p.pending_out) {
    peer_push(p, p.acc_prev, p.pending_step, tid, nthreads);
    peer_collect(p, p.pending_step, p.pending_out, tid, nthreads);
  }
}

// Blocking exchange at the tail of a launch: called by every thread of every CTA after the CTA's accumulator
// atomics.  The last CTA to arrive pushes this step's totals, waits for every peer's and sums them: one NVLink
// round trip plus the rank skew on the kernel's tail, `out` complete when the kernel completes.
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
  peer_push(p, p.acc, p.step, (int)threadIdx.x, T);
  peer_collect(p, p.step, p.out, (int)threadIdx.x, T);
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
    if (a.xf.flags) xform_event<D>(a.xf, tile * T + threadIdx.x, z);
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
            if (a.xf.flags) xform_event<D>(a.xf, j, zg, false);
            float zs[Spec::KA][D];
            LogDetAcc<M> ld;
            FwdSweep<Spec, M, V, false, 0>::run(row, zg, zs, ld);
            a.logp[(long long)j * a.B + r] = xform_out<M>(a.xf, Base::log_prob(bth, zg) + ld.nat());
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
      const float lpo = xform_out<M>(a.xf, lp);   // + the normalisation Jacobian (or the density itself)
      a.logp[r] = lpo;
      lsum += (double)lpo;
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
  // (split-phase mode is served by a separate one-CTA exchange launch for this kernel generation, see chain_dispatch)
  if (a.peer.world > 0 && !a.peer.deferred) peer_allreduce<T>(a.peer);
}

template <class Spec, bool BWD, class M, int T, int NB, int MINB>
__global__ void __launch_bounds__(T, MINB) chain_kernel(const ChainArgs a) {
  chain_body<Spec, BWD, M, T, NB>(a);
}


// ================================================================ warp-tile kernels (bulk-copy engine)
// Second generation of the chain kernel: every WARP owns its own tile pipeline, there is no CTA
// barrier in the tile loop, and the tile moves through the bulk-copy engine instead of through
// per-thread cp.async / LDS / STG instructions:
//
//   global t  --one elected lane: cp.async.bulk (1-D) or cp.async.bulk.tensor.2d (TMA), mbarrier
//               complete_tx-->  smem warp tile [32 rows x P], UNPADDED
//   lane r reads row r at compile-time offsets, runs the chain, writes d logp / d theta in place
//   smem warp tile  --fence.proxy.async + one elected lane: bulk store (bulk_group)-->  global dt
//
// A warp tile is 32 consecutive rows = one contiguous 128*P-byte span.  Unpadded rows are
// bank-conflict free for the per-thread V-wide reads exactly when P/V is odd (see row_stride);
// otherwise (P % 8 == 0) the tile is loaded as boxes of [32 rows x W columns], W*4 = 32/64/128
// bytes, through a 2-D tensor map with the matching 32B/64B/128B swizzle: the hardware XORs
// the 16-byte chunk index with the row's address bits, thread r undoes it with a per-thread XOR key
// that is folded into CW = W/4 base registers (SRow), so every parameter access still is one
// LDS/STS at [register + immediate] and 8 consecutive lanes hit 8 distinct bank groups.
// Out-of-range rows of the last tile are zero-filled / clipped by the tensor unit itself.
struct alignas(64) TensorMap {   // CUtensorMap (128 opaque bytes), declared here so that NVRTC needs no cuda.h
  unsigned long long opaque[16];
};

// box width (floats) of the swizzled layout for row width P; 0: linear layout (1-D bulk copies)
__host__ __device__ constexpr int warp_tile_box(int P) {
  return (P % 8 != 0) ? 0 : (P % 32 == 0 ? 32 : (P % 16 == 0 ? 16 : 8));
}

template <int P>
struct WarpTile {
  static constexpr int V = row_vec(P);
  static constexpr bool kSwz = warp_tile_box(P) > 0;
  static constexpr int W = kSwz ? warp_tile_box(P) : 4;
  static constexpr int CW = kSwz ? W / 4 : 1;
  static constexpr int NBX = kSwz ? P / W : 1;                 // boxes per tile
  static constexpr int kBoxBytes = kSwz ? 32 * W * 4 : 16;     // linear layout: "boxes" of one chunk
  static constexpr int kRowBytes = kSwz ? W * 4 : P * 4;       // smem pitch of this thread's row
  static constexpr unsigned kTileBytes = 128u * P;
  using Row = SRow<CW, kBoxBytes>;

  NFN_DEVI static int key(int r) {
    if constexpr (!kSwz) return 0;
    else if constexpr (W == 32) return r & 7;
    else if constexpr (W == 16) return (r >> 1) & 3;
    else return (r >> 2) & 1;
  }
  // accessor of row r (0..31) of the tile at byte offset `tile` of the dynamic shared memory
  NFN_DEVI static Row row(unsigned tile, int r) {
    Row x;
    const unsigned base = tile + (unsigned)(r * kRowBytes);
#pragma unroll
    for (int c = 0; c < CW; ++c) x.x[c] = base + (unsigned)((c ^ key(r)) << 4);
    return x;
  }
  // byte offset (inside the tile) of V-wide column group g of row r
  NFN_DEVI static unsigned group_off(int r, int g) {
    if constexpr (!kSwz) {
      return (unsigned)((r * P + g * V) * 4);
    } else {
      const int j = g / CW, c = g % CW;
      return (unsigned)(j * kBoxBytes + r * kRowBytes + ((c ^ key(r)) << 4));
    }
  }
};

// ---- mbarrier / bulk-copy PTX
NFN_DEVI void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
NFN_DEVI void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
NFN_DEVI void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "NFN_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra NFN_MBAR_DONE;\n"
      "bra NFN_MBAR_WAIT;\n"
      "NFN_MBAR_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
NFN_DEVI void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
NFN_DEVI bool elect_one() {
  unsigned pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
NFN_DEVI void bulk_load_1d(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
NFN_DEVI void bulk_store_1d(void* dst, unsigned src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
NFN_DEVI void tma_load_2d(unsigned dst, const TensorMap* tm, int c0, int c1, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(tm), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
NFN_DEVI void tma_store_2d(const TensorMap* tm, int c0, int c1, unsigned src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(c0),
               "r"(c1), "r"(src)
               : "memory");
}
NFN_DEVI void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
NFN_DEVI void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// ---- column sums of dt (bias gradient of the emitting layer), warp-private partial sums kept in REGISTERS
// across all tiles of the warp and combined once, at the very end of the kernel.  Three strategies:
//   kOwnRow  (linear layout, P <= 32): every lane adds its own gradient row (P accumulators per lane; the row
//            was written a moment ago, so the values usually never leave the registers);
//   kSwizzled (TMA layout): CW lanes share one box row (its CW 16-byte chunks), 32/CW rows per step -- a step
//            reads whole 128-byte lines, conflict-free, at [lane register + immediate] addresses; a lane owns the
//            same logical chunk of every box, i.e. 4 * NBX accumulators;
//   kColumn  (linear layout, wide rows): lane l owns columns l, l + 32, ... and walks the 32 rows.
template <int P>
struct WarpColSum {
  using L = WarpTile<P>;
  static constexpr int V = L::V;
  enum { kOwnRow = 0, kSwizzled = 1, kColumn = 2 };
  static constexpr int kMode = L::kSwz ? kSwizzled : (P <= 32 ? kOwnRow : kColumn);
  static constexpr int CW = L::CW;                       // lanes per box row (swizzled)
  static constexpr int RPS = 32 / CW;                    // rows per step (swizzled)
  static constexpr int NCOL = (P + 31) / 32;             // columns per lane (kColumn)
  static constexpr int NACC = kMode == kOwnRow ? P : (kMode == kSwizzled ? 4 * L::NBX : NCOL);
  float acc[NACC];

  NFN_DEVI void clear() {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.0f;
  }

  // kOwnRow: called by the lanes that own a valid row, right after their reverse sweep
  NFN_DEVI void add_own_row(typename L::Row row) {
    if constexpr (kMode == kOwnRow) {
      float g[P];
      Span<0, P, V>::load(row, g);
#pragma unroll
      for (int i = 0; i < P; ++i) acc[i] += g[i];
    }
  }

  // kSwizzled / kColumn: called by the whole warp once the tile's gradients are in shared memory; rows >= nvalid
  // of a swizzled tile hold the zero fill of the tensor unit (their lanes wrote nothing)
  NFN_DEVI void add_tile(unsigned tile, int lane, int nvalid) {
    const char* base = dyn_smem() + tile;
    if constexpr (kMode == kSwizzled) {
      const int c = lane % CW, q = lane / CW;
#pragma unroll
      for (int j = 0; j < L::NBX; ++j) {
#pragma unroll
        for (int s = 0; s < CW; ++s) {                   // 32 rows = CW steps of RPS rows
          const int r = s * RPS + q;
          const float4 x = *reinterpret_cast<const float4*>(base + j * L::kBoxBytes + r * L::kRowBytes +
                                                            ((c ^ L::key(r)) << 4));
          acc[4 * j + 0] += x.x; acc[4 * j + 1] += x.y; acc[4 * j + 2] += x.z; acc[4 * j + 3] += x.w;
        }
      }
    } else if constexpr (kMode == kColumn) {
#pragma unroll
      for (int m = 0; m < NCOL; ++m) {
        const int col = lane + 32 * m;
        if (col < P) {
#pragma unroll 8
          for (int r = 0; r < 32; ++r)
            if (r < nvalid) acc[m] += *reinterpret_cast<const float*>(base + (r * P + col) * 4);
        }
      }
    }
  }

  // warp-level totals -> s_part[P] (one row per warp); every lane of the warp must call
  NFN_DEVI void deposit(float* s_part, int lane) {
    if constexpr (kMode == kOwnRow) {
#pragma unroll
      for (int i = 0; i < P; ++i) {
        float v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_part[i] = v;
      }
    } else if constexpr (kMode == kSwizzled) {
      const int c = lane % CW;
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        float v = acc[i];
#pragma unroll
        for (int o = 16; o >= CW; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);   // over the row groups q
        if (lane < CW) s_part[(i / 4) * L::W + 4 * c + (i % 4)] = v;                  // box i/4, chunk c
      }
    } else {
#pragma unroll
      for (int m = 0; m < NCOL; ++m) {
        const int col = lane + 32 * m;
        if (col < P) s_part[col] = acc[m];
      }
    }
  }
};

// Launch geometry of the warp-tile kernels: NW warps per CTA (T = 32 NW threads), NB tile buffers per
// warp, MINB resident CTAs per SM (the grid is SMs x MINB, never more).  `hist` = K * D, the floats of z history
// a thread keeps for the reverse sweep (a proxy for its register footprint).  nb / warps > 0 override the
// defaults (A/B sweeps through the runtime specialiser: options "tune_wnb", "tune_wwarps").
//
// Defaults from the sweeps on B200 (profiles/tuning_r02.md).  What the kernels need is BYTES IN FLIGHT: the
// loaded DRAM latency is ~2 us, so ~90 KB of loads per SM must be outstanding to stream at the HBM rate.  The
// fused kernel refills a buffer only after its store has drained (at the END of an iteration), so of NB buffers
// one is in the registers' hands, one is draining and NB - 2 are loading: small tiles take 16 warps x 3 buffers;
// when that does not fit the SM's 227 KB, depth beats warps (4 buffers, as many warps as fit, whole CTAs of 4).
// The forward kernel (nothing to store) keeps NB - 1 loads in flight per warp.
__host__ __device__ constexpr ChainGeometry warp_tile_geometry(int P, bool bwd, int hist, int nb = 0, int warps = 0) {
  const unsigned tile = 128u * (unsigned)(P > 0 ? P : 1);
  const unsigned budget = 227u * 1024u - 4u * 2560u;   // per SM, minus per-CTA reservations / alignment slack
  (void)hist;
  int want = warps;
  if (bwd) {
    if (nb <= 0) nb = (16u * 3u * tile <= budget) ? 3 : 4;
    if (want <= 0) want = 16;
  } else {
    // ~64 KB of loads in flight per SM: 4 warps x 16 KB tiles (cfg3), 8 x 6 KB (cfg2), 16 x 4 KB; tiles under 3 KB
    // (1-D chains, a handful of flows) need every warp slot and a third buffer: 32 warps x 3
    if (nb <= 0) nb = tile < 3072u ? 3 : 2;
    if (want <= 0) {
      want = (int)(65536u / tile) / 4 * 4;
      want = want < 4 ? 4 : (want > 32 ? 32 : want);
    }
  }
  int w = (int)(budget / ((unsigned)nb * tile));
  while (w < 4 && nb > 2) {   // very wide rows: give depth back until at least one CTA of 4 warps fits
    --nb;
    w = (int)(budget / ((unsigned)nb * tile));
  }
  if (w > want) w = want;
  if (warps <= 0 && w >= 8) w = w / 4 * 4;   // defaults use whole CTAs of 4 warps
  if (w < 1) w = 1;
  const int ctas = (w + 3) / 4;
  const int nw = w / ctas;
  return ChainGeometry{nw * 32, nb, ctas, (unsigned)(nw * nb) * tile + 1024u, 0, ctas, 0};
}

template <class Spec, bool BWD, class M, int NW, int NB>
NFN_DEVI void chain_body_w(const ChainArgs& a, const TensorMap* tm_t, const TensorMap* tm_dt) {
  constexpr int D = Spec::D;
  constexpr int P = Spec::P();
  static_assert(P > 0, "the warp-tile kernel needs a parameter row");
  using L = WarpTile<P>;
  using CS = WarpColSum<P>;
  constexpr int V = L::V;
  constexpr int T = NW * 32;
  static_assert(NB >= 2 && NB <= 4, "2..4 tile buffers per warp");

  __shared__ __align__(8) unsigned long long full[NW * NB];
  __shared__ double red[NW];

  // warp-uniform by construction (the compiler keeps the tile bookkeeping in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = (int)(threadIdx.x & 31);
  const unsigned dyn_u32 = smem_u32(dyn_smem());
  const unsigned pad = (1024u - (dyn_u32 & 1023u)) & 1023u;   // swizzle patterns are functions of the address
  const unsigned tile0 = pad + (unsigned)(warp * NB) * L::kTileBytes;  // byte offset of this warp's buffer 0
  const unsigned bar0 = smem_u32(full) + (unsigned)(warp * NB) * 8u;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NW * NB; ++i) mbar_init(smem_u32(full) + 8u * i, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // programmatic dependent launch, see chain_body
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // split-phase exchange: the LAST CTA of the grid carries no tiles; it exchanges the previous launch's totals with
  // the peers while the other CTAs stream, and exits.  (A grid of one CTA does both, one after the other.)
  const bool xch = a.peer.world > 0 && a.peer.deferred != 0;
  const int tile_ctas = (xch && gridDim.x > 1) ? (int)gridDim.x - 1 : (int)gridDim.x;
  if (xch) {
    if ((int)blockIdx.x == tile_ctas || gridDim.x == 1) peer_exchange_prev(a.peer, (int)threadIdx.x, T);
    if ((int)blockIdx.x == tile_ctas) return;
  }

  const long long nwt = (a.B + 31) / 32;                        // warp tiles
  const long long GW = (long long)tile_ctas * NW;               // tile-carrying warps in the grid
  long long wt = (long long)blockIdx.x * NW + warp;
  // the last tile of a linear-layout kernel may be ragged: its bytes are not a whole number of 16-byte
  // chunks in general, so it is moved with plain loads / stores by the lanes (one warp, once per launch)
  const long long ragged = (!L::kSwz && (a.B & 31)) ? nwt - 1 : -1;

  auto issue_load = [&](int slot, long long tl) {
    if (tl == ragged) return;
    if (elect_one()) {
      const unsigned bar = bar0 + 8u * slot, dst = dyn_u32 + tile0 + (unsigned)slot * L::kTileBytes;
      mbar_expect_tx(bar, L::kTileBytes);
      if constexpr (L::kSwz) {
#pragma unroll
        for (int j = 0; j < L::NBX; ++j) tma_load_2d(dst + (unsigned)(j * L::kBoxBytes), tm_t, j * L::W, (int)(tl * 32), bar);
      } else {
        bulk_load_1d(dst, a.t + tl * (32 * P), L::kTileBytes, bar);
      }
    }
  };

  // prologue: NB-1 tiles in flight (forward) / NB-2 + the refill at the end of the first iteration (fused)
  constexpr int kAhead = BWD ? NB - 2 : NB - 1;   // loads issued ahead of the tile being consumed, at its start
#pragma unroll
  for (int s = 0; s < NB - 1; ++s) {
    const long long tl = wt + (long long)s * GW;
    if (s <= kAhead && tl < nwt) issue_load(s, tl);
  }

  float y_nxt[D];
  float g_nxt = 1.0f;
  {
    const long long r0 = wt * 32 + lane;
#pragma unroll
    for (int i = 0; i < D; ++i) y_nxt[i] = 0.0f;
    if (wt < nwt && r0 < a.B) {
      load_event<D>(a.y, a.y_broadcast ? 0 : r0, y_nxt);
      if constexpr (BWD) { if (a.g_logp) g_nxt = __ldg(a.g_logp + r0); }
    }
  }

  CS cs;
  if constexpr (BWD) cs.clear();
  double lsum = 0.0;
  int slot = 0;
  unsigned parity = 0;
  for (; wt < nwt; wt += GW) {
    const unsigned tile = tile0 + (unsigned)slot * L::kTileBytes;
    float z[D];
#pragma unroll
    for (int i = 0; i < D; ++i) z[i] = y_nxt[i];
    if (a.xf.flags) xform_event<D>(a.xf, wt * 32 + lane, z);
    const float g_cur = g_nxt;
    {
      const long long rn = (wt + GW) * 32 + lane;
      if (rn < a.B) {
        load_event<D>(a.y, a.y_broadcast ? 0 : rn, y_nxt);
        if constexpr (BWD) { if (a.g_logp) g_nxt = __ldg(a.g_logp + rn); }
      }
    }
    if constexpr (!BWD) {
      // forward: nothing is written back, the buffer of the previous tile is free as soon as every lane
      // has left it (the __syncwarp that ended the previous iteration): refill it right away
      const long long nxt = wt + (long long)(NB - 1) * GW;
      if (nxt < nwt) issue_load(slot == 0 ? NB - 1 : slot - 1, nxt);
    }
    const int nvalid = (a.B - wt * 32 >= 32) ? 32 : (int)(a.B - wt * 32);
    if (wt == ragged) {
      float* dstf = reinterpret_cast<float*>(dyn_smem() + tile);
      const float* src = a.t + wt * (32 * P);
      for (int e = lane; e < nvalid * P; e += 32) dstf[e] = __ldg(src + e);
      __syncwarp();
    } else {
      mbar_wait(bar0 + 8u * slot, parity);
    }

    const long long r = wt * 32 + lane;
    const typename L::Row row = L::row(tile, lane);
    bool done = false;
    if constexpr (!BWD) {
      if (a.grid_ny > 0) {
        if (r < a.B) {
          using Base = BaseDist<D, Spec::BASE, M>;
          float bth[Base::NA];
          if constexpr (Spec::BASE) Span<0, 2 * D, V>::load(row, bth);
          for (int j = 0; j < a.grid_ny; ++j) {
            float zg[D];
            load_event<D>(a.y, j, zg);  // same address for the whole warp: one broadcast load
            if (a.xf.flags) xform_event<D>(a.xf, j, zg, false);
            float zs[Spec::KA][D];
            LogDetAcc<M> ld;
            FwdSweep<Spec, M, V, false, 0>::run(row, zg, zs, ld);
            a.logp[(long long)j * a.B + r] = xform_out<M>(a.xf, Base::log_prob(bth, zg) + ld.nat());
          }
        }
        done = true;
      }
    }
    if (!done && r < a.B) {
      float zs[Spec::KA][D];
      LogDetAcc<M> ld;
      FwdSweep<Spec, M, V, BWD, 0>::run(row, z, zs, ld);
      using Base = BaseDist<D, Spec::BASE, M>;
      float bth[Base::NA];
      if constexpr (Spec::BASE) Span<0, 2 * D, V>::load(row, bth);
      const float lp = (BWD ? Base::log_prob_save(bth, z) : Base::log_prob(bth, z)) + ld.nat();
      const float lpo = xform_out<M>(a.xf, lp);   // + the normalisation Jacobian (or the density itself)
      a.logp[r] = lpo;
      lsum += (double)lpo;
      if constexpr (BWD) {
        const float cot = a.g_scale * g_cur;
        float G[D];
        float gb[Base::NA];
        Base::bwd_saved(bth, z, cot, G, gb);
        if constexpr (Spec::BASE) Span<0, 2 * D, V>::store(row, gb);
        BwdSweep<Spec, M, V, Spec::K - 1>::run(row, zs, G, cot);
        if (a.dy) store_event<D>(a.dy, r, G);
        if constexpr (CS::kMode == CS::kOwnRow) { if (a.dt_colsum) cs.add_own_row(row); }
      }
    }

    if constexpr (BWD) {
      fence_async_smem();   // this lane's gradient writes -> visible to the bulk-copy engine
      __syncwarp();
      if (wt == ragged) {
        const float* srcf = reinterpret_cast<const float*>(dyn_smem() + tile);
        float* dst = a.dt + wt * (32 * P);
        for (int e = lane; e < nvalid * P; e += 32) dst[e] = srcf[e];
      } else if (elect_one()) {
        const unsigned src = dyn_u32 + tile;
        if constexpr (L::kSwz) {
#pragma unroll
          for (int j = 0; j < L::NBX; ++j) tma_store_2d(tm_dt, j * L::W, (int)(wt * 32), src + (unsigned)(j * L::kBoxBytes));
        } else {
          bulk_store_1d(a.dt + wt * (32 * P), src, L::kTileBytes);
        }
        bulk_commit();
      }
      if constexpr (CS::kMode != CS::kOwnRow) { if (a.dt_colsum) cs.add_tile(tile, lane, nvalid); }
      // refill the buffer of the PREVIOUS tile: its store was committed one whole iteration ago, so waiting
      // for "all but the newest group have been read" costs nothing
      const long long nxt = wt + (long long)(NB - 1) * GW;
      if (nxt < nwt) {
        if (elect_one()) bulk_wait_read<1>();
        __syncwarp();
        issue_load(slot == 0 ? NB - 1 : slot - 1, nxt);
      }
    } else {
      __syncwarp();
    }
    if (++slot == NB) { slot = 0; parity ^= 1u; }
  }

  if constexpr (BWD) {
    // the tile buffers must outlive the bulk stores that read them
    bulk_wait_read<0>();
  }
  if (a.logp_sum) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if (lane == 0) red[warp] = lsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < NW; ++i) s += red[i];
      atomicAdd(a.logp_sum, s);
    }
  }
  if constexpr (BWD) {
    if (a.dt_colsum) {
      __syncthreads();      // every warp is done with its tile buffers: reuse them as scratch [NW][P]
      float* scratch = reinterpret_cast<float*>(dyn_smem() + pad);
      cs.deposit(scratch + warp * P, lane);
      __syncthreads();
      for (int j = threadIdx.x; j < P; j += T) {
        float s = 0.0f;
#pragma unroll
        for (int q = 0; q < NW; ++q) s += scratch[q * P + j];
        atomicAdd(a.dt_colsum + j, (double)s);
      }
    }
  }
  if (a.peer.world > 0 && !a.peer.deferred) peer_allreduce<T>(a.peer);   // blocking exchange at the tail
}

template <class Spec, bool BWD, class M, int NW, int NB, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
chain_kernel_w(const ChainArgs a, const __grid_constant__ TensorMap tm_t, const __grid_constant__ TensorMap tm_dt) {
  chain_body_w<Spec, BWD, M, NW, NB>(a, &tm_t, &tm_dt);
}

}  // namespace nfn
