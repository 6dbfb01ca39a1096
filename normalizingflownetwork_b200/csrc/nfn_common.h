// nfn_common.h -- host-side plumbing shared by the translation units of libnfn_b200.so:
// thread-local error string, launch counter, device properties, and the registry that maps
// a chain descriptor to its compile-time specialised launchers.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../include/nfn_b200.h"
#include "nfn_chain_kernel.cuh"
#include "nfn_dense_chain.cuh"
#include "nfn_dense_tc5.cuh"

struct nfn_peer_comm;  // opaque handle of the C ABI (nfn_peer.cu)

namespace nfn {

int set_error(int code, const char* fmt, ...);
int cuda_error(cudaError_t e, const char* what);
void count_launch();

struct DeviceInfo {
  int device = -1;
  int sm_count = 0;
  int smem_optin = 0;
};
const DeviceInfo& device_info();

// Run-time switches: environment defaults read once, nfn_set_option() afterwards (no getenv on a launch path).
enum Opt {
  kOptMath,          // 0 fast (MUFU-based, default) / 1 accurate (CUDA libm + IEEE division)   NFN_B200_MATH
  kOptForceGeneric,  // every chain through the runtime-chain kernel                          NFN_B200_FORCE_GENERIC
  kOptForceJit,      // skip the ahead-of-time instances                                      NFN_B200_FORCE_JIT
  kOptJit,           // runtime specialiser enabled (default 1)                               NFN_B200_JIT=0
  kOptChainIo,       // -1 measured default / 0 cp.async CTA tiles / 1 bulk copy, TMA warp tiles  NFN_B200_CHAIN_IO
  kOptDenseMma,      // 0 auto / 1 tcgen05 / 2 mma.sync                                       NFN_B200_DENSE_MMA
  kOptPdl,           // programmatic dependent launch (default 1)                             NFN_B200_PDL=0
  kOptDebug,         // NFN_B200_DEBUG
  kOptHostChunkMb,   // NFN_B200_HOST_CHUNK_MB
  kOptTuneWnb,       // warp-tile kernels through the runtime specialiser: tile buffers per warp (0: default)
  kOptTuneWwarps,    // ... and warps per SM (0: default)                                     A/B sweeps only
  kOptMlpMma,        // hidden layers' backward on the tensor cores (default 1)               NFN_B200_MLP_MMA=0
  kOptCount
};
int option(Opt o);
int math_mode();
// Programmatic dependent launch for the specialised chain kernels (NFN_B200_PDL=0 turns it off).
bool pdl_enabled();

typedef cudaError_t (*ChainLaunchFn)(const ChainArgs&, cudaStream_t);

struct ChainKernels {
  // [math mode][bwd]
  ChainLaunchFn fn[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // cp.async CTA-tile kernels
  ChainLaunchFn fnw[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // bulk-copy / TMA warp-tile kernels
  int P = 0;
};

// Which of the two kernel generations serves a launch: NFN_B200_CHAIN_IO=tma|cpasync forces one (A/B runs,
// tests); otherwise the measured default for the row width / direction (chain_prefers_warp_tile).
int chain_io_override();   // -1: none, 0: cp.async, 1: bulk copy / TMA
// Measured on B200 (profiles/tuning_r02.md): the warp-tile kernels win on every BASELINE chain, both directions
// (cfg2 fwd+bwd with column sums 69 vs 83 us, cfg3 fwd 0.63 vs 0.72 ms, cfg4 27 vs 31 us).
__host__ __device__ constexpr bool chain_prefers_warp_tile(int P, bool bwd) {
  (void)bwd;
  return P > 0;
}

// 2-D tensor map over a row-major fp32 matrix [rows, P] with boxes of [32 rows x W columns] and the
// 32B/64B/128B swizzle that matches W (WarpTile<P>); cuTensorMapEncodeTiled through the runtime's
// driver entry point (the library does not link libcuda).
int encode_row_tensor_map(TensorMap* out, const float* base, long long rows, int P, int W);

std::string chain_key(int d, bool base, int k, const uint8_t* types);
void register_chain(const std::string& key, const ChainKernels& k);
const ChainKernels* find_chain(const std::string& key);

// ------------------------------------------------------------------ specialised launcher
// Geometry of an AOT-specialised kernel; NFN_TUNE_* macros override it for A/B variant builds.
template <class Spec, bool BWD>
struct ChainTune {
  static constexpr ChainGeometry kGeo = chain_geometry(Spec::P(), BWD);
#ifdef NFN_TUNE_T
  static constexpr int T = NFN_TUNE_T;
#else
  static constexpr int T = kGeo.T;
#endif
#ifdef NFN_TUNE_NB
  static constexpr int NB = NFN_TUNE_NB;
#else
  static constexpr int NB = kGeo.NB;
#endif
#ifdef NFN_TUNE_MINB
  static constexpr int MINB = NFN_TUNE_MINB;
#else
  static constexpr int MINB = kGeo.MINB;
#endif
  static constexpr size_t kTile =
      Spec::P() > 0 ? (size_t)T * row_stride(Spec::P()) * sizeof(float) : 0;
};

template <class Spec, bool BWD, class M>
cudaError_t launch_chain(const ChainArgs& a, cudaStream_t st) {
  using Tune = ChainTune<Spec, BWD>;
  constexpr int T = Tune::T;
  constexpr int NB = Tune::NB;
  constexpr int MINB = Tune::MINB;
  constexpr size_t kSmem = Tune::kTile * NB;
  auto kern = chain_kernel<Spec, BWD, M, T, NB, MINB>;

  struct Cfg {
    int device = -1;
    int ctas_per_sm = 0;
  };
  static thread_local Cfg cfg;
  const DeviceInfo& di = device_info();
  if (cfg.device != di.device) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, kSmem);
    if (e != cudaSuccess) return e;
    cfg.ctas_per_sm = occ > 0 ? occ : 1;
    cfg.device = di.device;
  }
  const long long ntiles = (a.B + T - 1) / T;
  long long grid = (long long)di.sm_count * cfg.ctas_per_sm;
  if (grid > ntiles) grid = ntiles;
  if (pdl_enabled()) {
    // programmatic dependent launch: this grid's CTAs may be scheduled while the previous launch in the
    // stream is still retiring; the kernel waits (griddepcontrol.wait) before its first global access
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)grid);
    lc.blockDim = dim3((unsigned)T);
    lc.dynamicSmemBytes = kSmem;
    lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&lc, kern, a);
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
  }
  kern<<<(unsigned)grid, T, kSmem, st>>>(a);
  count_launch();
  return cudaGetLastError();
}

template <class Spec, bool BWD, class M>
cudaError_t launch_chain_w(const ChainArgs& a, cudaStream_t st) {
  constexpr int P = Spec::P();
  if constexpr (P <= 0) {
    return launch_chain<Spec, BWD, M>(a, st);
  } else {
    constexpr ChainGeometry kGeo = warp_tile_geometry(P, BWD, Spec::K * Spec::D);
    constexpr int NW = kGeo.T / 32;
    constexpr int NB = kGeo.NB;
    constexpr int MINB = kGeo.MINB;
    constexpr size_t kSmem = kGeo.smem_bytes;
    using L = WarpTile<P>;
    auto kern = chain_kernel_w<Spec, BWD, M, NW, NB, MINB>;
    struct Cfg {
      int device = -1;
      int ctas_per_sm = 0;
    };
    static thread_local Cfg cfg;
    const DeviceInfo& di = device_info();
    if (cfg.device != di.device) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
      if (e != cudaSuccess) return e;
      int occ = 0;
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NW * 32, kSmem);
      if (e != cudaSuccess) return e;
      cfg.ctas_per_sm = occ > 0 ? (occ < MINB ? occ : MINB) : 1;   // the geometry's CTA count, never more
      cfg.device = di.device;
    }
    TensorMap tm_t{}, tm_dt{};
    if constexpr (L::kSwz) {
      if (encode_row_tensor_map(&tm_t, a.t, a.B, P, L::W) != NFN_OK) return cudaErrorInvalidValue;
      if (BWD && encode_row_tensor_map(&tm_dt, a.dt, a.B, P, L::W) != NFN_OK) return cudaErrorInvalidValue;
    }
    // split-phase peer exchange: the last CTA of the grid carries the exchange instead of tiles
    const long long ntiles = (a.B + NW * 32 - 1) / (NW * 32) + ((a.peer.world > 0 && a.peer.deferred) ? 1 : 0);
    long long grid = (long long)di.sm_count * cfg.ctas_per_sm;
    if (grid > ntiles) grid = ntiles;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)grid);
    lc.blockDim = dim3((unsigned)(NW * 32));
    lc.dynamicSmemBytes = kSmem;
    lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = pdl_enabled() ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&lc, kern, a, tm_t, tm_dt);
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
  }
}

template <class Spec>
struct ChainRegistrar {
  explicit ChainRegistrar() {
    ChainKernels k;
    k.fn[0][0] = &launch_chain<Spec, false, MathFast>;
    k.fn[0][1] = &launch_chain<Spec, true, MathFast>;
    k.fn[1][0] = &launch_chain<Spec, false, MathAccurate>;
    k.fn[1][1] = &launch_chain<Spec, true, MathAccurate>;
    k.P = Spec::P();
    if constexpr (Spec::P() > 0) {
      k.fnw[0][0] = &launch_chain_w<Spec, false, MathFast>;
      k.fnw[0][1] = &launch_chain_w<Spec, true, MathFast>;
      k.fnw[1][0] = &launch_chain_w<Spec, false, MathAccurate>;
      k.fnw[1][1] = &launch_chain_w<Spec, true, MathAccurate>;
    }
    uint8_t types[Spec::KA];
    for (int i = 0; i < Spec::K; ++i) types[i] = (uint8_t)Spec::type(i);
    register_chain(chain_key(Spec::D, Spec::BASE, Spec::K, types), k);
  }
};

// ------------------------------------------------------------------ fused Dense(P) + chain
typedef cudaError_t (*DenseLaunchFn)(const DenseArgs&, cudaStream_t);
struct DenseKernels {
  DenseLaunchFn fn[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [math mode][bwd]  mma.sync version
  DenseLaunchFn fn5[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [math mode][bwd]  tcgen05 / TMEM version
};
void register_dense(const std::string& key, const DenseKernels& k);
const DenseKernels* find_dense(const std::string& key);

template <class Spec, int H, bool BWD, class M>
cudaError_t launch_dense(const DenseArgs& a, cudaStream_t st) {
  constexpr int T = 128;
  constexpr unsigned kSmem = dense_smem_bytes(Spec::P(), H, T, BWD);
  constexpr int kBySmem = (int)((227u * 1024u) / (kSmem + 1024u));
  constexpr int kWant = BWD ? 2 : 3;
  constexpr int MINB = kBySmem < 1 ? 1 : (kBySmem < kWant ? kBySmem : kWant);
  auto kern = dense_chain_kernel<Spec, H, BWD, M, T, MINB>;
  struct Cfg {
    int device = -1;
    int ctas_per_sm = 0;
  };
  static thread_local Cfg cfg;
  const DeviceInfo& di = device_info();
  if (cfg.device != di.device) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, kSmem);
    if (e != cudaSuccess) return e;
    cfg.ctas_per_sm = occ > 0 ? occ : 1;
    cfg.device = di.device;
  }
  const long long ntiles = (a.B + T - 1) / T;
  long long grid = (long long)di.sm_count * cfg.ctas_per_sm;
  if (grid > ntiles) grid = ntiles;
  kern<<<(unsigned)grid, T, kSmem, st>>>(a);
  count_launch();
  return cudaGetLastError();
}

// tcgen05 / TMEM version (nfn_dense_tc5.cuh): 128 threads per CTA, one TMEM allocation per CTA
template <class Spec, int H, bool BWD, class M>
cudaError_t launch_dense_tc5(const DenseArgs& a, cudaStream_t st) {
  using G = tc5::Geo<Spec::P(), H, BWD>;
  constexpr int T = tc5::kRows;          // rows per tile
  constexpr int NT = tc5::kThreads;      // threads per CTA
  constexpr unsigned kSmem = G::kBytes;
  constexpr int kByTmem = (int)(512u / G::kCols);
#ifdef NFN_TUNE_TC5_MINB
  constexpr int MINB = NFN_TUNE_TC5_MINB;
#else
  constexpr int MINB = tc5::min_blocks(Spec::P(), H, BWD);
#endif
  auto kern = tc5::dense_tc5_kernel<Spec, H, BWD, M, MINB>;
  struct Cfg {
    int device = -1;
    int ctas_per_sm = 0;
  };
  static thread_local Cfg cfg;
  const DeviceInfo& di = device_info();
  if (cfg.device != di.device) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
    if (e != cudaSuccess) return e;
    // the occupancy API reports 1 for kernels that allocate tensor memory: size the grid from our own geometry
#ifdef NFN_TUNE_TC5_MINB
    int occ = MINB < kByTmem ? MINB : kByTmem;
#else
    int occ = tc5::resident_ctas(Spec::P(), H, BWD);
#endif
    if (option(kOptDebug))
      fprintf(stderr, "[nfn_b200] dense_tc5 P=%d H=%d bwd=%d: %d CTAs/SM, TMEM cap %d, smem %u B, %d SMs\n", Spec::P(), H,
              (int)BWD, occ, kByTmem, kSmem, di.sm_count);
    cfg.ctas_per_sm = occ > 0 ? occ : 1;
    cfg.device = di.device;
  }
  const long long ntiles = (a.B + T - 1) / T;
  long long grid = (long long)di.sm_count * cfg.ctas_per_sm;
  if (grid > ntiles) grid = ntiles;
  kern<<<(unsigned)grid, NT, kSmem, st>>>(a);
  count_launch();
  return cudaGetLastError();
}

template <class Spec, int H>
struct DenseRegistrar {
  explicit DenseRegistrar() {
    DenseKernels k;
    k.fn[0][0] = &launch_dense<Spec, H, false, MathFast>;
    k.fn[0][1] = &launch_dense<Spec, H, true, MathFast>;
    k.fn[1][0] = &launch_dense<Spec, H, false, MathAccurate>;
    k.fn[1][1] = &launch_dense<Spec, H, true, MathAccurate>;
    k.fn5[0][0] = &launch_dense_tc5<Spec, H, false, MathFast>;
    k.fn5[0][1] = &launch_dense_tc5<Spec, H, true, MathFast>;
    k.fn5[1][0] = &launch_dense_tc5<Spec, H, false, MathAccurate>;
    k.fn5[1][1] = &launch_dense_tc5<Spec, H, true, MathAccurate>;
    uint8_t types[Spec::KA];
    for (int i = 0; i < Spec::K; ++i) types[i] = (uint8_t)Spec::type(i);
    register_dense(chain_key(Spec::D, Spec::BASE, Spec::K, types) + "|h" + std::to_string(H), k);
  }
};

// ------------------------------------------------------------------ fused Dense(P) + MDN head (mma.sync body)
template <int K, int D, int H, bool BWD, class M>
cudaError_t launch_dense_mdn(const DenseArgs& a, cudaStream_t st) {
  constexpr int T = 128;
  constexpr int P = MdnHead<K, D>::P;
  constexpr unsigned kSmem = dense_smem_bytes(P, H, T, BWD);
  constexpr int kBySmem = (int)((227u * 1024u) / (kSmem + 1024u));
  constexpr int kWant = BWD ? 2 : 3;
  constexpr int MINB = kBySmem < 1 ? 1 : (kBySmem < kWant ? kBySmem : kWant);
  auto kern = dense_mdn_kernel<K, D, H, BWD, M, T, MINB>;
  struct Cfg {
    int device = -1;
    int ctas_per_sm = 0;
  };
  static thread_local Cfg cfg;
  const DeviceInfo& di = device_info();
  if (cfg.device != di.device) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, kSmem);
    if (e != cudaSuccess) return e;
    cfg.ctas_per_sm = occ > 0 ? occ : 1;
    cfg.device = di.device;
  }
  const long long ntiles = (a.B + T - 1) / T;
  long long grid = (long long)di.sm_count * cfg.ctas_per_sm;
  if (grid > ntiles) grid = ntiles;
  kern<<<(unsigned)grid, T, kSmem, st>>>(a);
  count_launch();
  return cudaGetLastError();
}

inline std::string dense_mdn_key(int K, int D, int H) {
  return "mdn|k" + std::to_string(K) + "d" + std::to_string(D) + "|h" + std::to_string(H);
}

template <int K, int D, int H>
struct DenseMdnRegistrar {
  explicit DenseMdnRegistrar() {
    DenseKernels k;
    k.fn[0][0] = &launch_dense_mdn<K, D, H, false, MathFast>;
    k.fn[0][1] = &launch_dense_mdn<K, D, H, true, MathFast>;
    k.fn[1][0] = &launch_dense_mdn<K, D, H, false, MathAccurate>;
    k.fn[1][1] = &launch_dense_mdn<K, D, H, true, MathAccurate>;
    register_dense(dense_mdn_key(K, D, H), k);
  }
};

// ------------------------------------------------------------------ fused Dense(P) + KMN head (mma.sync body)
template <int MC, int D, int H, bool BWD, class M>
cudaError_t launch_dense_kmn(const DenseArgs& a, cudaStream_t st) {
  constexpr int T = 128;
  constexpr unsigned kSmem = dense_smem_bytes(MC, H, T, BWD, kmn_extra_floats(MC, D, T, BWD));
  constexpr int kBySmem = (int)((227u * 1024u) / (kSmem + 1024u));
  constexpr int kWant = BWD ? 2 : 3;
  constexpr int MINB = kBySmem < 1 ? 1 : (kBySmem < kWant ? kBySmem : kWant);
  auto kern = dense_kmn_kernel<MC, D, H, BWD, M, T, MINB>;
  struct Cfg {
    int device = -1;
    int ctas_per_sm = 0;
  };
  static thread_local Cfg cfg;
  const DeviceInfo& di = device_info();
  if (cfg.device != di.device) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, kSmem);
    if (e != cudaSuccess) return e;
    cfg.ctas_per_sm = occ > 0 ? occ : 1;
    cfg.device = di.device;
  }
  const long long ntiles = (a.B + T - 1) / T;
  long long grid = (long long)di.sm_count * cfg.ctas_per_sm;
  if (grid > ntiles) grid = ntiles;
  kern<<<(unsigned)grid, T, kSmem, st>>>(a);
  count_launch();
  return cudaGetLastError();
}

inline std::string dense_kmn_key(int MC, int D, int H) {
  return "kmn|m" + std::to_string(MC) + "d" + std::to_string(D) + "|h" + std::to_string(H);
}

template <int MC, int D, int H>
struct DenseKmnRegistrar {
  explicit DenseKmnRegistrar() {
    DenseKernels k;
    k.fn[0][0] = &launch_dense_kmn<MC, D, H, false, MathFast>;
    k.fn[0][1] = &launch_dense_kmn<MC, D, H, true, MathFast>;
    k.fn[1][0] = &launch_dense_kmn<MC, D, H, false, MathAccurate>;
    k.fn[1][1] = &launch_dense_kmn<MC, D, H, true, MathAccurate>;
    register_dense(dense_kmn_key(MC, D, H), k);
  }
};

cudaError_t launch_dense_jit(const nfn_chain_desc* desc, int H, const std::string& key, const DenseArgs& a,
                             bool bwd, int mode, cudaStream_t st, bool* served);
cudaError_t launch_dense_kmn_jit(int MC, int D, int H, const DenseArgs& a, bool bwd, int mode, cudaStream_t st,
                                 bool* served);
long long jit_dense_kmn_compile_check(int MC, int D, int H, int mode, std::string& log);
cudaError_t launch_dense_mdn_jit(int K, int D, int H, const DenseArgs& a, bool bwd, int mode, cudaStream_t st,
                                 bool* served);
long long jit_dense_mdn_compile_check(int K, int D, int H, int mode, std::string& log);
cudaError_t launch_dense_tc5_jit(const nfn_chain_desc* desc, int H, const std::string& key, const DenseArgs& a,
                                 bool bwd, int mode, cudaStream_t st, bool* served);

// runtime specialiser (nfn_jit.cu): NVRTC-compiled chain_kernel for chains without an AOT
// instance.  Returns cudaErrorNotSupported when the chain should go to the generic kernel.
cudaError_t launch_chain_jit(const nfn_chain_desc* desc, const std::string& key, const ChainArgs& a, bool bwd,
                             int mode, cudaStream_t st, bool* served);

// peer-memory communicator (nfn_peer.cu)
PeerArgs make_peer_args(::nfn_peer_comm* c, double* out);   // reads the communicator, changes nothing
void peer_commit(::nfn_peer_comm* c, double* out);          // after a successful launch: advance the sequence
int launch_peer_allreduce(const PeerArgs& p, cudaStream_t st);

// mixture heads (nfn_mixture.cu)
struct MixArgs {
  const float* t;
  const float* y;
  const float* g_logp;
  const float* locs;    // KMN only
  const float* scales;  // KMN only
  float* logp;
  float* dt;
  float* dy;
  float* dscales;       // KMN only
  double* logp_sum;
  double* dt_colsum;
  long long B;
  float g_scale;
  int y_broadcast;
  int K;                // components
  EventXform xf;
};
int launch_mdn(int d, bool bwd, const MixArgs& a, cudaStream_t st);
int launch_kmn(int d, bool bwd, const MixArgs& a, cudaStream_t st);
int launch_logmeanexp(const float* in, long long S, long long B, float* out, cudaStream_t st);
int launch_colsum(const float* dt, long long B, int P, double* out, cudaStream_t st);

// hidden layers of the conditioning network (nfn_mlp.cu)
int mlp_layer_supported(int K, int N, int act);
// xmean / xstd (device, [K], nullable together): the estimators' input normalisation fused into the first layer
int launch_dense_act_forward(const float* x, const float* xmean, const float* xstd, const float* w, const float* b,
                             float* out, long long B, int K, int N, int act, cudaStream_t st);
int launch_dense_act_backward(const float* x, const float* xmean, const float* xstd, const float* out, const float* dout,
                              const float* w, float* dx, float* dW, float* db, long long B, int K, int N, int act,
                              cudaStream_t st);

// first variational layer with S folded weight draws (nfn_mlp.cu)
int mlp_draws_supported(int K, int N, int NP, int act);
int launch_dense_act_draws(bool bwd, const float* x, const float* xmean, const float* xstd, const float* w,
                           const float* out_in, const float* dout, float* out, float* dw, int S, long long Bd, int K, int N,
                           int NP, int act, cudaStream_t st);

// mean-field weight posterior: S samples + exact KL, and their gradient (nfn_variational.cu)
int launch_variational(bool bwd, const float* params, const float* prior_loc, float prior_scale, const float* eps,
                       const float* dw, const float* gkl, float gkl_value, int n, int S, float* w, double* kl, float* dparams,
                       float* dprior_loc, cudaStream_t st);

}  // namespace nfn
