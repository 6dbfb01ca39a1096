// nfn_api.cu -- the C ABI of libnfn_b200.so (include/nfn_b200.h): descriptor validation,
// kernel registry and dispatch, and the chunked host-buffer pipeline.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <unordered_map>
#include <vector>

#include <cuda.h>

#include "nfn_common.h"

namespace nfn {

cudaError_t launch_chain_generic(const nfn_chain_desc* desc, const ChainArgs& a, bool bwd, int mode,
                                 cudaStream_t st);
cudaError_t launch_flow_single(int type, int d, const float* t, const float* z, int zb, float* zo,
                               float* f, long long B, cudaStream_t st);

// ------------------------------------------------------------------ errors, counters
static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_error(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return NFN_OK;
  return set_error(NFN_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
}

void count_launch() { ++g_launches; }

const DeviceInfo& device_info() {
  static thread_local DeviceInfo info;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = -1;
  if (dev != info.device) {
    info.device = dev;
    info.sm_count = 148;
    info.smem_optin = 227 * 1024;
    if (dev >= 0) {
      cudaDeviceGetAttribute(&info.sm_count, cudaDevAttrMultiProcessorCount, dev);
      cudaDeviceGetAttribute(&info.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    }
  }
  return info;
}

// ------------------------------------------------------------------ options
// Every switch is read from the environment ONCE (first use) and can be changed afterwards with
// nfn_set_option(); a launch never calls getenv.  `g_opt_gen` invalidates the per-thread descriptor cache.
static int g_opt[kOptCount];
static std::once_flag g_opt_once;
static std::atomic<unsigned> g_opt_gen{1};

static bool env_is(const char* name, const char* a, const char* b = nullptr) {
  const char* e = getenv(name);
  return e && (!strcmp(e, a) || (b && !strcmp(e, b)));
}

static void init_options() {
  g_opt[kOptMath] = env_is("NFN_B200_MATH", "accurate", "1") ? 1 : 0;
  g_opt[kOptForceGeneric] = getenv("NFN_B200_FORCE_GENERIC") ? 1 : 0;
  g_opt[kOptForceJit] = getenv("NFN_B200_FORCE_JIT") ? 1 : 0;
  g_opt[kOptJit] = env_is("NFN_B200_JIT", "0") ? 0 : 1;
  g_opt[kOptChainIo] = env_is("NFN_B200_CHAIN_IO", "tma", "bulk") ? 1 : (env_is("NFN_B200_CHAIN_IO", "cpasync", "ldgsts") ? 0 : -1);
  g_opt[kOptDenseMma] = env_is("NFN_B200_DENSE_MMA", "tc5") ? 1 : (env_is("NFN_B200_DENSE_MMA", "sync") ? 2 : 0);
  g_opt[kOptPdl] = env_is("NFN_B200_PDL", "0") ? 0 : 1;
  g_opt[kOptDebug] = getenv("NFN_B200_DEBUG") ? 1 : 0;
  long mb = 16;
  if (const char* ev = getenv("NFN_B200_HOST_CHUNK_MB")) {
    const long v = atol(ev);
    if (v >= 1 && v <= 1024) mb = v;
  }
  g_opt[kOptHostChunkMb] = (int)mb;
  g_opt[kOptTuneWnb] = getenv("NFN_B200_TUNE_WNB") ? atoi(getenv("NFN_B200_TUNE_WNB")) : 0;
  g_opt[kOptTuneWwarps] = getenv("NFN_B200_TUNE_WWARPS") ? atoi(getenv("NFN_B200_TUNE_WWARPS")) : 0;
  g_opt[kOptMlpMma] = env_is("NFN_B200_MLP_MMA", "0") ? 0 : 1;
}

int option(Opt o) {
  std::call_once(g_opt_once, init_options);
  return g_opt[o];
}

static const char* const kOptNames[kOptCount] = {"math", "force_generic", "force_jit", "jit", "chain_io",
                                                 "dense_mma", "pdl", "debug", "host_chunk_mb", "tune_wnb", "tune_wwarps",
                                                 "mlp_mma"};

int math_mode() { return option(kOptMath); }
int chain_io_override() { return option(kOptChainIo); }
bool pdl_enabled() { return option(kOptPdl) != 0; }

int encode_row_tensor_map(TensorMap* out, const float* base, long long rows, int P, int W) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static const EncodeFn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return (EncodeFn)fn;
  }();
  if (!encode) return set_error(NFN_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  if (rows < 1 || rows >= (1ll << 31)) return set_error(NFN_ERR_SHAPE, "B=%lld outside the tensor map's range", rows);
  static_assert(sizeof(CUtensorMap) == sizeof(TensorMap), "CUtensorMap is 128 bytes");
  const CUtensorMapSwizzle swz = W == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : (W == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  const cuuint64_t dims[2] = {(cuuint64_t)P, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)P * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)W, 32u};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2u, const_cast<float*>(base),
                            dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(NFN_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d) for [%lld, %d], box [32, %d]", (int)r, rows, P, W);
  return NFN_OK;
}

// ------------------------------------------------------------------ registry
static std::unordered_map<std::string, ChainKernels>& registry() {
  static std::unordered_map<std::string, ChainKernels> r;
  return r;
}

std::string chain_key(int d, bool base, int k, const uint8_t* types) {
  std::string s = "d" + std::to_string(d) + (base ? "b1:" : "b0:");
  for (int i = 0; i < k; ++i) s.push_back("pra"[types[i]]);
  return s;
}

void register_chain(const std::string& key, const ChainKernels& k) { registry()[key] = k; }

int jit_cache_size();
long long jit_compile_check(const nfn_chain_desc* desc, int mode, std::string& log);

const ChainKernels* find_chain(const std::string& key) {
  // test / tuning switches: FORCE_GENERIC skips both specialised paths, FORCE_JIT skips the
  // ahead-of-time instances so that the runtime specialiser serves every chain
  if (option(kOptForceGeneric) || option(kOptForceJit)) return nullptr;
  auto it = registry().find(key);
  return it == registry().end() ? nullptr : &it->second;
}

static std::unordered_map<std::string, DenseKernels>& dense_registry() {
  static std::unordered_map<std::string, DenseKernels> r;
  return r;
}
void register_dense(const std::string& key, const DenseKernels& k) { dense_registry()[key] = k; }
const DenseKernels* find_dense(const std::string& key) {
  if (option(kOptForceJit)) return nullptr;
  auto it = dense_registry().find(key);
  return it == dense_registry().end() ? nullptr : &it->second;
}
long long jit_dense_compile_check(const nfn_chain_desc* desc, int H, int mode, std::string& log);
long long jit_dense_tc5_compile_check(const nfn_chain_desc* desc, int H, int mode, std::string& log);

// ------------------------------------------------------------------ validation
static int check_desc(const nfn_chain_desc* d) {
  if (!d) return set_error(NFN_ERR_NULL, "chain descriptor is NULL");
  if (d->n_dims < 1 || d->n_dims > NFN_MAX_DIMS)
    return set_error(NFN_ERR_DESC, "n_dims=%d outside 1..%d", d->n_dims, NFN_MAX_DIMS);
  if (d->n_flows < 0 || d->n_flows > NFN_MAX_FLOWS)
    return set_error(NFN_ERR_DESC, "n_flows=%d outside 0..%d", d->n_flows, NFN_MAX_FLOWS);
  if (d->trainable_base != 0 && d->trainable_base != 1)
    return set_error(NFN_ERR_DESC, "trainable_base=%d is not 0/1", d->trainable_base);
  for (int k = 0; k < d->n_flows; ++k)
    if (d->flow_type[k] > NFN_FLOW_AFFINE)
      return set_error(NFN_ERR_DESC, "flow_type[%d]=%d is not planar(0)/radial(1)/affine(2)", k,
                       (int)d->flow_type[k]);
  return NFN_OK;
}

static int param_size(const nfn_chain_desc* d) {
  int p = d->trainable_base ? 2 * d->n_dims : 0;
  for (int k = 0; k < d->n_flows; ++k) p += flow_param_size(d->flow_type[k], d->n_dims);
  return p;
}

static bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

static size_t event_align(int d) { return d == 4 ? 16 : (d == 2 ? 8 : 4); }

static int check_rows(int64_t B, int64_t y_rows) {
  if (B < 0) return set_error(NFN_ERR_SHAPE, "B=%lld is negative", (long long)B);
  if (y_rows != B && y_rows != 1)
    return set_error(NFN_ERR_SHAPE, "y_rows=%lld must equal B=%lld or 1", (long long)y_rows,
                     (long long)B);
  return NFN_OK;
}

// Per-thread memo of the last descriptor's registry lookup: a launch of the same chain as the previous one
// (the normal case) builds no key string and touches no map.
struct DescMemo {
  nfn_chain_desc desc;
  unsigned gen = 0;
  std::string key;
  const ChainKernels* kernels = nullptr;
};
static thread_local DescMemo g_memo;

static const DescMemo& lookup_chain(const nfn_chain_desc* desc) {
  DescMemo& m = g_memo;
  const unsigned gen = g_opt_gen.load(std::memory_order_relaxed);
  const size_t used = offsetof(nfn_chain_desc, flow_type) + (size_t)desc->n_flows;
  if (m.gen != gen || m.desc.n_flows != desc->n_flows || memcmp(&m.desc, desc, used) != 0) {
    memset(&m.desc, 0, sizeof(m.desc));
    memcpy(&m.desc, desc, used);
    m.key = chain_key(desc->n_dims, desc->trainable_base != 0, desc->n_flows, desc->flow_type);
    m.kernels = find_chain(m.key);
    m.gen = gen;
  }
  return m;
}

static_assert(sizeof(nfn_event_xform) == sizeof(EventXform), "nfn_event_xform mirrors nfn::EventXform");

// validates the caller's event transform for an n_dims-dimensional head and copies it into the kernel arguments
static int set_xform(EventXform& dst, const nfn_event_xform* xf, int n_dims) {
  memset(&dst, 0, sizeof(dst));
  if (!xf) return NFN_OK;
  if (xf->flags & ~(NFN_XF_NORMALISE | NFN_XF_NOISE | NFN_XF_EXP))
    return set_error(NFN_ERR_DESC, "event transform: unknown flag bits 0x%x", xf->flags);
  if (xf->flags & NFN_XF_NORMALISE)
    for (int i = 0; i < n_dims; ++i)
      if (!(xf->std[i] != 0.0f)) return set_error(NFN_ERR_DESC, "event transform: std[%d] is zero or NaN", i);
  if ((xf->flags & NFN_XF_NOISE) && !(xf->noise_std >= 0.0f))
    return set_error(NFN_ERR_DESC, "event transform: noise_std=%g", (double)xf->noise_std);
  memcpy(&dst, xf, sizeof(dst));
  return NFN_OK;
}

static int chain_dispatch(const nfn_chain_desc* desc, const ChainArgs& a, bool bwd, cudaStream_t st) {
  const DescMemo& memo = lookup_chain(desc);
  const std::string& key = memo.key;
  const int mode = math_mode();
  const ChainKernels* k = memo.kernels;
  cudaError_t e;
  // Split-phase peer exchange: the warp-tile kernels carry it in one CTA of their own grid; every other kernel
  // runs without a peer epilogue and is followed by the one-CTA exchange launch.
  const bool split = a.peer.world > 0 && a.peer.deferred;
  const int io = chain_io_override();
  const int P = param_size(desc);
  const bool want_w = P > 0 && (io >= 0 ? io == 1 : chain_prefers_warp_tile(P, bwd));
  if (k && k->fn[mode][bwd ? 1 : 0]) {
    const bool warp_tile = k->fnw[mode][bwd ? 1 : 0] && want_w;
    if (split && !warp_tile) {
      ChainArgs c = a;
      c.peer.world = 0;
      e = k->fn[mode][bwd ? 1 : 0](c, st);
      if (e == cudaSuccess) {
        int rc = launch_peer_allreduce(a.peer, st);
        if (rc != NFN_OK) return rc;
      }
    } else {
      e = (warp_tile ? k->fnw : k->fn)[mode][bwd ? 1 : 0](a, st);
    }
  } else {
    if (!option(kOptForceGeneric)) {
      bool served = false;
      ChainArgs c = a;
      if (split && !want_w) c.peer.world = 0;
      e = launch_chain_jit(desc, key, c, bwd, mode, st, &served);
      if (e == cudaSuccess && served && split && !want_w) {
        int rc = launch_peer_allreduce(a.peer, st);
        if (rc != NFN_OK) return rc;
      }
      if (e != cudaSuccess || served) return cuda_error(e, key.c_str());
    }
    ChainArgs g = a;
    g.dt_colsum = nullptr;  // the generic kernel leaves the column sums to a second pass
    e = launch_chain_generic(desc, g, bwd, mode, st);
    if (e == cudaSuccess && bwd && a.dt_colsum) {
      int rc = launch_colsum(a.dt, a.B, param_size(desc), a.dt_colsum, st);
      if (rc != NFN_OK) return rc;
    }
    if (e == cudaSuccess && a.peer.world > 0) {  // no fused epilogue in the generic kernel
      int rc = launch_peer_allreduce(a.peer, st);
      if (rc != NFN_OK) return rc;
    }
  }
  return cuda_error(e, key.c_str());
}

}  // namespace nfn

using namespace nfn;

// =================================================================== C ABI
extern "C" {

int nfn_version(void) { return NFN_B200_VERSION; }

const char* nfn_last_error(void) { return g_err; }

int64_t nfn_launch_count_reset(void) {
  const long long n = g_launches;
  g_launches = 0;
  return n;
}

int nfn_set_math_mode(int accurate) { return nfn_set_option("math", accurate ? 1 : 0); }

int nfn_set_option(const char* name, int value) {
  if (!name) return set_error(NFN_ERR_NULL, "option name is NULL");
  option(kOptMath);  // environment defaults first
  for (int i = 0; i < kOptCount; ++i)
    if (!strcmp(name, kOptNames[i])) {
      g_opt[i] = value;
      g_opt_gen.fetch_add(1);
      return NFN_OK;
    }
  return set_error(NFN_ERR_DESC, "unknown option '%s'", name);
}

int nfn_get_option(const char* name) {
  if (!name) return set_error(NFN_ERR_NULL, "option name is NULL");
  for (int i = 0; i < kOptCount; ++i)
    if (!strcmp(name, kOptNames[i])) return option((Opt)i);
  return set_error(NFN_ERR_DESC, "unknown option '%s'", name);
}

int nfn_chain_param_size(const nfn_chain_desc* desc) {
  int rc = check_desc(desc);
  if (rc != NFN_OK) return rc;
  return param_size(desc);
}

int nfn_jit_cache_size(void) { return jit_cache_size(); }

int64_t nfn_jit_compile_check(const nfn_chain_desc* desc, int accurate) {
  int rc = check_desc(desc);
  if (rc != NFN_OK) return rc;
  std::string log;
  const long long n = jit_compile_check(desc, accurate ? 1 : 0, log);
  if (n < 0) return set_error(NFN_ERR_UNSUPPORTED, "NVRTC: %s", log.substr(0, 400).c_str());
  return n;
}

int nfn_chain_is_specialized(const nfn_chain_desc* desc) {
  int rc = check_desc(desc);
  if (rc != NFN_OK) return rc;
  return lookup_chain(desc).kernels ? 1 : 0;
}

int nfn_chain_forward(const nfn_chain_desc* desc, const float* t, const float* y, int64_t y_rows,
                      float* logp, int64_t B, void* stream) {
  return nfn_chain_forward_x(desc, t, y, y_rows, logp, B, nullptr, stream);
}

int nfn_chain_forward_x(const nfn_chain_desc* desc, const float* t, const float* y, int64_t y_rows,
                        float* logp, int64_t B, const nfn_event_xform* xf, void* stream) {
  int rc = check_desc(desc);
  if (rc != NFN_OK) return rc;
  if ((rc = check_rows(B, y_rows)) != NFN_OK) return rc;
  if (B == 0) return NFN_OK;
  const int P = param_size(desc);
  if (!y || !logp || (P > 0 && !t)) return set_error(NFN_ERR_NULL, "t, y and logp must be non-NULL");
  if (!aligned(t, 16)) return set_error(NFN_ERR_ALIGN, "t must be 16-byte aligned");
  if (!aligned(y, event_align(desc->n_dims)))
    return set_error(NFN_ERR_ALIGN, "y must be %zu-byte aligned", event_align(desc->n_dims));
  ChainArgs a{};
  a.t = t; a.y = y; a.logp = logp; a.B = B; a.g_scale = 1.0f; a.y_broadcast = (y_rows == 1 && B != 1);
  if ((rc = set_xform(a.xf, xf, desc->n_dims)) != NFN_OK) return rc;
  return chain_dispatch(desc, a, false, (cudaStream_t)stream);
}

int nfn_chain_forward_grid(const nfn_chain_desc* desc, const float* t, const float* y_grid, int64_t n_y,
                           float* logp, int64_t B, void* stream) {
  return nfn_chain_forward_grid_x(desc, t, y_grid, n_y, logp, B, nullptr, stream);
}

int nfn_chain_forward_grid_x(const nfn_chain_desc* desc, const float* t, const float* y_grid, int64_t n_y,
                             float* logp, int64_t B, const nfn_event_xform* xf, void* stream) {
  int rc = check_desc(desc);
  if (rc != NFN_OK) return rc;
  if (B < 0 || n_y < 0 || n_y > (1 << 30)) return set_error(NFN_ERR_SHAPE, "B=%lld, n_y=%lld", (long long)B, (long long)n_y);
  if (B == 0 || n_y == 0) return NFN_OK;
  const int P = param_size(desc);
  if (!y_grid || !logp || (P > 0 && !t)) return set_error(NFN_ERR_NULL, "t, y_grid and logp must be non-NULL");
  if (!aligned(t, 16)) return set_error(NFN_ERR_ALIGN, "t must be 16-byte aligned");
  if (!aligned(y_grid, event_align(desc->n_dims)))
    return set_error(NFN_ERR_ALIGN, "y_grid must be %zu-byte aligned", event_align(desc->n_dims));
  ChainArgs a{};
  a.t = t; a.y = y_grid; a.logp = logp; a.B = B; a.g_scale = 1.0f; a.y_broadcast = 1; a.grid_ny = (int)n_y;
  if ((rc = set_xform(a.xf, xf, desc->n_dims)) != NFN_OK) return rc;
  return chain_dispatch(desc, a, false, (cudaStream_t)stream);
}

int nfn_chain_forward_backward(const nfn_chain_desc* desc, const float* t, const float* y,
                               int64_t y_rows, const float* g_logp, float g_scale, float* logp,
                               float* dt, float* dy, double* logp_sum, double* dt_colsum, int64_t B,
                               void* stream) {
  return nfn_chain_forward_backward_x(desc, t, y, y_rows, g_logp, g_scale, logp, dt, dy, logp_sum, dt_colsum, B, nullptr,
                                      stream);
}

int nfn_chain_forward_backward_x(const nfn_chain_desc* desc, const float* t, const float* y,
                                 int64_t y_rows, const float* g_logp, float g_scale, float* logp,
                                 float* dt, float* dy, double* logp_sum, double* dt_colsum, int64_t B,
                                 const nfn_event_xform* xf, void* stream) {
  int rc = check_desc(desc);
  if (rc != NFN_OK) return rc;
  if ((rc = check_rows(B, y_rows)) != NFN_OK) return rc;
  if (B == 0) return NFN_OK;
  const int P = param_size(desc);
  if (!y || !logp || (P > 0 && (!t || !dt)))
    return set_error(NFN_ERR_NULL, "t, y, logp and dt must be non-NULL");
  if (!aligned(t, 16) || !aligned(dt, 16))
    return set_error(NFN_ERR_ALIGN, "t and dt must be 16-byte aligned");
  if (!aligned(y, event_align(desc->n_dims)))
    return set_error(NFN_ERR_ALIGN, "y must be %zu-byte aligned", event_align(desc->n_dims));
  if (dy && y_rows != B) return set_error(NFN_ERR_SHAPE, "dy requires y_rows == B");
  ChainArgs a{};
  a.t = t; a.y = y; a.g_logp = g_logp; a.logp = logp; a.dt = dt; a.dy = dy;
  a.logp_sum = logp_sum; a.dt_colsum = dt_colsum; a.B = B; a.g_scale = g_scale;
  a.y_broadcast = (y_rows == 1 && B != 1);
  if ((rc = set_xform(a.xf, xf, desc->n_dims)) != NFN_OK) return rc;
  return chain_dispatch(desc, a, true, (cudaStream_t)stream);
}

int nfn_chain_forward_backward_peer(const nfn_chain_desc* desc, const float* t, const float* y,
                                    int64_t y_rows, const float* g_logp, float g_scale, float* logp,
                                    float* dt, float* dy, int want_colsum, nfn_peer_comm* comm,
                                    double* reduced, int64_t B, void* stream) {
  int rc = check_desc(desc);
  if (rc != NFN_OK) return rc;
  if ((rc = check_rows(B, y_rows)) != NFN_OK) return rc;
  if (!comm || !reduced) return set_error(NFN_ERR_NULL, "comm and reduced must be non-NULL");
  const int P = param_size(desc);
  if (B > 0 && (!y || !logp || (P > 0 && (!t || !dt))))
    return set_error(NFN_ERR_NULL, "t, y, logp and dt must be non-NULL");
  if (!aligned(t, 16) || !aligned(dt, 16))
    return set_error(NFN_ERR_ALIGN, "t and dt must be 16-byte aligned");
  if (!aligned(y, event_align(desc->n_dims)))
    return set_error(NFN_ERR_ALIGN, "y must be %zu-byte aligned", event_align(desc->n_dims));
  if (dy && y_rows != B) return set_error(NFN_ERR_SHAPE, "dy requires y_rows == B");
  ChainArgs a{};
  a.peer = make_peer_args(comm, reduced);
  if (a.peer.n_values != P + 1)
    return set_error(NFN_ERR_SHAPE, "communicator carries %d values, the chain needs P + 1 = %d", a.peer.n_values,
                     P + 1);
  if (B == 0) {  // nothing local to add, but every rank must still take part in the exchange
    rc = launch_peer_allreduce(a.peer, (cudaStream_t)stream);
  } else {
    a.t = t; a.y = y; a.g_logp = g_logp; a.logp = logp; a.dt = dt; a.dy = dy;
    a.logp_sum = a.peer.acc + P;
    a.dt_colsum = want_colsum ? a.peer.acc : nullptr;
    a.B = B; a.g_scale = g_scale; a.y_broadcast = (y_rows == 1 && B != 1);
    rc = chain_dispatch(desc, a, true, (cudaStream_t)stream);
  }
  // the exchange's sequence number advances only once the launch is in the stream: an error return leaves this
  // rank in step with its peers
  if (rc == NFN_OK) peer_commit(comm, reduced);
  return rc;
}

// ------------------------------------------------------------------ fused Dense(P) + chain
static int dense_dispatch(const nfn_chain_desc* desc, int hidden, const DenseArgs& a, bool bwd, cudaStream_t st) {
  const std::string key = chain_key(desc->n_dims, desc->trainable_base != 0, desc->n_flows, desc->flow_type);
  const int mode = math_mode();
  const DenseKernels* k = find_dense(key + "|h" + std::to_string(hidden));
  // Two implementations of the same contract: tcgen05 / TMEM (nfn_dense_tc5.cuh, the default wherever it
  // exists) and warp-level mma.sync (nfn_dense_chain.cuh: the baseline it is measured against, and the
  // fallback if the tcgen05 kernel cannot be built for a chain).  Measured, B = 2^20, H = 16, tcgen05 vs mma.sync:
  //   fwd+bwd  P = 48: 88-92 vs 172 us   P = 32: 73 vs 131 us   P = 17: 68 vs 87 us   P = 11: 57 vs 66 us
  //   forward  P = 48: 33 vs  72 us   P = 32: 33 vs  53 us   P = 17: 27 vs 40 us   P = 11: 24 vs 31 us
  // Small launches are latency-bound and the tcgen05 kernel has the longer prologue (TMEM allocation, split
  // weight tiles, a one-tile pipeline delay): B = 2048..16384 rows 8.4 vs 6.2 us (P = 11), 10.8 vs 10.3 us
  // (P = 48); the crossover tracks rows x parameters, so tcgen05 takes launches with B * P >= 4 M (fwd+bwd) /
  // 1 M (forward).  NFN_B200_DENSE_MMA=tc5|sync forces one of them (A/B comparisons, tests).
  const bool force5 = option(kOptDenseMma) == 1, force_sync = option(kOptDenseMma) == 2;
  const bool want5 = force5 || (!force_sync && a.B * (long long)param_size(desc) >= (bwd ? (4ll << 20) : (1ll << 20)));
  if (want5 && k && k->fn5[mode][bwd ? 1 : 0]) return cuda_error(k->fn5[mode][bwd ? 1 : 0](a, st), key.c_str());
  bool served = false;
  if (want5 && !(k && k->fn[mode][bwd ? 1 : 0] && !force5)) {
    // no ahead-of-time instance of either kind (or tc5 forced): runtime-specialise the tcgen05 kernel
    cudaError_t e5 = launch_dense_tc5_jit(desc, hidden, key, a, bwd, mode, st, &served);
    if (e5 != cudaSuccess) return cuda_error(e5, key.c_str());
    if (served) return NFN_OK;
  }
  if (k && k->fn[mode][bwd ? 1 : 0]) return cuda_error(k->fn[mode][bwd ? 1 : 0](a, st), key.c_str());
  cudaError_t e = launch_dense_jit(desc, hidden, key, a, bwd, mode, st, &served);
  if (e != cudaSuccess) return cuda_error(e, key.c_str());
  if (!served)
    return set_error(NFN_ERR_UNSUPPORTED,
                     "no fused dense kernel for chain %s with hidden width %d (needs a multiple of 16 <= 64 and an "
                     "ahead-of-time instance or NVRTC): compose the layer and nfn_chain_forward_backward instead",
                     key.c_str(), hidden);
  return NFN_OK;
}

static int dense_common(const nfn_chain_desc* desc, int hidden, const float* h, const float* W, const float* bias,
                        const float* y, int64_t y_rows, float* logp, int64_t B) {
  int rc = check_desc(desc);
  if (rc != NFN_OK) return rc;
  if ((rc = check_rows(B, y_rows)) != NFN_OK) return rc;
  if (hidden < 1) return set_error(NFN_ERR_SHAPE, "hidden=%d", hidden);
  if (param_size(desc) < 1) return set_error(NFN_ERR_UNSUPPORTED, "the chain has no parameters to emit");
  if (B == 0) return 1;
  if (!h || !W || !bias || !y || !logp) return set_error(NFN_ERR_NULL, "h, W, bias, y and logp must be non-NULL");
  if (!aligned(h, 16)) return set_error(NFN_ERR_ALIGN, "h must be 16-byte aligned");
  if (!aligned(y, event_align(desc->n_dims)))
    return set_error(NFN_ERR_ALIGN, "y must be %zu-byte aligned", event_align(desc->n_dims));
  return NFN_OK;
}

int nfn_dense_chain_forward(const nfn_chain_desc* desc, int hidden, const float* h, const float* W,
                            const float* bias, const float* y, int64_t y_rows, float* logp, int64_t B,
                            void* stream) {
  return nfn_dense_chain_forward_x(desc, hidden, h, W, bias, y, y_rows, logp, B, nullptr, stream);
}

int nfn_dense_chain_forward_x(const nfn_chain_desc* desc, int hidden, const float* h, const float* W,
                              const float* bias, const float* y, int64_t y_rows, float* logp, int64_t B,
                              const nfn_event_xform* xf, void* stream) {
  int rc = dense_common(desc, hidden, h, W, bias, y, y_rows, logp, B);
  if (rc != NFN_OK) return rc > 0 ? NFN_OK : rc;
  DenseArgs a{};
  a.h = h; a.W = W; a.bias = bias; a.y = y; a.logp = logp; a.B = B; a.g_scale = 1.0f;
  a.y_broadcast = (y_rows == 1 && B != 1);
  if ((rc = set_xform(a.xf, xf, desc->n_dims)) != NFN_OK) return rc;
  return dense_dispatch(desc, hidden, a, false, (cudaStream_t)stream);
}

int nfn_dense_chain_forward_backward(const nfn_chain_desc* desc, int hidden, const float* h, const float* W,
                                     const float* bias, const float* y, int64_t y_rows, const float* g_logp,
                                     float g_scale, float* logp, float* dh, float* dW, float* dbias,
                                     double* logp_sum, int64_t B, void* stream) {
  return nfn_dense_chain_forward_backward_x(desc, hidden, h, W, bias, y, y_rows, g_logp, g_scale, logp, dh, dW, dbias,
                                            logp_sum, B, nullptr, stream);
}

int nfn_dense_chain_forward_backward_x(const nfn_chain_desc* desc, int hidden, const float* h, const float* W,
                                       const float* bias, const float* y, int64_t y_rows, const float* g_logp,
                                       float g_scale, float* logp, float* dh, float* dW, float* dbias,
                                       double* logp_sum, int64_t B, const nfn_event_xform* xf, void* stream) {
  int rc = dense_common(desc, hidden, h, W, bias, y, y_rows, logp, B);
  if (rc != NFN_OK) return rc > 0 ? NFN_OK : rc;
  if (!dh || !dW || !dbias) return set_error(NFN_ERR_NULL, "dh, dW and dbias must be non-NULL");
  if (!aligned(dh, 16)) return set_error(NFN_ERR_ALIGN, "dh must be 16-byte aligned");
  DenseArgs a{};
  a.h = h; a.W = W; a.bias = bias; a.y = y; a.g_logp = g_logp; a.logp = logp; a.dh = dh; a.dW = dW;
  a.dbias = dbias; a.logp_sum = logp_sum; a.B = B; a.g_scale = g_scale;
  a.y_broadcast = (y_rows == 1 && B != 1);
  if ((rc = set_xform(a.xf, xf, desc->n_dims)) != NFN_OK) return rc;
  return dense_dispatch(desc, hidden, a, true, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ folded posterior draws (Bayesian estimators)
// mma.sync body only (its weights are re-staged per draw from a few hundred bytes; the tcgen05 kernel pre-splits
// its weight tiles once per CTA)
static int dense_draws_dispatch(const nfn_chain_desc* desc, int hidden, const DenseArgs& a, bool bwd, cudaStream_t st) {
  const std::string key = chain_key(desc->n_dims, desc->trainable_base != 0, desc->n_flows, desc->flow_type);
  const int mode = math_mode();
  const DenseKernels* k = find_dense(key + "|h" + std::to_string(hidden));
  if (k && k->fn[mode][bwd ? 1 : 0]) return cuda_error(k->fn[mode][bwd ? 1 : 0](a, st), key.c_str());
  bool served = false;
  cudaError_t e = launch_dense_jit(desc, hidden, key, a, bwd, mode, st, &served);
  if (e != cudaSuccess) return cuda_error(e, key.c_str());
  if (!served)
    return set_error(NFN_ERR_UNSUPPORTED, "no fused dense kernel for chain %s with hidden width %d", key.c_str(), hidden);
  return NFN_OK;
}

static int draws_common(int draws, int64_t rows_per_draw, int64_t y_rows) {
  if (draws < 1 || draws > 65535) return set_error(NFN_ERR_SHAPE, "draws=%d outside 1..65535", draws);
  if (rows_per_draw < 0) return set_error(NFN_ERR_SHAPE, "rows_per_draw=%lld", (long long)rows_per_draw);
  if (y_rows != rows_per_draw && y_rows != 1)
    return set_error(NFN_ERR_SHAPE, "y must have rows_per_draw=%lld rows (one per sample, not per folded row) or 1",
                     (long long)rows_per_draw);
  return NFN_OK;
}

int nfn_dense_chain_forward_draws_x(const nfn_chain_desc* desc, int hidden, int draws, int64_t rows_per_draw,
                                    const float* h, const float* W, const float* bias, const float* y, int64_t y_rows,
                                    float* logp, const nfn_event_xform* xf, void* stream) {
  int rc = draws_common(draws, rows_per_draw, y_rows);
  if (rc != NFN_OK) return rc;
  const int64_t B = (int64_t)draws * rows_per_draw;
  rc = dense_common(desc, hidden, h, W, bias, y, 1, logp, B);   // (y rows were checked against rows_per_draw above)
  if (rc != NFN_OK) return rc > 0 ? NFN_OK : rc;
  DenseArgs a{};
  a.h = h; a.W = W; a.bias = bias; a.y = y; a.logp = logp; a.B = B; a.g_scale = 1.0f;
  a.y_broadcast = (y_rows == 1 && rows_per_draw != 1);
  a.draws = draws; a.rows_per_draw = rows_per_draw;
  if ((rc = set_xform(a.xf, xf, desc->n_dims)) != NFN_OK) return rc;
  return dense_draws_dispatch(desc, hidden, a, false, (cudaStream_t)stream);
}

int nfn_dense_chain_forward_backward_draws_x(const nfn_chain_desc* desc, int hidden, int draws, int64_t rows_per_draw,
                                             const float* h, const float* W, const float* bias, const float* y,
                                             int64_t y_rows, const float* g_logp, float g_scale, float* logp, float* dh,
                                             float* dW, float* dbias, double* logp_sum, const nfn_event_xform* xf,
                                             void* stream) {
  int rc = draws_common(draws, rows_per_draw, y_rows);
  if (rc != NFN_OK) return rc;
  const int64_t B = (int64_t)draws * rows_per_draw;
  rc = dense_common(desc, hidden, h, W, bias, y, 1, logp, B);   // (y rows were checked against rows_per_draw above)
  if (rc != NFN_OK) return rc > 0 ? NFN_OK : rc;
  if (!dh || !dW || !dbias) return set_error(NFN_ERR_NULL, "dh, dW and dbias must be non-NULL");
  if (!aligned(dh, 16)) return set_error(NFN_ERR_ALIGN, "dh must be 16-byte aligned");
  DenseArgs a{};
  a.h = h; a.W = W; a.bias = bias; a.y = y; a.g_logp = g_logp; a.logp = logp; a.dh = dh; a.dW = dW;
  a.dbias = dbias; a.logp_sum = logp_sum; a.B = B; a.g_scale = g_scale;
  a.y_broadcast = (y_rows == 1 && rows_per_draw != 1);
  a.draws = draws; a.rows_per_draw = rows_per_draw;
  if ((rc = set_xform(a.xf, xf, desc->n_dims)) != NFN_OK) return rc;
  return dense_draws_dispatch(desc, hidden, a, true, (cudaStream_t)stream);
}

static int act_draws_common(const float* x, const float* x_mean, const float* x_std, int draws, int64_t rows_per_draw,
                            int in_features, int units, int out_width, int act) {
  if (draws < 1 || draws > 65535) return set_error(NFN_ERR_SHAPE, "draws=%d outside 1..65535", draws);
  if (rows_per_draw < 0) return set_error(NFN_ERR_SHAPE, "rows_per_draw=%lld", (long long)rows_per_draw);
  if (!mlp_draws_supported(in_features, units, out_width, act))
    return set_error(NFN_ERR_UNSUPPORTED, "folded first layer %d -> %d (row width %d, act %d) is outside the kernel's range "
                     "(<= 8 inputs, <= 64 units, row width a multiple of 8 <= 64)", in_features, units, out_width, act);
  if (rows_per_draw == 0) return 1;
  if (!x) return set_error(NFN_ERR_NULL, "x must be non-NULL");
  if ((x_mean == nullptr) != (x_std == nullptr)) return set_error(NFN_ERR_NULL, "x_mean and x_std go together");
  return NFN_OK;
}

int nfn_dense_act_forward_draws(const float* x, const float* x_mean, const float* x_std, const float* w, int draws,
                                int64_t rows_per_draw, int in_features, int units, int out_width, int act, float* out,
                                void* stream) {
  int rc = act_draws_common(x, x_mean, x_std, draws, rows_per_draw, in_features, units, out_width, act);
  if (rc != NFN_OK) return rc > 0 ? NFN_OK : rc;
  if (!w || !out) return set_error(NFN_ERR_NULL, "w and out must be non-NULL");
  if (!aligned(out, 16)) return set_error(NFN_ERR_ALIGN, "out must be 16-byte aligned");
  return launch_dense_act_draws(false, x, x_mean, x_std, w, nullptr, nullptr, out, nullptr, draws, rows_per_draw, in_features,
                                units, out_width, act, (cudaStream_t)stream);
}

int nfn_dense_act_backward_draws(const float* x, const float* x_mean, const float* x_std, const float* out,
                                 const float* dout, int draws, int64_t rows_per_draw, int in_features, int units,
                                 int out_width, int act, float* dw, void* stream) {
  int rc = act_draws_common(x, x_mean, x_std, draws, rows_per_draw, in_features, units, out_width, act);
  if (rc != NFN_OK) return rc > 0 ? NFN_OK : rc;
  if (!out || !dout || !dw) return set_error(NFN_ERR_NULL, "out, dout and dw must be non-NULL");
  if (!aligned(out, 16) || !aligned(dout, 16)) return set_error(NFN_ERR_ALIGN, "out and dout must be 16-byte aligned");
  return launch_dense_act_draws(true, x, x_mean, x_std, nullptr, out, dout, nullptr, dw, draws, rows_per_draw, in_features,
                                units, out_width, act, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ mean-field weight posterior (Bayesian estimators)
int nfn_variational_sample(const float* params, const float* prior_loc, float prior_scale, const float* eps, int n,
                           int draws, float* w, double* kl, void* stream) {
  if (n < 1 || draws < 1) return set_error(NFN_ERR_SHAPE, "n=%d draws=%d", n, draws);
  if (!(prior_scale > 0.0f)) return set_error(NFN_ERR_DESC, "prior_scale must be positive");
  if (!params || !prior_loc || !eps || !w) return set_error(NFN_ERR_NULL, "params, prior_loc, eps and w must be non-NULL");
  return launch_variational(false, params, prior_loc, prior_scale, eps, nullptr, nullptr, 0.0f, n, draws, w, kl, nullptr,
                            nullptr, (cudaStream_t)stream);
}

int nfn_variational_sample_backward(const float* params, const float* prior_loc, float prior_scale, const float* eps,
                                    const float* dw, const float* g_kl, int n, int draws, float* dparams,
                                    float* dprior_loc, void* stream) {
  if (n < 1 || draws < 1) return set_error(NFN_ERR_SHAPE, "n=%d draws=%d", n, draws);
  if (!(prior_scale > 0.0f)) return set_error(NFN_ERR_DESC, "prior_scale must be positive");
  if (!params || !prior_loc || !dparams) return set_error(NFN_ERR_NULL, "params, prior_loc and dparams must be non-NULL");
  if (dw && !eps) return set_error(NFN_ERR_NULL, "eps must accompany dw");
  return launch_variational(true, params, prior_loc, prior_scale, eps, dw, g_kl, 0.0f, n, draws, nullptr, nullptr, dparams,
                            dprior_loc, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ one S-draw Bayesian training step, network part
static int dense_mdn_dispatch(int K, int d, int hidden, const DenseArgs& a, bool bwd, cudaStream_t st);

int nfn_bayes_train_step(const nfn_chain_desc* desc, int mdn_centers, int draws, int64_t rows_per_draw, int in_features,
                         int units, int hidden_width, int act, const float* x, const float* x_mean, const float* x_std,
                         const float* y, int64_t y_rows, const nfn_variational_layer* first,
                         const nfn_variational_layer* emitting, float g_scale, float* h, float* dh, float* logp,
                         double* logp_sum, const nfn_event_xform* xf, void* stream) {
  int rc = check_desc(desc);
  if (rc != NFN_OK) return rc;
  if ((rc = draws_common(draws, rows_per_draw, y_rows)) != NFN_OK) return rc;
  if ((rc = act_draws_common(x, x_mean, x_std, draws, rows_per_draw, in_features, units, hidden_width, act)) < 0) return rc;
  if (hidden_width % 16 != 0 || hidden_width < 16 || hidden_width > 64)
    return set_error(NFN_ERR_UNSUPPORTED, "hidden_width=%d must be 16, 32, 48 or 64", hidden_width);
  if (mdn_centers < 0 || mdn_centers > 4096) return set_error(NFN_ERR_DESC, "mdn_centers=%d", mdn_centers);
  const int P = mdn_centers > 0 ? mdn_centers * (2 * desc->n_dims + 1) : param_size(desc);
  if (P < 1) return set_error(NFN_ERR_UNSUPPORTED, "the head has no parameters to emit");
  if (!first || !emitting) return set_error(NFN_ERR_NULL, "both layers must be given");
  if (first->n != in_features * units + units || emitting->n != units * P + P)
    return set_error(NFN_ERR_SHAPE, "layer sizes %d / %d do not match %d -> %d -> %d", first->n, emitting->n, in_features, units, P);
  const nfn_variational_layer* ls[2] = {first, emitting};
  for (const nfn_variational_layer* l : ls) {
    if (!l->posterior || !l->prior_loc || !l->eps || !l->w || !l->dw || !l->dposterior)
      return set_error(NFN_ERR_NULL, "posterior, prior_loc, eps, w, dw and dposterior must be non-NULL");
    if (!(l->prior_scale > 0.0f)) return set_error(NFN_ERR_DESC, "prior_scale must be positive");
  }
  if (rows_per_draw == 0) return NFN_OK;
  if (!y || !h || !dh || !logp) return set_error(NFN_ERR_NULL, "y, h, dh and logp must be non-NULL");
  if (!aligned(h, 16) || !aligned(dh, 16)) return set_error(NFN_ERR_ALIGN, "h and dh must be 16-byte aligned");
  if (!aligned(y, event_align(desc->n_dims)))
    return set_error(NFN_ERR_ALIGN, "y must be %zu-byte aligned", event_align(desc->n_dims));
  cudaStream_t st = (cudaStream_t)stream;
  // 1. the samples and the KL terms
  for (const nfn_variational_layer* l : ls)
    if ((rc = launch_variational(false, l->posterior, l->prior_loc, l->prior_scale, l->eps, nullptr, nullptr, 0.0f, l->n, draws,
                                 l->w, l->kl, nullptr, nullptr, st)) != NFN_OK)
      return rc;
  // 2. first layer over the folded rows
  if ((rc = launch_dense_act_draws(false, x, x_mean, x_std, first->w, nullptr, nullptr, h, nullptr, draws, rows_per_draw,
                                   in_features, units, hidden_width, act, st)) != NFN_OK)
    return rc;
  // 3. emitting layer + head, weights and their gradient in the layer's own flat layout
  cudaError_t ce = cudaMemsetAsync(emitting->dw, 0, sizeof(float) * (size_t)draws * emitting->n, st);
  if (ce == cudaSuccess) ce = cudaMemsetAsync(first->dw, 0, sizeof(float) * (size_t)draws * first->n, st);
  if (ce != cudaSuccess) return cuda_error(ce, "bayes_train_step memset");
  DenseArgs a{};
  a.h = h; a.W = emitting->w; a.bias = emitting->w; a.y = y; a.logp = logp; a.dh = dh; a.dW = emitting->dw;
  a.dbias = emitting->dw; a.logp_sum = logp_sum; a.B = (int64_t)draws * rows_per_draw; a.g_scale = g_scale;
  a.y_broadcast = (y_rows == 1 && rows_per_draw != 1);
  a.draws = draws; a.rows_per_draw = rows_per_draw;
  a.flat_rows = units; a.flat_stride = emitting->n;
  if ((rc = set_xform(a.xf, xf, desc->n_dims)) != NFN_OK) return rc;
  rc = mdn_centers > 0 ? dense_mdn_dispatch(mdn_centers, desc->n_dims, hidden_width, a, true, st)
                       : dense_draws_dispatch(desc, hidden_width, a, true, st);
  if (rc != NFN_OK) return rc;
  // 4. first layer's per-draw weight gradient
  if ((rc = launch_dense_act_draws(true, x, x_mean, x_std, nullptr, h, dh, nullptr, first->dw, draws, rows_per_draw,
                                   in_features, units, hidden_width, act, st)) != NFN_OK)
    return rc;
  // 5. through the samples (and the KL) to the posterior parameters
  for (const nfn_variational_layer* l : ls)
    if ((rc = launch_variational(true, l->posterior, l->prior_loc, l->prior_scale, l->eps, l->dw, nullptr, l->kl_grad, l->n,
                                 draws, nullptr, nullptr, l->dposterior, l->dprior_loc, st)) != NFN_OK)
      return rc;
  return NFN_OK;
}

int64_t nfn_jit_dense_tc5_compile_check(const nfn_chain_desc* desc, int hidden, int accurate) {
  int rc = check_desc(desc);
  if (rc != NFN_OK) return rc;
  std::string log;
  const long long n = jit_dense_tc5_compile_check(desc, hidden, accurate ? 1 : 0, log);
  if (n < 0) return set_error(NFN_ERR_UNSUPPORTED, "NVRTC: %s", log.substr(0, 400).c_str());
  return n;
}

int64_t nfn_jit_dense_compile_check(const nfn_chain_desc* desc, int hidden, int accurate) {
  int rc = check_desc(desc);
  if (rc != NFN_OK) return rc;
  std::string log;
  const long long n = jit_dense_compile_check(desc, hidden, accurate ? 1 : 0, log);
  if (n < 0) return set_error(NFN_ERR_UNSUPPORTED, "NVRTC: %s", log.substr(0, 400).c_str());
  return n;
}

int nfn_dense_act_supported(int in_features, int units, int act) {
  return mlp_layer_supported(in_features, units, act);
}

int nfn_dense_act_forward(const float* x, const float* weight, const float* bias, int64_t B, int in_features,
                          int units, int act, float* out, void* stream) {
  return nfn_dense_act_forward_x(x, nullptr, nullptr, weight, bias, B, in_features, units, act, out, stream);
}

int nfn_dense_act_forward_x(const float* x, const float* x_mean, const float* x_std, const float* weight, const float* bias,
                            int64_t B, int in_features, int units, int act, float* out, void* stream) {
  if ((x_mean == nullptr) != (x_std == nullptr)) return set_error(NFN_ERR_NULL, "x_mean and x_std go together");
  if (!mlp_layer_supported(in_features, units, act))
    return set_error(NFN_ERR_UNSUPPORTED, "dense layer %d -> %d (act %d) is outside the fused kernels' range", in_features,
                     units, act);
  if (B < 0) return set_error(NFN_ERR_SHAPE, "B=%lld", (long long)B);
  if (B == 0) return NFN_OK;
  if (!x || !weight || !bias || !out) return set_error(NFN_ERR_NULL, "x, weight, bias and out must be non-NULL");
  if (!aligned(out, 16) || (in_features % 4 == 0 && !aligned(x, 16)))
    return set_error(NFN_ERR_ALIGN, "x and out must be 16-byte aligned");
  return launch_dense_act_forward(x, x_mean, x_std, weight, bias, out, B, in_features, units, act, (cudaStream_t)stream);
}

int nfn_dense_act_backward(const float* x, const float* out, const float* dout, const float* weight, int64_t B,
                           int in_features, int units, int act, float* dx, float* dweight, float* dbias,
                           void* stream) {
  return nfn_dense_act_backward_x(x, nullptr, nullptr, out, dout, weight, B, in_features, units, act, dx, dweight, dbias,
                                  stream);
}

int nfn_dense_act_backward_x(const float* x, const float* x_mean, const float* x_std, const float* out, const float* dout,
                             const float* weight, int64_t B, int in_features, int units, int act, float* dx,
                             float* dweight, float* dbias, void* stream) {
  if ((x_mean == nullptr) != (x_std == nullptr)) return set_error(NFN_ERR_NULL, "x_mean and x_std go together");
  if (x_mean && (dx != nullptr || in_features > 4 || units > 32))
    return set_error(NFN_ERR_UNSUPPORTED, "the fused input normalisation serves the first layer only (no dx, <= 4 inputs, <= 32 units)");
  if (!mlp_layer_supported(in_features, units, act))
    return set_error(NFN_ERR_UNSUPPORTED, "dense layer %d -> %d (act %d) is outside the fused kernels' range", in_features,
                     units, act);
  if (B < 0) return set_error(NFN_ERR_SHAPE, "B=%lld", (long long)B);
  if (B == 0) return NFN_OK;
  if (!x || !out || !dout || !weight || !dweight || !dbias)
    return set_error(NFN_ERR_NULL, "x, out, dout, weight, dweight and dbias must be non-NULL");
  if (!aligned(out, 16) || !aligned(dout, 16) || (in_features % 4 == 0 && !aligned(x, 16)))
    return set_error(NFN_ERR_ALIGN, "x, out and dout must be 16-byte aligned");
  return launch_dense_act_backward(x, x_mean, x_std, out, dout, weight, dx, dweight, dbias, B, in_features, units, act,
                                   (cudaStream_t)stream);
}

int nfn_flow_forward(int flow_type, int n_dims, const float* t, const float* z, int64_t z_rows,
                     float* z_out, float* fldj, int64_t B, void* stream) {
  if (flow_type < 0 || flow_type > NFN_FLOW_AFFINE)
    return set_error(NFN_ERR_DESC, "flow_type=%d is not planar(0)/radial(1)/affine(2)", flow_type);
  if (n_dims < 1 || n_dims > NFN_MAX_DIMS)
    return set_error(NFN_ERR_DESC, "n_dims=%d outside 1..%d", n_dims, NFN_MAX_DIMS);
  int rc = check_rows(B, z_rows);
  if (rc != NFN_OK) return rc;
  if (B == 0) return NFN_OK;
  if (!t || !z) return set_error(NFN_ERR_NULL, "t and z must be non-NULL");
  if (!aligned(z, event_align(n_dims)))
    return set_error(NFN_ERR_ALIGN, "z must be %zu-byte aligned", event_align(n_dims));
  return cuda_error(launch_flow_single(flow_type, n_dims, t, z, (z_rows == 1 && B != 1), z_out, fldj, B,
                                       (cudaStream_t)stream),
                    "flow_single_kernel");
}

// ------------------------------------------------------------------ mixture heads
static int mix_common(int K, int d, const float* t, const float* y, int64_t y_rows, float* logp,
                      int64_t B) {
  if (K < 1 || K > 4096) return set_error(NFN_ERR_DESC, "n_centers=%d outside 1..4096", K);
  if (d < 1 || d > NFN_MAX_DIMS) return set_error(NFN_ERR_DESC, "n_dims=%d outside 1..%d", d, NFN_MAX_DIMS);
  int rc = check_rows(B, y_rows);
  if (rc != NFN_OK) return rc;
  if (B == 0) return 1;  // nothing to do
  if (!t || !y || !logp) return set_error(NFN_ERR_NULL, "t, y and logp must be non-NULL");
  if (!aligned(t, 16)) return set_error(NFN_ERR_ALIGN, "t must be 16-byte aligned");
  if (!aligned(y, event_align(d))) return set_error(NFN_ERR_ALIGN, "y must be %zu-byte aligned", event_align(d));
  return NFN_OK;
}

int nfn_mdn_forward(int n_centers, int n_dims, const float* t, const float* y, int64_t y_rows,
                    float* logp, int64_t B, void* stream) {
  return nfn_mdn_forward_x(n_centers, n_dims, t, y, y_rows, logp, B, nullptr, stream);
}

int nfn_mdn_forward_x(int n_centers, int n_dims, const float* t, const float* y, int64_t y_rows,
                      float* logp, int64_t B, const nfn_event_xform* xf, void* stream) {
  int rc = mix_common(n_centers, n_dims, t, y, y_rows, logp, B);
  if (rc != NFN_OK) return rc > 0 ? NFN_OK : rc;
  MixArgs a{};
  a.t = t; a.y = y; a.logp = logp; a.B = B; a.g_scale = 1.0f; a.K = n_centers;
  a.y_broadcast = (y_rows == 1 && B != 1);
  if ((rc = set_xform(a.xf, xf, n_dims)) != NFN_OK) return rc;
  return launch_mdn(n_dims, false, a, (cudaStream_t)stream);
}

int nfn_mdn_forward_backward(int n_centers, int n_dims, const float* t, const float* y,
                             int64_t y_rows, const float* g_logp, float g_scale, float* logp,
                             float* dt, float* dy, double* logp_sum, double* dt_colsum, int64_t B,
                             void* stream) {
  return nfn_mdn_forward_backward_x(n_centers, n_dims, t, y, y_rows, g_logp, g_scale, logp, dt, dy, logp_sum, dt_colsum, B,
                                    nullptr, stream);
}

int nfn_mdn_forward_backward_x(int n_centers, int n_dims, const float* t, const float* y,
                               int64_t y_rows, const float* g_logp, float g_scale, float* logp,
                               float* dt, float* dy, double* logp_sum, double* dt_colsum, int64_t B,
                               const nfn_event_xform* xf, void* stream) {
  int rc = mix_common(n_centers, n_dims, t, y, y_rows, logp, B);
  if (rc != NFN_OK) return rc > 0 ? NFN_OK : rc;
  if (!dt) return set_error(NFN_ERR_NULL, "dt must be non-NULL");
  if (!aligned(dt, 16)) return set_error(NFN_ERR_ALIGN, "dt must be 16-byte aligned");
  if (dy && y_rows != B) return set_error(NFN_ERR_SHAPE, "dy requires y_rows == B");
  MixArgs a{};
  a.t = t; a.y = y; a.g_logp = g_logp; a.logp = logp; a.dt = dt; a.dy = dy; a.logp_sum = logp_sum;
  a.dt_colsum = dt_colsum; a.B = B; a.g_scale = g_scale; a.K = n_centers;
  a.y_broadcast = (y_rows == 1 && B != 1);
  if ((rc = set_xform(a.xf, xf, n_dims)) != NFN_OK) return rc;
  return launch_mdn(n_dims, true, a, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ fused Dense(P) + MDN head
static int dense_mdn_common(int K, int d, int hidden, const float* h, const float* W, const float* bias, const float* y,
                            int64_t y_rows, float* logp, int64_t B) {
  if (K < 1 || K > 4096) return set_error(NFN_ERR_DESC, "n_centers=%d outside 1..4096", K);
  if (d < 1 || d > NFN_MAX_DIMS) return set_error(NFN_ERR_DESC, "n_dims=%d outside 1..%d", d, NFN_MAX_DIMS);
  if (hidden < 1) return set_error(NFN_ERR_SHAPE, "hidden=%d", hidden);
  int rc = check_rows(B, y_rows);
  if (rc != NFN_OK) return rc;
  if (B == 0) return 1;
  if (!h || !W || !bias || !y || !logp) return set_error(NFN_ERR_NULL, "h, W, bias, y and logp must be non-NULL");
  if (!aligned(h, 16)) return set_error(NFN_ERR_ALIGN, "h must be 16-byte aligned");
  if (!aligned(y, event_align(d))) return set_error(NFN_ERR_ALIGN, "y must be %zu-byte aligned", event_align(d));
  return NFN_OK;
}

static int dense_mdn_dispatch(int K, int d, int hidden, const DenseArgs& a, bool bwd, cudaStream_t st) {
  const int mode = math_mode();
  const std::string key = dense_mdn_key(K, d, hidden);
  const DenseKernels* k = find_dense(key);
  if (k && k->fn[mode][bwd ? 1 : 0]) return cuda_error(k->fn[mode][bwd ? 1 : 0](a, st), key.c_str());
  bool served = false;
  cudaError_t e = launch_dense_mdn_jit(K, d, hidden, a, bwd, mode, st, &served);
  if (e != cudaSuccess) return cuda_error(e, key.c_str());
  if (!served)
    return set_error(NFN_ERR_UNSUPPORTED,
                     "no fused dense kernel for a %d-component %d-D mixture with hidden width %d (needs a multiple of "
                     "16 <= 64 and an ahead-of-time instance or NVRTC): compose the layer and nfn_mdn_forward_backward instead",
                     K, d, hidden);
  return NFN_OK;
}

int nfn_dense_mdn_forward_x(int n_centers, int n_dims, int hidden, const float* h, const float* W, const float* bias,
                            const float* y, int64_t y_rows, float* logp, int64_t B, const nfn_event_xform* xf,
                            void* stream) {
  int rc = dense_mdn_common(n_centers, n_dims, hidden, h, W, bias, y, y_rows, logp, B);
  if (rc != NFN_OK) return rc > 0 ? NFN_OK : rc;
  DenseArgs a{};
  a.h = h; a.W = W; a.bias = bias; a.y = y; a.logp = logp; a.B = B; a.g_scale = 1.0f;
  a.y_broadcast = (y_rows == 1 && B != 1);
  if ((rc = set_xform(a.xf, xf, n_dims)) != NFN_OK) return rc;
  return dense_mdn_dispatch(n_centers, n_dims, hidden, a, false, (cudaStream_t)stream);
}

int nfn_dense_mdn_forward_backward_x(int n_centers, int n_dims, int hidden, const float* h, const float* W,
                                     const float* bias, const float* y, int64_t y_rows, const float* g_logp,
                                     float g_scale, float* logp, float* dh, float* dW, float* dbias, double* logp_sum,
                                     int64_t B, const nfn_event_xform* xf, void* stream) {
  int rc = dense_mdn_common(n_centers, n_dims, hidden, h, W, bias, y, y_rows, logp, B);
  if (rc != NFN_OK) return rc > 0 ? NFN_OK : rc;
  if (!dh || !dW || !dbias) return set_error(NFN_ERR_NULL, "dh, dW and dbias must be non-NULL");
  if (!aligned(dh, 16)) return set_error(NFN_ERR_ALIGN, "dh must be 16-byte aligned");
  DenseArgs a{};
  a.h = h; a.W = W; a.bias = bias; a.y = y; a.g_logp = g_logp; a.logp = logp; a.dh = dh; a.dW = dW;
  a.dbias = dbias; a.logp_sum = logp_sum; a.B = B; a.g_scale = g_scale;
  a.y_broadcast = (y_rows == 1 && B != 1);
  if ((rc = set_xform(a.xf, xf, n_dims)) != NFN_OK) return rc;
  return dense_mdn_dispatch(n_centers, n_dims, hidden, a, true, (cudaStream_t)stream);
}

int nfn_dense_mdn_forward_draws_x(int n_centers, int n_dims, int hidden, int draws, int64_t rows_per_draw, const float* h,
                                  const float* W, const float* bias, const float* y, int64_t y_rows, float* logp,
                                  const nfn_event_xform* xf, void* stream) {
  int rc = draws_common(draws, rows_per_draw, y_rows);
  if (rc != NFN_OK) return rc;
  const int64_t B = (int64_t)draws * rows_per_draw;
  rc = dense_mdn_common(n_centers, n_dims, hidden, h, W, bias, y, 1, logp, B);
  if (rc != NFN_OK) return rc > 0 ? NFN_OK : rc;
  DenseArgs a{};
  a.h = h; a.W = W; a.bias = bias; a.y = y; a.logp = logp; a.B = B; a.g_scale = 1.0f;
  a.y_broadcast = (y_rows == 1 && rows_per_draw != 1);
  a.draws = draws; a.rows_per_draw = rows_per_draw;
  if ((rc = set_xform(a.xf, xf, n_dims)) != NFN_OK) return rc;
  return dense_mdn_dispatch(n_centers, n_dims, hidden, a, false, (cudaStream_t)stream);
}

int nfn_dense_mdn_forward_backward_draws_x(int n_centers, int n_dims, int hidden, int draws, int64_t rows_per_draw,
                                           const float* h, const float* W, const float* bias, const float* y,
                                           int64_t y_rows, const float* g_logp, float g_scale, float* logp, float* dh,
                                           float* dW, float* dbias, double* logp_sum, const nfn_event_xform* xf,
                                           void* stream) {
  int rc = draws_common(draws, rows_per_draw, y_rows);
  if (rc != NFN_OK) return rc;
  const int64_t B = (int64_t)draws * rows_per_draw;
  rc = dense_mdn_common(n_centers, n_dims, hidden, h, W, bias, y, 1, logp, B);
  if (rc != NFN_OK) return rc > 0 ? NFN_OK : rc;
  if (!dh || !dW || !dbias) return set_error(NFN_ERR_NULL, "dh, dW and dbias must be non-NULL");
  if (!aligned(dh, 16)) return set_error(NFN_ERR_ALIGN, "dh must be 16-byte aligned");
  DenseArgs a{};
  a.h = h; a.W = W; a.bias = bias; a.y = y; a.g_logp = g_logp; a.logp = logp; a.dh = dh; a.dW = dW;
  a.dbias = dbias; a.logp_sum = logp_sum; a.B = B; a.g_scale = g_scale;
  a.y_broadcast = (y_rows == 1 && rows_per_draw != 1);
  a.draws = draws; a.rows_per_draw = rows_per_draw;
  if ((rc = set_xform(a.xf, xf, n_dims)) != NFN_OK) return rc;
  return dense_mdn_dispatch(n_centers, n_dims, hidden, a, true, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ fused Dense(P) + KMN head
static int dense_kmn_dispatch(int MC, int d, int hidden, const DenseArgs& a, bool bwd, cudaStream_t st) {
  const int mode = math_mode();
  const std::string key = dense_kmn_key(MC, d, hidden);
  const DenseKernels* k = find_dense(key);
  if (k && k->fn[mode][bwd ? 1 : 0]) return cuda_error(k->fn[mode][bwd ? 1 : 0](a, st), key.c_str());
  bool served = false;
  cudaError_t e = launch_dense_kmn_jit(MC, d, hidden, a, bwd, mode, st, &served);
  if (e != cudaSuccess) return cuda_error(e, key.c_str());
  if (!served)
    return set_error(NFN_ERR_UNSUPPORTED,
                     "no fused dense kernel for %d %d-D Gaussian kernels with hidden width %d (needs a multiple of 16 <= 64 "
                     "and an ahead-of-time instance or NVRTC): compose the layer and nfn_kmn_forward_backward instead",
                     MC, d, hidden);
  return NFN_OK;
}

int nfn_dense_kmn_forward_x(int n_components, int n_dims, int hidden, const float* h, const float* W, const float* bias,
                            const float* y, int64_t y_rows, const float* locs, const float* scales, float* logp, int64_t B,
                            const nfn_event_xform* xf, void* stream) {
  int rc = dense_mdn_common(n_components, n_dims, hidden, h, W, bias, y, y_rows, logp, B);
  if (rc != NFN_OK) return rc > 0 ? NFN_OK : rc;
  if (!locs || !scales) return set_error(NFN_ERR_NULL, "locs and scales must be non-NULL");
  DenseArgs a{};
  a.h = h; a.W = W; a.bias = bias; a.y = y; a.logp = logp; a.B = B; a.g_scale = 1.0f;
  a.y_broadcast = (y_rows == 1 && B != 1);
  a.locs = locs; a.scales = scales;
  if ((rc = set_xform(a.xf, xf, n_dims)) != NFN_OK) return rc;
  return dense_kmn_dispatch(n_components, n_dims, hidden, a, false, (cudaStream_t)stream);
}

int nfn_dense_kmn_forward_backward_x(int n_components, int n_dims, int hidden, const float* h, const float* W,
                                     const float* bias, const float* y, int64_t y_rows, const float* locs,
                                     const float* scales, const float* g_logp, float g_scale, float* logp, float* dh,
                                     float* dW, float* dbias, float* dscales, double* logp_sum, int64_t B,
                                     const nfn_event_xform* xf, void* stream) {
  int rc = dense_mdn_common(n_components, n_dims, hidden, h, W, bias, y, y_rows, logp, B);
  if (rc != NFN_OK) return rc > 0 ? NFN_OK : rc;
  if (!locs || !scales) return set_error(NFN_ERR_NULL, "locs and scales must be non-NULL");
  if (!dh || !dW || !dbias) return set_error(NFN_ERR_NULL, "dh, dW and dbias must be non-NULL");
  if (!aligned(dh, 16)) return set_error(NFN_ERR_ALIGN, "dh must be 16-byte aligned");
  DenseArgs a{};
  a.h = h; a.W = W; a.bias = bias; a.y = y; a.g_logp = g_logp; a.logp = logp; a.dh = dh; a.dW = dW;
  a.dbias = dbias; a.logp_sum = logp_sum; a.B = B; a.g_scale = g_scale;
  a.y_broadcast = (y_rows == 1 && B != 1);
  a.locs = locs; a.scales = scales; a.dscales = dscales;
  if ((rc = set_xform(a.xf, xf, n_dims)) != NFN_OK) return rc;
  return dense_kmn_dispatch(n_components, n_dims, hidden, a, true, (cudaStream_t)stream);
}

int64_t nfn_jit_dense_kmn_compile_check(int n_components, int n_dims, int hidden, int accurate) {
  std::string log;
  const long long n = jit_dense_kmn_compile_check(n_components, n_dims, hidden, accurate ? 1 : 0, log);
  if (n < 0) return set_error(NFN_ERR_UNSUPPORTED, "dense+kmn runtime specialisation failed: %s", log.c_str());
  return n;
}

int64_t nfn_jit_dense_mdn_compile_check(int n_centers, int n_dims, int hidden, int accurate) {
  std::string log;
  const long long n = jit_dense_mdn_compile_check(n_centers, n_dims, hidden, accurate ? 1 : 0, log);
  if (n < 0) return set_error(NFN_ERR_UNSUPPORTED, "dense+mdn runtime specialisation failed: %s", log.c_str());
  return n;
}

int nfn_kmn_forward(int n_components, int n_dims, const float* t, const float* y, int64_t y_rows,
                    const float* locs, const float* scales, float* logp, int64_t B, void* stream) {
  return nfn_kmn_forward_x(n_components, n_dims, t, y, y_rows, locs, scales, logp, B, nullptr, stream);
}

int nfn_kmn_forward_x(int n_components, int n_dims, const float* t, const float* y, int64_t y_rows,
                      const float* locs, const float* scales, float* logp, int64_t B, const nfn_event_xform* xf,
                      void* stream) {
  int rc = mix_common(n_components, n_dims, t, y, y_rows, logp, B);
  if (rc != NFN_OK) return rc > 0 ? NFN_OK : rc;
  if (!locs || !scales) return set_error(NFN_ERR_NULL, "locs and scales must be non-NULL");
  MixArgs a{};
  a.t = t; a.y = y; a.locs = locs; a.scales = scales; a.logp = logp; a.B = B; a.g_scale = 1.0f;
  a.K = n_components; a.y_broadcast = (y_rows == 1 && B != 1);
  if ((rc = set_xform(a.xf, xf, n_dims)) != NFN_OK) return rc;
  return launch_kmn(n_dims, false, a, (cudaStream_t)stream);
}

int nfn_kmn_forward_backward(int n_components, int n_dims, const float* t, const float* y,
                             int64_t y_rows, const float* locs, const float* scales,
                             const float* g_logp, float g_scale, float* logp, float* dt, float* dy,
                             float* dscales, double* logp_sum, int64_t B, void* stream) {
  return nfn_kmn_forward_backward_x(n_components, n_dims, t, y, y_rows, locs, scales, g_logp, g_scale, logp, dt, dy, dscales,
                                    logp_sum, B, nullptr, stream);
}

int nfn_kmn_forward_backward_x(int n_components, int n_dims, const float* t, const float* y,
                               int64_t y_rows, const float* locs, const float* scales,
                               const float* g_logp, float g_scale, float* logp, float* dt, float* dy,
                               float* dscales, double* logp_sum, int64_t B, const nfn_event_xform* xf, void* stream) {
  int rc = mix_common(n_components, n_dims, t, y, y_rows, logp, B);
  if (rc != NFN_OK) return rc > 0 ? NFN_OK : rc;
  if (!locs || !scales || !dt) return set_error(NFN_ERR_NULL, "locs, scales and dt must be non-NULL");
  if (!aligned(dt, 16)) return set_error(NFN_ERR_ALIGN, "dt must be 16-byte aligned");
  if (dy && y_rows != B) return set_error(NFN_ERR_SHAPE, "dy requires y_rows == B");
  MixArgs a{};
  a.t = t; a.y = y; a.g_logp = g_logp; a.locs = locs; a.scales = scales; a.logp = logp; a.dt = dt;
  a.dy = dy; a.dscales = dscales; a.logp_sum = logp_sum; a.B = B; a.g_scale = g_scale;
  a.K = n_components; a.y_broadcast = (y_rows == 1 && B != 1);
  if ((rc = set_xform(a.xf, xf, n_dims)) != NFN_OK) return rc;
  return launch_kmn(n_dims, true, a, (cudaStream_t)stream);
}

int nfn_logmeanexp_draws(const float* logp_sb, int64_t S, int64_t B, float* out, void* stream) {
  if (S < 1 || B < 0) return set_error(NFN_ERR_SHAPE, "S=%lld, B=%lld", (long long)S, (long long)B);
  if (B == 0) return NFN_OK;
  if (!logp_sb || !out) return set_error(NFN_ERR_NULL, "logp_sb and out must be non-NULL");
  return launch_logmeanexp(logp_sb, S, B, out, (cudaStream_t)stream);
}

}  // extern "C"

// =================================================================== host-buffer pipeline
namespace nfn {

// Per-thread workspace: NS slots, each with its own stream and device staging buffers.
// Chunk i runs H2D -> kernel -> D2H on stream i % NS; chunks on different streams overlap
// (both copy engines + SMs busy), a slot is reused only by later chunks on the same stream.
struct HostPipe {
  static constexpr int NS = 3;
  int device = -1;
  cudaStream_t stream[NS] = {};
  void* buf[NS] = {};
  size_t cap[NS] = {};
  double* acc[NS] = {};    // device: logp_sum per slot
  double* colacc[NS] = {}; // device: dt_colsum per slot
  size_t colcap = 0;
  bool live = false;

  int ensure(size_t bytes, int P) {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_error(e, "cudaGetDevice");
    if (live && dev != device) release();
    if (!live) {
      for (int i = 0; i < NS; ++i) {
        if ((e = cudaStreamCreateWithFlags(&stream[i], cudaStreamNonBlocking)) != cudaSuccess)
          return cuda_error(e, "cudaStreamCreate");
        if ((e = cudaMalloc((void**)&acc[i], sizeof(double))) != cudaSuccess)
          return cuda_error(e, "cudaMalloc(acc)");
      }
      device = dev;
      live = true;
    }
    for (int i = 0; i < NS; ++i) {
      if (cap[i] < bytes) {
        if (buf[i]) cudaFree(buf[i]);
        buf[i] = nullptr;
        cap[i] = 0;
        if ((e = cudaMalloc(&buf[i], bytes)) != cudaSuccess) return cuda_error(e, "cudaMalloc(staging)");
        cap[i] = bytes;
      }
    }
    if (colcap < (size_t)P) {
      for (int i = 0; i < NS; ++i) {
        if (colacc[i]) cudaFree(colacc[i]);
        colacc[i] = nullptr;
        if ((e = cudaMalloc((void**)&colacc[i], (size_t)P * sizeof(double))) != cudaSuccess)
          return cuda_error(e, "cudaMalloc(colacc)");
      }
      colcap = (size_t)P;
    }
    return NFN_OK;
  }

  void release() {
    if (!live) return;
    for (int i = 0; i < NS; ++i) {
      if (stream[i]) { cudaStreamSynchronize(stream[i]); cudaStreamDestroy(stream[i]); }
      if (buf[i]) cudaFree(buf[i]);
      if (acc[i]) cudaFree(acc[i]);
      if (colacc[i]) cudaFree(colacc[i]);
      stream[i] = nullptr; buf[i] = nullptr; acc[i] = nullptr; colacc[i] = nullptr; cap[i] = 0;
    }
    colcap = 0;
    live = false;
  }
};

static thread_local HostPipe g_pipe;

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Generic chunked driver.  `run` enqueues the device work for one chunk.
struct ChunkPtrs {
  float* t; float* y; float* g; float* logp; float* dt;
};

template <class Run>
static int host_pipeline(int P, int d, bool bwd, const float* t, const float* y, int64_t y_rows,
                         const float* g_logp, float* logp, float* dt, double* logp_sum,
                         double* dt_colsum, int64_t B, Run run) {
  if (B == 0) {
    if (logp_sum) *logp_sum = 0.0;
    if (dt_colsum) memset(dt_colsum, 0, (size_t)P * sizeof(double));
    return NFN_OK;
  }
  // ~16 MiB of parameters per chunk (NFN_B200_HOST_CHUNK_MB overrides, tuning only), at least 3
  // chunks in flight when B allows it
  const size_t chunk_mb = (size_t)option(kOptHostChunkMb);
  int64_t rows = (int64_t)((chunk_mb << 20) / ((size_t)(P > 0 ? P : 1) * sizeof(float)));
  rows = rows / 1024 * 1024;
  if (rows < 1024) rows = 1024;
  if (rows > B) rows = B;
  const size_t t_b = align_up((size_t)rows * P * sizeof(float), 256);
  const size_t y_b = align_up((size_t)rows * d * sizeof(float), 256);
  const size_t r_b = align_up((size_t)rows * sizeof(float), 256);
  const size_t total = t_b + y_b + r_b /*g*/ + r_b /*logp*/ + (bwd ? t_b : 0);
  HostPipe& hp = g_pipe;
  int rc = hp.ensure(total, P > 0 ? P : 1);
  if (rc != NFN_OK) return rc;
  cudaError_t e;
  for (int s = 0; s < HostPipe::NS; ++s) {
    if ((e = cudaMemsetAsync(hp.acc[s], 0, sizeof(double), hp.stream[s])) != cudaSuccess)
      return cuda_error(e, "cudaMemsetAsync");
    if (bwd && dt_colsum && P > 0 &&
        (e = cudaMemsetAsync(hp.colacc[s], 0, (size_t)P * sizeof(double), hp.stream[s])) != cudaSuccess)
      return cuda_error(e, "cudaMemsetAsync");
  }
  int64_t done = 0;
  for (int64_t c = 0; done < B; ++c, done += rows) {
    const int s = (int)(c % HostPipe::NS);
    const int64_t n = (B - done < rows) ? (B - done) : rows;
    cudaStream_t st = hp.stream[s];
    char* base = (char*)hp.buf[s];
    ChunkPtrs p;
    p.t = (float*)base;
    p.y = (float*)(base + t_b);
    p.g = (float*)(base + t_b + y_b);
    p.logp = (float*)(base + t_b + y_b + r_b);
    p.dt = bwd ? (float*)(base + t_b + y_b + 2 * r_b) : nullptr;
    if (P > 0 && (e = cudaMemcpyAsync(p.t, t + done * P, (size_t)n * P * sizeof(float),
                                      cudaMemcpyHostToDevice, st)) != cudaSuccess)
      return cuda_error(e, "H2D t");
    const bool ybc = (y_rows == 1 && B != 1);
    if ((e = cudaMemcpyAsync(p.y, ybc ? y : y + done * d, (size_t)(ybc ? 1 : n) * d * sizeof(float),
                             cudaMemcpyHostToDevice, st)) != cudaSuccess)
      return cuda_error(e, "H2D y");
    if (bwd && g_logp &&
        (e = cudaMemcpyAsync(p.g, g_logp + done, (size_t)n * sizeof(float), cudaMemcpyHostToDevice,
                             st)) != cudaSuccess)
      return cuda_error(e, "H2D g_logp");
    rc = run(p, n, ybc ? 1 : n, (bwd && g_logp) ? p.g : nullptr, hp.acc[s],
             (bwd && dt_colsum) ? hp.colacc[s] : nullptr, st);
    if (rc != NFN_OK) return rc;
    if ((e = cudaMemcpyAsync(logp + done, p.logp, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost,
                             st)) != cudaSuccess)
      return cuda_error(e, "D2H logp");
    if (bwd && P > 0 &&
        (e = cudaMemcpyAsync(dt + done * P, p.dt, (size_t)n * P * sizeof(float), cudaMemcpyDeviceToHost,
                             st)) != cudaSuccess)
      return cuda_error(e, "D2H dt");
  }
  double sum = 0.0;
  std::vector<double> col((size_t)(P > 0 ? P : 1));
  if (dt_colsum) memset(dt_colsum, 0, (size_t)P * sizeof(double));
  for (int s = 0; s < HostPipe::NS; ++s) {
    if ((e = cudaStreamSynchronize(hp.stream[s])) != cudaSuccess) return cuda_error(e, "stream sync");
    double part = 0.0;
    if ((e = cudaMemcpy(&part, hp.acc[s], sizeof(double), cudaMemcpyDeviceToHost)) != cudaSuccess)
      return cuda_error(e, "D2H logp_sum");
    sum += part;
    if (bwd && dt_colsum && P > 0) {
      if ((e = cudaMemcpy(col.data(), hp.colacc[s], (size_t)P * sizeof(double), cudaMemcpyDeviceToHost)) !=
          cudaSuccess)
        return cuda_error(e, "D2H dt_colsum");
      for (int j = 0; j < P; ++j) dt_colsum[j] += col[j];
    }
  }
  if (logp_sum) *logp_sum = sum;
  return NFN_OK;
}

}  // namespace nfn

extern "C" {

int nfn_chain_forward_host(const nfn_chain_desc* desc, const float* t, const float* y, int64_t y_rows,
                           float* logp, int64_t B) {
  int rc = check_desc(desc);
  if (rc != NFN_OK) return rc;
  if ((rc = check_rows(B, y_rows)) != NFN_OK) return rc;
  const int P = param_size(desc);
  if (B > 0 && (!y || !logp || (P > 0 && !t))) return set_error(NFN_ERR_NULL, "t, y and logp must be non-NULL");
  return host_pipeline(P, desc->n_dims, false, t, y, y_rows, nullptr, logp, nullptr, nullptr, nullptr, B,
                       [&](const ChunkPtrs& p, int64_t n, int64_t yr, const float*, double*, double*,
                           cudaStream_t st) {
                         return nfn_chain_forward(desc, p.t, p.y, yr, p.logp, n, st);
                       });
}

int nfn_chain_forward_backward_host(const nfn_chain_desc* desc, const float* t, const float* y,
                                    int64_t y_rows, const float* g_logp, float g_scale, float* logp,
                                    float* dt, double* logp_sum, double* dt_colsum, int64_t B) {
  int rc = check_desc(desc);
  if (rc != NFN_OK) return rc;
  if ((rc = check_rows(B, y_rows)) != NFN_OK) return rc;
  const int P = param_size(desc);
  if (B > 0 && (!y || !logp || (P > 0 && (!t || !dt))))
    return set_error(NFN_ERR_NULL, "t, y, logp and dt must be non-NULL");
  return host_pipeline(P, desc->n_dims, true, t, y, y_rows, g_logp, logp, dt, logp_sum, dt_colsum, B,
                       [&](const ChunkPtrs& p, int64_t n, int64_t yr, const float* g, double* acc,
                           double* col, cudaStream_t st) {
                         return nfn_chain_forward_backward(desc, p.t, p.y, yr, g, g_scale, p.logp, p.dt,
                                                           nullptr, acc, col, n, st);
                       });
}

int nfn_mdn_forward_backward_host(int n_centers, int n_dims, const float* t, const float* y,
                                  int64_t y_rows, const float* g_logp, float g_scale, float* logp,
                                  float* dt, double* logp_sum, int64_t B) {
  if (n_centers < 1 || n_centers > 4096) return set_error(NFN_ERR_DESC, "n_centers=%d outside 1..4096", n_centers);
  if (n_dims < 1 || n_dims > NFN_MAX_DIMS) return set_error(NFN_ERR_DESC, "n_dims=%d outside 1..%d", n_dims, NFN_MAX_DIMS);
  int rc = check_rows(B, y_rows);
  if (rc != NFN_OK) return rc;
  if (B > 0 && (!t || !y || !logp || !dt)) return set_error(NFN_ERR_NULL, "t, y, logp and dt must be non-NULL");
  const int P = 2 * n_centers * n_dims + n_centers;
  return host_pipeline(P, n_dims, true, t, y, y_rows, g_logp, logp, dt, logp_sum, nullptr, B,
                       [&](const ChunkPtrs& p, int64_t n, int64_t yr, const float* g, double* acc, double*,
                           cudaStream_t st) {
                         return nfn_mdn_forward_backward(n_centers, n_dims, p.t, p.y, yr, g, g_scale, p.logp,
                                                         p.dt, nullptr, acc, nullptr, n, st);
                       });
}

int nfn_host_release(void) {
  g_pipe.release();
  return NFN_OK;
}

}  // extern "C"
