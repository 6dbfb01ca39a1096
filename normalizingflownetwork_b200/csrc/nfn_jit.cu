// nfn_jit.cu -- runtime specialiser: any chain the descriptor can express gets the same
// compile-time specialised kernel as the AOT-listed ones.
//
// On a registry miss the chain's ChainSpec is instantiated from the embedded device headers
// (nfn_math.cuh / nfn_flows.cuh / nfn_chain_kernel.cuh, the exact sources the AOT kernels are
// built from) with NVRTC for sm_100a, the cubin is loaded with cudaLibraryLoadData and the
// kernels launched through their cudaKernel_t handles.  Compiled cubins are cached in memory
// (per process) and on disk ($NFN_B200_CACHE or ~/.cache/nfn_b200).  libnvrtc is dlopen'ed
// lazily, so the library loads (and everything AOT works) on machines without it; if NVRTC is
// unavailable or NFN_B200_JIT=0 the generic runtime-chain kernel serves the chain instead.
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "nfn_common.h"
#include "nfn_dense_chain.cuh"
#include "nfn_dense_tc5.cuh"

// embedded by build.py (csrc/_gen/nfn_jit_sources.cu)
extern const char* const nfn_jit_src_math;
extern const char* const nfn_jit_src_flows;
extern const char* const nfn_jit_src_chain;
extern const char* const nfn_jit_src_dense;
extern const char* const nfn_jit_src_dense_tc5;
extern const char* const nfn_jit_src_mixture_row;

namespace nfn {

namespace {

// ------------------------------------------------------------------ NVRTC via dlopen
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
  void* h = nullptr;
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

Nvrtc& nvrtc() {
  static Nvrtc n;
  static bool tried = false;
  if (tried) return n;
  tried = true;
  const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"};
  for (const char* nm : names) {
    n.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (n.h) break;
  }
  if (!n.h) return n;
#define NFN_SYM(field, sym) *(void**)(&n.field) = dlsym(n.h, sym)
  NFN_SYM(CreateProgram, "nvrtcCreateProgram");
  NFN_SYM(CompileProgram, "nvrtcCompileProgram");
  NFN_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
  NFN_SYM(GetCUBIN, "nvrtcGetCUBIN");
  NFN_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
  NFN_SYM(GetProgramLog, "nvrtcGetProgramLog");
  NFN_SYM(DestroyProgram, "nvrtcDestroyProgram");
  NFN_SYM(GetErrorString, "nvrtcGetErrorString");
#undef NFN_SYM
  n.ok = n.CreateProgram && n.CompileProgram && n.GetCUBINSize && n.GetCUBIN && n.GetProgramLogSize &&
         n.GetProgramLog && n.DestroyProgram;
  return n;
}

// ------------------------------------------------------------------ cache
struct JitEntry {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t kern[2] = {nullptr, nullptr};  // [bwd]
  ChainGeometry geo[2];                       // T, (NB), MINB, smem_bytes per kernel
  int ctas_per_sm[2] = {0, 0};
  int device = -1;
  bool failed = false;
};

std::mutex g_mu;
std::map<std::string, JitEntry> g_cache;  // key: "<chain key>|m<mode>|dev<id>"

unsigned long long fnv1a(const std::string& s, unsigned long long h = 1469598103934665603ull) {
  for (unsigned char c : s) {
    h ^= c;
    h *= 1099511628211ull;
  }
  return h;
}

std::string cache_dir() {
  const char* e = getenv("NFN_B200_CACHE");
  std::string d;
  if (e && *e) {
    d = e;
  } else {
    const char* home = getenv("HOME");
    d = std::string(home && *home ? home : "/tmp") + "/.cache/nfn_b200";
  }
  std::string cur;
  for (size_t i = 0; i <= d.size(); ++i) {  // mkdir -p
    if (i == d.size() || d[i] == '/') {
      if (!cur.empty()) mkdir(cur.c_str(), 0755);
    }
    if (i < d.size()) cur.push_back(d[i]);
  }
  return d;
}

bool read_file(const std::string& path, std::vector<char>& out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  out.resize(n > 0 ? (size_t)n : 0);
  const bool ok = n > 0 && fread(out.data(), 1, (size_t)n, f) == (size_t)n;
  fclose(f);
  return ok;
}

void write_file_atomic(const std::string& path, const std::vector<char>& data) {
  const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return;
  const bool ok = fwrite(data.data(), 1, data.size(), f) == data.size();
  fclose(f);
  if (ok) rename(tmp.c_str(), path.c_str()); else remove(tmp.c_str());
}

// ------------------------------------------------------------------ program text
std::string program_source(const nfn_chain_desc* d, int mode, const ChainGeometry (&geo)[2]) {
  std::string spec = "nfn::ChainSpec<" + std::to_string(d->n_dims) + ", " + (d->trainable_base ? "true" : "false");
  for (int k = 0; k < d->n_flows; ++k) spec += ", " + std::to_string((int)d->flow_type[k]);
  spec += ">";
  const char* math = mode == 0 ? "nfn::MathFast" : "nfn::MathAccurate";
  std::string s = "#include \"nfn_chain_kernel.cuh\"\n";
  s += "using Spec = " + spec + ";\n";
  const char* names[2] = {"nfn_jit_chain_fwd", "nfn_jit_chain_fwd_bwd"};
  for (int b = 0; b < 2; ++b) {
    s += "extern \"C\" __global__ void __launch_bounds__(" + std::to_string(geo[b].T) + ", " +
         std::to_string(geo[b].MINB) + ") " + names[b] + "(const nfn::ChainArgs a) {\n  nfn::chain_body<Spec, " +
         (b ? "true" : "false") + ", " + math + ", " + std::to_string(geo[b].T) + ", " + std::to_string(geo[b].NB) +
         ">(a);\n}\n";
  }
  return s;
}

// the warp-tile (bulk copy / TMA) generation of the same kernels
std::string program_source_w(const nfn_chain_desc* d, int mode, const ChainGeometry (&geo)[2]) {
  std::string spec = "nfn::ChainSpec<" + std::to_string(d->n_dims) + ", " + (d->trainable_base ? "true" : "false");
  for (int k = 0; k < d->n_flows; ++k) spec += ", " + std::to_string((int)d->flow_type[k]);
  spec += ">";
  const char* math = mode == 0 ? "nfn::MathFast" : "nfn::MathAccurate";
  std::string s = "#include \"nfn_chain_kernel.cuh\"\n";
  s += "using Spec = " + spec + ";\n";
  const char* names[2] = {"nfn_jit_chain_w_fwd", "nfn_jit_chain_w_fwd_bwd"};
  for (int b = 0; b < 2; ++b) {
    s += "extern \"C\" __global__ void __launch_bounds__(" + std::to_string(geo[b].T) + ", " +
         std::to_string(geo[b].MINB) + ") " + names[b] +
         "(const nfn::ChainArgs a, const __grid_constant__ nfn::TensorMap tm_t, const __grid_constant__ nfn::TensorMap "
         "tm_dt) {\n  nfn::chain_body_w<Spec, " + (b ? "true" : "false") + ", " + math + ", " +
         std::to_string(geo[b].T / 32) + ", " + std::to_string(geo[b].NB) + ">(a, &tm_t, &tm_dt);\n}\n";
  }
  return s;
}

bool compile_cubin(const std::string& src, std::vector<char>& cubin, std::string& log) {
  Nvrtc& n = nvrtc();
  if (!n.ok) {
    log = "libnvrtc.so.12 not found";
    return false;
  }
  const char* headers[6] = {nfn_jit_src_math, nfn_jit_src_flows, nfn_jit_src_chain, nfn_jit_src_dense,
                            nfn_jit_src_dense_tc5, nfn_jit_src_mixture_row};
  const char* hnames[6] = {"nfn_math.cuh", "nfn_flows.cuh", "nfn_chain_kernel.cuh", "nfn_dense_chain.cuh",
                           "nfn_dense_tc5.cuh", "nfn_mixture_row.cuh"};
  nvrtcProgram prog = nullptr;
  int rc = n.CreateProgram(&prog, src.c_str(), "nfn_jit_chain.cu", 6, headers, hnames);
  if (rc != 0) {
    log = "nvrtcCreateProgram failed";
    return false;
  }
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-default-device", "-lineinfo"};
  rc = n.CompileProgram(prog, 4, opts);
  size_t ls = 0;
  n.GetProgramLogSize(prog, &ls);
  if (ls > 1) {
    log.resize(ls);
    n.GetProgramLog(prog, &log[0]);
  }
  bool ok = false;
  if (rc == 0) {
    size_t cs = 0;
    if (n.GetCUBINSize(prog, &cs) == 0 && cs > 0) {
      cubin.resize(cs);
      ok = n.GetCUBIN(prog, cubin.data()) == 0;
    }
  }
  n.DestroyProgram(&prog);
  return ok;
}

// chains whose z history would not fit the register file are left to the generic kernel
bool jit_eligible(const nfn_chain_desc* d, int P) {
  if (d->n_flows * d->n_dims > 96) return false;
  if (P > 0 && chain_geometry(P, true).smem_bytes > 200u * 1024u) return false;
  return true;
}

// Looks up / builds the two kernels (forward, forward+backward) of one program.  Returns nullptr
// when the generic kernel should serve the request (NVRTC missing, compile or load failure).
static JitEntry* get_or_build(const std::string& ckey, const std::string& src, const char* const (&names)[2],
                              const ChainGeometry (&geo)[2]) {
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_cache.find(ckey);
  if (it == g_cache.end()) {
    JitEntry e;
    e.device = device_info().device;
    e.geo[0] = geo[0];
    e.geo[1] = geo[1];
    const std::string all = src + nfn_jit_src_math + nfn_jit_src_flows + nfn_jit_src_chain + nfn_jit_src_dense +
                            nfn_jit_src_dense_tc5 + nfn_jit_src_mixture_row + "|sm_100a|v5";
    char name[64];
    snprintf(name, sizeof(name), "/chain_%016llx.cubin", fnv1a(all));
    const std::string path = cache_dir() + name;
    std::vector<char> cubin;
    std::string log;
    bool have = read_file(path, cubin);
    if (!have) {
      have = compile_cubin(src, cubin, log);
      if (have) write_file_atomic(path, cubin);
      else if (getenv("NFN_B200_JIT_VERBOSE")) fprintf(stderr, "[nfn_b200 jit] %s: %s\n", ckey.c_str(), log.c_str());
    }
    if (have) {
      cudaError_t ce = cudaLibraryLoadData(&e.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
      for (int b = 0; b < 2 && ce == cudaSuccess; ++b) {
        ce = cudaLibraryGetKernel(&e.kern[b], e.lib, names[b]);
        if (ce == cudaSuccess)
          ce = cudaFuncSetAttribute((const void*)e.kern[b], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)e.geo[b].smem_bytes);
        int occ = 0;
        if (ce == cudaSuccess)
          ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)e.kern[b], e.geo[b].T,
                                                             e.geo[b].smem_bytes);
        if (e.geo[b].force_ctas > 0) occ = e.geo[b].force_ctas;
        if (e.geo[b].max_ctas > 0 && occ > e.geo[b].max_ctas) occ = e.geo[b].max_ctas;
        e.ctas_per_sm[b] = occ > 0 ? occ : 1;
      }
      if (ce != cudaSuccess) {
        cudaGetLastError();
        e.failed = true;
      }
    } else {
      e.failed = true;
    }
    it = g_cache.emplace(ckey, e).first;
  }
  return it->second.failed ? nullptr : &it->second;
}

static bool jit_enabled() { return option(kOptJit) != 0; }

static int desc_param_size(const nfn_chain_desc* desc) {
  int P = desc->trainable_base ? 2 * desc->n_dims : 0;
  for (int k = 0; k < desc->n_flows; ++k) P += flow_param_size(desc->flow_type[k], desc->n_dims);
  return P;
}

static std::string spec_text(const nfn_chain_desc* d) {
  std::string spec = "nfn::ChainSpec<" + std::to_string(d->n_dims) + ", " + (d->trainable_base ? "true" : "false");
  for (int k = 0; k < d->n_flows; ++k) spec += ", " + std::to_string((int)d->flow_type[k]);
  return spec + ">";
}

template <class Args>
static cudaError_t launch_entry(JitEntry* ent, int b, const Args& a, long long B, cudaStream_t st) {
  const int T = ent->geo[b].T;
  const int rows = ent->geo[b].rows > 0 ? ent->geo[b].rows : T;
  const long long ntiles = (B + rows - 1) / rows;
  long long grid = (long long)device_info().sm_count * ent->ctas_per_sm[b];
  if (grid > ntiles) grid = ntiles;
  Args args = a;
  void* params[] = {&args};
  cudaError_t ce = cudaLaunchKernel((const void*)ent->kern[b], dim3((unsigned)grid), dim3((unsigned)T), params,
                                    ent->geo[b].smem_bytes, st);
  if (ce == cudaSuccess) count_launch();
  return ce;
}

}  // namespace

cudaError_t launch_chain_jit(const nfn_chain_desc* desc, const std::string& key, const ChainArgs& a, bool bwd,
                             int mode, cudaStream_t st, bool* served) {
  *served = false;
  if (!jit_enabled()) return cudaSuccess;
  const int P = desc_param_size(desc);
  if (!jit_eligible(desc, P)) return cudaSuccess;
  const int io = chain_io_override();
  if (P > 0 && (io >= 0 ? io == 1 : chain_prefers_warp_tile(P, bwd))) {
    const int hist = desc->n_flows * desc->n_dims, tnb = option(kOptTuneWnb), tw = option(kOptTuneWwarps);
    const ChainGeometry geo[2] = {warp_tile_geometry(P, false, hist, tnb, tw), warp_tile_geometry(P, true, hist, tnb, tw)};
    const char* const names[2] = {"nfn_jit_chain_w_fwd", "nfn_jit_chain_w_fwd_bwd"};
    const std::string ckey = key + "|w" + std::to_string(tnb) + "." + std::to_string(tw) + "|m" + std::to_string(mode) +
                             "|dev" + std::to_string(device_info().device);
    JitEntry* ent = get_or_build(ckey, program_source_w(desc, mode, geo), names, geo);
    if (ent) {
      TensorMap tm_t{}, tm_dt{};
      const int W = warp_tile_box(P);
      if (W > 0) {
        if (encode_row_tensor_map(&tm_t, a.t, a.B, P, W) != NFN_OK) return cudaErrorInvalidValue;
        if (bwd && encode_row_tensor_map(&tm_dt, a.dt, a.B, P, W) != NFN_OK) return cudaErrorInvalidValue;
      }
      const int b = bwd ? 1 : 0, T = ent->geo[b].T;
      const long long ntiles = (a.B + T - 1) / T + ((a.peer.world > 0 && a.peer.deferred) ? 1 : 0);  // + the exchange CTA
      long long grid = (long long)device_info().sm_count * ent->ctas_per_sm[b];
      if (grid > ntiles) grid = ntiles;
      ChainArgs args = a;
      void* params[] = {&args, &tm_t, &tm_dt};
      cudaLaunchConfig_t lc = {};
      lc.gridDim = dim3((unsigned)grid);
      lc.blockDim = dim3((unsigned)T);
      lc.dynamicSmemBytes = ent->geo[b].smem_bytes;
      lc.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      lc.attrs = at;
      lc.numAttrs = pdl_enabled() ? 1 : 0;
      cudaError_t ce = cudaLaunchKernelExC(&lc, (const void*)ent->kern[b], params);
      if (ce == cudaSuccess) {
        count_launch();
        *served = true;
      }
      return ce;
    }
    // (the warp-tile build failed) a split-phase exchange needs its CTA: leave the launch to the generic path
    if (a.peer.world > 0 && a.peer.deferred) return cudaSuccess;
  }
  const ChainGeometry geo[2] = {chain_geometry(P, false), chain_geometry(P, true)};
  const char* const names[2] = {"nfn_jit_chain_fwd", "nfn_jit_chain_fwd_bwd"};
  const std::string ckey = key + "|m" + std::to_string(mode) + "|dev" + std::to_string(device_info().device);
  JitEntry* ent = get_or_build(ckey, program_source(desc, mode, geo), names, geo);
  if (!ent) return cudaSuccess;  // generic kernel takes over
  cudaError_t ce = launch_entry(ent, bwd ? 1 : 0, a, a.B, st);
  if (ce == cudaSuccess) *served = true;
  return ce;
}

// ------------------------------------------------------------------ fused Dense(P) + chain
static std::string dense_program_source(const nfn_chain_desc* d, int H, int mode, const ChainGeometry (&geo)[2]) {
  const char* math = mode == 0 ? "nfn::MathFast" : "nfn::MathAccurate";
  std::string s = "#include \"nfn_dense_chain.cuh\"\nusing Spec = " + spec_text(d) + ";\n";
  const char* names[2] = {"nfn_jit_dense_fwd", "nfn_jit_dense_fwd_bwd"};
  for (int b = 0; b < 2; ++b) {
    s += "extern \"C\" __global__ void __launch_bounds__(" + std::to_string(geo[b].T) + ", " +
         std::to_string(geo[b].MINB) + ") " + names[b] + "(const nfn::DenseArgs a) {\n  nfn::dense_chain_body<Spec, " +
         std::to_string(H) + ", " + (b ? "true" : "false") + ", " + math + ", " + std::to_string(geo[b].T) +
         ">(a);\n}\n";
  }
  return s;
}

static void dense_geometry(int P, int H, ChainGeometry (&geo)[2]) {
  for (int b = 0; b < 2; ++b) {
    geo[b].T = 128;
    geo[b].NB = 2;
    geo[b].smem_bytes = dense_smem_bytes(P, H, 128, b == 1);
    const int by_smem = (int)((227u * 1024u) / (geo[b].smem_bytes + 1024u));
    const int want = b ? 2 : 3;
    geo[b].MINB = by_smem < 1 ? 1 : (by_smem < want ? by_smem : want);
  }
}

// ---- the tcgen05 / TMEM version (nfn_dense_tc5.cuh): 256 threads, 128 rows per tile
static std::string dense_tc5_program_source(const nfn_chain_desc* d, int H, int mode, const ChainGeometry (&geo)[2]) {
  const char* math = mode == 0 ? "nfn::MathFast" : "nfn::MathAccurate";
  std::string s = "#include \"nfn_dense_tc5.cuh\"\nusing Spec = " + spec_text(d) + ";\n";
  const char* names[2] = {"nfn_jit_dense_tc5_fwd", "nfn_jit_dense_tc5_fwd_bwd"};
  for (int b = 0; b < 2; ++b) {
    s += "extern \"C\" __global__ void __launch_bounds__(" + std::to_string(geo[b].T) + ", " +
         std::to_string(geo[b].MINB) + ") " + names[b] + "(const nfn::DenseArgs a) {\n  nfn::tc5::dense_tc5_body<Spec, " +
         std::to_string(H) + ", " + (b ? "true" : "false") + ", " + math + ", " + std::to_string(geo[b].MINB) +
         ">(a);\n}\n";
  }
  return s;
}

static void dense_tc5_geometry(int P, int H, ChainGeometry (&geo)[2]) {
  for (int b = 0; b < 2; ++b) {
    geo[b].T = tc5::kThreads;
    geo[b].rows = tc5::kRows;
    geo[b].NB = 1;
    geo[b].smem_bytes = tc5::smem_bytes(P, H, b == 1);
    geo[b].MINB = tc5::min_blocks(P, H, b == 1);
    geo[b].max_ctas = (int)(512u / tc5::tmem_cols(P, H, b == 1));
    geo[b].force_ctas = tc5::resident_ctas(P, H, b == 1);   // the occupancy API says 1 for TMEM kernels
  }
}

static bool dense_tc5_eligible(const nfn_chain_desc* desc, int P, int H) {
  // P <= 128: one accumulator row per thread; the operand over-read of GEMM 3 must stay inside the CTA's smem
  return P >= 1 && P <= 128 && H % 16 == 0 && H >= 16 && H <= 64 && jit_eligible(desc, P) &&
         6 * tc5::g_kW(P, H) >= (unsigned)(16 * tc5::g_NP3(P) - 3 * (tc5::round16(P) / 8)) * 128u &&
         tc5::smem_bytes(P, H, true) <= 220u * 1024u;
}

// served == false (with cudaSuccess): the caller falls back to the mma.sync version
cudaError_t launch_dense_tc5_jit(const nfn_chain_desc* desc, int H, const std::string& key, const DenseArgs& a,
                                 bool bwd, int mode, cudaStream_t st, bool* served) {
  *served = false;
  if (!jit_enabled()) return cudaSuccess;
  const int P = desc_param_size(desc);
  if (!dense_tc5_eligible(desc, P, H)) return cudaSuccess;
  ChainGeometry geo[2];
  dense_tc5_geometry(P, H, geo);
  const char* const names[2] = {"nfn_jit_dense_tc5_fwd", "nfn_jit_dense_tc5_fwd_bwd"};
  const std::string ckey =
      key + "|tc5dense" + std::to_string(H) + "|m" + std::to_string(mode) + "|dev" + std::to_string(device_info().device);
  JitEntry* ent = get_or_build(ckey, dense_tc5_program_source(desc, H, mode, geo), names, geo);
  if (!ent) return cudaSuccess;
  cudaError_t ce = launch_entry(ent, bwd ? 1 : 0, a, a.B, st);
  if (ce == cudaSuccess) *served = true;
  return ce;
}

long long jit_dense_tc5_compile_check(const nfn_chain_desc* desc, int H, int mode, std::string& log) {
  const int P = desc_param_size(desc);
  if (!dense_tc5_eligible(desc, P, H)) {
    log = "chain / hidden width not eligible for the tcgen05 kernel";
    return -1;
  }
  ChainGeometry geo[2];
  dense_tc5_geometry(P, H, geo);
  std::vector<char> cubin;
  if (!compile_cubin(dense_tc5_program_source(desc, H, mode, geo), cubin, log)) return -1;
  return (long long)cubin.size();
}

// served == false (with cudaSuccess): the caller must use the unfused path
cudaError_t launch_dense_jit(const nfn_chain_desc* desc, int H, const std::string& key, const DenseArgs& a,
                             bool bwd, int mode, cudaStream_t st, bool* served) {
  *served = false;
  if (!jit_enabled()) return cudaSuccess;
  const int P = desc_param_size(desc);
  if (P < 1 || H % 16 != 0 || H < 16 || H > 64 || !jit_eligible(desc, P)) return cudaSuccess;
  ChainGeometry geo[2];
  dense_geometry(P, H, geo);
  if (geo[1].smem_bytes > 220u * 1024u) return cudaSuccess;
  const char* const names[2] = {"nfn_jit_dense_fwd", "nfn_jit_dense_fwd_bwd"};
  const std::string ckey =
      key + "|dense" + std::to_string(H) + "|m" + std::to_string(mode) + "|dev" + std::to_string(device_info().device);
  JitEntry* ent = get_or_build(ckey, dense_program_source(desc, H, mode, geo), names, geo);
  if (!ent) return cudaSuccess;
  cudaError_t ce = launch_entry(ent, bwd ? 1 : 0, a, a.B, st);
  if (ce == cudaSuccess) *served = true;
  return ce;
}

long long jit_dense_compile_check(const nfn_chain_desc* desc, int H, int mode, std::string& log) {
  ChainGeometry geo[2];
  dense_geometry(desc_param_size(desc), H, geo);
  std::vector<char> cubin;
  if (!compile_cubin(dense_program_source(desc, H, mode, geo), cubin, log)) return -1;
  return (long long)cubin.size();
}

// compile only (no device needed): returns the cubin size, or -1 with the NVRTC log in `log`
long long jit_compile_check(const nfn_chain_desc* desc, int mode, std::string& log) {
  int P = desc->trainable_base ? 2 * desc->n_dims : 0;
  for (int k = 0; k < desc->n_flows; ++k) P += flow_param_size(desc->flow_type[k], desc->n_dims);
  ChainGeometry geo[2] = {chain_geometry(P, false), chain_geometry(P, true)};
  std::vector<char> cubin;
  if (!compile_cubin(program_source(desc, mode, geo), cubin, log)) return -1;
  long long total = (long long)cubin.size();
  if (P > 0) {  // the warp-tile generation must build from the same embedded headers
    const int hist = desc->n_flows * desc->n_dims;
    ChainGeometry geow[2] = {warp_tile_geometry(P, false, hist), warp_tile_geometry(P, true, hist)};
    std::vector<char> cubin_w;
    if (!compile_cubin(program_source_w(desc, mode, geow), cubin_w, log)) return -1;
    total += (long long)cubin_w.size();
  }
  return total;
}

// number of chains compiled (or loaded from the disk cache) by this process
int jit_cache_size() {
  std::lock_guard<std::mutex> lock(g_mu);
  int n = 0;
  for (auto& kv : g_cache) n += kv.second.failed ? 0 : 1;
  return n;
}

// ------------------------------------------------------------------ fused Dense(P) + MDN head
static std::string dense_mdn_program_source(int K, int D, int H, int mode, const ChainGeometry (&geo)[2]) {
  const char* math = mode == 0 ? "nfn::MathFast" : "nfn::MathAccurate";
  std::string s = "#include \"nfn_dense_chain.cuh\"\nusing Head = nfn::MdnHead<" + std::to_string(K) + ", " +
                  std::to_string(D) + ">;\n";
  const char* names[2] = {"nfn_jit_dense_mdn_fwd", "nfn_jit_dense_mdn_fwd_bwd"};
  for (int b = 0; b < 2; ++b) {
    s += "extern \"C\" __global__ void __launch_bounds__(" + std::to_string(geo[b].T) + ", " +
         std::to_string(geo[b].MINB) + ") " + names[b] + "(const nfn::DenseArgs a) {\n  nfn::dense_head_body<Head, " +
         std::to_string(H) + ", " + (b ? "true" : "false") + ", " + math + ", " + std::to_string(geo[b].T) +
         ">(a);\n}\n";
  }
  return s;
}

static bool dense_mdn_eligible(int K, int D, int H, const ChainGeometry (&geo)[2]) {
  return K >= 1 && D >= 1 && D <= 8 && H % 16 == 0 && H >= 16 && H <= 64 && geo[1].smem_bytes <= 220u * 1024u;
}

// served == false (with cudaSuccess): the caller must use the unfused path
cudaError_t launch_dense_mdn_jit(int K, int D, int H, const DenseArgs& a, bool bwd, int mode, cudaStream_t st,
                                 bool* served) {
  *served = false;
  if (!jit_enabled()) return cudaSuccess;
  ChainGeometry geo[2];
  dense_geometry(K * (2 * D + 1), H, geo);
  if (!dense_mdn_eligible(K, D, H, geo)) return cudaSuccess;
  const char* const names[2] = {"nfn_jit_dense_mdn_fwd", "nfn_jit_dense_mdn_fwd_bwd"};
  const std::string ckey = dense_mdn_key(K, D, H) + "|m" + std::to_string(mode) + "|dev" + std::to_string(device_info().device);
  JitEntry* ent = get_or_build(ckey, dense_mdn_program_source(K, D, H, mode, geo), names, geo);
  if (!ent) return cudaSuccess;
  cudaError_t ce = launch_entry(ent, bwd ? 1 : 0, a, a.B, st);
  if (ce == cudaSuccess) *served = true;
  return ce;
}

long long jit_dense_mdn_compile_check(int K, int D, int H, int mode, std::string& log) {
  ChainGeometry geo[2];
  dense_geometry(K * (2 * D + 1), H, geo);
  if (!dense_mdn_eligible(K, D, H, geo)) {
    log = "mixture / hidden width not eligible for the fused kernel";
    return -1;
  }
  std::vector<char> cubin;
  if (!compile_cubin(dense_mdn_program_source(K, D, H, mode, geo), cubin, log)) return -1;
  return (long long)cubin.size();
}

// ------------------------------------------------------------------ fused Dense(P) + KMN head
static std::string dense_kmn_program_source(int MC, int D, int H, int mode, const ChainGeometry (&geo)[2]) {
  const char* math = mode == 0 ? "nfn::MathFast" : "nfn::MathAccurate";
  std::string s = "#include \"nfn_dense_chain.cuh\"\nusing Head = nfn::KmnHead<" + std::to_string(MC) + ", " +
                  std::to_string(D) + ">;\n";
  const char* names[2] = {"nfn_jit_dense_kmn_fwd", "nfn_jit_dense_kmn_fwd_bwd"};
  for (int b = 0; b < 2; ++b) {
    s += "extern \"C\" __global__ void __launch_bounds__(" + std::to_string(geo[b].T) + ", " +
         std::to_string(geo[b].MINB) + ") " + names[b] + "(const nfn::DenseArgs a) {\n  nfn::dense_head_body<Head, " +
         std::to_string(H) + ", " + (b ? "true" : "false") + ", " + math + ", " + std::to_string(geo[b].T) +
         ">(a);\n}\n";
  }
  return s;
}

static void dense_kmn_geometry(int MC, int D, int H, ChainGeometry (&geo)[2]) {
  for (int b = 0; b < 2; ++b) {
    geo[b].T = 128;
    geo[b].NB = 2;
    geo[b].smem_bytes = dense_smem_bytes(MC, H, 128, b == 1, kmn_extra_floats(MC, D, 128, b == 1));
    const int by_smem = (int)((227u * 1024u) / (geo[b].smem_bytes + 1024u));
    const int want = b ? 2 : 3;
    geo[b].MINB = by_smem < 1 ? 1 : (by_smem < want ? by_smem : want);
  }
}

static bool dense_kmn_eligible(int MC, int D, int H, const ChainGeometry (&geo)[2]) {
  return MC >= 1 && D >= 1 && D <= 8 && H % 16 == 0 && H >= 16 && H <= 64 && geo[1].smem_bytes <= 220u * 1024u;
}

// served == false (with cudaSuccess): the caller must use the unfused path
cudaError_t launch_dense_kmn_jit(int MC, int D, int H, const DenseArgs& a, bool bwd, int mode, cudaStream_t st,
                                 bool* served) {
  *served = false;
  if (!jit_enabled()) return cudaSuccess;
  ChainGeometry geo[2];
  dense_kmn_geometry(MC, D, H, geo);
  if (!dense_kmn_eligible(MC, D, H, geo)) return cudaSuccess;
  const char* const names[2] = {"nfn_jit_dense_kmn_fwd", "nfn_jit_dense_kmn_fwd_bwd"};
  const std::string ckey = dense_kmn_key(MC, D, H) + "|m" + std::to_string(mode) + "|dev" + std::to_string(device_info().device);
  JitEntry* ent = get_or_build(ckey, dense_kmn_program_source(MC, D, H, mode, geo), names, geo);
  if (!ent) return cudaSuccess;
  cudaError_t ce = launch_entry(ent, bwd ? 1 : 0, a, a.B, st);
  if (ce == cudaSuccess) *served = true;
  return ce;
}

long long jit_dense_kmn_compile_check(int MC, int D, int H, int mode, std::string& log) {
  ChainGeometry geo[2];
  dense_kmn_geometry(MC, D, H, geo);
  if (!dense_kmn_eligible(MC, D, H, geo)) {
    log = "kernel mixture / hidden width not eligible for the fused kernel";
    return -1;
  }
  std::vector<char> cubin;
  if (!compile_cubin(dense_kmn_program_source(MC, D, H, mode, geo), cubin, log)) return -1;
  return (long long)cubin.size();
}

}  // namespace nfn
