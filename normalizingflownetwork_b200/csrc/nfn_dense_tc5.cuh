// nfn_dense_tc5.cuh -- Dense(P) + flow chain with the GEMMs on tcgen05 / TMEM (sm_100a).
//
// Same contract as dense_chain_kernel (nfn_dense_chain.cuh): per tile of 128 rows
//     t  = h W + b ;  flows forward + reverse sweep per row ;  dh = dt W^T ;  dW += h^T dt ;  db += 1^T dt
// but the GEMMs are single-thread-issued tcgen05.mma (kind::f16 on exact three-level bf16 splits,
// M = 128) with the accumulators in tensor memory, so the per-thread mma.sync fragment work (loads,
// splits, shuffles: >60 % of the instructions of the warp-level version) disappears:
//
//   the h row is split into 3 bf16 levels -> K-major operand tiles in smem
//   GEMM 1   D1[128 x PN]  = sum of the level products  A_i[128 x H] * W_j[H x PN]  +  1 * b   (fp32-grade t;
//                            the bias rides in as one more product: a ones block of the h tile against a
//                            three-level bias operand)
//   tcgen05.ld 32x32b hands thread r exactly row r of D1 -> per-row flows (nfn_flows.cuh)
//   the dt row is split into 3 bf16 levels -> operand tiles D_i [128 x PN]
//   GEMM 2   D2[128 x 3H]  = [dt W0^T | dt W1^T | dt W2^T]   (A = D_i K-major,  B = W[H][P] as N x K)
//   GEMM 3   D3[.. x 3H+16] = [dt^T h0 | dt^T 1 | dt^T h1 | dt^T h2]  (A = the dt tile read MN-major: one M = 128
//                              instruction spans 16 chunks of the [level 0 | level 1 | level 2] row groups, so
//                              the dt levels ride in different TMEM lanes; B = the h tile read MN-major with
//                              its three levels and a ones block side by side: dW^T and db in one pass; a pass
//                              that only holds level-2 chunks of dt stops after [h0 | 1]: N = H + 16)
//   D2 row r -> dh[r, :] ; the D3 lanes (level, p) leave as atomics into dW[:, p], db[p].
//
// Two pipelines share these pieces:
//   * pipe 1 (forward-only kernels, and backward shapes whose accumulators do not fit twice): the 128 compute
//     threads do everything but the issuing; a second warpgroup's first lane issues the MMAs.
//   * pipe 2 (backward, the default): the second warpgroup also WORKS.  Its thread r loads and splits the h row,
//     splits the dt row, drains dh and dW; the compute threads only run the flows and hand their dt row over
//     through TENSOR MEMORY (tcgen05.st into the D1 buffer their t row came from, double-buffered).  The compute
//     warps lose a third of their instructions and the SM gains eight more warps to issue from.
//
// Why bf16 levels and not TF32: the backward GEMMs contract over the ROWS of the tile, i.e. they need the
// dt and h tiles transposed.  With the no-swizzle canonical layout (8 x 16-byte core matrices) an
// X[128][C] tile written K-major is, byte for byte, X^T in the MN-major layout -- but tcgen05 reads
// MN-major TF32 operands only in the 128B/32B-swizzled layout (measured: tools/umma_probe.cu returns
// zeros otherwise), which no K-major layout matches.  16-bit operands have no such restriction.  An fp32
// value is EXACTLY x0 + x1 + x2 with three truncated bf16 levels (8 + 8 + 8 significand bits); products
// of levels i + j <= 2 carry everything above 2^-24 relative, and K = 16 per instruction halves the
// instruction count.
#pragma once
#ifndef __CUDACC_RTC__
#include <cstdint>
#else
typedef unsigned long long uint64_t;   // the runtime specialiser (NVRTC) has no host headers
typedef unsigned int uint32_t;
#endif

#include "nfn_dense_chain.cuh"

#ifndef NFN_TC5_PIPE2
#define NFN_TC5_PIPE2 1   // 0: every backward kernel on pipe 1 (A/B builds)
#endif
#ifndef NFN_TC5_HREGS
#define NFN_TC5_HREGS 80  // registers per thread of the staging warpgroup on pipe 2 (the compute warpgroup gets 256 - this)
#endif

namespace nfn {
namespace tc5 {

constexpr int kRows = 128;     // rows per tile == compute threads per CTA == TMEM lanes
// + a second warpgroup: its first lane issues every tcgen05.mma, and on pipe 2 its threads do the operand
// staging.  A whole warpgroup because registers are handed out per warpgroup (setmaxnreg).
constexpr int kThreads = 256;
__host__ __device__ constexpr int round16(int x) { return (x + 15) / 16 * 16; }
__host__ __device__ constexpr unsigned pow2_cols(int c) { return c <= 32 ? 32u : c <= 64 ? 64u : c <= 128 ? 128u : c <= 256 ? 256u : 512u; }

// byte offset of element (row, col) in a K-major no-swizzle bf16 operand tile with C columns:
// 8-row x 16-byte core matrices, the C/8 matrices of one 8-row group contiguous
__host__ __device__ constexpr unsigned tile_off(int row, int col, int C) {
  return (unsigned)((row >> 3) * (C / 8 * 128) + (col >> 3) * 128 + (row & 7) * 16 + (col & 7) * 2);
}

// ---- geometry as plain constexpr functions of (P, H, backward?): shared by the kernel (Geo below), the
// ahead-of-time launcher and the runtime specialiser, which only knows P and H at run time
// h tile, per 8-row group: [level 0 | ones | level 1 | level 2] x (H/8 chunks each, 2 for the ones block)
__host__ __device__ constexpr unsigned g_kLvlA(int H) { return (unsigned)(H / 8 * 128); }          // one h level inside an 8-row group
__host__ __device__ constexpr unsigned g_kGrpA(int H) { return 3 * g_kLvlA(H) + 256; }
__host__ __device__ constexpr unsigned g_lvlA(int H, int i) { return i == 0 ? 0u : (unsigned)i * g_kLvlA(H) + 256u; }
__host__ __device__ constexpr unsigned g_onesA(int H) { return g_kLvlA(H); }
__host__ __device__ constexpr unsigned g_kA(int H) { return (unsigned)(kRows / 8) * g_kGrpA(H); }  // one h tile
__host__ __device__ constexpr unsigned g_kD(int P) { return (unsigned)(kRows * round16(P) * 2); }  // one bf16 level of the dt tile
__host__ __device__ constexpr unsigned g_kW(int P, int H) { return (unsigned)(round16(P) * H * 2); }
__host__ __device__ constexpr unsigned g_kBiasT(int P) { return (unsigned)(round16(P) * 32); }     // bias operand [PN x 16] bf16
__host__ __device__ constexpr int g_NA(bool bwd) { return bwd ? 3 : 2; }
__host__ __device__ constexpr unsigned g_oD(int H, bool bwd) { return (unsigned)g_NA(bwd) * g_kA(H); }
__host__ __device__ constexpr unsigned g_oW1(int P, int H, bool bwd) { return g_oD(H, bwd) + (bwd ? 3 * g_kD(P) : 0); }
__host__ __device__ constexpr unsigned g_oW2(int P, int H, bool bwd) { return g_oW1(P, H, bwd) + 3 * g_kW(P, H); }
__host__ __device__ constexpr unsigned g_oBias(int P, int H, bool bwd) { return g_oW2(P, H, bwd) + (bwd ? 3 * g_kW(P, H) : 0); }
__host__ __device__ constexpr unsigned g_oBar(int P, int H, bool bwd) { return g_oBias(P, H, bwd) + g_kBiasT(P); }
__host__ __device__ constexpr unsigned smem_bytes(int P, int H, bool bwd) { return g_oBar(P, H, bwd) + 80; }
// GEMM 3 reads the dt tile MN-major with M = 128 = 16 eight-column chunks per pass; the three levels of a row
// group are 3 * PN/8 consecutive chunks, so ceil(3 PN / 128) passes cover them (1 for P <= 32, 2 for P <= 80).
// A pass that starts inside level 2 only needs the [h0 | ones] columns: level 2 x (h1, h2) is below 2^-24.
__host__ __device__ constexpr int g_NP3(int P) { return (3 * (round16(P) / 8) + 15) / 16; }
__host__ __device__ constexpr bool g_short3(int P, int j) { return 16 * j >= 2 * (round16(P) / 8); }
__host__ __device__ constexpr int g_N3(int P, int H, int j) { return g_short3(P, j) ? H + 16 : 3 * H + 16; }
__host__ __device__ constexpr int g_cols3(int P, int H, int npass) {   // TMEM columns of passes [0, npass)
  int c = 0;
  for (int j = 0; j < npass; ++j) c += g_N3(P, H, j);
  return c;
}
// pipe 2 needs two D1 buffers next to D2 and D3
__host__ __device__ constexpr bool g_pipe2(int P, int H, bool bwd) {
  return bwd && NFN_TC5_PIPE2 != 0 && 2 * round16(P) + 3 * H + g_cols3(P, H, g_NP3(P)) <= 256;
}
__host__ __device__ constexpr int g_nD1(int P, int H, bool bwd) { return g_pipe2(P, H, bwd) ? 2 : 1; }
__host__ __device__ constexpr int g_cD2(int P, int H, bool bwd) { return g_nD1(P, H, bwd) * round16(P); }
__host__ __device__ constexpr int g_cD3(int P, int H, bool bwd) { return g_cD2(P, H, bwd) + 3 * H; }
__host__ __device__ constexpr unsigned tmem_cols(int P, int H, bool bwd) {
  return pow2_cols(bwd ? g_cD3(P, H, bwd) + g_cols3(P, H, g_NP3(P)) : round16(P));
}
// Register budget for MINB resident CTAs: the launch-time allocation is 65536 / (MINB * 256) per thread
// (rounded down to 8); the second warpgroup keeps regs_issuer of it and the compute warpgroup gets the rest.
// The sum must never exceed the CTA's pool, or setmaxnreg.inc would wait forever.
__host__ __device__ constexpr int regs_launch(int minb) { return 65536 / (minb * kThreads) / 8 * 8; }
__host__ __device__ constexpr int regs_issuer(int minb, bool pipe2) { return pipe2 ? NFN_TC5_HREGS : (minb <= 2 ? 56 : 32); }
__host__ __device__ constexpr int regs_compute(int minb, bool pipe2) { return 2 * regs_launch(minb) - regs_issuer(minb, pipe2); }
// resident CTAs per SM the kernel's register plan is built for (shared memory and TMEM columns permitting)
__host__ __device__ constexpr int min_blocks(int P, int H, bool bwd) {
  const int by_smem = (int)((227u * 1024u) / (smem_bytes(P, H, bwd) + 1024u));
  const int by_tmem = (int)(512u / tmem_cols(P, H, bwd));
  const int want = bwd ? 2 : 3;   // compute warpgroup registers: 184-200 (fwd+bwd), 128 (forward)
  const int cap = by_smem < by_tmem ? by_smem : by_tmem;
  // never below 2: this value fixes the register plan (256 threads x 128 registers at launch, re-split by
  // setmaxnreg); when shared memory allows a single CTA the plan is simply the 2-CTA one
  return cap < 2 ? 2 : (cap < want ? cap : want);
}

// Resident CTAs per SM to size the grid with.  Computed here, not asked of the occupancy API:
// cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for every kernel that allocates tensor memory
// (measured), although shared memory, registers and TMEM columns all allow more.
__host__ __device__ constexpr int resident_ctas(int P, int H, bool bwd) {
  const int by_smem = (int)((227u * 1024u) / (smem_bytes(P, H, bwd) + 1024u));
  const int by_tmem = (int)(512u / tmem_cols(P, H, bwd));
  const int by_regs = min_blocks(P, H, bwd);   // the register plan: 65536 / (256 threads x launch registers)
  const int c = by_smem < by_tmem ? (by_smem < by_regs ? by_smem : by_regs) : (by_tmem < by_regs ? by_tmem : by_regs);
  return c < 1 ? 1 : c;
}

// shared-memory carve-up (bytes); every tile is 128-byte aligned
template <int P, int H, bool BWD>
struct Geo {
  static constexpr int PN = round16(P);            // parameter columns padded to the MMA N / K granule
  // h tile: per 8-row group [level 0 | ones | level 1 | level 2], so that read MN-major it is ONE operand
  // [h0 | 1 1 1 0 .. | h1 | h2] with N = 3H + 16 columns (GEMM 3 + bias gradient in one pass; the first H + 16
  // of them for a level-2-only pass), and read K-major level i is the tile at byte offset lvlA(i) (GEMM 1)
  static constexpr unsigned kLvlA = g_kLvlA(H);
  static constexpr unsigned kGrpA = g_kGrpA(H);
  static constexpr unsigned kOnesA = g_onesA(H);
  __host__ __device__ static constexpr unsigned lvlA(int i) { return g_lvlA(H, i); }
  static constexpr unsigned kA = g_kA(H);
  // dt tile: per 8-row group [level 0 | level 1 | level 2] x (PN/8 chunks each).  Read K-major, level i is the
  // tile at byte offset i * kLvlD (GEMM 2); read MN-major with M = 128, one instruction sees 16 consecutive
  // chunks = SEVERAL LEVELS AT ONCE, each landing in its own TMEM lanes (GEMM 3 needs NP3 passes, not 3)
  static constexpr int CL = PN / 8;                                // chunks per level
  static constexpr unsigned kLvlD = CL * 128;
  static constexpr unsigned kGrpD = 3 * kLvlD;
  static constexpr int NP3 = g_NP3(P);
  __host__ __device__ static constexpr bool short3(int j) { return g_short3(P, j); }
  __host__ __device__ static constexpr int N3(int j) { return g_N3(P, H, j); }
  static constexpr unsigned kD = g_kD(P);          // bytes of one level over the whole tile (tile = 3 kD)
  static constexpr unsigned kW = g_kW(P, H);       // one level of W, either orientation
  // A ring of h tiles: GEMM 1 of tile i+1 is issued while tile i is still in its flows, and (BWD) GEMM 3 of
  // tile i-1, which reads h(i-1), may still be running then
  static constexpr int NA = g_NA(BWD);
  static constexpr unsigned oA = 0;
  static constexpr unsigned oD = g_oD(H, BWD);                     // 3 levels (BWD)
  static constexpr unsigned oW1 = g_oW1(P, H, BWD);                // W as [N = PN][K = H], 3 levels
  // W as [N = H][K = PN], 3 levels back to back == ONE operand [W0; W1; W2] with N = 3H rows (GEMM 2)
  static constexpr unsigned oW2 = g_oW2(P, H, BWD);
  static constexpr unsigned oBias = g_oBias(P, H, BWD);            // [N = PN][K = 16]: bias levels in k = 0, 1, 2
  static constexpr unsigned oBar = g_oBar(P, H, BWD);              // mbarriers + tmem base
  static constexpr unsigned kBytes = smem_bytes(P, H, BWD);
  // MN-major reads of the dt tiles with M = 128 run (16 - PN/8) chunks past the tile: what follows must be ours
  static_assert(!BWD || 6 * kW >= (unsigned)(16 * NP3 - 3 * CL) * 128, "operand over-read must stay in the CTA's smem");
  // TMEM columns: D1 [PN] (x 2 on pipe 2) | D2 [3H: dt W0^T | dt W1^T | dt W2^T] | D3 [passes: dt^T h0 | db .. | dt^T h1 | dt^T h2]
  static constexpr bool kPipe2 = g_pipe2(P, H, BWD);
  static constexpr int cD1 = 0, cD2 = g_cD2(P, H, BWD), cD3 = g_cD3(P, H, BWD);
  __host__ __device__ static constexpr int cD3p(int j) { return cD3 + g_cols3(P, H, j); }
  static constexpr unsigned kCols = tmem_cols(P, H, BWD);
  static_assert(3 * H <= 256 && 3 * H + 16 <= 256, "MMA N limit");
};

// ------------------------------------------------------------------ PTX wrappers
NFN_DEVI void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
NFN_DEVI void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
NFN_DEVI void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
NFN_DEVI void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
NFN_DEVI void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
NFN_DEVI bool mbar_try_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded (2 s): a tensor-core op that never completes must not hang the GPU (the launch fails instead)
NFN_DEVI void mbar_wait(unsigned bar, unsigned parity) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
      if (mbar_try_wait(bar, parity)) return;
    }
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 2000000000ull) __trap();
  }
}
NFN_DEVI void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// the 128 threads of the second warpgroup (named barrier 1; barrier 0 is __syncthreads)
NFN_DEVI void helper_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
NFN_DEVI void tmem_alloc(unsigned smem_dst, unsigned cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
NFN_DEVI void tmem_dealloc(unsigned taddr, unsigned cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// shared-memory matrix descriptor, no swizzle, sm_100 version field set.
//   K-major : lbo = bytes between the two 16-byte K chunks of one instruction, sbo = between 8-row groups
//   MN-major: lbo = bytes between 8-deep K groups,                           sbo = between 8-wide MN chunks
// (measured with tools/umma_probe.cu)
NFN_DEVI uint64_t smem_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes) {
  return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor, kind::f16 with bf16 operands, fp32 accumulate
__host__ __device__ constexpr uint32_t instr_desc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
NFN_DEVI void mma_bf16(unsigned d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one lane of a converged warp
NFN_DEVI bool elect_one() {
  unsigned p;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(p));
  return p != 0;
}
NFN_DEVI void mma_commit(unsigned bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 16 / 8 consecutive columns of this thread's TMEM lane
NFN_DEVI void tmem_ld16(unsigned taddr, unsigned (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
NFN_DEVI void tmem_ld8(unsigned taddr, unsigned (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
NFN_DEVI void tmem_st16(unsigned taddr, const unsigned (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
NFN_DEVI void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
NFN_DEVI void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// N columns (multiple of 16) of this thread's lane -> floats; the values are pinned behind the wait
template <int N>
NFN_DEVI void tmem_load_row(unsigned taddr, float (&out)[N]) {
  static_assert(N % 16 == 0, "16-column granules");
  unsigned r[N / 16][16];
#pragma unroll
  for (int i = 0; i < N / 16; ++i) tmem_ld16(taddr + 16u * i, r[i]);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < N / 16; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      asm volatile("" : "+r"(r[i][j]));  // a use the compiler cannot hoist above the wait
      out[16 * i + j] = __uint_as_float(r[i][j]);
    }
}
// 8 columns -> floats (one wait per call)
NFN_DEVI void tmem_load8(unsigned taddr, float (&out)[8]) {
  unsigned r[8];
  tmem_ld8(taddr, r);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    asm volatile("" : "+r"(r[j]));
    out[j] = __uint_as_float(r[j]);
  }
}
// three 8-column blocks with one wait
NFN_DEVI void tmem_load8x3(unsigned t0, unsigned t1, unsigned t2, float (&o0)[8], float (&o1)[8], float (&o2)[8]) {
  unsigned r0[8], r1[8], r2[8];
  tmem_ld8(t0, r0);
  tmem_ld8(t1, r1);
  tmem_ld8(t2, r2);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    asm volatile("" : "+r"(r0[j]), "+r"(r1[j]), "+r"(r2[j]));
    o0[j] = __uint_as_float(r0[j]);
    o1[j] = __uint_as_float(r1[j]);
    o2[j] = __uint_as_float(r2[j]);
  }
}
// floats -> N columns of this thread's lane (the caller waits: tmem_st_wait)
template <int N>
NFN_DEVI void tmem_store_row(unsigned taddr, const float (&in)[N]) {
  static_assert(N % 16 == 0, "16-column granules");
#pragma unroll
  for (int i = 0; i < N / 16; ++i) {
    unsigned r[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(in[16 * i + j]);
    tmem_st16(taddr + 16u * i, r);
  }
}
NFN_DEVI void sts_u4(unsigned saddr, unsigned a, unsigned b, unsigned c, unsigned d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
NFN_DEVI void sts_u16(unsigned saddr, unsigned v) {
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(saddr), "h"((unsigned short)v) : "memory");
}

// ---- packed fp32 pairs (FADD2 on sm_100): one instruction for two independent IEEE additions
NFN_DEVI void sub2(float a0, float a1, float b0, float b1, float& r0, float& r1) {
  unsigned long long a, b, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(r));
}
NFN_DEVI void add2(float a0, float a1, float b0, float b1, float& r0, float& r1) {
  unsigned long long a, b, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(r));
}

// ---- exact three-level bf16 split of an fp32 value: x = x0 + x1 + x2, every level a truncated bf16
// (8 significand bits each; the third remainder has at most 8 bits left, so nothing is lost)
struct Bf3 {
  unsigned b0, b1, b2;  // fp32 bit patterns whose low 16 bits are zero
};
NFN_DEVI Bf3 split_bf3(float x) {
  Bf3 s;
  s.b0 = __float_as_uint(x) & 0xffff0000u;
  const float r1 = x - __uint_as_float(s.b0);
  s.b1 = __float_as_uint(r1) & 0xffff0000u;
  s.b2 = __float_as_uint(r1 - __uint_as_float(s.b1)) & 0xffff0000u;
  return s;
}
// two bf16 (high halves of two fp32 patterns) -> one 32-bit word, element `even` at the lower address
NFN_DEVI unsigned pack_bf(unsigned even, unsigned odd) { return __byte_perm(even, odd, 0x7632); }
// the same split for two values at once, straight to the packed words of the three level tiles: the byte
// permute takes the high halves of the raw patterns (no mask needed for the word), the remainders of both
// values come from one packed subtraction: 3 PRMT + 4 LOP + 2 FADD2 per pair instead of 3 + 6 + 4
NFN_DEVI void split_pair(float e, float o, unsigned& w0, unsigned& w1, unsigned& w2) {
  const unsigned eb = __float_as_uint(e), ob = __float_as_uint(o);
  w0 = pack_bf(eb, ob);
  float r1e, r1o, r2e, r2o;
  sub2(e, o, __uint_as_float(eb & 0xffff0000u), __uint_as_float(ob & 0xffff0000u), r1e, r1o);
  const unsigned e1 = __float_as_uint(r1e), o1 = __float_as_uint(r1o);
  w1 = pack_bf(e1, o1);
  sub2(r1e, r1o, __uint_as_float(e1 & 0xffff0000u), __uint_as_float(o1 & 0xffff0000u), r2e, r2o);
  w2 = pack_bf(__float_as_uint(r2e), __float_as_uint(r2o));
}

// 8 consecutive fp32 values -> one 16-byte chunk in each of the three level tiles (byte offsets o1, o2 of
// levels 1 and 2 relative to level 0)
NFN_DEVI void store_levels8(unsigned saddr, unsigned o1, unsigned o2, const float* v) {
  unsigned w0[4], w1[4], w2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_pair(v[2 * j], v[2 * j + 1], w0[j], w1[j], w2[j]);
  sts_u4(saddr, w0[0], w0[1], w0[2], w0[3]);
  sts_u4(saddr + o1, w1[0], w1[1], w1[2], w1[3]);
  sts_u4(saddr + o2, w2[0], w2[1], w2[2], w2[3]);
}

// level pairs (i, j) of a split product, smallest contribution first.  All nine reproduce the fp32
// product exactly; the six with i + j <= 2 carry everything above 2^-24 relative.
__device__ constexpr int kPairs9[9][2] = {{2, 2}, {1, 2}, {2, 1}, {0, 2}, {2, 0}, {1, 1}, {0, 1}, {1, 0}, {0, 0}};

// ------------------------------------------------------------------ pieces shared by the two pipelines
// One-time staging by all 256 threads: split weight tiles in both orientations, the bias operand and the
// ones blocks of the h tiles (written once; the level blocks around them are rewritten per tile).
template <int P, int H, bool BWD>
NFN_DEVI void stage_constants(const DenseArgs& a, unsigned sbase, int tid) {
  using G = Geo<P, H, BWD>;
  constexpr int PN = G::PN, NT = kThreads, T = kRows;
  for (int i = tid; i < PN * H; i += NT) {   // W1[n][k] = W[k][n]
    const int n = i / H, k = i % H;
    const Bf3 w = split_bf3((n < P) ? __ldg(a.W + k * P + n) : 0.0f);
    const unsigned o = sbase + G::oW1 + tile_off(n, k, H);
    sts_u16(o, w.b0 >> 16);
    sts_u16(o + G::kW, w.b1 >> 16);
    sts_u16(o + 2 * G::kW, w.b2 >> 16);
  }
  // bias operand [n][k]: the three levels of b[n] in k = 0, 1, 2 against the ones block's 1 1 1 0 .. 0
  for (int i = tid; i < PN * 16; i += NT) {
    const int n = i / 16, k = i % 16;
    unsigned v = 0;
    if (k < 3 && n < P) {
      const Bf3 b = split_bf3(__ldg(a.bias + n));
      v = (k == 0 ? b.b0 : (k == 1 ? b.b1 : b.b2)) >> 16;
    }
    sts_u16(sbase + G::oBias + tile_off(n, k, 16), v);
  }
  // ones block of every h tile: per row 16 bf16 = 1 1 1 0 .. 0 (read K-major it multiplies the bias operand in
  // GEMM 1; read MN-major its first column yields the bias gradient in GEMM 3, the next two repeat it, unused)
  for (int i = tid; i < G::NA * T * 16; i += NT) {
    const int b = i / (T * 16), r = (i / 16) % T, k = i % 16;
    sts_u16(sbase + G::oA + b * G::kA + (r >> 3) * G::kGrpA + G::kOnesA + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2,
            k < 3 ? 0x3f80u : 0u);
  }
  if constexpr (BWD) {
    for (int i = tid; i < H * PN; i += NT) {  // W2[n][k] = W[n][k]
      const int n = i / PN, k = i % PN;
      const Bf3 w = split_bf3((k < P) ? __ldg(a.W + n * P + k) : 0.0f);
      const unsigned o = sbase + G::oW2 + tile_off(n, k, PN);
      sts_u16(o, w.b0 >> 16);
      sts_u16(o + G::kW, w.b1 >> 16);
      sts_u16(o + 2 * G::kW, w.b2 >> 16);
    }
  }
}

NFN_DEVI unsigned uniform_base(unsigned x) {
  asm volatile("" : "+r"(x));
  return __shfl_sync(0xffffffffu, x, 0);
}

// GEMM 1 of one tile: the six level products with i + j <= 2 and the bias product, into D1 at column c1.
// Runs warp-converged on warp-uniform values (so the operand descriptors live in uniform registers); only
// the tcgen05 instructions themselves are predicated on the elected lane.
template <int P, int H, bool BWD>
NFN_DEVI void issue_gemm1(unsigned d1, unsigned sbase, unsigned a_off, unsigned bar) {
  using G = Geo<P, H, BWD>;
  const bool leader = elect_one();   // elected HERE: a predicate carried in from a merge point makes ptxas wrap every MMA in an election loop
  constexpr uint32_t kI1 = instr_desc(128, G::PN, 0, 0);
  // operand descriptors of level 0 / k-step 0; the others differ by a constant in the address field.  Built
  // HERE, per call, from a base the compiler must re-read (the empty asm) and knows to be warp-uniform (the
  // broadcast shuffle): hoisted out of the tile loop they would sit in -- and spill from -- the vector
  // registers of all 128 threads of the warpgroup, and every MMA would need a lane-uniformisation loop.
  sbase = uniform_base(sbase);
  const uint64_t dA_k = smem_desc(sbase + G::oA, 128, G::kGrpA);      // h tile, one level K-major
  const uint64_t dW1 = smem_desc(sbase + G::oW1, 128, H / 8 * 128);   // W as [PN][H], one level
  const uint64_t dB = smem_desc(sbase + G::oBias, 128, 256);          // bias operand [PN x 16]
  tc_fence_after();
  unsigned acc = 0;
#pragma unroll
  for (int q = 3; q < 9; ++q) {   // smallest first; the bias goes in just before the leading product
    if (q == 8) {
      if (leader) mma_bf16(d1, dA_k + (uint64_t)((a_off + G::kOnesA) >> 4), dB, kI1, acc);
      acc = 1;
    }
#pragma unroll
    for (int ks = 0; ks < H / 16; ++ks) {
      const uint64_t ad = dA_k + (uint64_t)((a_off + G::lvlA(kPairs9[q][0]) + ks * 256) >> 4);
      const uint64_t bd = dW1 + (uint64_t)((kPairs9[q][1] * G::kW + ks * 256) >> 4);
      if (leader) mma_bf16(d1, ad, bd, kI1, acc);
      acc = 1;
    }
  }
  if (leader) mma_commit(bar);
  __syncwarp();
}

// GEMM 2 + GEMM 3 of one tile (backward); `fresh` starts a new accumulation window in D3
template <int P, int H>
NFN_DEVI void issue_gemm23(unsigned tmem_base, unsigned sbase, unsigned a_off, bool fresh, unsigned bar) {
  using G = Geo<P, H, true>;
  const bool leader = elect_one();
  constexpr int PN = G::PN;
  constexpr uint32_t kI2 = instr_desc(128, 3 * H, 0, 0);
  sbase = uniform_base(sbase);   // see issue_gemm1
  const uint64_t dA_mn = smem_desc(sbase + G::oA, G::kGrpA, 128);     // h tile, all levels + ones MN-major (GEMM 3)
  const uint64_t dD_k = smem_desc(sbase + G::oD, 128, G::kGrpD);      // dt tile, one level K-major (GEMM 2)
  const uint64_t dD_mn = smem_desc(sbase + G::oD, G::kGrpD, 128);     // dt tile, 16 chunks across levels MN-major (GEMM 3)
  const uint64_t dW2 = smem_desc(sbase + G::oW2, 128, PN / 8 * 128);  // [W0; W1; W2], N = 3H
  tc_fence_after();
  // GEMM 2: [dt W0^T | dt W1^T | dt W2^T] = dt_i [W0; W1; W2]^T, dt levels smallest first (all 9 products)
  // (the level loop stays rolled: this thread has few registers; the k-steps are unrolled so the descriptor
  // moves into uniform registers pipeline)
  {
    unsigned acc = 0;
#pragma unroll 1
    for (int lv = 2; lv >= 0; --lv) {
#pragma unroll
      for (int ks = 0; ks < PN / 16; ++ks) {
        if (leader)
          mma_bf16(tmem_base + G::cD2, dD_k + (uint64_t)((lv * G::kLvlD + ks * 256) >> 4), dW2 + (uint64_t)((ks * 256) >> 4),
                   kI2, acc);
        acc = 1;
      }
    }
  }
  // GEMM 3: [dt^T h0 | dt^T 1 .. | dt^T h1 | dt^T h2] over the tile's 128 rows, 16 rows (two 8-row groups) per
  // instruction: dW^T and the bias gradient in one pass.  M = 128 spans 16 chunks of the [level 0 | level 1 |
  // level 2] groups, so the levels ride in different lanes of the same instruction.
#pragma unroll
  for (int j = 0; j < G::NP3; ++j) {
    const uint32_t kI3 = instr_desc(128, G::N3(j), 1, 1);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      if (leader)
        mma_bf16(tmem_base + G::cD3p(j), dD_mn + (uint64_t)((j * 2048 + ks * 2 * G::kGrpD) >> 4),
                 dA_mn + (uint64_t)((a_off + ks * 2 * G::kGrpA) >> 4), kI3, (ks > 0 || !fresh) ? 1u : 0u);
    }
  }
  if (leader) mma_commit(bar);
  __syncwarp();
}

// ------------------------------------------------------------------ pipe 1: compute threads do the staging
template <class Spec, int H, bool BWD, class M, int MINB>
NFN_DEVI void dense_tc5_body1(const DenseArgs& a) {
  constexpr int kRegsIssuer = regs_issuer(MINB, false), kRegsCompute = regs_compute(MINB, false);
  static_assert(kRows * (kRegsIssuer + kRegsCompute) <= kThreads * regs_launch(MINB), "register pool");
  static_assert(kRegsCompute <= 232 && kRegsCompute % 8 == 0 && kRegsIssuer % 8 == 0, "setmaxnreg range");
  constexpr int D = Spec::D;
  constexpr int P = Spec::P();
  using G = Geo<P, H, BWD>;
  constexpr int PN = G::PN, T = kRows, NT = kThreads;
  // the parameter row lives in REGISTERS (straight out of TMEM): every access of the flow code is a
  // compile-time index, scalar "vector width" 1 keeps it that way
  constexpr int V = 1;
  constexpr unsigned kGrpA = G::kGrpA;       // bytes between 8-row groups of the h tile
  constexpr unsigned kGrpD = G::kGrpD;       // ... of the dt tile (three levels side by side)

  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double red[NT / 32];
  const unsigned sbase = smem_u32(smem_raw);
  // bar1 / bar2: tensor pipe -> compute threads (GEMM 1 done / GEMM 2+3 done);
  // bar_h / bar_d: compute threads -> issuing thread (h tile written and D1 consumed / dt tiles written and D2, D3 drained)
  const unsigned bar1 = sbase + G::oBar, bar2 = bar1 + 8, bar_h = bar1 + 16, bar_d = bar1 + 24, tmem_slot = bar1 + 64;

  const int tid = threadIdx.x, warp = tid >> 5;
  const long long ntiles = (a.B + T - 1) / T;

  // ---- one-time set-up: barriers, TMEM, split weight tiles, bias operand, ones blocks
  if (tid == 0) {
    mbar_init(bar1, 1);
    mbar_init(bar2, 1);
    mbar_init(bar_h, T);
    mbar_init(bar_d, T);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, G::kCols);
  stage_constants<P, H, BWD>(a, sbase, tid);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  unsigned tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);   // the same value in every lane: say so
  const unsigned lane_base = tmem_base + ((unsigned)(warp * 32) << 16);

  // this thread's row of the next tile: h (registers), y, upstream cotangent
  float h_nxt[H];
  float y_nxt[D];
  float g_nxt = 1.0f;
  auto fetch_h = [&](long long tile) {
    const long long r = tile * T + tid;
#pragma unroll
    for (int i = 0; i < H; ++i) h_nxt[i] = 0.0f;
    if (tile < ntiles && r < a.B) {
      const float4* src = reinterpret_cast<const float4*>(a.h + r * H);
#pragma unroll
      for (int c = 0; c < H / 4; ++c) {
        const float4 v = __ldg(src + c);
        h_nxt[4 * c] = v.x; h_nxt[4 * c + 1] = v.y; h_nxt[4 * c + 2] = v.z; h_nxt[4 * c + 3] = v.w;
      }
    }
  };
  auto fetch_y = [&](long long tile) {
    const long long r = tile * T + tid;
    if (tile < ntiles && r < a.B) {
      load_event<D>(a.y, a.y_broadcast ? 0 : r, y_nxt);
      if constexpr (BWD) { if (a.g_logp) g_nxt = __ldg(a.g_logp + r); }
    }
  };
#pragma unroll
  for (int i = 0; i < D; ++i) y_nxt[i] = 0.0f;
  long long tile = blockIdx.x;

  const unsigned a_row = sbase + G::oA + (tid >> 3) * kGrpA + (tid & 7) * 16;  // this thread's row in the h tile
  const unsigned d_row = sbase + G::oD + (tid >> 3) * kGrpD + (tid & 7) * 16;   // ... in the dt tile
  float ls_hi = 0.0f, ls_lo = 0.0f;   // this thread's sum of logp (compensated)
  // GEMM 3, pass j: TMEM lane m holds chunk 16 j + m / 8 of the [level 0 | level 1 | level 2] row groups, i.e.
  // the contribution of ONE bf16 level of dt[:, p] to dW[:, p] and db[p].  It accumulates in tensor memory over
  // windows of kFlush tiles (bounding the number of in-place fp32 accumulations), then goes out as atomics,
  // which also sum the levels and the CTAs.
  // Narrow chains (one pass) have registers to spare: they drain D3 every tile into register accumulators and
  // issue their atomics once, at the end (measured 69 vs 74 us at P = 17); wide ones use the TMEM windows
  // (88 vs 94 us at P = 48, where the second accumulator set spilled).
  constexpr bool kRegAcc = (G::NP3 == 1);
  constexpr unsigned kFlush = kRegAcc ? 1u : 16u;
  float dw_acc[kRegAcc ? H : 1], db_acc = 0.0f;
#pragma unroll
  for (int k = 0; k < (kRegAcc ? H : 1); ++k) dw_acc[k] = 0.0f;
  auto split_h = [&](int buf) {
#pragma unroll
    for (int c = 0; c < H / 8; ++c) store_levels8(a_row + buf * G::kA + c * 128, G::lvlA(1), G::lvlA(2), h_nxt + 8 * c);
  };
  // dh row of a finished tile out of TMEM (sum of the three W-level blocks, smallest first) -> global
  auto drain_backward = [&](long long r_done) {
    // 16 hidden columns at a time: three level blocks in, one sum out (bounded register footprint for wide H)
#pragma unroll
    for (int c = 0; c < H / 16; ++c) {
      float b0[16], b1[16], b2[16];
      tmem_load_row<16>(lane_base + G::cD2 + 16 * c, b0);
      tmem_load_row<16>(lane_base + G::cD2 + H + 16 * c, b1);
      tmem_load_row<16>(lane_base + G::cD2 + 2 * H + 16 * c, b2);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        add2(b2[2 * q], b2[2 * q + 1], b1[2 * q], b1[2 * q + 1], b1[2 * q], b1[2 * q + 1]);
        add2(b1[2 * q], b1[2 * q + 1], b0[2 * q], b0[2 * q + 1], b0[2 * q], b0[2 * q + 1]);
      }
      if (r_done < a.B) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          st_stream_f4(a.dh + r_done * H + 16 * c + 4 * q, make_float4(b0[4 * q], b0[4 * q + 1], b0[4 * q + 2], b0[4 * q + 3]));
      }
    }
  };
  // the dW / db lanes (TMEM lane = (level, p)) -> register accumulators or atomics; warps past the last chunk skip
  auto flush_dw = [&]() {
#pragma unroll
    for (int j = 0; j < G::NP3; ++j) {
      if (16 * j + 4 * warp < 3 * G::CL) {   // this warp's 4 chunks of pass j exist (warp-uniform)
        const int q = 16 * j + (tid >> 3), p = (q % G::CL) * 8 + (tid & 7);   // lane -> (level, column p)
        const bool live = q < 3 * G::CL && p < P;
        const unsigned c0 = lane_base + G::cD3p(j);
#pragma unroll
        for (int c = 0; c < H / 16; ++c) {
          float b0[16];
          tmem_load_row<16>(c0 + 16 * c, b0);
          if (!G::short3(j)) {   // compile-time after unrolling
            float b1[16], b2[16];
            tmem_load_row<16>(c0 + H + 16 + 16 * c, b1);
            tmem_load_row<16>(c0 + 2 * H + 16 + 16 * c, b2);
#pragma unroll
            for (int k = 0; k < 16; ++k) b0[k] = (b2[k] + b1[k]) + b0[k];
          }
          if constexpr (kRegAcc) {
#pragma unroll
            for (int k = 0; k < 16; ++k) dw_acc[16 * c + k] += b0[k];
          } else if (live) {
#pragma unroll
            for (int k = 0; k < 16; ++k) atomicAdd(a.dW + (16 * c + k) * P + p, b0[k]);
          }
        }
        float bv[16];
        tmem_load_row<16>(c0 + H, bv);
        if constexpr (kRegAcc) db_acc += bv[0];
        else if (live) atomicAdd(a.dbias + p, bv[0]);
      }
    }
  };

  // Software pipeline over this CTA's tiles.  The tensor pipe runs in issue order (across the CTAs of
  // the SM too) and every tcgen05.mma costs ~46 cycles whatever the shape (tools/umma_probe.cu), so
  // (a) a separate warp does all the issuing and the 128 compute threads never wait for it at a CTA barrier;
  // (b) GEMM 1 of the NEXT tile is issued as soon as this tile's t row has left TMEM, a whole flow sweep
  //     before its result is needed, and GEMM 2 / 3 of this tile complete behind the next tile's flows:
  //     dh / dW are collected one tile late.
  //   compute, tile i: wait GEMM 1(i) -> t row -> split h(i+1) -> arrive bar_h -> flows -> [wait GEMM 2/3(i-1),
  //                    drain] -> split dt(i) -> arrive bar_d
  //   issuer,  tile i: wait bar_h -> GEMM 1(i+1) -> wait bar_d -> GEMM 2/3(i)
  if (tid >= T) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsIssuer));
    if (warp == T / 32) {
      unsigned k = 0;
      mbar_wait(bar_h, 0);
      issue_gemm1<P, H, BWD>(tmem_base + G::cD1, sbase, 0u, bar1);
      for (long long tl = blockIdx.x; tl < ntiles; tl += gridDim.x, ++k) {
        const int buf = (int)(k % G::NA);
        mbar_wait(bar_h, (k + 1) & 1);
        if (tl + gridDim.x < ntiles)
          issue_gemm1<P, H, BWD>(tmem_base + G::cD1, sbase, (unsigned)((k + 1) % G::NA) * G::kA, bar1);
        if constexpr (BWD) {
          mbar_wait(bar_d, k & 1);
          issue_gemm23<P, H>(tmem_base, sbase, (unsigned)buf * G::kA, (k % kFlush) == 0, bar2);
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsCompute));
    fetch_h(tile);
    split_h(0);
    fence_proxy_async();
    mbar_arrive(bar_h);
    fetch_h(tile + gridDim.x);
    fetch_y(tile);

    unsigned it = 0;          // tiles done by this CTA: mbarrier phase parity, h tile in use
    long long r_prev = a.B;   // this thread's row of the previous tile (B: none)
    for (; tile < ntiles; tile += gridDim.x, ++it) {
      float z[D];
#pragma unroll
      for (int i = 0; i < D; ++i) z[i] = y_nxt[i];
      if (a.xf.flags) xform_event<D>(a.xf, tile * T + tid, z);
      const float g_cur = g_nxt;

      // ---- t row out of TMEM (bias included): thread r owns row r of the accumulator
      mbar_wait(bar1, it & 1);
      tc_fence_after();
      float row[PN];
      tmem_load_row<PN>(lane_base + G::cD1, row);

      // ---- next tile's h row (fetched one tile ago) -> its h tile; D1 has been read: GEMM 1(i+1) may go
      split_h((int)((it + 1) % G::NA));
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar_h);
      fetch_h(tile + 2 * (long long)gridDim.x);   // consumed at the top of the next iteration
      fetch_y(tile + gridDim.x);

      // ---- per-row flow chain (registers), dt written in place over t
      const long long r = tile * T + tid;
      if (r < a.B) {
        float zs[Spec::KA][D];
        LogDetAcc<M> ld;
        FwdSweep<Spec, M, V, BWD, 0>::run(row, z, zs, ld);
        using Base = BaseDist<D, Spec::BASE, M>;
        float bth[Base::NA];
        if constexpr (Spec::BASE) Span<0, 2 * D, V>::load(row, bth);
        const float lp = xform_out<M>(a.xf, (BWD ? Base::log_prob_save(bth, z) : Base::log_prob(bth, z)) + ld.nat());
        a.logp[r] = lp;
        {  // compensated fp32 sum (fp64 adds are 1/64 rate here): ls_hi - ls_lo carries ~48 bits
          const float yv = lp - ls_lo, tv = ls_hi + yv;
          ls_lo = (tv - ls_hi) - yv;
          ls_hi = tv;
        }
        if constexpr (BWD) {
          const float cot = a.g_scale * g_cur;
          float Gz[D];
          float gb[Base::NA];
          Base::bwd_saved(bth, z, cot, Gz, gb);
          if constexpr (Spec::BASE) Span<0, 2 * D, V>::store(row, gb);
          BwdSweep<Spec, M, V, Spec::K - 1>::run(row, zs, Gz, cot);
        }
      }

      if constexpr (BWD) {
        // ---- the previous tile's GEMM 2 / 3 were issued before this tile's flows began: normally complete
        if (it > 0) {
          mbar_wait(bar2, (it - 1) & 1);
          tc_fence_after();
          drain_backward(r_prev);
          if (it % kFlush == 0) flush_dw();   // tile it starts a new accumulation window in D3
        }
        r_prev = r;
        // ---- split dt row -> three level tiles (rows past B and the pad columns are zero)
        if (r >= a.B) {
#pragma unroll
          for (int j = 0; j < P; ++j) row[j] = 0.0f;
        }
#pragma unroll
        for (int j = P; j < PN; ++j) row[j] = 0.0f;
#pragma unroll
        for (int c = 0; c < PN / 8; ++c) store_levels8(d_row + c * 128, G::kLvlD, 2 * G::kLvlD, row + 8 * c);
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(bar_d);
      }
    }
    if constexpr (BWD) {
      if (it > 0) {
        mbar_wait(bar2, (it - 1) & 1);
        tc_fence_after();
        drain_backward(r_prev);
        flush_dw();
      }
      if constexpr (kRegAcc) {   // one pass: lane -> chunk tid / 8 of the [level 0 | level 1 | level 2] groups
        const int q = tid >> 3, p = (q % G::CL) * 8 + (tid & 7);
        if (q < 3 * G::CL && p < P) {
#pragma unroll
          for (int k = 0; k < H; ++k) atomicAdd(a.dW + k * P + p, dw_acc[k]);
          atomicAdd(a.dbias + p, db_acc);
        }
      }
    }
  }
  __syncwarp();
  const double lsum = (double)ls_hi - (double)ls_lo;

  if (a.logp_sum) {
    const double sblk = block_sum<NT>(lsum, red);
    if (tid == 0) atomicAdd(a.logp_sum, sblk);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, G::kCols);
}

// ------------------------------------------------------------------ pipe 2: warp-specialised backward
// Per CTA tile k (this CTA's k-th tile), three actors:
//   compute thread r (warpgroup 0, most of the registers):
//       wait GEMM 1(k) -> t row out of D1[k & 1] -> flows forward + reverse -> dt row back INTO D1[k & 1]
//       (tcgen05.st) -> arrive bar_t[k & 1].  Nothing else: no splitting, no shared-memory tile traffic.
//   helper thread r (warpgroup 1), in this order:
//       wait GEMM 2/3(k-1) -> dh row of tile k-1 out of D2 -> global (dW / db every 16 tiles)
//       split the h row of tile k+2 (prefetched one tile earlier still) -> h ring -> arrive bar_hw
//       wait bar_t[k & 1] -> dt row out of D1[k & 1] -> three bf16 levels -> dt tile -> arrive bar_d
//   issuing lane (first warp of warpgroup 1, after its helper duties):
//       wait bar_d -> GEMM 1(k+2) into D1[k & 1], then GEMM 2/3(k)
// Barrier phases cannot alias: bar1 and bar_t alternate between two barriers (a producer would have to be two
// tiles ahead of a consumer, which the chain GEMM 1(k+2) <- bar_d(k) <- bar_t(k) forbids), and the helper
// warpgroup meets at a named barrier once per tile, so no helper warp arrives twice in one phase of bar_hw / bar_d.
template <class Spec, int H, class M, int MINB>
NFN_DEVI void dense_tc5_body2(const DenseArgs& a) {
  constexpr int kRegsHelper = regs_issuer(MINB, true), kRegsCompute = regs_compute(MINB, true);
  static_assert(kRows * (kRegsHelper + kRegsCompute) <= kThreads * regs_launch(MINB), "register pool");
  static_assert(kRegsCompute <= 232 && kRegsCompute % 8 == 0 && kRegsHelper % 8 == 0, "setmaxnreg range");
  constexpr int D = Spec::D;
  constexpr int P = Spec::P();
  using G = Geo<P, H, true>;
  static_assert(G::kPipe2, "two D1 buffers must fit");
  constexpr int PN = G::PN, T = kRows, NT = kThreads;
  constexpr int V = 1;
  constexpr unsigned kGrpA = G::kGrpA, kGrpD = G::kGrpD;
  constexpr unsigned kFlush = 16u;   // tiles per accumulation window of dW / db in tensor memory

  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double red[NT / 32];
  const unsigned sbase = smem_u32(smem_raw);
  // bar1[2]: tensor pipe -> compute (GEMM 1 done);  bar2: tensor pipe -> helpers (GEMM 2 + 3 done);
  // bar_hw: helpers -> issuer (h tile written);  bar_t[2]: compute -> helpers (dt row in D1);
  // bar_d: helpers -> issuer (dt tile written, D1 and D2 free)
  const unsigned bar1 = sbase + G::oBar, bar2 = bar1 + 16, bar_hw = bar1 + 24, bar_t = bar1 + 32, bar_d = bar1 + 48,
                 tmem_slot = bar1 + 64;

  const int tid = threadIdx.x, warp = tid >> 5;
  const long long ntiles = (a.B + T - 1) / T;

  if (tid == 0) {
    mbar_init(bar1, 1);
    mbar_init(bar1 + 8, 1);
    mbar_init(bar2, 1);
    mbar_init(bar_hw, T);
    mbar_init(bar_t, T);
    mbar_init(bar_t + 8, T);
    mbar_init(bar_d, T);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, G::kCols);
  stage_constants<P, H, true>(a, sbase, tid);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  unsigned tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);
  // a warp reaches the TMEM lanes 32 (warp % 4) .. + 31: helper warp 4 + q shares them with compute warp q
  const unsigned lane_base = tmem_base + ((unsigned)((warp & 3) * 32) << 16);
  float ls_hi = 0.0f, ls_lo = 0.0f;

  if (tid >= T) {
    // ================================================================ helpers + issuer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsHelper));
    const int hr = tid - T;                      // this thread's row of every tile
    const bool issuer = (warp == T / 32);
    const unsigned a_row = sbase + G::oA + (hr >> 3) * kGrpA + (hr & 7) * 16;
    const unsigned d_row = sbase + G::oD + (hr >> 3) * kGrpD + (hr & 7) * 16;

    // h row of a tile: prefetched into registers one tile ahead when it is narrow, else loaded chunk by chunk
    constexpr bool kPrefetchH = (H <= 32);
    float h_nxt[kPrefetchH ? H : 1];
    auto fetch_h = [&](long long tile) {
      if constexpr (kPrefetchH) {
        const long long r = tile * T + hr;
#pragma unroll
        for (int i = 0; i < H; ++i) h_nxt[i] = 0.0f;
        if (tile < ntiles && r < a.B) {
          const float4* src = reinterpret_cast<const float4*>(a.h + r * H);
#pragma unroll
          for (int c = 0; c < H / 4; ++c) {
            const float4 v = __ldg(src + c);
            h_nxt[4 * c] = v.x; h_nxt[4 * c + 1] = v.y; h_nxt[4 * c + 2] = v.z; h_nxt[4 * c + 3] = v.w;
          }
        }
      }
    };
    auto split_h = [&](int buf, long long tile) {
      if constexpr (kPrefetchH) {
#pragma unroll
        for (int c = 0; c < H / 8; ++c) store_levels8(a_row + buf * G::kA + c * 128, G::lvlA(1), G::lvlA(2), h_nxt + 8 * c);
      } else {
        const long long r = tile * T + hr;
        const bool live = r < a.B;
        const float4* src = reinterpret_cast<const float4*>(a.h + (live ? r : 0) * H);
#pragma unroll 2
        for (int c = 0; c < H / 8; ++c) {
          float v[8];
          const float4 v0 = live ? __ldg(src + 2 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 v1 = live ? __ldg(src + 2 * c + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[0] = v0.x; v[1] = v0.y; v[2] = v0.z; v[3] = v0.w; v[4] = v1.x; v[5] = v1.y; v[6] = v1.z; v[7] = v1.w;
          store_levels8(a_row + buf * G::kA + c * 128, G::lvlA(1), G::lvlA(2), v);
        }
      }
    };
    // dh row of a finished tile: the three W-level blocks of D2, smallest first -> global
    auto drain_dh = [&](long long r_done) {
#pragma unroll
      for (int c = 0; c < H / 8; ++c) {
        float b0[8], b1[8], b2[8];
        tmem_load8x3(lane_base + G::cD2 + 8 * c, lane_base + G::cD2 + H + 8 * c, lane_base + G::cD2 + 2 * H + 8 * c, b0, b1, b2);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          add2(b2[2 * q], b2[2 * q + 1], b1[2 * q], b1[2 * q + 1], b1[2 * q], b1[2 * q + 1]);
          add2(b1[2 * q], b1[2 * q + 1], b0[2 * q], b0[2 * q + 1], b0[2 * q], b0[2 * q + 1]);
        }
        if (r_done < a.B) {
          st_stream_f4(a.dh + r_done * H + 8 * c, make_float4(b0[0], b0[1], b0[2], b0[3]));
          st_stream_f4(a.dh + r_done * H + 8 * c + 4, make_float4(b0[4], b0[5], b0[6], b0[7]));
        }
      }
    };
    // dW / db lanes of D3 (TMEM lane = (level, p) of pass j) -> atomics, which also sum the levels and the CTAs
    auto flush_dw = [&]() {
#pragma unroll
      for (int j = 0; j < G::NP3; ++j) {
        if (16 * j + 4 * (warp & 3) < 3 * G::CL) {   // this warp's 4 chunks of pass j exist (warp-uniform)
          const int q = 16 * j + (hr >> 3), p = (q % G::CL) * 8 + (hr & 7);
          const bool live = q < 3 * G::CL && p < P;
          const unsigned c0 = lane_base + G::cD3p(j);
#pragma unroll
          for (int c = 0; c < H / 8; ++c) {
            float b0[8];
            if (G::short3(j)) {   // compile-time after unrolling
              tmem_load8(c0 + 8 * c, b0);
            } else {
              float b1[8], b2[8];
              tmem_load8x3(c0 + 8 * c, c0 + H + 16 + 8 * c, c0 + 2 * H + 16 + 8 * c, b0, b1, b2);
#pragma unroll
              for (int k = 0; k < 8; ++k) b0[k] = (b2[k] + b1[k]) + b0[k];
            }
            if (live) {
#pragma unroll
              for (int k = 0; k < 8; ++k) atomicAdd(a.dW + (8 * c + k) * P + p, b0[k]);
            }
          }
          float bv[8];
          tmem_load8(c0 + H, bv);
          if (live) atomicAdd(a.dbias + p, bv[0]);
        }
      }
    };
    // dt row of tile k out of D1[k & 1] -> three level tiles; the load of the next 16 columns is in flight while
    // the current ones are split (this is the one stretch between the end of a tile's flows and its GEMMs)
    auto split_dt = [&](unsigned d1) {
      unsigned r[2][16];
      tmem_ld16(d1, r[0]);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < PN / 16; ++c) {
        if (c + 1 < PN / 16) tmem_ld16(d1 + 16 * (c + 1), r[(c + 1) & 1]);
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          asm volatile("" : "+r"(r[c & 1][j]));
          v[j] = __uint_as_float(r[c & 1][j]);
        }
        store_levels8(d_row + (2 * c) * 128, G::kLvlD, 2 * G::kLvlD, v);
        store_levels8(d_row + (2 * c + 1) * 128, G::kLvlD, 2 * G::kLvlD, v + 8);
        if (c + 1 < PN / 16) tmem_ld_wait();
      }
    };

    // The h rows run TWO tiles ahead of the flows (ring of three tiles) and everything that does not depend on
    // the compute threads -- the previous tile's dh / dW drain, the next-but-one tile's h split -- happens BEFORE
    // the wait for their dt rows, i.e. in the helpers' idle time: after bar_t only the dt split stands between
    // the end of tile k's flows and GEMM 1(k+2).
    long long tl = blockIdx.x;
    const long long step = gridDim.x;
    fetch_h(tl);
    split_h(0, tl);
    fence_proxy_async();
    mbar_arrive(bar_hw);                       // phase 0: h(0)
    fetch_h(tl + step);
    if (issuer) {
      mbar_wait(bar_hw, 0);
      issue_gemm1<P, H, true>(tmem_base + G::cD1, sbase, 0u, bar1);
    }
    helper_sync();                             // phase 0 has been seen complete before anyone arrives for phase 1
    split_h(1, tl + step);
    fence_proxy_async();
    mbar_arrive(bar_hw);                       // phase 1: h(1) (zeros when this CTA has a single tile)
    fetch_h(tl + 2 * step);
    if (issuer) {
      mbar_wait(bar_hw, 1);
      if (tl + step < ntiles) issue_gemm1<P, H, true>(tmem_base + G::cD1 + PN, sbase, G::kA, bar1 + 8);
    }
    unsigned k = 0;
    long long r_prev = a.B;
    for (; tl < ntiles; tl += step, ++k) {
      helper_sync();   // every helper warp has finished tile k-1 (see the phase argument above)
      if (k > 0) {     // GEMM 2 / 3 of tile k-1 were issued a whole flow sweep ago: dh out; h(k-1), the dt tile and D2 free
        mbar_wait(bar2, (k - 1) & 1);
        tc_fence_after();
        drain_dh(r_prev);
        if (k % kFlush == 0) flush_dw();   // tile k starts a new accumulation window in D3
      }
      r_prev = tl * T + hr;
      split_h((int)((k + 2) % G::NA), tl + 2 * step);   // into the tile h(k-1) just left
      fence_proxy_async();
      mbar_arrive(bar_hw);                               // phase k + 2
      fetch_h(tl + 3 * step);
      // ---- the compute threads' dt rows of tile k are in D1[k & 1]
      mbar_wait(bar_t + 8 * (k & 1), (k >> 1) & 1);
      tc_fence_after();
      split_dt(lane_base + G::cD1 + (k & 1) * PN);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar_d);
      if (issuer) {
        mbar_wait(bar_d, k & 1);      // dt tile written; D1[k & 1] and D2 are free
        mbar_wait(bar_hw, k & 1);     // phase k + 2 (complete long ago; waited so that the phases are consumed in order)
        if (tl + 2 * step < ntiles)
          issue_gemm1<P, H, true>(tmem_base + G::cD1 + (k & 1) * PN, sbase, (unsigned)((k + 2) % G::NA) * G::kA,
                                  bar1 + 8 * (k & 1));
        issue_gemm23<P, H>(tmem_base, sbase, (unsigned)(k % G::NA) * G::kA, (k % kFlush) == 0, bar2);
      }
    }
    if (k > 0) {
      mbar_wait(bar2, (k - 1) & 1);
      tc_fence_after();
      drain_dh(r_prev);
      flush_dw();
    }
  } else {
    // ================================================================ compute: the flows, nothing else
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsCompute));
    float y_nxt[D];
    float g_nxt = 1.0f;
    auto fetch_y = [&](long long tile) {
      const long long r = tile * T + tid;
      if (tile < ntiles && r < a.B) {
        load_event<D>(a.y, a.y_broadcast ? 0 : r, y_nxt);
        if (a.g_logp) g_nxt = __ldg(a.g_logp + r);
      }
    };
#pragma unroll
    for (int i = 0; i < D; ++i) y_nxt[i] = 0.0f;
    long long tile = blockIdx.x;
    fetch_y(tile);
    unsigned it = 0;
    for (; tile < ntiles; tile += gridDim.x, ++it) {
      float z[D];
#pragma unroll
      for (int i = 0; i < D; ++i) z[i] = y_nxt[i];
      if (a.xf.flags) xform_event<D>(a.xf, tile * T + tid, z);
      const float g_cur = g_nxt;
      fetch_y(tile + gridDim.x);

      const unsigned d1 = lane_base + G::cD1 + (it & 1) * PN;
      mbar_wait(bar1 + 8 * (it & 1), (it >> 1) & 1);
      tc_fence_after();
      float row[PN];
      tmem_load_row<PN>(d1, row);

      const long long r = tile * T + tid;
      if (r < a.B) {
        float zs[Spec::KA][D];
        LogDetAcc<M> ld;
        FwdSweep<Spec, M, V, true, 0>::run(row, z, zs, ld);
        using Base = BaseDist<D, Spec::BASE, M>;
        float bth[Base::NA];
        if constexpr (Spec::BASE) Span<0, 2 * D, V>::load(row, bth);
        const float lp = xform_out<M>(a.xf, Base::log_prob_save(bth, z) + ld.nat());
        a.logp[r] = lp;
        {
          const float yv = lp - ls_lo, tv = ls_hi + yv;
          ls_lo = (tv - ls_hi) - yv;
          ls_hi = tv;
        }
        const float cot = a.g_scale * g_cur;
        float Gz[D];
        float gb[Base::NA];
        Base::bwd_saved(bth, z, cot, Gz, gb);
        if constexpr (Spec::BASE) Span<0, 2 * D, V>::store(row, gb);
        BwdSweep<Spec, M, V, Spec::K - 1>::run(row, zs, Gz, cot);
      } else {
        // rows past B: h was zero, t = bias; their dt must not reach dW / db
#pragma unroll
        for (int j = 0; j < PN; ++j) row[j] = 0.0f;
      }
      // (the pad columns P .. PN of a live row are exact zeros out of GEMM 1 and no flow touches them)
      tmem_store_row<PN>(d1, row);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_t + 8 * (it & 1));
    }
  }
  __syncwarp();
  const double lsum = (double)ls_hi - (double)ls_lo;

  if (a.logp_sum) {
    const double sblk = block_sum<NT>(lsum, red);
    if (tid == 0) atomicAdd(a.logp_sum, sblk);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, G::kCols);
}

// ------------------------------------------------------------------ entry: the pipeline the shape gets
template <class Spec, int H, bool BWD, class M, int MINB>
NFN_DEVI void dense_tc5_body(const DenseArgs& a) {
  static_assert(Spec::P() > 0 && Spec::P() <= 128, "1..128 parameter columns");
  static_assert(H % 16 == 0 && H >= 16 && H <= 64, "hidden width must be 16, 32, 48 or 64");
  if constexpr (g_pipe2(Spec::P(), H, BWD)) dense_tc5_body2<Spec, H, M, MINB>(a);
  else dense_tc5_body1<Spec, H, BWD, M, MINB>(a);
}

template <class Spec, int H, bool BWD, class M, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) dense_tc5_kernel(const DenseArgs a) {
  dense_tc5_body<Spec, H, BWD, M, MINB>(a);
}

}  // namespace tc5
}  // namespace nfn
