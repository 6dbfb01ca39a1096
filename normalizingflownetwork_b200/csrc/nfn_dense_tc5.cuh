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
// Roles: the 128 threads of warpgroup 0 own one row each (t row out of TMEM, flows, dt split, dh drain); in
// warpgroup 1 the first warp's elected lane issues every MMA, and the other three warps stage the h tiles
// (global -> three bf16 levels -> smem ring, one tile ahead), so that no compute thread touches h at all.
// (A second pipeline, in which warpgroup 1 also split the dt rows handed over through tensor memory, was built
// and measured slower -- the hand-over chain dt -> split -> GEMM 2/3 -> drain serialises behind the issuing
// warp; profiles/tuning_r02.md section 6, commit "experimental warp-specialised backward pipeline".)
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

namespace nfn {
namespace tc5 {

constexpr int kRows = 128;     // rows per tile == compute threads per CTA == TMEM lanes
// + a second warpgroup: its first warp issues every tcgen05.mma, its other three warps stage the h tiles.
// A whole warpgroup because registers are handed out per warpgroup (setmaxnreg).
constexpr int kThreads = 256;
constexpr int kStagers = 96;    // threads of warps 5..7, which stage the h tiles
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
__host__ __device__ constexpr int g_cD2(int P, int H, bool bwd) { return round16(P); }
__host__ __device__ constexpr int g_cD3(int P, int H, bool bwd) { return g_cD2(P, H, bwd) + 3 * H; }
__host__ __device__ constexpr unsigned tmem_cols(int P, int H, bool bwd) {
  return pow2_cols(bwd ? g_cD3(P, H, bwd) + g_cols3(P, H, g_NP3(P)) : round16(P));
}
// Register budget for MINB resident CTAs: the launch-time allocation is 65536 / (MINB * 256) per thread
// (rounded down to 8); the second warpgroup keeps regs_issuer of it and the compute warpgroup gets the rest.
// The sum must never exceed the CTA's pool, or setmaxnreg.inc would wait forever.
__host__ __device__ constexpr int regs_launch(int minb) { return 65536 / (minb * kThreads) / 8 * 8; }
__host__ __device__ constexpr int regs_issuer(int minb) { return 56; }   // issuing warp + the three h-staging warps (rows prefetched in registers)
__host__ __device__ constexpr int regs_compute(int minb) { return 2 * regs_launch(minb) - regs_issuer(minb); }
// resident CTAs per SM the kernel's register plan is built for (shared memory and TMEM columns permitting)
__host__ __device__ constexpr int min_blocks(int P, int H, bool bwd) {
  const int by_smem = (int)((227u * 1024u) / (smem_bytes(P, H, bwd) + 1024u));
  const int by_tmem = (int)(512u / tmem_cols(P, H, bwd));
  const int want = bwd ? 2 : 3;   // compute warpgroup registers: 200 (fwd+bwd), 104 (forward)
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
  // TMEM columns: D1 [PN] | D2 [3H: dt W0^T | dt W1^T | dt W2^T] | D3 [passes: dt^T h0 | db .. | dt^T h1 | dt^T h2]
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
// 16 consecutive columns of this thread's TMEM lane
NFN_DEVI void tmem_ld16(unsigned taddr, unsigned (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
NFN_DEVI void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
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

// ------------------------------------------------------------------ the fused body
template <class Spec, int H, bool BWD, class M, int MINB>
NFN_DEVI void dense_tc5_body(const DenseArgs& a) {
  constexpr int kRegsIssuer = regs_issuer(MINB), kRegsCompute = regs_compute(MINB);
  static_assert(kRows * (kRegsIssuer + kRegsCompute) <= kThreads * regs_launch(MINB), "register pool");
  static_assert(kRegsCompute <= 232 && kRegsCompute % 8 == 0 && kRegsIssuer % 8 == 0, "setmaxnreg range");
  constexpr int D = Spec::D;
  constexpr int P = Spec::P();
  static_assert(P > 0 && P <= 128, "1..128 parameter columns");
  static_assert(H % 16 == 0 && H >= 16 && H <= 64, "hidden width must be 16, 32, 48 or 64");
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
  // bar1 / bar2: tensor pipe -> compute threads and stagers (GEMM 1 done / GEMM 2+3 done);
  // bar_h / bar_d: compute threads -> issuing thread (D1 consumed / dt tiles written and D2, D3 drained);
  // bar_hw: stagers -> issuing thread (h tile written)
  const unsigned bar1 = sbase + G::oBar, bar2 = bar1 + 8, bar_h = bar1 + 16, bar_d = bar1 + 24, bar_hw = bar1 + 32,
                 tmem_slot = bar1 + 64;

  const int tid = threadIdx.x, warp = tid >> 5;
  const long long ntiles = (a.B + T - 1) / T;

  // ---- one-time set-up: barriers, TMEM, split weight tiles, bias operand, ones blocks
  if (tid == 0) {
    mbar_init(bar1, 1);
    mbar_init(bar2, 1);
    mbar_init(bar_h, T);
    mbar_init(bar_d, T);
    mbar_init(bar_hw, kStagers);
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

  // this thread's row of the next tile: y, upstream cotangent
  float y_nxt[D];
  float g_nxt = 1.0f;
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
  //   compute, tile i: wait GEMM 1(i) -> t row -> arrive bar_h -> flows -> [wait GEMM 2/3(i-1), drain] ->
  //                    split dt(i) -> arrive bar_d
  //   stagers, tile i: (loads of h(i) in flight) wait GEMM 1(i-1) -> split h(i) -> ring slot i % NA -> arrive bar_hw
  //   issuer,  tile i: wait bar_h, bar_hw -> GEMM 1(i+1) -> wait bar_d -> GEMM 2/3(i)
  if (tid >= T) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsIssuer));
    if (warp == T / 32) {
      unsigned k = 0;
      mbar_wait(bar_h, 0);
      mbar_wait(bar_hw, 0);
      issue_gemm1<P, H, BWD>(tmem_base + G::cD1, sbase, 0u, bar1);
      for (long long tl = blockIdx.x; tl < ntiles; tl += gridDim.x, ++k) {
        const int buf = (int)(k % G::NA);
        mbar_wait(bar_h, (k + 1) & 1);
        if (tl + gridDim.x < ntiles) {
          mbar_wait(bar_hw, (k + 1) & 1);
          issue_gemm1<P, H, BWD>(tmem_base + G::cD1, sbase, (unsigned)((k + 1) % G::NA) * G::kA, bar1);
        }
        if constexpr (BWD) {
          mbar_wait(bar_d, k & 1);
          issue_gemm23<P, H>(tmem_base, sbase, (unsigned)buf * G::kA, (k % kFlush) == 0, bar2);
        }
      }
    } else {
      // ---- stagers (warps 5..7): the h tile of every tile of this CTA, global -> three bf16 levels -> ring slot.
      // Work item = (8-column chunk c, row r), chunk-major, so that 8 consecutive threads write 8 consecutive
      // rows of one chunk: 128 contiguous bytes per level, conflict-free; each reads 32 bytes of its row.
      // Ring slot p % NA was last read by GEMM 1 / GEMM 3 of tile p - NA, all issued before GEMM 1(p-1): the
      // commit behind GEMM 1(p-1) (bar1) covers every earlier MMA of the issuing thread, so one wait frees the slot
      // -- and keeps this loop exactly one tile ahead of the flows, which is also what orders the phases of bar_hw.
      constexpr int kItems = kRows * (H / 8);
      constexpr int kPer = (kItems + kStagers - 1) / kStagers;   // items per thread: 3 (H = 16) .. 11 (H = 64)
      constexpr bool kPrefetch = kPer <= 3;                       // rows held in registers across the wait
      const int ht = tid - (kThreads - kStagers);
      unsigned p = 0;
      for (long long tl = blockIdx.x; tl < ntiles; tl += gridDim.x, ++p) {
        const unsigned slot = sbase + G::oA + (p % G::NA) * G::kA;
        float v[kPrefetch ? kPer : 1][8];
        auto load_item = [&](int item, float (&o)[8]) {
          const long long r = tl * T + (item & (kRows - 1));
          if (item < kItems && r < a.B) {
            const float4* src = reinterpret_cast<const float4*>(a.h + r * H + 8 * (item >> 7));
            const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
            o[0] = v0.x; o[1] = v0.y; o[2] = v0.z; o[3] = v0.w; o[4] = v1.x; o[5] = v1.y; o[6] = v1.z; o[7] = v1.w;
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = 0.0f;   // rows past B stage zeros: they add nothing to dW / db
          }
        };
        auto store_item = [&](int item, const float (&o)[8]) {
          if (item < kItems) {
            const int r = item & (kRows - 1), c = item >> 7;
            store_levels8(slot + (r >> 3) * kGrpA + c * 128 + (r & 7) * 16, G::lvlA(1), G::lvlA(2), o);
          }
        };
        if constexpr (kPrefetch) {
#pragma unroll
          for (int i = 0; i < kPer; ++i) load_item(ht + i * kStagers, v[i]);
        }
        if (p > 0) mbar_wait(bar1, (p - 1) & 1);
        if constexpr (kPrefetch) {
#pragma unroll
          for (int i = 0; i < kPer; ++i) store_item(ht + i * kStagers, v[i]);
        } else {
#pragma unroll 1
          for (int i = 0; i < kPer; ++i) {
            load_item(ht + i * kStagers, v[0]);
            store_item(ht + i * kStagers, v[0]);
          }
        }
        fence_proxy_async();
        mbar_arrive(bar_hw);
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsCompute));
    mbar_arrive(bar_h);   // phase 0: D1 is free
    fetch_y(tile);

    unsigned it = 0;          // tiles done by this CTA: mbarrier phase parity
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

      // ---- D1 has been read: GEMM 1(i+1) may go
      tc_fence_before();
      mbar_arrive(bar_h);
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

template <class Spec, int H, bool BWD, class M, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) dense_tc5_kernel(const DenseArgs a) {
  dense_tc5_body<Spec, H, BWD, M, MINB>(a);
}

}  // namespace tc5
}  // namespace nfn
