// tools/umma_probe.cu -- descriptor semantics of tcgen05.mma kind::tf32 with MN-major operands in the
// no-swizzle layout, measured on the device (development aid for csrc/nfn_dense_tc5.cuh).
//
// X[128][CA] and Y[128][CB] are written "K-major" (8-row x 16-byte core matrices).  The probe asks the
// tensor core for D[m][n] = sum_k X[k][m] * Y[k][n] (both operands read MN-major, K = 128 in 16 steps)
// and compares with the host.  argv: variant (bit 0: swap LBO/SBO of A, bit 1: swap for B,
// bit 2: B K-major "ones" style instead), M (64|128).
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/umma_probe tools/umma_probe.cu
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int CA = 48, CB = 16;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
  return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t instr_desc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ constexpr unsigned kmajor_off(int row, int col, int C) {
  return (unsigned)((row >> 3) * (C / 4 * 128) + (col >> 2) * 128 + (row & 7) * 16 + (col & 3) * 4);
}

__global__ void __launch_bounds__(128) probe(const float* X, const float* Y, float* D, int variant, int M, unsigned* status) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ unsigned tslot;
  unsigned char* sX = smem;                     // 128*CA*4 = 24576
  unsigned char* sY = smem + 128 * CA * 4 + 8192;  // slack after X for over-reads
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * CA; i += 128) *reinterpret_cast<float*>(sX + kmajor_off(i / CA, i % CA, CA)) = X[i];
  if (variant & 4) {  // B = constant K-major [N = 16][K = 8] tile whose row n = 0 is all ones: D[m][0] = sum_k X[k][m]
    for (int i = tid; i < 128; i += 128) *reinterpret_cast<float*>(sY + kmajor_off(i / 8, i % 8, 8)) = (i / 8 == 0) ? 1.0f : 0.0f;
  } else {
    for (int i = tid; i < 128 * CB; i += 128) *reinterpret_cast<float*>(sY + kmajor_off(i / CB, i % CB, CB)) = Y[i];
  }
  for (int i = tid; i < 2048; i += 128) reinterpret_cast<float*>(sX + 128 * CA * 4)[i] = 0.0f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tb = tslot;
  // sentinel in D so a dropped MMA is visible
  {
    const unsigned taddr = tb + ((unsigned)(warp * 32) << 16);
    const unsigned s = __float_as_uint(-777.0f);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(s) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned gA = CA / 4 * 128, gB = CB / 4 * 128;  // stride between 8-row groups
    const bool b_ones = variant & 4;
    const uint32_t idesc = instr_desc(M, 16, 1, b_ones ? 0 : 1);
    for (int ks = 0; ks < 16; ++ks) {
      const unsigned a_addr = smem_u32(sX) + ks * gA, b_addr = smem_u32(sY) + ks * gB;
      const uint64_t ad = (variant & 1) ? smem_desc(a_addr, 128, gA) : smem_desc(a_addr, gA, 128);
      uint64_t bd = (variant & 2) ? smem_desc(b_addr, 128, gB) : smem_desc(b_addr, gB, 128);
      if (b_ones) bd = smem_desc(smem_u32(sY), 128, 256);
      const unsigned acc = ks > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb), "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // bounded wait
  {
    unsigned ok = 0;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 1000000000ull) { if (tid == 0) *status = 1; break; }
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    unsigned r[16];
    const unsigned taddr = tb + ((unsigned)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) D[tid * 16 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(32u) : "memory");
}

// ---- bf16 variant: same experiment with 16-bit operands (K = 16 per instruction, 8 elements per 16-byte chunk)
__host__ __device__ constexpr unsigned kmajor_off16(int row, int col, int C) {
  return (unsigned)((row >> 3) * (C / 8 * 128) + (col >> 3) * 128 + (row & 7) * 16 + (col & 7) * 2);
}
__host__ __device__ constexpr uint32_t instr_desc16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__global__ void __launch_bounds__(128) probe16(const float* X, const float* Y, float* D, int variant, int M, unsigned* status) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ unsigned tslot;
  unsigned char* sX = smem;
  unsigned char* sY = smem + 128 * CA * 2 + 8192;
  const int tid = threadIdx.x, warp = tid >> 5;
  auto bf = [](float v) { return (unsigned short)(__float_as_uint(v) >> 16); };
  for (int i = tid; i < 128 * CA; i += 128) *reinterpret_cast<unsigned short*>(sX + kmajor_off16(i / CA, i % CA, CA)) = bf(X[i]);
  if (variant & 4) {
    for (int i = tid; i < 256; i += 128) *reinterpret_cast<unsigned short*>(sY + kmajor_off16(i / 16, i % 16, 16)) = (i / 16 == 0) ? bf(1.0f) : bf(0.0f);
  } else {
    for (int i = tid; i < 128 * CB; i += 128) *reinterpret_cast<unsigned short*>(sY + kmajor_off16(i / CB, i % CB, CB)) = bf(Y[i]);
  }
  for (int i = tid; i < 2048; i += 128) reinterpret_cast<float*>(sX + 128 * CA * 2)[i] = 0.0f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tb = tslot;
  {
    const unsigned taddr = tb + ((unsigned)(warp * 32) << 16);
    const unsigned s = __float_as_uint(-777.0f);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(s) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned gA = CA / 8 * 128, gB = CB / 8 * 128;  // stride between 8-row groups
    const bool b_ones = variant & 4;
    const uint32_t idesc = instr_desc16(M, 16, 1, b_ones ? 0 : 1);
    for (int ks = 0; ks < 8; ++ks) {  // 16 rows = two 8-row groups per step
      const unsigned a_addr = smem_u32(sX) + ks * 2 * gA, b_addr = smem_u32(sY) + ks * 2 * gB;
      const uint64_t ad = (variant & 1) ? smem_desc(a_addr, 128, gA) : smem_desc(a_addr, gA, 128);
      uint64_t bd = (variant & 2) ? smem_desc(b_addr, 128, gB) : smem_desc(b_addr, gB, 128);
      if (b_ones) bd = smem_desc(smem_u32(sY), 128, 256);
      const unsigned acc = ks > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb), "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    unsigned ok = 0;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 1000000000ull) { if (tid == 0) *status = 1; break; }
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    unsigned r[16];
    const unsigned taddr = tb + ((unsigned)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) D[tid * 16 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(32u) : "memory");
}

// ---- issue-rate microbenchmark: `n` back-to-back bf16 MMAs of one shape from one thread, cycles per MMA
__global__ void __launch_bounds__(128) bench16(int M, int N, int a_mn, int b_mn, int n, long long* cycles) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ unsigned tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 16384; i += 128) reinterpret_cast<float*>(smem)[i] = 0.0f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tb = tslot;
  if (tid == 0) {
    const uint32_t idesc = instr_desc16(M, N, a_mn, b_mn);
    const uint64_t ad = a_mn ? smem_desc(smem_u32(smem), 768, 128) : smem_desc(smem_u32(smem), 128, 768);
    const uint64_t bd = b_mn ? smem_desc(smem_u32(smem) + 32768, 256, 128) : smem_desc(smem_u32(smem) + 32768, 128, 256);
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb + (unsigned)((i & 1) * 64)), "l"(ad), "l"(bd), "r"(idesc), "r"(1u)
          : "memory");
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    unsigned ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    const long long t2 = clock64();
    cycles[0] = t1 - t0;
    cycles[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256u) : "memory");
}

static int run_bench() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench16, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const int shapes[][4] = {{128, 16, 0, 0}, {128, 48, 0, 0}, {128, 16, 1, 1}, {64, 16, 1, 1}, {64, 16, 0, 0}, {128, 32, 1, 1},
                           {64, 32, 1, 1}, {128, 64, 0, 0}, {128, 128, 0, 0}, {128, 256, 0, 0}, {64, 48, 1, 1}, {64, 64, 1, 1}};
  for (auto& sh : shapes) {
    for (int n : {64, 512}) {
      bench16<<<1, 128, 65536>>>(sh[0], sh[1], sh[2], sh[3], n, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long c[2] = {0, 0};
      cudaMemcpy(c, d, 16, cudaMemcpyDeviceToHost);
      printf("M %3d N %3d a_mn %d b_mn %d  n %4d: issue %7.1f cyc/mma, complete %7.1f cyc/mma  (%s)\n", sh[0], sh[1], sh[2], sh[3], n,
             (double)c[0] / n, (double)c[1] / n, cudaGetErrorString(e));
    }
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && argv[1][0] == 'b') return run_bench();
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int M = argc > 2 ? atoi(argv[2]) : 128;
  const int bf16 = argc > 3 ? atoi(argv[3]) : 0;
  std::vector<float> X(128 * CA), Y(128 * CB), D(128 * 16);
  // tf32-exact small integers: products and sums are exact
  for (int k = 0; k < 128; ++k) {
    for (int m = 0; m < CA; ++m) X[k * CA + m] = (float)((k * 7 + m * 3) % 11 - 5);
    for (int n = 0; n < CB; ++n) Y[k * CB + n] = (float)((k * 5 + n * 2) % 7 - 3);
  }
  float *dX, *dY, *dD;
  unsigned* dS;
  cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dY, Y.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dS, 4);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dY, Y.data(), Y.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dS, 0, 4);
  const int smem = 128 * CA * 4 + 8192 + 128 * CB * 4 + 8192;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe16, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (bf16) probe16<<<1, 128, smem>>>(dX, dY, dD, variant, M, dS);
  else probe<<<1, 128, smem>>>(dX, dY, dD, variant, M, dS);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned st = 0;
  cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
  printf("%s variant %d M %d: %s, timeout %u\n", bf16 ? "bf16" : "tf32", variant, M, cudaGetErrorString(e), st);
  if (e != cudaSuccess) return 1;
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  // reference (MN-major both)
  double maxerr = 0;
  int bad = 0, sentinel = 0;
  for (int m = 0; m < CA; ++m)
    for (int n = 0; n < CB; ++n) {
      double ref = 0;
      for (int k = 0; k < 128; ++k) ref += (double)X[k * CA + m] * ((variant & 4) ? (n == 0 ? 1.0 : 0.0) : (double)Y[k * CB + n]);
      const double got = D[m * 16 + n];
      if (got == -777.0) ++sentinel;
      const double err = std::fabs(got - ref);
      if (err > maxerr) maxerr = err;
      if (err > 1e-3) ++bad;
    }
  printf("  lanes 0..%d x 16 cols: max err %.3f, %d bad, %d sentinel\n", CA - 1, maxerr, bad, sentinel);
  for (int m = 0; m < 4; ++m) {
    printf("  lane %d:", m);
    for (int n = 0; n < 8; ++n) printf(" %8.1f", D[m * 16 + n]);
    printf("   ref:");
    for (int n = 0; n < 4; ++n) {
      double ref = 0;
      for (int k = 0; k < 128; ++k) ref += (double)X[k * CA + m] * ((variant & 4) ? (n == 0 ? 1.0 : 0.0) : (double)Y[k * CB + n]);
      printf(" %8.1f", ref);
    }
    printf("\n");
  }
  printf("  lane 16: %8.1f %8.1f   lane 32: %8.1f %8.1f  lane 64: %8.1f %8.1f\n", D[16 * 16], D[16 * 16 + 1], D[32 * 16], D[32 * 16 + 1],
         D[64 * 16], D[64 * 16 + 1]);
  return 0;
}
