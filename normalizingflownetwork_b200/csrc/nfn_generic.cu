// nfn_generic.cu -- runtime-chain fallback kernel, single-bijector kernel, column sums.
//
// The generic kernel serves any chain the descriptor can express (K <= 64 flows of any
// mix, d <= 8) when no compile-time specialisation is registered for it.  Same per-flow
// arithmetic (nfn_flows.cuh); the chain structure is a warp-uniform runtime switch, each
// thread walks its own parameter row straight from global memory, and the z history for
// the reverse sweep lives in local memory.  It is the coverage path, not the fast path.
#include "nfn_common.h"

namespace nfn {

struct GenericChain {
  int K;
  int base;
  int P;
  unsigned char type[NFN_MAX_FLOWS];
  short off[NFN_MAX_FLOWS];
};

template <int N>
NFN_DEVI void ld_row(const float* __restrict__ p, float (&th)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) th[i] = __ldg(p + i);
}
template <int N>
NFN_DEVI void st_row(float* __restrict__ p, const float (&g)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = g[i];
}

template <int D, bool BWD, class M>
__global__ void __launch_bounds__(128) chain_generic_kernel(const ChainArgs a, const GenericChain c) {
  __shared__ double red[4];
  double lsum = 0.0;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < a.B;
       r += (long long)gridDim.x * blockDim.x) {
    const float* row = a.t + r * c.P;
    if constexpr (!BWD) {
      if (a.grid_ny > 0) {  // outer-product scoring: B parameter rows x grid_ny events
        for (int j = 0; j < a.grid_ny; ++j) {
          float zg[D];
          load_event<D>(a.y, j, zg);
          if (a.xf.flags) xform_event<D>(a.xf, j, zg, false);
          LogDetAcc<M> ldg;
          for (int k = 0; k < c.K; ++k) {
            const float* p = row + c.off[k];
            switch (c.type[k]) {
              case kPlanar: { float th[2 * D + 1]; ld_row(p, th); PlanarFlow<D, M>::fwd(th, zg, ldg); } break;
              case kRadial: { float th[D + 2]; ld_row(p, th); RadialFlow<D, M>::fwd(th, zg, ldg); } break;
              default: { float th[2 * D]; ld_row(p, th); AffineFlow<D, M>::fwd(th, zg, ldg); } break;
            }
          }
          float lpg;
          if (c.base) {
            float bt[2 * D];
            ld_row(row, bt);
            lpg = BaseDist<D, true, M>::log_prob(bt, zg) + ldg.nat();
          } else {
            float dummy[1] = {0.0f};
            lpg = BaseDist<D, false, M>::log_prob(dummy, zg) + ldg.nat();
          }
          a.logp[(long long)j * a.B + r] = xform_out<M>(a.xf, lpg);
        }
        continue;
      }
    }
    float z[D];
    load_event<D>(a.y, a.y_broadcast ? 0 : r, z);
    if (a.xf.flags) xform_event<D>(a.xf, r, z);
    float zs[BWD ? NFN_MAX_FLOWS * D : 1];
    LogDetAcc<M> ld;
    for (int k = 0; k < c.K; ++k) {
      if constexpr (BWD) {
#pragma unroll
        for (int i = 0; i < D; ++i) zs[k * D + i] = z[i];
      }
      const float* p = row + c.off[k];
      switch (c.type[k]) {
        case kPlanar: {
          float th[2 * D + 1];
          ld_row(p, th);
          PlanarFlow<D, M>::fwd(th, z, ld);
        } break;
        case kRadial: {
          float th[D + 2];
          ld_row(p, th);
          RadialFlow<D, M>::fwd(th, z, ld);
        } break;
        default: {
          float th[2 * D];
          ld_row(p, th);
          AffineFlow<D, M>::fwd(th, z, ld);
        } break;
      }
    }
    float bth[2 * D];
    float lp;
    if (c.base) {
      ld_row(row, bth);
      lp = BaseDist<D, true, M>::log_prob(bth, z) + ld.nat();
    } else {
      float dummy[1] = {0.0f};
      lp = BaseDist<D, false, M>::log_prob(dummy, z) + ld.nat();
    }
    const float lpo = xform_out<M>(a.xf, lp);
    a.logp[r] = lpo;
    lsum += (double)lpo;
    if constexpr (BWD) {
      float* drow = a.dt + r * c.P;
      const float cot = a.g_scale * (a.g_logp ? __ldg(a.g_logp + r) : 1.0f);
      float G[D];
      if (c.base) {
        float gb[2 * D];
        BaseDist<D, true, M>::bwd(bth, z, cot, G, gb);
        st_row(drow, gb);
      } else {
        float dummy[1] = {0.0f}, gd[1];
        BaseDist<D, false, M>::bwd(dummy, z, cot, G, gd);
      }
      for (int k = c.K - 1; k >= 0; --k) {
        float zin[D];
#pragma unroll
        for (int i = 0; i < D; ++i) zin[i] = zs[k * D + i];
        const float* p = row + c.off[k];
        float* q = drow + c.off[k];
        switch (c.type[k]) {
          case kPlanar: {
            float th[2 * D + 1], g[2 * D + 1];
            ld_row(p, th);
            PlanarFlow<D, M>::bwd(th, zin, G, cot, g);
            st_row(q, g);
          } break;
          case kRadial: {
            float th[D + 2], g[D + 2];
            ld_row(p, th);
            RadialFlow<D, M>::bwd(th, zin, G, cot, g);
            st_row(q, g);
          } break;
          default: {
            float th[2 * D], g[2 * D];
            ld_row(p, th);
            AffineFlow<D, M>::bwd(th, zin, G, cot, g);
            st_row(q, g);
          } break;
        }
      }
      if (a.dy) store_event<D>(a.dy, r, G);
    }
  }
  if (a.logp_sum) {
    const double s = block_sum<128>(lsum, red);
    if (threadIdx.x == 0) atomicAdd(a.logp_sum, s);
  }
}

template <int D, bool BWD, class M>
static cudaError_t launch_generic_t(const ChainArgs& a, const GenericChain& c, cudaStream_t st) {
  const DeviceInfo& di = device_info();
  long long blocks = (a.B + 127) / 128;
  const long long cap = (long long)di.sm_count * 16;
  if (blocks > cap) blocks = cap;
  chain_generic_kernel<D, BWD, M><<<(unsigned)blocks, 128, 0, st>>>(a, c);
  count_launch();
  return cudaGetLastError();
}

template <int D>
static cudaError_t launch_generic_d(const ChainArgs& a, const GenericChain& c, bool bwd, int mode,
                                    cudaStream_t st) {
  if (mode == 0) {
    return bwd ? launch_generic_t<D, true, MathFast>(a, c, st)
               : launch_generic_t<D, false, MathFast>(a, c, st);
  }
  return bwd ? launch_generic_t<D, true, MathAccurate>(a, c, st)
             : launch_generic_t<D, false, MathAccurate>(a, c, st);
}

cudaError_t launch_chain_generic(const nfn_chain_desc* desc, const ChainArgs& a, bool bwd, int mode,
                                 cudaStream_t st) {
  GenericChain c;
  const int d = desc->n_dims;
  c.K = desc->n_flows;
  c.base = desc->trainable_base ? 1 : 0;
  int off = c.base ? 2 * d : 0;
  for (int k = c.K - 1; k >= 0; --k) {  // last flow owns the first columns
    c.type[k] = desc->flow_type[k];
    c.off[k] = (short)off;
    off += flow_param_size(desc->flow_type[k], d);
  }
  c.P = off;
  switch (d) {
    case 1: return launch_generic_d<1>(a, c, bwd, mode, st);
    case 2: return launch_generic_d<2>(a, c, bwd, mode, st);
    case 3: return launch_generic_d<3>(a, c, bwd, mode, st);
    case 4: return launch_generic_d<4>(a, c, bwd, mode, st);
    case 5: return launch_generic_d<5>(a, c, bwd, mode, st);
    case 6: return launch_generic_d<6>(a, c, bwd, mode, st);
    case 7: return launch_generic_d<7>(a, c, bwd, mode, st);
    case 8: return launch_generic_d<8>(a, c, bwd, mode, st);
  }
  return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------ single bijector
template <int TYPE, int D>
__global__ void __launch_bounds__(128) flow_single_kernel(const float* __restrict__ t,
                                                          const float* __restrict__ z_in, int z_bcast,
                                                          float* __restrict__ z_out,
                                                          float* __restrict__ fldj, long long B) {
  using M = MathAccurate;
  constexpr int N = flow_param_size(TYPE, D);
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < B;
       r += (long long)gridDim.x * blockDim.x) {
    float th[N];
    ld_row(t + r * N, th);
    float z[D];
    load_event<D>(z_in, z_bcast ? 0 : r, z);
    LogDetAcc<M> ld;
    if constexpr (TYPE == kPlanar) PlanarFlow<D, M>::fwd(th, z, ld);
    else if constexpr (TYPE == kRadial) RadialFlow<D, M>::fwd(th, z, ld);
    else AffineFlow<D, M>::fwd(th, z, ld);
    if (z_out) store_event<D>(z_out, r, z);
    if (fldj) fldj[r] = ld.nat();
  }
}

template <int TYPE, int D>
static cudaError_t launch_flow_td(const float* t, const float* z, int zb, float* zo, float* f,
                                  long long B, cudaStream_t st) {
  long long blocks = (B + 127) / 128;
  const long long cap = (long long)device_info().sm_count * 16;
  if (blocks > cap) blocks = cap;
  flow_single_kernel<TYPE, D><<<(unsigned)blocks, 128, 0, st>>>(t, z, zb, zo, f, B);
  count_launch();
  return cudaGetLastError();
}

template <int TYPE>
static cudaError_t launch_flow_t(int d, const float* t, const float* z, int zb, float* zo, float* f,
                                 long long B, cudaStream_t st) {
  switch (d) {
    case 1: return launch_flow_td<TYPE, 1>(t, z, zb, zo, f, B, st);
    case 2: return launch_flow_td<TYPE, 2>(t, z, zb, zo, f, B, st);
    case 3: return launch_flow_td<TYPE, 3>(t, z, zb, zo, f, B, st);
    case 4: return launch_flow_td<TYPE, 4>(t, z, zb, zo, f, B, st);
    case 5: return launch_flow_td<TYPE, 5>(t, z, zb, zo, f, B, st);
    case 6: return launch_flow_td<TYPE, 6>(t, z, zb, zo, f, B, st);
    case 7: return launch_flow_td<TYPE, 7>(t, z, zb, zo, f, B, st);
    case 8: return launch_flow_td<TYPE, 8>(t, z, zb, zo, f, B, st);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_flow_single(int type, int d, const float* t, const float* z, int zb, float* zo,
                               float* f, long long B, cudaStream_t st) {
  switch (type) {
    case kPlanar: return launch_flow_t<kPlanar>(d, t, z, zb, zo, f, B, st);
    case kRadial: return launch_flow_t<kRadial>(d, t, z, zb, zo, f, B, st);
    case kAffine: return launch_flow_t<kAffine>(d, t, z, zb, zo, f, B, st);
  }
  return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------ column sums of dt
// out[j] += sum_b dt[b, j]; used by the generic chain path and the mixture heads.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dt, long long B, int P,
                                                     double* __restrict__ out) {
  extern __shared__ float acc[];
  for (int j = threadIdx.x; j < P; j += blockDim.x) acc[j] = 0.0f;
  __syncthreads();
  const long long total = B * (long long)P;
  const long long per = ((total + gridDim.x - 1) / gridDim.x + 255) / 256 * 256;
  const long long lo = per * blockIdx.x;
  long long hi = lo + per;
  if (hi > total) hi = total;
  for (long long e = lo + threadIdx.x; e < hi; e += blockDim.x) atomicAdd(&acc[e % P], __ldg(dt + e));
  __syncthreads();
  for (int j = threadIdx.x; j < P; j += blockDim.x) atomicAdd(out + j, (double)acc[j]);
}

int launch_colsum(const float* dt, long long B, int P, double* out, cudaStream_t st) {
  if (B <= 0 || P <= 0) return NFN_OK;
  long long blocks = (B * (long long)P + 65535) / 65536;
  const long long cap = (long long)device_info().sm_count * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  colsum_kernel<<<(unsigned)blocks, 256, P * sizeof(float), st>>>(dt, B, P, out);
  count_launch();
  return cuda_error(cudaGetLastError(), "colsum_kernel");
}

}  // namespace nfn
