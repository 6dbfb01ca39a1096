// tools/dense_probe.cu -- A/B harness for the fused Dense(P)+chain kernels through the C ABI, no Python:
// runs the warp-level mma.sync version (NFN_B200_DENSE_MMA=sync) and the tcgen05 / TMEM version on the
// same random inputs, compares every output, and times both with CUDA events.
//
//   nvcc -O2 -std=c++17 -o build/dense_probe tools/dense_probe.cu -Lnormalizingflownetwork_b200 -lnfn_b200 \
//        -Xlinker -rpath -Xlinker '$ORIGIN/../normalizingflownetwork_b200'
//   build/dense_probe [rows=1048576] [cfg=2] [reps=20] [hidden=16]
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../include/nfn_b200.h"

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 2;                                                                       \
    }                                                                                 \
  } while (0)

static double max_rel(const std::vector<float>& a, const std::vector<float>& b, double* where = nullptr) {
  double m = 0;
  for (size_t i = 0; i < a.size(); ++i) {
    const double e = std::fabs((double)a[i] - (double)b[i]) / std::max(1.0, std::fabs((double)b[i]));
    if (!(e <= m)) { m = e; if (where) *where = (double)i; }
  }
  return m;
}

int main(int argc, char** argv) {
  const long long B = argc > 1 ? atoll(argv[1]) : (1 << 20);
  const int cfg = argc > 2 ? atoi(argv[2]) : 2;
  const int reps = argc > 3 ? atoi(argv[3]) : 20;
  const int H = argc > 4 ? atoi(argv[4]) : 16;
  nfn_chain_desc desc;
  memset(&desc, 0, sizeof(desc));
  desc.trainable_base = 1;
  if (cfg == 2) {
    desc.n_dims = 2;
    desc.n_flows = 10;
    const uint8_t ft[10] = {0, 1, 2, 0, 1, 2, 0, 1, 2, 0};
    memcpy(desc.flow_type, ft, 10);
  } else if (cfg == 3) {    // BASELINE config 3: 16 flows (radial, planar) x 8, 4-D y (P = 128)
    desc.n_dims = 4;
    desc.n_flows = 16;
    for (int i = 0; i < 16; ++i) desc.flow_type[i] = (i % 2 == 0) ? 1 : 0;
  } else if (cfg == 10) {   // NormalizingFlowNetwork's default chain: 10 radial flows, 1-D y (P = 32)
    desc.n_dims = 1;
    desc.n_flows = 10;
    for (int i = 0; i < 10; ++i) desc.flow_type[i] = 1;
  } else if (cfg == 1) {
    desc.n_dims = 1;
    desc.n_flows = 3;
    desc.flow_type[0] = desc.flow_type[1] = desc.flow_type[2] = 1;
  } else {
    desc.n_dims = 1;
    desc.n_flows = 5;
    for (int i = 0; i < 5; ++i) desc.flow_type[i] = 1;
  }
  const int P = nfn_chain_param_size(&desc), d = desc.n_dims;
  printf("rows %lld  P %d  H %d  d %d\n", B, P, H, d);

  std::mt19937 rng(22);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::vector<float> h(B * H), W(H * P), bias(P), y(B * d);
  for (auto& v : h) v = std::tanh(nd(rng));
  for (auto& v : W) v = 0.5f * nd(rng) / std::sqrt((float)H) * 2.0f;
  for (auto& v : bias) v = 0.1f * nd(rng);
  for (auto& v : y) v = nd(rng);

  float *dh_in, *dW_in, *db_in, *dy_in, *d_logp, *d_dh, *d_dW, *d_db;
  double* d_sum;
  CK(cudaMalloc(&dh_in, B * H * 4));
  CK(cudaMalloc(&dW_in, H * P * 4));
  CK(cudaMalloc(&db_in, P * 4));
  CK(cudaMalloc(&dy_in, B * d * 4));
  CK(cudaMalloc(&d_logp, B * 4));
  CK(cudaMalloc(&d_dh, B * H * 4));
  CK(cudaMalloc(&d_dW, H * P * 4));
  CK(cudaMalloc(&d_db, P * 4));
  CK(cudaMalloc(&d_sum, 8));
  CK(cudaMemcpy(dh_in, h.data(), B * H * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW_in, W.data(), H * P * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db_in, bias.data(), P * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dy_in, y.data(), B * d * 4, cudaMemcpyHostToDevice));

  struct Out {
    std::vector<float> logp, dh, dW, db;
    double sum = 0;
    float ms_fb = 0, ms_f = 0;
  } out[2];
  const char* names[2] = {"mma.sync", "tcgen05"};
  for (int v = 0; v < 2; ++v) {
    setenv("NFN_B200_DENSE_MMA", v == 0 ? "sync" : "tc5", 1);
    auto run = [&]() {
      return nfn_dense_chain_forward_backward(&desc, H, dh_in, dW_in, db_in, dy_in, B, nullptr, -1.0f / (float)B, d_logp,
                                              d_dh, d_dW, d_db, d_sum, B, nullptr);
    };
    CK(cudaMemset(d_dW, 0, H * P * 4));
    CK(cudaMemset(d_db, 0, P * 4));
    CK(cudaMemset(d_sum, 0, 8));
    CK(cudaMemset(d_logp, 0xff, B * 4));
    CK(cudaMemset(d_dh, 0xff, B * H * 4));
    int rc = run();
    if (rc != 0) { printf("%s: rc %d: %s\n", names[v], rc, nfn_last_error()); return 1; }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: kernel failed: %s\n", names[v], cudaGetErrorString(e)); return 1; }
    Out& o = out[v];
    o.logp.resize(B); o.dh.resize(B * H); o.dW.resize(H * P); o.db.resize(P);
    CK(cudaMemcpy(o.logp.data(), d_logp, B * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(o.dh.data(), d_dh, B * H * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(o.dW.data(), d_dW, H * P * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(o.db.data(), d_db, P * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&o.sum, d_sum, 8, cudaMemcpyDeviceToHost));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) run();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) run();
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&o.ms_fb, e0, e1);
    o.ms_fb /= reps;
    auto runf = [&]() { return nfn_dense_chain_forward(&desc, H, dh_in, dW_in, db_in, dy_in, B, d_logp, B, nullptr); };
    for (int i = 0; i < 3; ++i) runf();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) runf();
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&o.ms_f, e0, e1);
    o.ms_f /= reps;
    std::vector<float> lf(B);
    CK(cudaMemcpy(lf.data(), d_logp, B * 4, cudaMemcpyDeviceToHost));
    printf("%-9s fwd+bwd %8.2f us   fwd %8.2f us   sum logp %.9g   fwd-only vs fused logp diff %.2e\n", names[v],
           o.ms_fb * 1e3, o.ms_f * 1e3, o.sum, max_rel(lf, o.logp));
  }
  double w = 0;
  printf("tcgen05 vs mma.sync: logp %.3e", max_rel(out[1].logp, out[0].logp, &w));
  printf(" (row %.0f)", w);
  // gradients carry the -1/B cotangent: compare scaled by B
  auto scaled = [&](std::vector<float> v) { for (auto& x : v) x *= (float)B; return v; };
  printf("  dh*B %.3e", max_rel(scaled(out[1].dh), scaled(out[0].dh), &w));
  printf(" (elem %.0f)", w);
  printf("  dW %.3e", max_rel(out[1].dW, out[0].dW, &w));
  printf(" (elem %.0f)", w);
  printf("  db %.3e\n", max_rel(out[1].db, out[0].db));
  printf("first values: logp %.6f %.6f | dh*B %.6f %.6f | dW %.6f %.6f | db %.6f %.6f\n", out[1].logp[0], out[0].logp[0],
         out[1].dh[0] * B, out[0].dh[0] * B, out[1].dW[0], out[0].dW[0], out[1].db[0], out[0].db[0]);
  return 0;
}
