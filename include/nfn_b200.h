/*
 * nfn_b200.h -- C ABI of libnfn_b200.so: the B200 (sm_100a) flow-chain / mixture-head
 * log-likelihood hot path of siboehm/NormalizingFlowNetwork.
 *
 * The reference has no native layer at all (pure Python on TensorFlow/TFP), so there is
 * no existing FFI to mirror; each entry point below names the reference interface
 * (path:line under the reference tree) whose arithmetic it replaces.  Python binds these
 * with ctypes (normalizingflownetwork_b200/_lib.py); see INTEGRATION.md for the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns NFN_OK (0) or a negative nfn_status; it never throws and
 *     never terminates the process.  nfn_last_error() returns a thread-local message.
 *   - "device" pointers are CUDA device pointers owned by the caller; fp32, row-major,
 *     dense (row stride == row width).  t/dt base pointers must be 16-byte aligned
 *     (NFN_ERR_ALIGN otherwise).  The library allocates no persistent device memory in
 *     the device entry points, never synchronises the host, and is CUDA-graph
 *     capturable.  `stream` is a cudaStream_t passed as void* (NULL = legacy default).
 *   - the *_host entry points take HOST pointers (pinned memory recommended), run a
 *     chunked H2D -> kernel -> D2H pipeline on library-owned streams/staging buffers
 *     (per calling thread, released by nfn_host_release()), and return after the
 *     results are in the host buffers.
 *   - one process (or thread) per GPU; the current CUDA device of the calling thread is
 *     used.  No global mutable state apart from the thread-local error string, the
 *     read-only kernel registry and the per-thread host-pipeline workspace.
 */
#ifndef NFN_B200_H
#define NFN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NFN_B200_VERSION 100 /* 0.1.0 */

#define NFN_MAX_FLOWS 64
#define NFN_MAX_DIMS 8

typedef enum nfn_status {
  NFN_OK = 0,
  NFN_ERR_NULL = -1,        /* required pointer is NULL */
  NFN_ERR_DESC = -2,        /* malformed descriptor (n_dims, n_flows, flow type code) */
  NFN_ERR_SHAPE = -3,       /* bad B / y_rows / width */
  NFN_ERR_ALIGN = -4,       /* t or dt not 16-byte aligned */
  NFN_ERR_CUDA = -5,        /* CUDA runtime error, see nfn_last_error() */
  NFN_ERR_UNSUPPORTED = -6, /* valid request this build cannot serve */
  NFN_ERR_PEER_TIMEOUT = -7 /* a rank did not arrive at a peer exchange in time (nfn_peer_status) */
} nfn_status;

/* Flow type codes: the keys of FLOWS in estimators/normalizing_flows/__init__.py:5 */
enum { NFN_FLOW_PLANAR = 0, NFN_FLOW_RADIAL = 1, NFN_FLOW_AFFINE = 2 };

/*
 * Chain descriptor == the constructor arguments of InverseNormalizingFlowLayer
 * (estimators/DistributionLayers.py:220-243).  flow_type[] is in `flow_types` order:
 * data y passes through flow_type[0] first.  The parameter-row layout is the
 * reference's (DistributionLayers.py:252, :267-278, :283-288):
 *   [ mu(d) | sigma_raw(d) ]  (only if trainable_base)   then
 *   [ theta(flow K-1) | theta(flow K-2) | ... | theta(flow 0) ]   (REVERSED order)
 * with theta = planar [u(d), w_raw(d), b], radial [alpha_raw, beta_raw, gamma(d)],
 * affine [shift(d), scale_raw(d)].
 */
typedef struct nfn_chain_desc {
  int32_t n_dims;         /* d, 1..NFN_MAX_DIMS */
  int32_t n_flows;        /* K, 0..NFN_MAX_FLOWS */
  int32_t trainable_base; /* 0 | 1 */
  uint8_t flow_type[NFN_MAX_FLOWS];
} nfn_chain_desc;

/* InverseNormalizingFlowLayer.get_total_param_size (DistributionLayers.py:257-265).
 * Returns P > 0 or a negative nfn_status. */
int nfn_chain_param_size(const nfn_chain_desc* desc);

/* 1 if an ahead-of-time specialised kernel is built in for this chain, 0 otherwise (the chain
 * is then specialised at first use by the runtime compiler -- NVRTC, cached in memory and under
 * $NFN_B200_CACHE or ~/.cache/nfn_b200 -- or, if NVRTC is unavailable / NFN_B200_JIT=0 / the
 * chain is too long for registers, served by the generic runtime-chain kernel).  Negative
 * nfn_status on a bad descriptor. */
int nfn_chain_is_specialized(const nfn_chain_desc* desc);

/* Number of chains the runtime specialiser has compiled or loaded in this process. */
int nfn_jit_cache_size(void);

/* Compile (but do not load) the runtime-specialised kernels of a chain: needs NVRTC but no GPU.
 * Returns the cubin size in bytes, or a negative nfn_status with the compiler log in
 * nfn_last_error(). */
int64_t nfn_jit_compile_check(const nfn_chain_desc* desc, int accurate);

/*
 * log_prob of TransformedDistribution(base, Invert(Chain(flows))) for B rows:
 * replaces dist.log_prob(y) built by InverseNormalizingFlowLayer._get_distribution_fn
 * (DistributionLayers.py:245-255) as called from BaseEstimator.py:57,75,86.
 *   t      [B, P] device      y [y_rows, d] device, y_rows == B or 1 (broadcast)
 *   logp   [B]    device out
 */
int nfn_chain_forward(const nfn_chain_desc* desc, const float* t, const float* y,
                      int64_t y_rows, float* logp, int64_t B, void* stream);

/*
 * Outer-product density grid: every parameter row against every event of y_grid[n_y, d] -- the loop
 * `for i in range(y_num): dist.prob(y[i])` of plot_model (evaluation/visualization/flow_plotting.py:33-53).
 * logp is [n_y, B] (event-major).  Each parameter tile is staged once and reused for all n_y events,
 * so the parameter tensor is read from HBM once instead of n_y times.
 */
int nfn_chain_forward_grid(const nfn_chain_desc* desc, const float* t, const float* y_grid, int64_t n_y,
                           float* logp, int64_t B, void* stream);

/*
 * Fused forward + reverse sweep: replaces log_prob + tape.gradient of the Keras train
 * step (BaseEstimator.py:55-59 loss closure under Sequential.fit).
 *   g_logp     [B] device, nullable: upstream cotangent of logp per row
 *   g_scale    scalar multiplied into the cotangent (cot_b = g_scale * (g_logp ? g_logp[b] : 1));
 *              mean-NLL training passes g_logp = NULL, g_scale = -1/B_global
 *   logp       [B] device out
 *   dt         [B, P] device out:  cot_b * d logp_b / d t[b, :]
 *   dy         [B, d] device out, nullable: cot_b * d logp_b / d y[b, :] (requires y_rows == B)
 *   logp_sum   device double*, nullable: += sum_b logp_b  (one atomic per CTA)
 *   dt_colsum  device double[P], nullable: += sum_b dt[b, :]  (gradient of the bias of the
 *              Dense(P) layer that emits t, MaximumLikelihoodNNEstimator.py:43).  logp_sum and
 *              dt_colsum may be adjacent in ONE fp64 buffer so that a data-parallel step
 *              reduces both with a single all-reduce.
 */
int nfn_chain_forward_backward(const nfn_chain_desc* desc, const float* t, const float* y,
                               int64_t y_rows, const float* g_logp, float g_scale, float* logp,
                               float* dt, float* dy, double* logp_sum, double* dt_colsum,
                               int64_t B, void* stream);

/*
 * Fused compute + all-reduce for data-parallel training on the GPUs of one box (one process
 * per GPU): as nfn_chain_forward_backward, and the LAST CTA of the launch exchanges this
 * rank's fp64 totals [dt column sums (P) | sum logp] with every peer through NVLink peer
 * memory (push into each peer's IPC-mapped region, flag, wait, sum in rank order --
 * deterministic), so the step needs no separate collective launch.  `reduced` (device
 * double[P + 1]) holds the sums over all ranks when the kernel completes.  Every rank of the
 * communicator must make the same sequence of calls.  A rank that does not arrive within the
 * time-out (30 s; NFN_B200_PEER_TIMEOUT_S at communicator creation) is abandoned instead of hanging
 * the GPU: `reduced` gets NaN AND a sticky device flag is raised that nfn_peer_status() reports as
 * NFN_ERR_PEER_TIMEOUT -- never a silent NaN.  A call that fails does not advance the exchange
 * sequence.  want_colsum = 0 skips the in-kernel column sums (their slots stay 0).
 *
 * Split-phase mode (nfn_peer_set_deferred(comm, 1)): a launch leaves its totals in local accumulators;
 * ONE CTA of the NEXT launch's grid on the communicator (it carries no tiles) pushes them to the peers, collects
 * the peers' totals, writes the cross-rank sums into the `reduced` pointer the earlier call was given, and exits.
 * `reduced` of call k is therefore complete when call k+1 (or nfn_peer_flush) completes.  The NVLink round trip
 * and up to one kernel duration of rank skew overlap with a whole kernel of tile work on the other CTAs, and
 * nothing sits at any kernel's tail: no kernel's completion waits for remote stores of its own.
 *
 * Communicator set-up (see normalizingflownetwork_b200/parallel.py:PeerComm):
 *   nfn_peer_alloc        cudaMalloc + zero one region, export its 64-byte cudaIpc handle
 *   (exchange the handles between the processes, e.g. torch.distributed.all_gather)
 *   nfn_peer_open         map a peer's region (cudaIpcOpenMemHandle, lazy peer access)
 *   nfn_peer_comm_create  regions[world] in rank order (own region = the nfn_peer_alloc pointer)
 *   nfn_peer_allreduce    stand-alone exchange of device double[n_values] (one tiny kernel)
 */
typedef struct nfn_peer_comm nfn_peer_comm;
int64_t nfn_peer_region_bytes(int world, int n_values);
int nfn_peer_alloc(int world, int n_values, void** region, unsigned char* handle64);
int nfn_peer_open(const unsigned char* handle64, void** mapped);
int nfn_peer_close(void* mapped);
int nfn_peer_free(void* region);
int nfn_peer_comm_create(int world, int rank, int n_values, void* const* regions, nfn_peer_comm** comm);
int nfn_peer_comm_destroy(nfn_peer_comm* comm);
int nfn_peer_allreduce(nfn_peer_comm* comm, const double* values, double* reduced, void* stream);
int nfn_peer_set_deferred(nfn_peer_comm* comm, int deferred);   /* 1: split-phase mode (no exchange may be pending) */
int nfn_peer_flush(nfn_peer_comm* comm, void* stream);          /* split-phase: complete the pending exchange */
int nfn_peer_status(nfn_peer_comm* comm);                       /* synchronises; NFN_OK or NFN_ERR_PEER_TIMEOUT */
int nfn_chain_forward_backward_peer(const nfn_chain_desc* desc, const float* t, const float* y,
                                    int64_t y_rows, const float* g_logp, float g_scale, float* logp,
                                    float* dt, float* dy, int want_colsum, nfn_peer_comm* comm,
                                    double* reduced, int64_t B, void* stream);

/*
 * The emitting Dense(P) layer fused into the flow-chain kernel: replaces
 * `Dense(output_size, "linear")` (MaximumLikelihoodNNEstimator.py:43) + the NF layer's log_prob +
 * their tape gradients.  t = h W + bias is formed tile by tile in shared memory on the tensor cores
 * (3xTF32, fp32-level accuracy) and never written to HBM.
 *   h [B, hidden] device    W [hidden, P] row-major (Keras kernel layout)    bias [P]
 *   dh [B, hidden] out      dW [hidden, P] +=  h^T dt      dbias [P] += sum_b dt
 * hidden must be 16, 32, 48 or 64.  Returns NFN_ERR_UNSUPPORTED when no fused kernel can serve the
 * request (other widths; no ahead-of-time instance and NVRTC unavailable): compose
 * t = h W + bias with nfn_chain_forward_backward then.
 */
int nfn_dense_chain_forward(const nfn_chain_desc* desc, int hidden, const float* h, const float* W,
                            const float* bias, const float* y, int64_t y_rows, float* logp, int64_t B,
                            void* stream);
int nfn_dense_chain_forward_backward(const nfn_chain_desc* desc, int hidden, const float* h, const float* W,
                                     const float* bias, const float* y, int64_t y_rows, const float* g_logp,
                                     float g_scale, float* logp, float* dh, float* dW, float* dbias,
                                     double* logp_sum, int64_t B, void* stream);
int64_t nfn_jit_dense_compile_check(const nfn_chain_desc* desc, int hidden, int accurate);
/* the same check for the tcgen05 / TMEM implementation (csrc/nfn_dense_tc5.cuh) */
int64_t nfn_jit_dense_tc5_compile_check(const nfn_chain_desc* desc, int hidden, int accurate);

/*
 * One hidden layer of the conditioning network, `Dense(units, activation)` (reference
 * estimators/MaximumLikelihoodNNEstimator.py:37-44), as one kernel each way:
 *   forward   out[B, N] = act(x[B, K] weight^T + bias)
 *   backward  dpre = dout * act'(out);  dx[B, K] = dpre weight (NULL: not wanted);  dweight += dpre^T x;
 *             dbias += sum_b dpre
 * weight is [N, K] row-major (out_features x in_features); act: 0 linear, 1 tanh, 2 relu, 3 sigmoid, 4 elu.
 * 1 <= K <= 64 and N in {8, 16, 32, 64}: nfn_dense_act_supported tells; other shapes stay with the caller's
 * GEMM library.  All row pointers 16-byte aligned when K (resp. N) is a multiple of 4.
 */
int nfn_dense_act_supported(int in_features, int units, int act);
int nfn_dense_act_forward(const float* x, const float* weight, const float* bias, int64_t B, int in_features,
                          int units, int act, float* out, void* stream);
int nfn_dense_act_backward(const float* x, const float* out, const float* dout, const float* weight, int64_t B,
                           int in_features, int units, int act, float* dx, float* dweight, float* dbias,
                           void* stream);
/* The same with the estimators' input normalisation fused into the layer (the FIRST layer of the network):
 * x is the raw conditioning input and every read of it becomes (x - x_mean) / (x_std + 1e-8), the Lambda layer of
 * reference estimators/MaximumLikelihoodNNEstimator.py:40.  x_mean / x_std are device float[K] (both or neither).
 * The backward form serves the first layer only (dx == NULL, K <= 4, N <= 32), else NFN_ERR_UNSUPPORTED. */
int nfn_dense_act_forward_x(const float* x, const float* x_mean, const float* x_std, const float* weight,
                            const float* bias, int64_t B, int in_features, int units, int act, float* out, void* stream);
int nfn_dense_act_backward_x(const float* x, const float* x_mean, const float* x_std, const float* out, const float* dout,
                             const float* weight, int64_t B, int in_features, int units, int act, float* dx,
                             float* dweight, float* dbias, void* stream);

/*
 * One bijector on its own: Flow(t, n_dims).forward(z) and ._forward_log_det_jacobian(z)
 * (PlanarFlow.py:68-80, RadialFlow.py:51-70, AffineFlow.py:7-9), the calls made by
 * tests/test_flows.py:19-41.  t [B, size(flow)], z [z_rows, d] (z_rows == B or 1),
 * z_out [B, d] nullable, fldj [B] nullable.  Always uses the accurate math path.
 */
int nfn_flow_forward(int flow_type, int n_dims, const float* t, const float* z, int64_t z_rows,
                     float* z_out, float* fldj, int64_t B, void* stream);

/*
 * MDN head: tfd.Mixture.log_prob built by GaussianMixtureLayer._get_distribution_fn
 * (DistributionLayers.py:196-212).  t [B, 2*K*d + K] =
 * [ (mu_k(d), sigma_raw_k(d))_{k<K} | logits(K) ];  same cotangent / output conventions
 * as nfn_chain_forward_backward.
 */
int nfn_mdn_forward(int n_centers, int n_dims, const float* t, const float* y, int64_t y_rows,
                    float* logp, int64_t B, void* stream);
int nfn_mdn_forward_backward(int n_centers, int n_dims, const float* t, const float* y,
                             int64_t y_rows, const float* g_logp, float g_scale, float* logp,
                             float* dt, float* dy, double* logp_sum, double* dt_colsum, int64_t B,
                             void* stream);

/*
 * KMN head: tfd.MixtureSameFamily.log_prob built by
 * GaussianKernelsLayer._get_distribution_fn (DistributionLayers.py:118-133).
 *   t      [B, M] logits        locs [M, d] device (shared centres)
 *   scales [M] device (shared isotropic bandwidths; may be negative, |.| is used)
 *   dscales [M] device, nullable: += sum_b cot_b * d logp_b / d scales[m]
 */
int nfn_kmn_forward(int n_components, int n_dims, const float* t, const float* y, int64_t y_rows,
                    const float* locs, const float* scales, float* logp, int64_t B, void* stream);
int nfn_kmn_forward_backward(int n_components, int n_dims, const float* t, const float* y,
                             int64_t y_rows, const float* locs, const float* scales,
                             const float* g_logp, float g_scale, float* logp, float* dt, float* dy,
                             float* dscales, double* logp_sum, int64_t B, void* stream);

/*
 * Event transform: the estimators' y pipeline fused into the heads as a prologue / epilogue,
 * so that normalisation, noise regularisation, the normalisation Jacobian and `pdf`'s exp cost no extra pass over
 * y or logp.  Replaces, per call (reference estimators/BaseEstimator.py):
 *   NFN_XF_NORMALISE  y' = (y - mean) / std                      :61-69 (`(x - y_mean) / y_std`; also :73, :82)
 *   NFN_XF_NOISE      y' += noise_std * N(0, 1), training only   :66-68 (GaussianNoise) -- Philox4x32-10, key `seed`,
 *                     counter (global row, offset + *offset_dev): the caller advances the offset once per call;
 *                     `offset_dev` (device uint64, nullable) lets a captured CUDA graph draw fresh noise per replay
 *   logp_shift        added to every log-prob: -sum(log y_std)    :57 (`+ tf.reduce_sum(tf.math.log(y_std))` of the
 *                     NLL), :86 (`log_prob - sum log y_std`)
 *   NFN_XF_EXP        logp[] receives exp(log-prob + logp_shift)  :75 (`prob / prod(y_std)`)
 * The `_x` entry points are the plain ones with one more argument (NULL = identity).  logp_sum accumulates what
 * is written to logp[]; dy (where requested) is the gradient with respect to the transformed event y'.
 */
enum { NFN_XF_NORMALISE = 1, NFN_XF_NOISE = 2, NFN_XF_EXP = 4 };
typedef struct nfn_event_xform {
  float mean[NFN_MAX_DIMS];
  float std[NFN_MAX_DIMS];
  float noise_std;
  float logp_shift;
  uint64_t seed;
  uint64_t offset;
  const uint64_t* offset_dev;
  int32_t flags;
  int32_t reserved;
} nfn_event_xform;

int nfn_chain_forward_x(const nfn_chain_desc* desc, const float* t, const float* y, int64_t y_rows,
                        float* logp, int64_t B, const nfn_event_xform* xf, void* stream);
int nfn_chain_forward_grid_x(const nfn_chain_desc* desc, const float* t, const float* y_grid, int64_t n_y,
                             float* logp, int64_t B, const nfn_event_xform* xf, void* stream);
int nfn_chain_forward_backward_x(const nfn_chain_desc* desc, const float* t, const float* y,
                                 int64_t y_rows, const float* g_logp, float g_scale, float* logp,
                                 float* dt, float* dy, double* logp_sum, double* dt_colsum, int64_t B,
                                 const nfn_event_xform* xf, void* stream);
int nfn_dense_chain_forward_x(const nfn_chain_desc* desc, int hidden, const float* h, const float* W,
                              const float* bias, const float* y, int64_t y_rows, float* logp, int64_t B,
                              const nfn_event_xform* xf, void* stream);
int nfn_dense_chain_forward_backward_x(const nfn_chain_desc* desc, int hidden, const float* h, const float* W,
                                       const float* bias, const float* y, int64_t y_rows, const float* g_logp,
                                       float g_scale, float* logp, float* dh, float* dW, float* dbias,
                                       double* logp_sum, int64_t B, const nfn_event_xform* xf, void* stream);
int nfn_mdn_forward_x(int n_centers, int n_dims, const float* t, const float* y, int64_t y_rows,
                      float* logp, int64_t B, const nfn_event_xform* xf, void* stream);
int nfn_mdn_forward_backward_x(int n_centers, int n_dims, const float* t, const float* y,
                               int64_t y_rows, const float* g_logp, float g_scale, float* logp,
                               float* dt, float* dy, double* logp_sum, double* dt_colsum, int64_t B,
                               const nfn_event_xform* xf, void* stream);
/*
 * S posterior weight draws folded into the batch: what the Bayesian estimators' Monte-Carlo loop
 * (BayesianNNEstimator.py:65-76: `for _ in range(posterior_draws): self(x).log_prob(y)`, and the S-draw training
 * step of BASELINE config 4) becomes when every draw is a slab of ONE launch.  Rows are draw-major: row
 * s * rows_per_draw + b is sample b under weight draw s.
 *   nfn_dense_act_*_draws         the first tfp.layers.DenseVariational layer (BayesianNNEstimator.py:103-118):
 *                                 x [rows_per_draw, in] is read per SAMPLE, w [draws, in*units + units] is the
 *                                 layer's flat per-draw sample [kernel (in x units) | bias]; out / dout
 *                                 [draws * rows_per_draw, out_width] with columns units .. out_width zero (the row
 *                                 width the fused head wants); dw [draws, in*units + units] += per-draw gradient.
 *                                 x_mean / x_std (nullable together) fuse the input normalisation
 *                                 (MaximumLikelihoodNNEstimator.py:40).  in <= 8, units <= 64, out_width % 8 == 0.
 *   nfn_dense_chain_*_draws_x     the emitting DenseVariational layer + the flow chain, per-draw weights
 *                                 W [draws, hidden, P], bias [draws, P], dW / dbias likewise (+=); y [rows_per_draw, d]
 *                                 per SAMPLE (not repeated); h, logp, dh, g_logp are folded.  Served by the
 *                                 mma.sync kernel (weights re-staged per draw); B = draws * rows_per_draw.
 */
int nfn_dense_act_forward_draws(const float* x, const float* x_mean, const float* x_std, const float* w, int draws,
                                int64_t rows_per_draw, int in_features, int units, int out_width, int act, float* out,
                                void* stream);
int nfn_dense_act_backward_draws(const float* x, const float* x_mean, const float* x_std, const float* out,
                                 const float* dout, int draws, int64_t rows_per_draw, int in_features, int units,
                                 int out_width, int act, float* dw, void* stream);
int nfn_dense_chain_forward_draws_x(const nfn_chain_desc* desc, int hidden, int draws, int64_t rows_per_draw,
                                    const float* h, const float* W, const float* bias, const float* y, int64_t y_rows,
                                    float* logp, const nfn_event_xform* xf, void* stream);
int nfn_dense_chain_forward_backward_draws_x(const nfn_chain_desc* desc, int hidden, int draws, int64_t rows_per_draw,
                                             const float* h, const float* W, const float* bias, const float* y,
                                             int64_t y_rows, const float* g_logp, float g_scale, float* logp, float* dh,
                                             float* dW, float* dbias, double* logp_sum, const nfn_event_xform* xf,
                                             void* stream);
/*
 * Weight-space arithmetic of one tfp.layers.DenseVariational layer with the reference's mean-field posterior and
 * normal prior (estimators/DistributionLayers.py:17-71, BayesianNNEstimator.py:78-118), one launch each way instead
 * of ~40 elementwise launches: params [2 n] = (loc | raw scale), sigma = 1e-3 + softplus(log(e - 1) + 0.05 raw).
 *   nfn_variational_sample           w [draws, n] = loc + sigma * eps[draws, n];  *kl += KL(q || N(prior_loc, prior_scale))
 *                                    (kl device double, nullable)
 *   nfn_variational_sample_backward  dparams [2 n] += d/dparams of (sum dw . w + g_kl * KL), dprior_loc [n] += likewise
 *                                    (nullable: fixed prior); dw nullable (KL only), g_kl device float (nullable = 0)
 */
int nfn_variational_sample(const float* params, const float* prior_loc, float prior_scale, const float* eps, int n,
                           int draws, float* w, double* kl, void* stream);
int nfn_variational_sample_backward(const float* params, const float* prior_loc, float prior_scale, const float* eps,
                                    const float* dw, const float* g_kl, int n, int draws, float* dparams,
                                    float* dprior_loc, void* stream);
/*
 * One S-draw training step of a Bayesian estimator with ONE hidden variational layer, network part, in a single call
 * (BayesianNNEstimator.py:103-146 with the Monte-Carlo draws of :65-76 folded into the batch; BASELINE config 4):
 *   weight samples + exact KL of both tfp.layers.DenseVariational layers   (nfn_variational_sample)
 *   first layer over the draws * rows_per_draw folded rows                  (nfn_dense_act_forward_draws)
 *   emitting layer + density head + both backward GEMMs, per-draw weights   (nfn_dense_chain / _mdn _draws_x)
 *   first layer's per-draw weight gradient                                  (nfn_dense_act_backward_draws)
 *   gradients of the posterior parameters through the samples and the KL    (nfn_variational_sample_backward)
 * Seven launches, no host work between them -- what an eager framework spends ~0.7 ms of dispatch on.  The caller
 * supplies the standard-normal draws eps (so that its own generator stays the source of randomness), the workspaces
 * and the gradient buffers (+=).  head: mdn_centers == 0 -> the flow chain `desc`, else an MDN with that many centres
 * of desc->n_dims dimensions.  hidden_width: row width of h / dh (units rounded up to 16).
 */
typedef struct nfn_variational_layer {
  const float* posterior;   /* [2 n]: loc | raw scale */
  const float* prior_loc;   /* [n] */
  const float* eps;         /* [draws, n] */
  float* w;                 /* [draws, n]  workspace: the samples, flat [kernel | bias] per draw */
  float* dw;                /* [draws, n]  workspace: their gradient (zeroed by the call) */
  float* dposterior;        /* [2 n] += */
  float* dprior_loc;        /* [n] += , or NULL (fixed prior) */
  double* kl;               /* device scalar += KL(q || prior), or NULL */
  float prior_scale;
  float kl_grad;            /* d loss / d KL of this layer (kl_weight / world size) */
  int32_t n;                /* inputs * units + units */
  int32_t reserved;
} nfn_variational_layer;

int nfn_bayes_train_step(const nfn_chain_desc* desc, int mdn_centers, int draws, int64_t rows_per_draw, int in_features,
                         int units, int hidden_width, int act, const float* x, const float* x_mean, const float* x_std,
                         const float* y, int64_t y_rows, const nfn_variational_layer* first,
                         const nfn_variational_layer* emitting, float g_scale, float* h, float* dh, float* logp,
                         double* logp_sum, const nfn_event_xform* xf, void* stream);
/*
 * The emitting Dense(P) layer fused into the MDN head: replaces `Dense(output_size, "linear")`
 * (MaximumLikelihoodNNEstimator.py:43) + GaussianMixtureLayer's log_prob (DistributionLayers.py:196-212) + their
 * tape gradients, P = n_centers * (2 n_dims + 1).  Same contract as nfn_dense_chain_*_x: t = h W + bias is formed
 * tile by tile in shared memory (3xTF32 mma.sync, fp32-level accuracy), the mixture's per-row arithmetic is the
 * one of nfn_mdn_forward_backward, and per row the kernel moves 4 (2 hidden + n_dims + 1) bytes instead of
 * 4 (2 P + n_dims + 1).  hidden must be 16, 32, 48 or 64; NFN_ERR_UNSUPPORTED when no kernel can serve the request
 * (no ahead-of-time instance and NVRTC unavailable, or the tile does not fit shared memory): compose then.
 */
int nfn_dense_mdn_forward_x(int n_centers, int n_dims, int hidden, const float* h, const float* W, const float* bias,
                            const float* y, int64_t y_rows, float* logp, int64_t B, const nfn_event_xform* xf,
                            void* stream);
int nfn_dense_mdn_forward_backward_x(int n_centers, int n_dims, int hidden, const float* h, const float* W,
                                     const float* bias, const float* y, int64_t y_rows, const float* g_logp,
                                     float g_scale, float* logp, float* dh, float* dW, float* dbias, double* logp_sum,
                                     int64_t B, const nfn_event_xform* xf, void* stream);
/* ... with S folded weight draws (per-draw W [draws, hidden, P], bias [draws, P]; see nfn_dense_chain_*_draws_x):
 * the Bayesian mixture density network's Monte-Carlo loop (BayesMixtureDensityNetwork, BayesianNNEstimator.py:65-76) */
int nfn_dense_mdn_forward_draws_x(int n_centers, int n_dims, int hidden, int draws, int64_t rows_per_draw, const float* h,
                                  const float* W, const float* bias, const float* y, int64_t y_rows, float* logp,
                                  const nfn_event_xform* xf, void* stream);
int nfn_dense_mdn_forward_backward_draws_x(int n_centers, int n_dims, int hidden, int draws, int64_t rows_per_draw,
                                           const float* h, const float* W, const float* bias, const float* y,
                                           int64_t y_rows, const float* g_logp, float g_scale, float* logp, float* dh,
                                           float* dW, float* dbias, double* logp_sum, const nfn_event_xform* xf,
                                           void* stream);
/*
 * ... and into the KMN head: `Dense(output_size)` (MaximumLikelihoodNNEstimator.py:43) + GaussianKernelsLayer's
 * log_prob (DistributionLayers.py:118-133) + their tape gradients.  The layer emits the logits of n_components fixed
 * kernels; locs [n_components, n_dims] and scales [n_components] (negative bandwidths are legal, DistributionLayers.py
 * :98-116) are read once per CTA; dscales [n_components] += d sum(cot * logp) / d scales (nullable: fixed bandwidths).
 */
int nfn_dense_kmn_forward_x(int n_components, int n_dims, int hidden, const float* h, const float* W, const float* bias,
                            const float* y, int64_t y_rows, const float* locs, const float* scales, float* logp, int64_t B,
                            const nfn_event_xform* xf, void* stream);
int nfn_dense_kmn_forward_backward_x(int n_components, int n_dims, int hidden, const float* h, const float* W,
                                     const float* bias, const float* y, int64_t y_rows, const float* locs,
                                     const float* scales, const float* g_logp, float g_scale, float* logp, float* dh,
                                     float* dW, float* dbias, float* dscales, double* logp_sum, int64_t B,
                                     const nfn_event_xform* xf, void* stream);
int64_t nfn_jit_dense_kmn_compile_check(int n_components, int n_dims, int hidden, int accurate);
/* bytes of the NVRTC-built cubin for this mixture / hidden width (>= 0), or a negative nfn_status */
int64_t nfn_jit_dense_mdn_compile_check(int n_centers, int n_dims, int hidden, int accurate);
int nfn_kmn_forward_x(int n_components, int n_dims, const float* t, const float* y, int64_t y_rows,
                      const float* locs, const float* scales, float* logp, int64_t B, const nfn_event_xform* xf,
                      void* stream);
int nfn_kmn_forward_backward_x(int n_components, int n_dims, const float* t, const float* y,
                               int64_t y_rows, const float* locs, const float* scales,
                               const float* g_logp, float g_scale, float* logp, float* dt, float* dy,
                               float* dscales, double* logp_sum, int64_t B, const nfn_event_xform* xf, void* stream);

/*
 * Posterior-predictive epilogue of BayesianNNEstimator.score
 * (BayesianNNEstimator.py:65-76, evaluation/scorers.py:13-27) for S draws folded into
 * the batch: logp_sb [S, B] -> out[b] = logsumexp_s(logp_sb[s, b]) - log(S).
 */
int nfn_logmeanexp_draws(const float* logp_sb, int64_t S, int64_t B, float* out, void* stream);

/*
 * HOST-buffer entry points (the end-to-end call): same semantics as the device entry
 * points with every pointer a HOST pointer.  logp_sum is a host double* (overwritten,
 * nullable); dt_colsum a host double[P] (overwritten, nullable).  Copies are chunked and
 * overlapped with compute on library-owned streams.  Blocking.
 */
int nfn_chain_forward_host(const nfn_chain_desc* desc, const float* t, const float* y,
                           int64_t y_rows, float* logp, int64_t B);
int nfn_chain_forward_backward_host(const nfn_chain_desc* desc, const float* t, const float* y,
                                    int64_t y_rows, const float* g_logp, float g_scale,
                                    float* logp, float* dt, double* logp_sum, double* dt_colsum,
                                    int64_t B);
int nfn_mdn_forward_backward_host(int n_centers, int n_dims, const float* t, const float* y,
                                  int64_t y_rows, const float* g_logp, float g_scale, float* logp,
                                  float* dt, double* logp_sum, int64_t B);
/* Free this thread's host-pipeline workspace (streams, events, staging buffers). */
int nfn_host_release(void);

/* Number of kernel launches issued by this thread since the last call (then reset). */
int64_t nfn_launch_count_reset(void);

/* Math path of the chain / mixture kernels: 0 = fast (default; one MUFU per transcendental,
 * cancellation-free forms), 1 = accurate (CUDA libm, IEEE division).  Process-wide; the
 * environment variable NFN_B200_MATH=accurate selects 1 at first use. */
int nfn_set_math_mode(int accurate);

/* Process-wide run-time switches (tests, A/B tuning).  Each has an environment default that is read once,
 * at first use; no launch path calls getenv.  Names and values:
 *   "math"          0 fast | 1 accurate                                   (NFN_B200_MATH=accurate)
 *   "force_generic" 1: every chain through the runtime-chain kernel       (NFN_B200_FORCE_GENERIC)
 *   "force_jit"     1: skip the ahead-of-time kernel instances            (NFN_B200_FORCE_JIT)
 *   "jit"           0: runtime specialiser off                            (NFN_B200_JIT=0)
 *   "chain_io"      -1 measured default | 0 cp.async CTA-tile kernels | 1 bulk-copy / TMA warp-tile kernels
 *                                                                         (NFN_B200_CHAIN_IO=cpasync|tma)
 *   "dense_mma"     0 auto | 1 tcgen05 | 2 mma.sync                       (NFN_B200_DENSE_MMA=tc5|sync)
 *   "pdl"           0: no programmatic dependent launch                   (NFN_B200_PDL=0)
 *   "mlp_mma"       0: hidden layers' backward on the scalar kernels only  (NFN_B200_MLP_MMA=0)
 *   "host_chunk_mb" chunk size of the *_host pipelines, 1..1024 (default 16) (NFN_B200_HOST_CHUNK_MB)
 *   "debug"         1: launch geometry of the fused kernels on stderr       (NFN_B200_DEBUG)
 *   "tune_wnb", "tune_wwarps"  geometry overrides of the runtime-specialised warp-tile kernels (tools only)
 * Unknown names return NFN_ERR_DESC. */
int nfn_set_option(const char* name, int value);
int nfn_get_option(const char* name);

const char* nfn_last_error(void);
int nfn_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NFN_B200_H */
