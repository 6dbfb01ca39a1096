// nfn_peer.cu -- peer-memory communicator for the fused in-kernel all-reduce of the hot
// path's fp64 accumulators ([dt column sums | sum logp]) across the GPUs of one box.
//
// One process per GPU.  Each rank cudaMallocs one region (words[2][world][n][2], 8 bytes each),
// exports it with cudaIpcGetMemHandle; the Python side exchanges the 64-byte handles over
// torch.distributed and every rank maps its peers with cudaIpcOpenMemHandle (NVLink P2P).
// The exchange itself runs inside the compute kernel's last CTA (peer_allreduce in
// nfn_chain_kernel.cuh) or, for kernels without that epilogue, in a one-CTA kernel here.
#include <cstdlib>
#include <cstring>
#include <new>

#include "nfn_common.h"

struct nfn_peer_comm {
  int world = 0, rank = 0, n_values = 0;
  double* base[nfn::kMaxPeers] = {};
  double* acc = nullptr;       // [2][n_values] local accumulators (self-resetting), one set per step parity
  unsigned* ticket = nullptr;  // CTA arrival counter (self-resetting)
  unsigned* status = nullptr;  // sticky device error word (time-outs)
  unsigned long long step = 0; // exchanges issued so far
  int deferred = 0;            // split-phase mode
  double* pending_out = nullptr;        // split-phase: where the sums of exchange step-1 still have to go
  long long timeout_cycles = 60000000000ll;  // ~30 s at 2 GHz
};

namespace nfn {

// one CTA: blocking mode exchanges THIS step's totals, split-phase mode the previous step's
__global__ void __launch_bounds__(128) peer_allreduce_kernel(const PeerArgs p) {
  if (p.deferred) peer_exchange_prev(p, (int)threadIdx.x, 128);
  else peer_allreduce<128>(p);
}
// split-phase: send the last launch's totals and collect their sums (one CTA)
__global__ void __launch_bounds__(128) peer_flush_kernel(const PeerArgs p) {
  if (p.world > 0 && p.deferred) peer_exchange_prev(p, (int)threadIdx.x, 128);
}

// Arguments of the NEXT exchange.  Nothing in the communicator changes here: the caller commits
// (peer_commit) only after its launch succeeded, so a failed call leaves this rank's sequence number, parity
// and pending result in step with its peers.
PeerArgs make_peer_args(nfn_peer_comm* c, double* out) {
  PeerArgs p{};
  for (int i = 0; i < kMaxPeers; ++i) p.base[i] = c->base[i];
  p.acc = c->acc + (size_t)(c->step & 1ull) * c->n_values;          // this exchange's accumulators
  p.acc_prev = c->acc + (size_t)((c->step + 1ull) & 1ull) * c->n_values;  // the previous exchange's
  p.out = out;
  p.ticket = c->ticket;
  p.step = c->step;
  p.world = c->world;
  p.rank = c->rank;
  p.n_values = c->n_values;
  p.deferred = c->deferred;
  p.pending_out = c->deferred ? c->pending_out : nullptr;
  p.pending_step = c->step - 1;
  p.status = c->status;
  p.timeout_cycles = c->timeout_cycles;
  return p;
}

void peer_commit(nfn_peer_comm* c, double* out) {
  ++c->step;
  c->pending_out = c->deferred ? out : nullptr;
}

int launch_peer_allreduce(const PeerArgs& p, cudaStream_t st) {
  peer_allreduce_kernel<<<1, 128, 0, st>>>(p);
  count_launch();
  return cuda_error(cudaGetLastError(), "peer_allreduce_kernel");
}

static size_t region_bytes(int world, int n) {
  return (size_t)2 * world * n * 2 * sizeof(unsigned long long);  // words[2][world][n][2]
}

}  // namespace nfn

using namespace nfn;

extern "C" {

int64_t nfn_peer_region_bytes(int world, int n_values) {
  if (world < 1 || world > kMaxPeers || n_values < 1)
    return set_error(NFN_ERR_SHAPE, "world=%d (1..%d), n_values=%d", world, kMaxPeers, n_values);
  return (int64_t)region_bytes(world, n_values);
}

int nfn_peer_alloc(int world, int n_values, void** region, unsigned char* handle64) {
  if (!region || !handle64) return set_error(NFN_ERR_NULL, "region and handle must be non-NULL");
  if (world < 1 || world > kMaxPeers || n_values < 1)
    return set_error(NFN_ERR_SHAPE, "world=%d (1..%d), n_values=%d", world, kMaxPeers, n_values);
  const size_t bytes = region_bytes(world, n_values);
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return cuda_error(e, "cudaMalloc(peer region)");
  if ((e = cudaMemset(p, 0, bytes)) != cudaSuccess) return cuda_error(e, "cudaMemset(peer region)");
  cudaIpcMemHandle_t h;
  if ((e = cudaIpcGetMemHandle(&h, p)) != cudaSuccess) {
    cudaFree(p);
    return cuda_error(e, "cudaIpcGetMemHandle");
  }
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  *region = p;
  return NFN_OK;
}

int nfn_peer_open(const unsigned char* handle64, void** mapped) {
  if (!handle64 || !mapped) return set_error(NFN_ERR_NULL, "handle and mapped must be non-NULL");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  return cuda_error(cudaIpcOpenMemHandle(mapped, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
}

int nfn_peer_close(void* mapped) { return cuda_error(cudaIpcCloseMemHandle(mapped), "cudaIpcCloseMemHandle"); }

int nfn_peer_free(void* region) { return cuda_error(cudaFree(region), "cudaFree(peer region)"); }

int nfn_peer_comm_create(int world, int rank, int n_values, void* const* regions, nfn_peer_comm** comm) {
  if (!regions || !comm) return set_error(NFN_ERR_NULL, "regions and comm must be non-NULL");
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world || n_values < 1)
    return set_error(NFN_ERR_SHAPE, "world=%d (1..%d), rank=%d, n_values=%d", world, kMaxPeers, rank, n_values);
  nfn_peer_comm* c = new (std::nothrow) nfn_peer_comm();
  if (!c) return set_error(NFN_ERR_CUDA, "out of host memory");
  c->world = world;
  c->rank = rank;
  c->n_values = n_values;
  for (int i = 0; i < world; ++i) {
    if (!regions[i]) {
      delete c;
      return set_error(NFN_ERR_NULL, "regions[%d] is NULL", i);
    }
    c->base[i] = (double*)regions[i];
  }
  cudaError_t e = cudaMalloc((void**)&c->acc, (size_t)2 * n_values * sizeof(double) + 16);
  if (e == cudaSuccess) e = cudaMemset(c->acc, 0, (size_t)2 * n_values * sizeof(double) + 16);
  if (const char* ev = getenv("NFN_B200_PEER_TIMEOUT_S")) {  // set-up time only, never on a launch path
    const double sec = atof(ev);
    if (sec > 0.0) c->timeout_cycles = (long long)(sec * 2.0e9);
  }
  if (e != cudaSuccess) {
    delete c;
    return cuda_error(e, "cudaMalloc(peer accumulators)");
  }
  c->ticket = (unsigned*)(c->acc + 2 * n_values);
  c->status = c->ticket + 1;
  *comm = c;
  return NFN_OK;
}

int nfn_peer_comm_destroy(nfn_peer_comm* comm) {
  if (!comm) return NFN_OK;
  cudaFree(comm->acc);
  delete comm;
  return NFN_OK;
}

int nfn_peer_allreduce(nfn_peer_comm* comm, const double* values, double* reduced, void* stream) {
  if (!comm || !values || !reduced) return set_error(NFN_ERR_NULL, "comm, values and reduced must be non-NULL");
  cudaStream_t st = (cudaStream_t)stream;
  // accumulate the caller's values into the (zero) local accumulators, then exchange
  const PeerArgs pa = make_peer_args(comm, reduced);
  cudaError_t e = cudaMemcpyAsync(pa.acc, values, (size_t)comm->n_values * sizeof(double), cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return cuda_error(e, "cudaMemcpyAsync(values)");
  int rc = launch_peer_allreduce(pa, st);
  if (rc == NFN_OK) peer_commit(comm, reduced);
  return rc;
}

int nfn_peer_set_deferred(nfn_peer_comm* comm, int deferred) {
  if (!comm) return set_error(NFN_ERR_NULL, "comm is NULL");
  if (comm->pending_out) return set_error(NFN_ERR_SHAPE, "an exchange is still pending: call nfn_peer_flush first");
  comm->deferred = deferred ? 1 : 0;
  return NFN_OK;
}

int nfn_peer_flush(nfn_peer_comm* comm, void* stream) {
  if (!comm) return set_error(NFN_ERR_NULL, "comm is NULL");
  if (!comm->deferred || !comm->pending_out) return NFN_OK;
  PeerArgs p = make_peer_args(comm, nullptr);
  peer_flush_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(p);
  count_launch();
  int rc = cuda_error(cudaGetLastError(), "peer_flush_kernel");
  if (rc == NFN_OK) comm->pending_out = nullptr;
  return rc;
}

int nfn_peer_status(nfn_peer_comm* comm) {
  if (!comm) return set_error(NFN_ERR_NULL, "comm is NULL");
  unsigned st = 0;
  cudaError_t e = cudaMemcpy(&st, comm->status, sizeof(st), cudaMemcpyDeviceToHost);  // synchronises
  if (e != cudaSuccess) return cuda_error(e, "cudaMemcpy(peer status)");
  if (st != 0)
    return set_error(NFN_ERR_PEER_TIMEOUT,
                     "a peer did not arrive within the exchange time-out (%.1f s): the reduced sums of that step are NaN "
                     "and the ranks may have diverged", (double)comm->timeout_cycles / 2.0e9);
  return NFN_OK;
}

}  // extern "C"
