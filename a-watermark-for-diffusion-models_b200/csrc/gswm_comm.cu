// Host side of gswm_comm (include/gswm.h, "Multi-GPU"): mailbox allocation, peer mapping (CUDA IPC between processes,
// peer access inside one), and the entry points that launch the exchange -- stand-alone, or fused into the extract kernel.
// The device side is gswm_comm.cuh.  Also gswm_allreduce_counters: the same sum through a caller-owned NCCL communicator,
// libnccl resolved with dlopen so that libgswm.so has no link-time dependency on it.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdint>
#include <cstring>
#include <mutex>
#include <new>

#include "../../include/gswm.h"
#include "gswm_comm.cuh"
#include "gswm_internal.h"

using gswm::CommDev;
using gswm::CommSlot;

static_assert(sizeof(cudaIpcMemHandle_t) == GSWM_COMM_HANDLE_BYTES, "handle size");
static_assert(GSWM_N_COUNTERS <= GSWM_COMM_MAX_VALUES, "counter vector must fit a mailbox slot");

struct gswm_comm {
  int device = 0, rank = 0, n_ranks = 0;
  unsigned char* base = nullptr;                 // own allocation: mailbox [2][GSWM_COMM_MAX_RANKS] slots, then the ticket word
  int* h_status = nullptr;                       // mapped host word the kernels write on a timeout
  void* peer[GSWM_COMM_MAX_RANKS] = {};          // every rank's mailbox in this rank's address space
  bool ipc[GSWM_COMM_MAX_RANKS] = {};            // opened with cudaIpcOpenMemHandle (to be closed)
  bool connected = false;
  unsigned long long epoch = 0;
};

namespace {

constexpr size_t kBoxBytes = 2 * GSWM_COMM_MAX_RANKS * sizeof(CommSlot);

#define GSWM_CUDA(expr)                      \
  do {                                       \
    cudaError_t e_ = (expr);                 \
    if (e_ != cudaSuccess) return (int)e_;   \
  } while (0)

int device_view(gswm_comm* c, CommDev* d) {
  if (!c) return GSWM_E_NULL;
  if (!c->connected) return GSWM_E_COMM;
  for (int r = 0; r < GSWM_COMM_MAX_RANKS; ++r) d->box[r] = reinterpret_cast<CommSlot*>(c->peer[r < c->n_ranks ? r : c->rank]);
  d->ticket = reinterpret_cast<unsigned*>(c->base + kBoxBytes);
  d->status = c->h_status;
  d->epoch = ++c->epoch;
  d->rank = c->rank;
  d->n_ranks = c->n_ranks;
  return GSWM_OK;
}

}  // namespace

extern "C" {

int gswm_comm_create(gswm_comm** out, int device, int rank, int n_ranks, uint8_t* handle_out) {
  if (!out) return GSWM_E_NULL;
  *out = nullptr;
  if (n_ranks < 1 || n_ranks > GSWM_COMM_MAX_RANKS || rank < 0 || rank >= n_ranks) return GSWM_E_COMM;
  GSWM_CUDA(cudaSetDevice(device));
  gswm_comm* c = new (std::nothrow) gswm_comm();
  if (!c) return (int)cudaErrorMemoryAllocation;
  c->device = device; c->rank = rank; c->n_ranks = n_ranks;
  int rc = (int)cudaMalloc((void**)&c->base, kBoxBytes + 128);
  if (rc == 0) rc = (int)cudaMemset(c->base, 0, kBoxBytes + 128);
  if (rc == 0) rc = (int)cudaHostAlloc((void**)&c->h_status, sizeof(int), cudaHostAllocMapped);
  if (rc == 0) *c->h_status = 0;
  if (rc == 0 && handle_out) {
    cudaIpcMemHandle_t h;
    rc = (int)cudaIpcGetMemHandle(&h, c->base);
    if (rc == 0) std::memcpy(handle_out, &h, sizeof(h));
  }
  if (rc == 0) rc = (int)cudaDeviceSynchronize();
  if (rc != 0) {
    gswm_comm_destroy(c);
    return rc;
  }
  c->peer[rank] = c->base;
  c->connected = n_ranks == 1;
  *out = c;
  return GSWM_OK;
}

int gswm_comm_connect(gswm_comm* c, const uint8_t* all_handles) {
  if (!c || !all_handles) return GSWM_E_NULL;
  GSWM_CUDA(cudaSetDevice(c->device));
  for (int r = 0; r < c->n_ranks; ++r) {
    if (r == c->rank || c->peer[r]) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, all_handles + (size_t)r * GSWM_COMM_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      return GSWM_E_COMM;
    }
    c->peer[r] = p;
    c->ipc[r] = true;
  }
  c->connected = true;
  return GSWM_OK;
}

int gswm_comm_connect_local(gswm_comm* const* comms, int n_ranks) {
  if (!comms) return GSWM_E_NULL;
  if (n_ranks < 1 || n_ranks > GSWM_COMM_MAX_RANKS) return GSWM_E_COMM;
  for (int i = 0; i < n_ranks; ++i)
    if (!comms[i] || comms[i]->n_ranks != n_ranks || comms[i]->rank != i) return GSWM_E_COMM;
  for (int i = 0; i < n_ranks; ++i) {
    gswm_comm* c = comms[i];
    GSWM_CUDA(cudaSetDevice(c->device));
    for (int r = 0; r < n_ranks; ++r) {
      if (r == i) continue;
      if (comms[r]->device != c->device) {
        int can = 0;
        GSWM_CUDA(cudaDeviceCanAccessPeer(&can, c->device, comms[r]->device));
        if (!can) return GSWM_E_COMM;
        const cudaError_t e = cudaDeviceEnablePeerAccess(comms[r]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return (int)e;
        cudaGetLastError();
      }
      c->peer[r] = comms[r]->base;
    }
    c->connected = true;
  }
  return GSWM_OK;
}

void gswm_comm_destroy(gswm_comm* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < GSWM_COMM_MAX_RANKS; ++r)
    if (c->ipc[r] && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
  cudaFree(c->base);
  cudaFreeHost(c->h_status);
  delete c;
}

int gswm_comm_status(gswm_comm* c) {
  if (!c) return GSWM_E_NULL;
  return *reinterpret_cast<volatile int*>(c->h_status);
}

int gswm_comm_allreduce_counters(gswm_comm* c, int64_t* d_counters, int32_t n, void* stream) {
  if (!c || !d_counters) return GSWM_E_NULL;
  if (n < 1 || n > GSWM_COMM_MAX_VALUES) return GSWM_E_RANGE;
  CommDev d;
  const int rc = device_view(c, &d);
  if (rc) return rc;
  return gswm::comm_allreduce_launch(d, d_counters, n, stream);
}

int gswm_extract_allreduce(const gswm_job* job, const void* d_z, int32_t z_dtype, uint8_t* d_msg_out, uint16_t* d_counts,
                           int32_t* d_matched, uint8_t* d_flags, int64_t* d_counters, gswm_comm* c, int64_t* d_reduced,
                           void* stream) {
  if (!c || !d_counters || !d_reduced) return GSWM_E_NULL;
  if (!c->connected) return GSWM_E_COMM;
  // argument errors must not consume an epoch: every rank's epochs advance in lock step
  int rc = gswm::check_job(job, true);
  if (rc) return rc;
  CommDev d;
  if (job->n_latents == 0) {
    // a rank whose shard is empty still takes part in the collective (the others wait for it): no extraction, the
    // accumulated counters go through the stand-alone exchange
    GSWM_CUDA(cudaMemcpyAsync(d_reduced, d_counters, sizeof(int64_t) * GSWM_N_COUNTERS, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    if ((rc = device_view(c, &d))) return rc;
    return gswm::comm_allreduce_launch(d, d_reduced, GSWM_N_COUNTERS, stream);
  }
  if ((rc = device_view(c, &d))) return rc;
  rc = gswm::extract_impl(job, d_z, z_dtype, d_msg_out, d_counts, d_matched, d_flags, d_counters, &d, d_reduced, stream);
  if (rc) --c->epoch;                                                  // nothing was launched
  return rc;
}

// ---- NCCL, for hosts that already own a communicator ---------------------------------------------------------------------
int gswm_allreduce_counters(void* nccl_comm, int64_t* d_counters, int32_t n, void* stream) {
  if (!nccl_comm || !d_counters) return GSWM_E_NULL;
  if (n < 1) return GSWM_E_RANGE;
  typedef int (*allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  static allreduce_fn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    // the copy already mapped into the process (torch's bundled one, or the host's own) if there is one: the communicator
    // must be used with the library that created it
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW);
    if (h) fn = reinterpret_cast<allreduce_fn>(dlsym(h, "ncclAllReduce"));
  });
  if (!fn) return GSWM_E_COMM;
  const int kNcclInt64 = 4, kNcclSum = 0;                               // nccl.h: ncclDataType_t / ncclRedOp_t
  const int r = fn(d_counters, d_counters, (size_t)n, kNcclInt64, kNcclSum, nccl_comm, (cudaStream_t)stream);
  return r == 0 ? GSWM_OK : GSWM_E_COMM;
}

}  // extern "C"
