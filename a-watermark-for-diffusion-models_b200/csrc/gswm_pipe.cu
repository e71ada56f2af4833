// Host-buffer layer of libgswm (gswm_pipe_*): streams batches that live in HOST memory through the
// device kernels in chunks, two slots deep, so that the PCIe copy of one chunk overlaps the kernel
// (and the copy) of the next.  This is the path a caller holding numpy / CPU torch buffers takes --
// the reference builds latents on the CPU and `.to(device)`s them (README.md:112) and extract.py:70
// returns a CPU tensor.
//
// Each slot owns one stream; everything for a chunk is enqueued in order on its slot's stream, so device
// buffer reuse two chunks later is ordered by the stream itself.  Nothing in the chunk loop blocks the
// host: key material is uploaded once per call through a pinned staging buffer, and the small per-chunk
// outputs (messages, counts, matched) are WRITTEN BY THE EXTRACT KERNEL ITSELF into per-slot pinned host
// staging (mapped memory, posted PCIe writes) and handed to the caller's (possibly pageable) arrays when
// the slot comes round again.  They must not go through the copy engine: a 32-byte D2H copy queues behind
// whatever 16 MB D2H copy another pipe (the embed side) has in flight on the same engine, and the slot's
// next H2D cannot start until it is through -- measured 7.0 ms instead of 5.4 ms per 4096-latent pair with
// both directions busy (tools/e2ebench.py).  The big latent buffers are copied straight from / to the
// caller's memory: pinned memory there gives the full PCIe rate.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <new>

#include "../../include/gswm.h"

namespace {

constexpr int kSlots = 2;
constexpr int kMaxMsgBytes = 1024;   // msg_bits <= 8192

struct Slot {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;  // recorded after the chunk's kernel (its outputs are then visible in the staging)
  void* d_in = nullptr;        // z (extract) or u (injected embed): chunk * max_elems * 8 bytes
  void* d_out = nullptr;       // latents out (embed): chunk * max_elems * 8 bytes
  // pinned, device-visible host staging the extract kernel writes its small outputs into
  uint8_t* h_msg_out = nullptr;   // chunk * kMaxMsgBytes
  uint16_t* h_counts = nullptr;   // chunk * 8192
  int32_t* h_matched = nullptr;   // chunk
  uint8_t* h_flags = nullptr;     // chunk
  // what the staging currently holds (to be delivered to the caller), -1 = nothing
  int64_t pending_first = -1, pending_n = 0;
};

}  // namespace

struct gswm_pipe {
  int device = 0;
  int64_t max_elems = 0;
  int64_t chunk = 0;
  Slot slot[kSlots];
  int64_t* d_counters = nullptr;
  int64_t* h_counters = nullptr;   // mapped host copy, written by publish_counters_kernel
  // key material of the current call, uploaded once: [keys | nonces | msgs], grown on demand
  uint8_t* d_km = nullptr;
  uint8_t* h_km = nullptr;      // pinned staging
  size_t km_capacity = 0;
};

namespace {

#define GSWM_CUDA(expr)                      \
  do {                                       \
    cudaError_t e_ = (expr);                 \
    if (e_ != cudaSuccess) return (int)e_;   \
  } while (0)

// The call's counters go back to the host the same way as the per-latent outputs: a one-warp kernel stores them into
// mapped host memory, so the result never waits in a copy-engine queue.
__global__ void publish_counters_kernel(const int64_t* __restrict__ d, int64_t* __restrict__ h) {
  if (threadIdx.x < GSWM_N_COUNTERS) h[threadIdx.x] = d[threadIdx.x];
}

int check_host_job(const gswm_pipe* p, const gswm_host_job* j, bool need_msg) {
  if (!p || !j || !j->h_keys || !j->h_nonces) return GSWM_E_NULL;
  if (need_msg && !j->h_msgs) return GSWM_E_NULL;
  if (j->n_latents < 0 || j->n_elems <= 0 || (j->n_elems % 4) != 0) return GSWM_E_SHAPE;
  if (j->n_elems > p->max_elems) return GSWM_E_RANGE;
  if (j->msg_bits <= 0 || j->msg_bits > j->n_elems) return GSWM_E_MSGLEN;
  if (need_msg && (j->msg_bits % 32) != 0) return GSWM_E_MSGLEN;      // embed (gswm_job rules); extract takes any divisor
  if (j->msg_bits > kMaxMsgBytes * 8) return GSWM_E_RANGE;
  return GSWM_OK;
}

inline int64_t msg_row_bytes(const gswm_host_job* j) { return (j->msg_bits + 7) / 8; }
inline bool host_per_latent(const gswm_host_job* j) { return (j->flags & GSWM_JOB_PER_LATENT) != 0; }

struct DeviceKeys {
  const uint8_t* keys;
  const uint8_t* nonces;
  const uint8_t* msgs;   // null when the host job has no messages
  int64_t msg_bytes;
};

// Upload the whole call's key material once (pinned staging -> device, on slot 0's stream; the other slots
// wait for it through an event).
int upload_keys(gswm_pipe* p, const gswm_host_job* hj, DeviceKeys* dk) {
  const int64_t rows = host_per_latent(hj) ? hj->n_latents : 1;
  const int64_t mb = msg_row_bytes(hj);
  const size_t need = (size_t)rows * (32 + 16 + (hj->h_msgs ? mb : 0));   // [keys | nonces | msgs]: msgs start 4-byte aligned (48 rows)
  if (need > p->km_capacity) {
    if (p->d_km) cudaFree(p->d_km);
    if (p->h_km) cudaFreeHost(p->h_km);
    p->d_km = nullptr; p->h_km = nullptr; p->km_capacity = 0;
    const size_t cap = std::max<size_t>(need, 1 << 16);
    GSWM_CUDA(cudaMalloc((void**)&p->d_km, cap));
    GSWM_CUDA(cudaMallocHost((void**)&p->h_km, cap));
    p->km_capacity = cap;
  }
  std::memcpy(p->h_km, hj->h_keys, (size_t)rows * 32);
  std::memcpy(p->h_km + rows * 32, hj->h_nonces, (size_t)rows * 16);
  if (hj->h_msgs) std::memcpy(p->h_km + rows * 48, hj->h_msgs, (size_t)rows * mb);
  GSWM_CUDA(cudaMemcpyAsync(p->d_km, p->h_km, need, cudaMemcpyHostToDevice, p->slot[0].stream));
  cudaEvent_t up;
  GSWM_CUDA(cudaEventCreateWithFlags(&up, cudaEventDisableTiming));
  int rc = (int)cudaEventRecord(up, p->slot[0].stream);
  for (int k = 1; k < kSlots && rc == 0; ++k) rc = (int)cudaStreamWaitEvent(p->slot[k].stream, up, 0);
  cudaEventDestroy(up);
  dk->keys = p->d_km;
  dk->nonces = p->d_km + rows * 32;
  dk->msgs = hj->h_msgs ? p->d_km + rows * 48 : nullptr;
  dk->msg_bytes = mb;
  return rc;
}

void chunk_job(const gswm_host_job* hj, const DeviceKeys& dk, int64_t first, int64_t n, gswm_job* dj) {
  const int64_t row0 = host_per_latent(hj) ? first : 0;
  dj->n_latents = n;
  dj->n_elems = hj->n_elems;
  dj->msg_bits = hj->msg_bits;
  dj->flags = hj->flags & GSWM_JOB_PER_LATENT;
  dj->d_keys = dk.keys + row0 * 32;
  dj->d_nonces = dk.nonces + row0 * 16;
  dj->d_msgs = dk.msgs ? dk.msgs + row0 * dk.msg_bytes : nullptr;
}

int sync_all(gswm_pipe* p, int rc) {
  for (auto& s : p->slot) {
    cudaError_t e = cudaStreamSynchronize(s.stream);
    if (rc == 0 && e != cudaSuccess) rc = (int)e;
  }
  return rc;
}

// Deliver what a slot's pinned staging holds to the caller's arrays (after the slot's last copy finished).
int drain_slot(Slot& s, int64_t msg_bytes, int64_t msg_bits, uint8_t* h_msg_out, uint16_t* h_counts, int32_t* h_matched,
               uint8_t* h_flags) {
  if (s.pending_first < 0) return GSWM_OK;
  GSWM_CUDA(cudaEventSynchronize(s.done));
  const int64_t f = s.pending_first, n = s.pending_n;
  std::memcpy(h_msg_out + f * msg_bytes, s.h_msg_out, (size_t)(n * msg_bytes));
  if (h_counts) std::memcpy(h_counts + f * msg_bits, s.h_counts, (size_t)(n * msg_bits) * sizeof(uint16_t));
  if (h_matched) std::memcpy(h_matched + f, s.h_matched, (size_t)n * sizeof(int32_t));
  if (h_flags) std::memcpy(h_flags + f, s.h_flags, (size_t)n);
  s.pending_first = -1;
  return GSWM_OK;
}

}  // namespace

extern "C" {

int gswm_pipe_create(gswm_pipe** out, int device, int64_t max_elems, int64_t max_latents_per_chunk) {
  if (!out) return GSWM_E_NULL;
  *out = nullptr;
  if (max_elems <= 0 || (max_elems % 4) != 0 || max_latents_per_chunk <= 0) return GSWM_E_SHAPE;
  GSWM_CUDA(cudaSetDevice(device));
  gswm_pipe* p = new (std::nothrow) gswm_pipe();
  if (!p) return (int)cudaErrorMemoryAllocation;
  p->device = device;
  p->max_elems = max_elems;
  p->chunk = max_latents_per_chunk;
  const size_t lat_bytes = (size_t)p->chunk * (size_t)max_elems * 8;
  int rc = 0;
  auto A = [&](void** ptr, size_t bytes) {
    if (rc == 0) rc = (int)cudaMalloc(ptr, bytes);
  };
  auto H = [&](void** ptr, size_t bytes) {   // pinned + mapped: under unified addressing the kernel uses the same pointer
    if (rc == 0) rc = (int)cudaHostAlloc(ptr, bytes, cudaHostAllocMapped);
  };
  for (auto& s : p->slot) {
    if (rc == 0) rc = (int)cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
    if (rc == 0) rc = (int)cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
    A(&s.d_in, lat_bytes);
    A(&s.d_out, lat_bytes);
    H((void**)&s.h_msg_out, (size_t)p->chunk * kMaxMsgBytes);
    H((void**)&s.h_counts, (size_t)p->chunk * 8192 * sizeof(uint16_t));
    H((void**)&s.h_matched, (size_t)p->chunk * sizeof(int32_t));
    H((void**)&s.h_flags, (size_t)p->chunk);
  }
  A((void**)&p->d_counters, GSWM_N_COUNTERS * sizeof(int64_t));
  H((void**)&p->h_counters, GSWM_N_COUNTERS * sizeof(int64_t));
  if (rc != 0) {
    gswm_pipe_destroy(p);
    return rc;
  }
  *out = p;
  return GSWM_OK;
}

void gswm_pipe_destroy(gswm_pipe* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  for (auto& s : p->slot) {
    if (s.stream) cudaStreamSynchronize(s.stream);
    cudaFree(s.d_in); cudaFree(s.d_out);
    cudaFreeHost(s.h_msg_out); cudaFreeHost(s.h_counts); cudaFreeHost(s.h_matched); cudaFreeHost(s.h_flags);
    if (s.done) cudaEventDestroy(s.done);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  cudaFree(p->d_counters);
  cudaFreeHost(p->h_counters);
  cudaFree(p->d_km);
  cudaFreeHost(p->h_km);
  delete p;
}

int gswm_pipe_embed(gswm_pipe* p, const gswm_host_job* job, uint64_t seed, uint64_t offset, int64_t first_latent,
                    float* h_out) {
  int rc = check_host_job(p, job, true);
  if (rc) return rc;
  if (!h_out) return GSWM_E_NULL;
  GSWM_CUDA(cudaSetDevice(p->device));
  DeviceKeys dk;
  if ((rc = upload_keys(p, job, &dk))) return sync_all(p, rc);
  const size_t row_bytes = (size_t)job->n_elems * sizeof(float);
  int c = 0;
  for (int64_t first = 0; first < job->n_latents && rc == 0; first += p->chunk, ++c) {
    Slot& s = p->slot[c % kSlots];
    const int64_t n = std::min(p->chunk, job->n_latents - first);
    gswm_job dj;
    chunk_job(job, dk, first, n, &dj);
    if ((rc = gswm_embed(&dj, seed, offset, first_latent + first, (float*)s.d_out, s.stream))) break;
    rc = (int)cudaMemcpyAsync(reinterpret_cast<char*>(h_out) + (size_t)first * row_bytes, s.d_out, (size_t)n * row_bytes,
                              cudaMemcpyDeviceToHost, s.stream);
  }
  return sync_all(p, rc);
}

int gswm_pipe_embed_injected(gswm_pipe* p, const gswm_host_job* job, const double* h_u, int32_t u_per_latent,
                             void* h_out, int32_t out_dtype) {
  int rc = check_host_job(p, job, true);
  if (rc) return rc;
  if (!h_out || !h_u) return GSWM_E_NULL;
  if (out_dtype != GSWM_F32 && out_dtype != GSWM_F64) return GSWM_E_DTYPE;
  GSWM_CUDA(cudaSetDevice(p->device));
  DeviceKeys dk;
  if ((rc = upload_keys(p, job, &dk))) return sync_all(p, rc);
  const size_t esz = out_dtype == GSWM_F32 ? 4 : 8;
  const size_t out_row = (size_t)job->n_elems * esz;
  const size_t u_row = (size_t)job->n_elems * sizeof(double);
  int c = 0;
  for (int64_t first = 0; first < job->n_latents && rc == 0; first += p->chunk, ++c) {
    Slot& s = p->slot[c % kSlots];
    const int64_t n = std::min(p->chunk, job->n_latents - first);
    gswm_job dj;
    chunk_job(job, dk, first, n, &dj);
    const char* u_src = reinterpret_cast<const char*>(h_u) + (u_per_latent ? (size_t)first * u_row : 0);
    if ((rc = (int)cudaMemcpyAsync(s.d_in, u_src, (u_per_latent ? (size_t)n : 1) * u_row, cudaMemcpyHostToDevice, s.stream))) break;
    if ((rc = gswm_embed_injected(&dj, (const double*)s.d_in, u_per_latent, s.d_out, out_dtype, s.stream))) break;
    rc = (int)cudaMemcpyAsync(reinterpret_cast<char*>(h_out) + (size_t)first * out_row, s.d_out, (size_t)n * out_row,
                              cudaMemcpyDeviceToHost, s.stream);
  }
  return sync_all(p, rc);
}

int gswm_pipe_extract(gswm_pipe* p, const gswm_host_job* job, const void* h_z, int32_t z_dtype, uint8_t* h_msg_out,
                      uint16_t* h_counts, int32_t* h_matched, uint8_t* h_flags, int64_t* h_counters) {
  int rc = check_host_job(p, job, false);
  if (rc) return rc;
  if (!h_z || !h_msg_out) return GSWM_E_NULL;
  if (z_dtype != GSWM_F32 && z_dtype != GSWM_F16 && z_dtype != GSWM_BF16 && z_dtype != GSWM_F64) return GSWM_E_DTYPE;
  if ((job->n_elems % job->msg_bits) != 0) return GSWM_E_MSGLEN;
  GSWM_CUDA(cudaSetDevice(p->device));
  const size_t esz = z_dtype == GSWM_F32 ? 4 : z_dtype == GSWM_F64 ? 8 : 2;
  const size_t z_row = (size_t)job->n_elems * esz;
  const int64_t mb = msg_row_bytes(job);
  const bool want_matched = h_matched && job->h_msgs;
  // counters are accumulated by both slots' kernels: zero them (with the key upload) on slot 0's stream, which the
  // other slots wait for inside upload_keys
  rc = (int)cudaMemsetAsync(p->d_counters, 0, GSWM_N_COUNTERS * sizeof(int64_t), p->slot[0].stream);
  DeviceKeys dk{};
  if (rc == 0) rc = upload_keys(p, job, &dk);
  for (auto& s : p->slot) s.pending_first = -1;
  int c = 0;
  for (int64_t first = 0; first < job->n_latents && rc == 0; first += p->chunk, ++c) {
    Slot& s = p->slot[c % kSlots];
    const int64_t n = std::min(p->chunk, job->n_latents - first);
    // the slot's staging still holds the outputs of chunk c - kSlots: hand them to the caller first
    if ((rc = drain_slot(s, mb, job->msg_bits, h_msg_out, h_counts, want_matched ? h_matched : nullptr, h_flags))) break;
    gswm_job dj;
    chunk_job(job, dk, first, n, &dj);
    if ((rc = (int)cudaMemcpyAsync(s.d_in, reinterpret_cast<const char*>(h_z) + (size_t)first * z_row, (size_t)n * z_row,
                                   cudaMemcpyHostToDevice, s.stream))) break;
    // small outputs: the kernel stores them straight into the slot's mapped host staging (no copy-engine work)
    if ((rc = gswm_extract(&dj, s.d_in, z_dtype, s.h_msg_out, h_counts ? s.h_counts : nullptr,
                           want_matched ? s.h_matched : nullptr, h_flags ? s.h_flags : nullptr, p->d_counters, s.stream))) break;
    if ((rc = (int)cudaEventRecord(s.done, s.stream))) break;
    s.pending_first = first;
    s.pending_n = n;
  }
  for (auto& s : p->slot) {
    const int r2 = drain_slot(s, mb, job->msg_bits, h_msg_out, h_counts, want_matched ? h_matched : nullptr, h_flags);
    if (rc == 0) rc = r2;
  }
  if (rc == 0 && h_counters) {               // after every slot's kernels: slot 0 waits for the others, then publishes
    for (int k = 1; k < kSlots && rc == 0; ++k) {
      rc = (int)cudaEventRecord(p->slot[k].done, p->slot[k].stream);
      if (rc == 0) rc = (int)cudaStreamWaitEvent(p->slot[0].stream, p->slot[k].done, 0);
    }
    if (rc == 0) {
      publish_counters_kernel<<<1, 32, 0, p->slot[0].stream>>>(p->d_counters, p->h_counters);
      rc = (int)cudaGetLastError();
    }
  }
  rc = sync_all(p, rc);
  if (rc == 0 && h_counters) std::memcpy(h_counters, p->h_counters, GSWM_N_COUNTERS * sizeof(int64_t));
  return rc;
}

}  // extern "C"
