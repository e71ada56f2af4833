// Host-buffer layer of libgswm (gswm_pipe_*): streams batches that live in HOST memory through the
// device kernels in chunks, two slots deep, so that the PCIe copy of one chunk overlaps the kernel
// (and the opposite-direction copy) of the next.  This is the path a caller holding numpy / CPU
// torch buffers takes -- the reference builds latents on the CPU and `.to(device)`s them
// (README.md:112) and extract.py:70 returns a CPU tensor.
//
// Each slot owns one stream; everything for a chunk is enqueued in order on its slot's stream, so
// buffer reuse two chunks later is ordered by the stream itself and no events are needed.
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
  void* d_in = nullptr;        // z (extract) or u (injected embed): chunk * max_elems * 8 bytes
  void* d_out = nullptr;       // latents out (embed): chunk * max_elems * 8 bytes
  uint8_t* d_keys = nullptr;   // chunk * 32
  uint8_t* d_nonces = nullptr; // chunk * 16
  uint8_t* d_msgs = nullptr;   // chunk * kMaxMsgBytes
  void* d_ws = nullptr;        // gswm_workspace_bytes for max_elems
  uint8_t* d_msg_out = nullptr;   // chunk * kMaxMsgBytes
  uint16_t* d_counts = nullptr;   // chunk * 8192
  int32_t* d_matched = nullptr;   // chunk
};

}  // namespace

struct gswm_pipe {
  int device = 0;
  int64_t max_elems = 0;
  int64_t chunk = 0;
  Slot slot[kSlots];
  int64_t* d_counters = nullptr;
};

namespace {

#define GSWM_CUDA(expr)                      \
  do {                                       \
    cudaError_t e_ = (expr);                 \
    if (e_ != cudaSuccess) return (int)e_;   \
  } while (0)

int check_host_job(const gswm_pipe* p, const gswm_host_job* j, bool need_msg) {
  if (!p || !j || !j->h_keys || !j->h_nonces) return GSWM_E_NULL;
  if (need_msg && !j->h_msgs) return GSWM_E_NULL;
  if (j->n_latents < 0 || j->n_elems <= 0 || (j->n_elems % 4) != 0) return GSWM_E_SHAPE;
  if (j->n_elems > p->max_elems) return GSWM_E_RANGE;
  if (j->msg_bits <= 0 || (j->msg_bits % 32) != 0 || j->msg_bits > j->n_elems) return GSWM_E_MSGLEN;
  if (j->msg_bits > kMaxMsgBytes * 8) return GSWM_E_RANGE;
  return GSWM_OK;
}

// Upload the key material of latents [first, first + n) of a host job into a slot and describe it
// as a device job.
int stage_job(const gswm_host_job* hj, Slot& s, int64_t first, int64_t n, gswm_job* dj) {
  const int64_t mb = hj->msg_bits / 8;
  const int64_t rows = hj->per_latent ? n : 1;
  const int64_t row0 = hj->per_latent ? first : 0;
  GSWM_CUDA(cudaMemcpyAsync(s.d_keys, hj->h_keys + row0 * 32, rows * 32, cudaMemcpyHostToDevice, s.stream));
  GSWM_CUDA(cudaMemcpyAsync(s.d_nonces, hj->h_nonces + row0 * 16, rows * 16, cudaMemcpyHostToDevice, s.stream));
  if (hj->h_msgs)
    GSWM_CUDA(cudaMemcpyAsync(s.d_msgs, hj->h_msgs + row0 * mb, rows * mb, cudaMemcpyHostToDevice, s.stream));
  dj->n_latents = n;
  dj->n_elems = hj->n_elems;
  dj->msg_bits = hj->msg_bits;
  dj->per_latent = hj->per_latent;
  dj->d_keys = s.d_keys;
  dj->d_nonces = s.d_nonces;
  dj->d_msgs = hj->h_msgs ? s.d_msgs : nullptr;
  return GSWM_OK;
}

int sync_all(gswm_pipe* p, int rc) {
  for (auto& s : p->slot) {
    cudaError_t e = cudaStreamSynchronize(s.stream);
    if (rc == 0 && e != cudaSuccess) rc = (int)e;
  }
  return rc;
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
  gswm_job probe{};
  probe.n_elems = max_elems;
  const size_t ws_bytes = gswm_workspace_bytes(&probe);
  int rc = 0;
  auto A = [&](void** ptr, size_t bytes) {
    if (rc == 0) rc = (int)cudaMalloc(ptr, bytes);
  };
  for (auto& s : p->slot) {
    if (rc == 0) rc = (int)cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
    A(&s.d_in, lat_bytes);
    A(&s.d_out, lat_bytes);
    A((void**)&s.d_keys, (size_t)p->chunk * 32);
    A((void**)&s.d_nonces, (size_t)p->chunk * 16);
    A((void**)&s.d_msgs, (size_t)p->chunk * kMaxMsgBytes);
    A(&s.d_ws, ws_bytes);
    A((void**)&s.d_msg_out, (size_t)p->chunk * kMaxMsgBytes);
    A((void**)&s.d_counts, (size_t)p->chunk * 8192 * sizeof(uint16_t));
    A((void**)&s.d_matched, (size_t)p->chunk * sizeof(int32_t));
  }
  A((void**)&p->d_counters, GSWM_N_COUNTERS * sizeof(int64_t));
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
    cudaFree(s.d_in); cudaFree(s.d_out); cudaFree(s.d_keys); cudaFree(s.d_nonces); cudaFree(s.d_msgs);
    cudaFree(s.d_ws); cudaFree(s.d_msg_out); cudaFree(s.d_counts); cudaFree(s.d_matched);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  cudaFree(p->d_counters);
  delete p;
}

int gswm_pipe_embed(gswm_pipe* p, const gswm_host_job* job, uint64_t seed, uint64_t offset, int64_t first_latent,
                    float* h_out) {
  int rc = check_host_job(p, job, true);
  if (rc) return rc;
  if (!h_out) return GSWM_E_NULL;
  GSWM_CUDA(cudaSetDevice(p->device));
  const size_t row_bytes = (size_t)job->n_elems * sizeof(float);
  int c = 0;
  for (int64_t first = 0; first < job->n_latents && rc == 0; first += p->chunk, ++c) {
    Slot& s = p->slot[c % kSlots];
    const int64_t n = std::min(p->chunk, job->n_latents - first);
    gswm_job dj;
    if ((rc = stage_job(job, s, first, n, &dj))) break;
    if ((rc = gswm_embed(&dj, seed, offset, first_latent + first, (float*)s.d_out, s.d_ws, s.stream))) break;
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
  const size_t esz = out_dtype == GSWM_F32 ? 4 : 8;
  const size_t out_row = (size_t)job->n_elems * esz;
  const size_t u_row = (size_t)job->n_elems * sizeof(double);
  int c = 0;
  for (int64_t first = 0; first < job->n_latents && rc == 0; first += p->chunk, ++c) {
    Slot& s = p->slot[c % kSlots];
    const int64_t n = std::min(p->chunk, job->n_latents - first);
    gswm_job dj;
    if ((rc = stage_job(job, s, first, n, &dj))) break;
    const char* u_src = reinterpret_cast<const char*>(h_u) + (u_per_latent ? (size_t)first * u_row : 0);
    if ((rc = (int)cudaMemcpyAsync(s.d_in, u_src, (u_per_latent ? (size_t)n : 1) * u_row, cudaMemcpyHostToDevice, s.stream))) break;
    if ((rc = gswm_embed_injected(&dj, (const double*)s.d_in, u_per_latent, s.d_out, out_dtype, s.d_ws, s.stream))) break;
    rc = (int)cudaMemcpyAsync(reinterpret_cast<char*>(h_out) + (size_t)first * out_row, s.d_out, (size_t)n * out_row,
                              cudaMemcpyDeviceToHost, s.stream);
  }
  return sync_all(p, rc);
}

int gswm_pipe_extract(gswm_pipe* p, const gswm_host_job* job, const void* h_z, int32_t z_dtype, uint8_t* h_msg_out,
                      uint16_t* h_counts, int32_t* h_matched, int64_t* h_counters) {
  int rc = check_host_job(p, job, false);
  if (rc) return rc;
  if (!h_z || !h_msg_out) return GSWM_E_NULL;
  if (z_dtype != GSWM_F32 && z_dtype != GSWM_F16 && z_dtype != GSWM_BF16) return GSWM_E_DTYPE;
  if ((job->n_elems % job->msg_bits) != 0) return GSWM_E_MSGLEN;
  GSWM_CUDA(cudaSetDevice(p->device));
  const size_t esz = z_dtype == GSWM_F32 ? 4 : 2;
  const size_t z_row = (size_t)job->n_elems * esz;
  const size_t mb = (size_t)job->msg_bits / 8;
  // counters are accumulated by both slots' kernels; zero them on slot 0 and make slot 1 wait for it
  cudaEvent_t zeroed;
  GSWM_CUDA(cudaEventCreateWithFlags(&zeroed, cudaEventDisableTiming));
  rc = (int)cudaMemsetAsync(p->d_counters, 0, GSWM_N_COUNTERS * sizeof(int64_t), p->slot[0].stream);
  if (rc == 0) rc = (int)cudaEventRecord(zeroed, p->slot[0].stream);
  for (int k = 1; k < kSlots && rc == 0; ++k) rc = (int)cudaStreamWaitEvent(p->slot[k].stream, zeroed, 0);
  int c = 0;
  for (int64_t first = 0; first < job->n_latents && rc == 0; first += p->chunk, ++c) {
    Slot& s = p->slot[c % kSlots];
    const int64_t n = std::min(p->chunk, job->n_latents - first);
    gswm_job dj;
    if ((rc = stage_job(job, s, first, n, &dj))) break;
    if ((rc = (int)cudaMemcpyAsync(s.d_in, reinterpret_cast<const char*>(h_z) + (size_t)first * z_row, (size_t)n * z_row,
                                   cudaMemcpyHostToDevice, s.stream))) break;
    if ((rc = gswm_extract(&dj, s.d_in, z_dtype, s.d_msg_out, h_counts ? s.d_counts : nullptr,
                           (h_matched && dj.d_msgs) ? s.d_matched : nullptr, p->d_counters, s.d_ws, s.stream))) break;
    if ((rc = (int)cudaMemcpyAsync(h_msg_out + (size_t)first * mb, s.d_msg_out, (size_t)n * mb, cudaMemcpyDeviceToHost, s.stream))) break;
    if (h_counts &&
        (rc = (int)cudaMemcpyAsync(h_counts + (size_t)first * job->msg_bits, s.d_counts,
                                   (size_t)n * job->msg_bits * sizeof(uint16_t), cudaMemcpyDeviceToHost, s.stream))) break;
    if (h_matched && dj.d_msgs &&
        (rc = (int)cudaMemcpyAsync(h_matched + first, s.d_matched, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream))) break;
  }
  rc = sync_all(p, rc);
  if (rc == 0 && h_counters)
    rc = (int)cudaMemcpy(h_counters, p->d_counters, GSWM_N_COUNTERS * sizeof(int64_t), cudaMemcpyDeviceToHost);
  cudaEventDestroy(zeroed);
  return rc;
}

}  // extern "C"
