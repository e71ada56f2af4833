// The reference's own uniform stream on the device: numpy's legacy MT19937 (`np.random.RandomState(seed).uniform(0, 1)`,
// nodes.py:52-53,114-117; v1.5.2:27,72-75), bit for bit, and the seeded embed built on it.
//
// One CTA per stream.  The 624-word state lives in shared memory, double-buffered; a regeneration is the textbook
// recurrence  new[i] = new_or_old[i + 397 mod 624] ^ twist(old[i], old[i + 1])  done in three barrier-separated phases
// ([0,227) reads only the old block, [227,454) reads what the first phase wrote, [454,624) what the second wrote), so
// 227 threads work per phase instead of one.  Tempered words are paired into 53-bit doubles exactly as genrand_res53
// does ((a >> 5) * 2^26 + (b >> 6)) / 2^53, 312 per regeneration, and queued in a small shared ring from which the CTA
// takes 256 at a time: element e of the latent gets the e-th double of the stream, as the reference's per-element
// `rng.uniform(0, 1)` calls do.  Seeding is init_genrand (numpy's path for an integer seed < 2^32): an inherently
// sequential 623-step recurrence, ~4 us on one lane.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/gswm.h"
#include "gswm_internal.h"
#include "gswm_math.cuh"
#include "gswm_tile.cuh"

namespace gswm {

constexpr int kMtN = 624, kMtM = 397;
constexpr int kMtPerRegen = kMtN / 2;          // doubles per regeneration
constexpr int kRing = 1024;                    // >= 255 left over + 312 new

__device__ __forceinline__ uint32_t mt_twist(uint32_t a, uint32_t b) {
  const uint32_t y = (a & 0x80000000u) | (b & 0x7FFFFFFFu);
  return (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9D2C5680u;
  y ^= (y << 15) & 0xEFC60000u;
  y ^= y >> 18;
  return y;
}

struct MtArgs {
  const uint8_t* keys;
  const uint8_t* nonces;
  const uint8_t* msgs;
  const uint32_t* seeds;       // per stream, or null: seed0 + stream
  void* out;
  int64_t n_elems;             // doubles per stream
  uint32_t seed0;
  uint32_t per_latent;
  uint32_t msg_words, tiled_words, msg_stride_bytes;
};

// kEmbed: z = Phi^-1((u + y) / 2) of stream b's uniforms into latent b (OutT fp32 / fp64); otherwise the uniforms themselves.
template <bool kEmbed, typename OutT>
__global__ void __launch_bounds__(kThreads)
mt19937_kernel(const MtArgs a) {
  __shared__ uint32_t s_mt[2][kMtN];
  __shared__ double s_ring[kRing];
  __shared__ __align__(16) uint32_t s_ks[kEmbed ? kTileWords : 4];
  const int64_t stream = blockIdx.x;
  const uint32_t tid = threadIdx.x;
  if (tid == 0) {                                                     // init_genrand(seed)
    uint32_t x = a.seeds ? a.seeds[stream] : a.seed0 + (uint32_t)stream;
    s_mt[0][0] = x;
#pragma unroll 4
    for (uint32_t i = 1; i < kMtN; ++i) {
      x = 1812433253u * (x ^ (x >> 30)) + i;
      s_mt[0][i] = x;
    }
  }
  __syncthreads();
  OutT* out = reinterpret_cast<OutT*>(a.out) + stream * a.n_elems;
  const uint8_t* s_bytes = reinterpret_cast<const uint8_t*>(s_ks);
  const int64_t row = a.per_latent ? stream : 0;
  int cur = 0;                                                        // buffer holding the previous block
  int64_t head = 0, tail = 0;                                         // doubles consumed / produced so far (uniform across the CTA)
  while (head < a.n_elems) {
    if (tail < a.n_elems) {
      const uint32_t* o = s_mt[cur];
      uint32_t* n = s_mt[cur ^ 1];
      if (tid < kMtN - kMtM) n[tid] = o[tid + kMtM] ^ mt_twist(o[tid], o[tid + 1]);                       // [0, 227)
      __syncthreads();
      if (tid < kMtN - kMtM) n[tid + 227] = n[tid] ^ mt_twist(o[tid + 227], o[tid + 228]);                // [227, 454)
      __syncthreads();
      if (tid < kMtN - 454) {                                                                              // [454, 624)
        const uint32_t i = tid + 454;
        n[i] = n[i - 227] ^ mt_twist(o[i], i + 1 < kMtN ? o[i + 1] : n[0]);
      }
      __syncthreads();
      for (uint32_t k = tid; k < kMtPerRegen; k += kThreads) {                                             // genrand_res53
        const uint32_t hi = mt_temper(n[2 * k]) >> 5, lo = mt_temper(n[2 * k + 1]) >> 6;
        s_ring[(tail + k) & (kRing - 1)] = ((double)hi * 67108864.0 + (double)lo) * (1.0 / 9007199254740992.0);
      }
      tail += kMtPerRegen;
      cur ^= 1;
      __syncthreads();
    }
    while (head < a.n_elems && (tail - head >= kThreads || tail >= a.n_elems)) {
      const int64_t e = head + tid;
      if constexpr (kEmbed) {
        if ((head % kTileElems) == 0) {                               // first batch of a tile: stage its keystream ^ message
          const uint32_t tile = (uint32_t)(head / kTileElems);
          __syncthreads();                                            // the previous tile's readers are done with s_ks
          compute_private_slice(s_ks, a.keys, a.nonces, a.msgs + row * (int64_t)a.msg_stride_bytes, row, tile,
                                tile_words(a.n_elems, tile), a.msg_words, a.tiled_words);
        }
      }
      if (e < a.n_elems && e < tail) {
        const double u = s_ring[e & (kRing - 1)];
        if constexpr (kEmbed) {
          const uint32_t et = (uint32_t)(e % kTileElems);
          const double y = (double)((s_bytes[et >> 3] >> (7 - (et & 7))) & 1u);
          out[e] = (OutT)norm_ppf_f64((u + y) / 2.0);                 // gs_insert.py:64, same roundings
        } else {
          out[e] = (OutT)u;
        }
      }
      head += kThreads;
    }
  }
}

}  // namespace gswm

using namespace gswm;

extern "C" {

int gswm_mt19937_uniform(const uint32_t* d_seeds, uint32_t seed0, int64_t n_streams, int64_t n_each, double* d_out,
                         void* stream) {
  if (!d_out) return GSWM_E_NULL;
  if (n_streams < 0 || n_each < 0 || n_streams > 0x7FFFFFFFll) return GSWM_E_SHAPE;
  if (n_streams == 0 || n_each == 0) return GSWM_OK;
  MtArgs a{};
  a.seeds = d_seeds; a.seed0 = seed0; a.out = d_out; a.n_elems = n_each;
  mt19937_kernel<false, double><<<(unsigned)n_streams, kThreads, 0, (cudaStream_t)stream>>>(a);
  count_launch();
  return (int)cudaGetLastError();
}

int gswm_embed_mt19937(const gswm_job* job, const uint32_t* d_seeds, uint32_t seed0, void* d_out, int32_t out_dtype,
                       void* stream) {
  int rc = check_job(job, false);
  if (rc) return rc;
  if (!d_out) return GSWM_E_NULL;
  if (out_dtype != GSWM_F32 && out_dtype != GSWM_F64) return GSWM_E_DTYPE;
  if (job->n_latents == 0) return GSWM_OK;
  MtArgs a{};
  a.keys = job->d_keys; a.nonces = job->d_nonces; a.msgs = job->d_msgs;
  a.seeds = d_seeds; a.seed0 = seed0; a.out = d_out; a.n_elems = job->n_elems;
  a.per_latent = (job->flags & GSWM_JOB_PER_LATENT) ? 1u : 0u;
  a.msg_words = (uint32_t)job->msg_bits / 32;
  a.tiled_words = (uint32_t)(job->n_elems / job->msg_bits) * a.msg_words;
  a.msg_stride_bytes = (uint32_t)job->msg_bits / 8;
  if (out_dtype == GSWM_F32) mt19937_kernel<true, float><<<(unsigned)job->n_latents, kThreads, 0, (cudaStream_t)stream>>>(a);
  else mt19937_kernel<true, double><<<(unsigned)job->n_latents, kThreads, 0, (cudaStream_t)stream>>>(a);
  count_launch();
  return (int)cudaGetLastError();
}

}  // extern "C"
