// ChaCha20 thread layouts on sm_100a: which one should produce a tile's keystream?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I a-watermark-for-diffusion-models_b200/csrc \
//        -o tools/chacha_layouts tools/chacha_layouts.cu && tools/chacha_layouts
//
//   lane   one block per THREAD (what libgswm uses, csrc/gswm_math.cuh:chacha20_block): the 16 state words live in the
//          thread's registers, a quarter-round is 12 register ops, no communication;
//   quad   one block per FOUR LANES, the "warp-cooperative" layout north_star sketches (a warp = 8 blocks): lane c of a
//          quad holds column c of the 4x4 state (one word of each row); the column round is 12 register ops per lane,
//          the diagonal round rotates rows 1..3 across the quad with three shuffles before and three after.
// Both are checked against each other word for word, then timed two ways:
//   throughput  n_streams x 32 blocks (BASELINE config 4: a distinct key per latent), whole GPU, blocks per second;
//   latency     ONE tile (32 blocks) by one CTA: 1 warp (lane) against 4 warps (quad) -- the prologue a CTA pays when
//               the keystream cannot be hidden behind a predecessor (cold launch).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "gswm_math.cuh"

using gswm::rotl32;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void load_kn(const uint32_t* keys, const uint32_t* nonces, int64_t s, uint32_t (&k)[8], uint32_t (&n)[4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) k[i] = __ldg(keys + s * 8 + i);
#pragma unroll
  for (int i = 0; i < 4; ++i) n[i] = __ldg(nonces + s * 4 + i);
}

// ---- lane layout: thread = (stream, block) ------------------------------------------------------------------------
template <bool kStore>
__global__ void __launch_bounds__(256) lane_kernel(const uint32_t* keys, const uint32_t* nonces, int64_t n_streams, uint32_t* out) {
  const int64_t gid = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (gid >= n_streams * 32) return;
  const int64_t s = gid >> 5;
  uint32_t k[8], n[4], ks[16];
  load_kn(keys, nonces, s, k, n);
  gswm::chacha20_block(k, n, (uint32_t)(gid & 31), ks);
  if (!kStore) {                                                      // compute only: one word per thread keeps the rounds alive
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= ks[i];
    out[gid] = acc;
    return;
  }
  uint4* dst = reinterpret_cast<uint4*>(out + gid * 16);
#pragma unroll
  for (int q = 0; q < 4; ++q) dst[q] = make_uint4(ks[4 * q], ks[4 * q + 1], ks[4 * q + 2], ks[4 * q + 3]);
}

// ---- quad layout: 4 lanes = one block; lane c holds (x[c], x[4+c], x[8+c], x[12+c]) --------------------------------
__device__ __forceinline__ void quad_block(const uint32_t* keys, const uint32_t* nonces, int64_t s, uint32_t block, uint32_t c,
                                           uint32_t& o0, uint32_t& o1, uint32_t& o2, uint32_t& o3) {
  const uint32_t sigma[4] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
  const uint32_t n0 = __ldg(nonces + s * 4), n1 = __ldg(nonces + s * 4 + 1);
  const uint32_t lo = n0 + block, hi = n1 + (lo < block ? 1u : 0u);
  const uint32_t i0 = sigma[c], i1 = __ldg(keys + s * 8 + c), i2 = __ldg(keys + s * 8 + 4 + c);
  const uint32_t i3 = c == 0 ? lo : c == 1 ? hi : __ldg(nonces + s * 4 + c);
  uint32_t a = i0, b = i1, cc = i2, d = i3;
  const uint32_t lane = threadIdx.x & 31u, base = lane & ~3u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    GSWM_QR(a, b, cc, d)                                              // column round: this lane's column
    b = __shfl_sync(0xFFFFFFFFu, b, base | ((c + 1) & 3));            // diagonalise: row 1 <<< 1, row 2 <<< 2, row 3 <<< 3
    cc = __shfl_sync(0xFFFFFFFFu, cc, base | ((c + 2) & 3));
    d = __shfl_sync(0xFFFFFFFFu, d, base | ((c + 3) & 3));
    GSWM_QR(a, b, cc, d)                                              // diagonal round
    b = __shfl_sync(0xFFFFFFFFu, b, base | ((c + 3) & 3));            // and back
    cc = __shfl_sync(0xFFFFFFFFu, cc, base | ((c + 2) & 3));
    d = __shfl_sync(0xFFFFFFFFu, d, base | ((c + 1) & 3));
  }
  o0 = a + i0; o1 = b + i1; o2 = cc + i2; o3 = d + i3;
}

template <bool kStore>
__global__ void __launch_bounds__(256) quad_kernel(const uint32_t* keys, const uint32_t* nonces, int64_t n_streams, uint32_t* out) {
  const int64_t gid = (int64_t)blockIdx.x * 256 + threadIdx.x;        // 4 threads per block (grid is a whole number of warps)
  const int64_t blk = gid >> 2;
  const bool live = blk < n_streams * 32;
  const int64_t s = live ? blk >> 5 : 0;
  const uint32_t c = (uint32_t)gid & 3u;
  uint32_t o0, o1, o2, o3;
  quad_block(keys, nonces, s, (uint32_t)(blk & 31), c, o0, o1, o2, o3);
  if (!live) return;
  if (!kStore) {
    out[gid] = o0 ^ o1 ^ o2 ^ o3;
    return;
  }
  uint32_t* dst = out + blk * 16 + c;                                 // word r*4 + c of the block: 16-byte runs per quad
  dst[0] = o0; dst[4] = o1; dst[8] = o2; dst[12] = o3;
}

// ---- one tile by one CTA, cycles from first instruction to keystream in shared memory ------------------------------
__global__ void __launch_bounds__(128) tile_latency_kernel(const uint32_t* keys, const uint32_t* nonces, int quad, uint32_t* out,
                                                           long long* cycles) {
  __shared__ uint32_t s_ks[512];
  const long long t0 = clock64();
  if (quad) {                                                         // 128 threads: 32 blocks x 4 lanes
    uint32_t o0, o1, o2, o3;
    const uint32_t blk = threadIdx.x >> 2, c = threadIdx.x & 3u;
    quad_block(keys, nonces, 0, blk, c, o0, o1, o2, o3);
    s_ks[blk * 16 + c] = o0; s_ks[blk * 16 + 4 + c] = o1; s_ks[blk * 16 + 8 + c] = o2; s_ks[blk * 16 + 12 + c] = o3;
  } else if (threadIdx.x < 32) {                                      // one warp: 32 blocks x 1 lane
    uint32_t k[8], n[4], ks[16];
    load_kn(keys, nonces, 0, k, n);
    gswm::chacha20_block(k, n, threadIdx.x, ks);
#pragma unroll
    for (int i = 0; i < 16; ++i) s_ks[threadIdx.x * 16 + i] = ks[i];
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) *cycles = t1 - t0;
  for (int i = threadIdx.x; i < 512; i += 128) out[i] = s_ks[i];
}

int main() {
  const int64_t n_streams = 65536;                                    // 2 M blocks = 128 MB of keystream
  const int64_t blocks = n_streams * 32;
  std::vector<uint32_t> hk(n_streams * 8), hn(n_streams * 4);
  uint64_t x = 0x9E3779B97F4A7C15ull;
  auto next = [&] { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return (uint32_t)(x >> 16); };
  for (auto& v : hk) v = next();
  for (auto& v : hn) v = next();
  for (int i = 0; i < 8; ++i) { hn[4 * i] = 0xFFFFFFF0u + i; hn[4 * i + 1] = 0xFFFFFFFFu; }   // counter carries
  uint32_t *dk, *dn, *da, *db;
  long long* dc;
  CK(cudaMalloc(&dk, hk.size() * 4)); CK(cudaMalloc(&dn, hn.size() * 4));
  CK(cudaMalloc(&da, blocks * 64)); CK(cudaMalloc(&db, blocks * 64)); CK(cudaMalloc(&dc, 8));
  CK(cudaMemcpy(dk, hk.data(), hk.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dn, hn.data(), hn.size() * 4, cudaMemcpyHostToDevice));
  const unsigned g_lane = (unsigned)((blocks + 255) / 256), g_quad = (unsigned)((blocks * 4 + 255) / 256);
  lane_kernel<true><<<g_lane, 256>>>(dk, dn, n_streams, da);
  quad_kernel<true><<<g_quad, 256>>>(dk, dn, n_streams, db);
  CK(cudaDeviceSynchronize());
  std::vector<uint32_t> ha(blocks * 16), hb(blocks * 16);
  CK(cudaMemcpy(ha.data(), da, blocks * 64, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hb.data(), db, blocks * 64, cudaMemcpyDeviceToHost));
  int64_t bad = 0;
  for (size_t i = 0; i < ha.size(); ++i) bad += ha[i] != hb[i];
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms[4] = {1e9f, 1e9f, 1e9f, 1e9f};                             // lane, quad with stores; lane, quad compute only
  for (int rep = 0; rep < 5; ++rep) {
    for (int which = 0; which < 4; ++which) {
      CK(cudaEventRecord(e0));
      for (int i = 0; i < 10; ++i) {
        if (which == 0) lane_kernel<true><<<g_lane, 256>>>(dk, dn, n_streams, da);
        else if (which == 1) quad_kernel<true><<<g_quad, 256>>>(dk, dn, n_streams, db);
        else if (which == 2) lane_kernel<false><<<g_lane, 256>>>(dk, dn, n_streams, da);
        else quad_kernel<false><<<g_quad, 256>>>(dk, dn, n_streams, db);
      }
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float t;
      CK(cudaEventElapsedTime(&t, e0, e1));
      if (t / 10 < ms[which]) ms[which] = t / 10;
    }
  }
  long long cyc[2] = {1 << 30, 1 << 30};
  for (int rep = 0; rep < 20; ++rep) {
    for (int which = 0; which < 2; ++which) {
      long long c;
      tile_latency_kernel<<<1, 128>>>(dk, dn, which, da, dc);
      CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
      if (rep >= 2 && c < cyc[which]) cyc[which] = c;
    }
  }
  printf("{\"blocks\": %lld, \"mismatching_words\": %lld, \"lane_ms\": %.4f, \"quad_ms\": %.4f, \"lane_Gblocks_per_s\": %.2f, "
         "\"quad_Gblocks_per_s\": %.2f, \"lane_keystream_GBps\": %.1f, \"quad_keystream_GBps\": %.1f, \"lane_compute_only_ms\": %.4f, \"quad_compute_only_ms\": %.4f, "
         "\"tile_latency_cycles_lane_1warp\": %lld, \"tile_latency_cycles_quad_4warps\": %lld}\n",
         (long long)blocks, (long long)bad, ms[0], ms[1], blocks / ms[0] / 1e6, blocks / ms[1] / 1e6, blocks * 64 / ms[0] / 1e6,
         blocks * 64 / ms[1] / 1e6, ms[2], ms[3], cyc[0], cyc[1]);
  return bad != 0;
}
