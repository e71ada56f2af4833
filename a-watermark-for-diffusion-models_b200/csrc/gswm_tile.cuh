// Tile-level building blocks shared by the embed / extract kernels (gswm_kernels.cu) and the MT19937 embed
// (gswm_mt19937.cu): the unit of work and the per-CTA ChaCha20 keystream staging.
//
// A TILE is 32 consecutive ChaCha20 blocks of one latent = 16384 latent elements = 2 KB of keystream = 64 KB of fp32
// latent.  One warp produces a tile's keystream with one ChaCha block per lane (no shuffles: a quarter-round is 12
// register ops); the keystream never touches HBM.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "gswm_math.cuh"

namespace gswm {

constexpr int kThreads = 256;
constexpr int kTileElems = 16384;            // 32 ChaCha blocks
constexpr int kTileWords = kTileElems / 32;  // 512 keystream words
constexpr int kTileF4 = kTileElems / 4;      // 4096 float4 per tile

__device__ __forceinline__ void load_key_nonce(const uint8_t* __restrict__ keys, const uint8_t* __restrict__ nonces,
                                               int64_t row, uint32_t (&k)[8], uint32_t (&n)[4]) {
  const uint32_t* kp = reinterpret_cast<const uint32_t*>(keys + row * 32);
  const uint32_t* np = reinterpret_cast<const uint32_t*>(nonces + row * 16);
#pragma unroll
  for (int i = 0; i < 8; ++i) k[i] = __ldg(kp + i);
#pragma unroll
  for (int i = 0; i < 4; ++i) n[i] = __ldg(np + i);
}

// Message word for keystream word index `wi` of a latent: the message tiled n_elems/msg_bits times,
// zero beyond the last whole copy (nodes.py:79-87).
__device__ __forceinline__ uint32_t tiled_msg_word(const uint8_t* __restrict__ msg, uint32_t wi, uint32_t msg_words,
                                                   uint32_t tiled_words) {
  if (msg == nullptr || wi >= tiled_words) return 0u;
  return __ldg(reinterpret_cast<const uint32_t*>(msg) + (wi % msg_words));
}

// ------------------------------------------------------------------------------------------------
// Tile keystream staging.  A CTA always computes the keystream it needs itself, into its own shared memory, one
// ChaCha20 block per lane:
//   shared key      : once per CTA, AHEAD of the grid dependency wait (key material is final before the call is
//                     enqueued, gswm.h), so that behind another kernel the predecessor's tail hides it.  An earlier
//                     version computed one table per launch in global memory and had every CTA wait on a ready flag:
//                     4.1 us of whole-GPU idle per launch (tools/embed_trace.py) against 0 (hidden) .. 2.4 us (cold).
//   per-latent keys : once per latent and tile, by warp 0 between two barriers.
// ------------------------------------------------------------------------------------------------
// lane `lane` of one warp: ChaCha block `tile*32 + lane` of stream `row`, XOR tiled message, 64 bytes to dst
__device__ __forceinline__ void chacha_tile_lane(uint32_t* __restrict__ dst, const uint8_t* __restrict__ keys,
                                                 const uint8_t* __restrict__ nonces, const uint8_t* __restrict__ msg,
                                                 int64_t row, uint32_t tile, uint32_t lane, uint32_t msg_words,
                                                 uint32_t tiled_words) {
  uint32_t k[8], n[4], ks[16];
  load_key_nonce(keys, nonces, row, k, n);
  const uint32_t blk = tile * 32 + lane;
  chacha20_block(k, n, blk, ks);
  uint4* d4 = reinterpret_cast<uint4*>(dst + lane * 16);     // lane l owns words [16 l, 16 l + 16)
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 v;
    v.x = ks[4 * q + 0] ^ tiled_msg_word(msg, blk * 16 + 4 * q + 0, msg_words, tiled_words);
    v.y = ks[4 * q + 1] ^ tiled_msg_word(msg, blk * 16 + 4 * q + 1, msg_words, tiled_words);
    v.z = ks[4 * q + 2] ^ tiled_msg_word(msg, blk * 16 + 4 * q + 2, msg_words, tiled_words);
    v.w = ks[4 * q + 3] ^ tiled_msg_word(msg, blk * 16 + 4 * q + 3, msg_words, tiled_words);
    d4[q] = v;
  }
}

// elements of tile `tile` of an n_elems-element latent (n_elems is a multiple of 4, not necessarily of the tile)
__device__ __forceinline__ uint32_t tile_elems(int64_t n_elems, uint32_t tile) {
  const int64_t remain = n_elems - (int64_t)tile * kTileElems;
  return (uint32_t)(remain < kTileElems ? remain : kTileElems);
}
// keystream words covering them (a trailing partial word / partial ChaCha block is computed whole)
__device__ __forceinline__ uint32_t tile_words(int64_t n_elems, uint32_t tile) { return (tile_elems(n_elems, tile) + 31u) >> 5; }

// Warp 0 computes tile `tile` of stream `row` in place (per-latent keys: row = latent; shared key: row = 0).
__device__ __forceinline__ void compute_private_slice(uint32_t* __restrict__ s_ks, const uint8_t* __restrict__ keys,
                                                      const uint8_t* __restrict__ nonces, const uint8_t* __restrict__ msg,
                                                      int64_t latent, uint32_t tile, uint32_t words, uint32_t msg_words,
                                                      uint32_t tiled_words) {
  if (threadIdx.x < 32 && threadIdx.x * 16 < words)
    chacha_tile_lane(s_ks, keys, nonces, msg, latent, tile, threadIdx.x, msg_words, tiled_words);
  __syncthreads();
}

}  // namespace gswm
