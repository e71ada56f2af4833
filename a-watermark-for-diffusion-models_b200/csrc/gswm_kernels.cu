// libgswm device entry points: ChaCha20 keystream (K1), embed (K2), extract (K3) for sm_100a.
//
// Work decomposition shared by K2/K3: a TILE is 32 consecutive ChaCha20 blocks of one latent =
// 16384 latent elements = 2 KB of keystream = 64 KB of fp32 latent.  One warp produces a tile's
// keystream with one ChaCha block per lane (no shuffles: a quarter-round is 12 register ops); the
// CTA's 256 threads then stream the tile as 16 fully coalesced 128-bit accesses per thread.
//
// HBM layout: latents are [n_latents][n_elems] contiguous (C-order (B,4,H/8,W/8)), 16-byte aligned;
// keys [.][32], nonces [.][16], messages [.][msg_bits/8] are byte arrays.  Keystream never touches HBM.
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include <atomic>
#include <cstdint>
#include <cstdlib>

#include "../../include/gswm.h"
#include "gswm_comm.cuh"
#include "gswm_internal.h"
#include "gswm_math.cuh"
#include "gswm_tile.cuh"

namespace gswm {

static std::atomic<int64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// Reference extract.py:83: int(norm.cdf(z) * 2) == 1  <=>  z >= -6.957291061679417e-17 (float64).
// Smallest fp32 that is >= that double (0xA4A06C98); bf16 inputs use its truncation 0xA4A0, fp16 inputs +0 (no fp16
// value other than -0.0 lies in [-6.96e-17, 0)) -- see NegatedBits16.
__device__ __forceinline__ float quantise_threshold() { return __uint_as_float(0xA4A06C98u); }

// Programmatic dependent launch (PDL): both throughput kernels are launched with the stream-serialisation attribute
// relaxed, tell the hardware at once that the NEXT kernel in the stream may start being scheduled as their CTAs
// retire, and do everything that touches global memory only after griddep_wait() -- which returns when every
// earlier kernel in the stream has completed and flushed.  What overlaps with the predecessor's tail is the launch
// latency and the CTA prologue (sign table, barrier init, counter reset): ~2-3 us per kernel of a ~60 us step.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// K1: keystream (optionally XOR tiled message) to global memory, one ChaCha block per thread.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
chacha20_keystream_kernel(const uint8_t* __restrict__ keys, const uint8_t* __restrict__ nonces,
                          const uint8_t* __restrict__ msgs, int64_t n_streams, uint32_t blocks_each,
                          uint32_t msg_words, uint32_t tiled_words, uint32_t msg_stride_bytes,
                          uint32_t* __restrict__ out) {
  const int64_t gid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (gid >= n_streams * (int64_t)blocks_each) return;
  const int64_t s = gid / blocks_each;
  const uint32_t blk = (uint32_t)(gid - s * blocks_each);
  uint32_t k[8], n[4], ks[16];
  load_key_nonce(keys, nonces, s, k, n);
  chacha20_block(k, n, blk, ks);
  const uint8_t* msg = msgs ? msgs + s * (int64_t)msg_stride_bytes : nullptr;
  uint4* dst = reinterpret_cast<uint4*>(out + gid * 16);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 v;
    v.x = ks[4 * q + 0] ^ tiled_msg_word(msg, blk * 16 + 4 * q + 0, msg_words, tiled_words);
    v.y = ks[4 * q + 1] ^ tiled_msg_word(msg, blk * 16 + 4 * q + 1, msg_words, tiled_words);
    v.z = ks[4 * q + 2] ^ tiled_msg_word(msg, blk * 16 + 4 * q + 2, msg_words, tiled_words);
    v.w = ks[4 * q + 3] ^ tiled_msg_word(msg, blk * 16 + 4 * q + 3, msg_words, tiled_words);
    dst[q] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// K2: embed.  grid = n_latents * tiles_per_latent CTAs, one tile each.
//   bits  <- keystream XOR tiled message                                (gs_insert.py:23,45-49)
//   u     <- 23 bits of Philox4x32 per element (3 calls feed 16 elements) (stands in for gs_insert.py:62)
//   z     <- Phi^-1((u + y)/2) = +-g(v) via bucket_quantile4_f32          (gs_insert.py:64)
//   store <- one 128-bit store per 4 elements, C-order flat index       (gs_insert.py:65)
// ------------------------------------------------------------------------------------------------
struct EmbedArgs {
  const uint8_t* keys;
  const uint8_t* nonces;
  const uint8_t* msgs;
  float* out;
  int64_t n_elems;
  int64_t n_latents;
  int64_t first_latent;      // global index of latent 0 (sharding)
  uint32_t tiles_per_latent;
  uint32_t msg_words;
  uint32_t tiled_words;
  uint32_t msg_stride_bytes;
  uint32_t seed_lo, seed_hi, off_lo, off_hi;   // off_hi holds (offset_hi << 2): its low 2 bits select the Philox call
  uint32_t keys_in_flight;   // GSWM_JOB_KEYS_IN_FLIGHT: read key material only behind the grid dependency wait
  PhiloxKeys rk;             // round keys of (seed_lo, seed_hi)
  PhiloxLaunch pl;           // round-1 products of the v4 counter layout (constants of the launch)
};

// Staging for the injected-uniform kernel, grid = (n_latents, tiles_per_latent): blockIdx.x = latent, blockIdx.y = tile.
template <bool kPerLatent>
__device__ __forceinline__ void embed_stage(uint32_t* s_ks, const EmbedArgs& a, int64_t latent, uint32_t tile, uint32_t words) {
  const int64_t row = kPerLatent ? latent : 0;
  compute_private_slice(s_ks, a.keys, a.nonces, a.msgs + row * (int64_t)a.msg_stride_bytes, row, tile, words,
                        a.msg_words, a.tiled_words);
}

// Sign look-up: for a byte of bucket bits, +-1.0f for the four elements of its high nibble (even float4)
// and of its low nibble (odd float4); +1 where the bucket bit is 1.  256 entries x 2 x float4 = 8 KB.
__device__ __forceinline__ void build_sign_lut(float4* lut) {
  const uint32_t b = threadIdx.x;                  // kThreads == 256: one byte value per thread
  lut[2 * b + 0] = make_float4((b & 0x80u) ? 1.f : -1.f, (b & 0x40u) ? 1.f : -1.f, (b & 0x20u) ? 1.f : -1.f, (b & 0x10u) ? 1.f : -1.f);
  lut[2 * b + 1] = make_float4((b & 0x08u) ? 1.f : -1.f, (b & 0x04u) ? 1.f : -1.f, (b & 0x02u) ? 1.f : -1.f, (b & 0x01u) ? 1.f : -1.f);
}

#ifndef GSWM_TOPCELL
#define GSWM_TOPCELL 1      // 1: the outermost grid cell is refined (since uniforms v3); 0: diagnostic build without (the cell's midpoint)
#endif
struct TopCellShared {      // the CTA's log of outermost-cell elements (gswm_math.cuh: TopCellLog)
  uint2 list[kTopCellSlots];
  uint32_t count;
};
#ifndef GSWM_PHILOX_CONST_KEYS
#define GSWM_PHILOX_CONST_KEYS 1
#endif
#ifndef GSWM_CTR_V4
#define GSWM_CTR_V4 1       // 1: lane index in counter word 1 (uniforms v4, gswm_math.cuh: philox4x32_v4_calls3); 0: in word 0 (v3)
#endif
// v4 counters: the part of a super-iteration's three Philox calls that does not depend on the lane, tabulated per CTA.
//   shared key : entry ((it % kUniIters) * 2 + (sidx - s0)) * 3 + call for the CTA's it-th latent; refilled every kUniIters
//                latents (one barrier pair per 42 latents; the first fill happens in the prologue, ahead of the grid
//                dependency wait).
//   per-latent : entry sidx * 3 + call of the CTA's current latent, filled next to its keystream.
constexpr uint32_t kUniIters = 42;
constexpr uint32_t kUniEntries = kUniIters * 2 * 3;                   // 252 entries, one thread each
static_assert(kUniEntries <= kThreads && kUniEntries >= 12, "one thread per table entry");

// One SUPER-ITERATION of a tile: 1024 float4 = 4 per thread, fed by three Philox calls.
// Philox counter word 0..1: G = ((global_latent * tiles + tile) * 4 + sidx) * 256 + tid; words 2..3: offset, call index.
// kHoisted (shared key, whole tile): the bucket nibbles of this thread's four float4 never change from latent to latent, so
// the caller passes them pre-scaled (nibble << 4 in byte k of `sign_pack`) and `my_sign` is the 16-entry nibble table:
// one PRMT + one LDS.128 per float4 instead of LDS.U8 + LEA + LDS.128 from the 8 KB byte table.
template <bool kGuard, bool kHoisted = false>
__device__ __forceinline__ void embed_super_iteration(const EmbedArgs& a, const uint8_t* __restrict__ s_bytes,
                                                      const float4* __restrict__ my_sign, float4* __restrict__ out4,
                                                      uint64_t g_tile, uint32_t sidx, uint32_t n_f4, TopCellShared* s_top,
                                                      const uint4* __restrict__ uni, uint32_t latent_rel, uint32_t sign_pack = 0) {
#if GSWM_CTR_V4
  // g_tile = (global latent * tiles + tile) * 4, warp-uniform; T = g_tile + sidx < 2^54 (checked by the host)
  const uint64_t T = g_tile + sidx;
  uint4 cc[3];
  philox4x32_v4_calls3(cc, threadIdx.x | ((uint32_t)T << 8), uni, a.pl, a.rk);
  const uint4 c0 = cc[0], c1 = cc[1], c2 = cc[2];
#else
  const uint64_t g = g_tile + (uint64_t)sidx * kThreads;
  const uint32_t glo = (uint32_t)g, ghi = (uint32_t)(g >> 32);
#if GSWM_PHILOX_CONST_KEYS
  const uint4 c0 = philox4x32_keys(make_uint4(glo, ghi, a.off_lo, a.off_hi + 0u), a.rk);
  const uint4 c1 = philox4x32_keys(make_uint4(glo, ghi, a.off_lo, a.off_hi + 1u), a.rk);
  const uint4 c2 = philox4x32_keys(make_uint4(glo, ghi, a.off_lo, a.off_hi + 2u), a.rk);
#else
  const uint4 c0 = philox4x32(make_uint4(glo, ghi, a.off_lo, a.off_hi + 0u), a.seed_lo, a.seed_hi);
  const uint4 c1 = philox4x32(make_uint4(glo, ghi, a.off_lo, a.off_hi + 1u), a.seed_lo, a.seed_hi);
  const uint4 c2 = philox4x32(make_uint4(glo, ghi, a.off_lo, a.off_hi + 2u), a.seed_lo, a.seed_hi);
#endif
#endif
  const uint32_t i0 = (4u * sidx) * kThreads + threadIdx.x;
  const uint32_t f0[4] = {fbits_top23(c0.x), fbits_top23(c0.y), fbits_top23(c0.z), fbits_top23(c0.w)};
  const uint32_t f1[4] = {fbits_top23(c1.x), fbits_top23(c1.y), fbits_top23(c1.z), fbits_top23(c1.w)};
  const uint32_t f2[4] = {fbits_top23(c2.x), fbits_top23(c2.y), fbits_top23(c2.z), fbits_top23(c2.w)};
  const uint32_t f3[4] = {fbits_low_bytes(c0.x, c1.x, c2.x), fbits_low_bytes(c0.y, c1.y, c2.y),
                          fbits_low_bytes(c0.z, c1.z, c2.z), fbits_low_bytes(c0.w, c1.w, c2.w)};
  auto emit = [&](uint32_t i, const uint32_t (&f)[4], uint32_t k) {
    if (kGuard && i >= n_f4) return;
#ifdef GSWM_WHATIF_NOSIGN
    const float4 sgn = make_float4(1.f, 1.f, 1.f, 1.f);                // diagnostic build: no bucket bits, no LUT (gswm_math.cuh)
#else
    const float4 sgn = kHoisted ? *reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(my_sign) +
                                                                    __byte_perm(sign_pack, 0u, 0x4440u | k))
                                : my_sign[2u * s_bytes[i >> 1]];
#endif
    // an element in the outermost grid cell is logged here and refined in the kernel's epilogue
#if GSWM_TOPCELL
    const TopCellLog top{&s_top->count, s_top->list, latent_rel, i};
#else
    const NoTopCell top;                                               // diagnostic build (uniforms v2: the cell's midpoint)
#endif
#ifdef GSWM_WHATIF_NOSTORE
    const float4 zq = bucket_quantile4_f32<!kGuard>(f[0], f[1], f[2], f[3], sgn, top);
    if (zq.x == 123.456f) __stcs(out4 + i, zq);                        // diagnostic build: arithmetic kept alive, nothing written
#else
    __stcs(out4 + i, bucket_quantile4_f32<!kGuard>(f[0], f[1], f[2], f[3], sgn, top));   // !kGuard: every lane of the warp is here
#endif
  };
  emit(i0, f0, 0);
  emit(i0 + kThreads, f1, 1);
  emit(i0 + 2 * kThreads, f2, 2);
  emit(i0 + 3 * kThreads, f3, 3);
}

// Grid (X, Y).  CTA (x, y) owns one fixed piece of the latent layout -- so its keystream, its sign table and its Philox
// offset are set up ONCE -- and walks the latents x, x + X, x + 2X, ...
//   shared key : Y = non-empty HALF tiles of a latent; CTA (x, h) produces super-iterations {2(h&1), 2(h&1)+1} of tile
//                h>>1 (two independent instruction streams).  X * Y is sized to the CTAs the GPU holds at once
//                (persistent); nothing in the latent loop synchronises, warps drift freely.  Every CTA computes the 16
//                ChaCha20 blocks of ITS half tile itself, into its own shared memory, ahead of the grid dependency wait:
//                there is no cross-CTA dependency of any kind (see "Tile keystream staging" above).
//   per-latent : Y = tiles, X = n_latents (one latent per CTA): warp 0 computes the tile's keystream between two barriers
//                while the SM's other CTAs keep its issue slots busy.
// Resident CTAs per SM.  Shared key: 4 (64 registers: two super-iterations in flight without spills; measured 69.7 us per
// 4096 SD-2.1 latents against 73.6 at 6 x 40 registers).  Per-latent keys: 6, one latent per CTA -- while warp 0 computes a
// tile's ChaCha20 blocks the other CTAs of the SM keep its issue slots busy, so more and smaller CTAs win there.
#ifndef GSWM_EMBED_MINB
#define GSWM_EMBED_MINB 4
#endif
#ifndef GSWM_PL_UNROLL
#define GSWM_PL_UNROLL 2    // super-iterations of a per-latent-key CTA unrolled together
#endif
constexpr int kPlUnroll = GSWM_PL_UNROLL;
#ifndef GSWM_EMBED_MINB_PER_LATENT
#define GSWM_EMBED_MINB_PER_LATENT 6
#endif
// -DGSWM_TRACE: per-CTA timeline of the embed kernel (globaltimer ns at entry, after the grid dependency wait, after the
// keystream slice is staged, at exit), read back with gswm_debug_trace_read -- tools/embed_trace.py.
#ifdef GSWM_TRACE
__device__ unsigned long long g_trace[2 * 4 * 8192];                  // two buffers: consecutive launches alternate
__device__ unsigned g_trace_buf;
__device__ __forceinline__ void trace_mark(int slot) {
  if (threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x;
    if (cta < 8192) g_trace[(g_trace_buf & 1u) * 4 * 8192 + 4 * cta + slot] = t;
  }
}
#else
__device__ __forceinline__ void trace_mark(int) {}
#endif

template <bool kPerLatent>
__global__ void __launch_bounds__(kThreads, kPerLatent ? GSWM_EMBED_MINB_PER_LATENT : GSWM_EMBED_MINB)
embed_kernel(const EmbedArgs a) {
  __shared__ __align__(16) uint32_t s_ks[kTileWords];
  __shared__ __align__(16) float4 s_sign[512];
  __shared__ __align__(16) float4 s_sign16[16];                       // nibble -> +-1.0f x4 (shared-key whole-tile path)
  __shared__ __align__(16) TopCellShared s_top;                       // outermost-cell elements seen by this CTA (refined in the epilogue)
#if GSWM_CTR_V4
  __shared__ __align__(16) uint4 s_uni[kUniEntries];                  // lane-independent halves of Philox rounds 1..3 (philox_v4_cta_part)
#else
  uint4* const s_uni = nullptr;
#endif
  const uint32_t tile = kPerLatent ? blockIdx.y : blockIdx.y >> 1;
  const uint32_t tiles = a.tiles_per_latent;
  const uint32_t words = tile_words(a.n_elems, tile);
  const uint32_t n_f4 = tile_elems(a.n_elems, tile) >> 2;
  trace_mark(0);
  griddep_launch_dependents();
  build_sign_lut(s_sign);
  if (threadIdx.x == 0) s_top.count = 0;
  if (threadIdx.x < 16) {
    const uint32_t n = threadIdx.x;
    s_sign16[n] = make_float4((n & 8u) ? 1.f : -1.f, (n & 4u) ? 1.f : -1.f, (n & 2u) ? 1.f : -1.f, (n & 1u) ? 1.f : -1.f);
  }
  if constexpr (!kPerLatent) {
    // Shared key: this CTA's half tile needs 16 ChaCha20 blocks, once.  Half a warp computes them here, AHEAD of the
    // grid dependency wait -- key material is final before the call is enqueued (gswm.h) -- so when the previous kernel
    // in the stream is still draining, its tail hides this prologue; no table in global memory, no flag to spin on.
    // (GSWM_JOB_KEYS_IN_FLIGHT: the key material may still be being written -- same thing, behind the wait.)
    const uint32_t lane = threadIdx.x;
    const bool mine = lane < 32 && (lane >> 4) == (blockIdx.y & 1u) && lane * 16 < words;
    if (mine && !a.keys_in_flight) chacha_tile_lane(s_ks, a.keys, a.nonces, a.msgs, 0, tile, lane, a.msg_words, a.tiled_words);
    if (a.keys_in_flight) {
      griddep_wait();
      if (mine) chacha_tile_lane(s_ks, a.keys, a.nonces, a.msgs, 0, tile, lane, a.msg_words, a.tiled_words);
    }
    __syncthreads();                                                  // LUT + keystream visible
  }
  griddep_wait();                                                     // nothing above writes global memory or reads a predecessor's output
  trace_mark(1);
  const uint8_t* s_bytes = reinterpret_cast<const uint8_t*>(s_ks);
  const float4* my_sign = s_sign + (threadIdx.x & 1u);               // i & 1 == threadIdx.x & 1 for every float4
#if GSWM_CTR_V4
  const uint64_t tile_stride = 4ull;                                  // super-iterations per tile (the lane index lives in another counter word)
#else
  const uint64_t tile_stride = 4ull * kThreads;                       // Philox counters per tile
#endif
  const uint64_t n_f4_latent = (uint64_t)(a.n_elems >> 2);
  float4* out4 = reinterpret_cast<float4*>(a.out) + (blockIdx.x * n_f4_latent + (uint64_t)tile * kTileF4);
#if GSWM_CTR_V4
  uint64_t g_tile = ((uint64_t)(a.first_latent + blockIdx.x) * tiles + tile) * tile_stride;
#else
  uint64_t g_tile = ((uint64_t)(a.first_latent + blockIdx.x) * tiles + tile) * tile_stride + threadIdx.x;
#endif
  const uint64_t out_step = gridDim.x * n_f4_latent;
  const uint64_t g_step = (uint64_t)gridDim.x * tiles * tile_stride;

  if constexpr (!kPerLatent) {
    trace_mark(2);
    const uint32_t s0 = (blockIdx.y & 1u) * 2u;                       // the launch only creates non-empty halves
    // Static split: CTA (x, h) walks latents x, x + X, ...  (A dynamic split through per-column atomic counters was tried,
    // because the warp schedulers share an SM unfairly and the CTAs of one launch finish anywhere between 27 and 60 us
    // -- tools/embed_trace.py -- but same-address atomics under this kernel's store traffic complete only every ~34 ns,
    // far too slow for one grab per latent: 67.6 us against 56.1 us.)
#if GSWM_CTR_V4
    // entry e = (it * 2 + sd) * 3 + call of the table, for the kUniIters latents starting at the CTA's `it0`-th
    auto fill_uni = [&](uint64_t g_tile_first) {
      const uint32_t e = threadIdx.x;
      if (e < kUniEntries) {
        const uint64_t T = g_tile_first + (uint64_t)(e / 6u) * g_step + s0 + (e % 6u) / 3u;
        s_uni[e] = philox_v4_cta_part(((uint32_t)(T >> 24) << 2) | (e % 3u), a.pl, a.rk);
      }
    };
    fill_uni(g_tile);                                                 // (registers and kernel parameters only: could sit ahead of the wait)
    __syncthreads();
    uint32_t it = 0;
#define GSWM_UNI_STEP()                                                                                    \
    if (++it == kUniIters) { __syncthreads(); fill_uni(g_tile + g_step); __syncthreads(); it = 0; }
#define GSWM_UNI(sd) (s_uni + (it * 2u + (sd)) * 3u)
#else
#define GSWM_UNI_STEP()
#define GSWM_UNI(sd) nullptr
#endif
    if (n_f4 == kTileF4) {
      // this thread's eight bucket nibbles (float4 (4 sidx + k) 256 + tid, sidx in {s0, s0 + 1}) are the same for every
      // latent: fetched once, kept as table offsets (nibble << 4) in the bytes of two registers
      uint32_t pack[2] = {0u, 0u};
#pragma unroll
      for (uint32_t j = 0; j < 8; ++j) {
        const uint32_t i = (4u * (s0 + (j >> 2)) + (j & 3u)) * kThreads + threadIdx.x;
        const uint32_t byte = s_bytes[i >> 1];
        const uint32_t nib = (threadIdx.x & 1u) ? (byte & 0xFu) : (byte >> 4);
        pack[j >> 2] |= (nib << 4) << (8u * (j & 3u));
      }
      for (int64_t latent = blockIdx.x; latent < a.n_latents; latent += gridDim.x, out4 += out_step, g_tile += g_step) {
        embed_super_iteration<false, true>(a, s_bytes, s_sign16, out4, g_tile, s0, n_f4, &s_top, GSWM_UNI(0), (uint32_t)latent, pack[0]);
        embed_super_iteration<false, true>(a, s_bytes, s_sign16, out4, g_tile, s0 + 1, n_f4, &s_top, GSWM_UNI(1), (uint32_t)latent, pack[1]);
        GSWM_UNI_STEP()
      }
    } else {
      for (int64_t latent = blockIdx.x; latent < a.n_latents; latent += gridDim.x, out4 += out_step, g_tile += g_step) {
        embed_super_iteration<true>(a, s_bytes, my_sign, out4, g_tile, s0, n_f4, &s_top, GSWM_UNI(0), (uint32_t)latent);
        if ((s0 + 1) * 4 * kThreads < n_f4) embed_super_iteration<true>(a, s_bytes, my_sign, out4, g_tile, s0 + 1, n_f4, &s_top, GSWM_UNI(1), (uint32_t)latent);
        GSWM_UNI_STEP()
      }
    }
#undef GSWM_UNI_STEP
#undef GSWM_UNI
    trace_mark(3);
  } else {
    for (int64_t latent = blockIdx.x; latent < a.n_latents; latent += gridDim.x, out4 += out_step, g_tile += g_step) {
      __syncthreads();                                                // previous latent's readers are done with s_ks (first pass: LUT built)
#if GSWM_CTR_V4
      if (threadIdx.x >= 32 && threadIdx.x < 44) {                    // warp 1, next to warp 0's ChaCha20: entry sidx * 3 + call
        const uint32_t e = threadIdx.x - 32u;
        const uint64_t T = g_tile + e / 3u;
        s_uni[e] = philox_v4_cta_part(((uint32_t)(T >> 24) << 2) | (e % 3u), a.pl, a.rk);
      }
#endif
      compute_private_slice(s_ks, a.keys, a.nonces, a.msgs + latent * (int64_t)a.msg_stride_bytes, latent, tile, words,
                            a.msg_words, a.tiled_words);
      if (n_f4 == kTileF4) {
#pragma unroll kPlUnroll
        for (uint32_t sidx = 0; sidx < 4; ++sidx) embed_super_iteration<false>(a, s_bytes, my_sign, out4, g_tile, sidx, n_f4, &s_top, s_uni + 3u * sidx, (uint32_t)latent);
      } else {
        for (uint32_t sidx = 0; sidx * 4 * kThreads < n_f4; ++sidx) embed_super_iteration<true>(a, s_bytes, my_sign, out4, g_tile, sidx, n_f4, &s_top, s_uni + 3u * sidx, (uint32_t)latent);
      }
    }
  }
#if GSWM_TOPCELL
  // Epilogue: refine the outermost-cell elements this CTA logged -- one in 8.4 M, so almost always none.
  __syncthreads();                                                    // all stores and log entries of the CTA are done and visible to it
  const uint32_t logged = s_top.count < kTopCellSlots ? s_top.count : kTopCellSlots;
  for (uint32_t e = threadIdx.x; e < logged; e += kThreads) {
    const uint64_t lat = s_top.list[e].x;                             // latent within this launch
    const uint32_t t = tile, f4 = s_top.list[e].y;                    // tile; float4 (4 sidx + k) * 256 + tid within the tile
    float4* where = reinterpret_cast<float4*>(a.out) + (lat * n_f4_latent + (uint64_t)t * kTileF4 + f4);
    const uint32_t sk = f4 / kThreads, tid = f4 % kThreads;
#if GSWM_CTR_V4
    const uint64_t T = (((uint64_t)a.first_latent + lat) * tiles + t) * 4u + (sk >> 2);
    uint4 ctr = make_uint4(a.off_lo, tid | ((uint32_t)T << 8), a.off_hi, (uint32_t)(T >> 24) << 2);
#else
    const uint64_t g = (((uint64_t)a.first_latent + lat) * tiles + t) * tile_stride + (uint64_t)(sk >> 2) * kThreads + tid;
    uint4 ctr = make_uint4((uint32_t)g, (uint32_t)(g >> 32), a.off_lo, a.off_hi);
#endif
    uint4 w[3];
#pragma unroll 1
    for (uint32_t c = 0; c < 3; ++c) w[c] = philox4x32(make_uint4(ctr.x, ctr.y, ctr.z, ctr.w + c), a.seed_lo, a.seed_hi);
    ctr.w += 3u;
    top_cell_fixup(where, w, sk & 3u, ctr, a.seed_lo, a.seed_hi);
  }
#endif
}

// Injected-uniform embed (fp64 arithmetic, parity / seeded drop-in mode; not the throughput path).
template <bool kPerLatent, typename OutT>
__global__ void __launch_bounds__(kThreads)
embed_injected_kernel(const EmbedArgs a, const double* __restrict__ u, int u_per_latent, OutT* __restrict__ out) {
  __shared__ __align__(16) uint32_t s_ks[kTileWords];
  const int64_t latent = blockIdx.x;
  const uint32_t tile = blockIdx.y;
  const int64_t tile_base = (int64_t)tile * kTileElems;
  const uint32_t words = tile_words(a.n_elems, tile);
  const uint32_t n_el = tile_elems(a.n_elems, tile);
  embed_stage<kPerLatent>(s_ks, a, latent, tile, words);
  const uint8_t* s_bytes = reinterpret_cast<const uint8_t*>(s_ks);
  const double* up = u + (u_per_latent ? latent * a.n_elems : 0) + tile_base;
  OutT* op = out + latent * a.n_elems + tile_base;
  // gridDim.z CTAs share a tile (each stages the tile's keystream itself: 2.4 us on one warp, in parallel): a single
  // latent -- every call of the reference-named drop-ins -- is 16 CTAs instead of one doing 16 384 float64 quantiles
  // (~45 us of a 128 us call, tools/dropin_breakdown.py)
  const uint32_t per = (n_el + gridDim.z - 1) / gridDim.z;
  const uint32_t lo = blockIdx.z * per, hi = lo + per < n_el ? lo + per : n_el;
  for (uint32_t e = lo + threadIdx.x; e < hi; e += kThreads) {
    const double y = (double)((s_bytes[e >> 3] >> (7 - (e & 7))) & 1u);
    const double p = (up[e] + y) / 2.0;                              // gs_insert.py:64, same roundings
    op[e] = (OutT)norm_ppf_f64(p);
  }
}

// ------------------------------------------------------------------------------------------------
// K3: extract.  Persistent grid (a few CTAs per SM); CTA c decodes latents c, c + grid, c + 2 grid, ...
//   bit    <- z >= threshold                                             (extract.py:82-84)
//   bit    ^= keystream bit                                              (extract.py:86-87)
//   count  <- per message position over the R copies                     (extract.py:91-98)
//   msg    <- count > R/2  (strict; tie -> 0)                            (extract.py:99)
//   score  <- popc(~(msg ^ reference))                                   (extract.py:103-109)
//
// The latents are streamed HBM -> shared memory by the TMA engine (cp.async.bulk, 1-D, completion on
// an mbarrier) in CHUNKS of 32 KB (8192 fp32 / 16384 fp16 elements) through a kStages-deep ring, issued by one thread kStages-1
// chunks ahead of the consumers -- across latent boundaries, so the memory system never drains while a
// latent's vote is being finalised.  The 256 threads read their four 4-element groups of the chunk with
// conflict-free 128-bit shared loads.
//
// Fast path (kPow2): msg_bits divides 1024, so a thread's four positions never change and the four
// counts ride in one register as byte lanes.  General path: shared-memory atomics per set bit.
// ------------------------------------------------------------------------------------------------
struct ExtractArgs {
  const uint8_t* keys;
  const uint8_t* nonces;
  const uint8_t* msgs;        // reference messages (may be null)
  const void* z;
  uint8_t* msg_out;
  uint16_t* counts;
  int32_t* matched;
  unsigned long long* counters;
  int64_t n_elems;
  int64_t n_latents;
  uint32_t tiles_per_latent;
  uint32_t chunks_per_latent;
  uint32_t msg_bits;
  uint32_t msg_stride_bytes;
  uint32_t copies;            // R = n_elems / msg_bits
  uint32_t ks_cache_tiles;    // shared-key mode: tiles of keystream kept resident in shared memory (0 = restage per tile)
  uint8_t* flags;             // per-latent GSWM_FLAG_* (may be null)
  uint32_t keys_in_flight;    // GSWM_JOB_KEYS_IN_FLIGHT: read key material only behind the grid dependency wait
  // gswm_extract_allreduce: the last CTA to retire sums `counters` over the ranks of `comm` into `reduced`
  long long* reduced;         // null = plain extract
  CommDev comm;
};

#ifndef GSWM_CHUNK_BYTES
#define GSWM_CHUNK_BYTES 32768
#endif
#ifndef GSWM_EXTRACT_MINB
#define GSWM_EXTRACT_MINB 3
#endif
// A chunk is a fixed number of BYTES (what the memory system has in flight per CTA is what matters): 8192 fp32 or
// 16384 fp16 / bf16 elements -- half a tile or a whole tile.
// Ring depth: fp32 input 3 stages (96 KB -> 2 CTAs per SM), 16-bit input 2 stages (64 KB -> 3 CTAs per SM).  Alone, both
// shapes keep ~192 KB per SM in flight and fp32 runs within 2 % of each other (41.4 / 42.2 us); 16-bit input is bound by
// the consumer warps and wants the third CTA (22.5 against 25.1 us), and so do per-latent keys, where warp 0's ChaCha20
// rounds stall a CTA once per tile (45.3 against 52.3 us).  But next to the embed kernel (the co-scheduled
// step, bench.py) only one or two extract CTAs find room on an SM, and then the bytes EACH of them keeps in flight
// decide how much of the idle HBM bandwidth gets used: 87.2 us per step with 3 stages against 93.1 us with 2
// (tools/cobench.py).  -DGSWM_STAGES=n overrides both.
constexpr int kStageBytes = GSWM_CHUNK_BYTES;
template <typename T>
struct Chunk {
  static constexpr int kElems = kStageBytes / (int)sizeof(T);
  static constexpr int kPerTile = kTileElems / kElems;
  static_assert(kPerTile >= 1 && kPerTile * kElems == kTileElems, "a chunk must divide a tile");
};
template <typename T, bool kPerLatent>
struct Ring {
#ifdef GSWM_STAGES
  static constexpr int kStages = GSWM_STAGES;
#else
  static constexpr int kStages = (sizeof(T) == 4 && !kPerLatent) ? 3 : 2;
#endif
  static constexpr int kMinBlocks = kStages * kStageBytes > 72 * 1024 ? 2 : GSWM_EXTRACT_MINB;
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on `bar`; streaming data: evict-first in L2
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_addr(dst_smem)), "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy) : "memory");
}

// Quantise the four elements of group g of a staged chunk to a word with the NEGATED reference bit of element j
// in bit 7 of byte j.  The reference's bit is int(norm.cdf(z) * 2) == (z >= T), T = quantise_threshold() (tiny, < 0).
//
// The same pass keeps, per thread, a NaN-PROPAGATING running maximum of every element it has seen of the current latent
// (`scan`): extract.py:83 raises for a NaN (int(nan)) and extract.py:86 for any z >= 8.292361075813597 (int(cdf * 2) == 2),
// and the end-of-latent code turns that maximum into the latent's GSWM_FLAG_* -- two FMNMX3.NAN per four fp32 elements
// (two HMNMX2.NAN for 16-bit inputs) instead of two extra passes over the tensor on the host side of the drop-in.
__device__ __forceinline__ float fmax3_nan(float a, float b, float c) {
  float r;
  asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
// smallest value of each type that is >= 8.292361075813597 (where int(norm.cdf(z) * 2) becomes 2)
__device__ __forceinline__ float range_limit_f32() { return __uint_as_float(0x4104AD83u); }   // 8.29236125946045
constexpr float kRangeLimitF16 = 8.296875f;                                                    // 0x4826
constexpr float kRangeLimitBF16 = 8.3125f;                                                     // 0x4105
__device__ __forceinline__ uint32_t range_flags(float running_max, float limit) {
  return running_max != running_max ? (uint32_t)GSWM_FLAG_NAN : (running_max >= limit ? (uint32_t)GSWM_FLAG_RANGE : 0u);
}

template <typename T>
struct NegatedBits;

// fp32: z - T is >= +0 exactly when z >= T (the sum of two floats is never rounded across zero, subnormals are
// kept, an exact zero comes out as +0), so after one FADD the reference's bit is the complement of the sign bit --
// including -0.0 and the [T, 0) sliver, which extract.py:83 maps to 1.  Three byte permutes gather the sign bytes.
template <>
struct NegatedBits<float> {
  struct Scan {
    float m;
    __device__ __forceinline__ void reset() { m = -CUDART_INF_F; }
    __device__ __forceinline__ uint32_t flags() const { return range_flags(m, range_limit_f32()); }
  };
  static __device__ __forceinline__ uint32_t word(const void* stage, uint32_t g, Scan& scan) {
    const float4 z = reinterpret_cast<const float4*>(stage)[g];
    scan.m = fmax3_nan(z.x, z.y, fmax3_nan(z.z, z.w, scan.m));
    const float c = -quantise_threshold();
    const uint32_t s0 = __float_as_uint(z.x + c), s1 = __float_as_uint(z.y + c);
    const uint32_t s2 = __float_as_uint(z.z + c), s3 = __float_as_uint(z.w + c);
    const uint32_t p01 = __byte_perm(s0, s1, 0x0073);    // bytes: s0.b3, s1.b3, -, -
    const uint32_t p23 = __byte_perm(s2, s3, 0x0073);
    return __byte_perm(p01, p23, 0x5410);                // bytes: s0.b3, s1.b3, s2.b3, s3.b3
  }
};

// fp16 / bf16: one packed compare (HSET2.LT) per two elements yields an all-ones halfword where z < T16, and one byte
// permute puts the four halfwords' high bytes in place.  T16 is the smallest value of the type that is >= T:
//   fp16 : nothing lies in [T, 0) except -0.0 (the smallest subnormal is 6e-8), so T16 = +0 -- and -0.0 < 0 is false, as
//          the reference needs (norm.cdf(-0.0) * 2 == 1);
//   bf16 : has fp32's exponent range, so tiny negatives in [T, 0) exist and must decode as 1: T16 = 0xA4A0
//          (-6.94e-17, the fp32 threshold's bit pattern truncated towards zero).  Subnormals are compared, not flushed.
template <typename T2, uint32_t kThresholdBits, bool kHalf>
struct NegatedBits16 {
  struct Scan {
    T2 m;
    __device__ __forceinline__ void reset() {
      const uint32_t ninf = kHalf ? 0xFC00FC00u : 0xFF80FF80u;       // -inf x2 (fp16 | bf16)
      memcpy(&m, &ninf, 4);
    }
    __device__ __forceinline__ uint32_t flags() const {
      float2 f;
      if constexpr (kHalf) f = __half22float2(m); else f = __bfloat1622float2(m);
      const float limit = kHalf ? kRangeLimitF16 : kRangeLimitBF16;
      return range_flags(f.x, limit) | range_flags(f.y, limit);      // NAN | RANGE collapses to NAN in the caller
    }
  };
  static __device__ __forceinline__ uint32_t word(const void* stage, uint32_t g, Scan& scan) {
    const uint2 r = reinterpret_cast<const uint2*>(stage)[g];
    const uint32_t tb = kThresholdBits | (kThresholdBits << 16);
    T2 zx, zy, thr;
    memcpy(&zx, &r.x, 4);
    memcpy(&zy, &r.y, 4);
    memcpy(&thr, &tb, 4);
    scan.m = __hmax2_nan(scan.m, __hmax2_nan(zx, zy));
    const uint32_t nx = __hlt2_mask(zx, thr);                         // 0xFFFF per halfword where z < T16
    const uint32_t ny = __hlt2_mask(zy, thr);
    return __byte_perm(nx, ny, 0x7531);                              // bytes: x.b1, x.b3, y.b1, y.b3
  }
};
// fp64 (what gs_insert.py:75 returns and a float64 caller hands to extract.py:83): the reference's own thresholds, compared
// in double.  Not a throughput path (fp64 compares run at 1/64 rate); it keeps float64 callers on the device.
template <>
struct NegatedBits<double> {
  struct Scan {
    uint32_t f;
    __device__ __forceinline__ void reset() { f = 0; }
    __device__ __forceinline__ uint32_t flags() const { return f; }
  };
  static __device__ __forceinline__ uint32_t word(const void* stage, uint32_t g, Scan& scan) {
    const double2 a = reinterpret_cast<const double2*>(stage)[2 * g], b = reinterpret_cast<const double2*>(stage)[2 * g + 1];
    const double t = -6.957291061679417e-17;                          // int(norm.cdf(z) * 2) == 1  <=>  z >= t
    const double big = 8.292361075813597;                             // int(norm.cdf(z) * 2) == 2  <=>  z >= big
    if (a.x != a.x || a.y != a.y || b.x != b.x || b.y != b.y) scan.f |= GSWM_FLAG_NAN;
    if (a.x >= big || a.y >= big || b.x >= big || b.y >= big) scan.f |= GSWM_FLAG_RANGE;
    return (a.x < t ? 0x00000080u : 0u) | (a.y < t ? 0x00008000u : 0u) | (b.x < t ? 0x00800000u : 0u) | (b.y < t ? 0x80000000u : 0u);
  }
};
template <>
struct NegatedBits<__half> : NegatedBits16<__half2, 0x0000u, true> {};
template <>
struct NegatedBits<__nv_bfloat16> : NegatedBits16<__nv_bfloat162, 0xA4A0u, false> {};

template <typename T, bool kPerLatent, bool kPow2>
__global__ void __launch_bounds__(kThreads, (Ring<T, kPerLatent>::kMinBlocks))
extract_kernel(const ExtractArgs a) {
  constexpr int kStages = Ring<T, kPerLatent>::kStages;
  constexpr int kChunkElems = Chunk<T>::kElems;
  constexpr int kChunksPerTile = Chunk<T>::kPerTile;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* s_stage = smem_raw;                                              // kStages x kStageBytes
  uint32_t* s_ks_all = reinterpret_cast<uint32_t*>(smem_raw + kStages * kStageBytes);  // max(1, ks_cache_tiles) x kTileWords
  uint32_t* s_cnt = s_ks_all + (a.ks_cache_tiles ? a.ks_cache_tiles : 1u) * kTileWords;  // msg_bits counters
  __shared__ __align__(8) uint64_t s_full[kStages];
  __shared__ int s_matched;
  __shared__ uint32_t s_flags;

  const uint32_t cpl = a.chunks_per_latent;
  // byte * magic puts the thread's keystream nibble (MSB = element 0) on bit 7 of bytes 0..3; the other nibble's
  // products land on bits that the 0x80808080 mask drops, and no two partial products share a bit (no carries).
  //   even group (high nibble b7..b4): shifts 0, 9, 18, 27;   odd group (low nibble b3..b0): shifts 4, 13, 22, 31
  const uint32_t spread_magic = (threadIdx.x & 1u) ? 0x80402010u : 0x08040201u;
  // latents owned by this CTA: blockIdx.x + k * gridDim.x
  const int64_t n_mine = a.n_latents > (int64_t)blockIdx.x ? (a.n_latents - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t total_chunks = n_mine * cpl;

  uint64_t policy = 0;
  // producer (thread 0): issue flat chunk q of this CTA's sequence into ring slot q % kStages
  auto issue = [&](int64_t q) {
    if (q >= total_chunks) return;
    const int64_t seq = q / cpl;
    const uint32_t within = (uint32_t)(q - seq * cpl);
    const int64_t latent = blockIdx.x + seq * (int64_t)gridDim.x;
    const int64_t e0 = (int64_t)within * kChunkElems;
    const int64_t rem = a.n_elems - e0;
    const uint32_t bytes = (uint32_t)(rem < kChunkElems ? rem : kChunkElems) * (uint32_t)sizeof(T);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(a.z) + ((size_t)latent * (size_t)a.n_elems + (size_t)e0) * sizeof(T);
    uint64_t* bar = &s_full[q % kStages];
    mbar_expect_tx(bar, bytes);
    tma_load_1d(s_stage + (size_t)(q % kStages) * kStageBytes, src, bytes, bar, policy);
  };
  // Shared key, a latent's whole keystream fits the cache: every CTA computes it once, warp w taking tiles w, w + 8, ...
  auto stage_resident_keystream = [&]() {
    for (uint32_t t = threadIdx.x >> 5; t < a.ks_cache_tiles; t += kThreads / 32) {
      const uint32_t lane = threadIdx.x & 31u;
      if (lane * 16 < tile_words(a.n_elems, t))
        chacha_tile_lane(s_ks_all + t * kTileWords, a.keys, a.nonces, nullptr, 0, t, lane, 0, 0);
    }
  };

  griddep_launch_dependents();
  if (threadIdx.x == 0) {
    for (int sidx = 0; sidx < kStages; ++sidx) mbar_init(&s_full[sidx], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    s_matched = 0;
    s_flags = 0;
  }
  for (uint32_t p = threadIdx.x; p < a.msg_bits; p += kThreads) s_cnt[p] = 0;
  // ... ahead of the grid dependency wait: key material is final before the call is enqueued (gswm.h) unless the job says
  // GSWM_JOB_KEYS_IN_FLIGHT
  if (!kPerLatent && !a.keys_in_flight) stage_resident_keystream();
  __syncthreads();
  griddep_wait();                                                     // nothing above writes global memory or reads a predecessor's output
  if (!kPerLatent && a.keys_in_flight) {
    stage_resident_keystream();
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    for (int64_t q = 0; q < kStages - 1; ++q) issue(q);               // kStages-1 chunks in flight from the start
  }

  unsigned long long acc_matched = 0, acc_exact = 0, acc_msgs = 0, acc_nan = 0, acc_range = 0;   // thread 0 only; flushed once per CTA
  const uint32_t row_bytes = a.msg_stride_bytes;                     // (msg_bits + 7) / 8
  const bool whole_words = (a.msg_bits & 31u) == 0;                  // message rows are whole, aligned 32-bit words

  int64_t q = 0;                                 // flat chunk index
  for (int64_t seq = 0; seq < n_mine; ++seq) {
    const int64_t latent = blockIdx.x + seq * (int64_t)gridDim.x;
    uint32_t packed = 0;                          // kPow2: byte lane k = count of element k
    uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;      // spilled byte lanes (only for very large latents)
    uint32_t adds_since_spill = 0;
    typename NegatedBits<T>::Scan scan;           // NaN-propagating running maximum of this thread's elements of the latent
    scan.reset();

    for (uint32_t within = 0; within < cpl; ++within, ++q) {
      // every thread has finished chunk q-1 (barrier at the end of the previous iteration): refill its slot
      if (threadIdx.x == 0) issue(q + kStages - 1);
      const uint32_t tile = within / kChunksPerTile;
      const uint32_t chunk_in_tile = within % kChunksPerTile;
      const bool ks_resident = !kPerLatent && a.ks_cache_tiles != 0;
      uint32_t* s_ks = ks_resident ? s_ks_all + tile * kTileWords : s_ks_all;
      if (chunk_in_tile == 0 && !ks_resident) {   // new tile: stage its keystream (both paths end with a barrier)
        const uint32_t words = tile_words(a.n_elems, tile);
        compute_private_slice(s_ks, a.keys, a.nonces, nullptr, kPerLatent ? latent : 0, tile, words, 0, 0);
      }
      const uint8_t* s_bytes = reinterpret_cast<const uint8_t*>(s_ks);
      const int64_t e0 = (int64_t)within * kChunkElems;
      const int64_t rem = a.n_elems - e0;
      const uint32_t n_grp = (uint32_t)(rem < kChunkElems ? rem : kChunkElems) >> 2;
      mbar_wait(&s_full[q % kStages], (uint32_t)((q / kStages) & 1));
      const unsigned char* stage = s_stage + (size_t)(q % kStages) * kStageBytes;
      // keystream nibble of group i of the tile lives in byte i>>1 (high nibble for even i); this thread's groups
      // of the chunk are k*256 + tid, so the byte is a compile-time offset from a per-thread base and the nibble
      // half never changes: one multiply spreads it to bit 7 of byte j for element j (see spread magic below).
      const uint8_t* ks_base = s_bytes + chunk_in_tile * (kChunkElems / 8) + (threadIdx.x >> 1);
      auto consume = [&](uint32_t k) {
        const uint32_t g = k * kThreads + threadIdx.x;
        const uint32_t nz = NegatedBits<T>::word(stage, g, scan);
        const uint32_t ks = (uint32_t)ks_base[k * (kThreads / 2)] * spread_magic;
        const uint32_t d = ~(nz ^ ks) & 0x80808080u;                   // decrypted bit of element j in bit 7 of byte j
        if constexpr (kPow2) {
          packed += d >> 7;
        } else {
          // any message length that divides the latent (extract.py:91-98): element e votes for position e % msg_bits
          uint32_t pos = (uint32_t)((e0 + 4 * (int64_t)g) % a.msg_bits);
#pragma unroll
          for (uint32_t j = 0; j < 4; ++j) {
            if (d & (0x80u << (8u * j))) atomicAdd(&s_cnt[pos], 1u);
            if (++pos == a.msg_bits) pos = 0;
          }
        }
      };
      if (n_grp == kChunkElems / 4) {
#pragma unroll
        for (uint32_t k = 0; k < kChunkElems / 4 / kThreads; ++k) consume(k);
      } else {
        for (uint32_t k = 0; k * kThreads + threadIdx.x < n_grp; ++k) consume(k);
      }
      if constexpr (kPow2) {
        constexpr uint32_t kAddsPerChunk = kChunkElems / 4 / kThreads;
        adds_since_spill += kAddsPerChunk;
        if (adds_since_spill + kAddsPerChunk > 255u) {                 // the next chunk could overflow a byte lane
          c0 += packed & 0xFF; c1 += (packed >> 8) & 0xFF; c2 += (packed >> 16) & 0xFF; c3 += packed >> 24;
          packed = 0; adds_since_spill = 0;
        }
      }
      __syncthreads();                            // chunk q fully consumed: its ring slot and s_ks may be reused
    }

    // ---- end of latent: majority vote, pack MSB-first, score against the reference message ----
    {
      // the inputs the reference rejects: any thread's running maximum is NaN or >= 8.2924 (one shared-memory OR per warp that saw one)
      const uint32_t f = scan.flags();
      const uint32_t any = __reduce_or_sync(0xFFFFFFFFu, f);
      if (any && (threadIdx.x & 31u) == 0) atomicOr(&s_flags, any);
    }
    if constexpr (kPow2) {
      c0 += packed & 0xFF; c1 += (packed >> 8) & 0xFF; c2 += (packed >> 16) & 0xFF; c3 += packed >> 24;
      const uint32_t pos = (4u * threadIdx.x) & (a.msg_bits - 1u);
      atomicAdd(&s_cnt[pos + 0], c0);
      atomicAdd(&s_cnt[pos + 1], c1);
      atomicAdd(&s_cnt[pos + 2], c2);
      atomicAdd(&s_cnt[pos + 3], c3);
    }
    __syncthreads();                              // counts and flags of the whole latent are in shared memory
    const uint8_t* ref = a.msgs ? a.msgs + (kPerLatent ? latent * (int64_t)row_bytes : 0) : nullptr;
    uint8_t* out_row = a.msg_out + latent * (int64_t)row_bytes;
    int my_matched = 0;
    for (uint32_t p0 = 0; p0 < a.msg_bits; p0 += kThreads) {          // uniform trip count: whole warps reach the ballot
      const uint32_t p = p0 + threadIdx.x;
      const bool valid = p < a.msg_bits;
      uint32_t cnt = 0;
      if (valid) {
        cnt = s_cnt[p];
        s_cnt[p] = 0;                                                  // ready for the next latent
        if (a.counts) a.counts[latent * a.msg_bits + p] = (uint16_t)cnt;
      }
      const bool bit = valid && 2u * cnt > a.copies;                   // count_1 > len(segments)/2; tie -> 0 (extract.py:99)
      const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, bit);         // lane l = position 32w + l
      // position l of the word -> byte l>>3, bit 7-(l&7): reverse all bits, then swap bytes back
      const uint32_t word = __byte_perm(__brev(ballot), 0u, 0x0123);
      const uint32_t wbase = p & ~31u;                                 // first position of this warp's word
      if ((threadIdx.x & 31) == 0 && wbase < a.msg_bits) {
        if (whole_words) {
          reinterpret_cast<uint32_t*>(out_row)[wbase >> 5] = word;
          if (ref) my_matched += __popc(~(word ^ __ldg(reinterpret_cast<const uint32_t*>(ref) + (wbase >> 5))));
        } else {
          // a row is (msg_bits + 7) / 8 bytes at any alignment: byte stores; bits past msg_bits are zero and not scored
          const uint32_t left = a.msg_bits - wbase;                    // valid positions in this word (>= 1)
          const uint32_t nb = left >= 32u ? 4u : (left + 7u) >> 3;
          uint32_t refw = 0;
          for (uint32_t b = 0; b < nb; ++b) {
            out_row[(wbase >> 3) + b] = (uint8_t)(word >> (8u * b));
            if (ref) refw |= (uint32_t)ref[(wbase >> 3) + b] << (8u * b);
          }
          if (ref) {
            uint32_t mask = 0;                                         // position l valid <=> l < left; byte b holds l = 8b .. 8b+7 MSB-first
            for (uint32_t b = 0; b < nb; ++b) {
              const uint32_t v = left - 8u * b >= 8u ? 8u : left - 8u * b;
              mask |= ((0xFF00u >> v) & 0xFFu) << (8u * b);
            }
            my_matched += __popc(~(word ^ refw) & mask);
          }
        }
      }
    }
    if (ref) {
      if ((threadIdx.x & 31) == 0 && my_matched) atomicAdd(&s_matched, my_matched);
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      uint32_t f = s_flags;
      s_flags = 0;
      if (f & GSWM_FLAG_NAN) f = GSWM_FLAG_NAN;                        // int(nan) raises before the digit string is parsed
      if (a.flags) a.flags[latent] = (uint8_t)f;
      acc_nan += (unsigned long long)(f == GSWM_FLAG_NAN);
      acc_range += (unsigned long long)(f == GSWM_FLAG_RANGE);
      if (ref) {
        const int m = s_matched;
        s_matched = 0;
        if (a.matched) a.matched[latent] = m;
        acc_matched += (unsigned long long)m;
        acc_exact += (unsigned long long)(m == (int)a.msg_bits);
      }
      ++acc_msgs;
    }
    __syncthreads();                              // s_cnt / s_matched / s_flags reset visible before the next latent
  }
  if (threadIdx.x == 0 && a.counters && acc_msgs) {
    if (a.msgs) {
      atomicAdd(&a.counters[GSWM_CTR_MATCHED_BITS], acc_matched);
      atomicAdd(&a.counters[GSWM_CTR_EXACT_MSGS], acc_exact);
    }
    atomicAdd(&a.counters[GSWM_CTR_TOTAL_BITS], acc_msgs * a.msg_bits);
    atomicAdd(&a.counters[GSWM_CTR_TOTAL_MSGS], acc_msgs);
    if (acc_nan) atomicAdd(&a.counters[GSWM_CTR_NAN_LATENTS], acc_nan);
    if (acc_range) atomicAdd(&a.counters[GSWM_CTR_RANGE_LATENTS], acc_range);
  }

  // ---- gswm_extract_allreduce: the collective that follows K3, inside K3 ------------------------------------------------
  // The last CTA to retire holds this rank's final counters; its first warp publishes them to every peer's mailbox over
  // NVLink and collects the peers' (gswm_comm.cuh).  The other CTAs are gone by then: the wait costs one warp of one SM.
  if (a.reduced != nullptr && threadIdx.x < 32) {
    unsigned ticket = 0;
    if (threadIdx.x == 0) {
      __threadfence();                                                 // this CTA's counter atomics before its ticket
      ticket = atomicAdd(a.comm.ticket, 1u);
    }
    ticket = __shfl_sync(0xFFFFFFFFu, ticket, 0);
    if (ticket == gridDim.x - 1) {
      __threadfence();
      long long mine[GSWM_COMM_MAX_VALUES], sum[GSWM_COMM_MAX_VALUES];
      const long long own = threadIdx.x < GSWM_N_COUNTERS ? (long long)atomicAdd(&a.counters[threadIdx.x], 0ull) : 0ll;
#pragma unroll
      for (int i = 0; i < GSWM_COMM_MAX_VALUES; ++i) mine[i] = __shfl_sync(0xFFFFFFFFu, own, i);
      comm_allreduce_warp(a.comm, mine, GSWM_N_COUNTERS, sum);
#pragma unroll
      for (int i = 0; i < GSWM_N_COUNTERS; ++i)
        if (threadIdx.x == i) a.reduced[i] = sum[i];
      if (threadIdx.x == 0) *a.comm.ticket = 0u;                       // ready for the next launch (stream-ordered)
    }
  }
}

// Stand-alone form of the same exchange (gswm_comm_allreduce_counters): one warp, in place.
__global__ void __launch_bounds__(32)
comm_allreduce_kernel(const CommDev c, long long* __restrict__ values, int n) {
  long long mine[GSWM_COMM_MAX_VALUES], sum[GSWM_COMM_MAX_VALUES];
#pragma unroll
  for (int i = 0; i < GSWM_COMM_MAX_VALUES; ++i) mine[i] = i < n ? values[i] : 0ll;
  comm_allreduce_warp(c, mine, n, sum);
#pragma unroll
  for (int i = 0; i < GSWM_COMM_MAX_VALUES; ++i)
    if ((int)threadIdx.x == i && i < n) values[i] = sum[i];
}

// Test hook: evaluate the fp32 bucket quantile on caller-supplied raw words (exhaustive accuracy test).
// m = w >> 9; bucket bit 1: z = +g((m + 1/2) 2^-23); bucket bit 0: z = -g(.) (so the reference's u is 1 - v).
__global__ void __launch_bounds__(kThreads)
debug_quantile_kernel(const uint32_t* __restrict__ w, int64_t n, uint32_t bucket_bit, int use_vec4, float* __restrict__ out) {
  const int64_t i = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 4;
  if (i >= n) return;
  const float s = bucket_bit ? 1.f : -1.f;
  const uint4 r = *reinterpret_cast<const uint4*>(w + i);
  float4 z;
  if (use_vec4) {
    z = bucket_quantile4_f32(fbits_top23(r.x), fbits_top23(r.y), fbits_top23(r.z), fbits_top23(r.w), make_float4(s, s, s, s));
  } else {
    z = make_float4(s * halfnormal_quantile(fbits_top23(r.x)), s * halfnormal_quantile(fbits_top23(r.y)),
                    s * halfnormal_quantile(fbits_top23(r.z)), s * halfnormal_quantile(fbits_top23(r.w)));
  }
  *reinterpret_cast<float4*>(out + i) = z;
}

// Test hook: |z| of the refined outermost cell for caller-supplied 32-bit refinement words.
__global__ void __launch_bounds__(kThreads)
debug_top_cell_kernel(const uint32_t* __restrict__ w, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i < n) out[i] = top_cell_quantile(w[i]);
}

// Test hook: the Philox4x32 core on caller-supplied counters / keys, 7 rounds (the product) or 10 (the variant with
// published known-answer vectors).
template <int kRounds>
__global__ void __launch_bounds__(kThreads)
debug_philox_kernel(const uint32_t* __restrict__ in, int64_t n, uint32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const uint32_t* p = in + 6 * i;
  const uint4 r = philox4x32<kRounds>(make_uint4(p[0], p[1], p[2], p[3]), p[4], p[5]);
  out[4 * i + 0] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
}

__global__ void __launch_bounds__(kThreads)
debug_ppf64_kernel(const double* __restrict__ p, int64_t n, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i < n) out[i] = norm_ppf_f64(p[i]);
}

// ------------------------------------------------------------------------------------------------
// host side of the device entry points
// ------------------------------------------------------------------------------------------------
static inline bool per_latent_of(const gswm_job* job) { return (job->flags & GSWM_JOB_PER_LATENT) != 0; }

int check_job(const gswm_job* job, bool for_extract) {
  if (!job || !job->d_keys || !job->d_nonces) return GSWM_E_NULL;
  if (!for_extract && !job->d_msgs) return GSWM_E_NULL;
  if ((reinterpret_cast<uintptr_t>(job->d_keys) | reinterpret_cast<uintptr_t>(job->d_nonces)) & 3u) return GSWM_E_ALIGN;   // read as 32-bit words
  if (job->n_latents < 0 || job->n_elems <= 0 || (job->n_elems % 4) != 0) return GSWM_E_SHAPE;
  if (job->msg_bits <= 0 || job->msg_bits > job->n_elems) return GSWM_E_MSGLEN;
  if (for_extract) {
    // extract.py:91-98 splits the bit string into message_length-sized segments: any divisor of the latent size works
    if ((job->n_elems % job->msg_bits) != 0) return GSWM_E_MSGLEN;
  } else if ((job->msg_bits % 32) != 0) {
    return GSWM_E_MSGLEN;
  }
  // message rows are read as 32-bit words whenever they are whole words (always for embed)
  if ((job->msg_bits % 32) == 0 && (reinterpret_cast<uintptr_t>(job->d_msgs) & 3u)) return GSWM_E_ALIGN;
  if (job->n_elems > ((int64_t)1 << 31)) return GSWM_E_RANGE;
  const int64_t tiles = (job->n_elems + kTileElems - 1) / kTileElems;
  if (job->n_latents > 0x7FFFFFFFll || tiles > 32767) return GSWM_E_RANGE;    // grid.y carries tiles or half tiles
  return GSWM_OK;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static uint32_t tiles_of(int64_t n_elems) { return (uint32_t)((n_elems + kTileElems - 1) / kTileElems); }

static unsigned halves_of(int64_t n_elems) { return (unsigned)((n_elems / 4 + 8 * kThreads - 1) / (8 * kThreads)); }

static EmbedArgs make_embed_args(const gswm_job* job) {
  EmbedArgs a{};
  a.keys = job->d_keys;
  a.nonces = job->d_nonces;
  a.msgs = job->d_msgs;
  a.n_elems = job->n_elems;
  a.n_latents = job->n_latents;
  a.tiles_per_latent = tiles_of(job->n_elems);
  a.msg_words = (uint32_t)job->msg_bits / 32;
  a.tiled_words = (uint32_t)(job->n_elems / job->msg_bits) * a.msg_words;
  a.msg_stride_bytes = (uint32_t)job->msg_bits / 8;
  a.keys_in_flight = (job->flags & GSWM_JOB_KEYS_IN_FLIGHT) ? 1u : 0u;
  return a;
}

// Launch with programmatic stream serialisation allowed (see griddep_wait above); GSWM_PDL=0 in the environment
// falls back to ordinary stream order (the device-side griddepcontrol instructions are then no-ops).
static bool pdl_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("GSWM_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

template <typename Args>
static int launch_pdl(void (*kernel)(const Args), dim3 grid, size_t smem, cudaStream_t st, const Args& a) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return (int)cudaLaunchKernelEx(&cfg, kernel, a);
}

// Tuning knob for co-scheduling experiments (tools/cobench.py): an environment variable may LOWER the number of
// CTAs per SM a persistent grid is sized for, leaving room for another kernel on the same SMs.
static int cap_ctas(int per_sm, const char* env_name) {
  const char* e = std::getenv(env_name);
  const int c = e ? std::atoi(e) : 0;
  return (c >= 1 && c < per_sm) ? c : per_sm;
}

// Persistent launch: grid = min(work items, SMs x resident CTAs of this kernel at this shared-memory size).
template <typename K, typename Args>
static int launch_persistent(K kernel, const Args& a, int64_t items, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int dev = 0, sms = 0, per_sm = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return (int)e;
  if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem)) != cudaSuccess) return (int)e;
  if (per_sm < 1) return (int)cudaErrorLaunchOutOfResources;
  per_sm = cap_ctas(per_sm, "GSWM_EXTRACT_CTAS_PER_SM");
  const int64_t resident = (int64_t)sms * per_sm;
  const unsigned grid = (unsigned)(items < resident ? items : resident);
  return launch_pdl(kernel, dim3(grid), smem, st, a);
}

// Embed grid (X, Y): Y fixed pieces of a latent (half tiles / tiles), X latent lanes.  Persistent: X * Y CTAs fill the GPU once.
template <typename K>
static int launch_embed(K kernel, const EmbedArgs& a, unsigned y, bool persistent, cudaStream_t st) {
  int dev = 0, sms = 0, per_sm = 0;
  cudaError_t e;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return (int)e;
  if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0)) != cudaSuccess) return (int)e;
  if (per_sm < 1) return (int)cudaErrorLaunchOutOfResources;
  per_sm = cap_ctas(per_sm, "GSWM_EMBED_CTAS_PER_SM");
  int64_t x = persistent ? ((int64_t)sms * per_sm) / y : a.n_latents;
  if (x < 1) x = 1;
  if (x > a.n_latents) x = a.n_latents;
  return launch_pdl(kernel, dim3((unsigned)x, y), 0, st, a);
}

template <typename K>
static int launch_extract_kernel(K kernel, const ExtractArgs& a, size_t smem, cudaStream_t st) {
  return launch_persistent(kernel, a, a.n_latents, smem, st);
}

template <typename T>
static int launch_extract(const ExtractArgs& a, bool per_latent, bool pow2, size_t smem, cudaStream_t st) {
  if (per_latent) {
    if (pow2) return launch_extract_kernel(extract_kernel<T, true, true>, a, smem, st);
    return launch_extract_kernel(extract_kernel<T, true, false>, a, smem, st);
  }
  if (pow2) return launch_extract_kernel(extract_kernel<T, false, true>, a, smem, st);
  return launch_extract_kernel(extract_kernel<T, false, false>, a, smem, st);
}

}  // namespace gswm

using namespace gswm;

extern "C" {

int gswm_abi_version(void) { return GSWM_ABI_VERSION; }

const char* gswm_strerror(int code) {
  switch (code) {
    case GSWM_OK: return "success";
    case GSWM_E_NULL: return "gswm: a required pointer is NULL";
    case GSWM_E_SHAPE: return "gswm: n_elems must be a positive multiple of 4 and n_latents >= 0";
    case GSWM_E_MSGLEN: return "gswm: msg_bits must be positive and <= n_elems; a multiple of 32 for embed, a divisor of n_elems for extract";
    case GSWM_E_DTYPE: return "gswm: unknown element type";
    case GSWM_E_RANGE: return "gswm: size out of range";
    case GSWM_E_COMM: return "gswm: communicator error (bad rank / size, peer memory not mappable, NCCL not loadable, or a peer timed out)";
    case GSWM_E_ALIGN: return "gswm: latent pointer or row not 16-byte aligned, or key / nonce / message pointer not 4-byte aligned";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "gswm: unknown error";
  }
}

int gswm_debug_bucket_quantile(const uint32_t* d_words, int64_t n, int32_t bucket_bit, int32_t use_vec4,
                               float* d_out, void* stream) {
  if (!d_words || !d_out) return GSWM_E_NULL;
  if (n < 0 || (n % 4) != 0) return GSWM_E_SHAPE;
  if (!aligned16(d_words) || !aligned16(d_out)) return GSWM_E_ALIGN;
  if (n == 0) return GSWM_OK;
  const int64_t grid = (n / 4 + kThreads - 1) / kThreads;
  debug_quantile_kernel<<<(unsigned)grid, kThreads, 0, (cudaStream_t)stream>>>(d_words, n, bucket_bit ? 1u : 0u,
                                                                               use_vec4, d_out);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

int gswm_debug_norm_ppf(const double* d_p, int64_t n, double* d_out, void* stream) {
  if (!d_p || !d_out) return GSWM_E_NULL;
  if (n <= 0) return n == 0 ? GSWM_OK : GSWM_E_SHAPE;
  debug_ppf64_kernel<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(d_p, n, d_out);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

int gswm_debug_top_cell(const uint32_t* d_words, int64_t n, float* d_out, void* stream) {
  if (!d_words || !d_out) return GSWM_E_NULL;
  if (n < 0) return GSWM_E_SHAPE;
  if (n == 0) return GSWM_OK;
  debug_top_cell_kernel<<<(unsigned)((n + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(d_words, n, d_out);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

int gswm_debug_philox4x32(const uint32_t* d_in, int64_t n, int32_t rounds, uint32_t* d_out, void* stream) {
  if (!d_in || !d_out) return GSWM_E_NULL;
  if (n < 0) return GSWM_E_SHAPE;
  if (rounds != 7 && rounds != 10) return GSWM_E_RANGE;
  if (n == 0) return GSWM_OK;
  const unsigned grid = (unsigned)((n + kThreads - 1) / kThreads);
  if (rounds == 7) debug_philox_kernel<7><<<grid, kThreads, 0, (cudaStream_t)stream>>>(d_in, n, d_out);
  else debug_philox_kernel<10><<<grid, kThreads, 0, (cudaStream_t)stream>>>(d_in, n, d_out);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

int64_t gswm_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int gswm_philox_rounds(void) { return GSWM_PHILOX_ROUNDS; }

#ifdef GSWM_TRACE
int gswm_debug_trace_read(unsigned long long* h_out, int buf) {      // h_out[4 * 8192]
  return (int)cudaMemcpyFromSymbol(h_out, g_trace, sizeof(unsigned long long) * 4 * 8192, sizeof(unsigned long long) * 4 * 8192 * (size_t)(buf & 1));
}
int gswm_debug_trace_select(int buf, void* stream) {                 // stream-ordered: later launches write buffer `buf`
  const unsigned b = (unsigned)buf;
  return (int)cudaMemcpyToSymbolAsync(g_trace_buf, &b, sizeof(b), 0, cudaMemcpyHostToDevice, (cudaStream_t)stream);
}
#endif

int gswm_chacha20_keystream(const uint8_t* d_keys, const uint8_t* d_nonces, int64_t n_streams,
                            int64_t n_bytes_each, uint8_t* d_out, void* stream) {
  if (!d_keys || !d_nonces || !d_out) return GSWM_E_NULL;
  if (n_streams < 0 || n_bytes_each <= 0 || (n_bytes_each % 64) != 0) return GSWM_E_SHAPE;
  if (!aligned16(d_out)) return GSWM_E_ALIGN;
  const int64_t blocks_each = n_bytes_each / 64;
  if (blocks_each > 0xFFFFFFFFll) return GSWM_E_RANGE;
  const int64_t total = n_streams * blocks_each;
  if (total == 0) return GSWM_OK;
  const int64_t grid = (total + kThreads - 1) / kThreads;
  if (grid > 0x7FFFFFFFll) return GSWM_E_RANGE;
  chacha20_keystream_kernel<<<(unsigned)grid, kThreads, 0, (cudaStream_t)stream>>>(
      d_keys, d_nonces, nullptr, n_streams, (uint32_t)blocks_each, 0, 0, 0, reinterpret_cast<uint32_t*>(d_out));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

int gswm_embed(const gswm_job* job, uint64_t seed, uint64_t offset, int64_t first_latent, float* d_out, void* stream) {
  int rc = check_job(job, false);
  if (rc) return rc;
  if (!d_out) return GSWM_E_NULL;
  if (!aligned16(d_out)) return GSWM_E_ALIGN;
  if ((offset >> 62) || first_latent < 0) return GSWM_E_RANGE;        // a negative index would wrap into another latent's counters
  if (job->n_latents == 0) return GSWM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  EmbedArgs a = make_embed_args(job);
  a.out = d_out;
  a.first_latent = first_latent;
  a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32);
#if GSWM_CTR_V4
  a.off_lo = (uint32_t)offset; a.off_hi = (uint32_t)(offset >> 32);
  {
    const unsigned __int128 last = ((unsigned __int128)first_latent + (unsigned __int128)job->n_latents) * a.tiles_per_latent * 4u;
    if (last > ((unsigned __int128)1 << 54)) return GSWM_E_RANGE;     // `last` is one past the largest T: the counter holds 54 bits of it
    const uint64_t pa = 0xD2511F53ull * a.off_lo, pb = 0xCD9E8D57ull * a.off_hi;
    a.pl.a_hi = (uint32_t)(pa >> 32); a.pl.b_lo = (uint32_t)pb;
    a.pl.x0 = (uint32_t)(pb >> 32) ^ a.seed_lo;                       // round keys 0: (seed_lo, seed_hi)
    a.pl.x3 = (uint32_t)pa ^ (a.seed_hi + 0xBB67AE85u);               // round keys 1, word 1
  }
#else
  a.off_lo = (uint32_t)offset; a.off_hi = (uint32_t)(offset >> 32) << 2;
#endif
  for (int r = 0; r < GSWM_PHILOX_ROUNDS; ++r) {
    a.rk.k[2 * r] = a.seed_lo + (uint32_t)r * 0x9E3779B9u;
    a.rk.k[2 * r + 1] = a.seed_hi + (uint32_t)r * 0xBB67AE85u;
  }
  rc = per_latent_of(job) ? launch_embed(embed_kernel<true>, a, a.tiles_per_latent, false, st)
                          : launch_embed(embed_kernel<false>, a, halves_of(job->n_elems), true, st);   // Y = non-empty half tiles
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return rc;
}

int gswm_embed_injected(const gswm_job* job, const double* d_u, int32_t u_per_latent, void* d_out,
                        int32_t out_dtype, void* stream) {
  int rc = check_job(job, false);
  if (rc) return rc;
  if (!d_out || !d_u) return GSWM_E_NULL;
  if (out_dtype != GSWM_F32 && out_dtype != GSWM_F64) return GSWM_E_DTYPE;
  if (job->n_latents == 0) return GSWM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  EmbedArgs a = make_embed_args(job);
  // small jobs: several CTAs per tile, so that one latent is not one CTA (see the kernel)
  const int64_t ctas = job->n_latents * (int64_t)tiles_of(job->n_elems);
  unsigned split = 1;
  while (split < 16 && ctas * split * 2 <= 296) split *= 2;
  const dim3 grid((unsigned)job->n_latents, tiles_of(job->n_elems), split);
  const int upl = u_per_latent ? 1 : 0;
  if (out_dtype == GSWM_F32) {
    if (per_latent_of(job)) embed_injected_kernel<true, float><<<grid, kThreads, 0, st>>>(a, d_u, upl, (float*)d_out);
    else embed_injected_kernel<false, float><<<grid, kThreads, 0, st>>>(a, d_u, upl, (float*)d_out);
  } else {
    if (per_latent_of(job)) embed_injected_kernel<true, double><<<grid, kThreads, 0, st>>>(a, d_u, upl, (double*)d_out);
    else embed_injected_kernel<false, double><<<grid, kThreads, 0, st>>>(a, d_u, upl, (double*)d_out);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

}  // extern "C"

namespace gswm {
// gswm_extract and gswm_extract_allreduce (the latter from gswm_comm.cu, which owns the communicator)
int extract_impl(const gswm_job* job, const void* d_z, int32_t z_dtype, uint8_t* d_msg_out, uint16_t* d_counts,
                 int32_t* d_matched, uint8_t* d_flags, int64_t* d_counters, const CommDev* comm, int64_t* d_reduced,
                 void* stream) {
  int rc = check_job(job, true);
  if (rc) return rc;
  if (!d_z || !d_msg_out) return GSWM_E_NULL;
  if (z_dtype != GSWM_F32 && z_dtype != GSWM_F16 && z_dtype != GSWM_BF16 && z_dtype != GSWM_F64) return GSWM_E_DTYPE;
  const bool whole_words = (job->msg_bits % 32) == 0;
  if (!aligned16(d_z) || (whole_words && (reinterpret_cast<uintptr_t>(d_msg_out) & 3u))) return GSWM_E_ALIGN;
  const size_t esz = z_dtype == GSWM_F32 ? 4 : z_dtype == GSWM_F64 ? 8 : 2;
  if (((size_t)job->n_elems * esz) % 16 != 0) return GSWM_E_ALIGN;     // every latent row is fetched with 16-byte bulk copies
  const int64_t copies = job->n_elems / job->msg_bits;
  if (d_counts && copies > 65535) return GSWM_E_RANGE;
  if (job->msg_bits > 8192) return GSWM_E_RANGE;
  if (comm && (!d_reduced || !d_counters)) return GSWM_E_NULL;
  if (job->n_latents == 0 && !comm) return GSWM_OK;
  if (job->n_latents == 0) return GSWM_E_SHAPE;                        // the fused exchange lives in the extract kernel: it needs one launch
  cudaStream_t st = (cudaStream_t)stream;
  ExtractArgs a{};
  const bool per_latent = per_latent_of(job);
  a.keys = job->d_keys; a.nonces = job->d_nonces; a.msgs = job->d_msgs;
  a.z = d_z; a.msg_out = d_msg_out; a.counts = d_counts; a.matched = d_matched; a.flags = d_flags;
  a.counters = reinterpret_cast<unsigned long long*>(d_counters);
  a.n_elems = job->n_elems;
  a.tiles_per_latent = tiles_of(job->n_elems);
  a.msg_bits = (uint32_t)job->msg_bits;
  a.msg_stride_bytes = ((uint32_t)job->msg_bits + 7u) / 8u;
  a.copies = (uint32_t)copies;
  a.n_latents = job->n_latents;
  a.keys_in_flight = (job->flags & GSWM_JOB_KEYS_IN_FLIGHT) ? 1u : 0u;
  if (comm) {
    a.comm = *comm;
    a.reduced = reinterpret_cast<long long*>(d_reduced);
  }
  const int64_t chunk_elems = z_dtype == GSWM_F32 ? Chunk<float>::kElems : z_dtype == GSWM_F64 ? Chunk<double>::kElems : Chunk<__half>::kElems;
  a.chunks_per_latent = (uint32_t)((job->n_elems + chunk_elems - 1) / chunk_elems);
  // fast path: a thread's four vote positions never change (msg_bits is a power of two in [4, 1024]); otherwise shared-memory atomics
  const bool pow2 = job->msg_bits >= 4 && (1024 % job->msg_bits) == 0;
  a.ks_cache_tiles = (!per_latent && a.tiles_per_latent <= 8) ? a.tiles_per_latent : 0;
  const int stages = z_dtype != GSWM_F32 ? Ring<__half, false>::kStages
                                         : (per_latent ? Ring<float, true>::kStages : Ring<float, false>::kStages);
  const size_t smem = (size_t)stages * kStageBytes +
                      (size_t)((a.ks_cache_tiles ? a.ks_cache_tiles : 1u) * kTileWords + job->msg_bits) * sizeof(uint32_t);
  if (z_dtype == GSWM_F32) rc = launch_extract<float>(a, per_latent, pow2, smem, st);
  else if (z_dtype == GSWM_F64) rc = launch_extract<double>(a, per_latent, pow2, smem, st);
  else if (z_dtype == GSWM_F16) rc = launch_extract<__half>(a, per_latent, pow2, smem, st);
  else rc = launch_extract<__nv_bfloat16>(a, per_latent, pow2, smem, st);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return rc;
}

int comm_allreduce_launch(const CommDev& c, int64_t* d_values, int n, void* stream) {
  comm_allreduce_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(c, reinterpret_cast<long long*>(d_values), n);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}
}  // namespace gswm

extern "C" {

int gswm_extract(const gswm_job* job, const void* d_z, int32_t z_dtype, uint8_t* d_msg_out, uint16_t* d_counts,
                 int32_t* d_matched, uint8_t* d_flags, int64_t* d_counters, void* stream) {
  return gswm::extract_impl(job, d_z, z_dtype, d_msg_out, d_counts, d_matched, d_flags, d_counters, nullptr, nullptr, stream);
}

}  // extern "C"
