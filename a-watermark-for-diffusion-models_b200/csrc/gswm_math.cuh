// Device-side building blocks of the Gaussian-Shading codec for sm_100a:
//   * ChaCha20 block function, one block per thread (integer add / xor / funnel-shift rotate);
//   * Philox4x32 counter-based generator (the product's uniform source);
//   * half-normal quantile g(v) = sqrt(2) erfinv(v) in fp32 (registers only: 1 MUFU.LG2 + FFMAs) and
//     fp64 (injected-uniform mode), coefficients from tools/fit_halfnormal_quantile.py.
// Nothing here touches memory.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "gswm_coeffs.inc"

namespace gswm {

// ------------------------------------------------------------------------------------------------
// ChaCha20 (DJB layout as `cryptography` exposes it: 64-bit block counter in words 12..13)
// Replaces Cipher(algorithms.ChaCha20(key, nonce)) at gs_insert.py:45-47 / extract.py:77-78.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int n) { return __funnelshift_l(x, x, n); }

#ifndef GSWM_CHACHA_UNROLL
#define GSWM_CHACHA_UNROLL 2     // double rounds unrolled together: code size counts (per-latent keys: 64.4 us with 2 against 69.0 us fully unrolled -- the CTAs of an SM sit in different parts of the kernel and share its instruction cache)
#endif
constexpr int kChachaUnroll = GSWM_CHACHA_UNROLL;
#define GSWM_QR(a, b, c, d)            \
  a += b; d ^= a; d = rotl32(d, 16);   \
  c += d; b ^= c; b = rotl32(b, 12);   \
  a += b; d ^= a; d = rotl32(d, 8);    \
  c += d; b ^= c; b = rotl32(b, 7);

// key: 8 LE words; nonce: 4 LE words (words 0..1 = initial 64-bit counter, 2..3 = nonce proper);
// block: block index added to the counter with carry.  out: 16 keystream words (LE byte order).
__device__ __forceinline__ void chacha20_block(const uint32_t (&key)[8], const uint32_t (&nonce)[4],
                                               uint32_t block, uint32_t (&out)[16]) {
  const uint32_t c0 = 0x61707865u, c1 = 0x3320646eu, c2 = 0x79622d32u, c3 = 0x6b206574u;
  const uint32_t ctr_lo = nonce[0] + block;
  const uint32_t ctr_hi = nonce[1] + (ctr_lo < block ? 1u : 0u);
  uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3;
  uint32_t x4 = key[0], x5 = key[1], x6 = key[2], x7 = key[3];
  uint32_t x8 = key[4], x9 = key[5], x10 = key[6], x11 = key[7];
  uint32_t x12 = ctr_lo, x13 = ctr_hi, x14 = nonce[2], x15 = nonce[3];
#pragma unroll kChachaUnroll
  for (int r = 0; r < 10; ++r) {
    GSWM_QR(x0, x4, x8, x12) GSWM_QR(x1, x5, x9, x13) GSWM_QR(x2, x6, x10, x14) GSWM_QR(x3, x7, x11, x15)
    GSWM_QR(x0, x5, x10, x15) GSWM_QR(x1, x6, x11, x12) GSWM_QR(x2, x7, x8, x13) GSWM_QR(x3, x4, x9, x14)
  }
  out[0] = x0 + c0;  out[1] = x1 + c1;  out[2] = x2 + c2;  out[3] = x3 + c3;
  out[4] = x4 + key[0];  out[5] = x5 + key[1];  out[6] = x6 + key[2];  out[7] = x7 + key[3];
  out[8] = x8 + key[4];  out[9] = x9 + key[5];  out[10] = x10 + key[6];  out[11] = x11 + key[7];
  out[12] = x12 + ctr_lo;  out[13] = x13 + ctr_hi;  out[14] = x14 + nonce[2];  out[15] = x15 + nonce[3];
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-R (Salmon, Moraes, Dror, Shaw, SC'11).  Takes the place of np.random.uniform at
// gs_insert.py:62; restated for the oracle in oracle/gs_oracle.py:philox4x32.
//
// R = 7 by default: Philox4x32-7 is the fewest rounds the authors report as Crush-resistant (passes TestU01
// SmallCrush, Crush and BigCrush; Table 2 of the paper -- 10 rounds is their "with safety margin" variant).  The
// generator it stands in for, numpy's MT19937, does not pass BigCrush (linear-complexity tests), so 7 rounds is not
// a step down from the reference; and the secrecy of the watermark rests on the ChaCha20 keystream that picks the
// bucket, not on the within-bucket uniform.  The embed kernel is bound by the FMA-heavy pipe that executes
// Philox's IMAD.WIDE, so rounds are time: 58.9 us (R = 7) against 66.9 us (R = 10) per 4096 SD-2.1 latents.
// -DGSWM_PHILOX_ROUNDS=10 builds the curand-compatible round count (oracle: GSWM_PHILOX_ROUNDS in gs_oracle.py;
// gswm_philox_rounds() reports what a library was built with).
// ------------------------------------------------------------------------------------------------
#ifndef GSWM_PHILOX_ROUNDS
#define GSWM_PHILOX_ROUNDS 7
#endif
template <int kRounds = GSWM_PHILOX_ROUNDS>
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < kRounds; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c.x;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c.z;
    uint4 n;
    n.x = (uint32_t)(p1 >> 32) ^ c.y ^ k0;
    n.y = (uint32_t)p1;
    n.z = (uint32_t)(p0 >> 32) ^ c.w ^ k1;
    n.w = (uint32_t)p0;
    c = n;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

// Same generator with the round keys (k0 + r W0, k1 + r W1) precomputed by the host into kernel-parameter space:
// the xor then takes its key straight from the constant bank instead of from a uniform register that has to be
// rematerialised with a UIADD3 per key per loop iteration.
struct PhiloxKeys {
  uint32_t k[2 * GSWM_PHILOX_ROUNDS];
};
__device__ __forceinline__ uint4 philox4x32_keys(uint4 c, const PhiloxKeys& rk) {
#pragma unroll
  for (int r = 0; r < GSWM_PHILOX_ROUNDS; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c.x;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c.z;
    uint4 n;
    n.x = (uint32_t)(p1 >> 32) ^ c.y ^ rk.k[2 * r];
    n.y = (uint32_t)p1;
    n.z = (uint32_t)(p0 >> 32) ^ c.w ^ rk.k[2 * r + 1];
    n.w = (uint32_t)p0;
    c = n;
  }
  return c;
}

// Three Philox4x32 calls for one super-iteration under the v4 counter layout (see "The product's uniform source" below):
//   ctr = (offset_lo, c1, offset_hi, c3),  c1 = tid | (T & 0xFFFFFF) << 8,  c3 = (T >> 24) << 2 | call,  call = 0..2
// with T = (global latent * tiles + tile) * 4 + super-iteration: the only word that differs from lane to lane is c1, and
// c1 is not multiplied in round 1.  Evaluated as written this is plain Philox4x32-R on those counters; evaluated with
// the data flow in mind, what does not depend on the lane is computed ONCE PER CTA (philox_v4_cta_part: a table in shared
// memory filled ahead of the latent loop, read back with one broadcast LDS.128 per call), and what depends on the lane
// but not on the call is shared by the three calls:
//   round 1: both products are constants of the launch (M0 * offset_lo, M1 * offset_hi, host-computed: PhiloxLaunch);
//   round 2: M0 * x is one multiply for all three calls; M1 * z does not depend on the lane (table);
//   round 3: M0 * x does not depend on the lane (table); M1 * z is one multiply for all three calls;
//   rounds 4..R: two multiplies per call.
// 2 + 6 (R - 3) = 26 IMAD.WIDE per lane for R = 7 instead of 35 with the lane index in word 0 (v3: 1 + 4 + 6 (R - 2)), and
// 9 fewer LOP3 -- about a tenth of the embed kernel's FMA-heavy-pipe time, the pipe that bounds it.  (Left to the
// compiler, the lane-independent products stay vector IMAD.WIDE -- it does not move them to the uniform datapath -- and
// nothing is gained: 68 IMAD.WIDE per 8 float4 either way, tools/sass_loop_hist.py.)
struct PhiloxLaunch {
  uint32_t a_hi;            // hi(M0 * offset_lo)
  uint32_t b_lo;            // lo(M1 * offset_hi)
  uint32_t x0;              // hi(M1 * offset_hi) ^ round key 0 (word 0)
  uint32_t x3;              // lo(M0 * offset_lo) ^ round key 1 (word 1)     [rk.k[3]]
};
// Lane-independent part of rounds 1..3 for counter word 3 = c3: {m1 ^ k4, hi(q) ^ k5, lo(q), -}
__device__ __forceinline__ uint4 philox_v4_cta_part(uint32_t c3, const PhiloxLaunch& L, const PhiloxKeys& rk) {
  const uint32_t n2 = L.a_hi ^ c3 ^ rk.k[1];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * n2;
  const uint32_t m0 = (uint32_t)(p1 >> 32) ^ L.b_lo ^ rk.k[2];
  const uint64_t q = (uint64_t)0xD2511F53u * m0;
  return make_uint4((uint32_t)p1 ^ rk.k[4], (uint32_t)(q >> 32) ^ rk.k[5], (uint32_t)q, 0u);
}
// uni[call]: the table entries of this super-iteration (shared memory, same address for every lane)
__device__ __forceinline__ void philox4x32_v4_calls3(uint4 (&out)[3], uint32_t c1, const uint4* __restrict__ uni,
                                                     const PhiloxLaunch& L, const PhiloxKeys& rk) {
  static_assert(GSWM_PHILOX_ROUNDS >= 3, "the shared evaluation covers rounds 1..3");
  const uint32_t n0 = L.x0 ^ c1;                                      // round 1, word 0
  const uint64_t p0 = (uint64_t)0xD2511F53u * n0;                     // round 2
  const uint32_t m2 = (uint32_t)(p0 >> 32) ^ L.x3;
  const uint32_t m3 = (uint32_t)p0;
  const uint64_t p1s = (uint64_t)0xCD9E8D57u * m2;                    // round 3
#pragma unroll
  for (uint32_t call = 0; call < 3; ++call) {
    const uint4 u = uni[call];
    uint4 c = make_uint4((uint32_t)(p1s >> 32) ^ u.x, (uint32_t)p1s, u.y ^ m3, u.z);
#pragma unroll
    for (int r = 3; r < GSWM_PHILOX_ROUNDS; ++r) {
      const uint64_t x0 = (uint64_t)0xD2511F53u * c.x;
      const uint64_t x1 = (uint64_t)0xCD9E8D57u * c.z;
      c = make_uint4((uint32_t)(x1 >> 32) ^ c.y ^ rk.k[2 * r], (uint32_t)x1, (uint32_t)(x0 >> 32) ^ c.w ^ rk.k[2 * r + 1], (uint32_t)x0);
    }
    out[call] = c;
  }
}

// ------------------------------------------------------------------------------------------------
// The product's uniform source ("gswm uniforms v4", restated in oracle/gs_oracle.py:gswm_uniforms).
//
// Every element gets a 23-bit integer m; its uniform is
//       u = v        if the element's bucket bit y is 1,        v = (m + 1/2) 2^-23
//       u = 1 - v    if y is 0                                  (again a grid point: m -> ~m)
// so that the reference's z = Phi^-1((u + y)/2) (gs_insert.py:64) is  +g(v)  resp.  -g(v)  with
// g(v) = sqrt(2) erfinv(v) the half-normal quantile: both buckets share ONE positive evaluation and the
// bucket only sets the sign.  (u is exactly uniform on the grid and independent of y either way, because
// m -> ~m is a bijection of the grid.)  v = (2m+1) 2^-24 is exact in fp32, which keeps small |z|
// accurate (rounding u breaks the 1e-6 tolerance otherwise).
//
// The m's come from Philox in groups: three Philox4x32 calls (384 bits) feed 16 elements -- the four
// float4 a thread stores in one "super-iteration".  float4 k = 0..2 take the top 23 bits of call k's
// four words; float4 3 is assembled from the otherwise unused low bytes of the three calls.
//
// v3: the outermost cell m = 2^23 - 1 (v in [1 - 2^-23, 1): probability 2^-23 per element, |z| >= 5.3) is SUBDIVIDED
// instead of being represented by its midpoint: a fourth Philox call (counter word 3 = call index 3, key word 1 + k
// for float4 k) yields a 28-bit m2 and v = 1 - (m2 + 1/2) 2^-51, so that the watermarked noise has the reference's
// support -- |z| up to 8.2095 = norm.ppf(1 - 2^-53), the largest value gs_insert.py:62-64 can produce for bucket
// bit 1 -- where the plain 23-bit grid stops at 5.42 (tail mass 6e-8 cut off).  The hot loop does not know about it: the
// rare tail block LOGS such an element in shared memory and the CTA refines its logged elements once, after its last
// latent (TopCellLog below, embed_kernel's epilogue).
//
// v4: the counter layout -- lane index in word 1, T = (latent, tile, super-iteration) split over words 1 and 3, the offset in
// words 0 and 2 (philox4x32_v4_calls3 above): same generator, same use of its output words, 26 instead of 35 wide multiplies
// per 16 elements.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}

// bit pattern of f = 1 + m 2^-23 from the top 23 bits of a word: one funnel shift drops the low 9 bits
// and shifts the exponent 0x7F in from the top
__device__ __forceinline__ uint32_t fbits_top23(uint32_t w) { return __funnelshift_r(w, 0x7Fu, 9); }

// ... and from the low bytes of three words: m = a.b0 | b.b0 << 8 | (c.b0 & 0x7F) << 16
__device__ __forceinline__ uint32_t fbits_low_bytes(uint32_t a, uint32_t b, uint32_t c) {
  const uint32_t ab = prmt(a, b, 0x0040u);            // bytes: a.b0, b.b0, a.b0, a.b0
  const uint32_t abc = prmt(ab, c, 0x0410u);          // bytes: a.b0, b.b0, c.b0, a.b0
  return (abc & 0x007FFFFFu) | 0x3F800000u;
}

__device__ __forceinline__ float horner_central(float x) {
  const float c[] = {GSWM_HNQ_CENTRAL_COEFFS};
  float p = c[0];
#pragma unroll
  for (int i = 1; i < (int)(sizeof(c) / sizeof(float)); ++i) p = fmaf(p, x, c[i]);
  return p;
}

__device__ __forceinline__ float horner_tail(float s) {
  const float c[] = {GSWM_HNQ_TAIL_COEFFS};
  float p = c[0];
#pragma unroll
  for (int i = 1; i < (int)(sizeof(c) / sizeof(float)); ++i) p = fmaf(p, s, c[i]);
  return p;
}

__device__ __forceinline__ float lg2_fast(float x) {   // bare MUFU.LG2 (argument is never subnormal)
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sqrt_fast(float x) {  // MUFU.SQRT
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float quantile_tail(float x) {           // m > M_SPLIT, ~0.12 % of elements
  return horner_tail(sqrt_fast(GSWM_HNQ_CSHIFT - x) - GSWM_HNQ_S0);
}

// One element: g(v) for f = 1 + m 2^-23 given as a bit pattern.
//   v = f - (1 - 2^-24) (exact),  x = c + lg2((1 - v)(1 + v - 2^-24)),  g = v P(x)  |  Q(sqrt(c - x))
__device__ __forceinline__ float halfnormal_quantile(uint32_t fbits) {
  const float f = __uint_as_float(fbits);
  const float v = f - __uint_as_float(0x3F7FFFFFu);
  const float t = fmaf(-GSWM_HNQ_KSCALE, v, GSWM_HNQ_KSCALE);
  const float x = lg2_fast(t * f);
  return (x >= GSWM_HNQ_XSPLIT) ? v * horner_central(x) : quantile_tail(x);
}

// Four elements as two packed pairs.  Blackwell's FFMA2 / FMUL2 / FADD2 (fma.rn.f32x2) do two fp32
// operations per issued instruction.  The central polynomial is evaluated unconditionally for all four
// elements and a single rarely-taken branch patches the elements that fell in the tail -- one
// compare-and-branch per float4 instead of a divergence region per element.
__device__ __forceinline__ float2 splat2(float a) { return make_float2(a, a); }

// -DGSWM_WHATIF_* builds (tools/whatif.sh) switch single ingredients of the embed kernel OFF to measure what each one costs
// in place; they produce wrong latents and exist for that measurement only.
#ifndef GSWM_WHATIF_POLY_SKIP
#define GSWM_WHATIF_POLY_SKIP 0      // Horner steps left out of the central polynomial (0 = the product)
#endif
__device__ __forceinline__ float2 horner_central2(float2 x) {
  const float c[] = {GSWM_HNQ_CENTRAL_COEFFS};
  float2 p = splat2(c[GSWM_WHATIF_POLY_SKIP]);
#pragma unroll
  for (int i = 1 + GSWM_WHATIF_POLY_SKIP; i < (int)(sizeof(c) / sizeof(float)); ++i) p = __ffma2_rn(p, x, splat2(c[i]));
  return p;
}

__device__ __forceinline__ void quantile_front2(uint32_t fa, uint32_t fb, float2& v, float2& x) {
  const float2 f = make_float2(__uint_as_float(fa), __uint_as_float(fb));
  v = __fadd2_rn(f, splat2(-__uint_as_float(0x3F7FFFFFu)));
  const float2 t = __ffma2_rn(v, splat2(-GSWM_HNQ_KSCALE), splat2(GSWM_HNQ_KSCALE));
  const float2 q = __fmul2_rn(t, f);
  x.x = lg2_fast(q.x);
  x.y = lg2_fast(q.y);
}

// Scalar (one element per instruction) evaluation of the same thing; used for half of the elements when
// GSWM_QUANTILE_MODE == 1 so that the FMA-heavy pipe (which also carries Philox's IMAD.WIDE) and the FMA-lite
// pipe share the polynomial work.
__device__ __forceinline__ void quantile_front1(uint32_t fa, float& v, float& x) {
  const float f = __uint_as_float(fa);
  v = f - __uint_as_float(0x3F7FFFFFu);
  const float t = fmaf(-GSWM_HNQ_KSCALE, v, GSWM_HNQ_KSCALE);
  x = lg2_fast(t * f);
}

#ifndef GSWM_QUANTILE_MODE
#define GSWM_QUANTILE_MODE 0      // 0: both pairs packed (FFMA2); 1: pair 0 packed, pair 1 scalar; 2: all scalar
#endif

// The outermost grid cell, refined (since uniforms v3).  w: 32 fresh Philox bits; m2 = w >> 4 (28 bits);
// p = P(|Z| > z) / 2 = (m2 + 1/2) 2^-52;  |z| = -Phi^-1(p) = R(sqrt(52 - lg2(m2 + 1/2)) - S1), R of degree 5 fitted like the
// others (tools/fit_halfnormal_quantile.py: 1.9e-7 relative in emulated fp32), |z| in [5.29, 8.21].  Reached about once
// per 8 M elements, from inside the rare tail block: ~15 instructions plus one Philox call there, nothing anywhere else.
__device__ __forceinline__ float top_cell_quantile(uint32_t w) {
  const float c[] = {GSWM_HNQ_FARTAIL_COEFFS};
  const float mf = (float)(w >> 4) + 0.5f;
  const float s = sqrt_fast(52.0f - lg2_fast(mf)) - GSWM_HNQ_S1;
  float p = c[0];
#pragma unroll
  for (int i = 1; i < (int)(sizeof(c) / sizeof(float)); ++i) p = fmaf(p, s, c[i]);
  return p;
}
struct NoTopCell {                                                   // test hook / callers without a counter: the cell keeps its midpoint
  static constexpr bool kLogs = false;
  __device__ __forceinline__ void operator()() const {}
};
// How the outermost cell gets refined without costing the hot loop anything.  Tried first, and measured (4096 SD-2.1
// latents, profiles/r02b_kbench.jsonl): the refinement inline in the patch block -- 69.6 us against 55.0 us, the loop body
// (2561 instructions, every patch block dragging a Philox call behind it) no longer fits the instruction cache, and what
// the hot path jumps over is fetched all the same; the rare formulas as out-of-line calls -- 58.2 .. 59.4 us, a call in
// the loop makes the compiler re-materialise its constants after every patch block; a per-thread mask tested after each
// super-iteration -- 58.5 us.  What costs nothing: the patch block -- which ~15 % of the warps' float4 enter anyway -- takes
// the minimum of its four x once more, and a float4 holding such an element (one in 2 M) appends its POSITION to a small
// list in SHARED memory; the CTA walks the list once, after its last latent, re-derives the float4's Philox counter from
// the position, finds the element(s) with m = 2^23 - 1 and refines them.  The list holds 64 float4 (a CTA sees one per
// 2 M float4 it produces; an overflowing one would keep the cell's midpoint -- still a sample of the right bucket).
constexpr uint32_t kTopCellSlots = 64;
struct TopCellLog {
  static constexpr bool kLogs = true;
  uint32_t* count;          // shared memory
  uint2* list;              // shared memory, kTopCellSlots entries: (latent within the launch, float4 within the tile) -- with the
                            // CTA's tile that says everything (which super-iteration, lane, position: hence which Philox counter
                            // and which address).  Both are values the loop has anyway (a uniform register and a constant + tid):
                            // logging the float4's ADDRESS instead kept two more registers alive across the polynomial
                            // (per-latent keys, 40 registers: 12 bytes of spills, 69.4 us against 63.1 us).
  uint32_t latent, f4;
  __device__ __forceinline__ void operator()() const {
    const uint32_t slot = atomicAdd(count, 1u);
    if (slot < kTopCellSlots) list[slot] = make_uint2(latent, f4);
  }
};
// The epilogue's half: float4 k of the super-iteration whose three Philox outputs are w[0..2] (recomputed by the caller
// from the address), `ctr3` = that super-iteration's counter with call index 3.  Every element of the float4 with
// m = 2^23 - 1 gets |z| from word j of Philox(ctr3) under key (seed_lo, seed_hi + k); the sign is the one it was stored with.
__device__ __forceinline__ void top_cell_fixup(float4* where, const uint4 (&w)[3], uint32_t k, uint4 ctr3, uint32_t seed_lo,
                                               uint32_t seed_hi) {
  const uint32_t w0[4] = {w[0].x, w[0].y, w[0].z, w[0].w}, w1[4] = {w[1].x, w[1].y, w[1].z, w[1].w}, w2[4] = {w[2].x, w[2].y, w[2].z, w[2].w};
  const uint4 r = philox4x32(ctr3, seed_lo, seed_hi + k);
  const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
  float* z = reinterpret_cast<float*>(where);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t fb = k == 0 ? fbits_top23(w0[j]) : k == 1 ? fbits_top23(w1[j]) : k == 2 ? fbits_top23(w2[j])
                                                                                         : fbits_low_bytes(w0[j], w1[j], w2[j]);
    if (fb == 0x3FFFFFFFu) z[j] = copysignf(top_cell_quantile(rr[j]), z[j]);
  }
}
constexpr float kTopCellX = -18.5f;

#ifndef GSWM_TAIL_INT
#define GSWM_TAIL_INT 1           // 1: central / tail decided on the integer f bits (before the MUFU); 0: on x (round-1 form)
#endif

// |z| of four elements from their f bit patterns; `sgn` (+-1 per element: +1 for bucket bit 1) gives z.
// kWarpUniformTail (callers whose whole warp is converged here): the rare tail patch is entered on a warp VOTE, i.e.
// through a uniform branch -- the warp executes the patch block when any lane needs it either way, but a uniform branch
// needs no reconvergence barrier (BSSY / BSYNC) around the common fall-through: 56.3 -> 55.3 us per 4096 SD-2.1 latents,
// bit-identical output.
// WHETHER the patch block is entered is decided on the integers: x < XSPLIT implies bits(f) > GSWM_HNQ_FGUARD_BITS (x is
// monotone in m up to rounding steps; the guard sits 64 grid points below the split, tools/fit_halfnormal_quantile.py
// checks the implication over all 2^23 m) -- one three-input integer maximum and two compares on values that exist
// before the MUFU.LG2 is even issued, instead of a float minimum tree behind it.  WHICH formula an element gets is its
// own x < XSPLIT, as before.
template <bool kWarpUniformTail = false, typename TopCell = NoTopCell>
__device__ __forceinline__ float4 bucket_quantile4_f32(uint32_t f0, uint32_t f1, uint32_t f2, uint32_t f3, float4 sgn,
                                                       const TopCell top = TopCell()) {
  float2 v01, v23, x01, x23, g01, g23;
#if GSWM_TAIL_INT && !defined(GSWM_WHATIF_NOTAIL)
  bool any_tail = __vimax3_u32(f0, f1, f2) > GSWM_HNQ_FGUARD_BITS || f3 > GSWM_HNQ_FGUARD_BITS;
  if (kWarpUniformTail) any_tail = __any_sync(0xFFFFFFFFu, any_tail);
#endif
#if GSWM_QUANTILE_MODE == 2
  quantile_front1(f0, v01.x, x01.x);
  quantile_front1(f1, v01.y, x01.y);
  g01 = make_float2(v01.x * horner_central(x01.x), v01.y * horner_central(x01.y));
#else
  quantile_front2(f0, f1, v01, x01);
  g01 = __fmul2_rn(v01, horner_central2(x01));
#endif
#if GSWM_QUANTILE_MODE >= 1
  quantile_front1(f2, v23.x, x23.x);
  quantile_front1(f3, v23.y, x23.y);
  g23 = make_float2(v23.x * horner_central(x23.x), v23.y * horner_central(x23.y));
#else
  quantile_front2(f2, f3, v23, x23);
  g23 = __fmul2_rn(v23, horner_central2(x23));
#endif
#ifndef GSWM_WHATIF_NOTAIL
#if !GSWM_TAIL_INT
  bool any_tail = fminf(fminf(x01.x, x01.y), fminf(x23.x, x23.y)) < GSWM_HNQ_XSPLIT;
  if (kWarpUniformTail) any_tail = __any_sync(0xFFFFFFFFu, any_tail);
#endif
  if (__builtin_expect(any_tail, 0)) {
    // per element the decision is its own x (the f bit patterns are dead by now: nothing is kept alive across the Horner
    // chains for this block); the outermost cell is x < -18.5 (m = 2^23 - 1: x = -19.0; m = 2^23 - 2: x = -17.4)
    if (x01.x < GSWM_HNQ_XSPLIT) g01.x = quantile_tail(x01.x);
    if (x01.y < GSWM_HNQ_XSPLIT) g01.y = quantile_tail(x01.y);
    if (x23.x < GSWM_HNQ_XSPLIT) g23.x = quantile_tail(x23.x);
    if (x23.y < GSWM_HNQ_XSPLIT) g23.y = quantile_tail(x23.y);
    // outermost cell somewhere in this float4: noted for the epilogue, midpoint for now
    if (TopCell::kLogs && fminf(fminf(x01.x, x01.y), fminf(x23.x, x23.y)) < kTopCellX) top();
  }
#endif
#ifdef GSWM_WHATIF_NOSIGN
  return make_float4(g01.x, g01.y, g23.x, g23.y);
#endif
#if GSWM_QUANTILE_MODE == 2
  return make_float4(g01.x * sgn.x, g01.y * sgn.y, g23.x * sgn.z, g23.y * sgn.w);
#else
  g01 = __fmul2_rn(g01, make_float2(sgn.x, sgn.y));
#if GSWM_QUANTILE_MODE == 1
  return make_float4(g01.x, g01.y, g23.x * sgn.z, g23.y * sgn.w);
#else
  g23 = __fmul2_rn(g23, make_float2(sgn.z, sgn.w));
  return make_float4(g01.x, g01.y, g23.x, g23.y);
#endif
#endif
}

// ------------------------------------------------------------------------------------------------
// fp64 path (injected uniforms): z = Phi^-1(p), p = (u + y) / 2 computed exactly as the reference
// does.  Same structure, natural log, degree-14 polynomials; relative error ~1e-12.
// ------------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ double horner64(const double (&c)[N], double x) {
  double p = c[0];
#pragma unroll
  for (int i = 1; i < N; ++i) p = fma(p, x, c[i]);
  return p;
}

__device__ inline double norm_ppf_f64(double p) {
  if (!(p >= 0.0 && p <= 1.0)) return CUDART_NAN;                   // outside [0, 1] (or NaN): scipy's ndtri returns nan
  // t = 2 min(p, 1-p) in (0,1]; both branches are exact in binary64 for p in [0,1]
  const bool upper = p > 0.5;
  const double t = upper ? 2.0 - 2.0 * p : 2.0 * p;
  if (!(t > 0.0)) return upper ? CUDART_INF : -CUDART_INF;          // ppf(0) = -inf, ppf(1) = +inf
  const double v = 1.0 - t;
  if (v == 0.0) return 0.0;                                         // ppf(0.5) = 0.0
  const double w = -log(t * (2.0 - t));
  double g;
  if (w > 38.0) {
    // Beyond anything a 53-bit uniform can reach (t < 2^-53).  Solve Q(z) = t/2 by fixed-point
    // iteration on the asymptotic series Q(z) = phi(z)/z sum_k (-1)^k (2k-1)!! / z^2k, k <= 8,
    // whose truncation error is < 1e-10 relative in z for z > 8.
    const double lq = log(0.5 * t);
    g = sqrt(-2.0 * lq);
#pragma unroll 1
    for (int it = 0; it < 12; ++it) {
      const double r = 1.0 / (g * g);
      const double series = 1.0 + r * (-1.0 + r * (3.0 + r * (-15.0 + r * (105.0 + r * (-945.0 + r * (10395.0 +
                            r * (-135135.0 + r * 2027025.0)))))));
      g = sqrt(-2.0 * (lq + log(g * 2.5066282746310002) - log(series)));
    }
  } else if (w < GSWM_HNQ64_WSPLIT) {
    const double c[] = {GSWM_HNQ64_CENTRAL_COEFFS};
    g = v * horner64(c, w - 0.5 * GSWM_HNQ64_WSPLIT);
  } else if (w < GSWM_HNQ64_WA) {
    const double c[] = {GSWM_HNQ64_TAILA_COEFFS};
    g = horner64(c, sqrt(w) - GSWM_HNQ64_SA0);
  } else {
    const double c[] = {GSWM_HNQ64_TAILB_COEFFS};
    g = horner64(c, sqrt(w) - GSWM_HNQ64_SB0);
  }
  return upper ? g : -g;
}

}  // namespace gswm
