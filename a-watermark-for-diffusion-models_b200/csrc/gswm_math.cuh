// Device-side building blocks of the Gaussian-Shading codec for sm_100a:
//   * ChaCha20 block function, one block per thread (integer add / xor / funnel-shift rotate);
//   * Philox4x32-10 counter-based generator (the product's uniform source);
//   * half-normal quantile g(v) = sqrt(2) erfinv(v) in fp32 (registers only: 1 MUFU.LG2 + FFMAs) and
//     fp64 (injected-uniform mode), coefficients from tools/fit_halfnormal_quantile.py.
// Nothing here touches memory.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "gswm_coeffs.inc"

namespace gswm {

// ------------------------------------------------------------------------------------------------
// ChaCha20 (DJB layout as `cryptography` exposes it: 64-bit block counter in words 12..13)
// Replaces Cipher(algorithms.ChaCha20(key, nonce)) at gs_insert.py:45-47 / extract.py:77-78.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int n) { return __funnelshift_l(x, x, n); }

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
#pragma unroll
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
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11).  Takes the place of np.random.uniform at
// gs_insert.py:62; restated for the oracle in oracle/gs_oracle.py:philox4x32.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
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

// ------------------------------------------------------------------------------------------------
// fp32 half-normal quantile with the bucket fused in.
//
//   w       raw 32-bit random word; the uniform is u = ((w >> 9) + 0.5) * 2^-23
//   flip    0x00000000 if the bucket bit y is 1, 0xFFFFFFFF if y is 0
//   returns Phi^-1((u + y) / 2)  =  +g(u) for y = 1,  -g(1 - u) for y = 0           (gs_insert.py:64)
//
// 1 - u is again on the grid (complement the 23 bits), so both buckets share one positive-half
// evaluation and the sign is OR-ed in at the end.  v = (2m+1) 2^-24 is exact in fp32, which is what
// keeps small |z| accurate (rounding u breaks the 1e-6 tolerance otherwise).
// ------------------------------------------------------------------------------------------------
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

// Front end shared by both branches: v = (2m+1) 2^-24 exactly, x = c + lg2((1 - v)(1 + v - 2^-24)).
__device__ __forceinline__ void quantile_front(uint32_t w, uint32_t flip, float& v, float& x) {
  // f = 1 + m 2^-23 with m = 23 random bits (complemented for bucket 0): one funnel shift drops the
  // low 9 bits and shifts the exponent 0x7F in from the top.
  const float f = __uint_as_float(__funnelshift_r(w ^ flip, 0x7Fu, 9));
  v = f - __uint_as_float(0x3F7FFFFFu);                             // exact: f - (1 - 2^-24)
  const float t = fmaf(-GSWM_HNQ_KSCALE, v, GSWM_HNQ_KSCALE);       // K (1 - v), K = 2^c
  x = lg2_fast(t * f);
}

__device__ __forceinline__ float quantile_tail(float x) {           // x < XSPLIT, ~0.3 % of elements
  return horner_tail(sqrt_fast(GSWM_HNQ_CSHIFT - x) - GSWM_HNQ_S0);
}

__device__ __forceinline__ float apply_sign(float g, uint32_t flip) {
  return __uint_as_float(__float_as_uint(g) | (flip & 0x80000000u));
}

// One element (used by tests / small paths).
__device__ __forceinline__ float bucket_quantile_f32(uint32_t w, uint32_t flip) {
  float v, x;
  quantile_front(w, flip, v, x);
  const float g = (x >= GSWM_HNQ_XSPLIT) ? v * horner_central(x) : quantile_tail(x);
  return apply_sign(g, flip);
}

// Four elements at once: the central polynomial is evaluated unconditionally for all four, and a
// single rarely-taken branch patches the elements that fell in the tail -- one compare-and-branch
// per float4 instead of a divergence region per element.
__device__ __forceinline__ float4 bucket_quantile4_f32(uint4 w, uint32_t f0, uint32_t f1, uint32_t f2, uint32_t f3) {
  float v0, v1, v2, v3, x0, x1, x2, x3;
  quantile_front(w.x, f0, v0, x0);
  quantile_front(w.y, f1, v1, x1);
  quantile_front(w.z, f2, v2, x2);
  quantile_front(w.w, f3, v3, x3);
  float g0 = v0 * horner_central(x0);
  float g1 = v1 * horner_central(x1);
  float g2 = v2 * horner_central(x2);
  float g3 = v3 * horner_central(x3);
  if (fminf(fminf(x0, x1), fminf(x2, x3)) < GSWM_HNQ_XSPLIT) {
    if (x0 < GSWM_HNQ_XSPLIT) g0 = quantile_tail(x0);
    if (x1 < GSWM_HNQ_XSPLIT) g1 = quantile_tail(x1);
    if (x2 < GSWM_HNQ_XSPLIT) g2 = quantile_tail(x2);
    if (x3 < GSWM_HNQ_XSPLIT) g3 = quantile_tail(x3);
  }
  return make_float4(apply_sign(g0, f0), apply_sign(g1, f1), apply_sign(g2, f2), apply_sign(g3, f3));
}

// Flip masks of four consecutive elements from their bucket-bit nibble (bit 3 = first element):
// spread the nibble to bit 7 of each byte lane with one multiply, then PRMT in sign-replicate mode
// turns byte lane j into 0x00000000 / 0xFFFFFFFF.  flip = ~(bit ? ~0 : 0).
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}
__device__ __forceinline__ void nibble_flip_masks(uint32_t nib, uint32_t& f0, uint32_t& f1, uint32_t& f2, uint32_t& f3) {
  // bit k of nib -> bit 7 of byte lane k ; inverted so that a SET bucket bit gives flip = 0
  const uint32_t lanes = ~((nib & 0xFu) * (0x00204081u << 7)) & 0x80808080u;
  f0 = prmt(lanes, 0u, 0xBBBBu);   // element 0 = nibble bit 3 = byte lane 3, sign-replicated
  f1 = prmt(lanes, 0u, 0xAAAAu);
  f2 = prmt(lanes, 0u, 0x9999u);
  f3 = prmt(lanes, 0u, 0x8888u);
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

__device__ __forceinline__ double norm_ppf_f64(double p) {
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
