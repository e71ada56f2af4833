// Issue-rate microbenchmarks for the instruction classes the embed kernel is made of (sm_100a) -- the denominators of
// bench.py's issue-utilisation figure, measured in the same run on the same GPU (MEASURED_PEAKS.json only has HBM and
// tensor peaks).  gswm_debug_issue_rate(kind) launches one 1024-thread CTA per SM (8 warps per
// SM sub-partition, 8 independent dependency chains per thread, so latency is hidden) and returns warp instructions
// issued per clock per sub-partition, from clock64() inside the kernel, plus the SM clock the run saw.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <vector>

#include "../../include/gswm.h"
#include "gswm_internal.h"

namespace {

constexpr int kIters = 1024;
constexpr int kMbThreads = 1024;

// 8 independent chains, 4 times over per loop iteration: 32 measured instructions (64 for the two-instruction mix) against
// ~3 of loop overhead (the first version had 8 per iteration and read 0.70 warp-inst/clk for a one-cycle FFMA: 8 / 11)
#define CHAINS8_ONCE(BODY) BODY(0) BODY(1) BODY(2) BODY(3) BODY(4) BODY(5) BODY(6) BODY(7)
#define CHAINS8(BODY) CHAINS8_ONCE(BODY) CHAINS8_ONCE(BODY) CHAINS8_ONCE(BODY) CHAINS8_ONCE(BODY)
constexpr int kRepeat = 4;

template <int kind>
__global__ void __launch_bounds__(kMbThreads) issue_rate_kernel(uint32_t* out, uint32_t seed, long long* cycles) {
  uint32_t a[8], b[8];
  float f[8], g[8];
  unsigned long long w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = seed + threadIdx.x * 8 + i; b[i] = a[i] * 2654435761u;
    f[i] = 1.0f + (float)i * 1e-3f + seed * 1e-9f; g[i] = 0.5f + threadIdx.x * 1e-6f; w[i] = a[i];
  }
  const float c = 1.0001f + seed * 1e-9f;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
    if (kind == GSWM_ISSUE_FFMA2) {            // packed fp32x2 FMA: two FMAs per lane per instruction
#define B(i) asm volatile("{.reg .b64 t,u,v; mov.b64 t,{%0,%1}; mov.b64 u,{%2,%2}; mov.b64 v,{%3,%3}; fma.rn.f32x2 t,t,u,v; mov.b64 {%0,%1},t;}" : "+f"(f[i]), "+f"(g[i]) : "f"(c), "f"(c));
      CHAINS8(B)
#undef B
    } else if (kind == GSWM_ISSUE_IMAD_WIDE) { // 32 x 32 -> 64 multiply (Philox)
      // w = lo(w) * M + w: both halves of every product stay live and every multiply depends on the one before (a chain of plain mul.wide whose high
      // halves are overwritten unread gets narrowed to 32-bit IMADs by ptxas and reads 1.5 cycles instead of 4)
#define B(i) asm volatile("{.reg .b32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, 0xD2511F53, %0;}" : "+l"(w[i]));
      CHAINS8(B)
#undef B
    } else if (kind == GSWM_ISSUE_LOP3) {      // three-input logic (Philox xor, bit assembly)
#define B(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
      CHAINS8(B)
#undef B
    } else if (kind == GSWM_ISSUE_MUFU) {      // MUFU.LG2
#define B(i) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      CHAINS8(B)
#undef B
    } else if (kind == GSWM_ISSUE_FFMA_IMM) {  // scalar FMA with immediate operands: the single-issue peak
#define B(i) asm volatile("fma.rn.f32 %0, %0, 0f3F800347, 0f3A83126F;" : "+f"(f[i]));
      CHAINS8(B)
#undef B
    } else if (kind == GSWM_ISSUE_PHILOX_MIX) { // IMAD.WIDE + one three-input LOP3 alternating, each feeding the other: the shape of a Philox round
#define B(i) asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(w[i]) : "r"(a[i])); \
             asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a[i]) : "r"((uint32_t)w[i]), "r"((uint32_t)(w[i] >> 32)), "r"(seed));
      CHAINS8(B)
#undef B
    } else if (kind == GSWM_ISSUE_FFMA2_LOP3) {  // packed FMA next to independent logic: do the FMA and ALU pipes overlap?
#define B(i) asm volatile("{.reg .b64 t,u,v; mov.b64 t,{%0,%1}; mov.b64 u,{%2,%2}; mov.b64 v,{%3,%3}; fma.rn.f32x2 t,t,u,v; mov.b64 {%0,%1},t;}" : "+f"(f[i]), "+f"(g[i]) : "f"(c), "f"(c)); \
             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
      CHAINS8(B)
#undef B
    } else if (kind == GSWM_ISSUE_FFMA_IMAD_WIDE) {  // scalar FMA next to an independent wide multiply: FMA-lite under the FMA-heavy pipe?
#define B(i) asm volatile("fma.rn.f32 %0, %0, 0f3F800347, 0f3A83126F;" : "+f"(f[i])); \
             asm volatile("{.reg .b32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, 0xD2511F53, %0;}" : "+l"(w[i]));
      CHAINS8(B)
#undef B
    } else if (kind == GSWM_ISSUE_FFMA2_IMAD_WIDE) { // packed FMA next to an independent wide multiply
#define B(i) asm volatile("{.reg .b64 t,u,v; mov.b64 t,{%0,%1}; mov.b64 u,{%2,%2}; mov.b64 v,{%3,%3}; fma.rn.f32x2 t,t,u,v; mov.b64 {%0,%1},t;}" : "+f"(f[i]), "+f"(g[i]) : "f"(c), "f"(c)); \
             asm volatile("{.reg .b32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, 0xD2511F53, %0;}" : "+l"(w[i]));
      CHAINS8(B)
#undef B
    } else if (kind == GSWM_ISSUE_LOP3_IMAD_WIDE) {  // independent logic next to an independent wide multiply
#define B(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed)); \
             asm volatile("{.reg .b32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, 0xD2511F53, %0;}" : "+l"(w[i]));
      CHAINS8(B)
#undef B
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= a[i] ^ b[i] ^ __float_as_uint(f[i]) ^ __float_as_uint(g[i]) ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
  out[blockIdx.x * kMbThreads + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int kind>
int run_kind(int per_step, int sms, uint32_t* d_out, long long* d_cycles, double* rate, double* ghz) {
  issue_rate_kernel<kind><<<sms, kMbThreads>>>(d_out, 1, d_cycles);              // warm-up
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  issue_rate_kernel<kind><<<sms, kMbThreads>>>(d_out, 2, d_cycles);
  cudaEventRecord(e1);
  cudaError_t err = cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (err != cudaSuccess) return (int)err;
  std::vector<long long> cyc(sms);
  if ((err = cudaMemcpy(cyc.data(), d_cycles, sizeof(long long) * sms, cudaMemcpyDeviceToHost)) != cudaSuccess) return (int)err;
  std::sort(cyc.begin(), cyc.end());
  const double warp_instr_per_smsp = (double)kIters * 8 * kRepeat * per_step * (kMbThreads / 32 / 4);
  *rate = warp_instr_per_smsp / (double)cyc[sms / 2];
  *ghz = (double)cyc[sms - 1] / ((double)ms * 1e6);
  gswm::count_launch(); gswm::count_launch();
  return GSWM_OK;
}

}  // namespace

extern "C" int gswm_debug_issue_rate(int32_t kind, double* warp_inst_per_clk_per_smsp, double* sm_ghz) {
  if (!warp_inst_per_clk_per_smsp || !sm_ghz) return GSWM_E_NULL;
  int dev = 0, sms = 0;
  cudaError_t e;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return (int)e;
  if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
  uint32_t* d_out = nullptr;
  long long* d_cycles = nullptr;
  if ((e = cudaMalloc(&d_out, (size_t)sms * kMbThreads * 4)) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&d_cycles, sizeof(long long) * sms)) != cudaSuccess) { cudaFree(d_out); return (int)e; }
  int rc;
  switch (kind) {
    case GSWM_ISSUE_FFMA2: rc = run_kind<GSWM_ISSUE_FFMA2>(1, sms, d_out, d_cycles, warp_inst_per_clk_per_smsp, sm_ghz); break;
    case GSWM_ISSUE_IMAD_WIDE: rc = run_kind<GSWM_ISSUE_IMAD_WIDE>(1, sms, d_out, d_cycles, warp_inst_per_clk_per_smsp, sm_ghz); break;
    case GSWM_ISSUE_LOP3: rc = run_kind<GSWM_ISSUE_LOP3>(1, sms, d_out, d_cycles, warp_inst_per_clk_per_smsp, sm_ghz); break;
    case GSWM_ISSUE_MUFU: rc = run_kind<GSWM_ISSUE_MUFU>(1, sms, d_out, d_cycles, warp_inst_per_clk_per_smsp, sm_ghz); break;
    case GSWM_ISSUE_FFMA_IMM: rc = run_kind<GSWM_ISSUE_FFMA_IMM>(1, sms, d_out, d_cycles, warp_inst_per_clk_per_smsp, sm_ghz); break;
    case GSWM_ISSUE_PHILOX_MIX: rc = run_kind<GSWM_ISSUE_PHILOX_MIX>(2, sms, d_out, d_cycles, warp_inst_per_clk_per_smsp, sm_ghz); break;
    case GSWM_ISSUE_FFMA2_LOP3: rc = run_kind<GSWM_ISSUE_FFMA2_LOP3>(2, sms, d_out, d_cycles, warp_inst_per_clk_per_smsp, sm_ghz); break;
    case GSWM_ISSUE_FFMA_IMAD_WIDE: rc = run_kind<GSWM_ISSUE_FFMA_IMAD_WIDE>(2, sms, d_out, d_cycles, warp_inst_per_clk_per_smsp, sm_ghz); break;
    case GSWM_ISSUE_FFMA2_IMAD_WIDE: rc = run_kind<GSWM_ISSUE_FFMA2_IMAD_WIDE>(2, sms, d_out, d_cycles, warp_inst_per_clk_per_smsp, sm_ghz); break;
    case GSWM_ISSUE_LOP3_IMAD_WIDE: rc = run_kind<GSWM_ISSUE_LOP3_IMAD_WIDE>(2, sms, d_out, d_cycles, warp_inst_per_clk_per_smsp, sm_ghz); break;
    default: rc = GSWM_E_RANGE;
  }
  cudaFree(d_out);
  cudaFree(d_cycles);
  return rc;
}
