// Issue-rate microbenchmarks for the instruction classes the embed kernel is made of (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench tools/microbench.cu && ./microbench
// Prints warp-instructions per cycle per SM sub-partition (SMSP) for each class and a few mixes; these are
// the denominators for the embed kernel's issue-utilisation figure (MEASURED_PEAKS.json only has HBM and
// tensor peaks).  Every kernel runs 8 independent dependency chains per thread so latency is hidden.
#include <algorithm>
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 4096;
constexpr int kThreads = 1024;   // one CTA per SM: 32 warps = 8 per SMSP, so a CTA's cycle count is its SM's busy time

#define CHAINS8(BODY) BODY(0) BODY(1) BODY(2) BODY(3) BODY(4) BODY(5) BODY(6) BODY(7)

template <int kind>
__global__ void __launch_bounds__(kThreads) bench(uint32_t* out, uint32_t seed, long long* cycles) {
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
    if (kind == 0) {          // FFMA
#define B(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(c), "f"(g[i]));
      CHAINS8(B)
#undef B
    } else if (kind == 1) {   // FFMA2
#define B(i) asm volatile("{.reg .b64 t,u,v; mov.b64 t,{%0,%1}; mov.b64 u,{%2,%2}; mov.b64 v,{%3,%3}; fma.rn.f32x2 t,t,u,v; mov.b64 {%0,%1},t;}" : "+f"(f[i]), "+f"(g[i]) : "f"(c), "f"(c));
      CHAINS8(B)
#undef B
    } else if (kind == 2) {   // IMAD.WIDE.U32
#define B(i) asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(w[i]) : "r"((uint32_t)w[i]));
      CHAINS8(B)
#undef B
    } else if (kind == 3) {   // IMAD (32-bit mad.lo)
#define B(i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
      CHAINS8(B)
#undef B
    } else if (kind == 4) {   // LOP3
#define B(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
      CHAINS8(B)
#undef B
    } else if (kind == 5) {   // SHF (funnel)
#define B(i) asm volatile("shf.r.wrap.b32 %0, %0, %1, 9;" : "+r"(a[i]) : "r"(b[i]));
      CHAINS8(B)
#undef B
    } else if (kind == 6) {   // MUFU.LG2
#define B(i) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      CHAINS8(B)
#undef B
    } else if (kind == 7) {   // mix: IMAD.WIDE + LOP3 alternating (Philox round shape)
#define B(i) asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(w[i]) : "r"((uint32_t)w[i] ^ a[i])); \
             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"((uint32_t)(w[i] >> 32)), "r"(seed));
      CHAINS8(B)
#undef B
    } else if (kind == 8) {   // mix: FFMA2 + LOP3
#define B(i) asm volatile("{.reg .b64 t,u,v; mov.b64 t,{%0,%1}; mov.b64 u,{%2,%2}; mov.b64 v,{%3,%3}; fma.rn.f32x2 t,t,u,v; mov.b64 {%0,%1},t;}" : "+f"(f[i]), "+f"(g[i]) : "f"(c), "f"(c)); \
             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
      CHAINS8(B)
#undef B
    } else if (kind == 9) {   // mix: FFMA + LOP3
#define B(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(c), "f"(g[i])); \
             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
      CHAINS8(B)
#undef B
    } else if (kind == 10) {  // mix: FFMA + IMAD.WIDE + LOP3
#define B(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(c), "f"(g[i])); \
             asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(w[i]) : "r"((uint32_t)w[i])); \
             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
      CHAINS8(B)
#undef B
    } else if (kind == 11) {  // PRMT
#define B(i) asm volatile("prmt.b32 %0, %0, %1, 0x8888;" : "+r"(a[i]) : "r"(b[i]));
      CHAINS8(B)
#undef B
    } else if (kind == 12) {  // mul.hi.u32 (IMAD.HI)
#define B(i) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
      CHAINS8(B)
#undef B
    } else if (kind == 13) {  // FFMA2 + IMAD.WIDE
#define B(i) asm volatile("{.reg .b64 t,u,v; mov.b64 t,{%0,%1}; mov.b64 u,{%2,%2}; mov.b64 v,{%3,%3}; fma.rn.f32x2 t,t,u,v; mov.b64 {%0,%1},t;}" : "+f"(f[i]), "+f"(g[i]) : "f"(c), "f"(c)); \
             asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(w[i]) : "r"((uint32_t)w[i]));
      CHAINS8(B)
#undef B
    } else if (kind == 14) {  // IADD3
#define B(i) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(a[i]));
      CHAINS8(B)
#undef B
    } else if (kind == 15) {  // FFMA + FFMA2 interleaved
#define B(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(c), "f"(c));
      CHAINS8(B)
#undef B
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= a[i] ^ b[i] ^ __float_as_uint(f[i]) ^ __float_as_uint(g[i]) ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
  out[blockIdx.x * kThreads + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int kind>
void run(const char* name, int instr_per_chain_step, uint32_t* d_out, long long* d_cycles, int sms) {
  bench<kind><<<sms, kThreads>>>(d_out, 1, d_cycles);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bench<kind><<<sms, kThreads>>>(d_out, 2, d_cycles);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  static long long cyc[1024];
  cudaMemcpy(cyc, d_cycles, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  std::sort(cyc, cyc + sms);
  const long long med = cyc[sms / 2];
  // warp-instructions issued per SMSP: 32 warps per SM = 8 per SMSP
  const double warp_instr_per_smsp = (double)kIters * 8 * instr_per_chain_step * (kThreads / 32 / 4);
  printf("{\"bench\": \"%s\", \"warp_instr_per_clk_per_smsp\": %.4f, \"cycles_median\": %lld, \"cycles_max\": %lld, \"ms\": %.4f, \"ghz\": %.3f}\n",
         name, warp_instr_per_smsp / (double)med, med, cyc[sms - 1], ms, (double)cyc[sms - 1] / (ms * 1e6));
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  uint32_t* d_out;
  long long* d_cycles;
  cudaMalloc(&d_out, (size_t)p.multiProcessorCount * kThreads * 4);
  cudaMalloc(&d_cycles, 8 * 1024);
  printf("{\"device\": \"%s\", \"sms\": %d}\n", p.name, p.multiProcessorCount);
  run<0>("FFMA", 1, d_out, d_cycles, p.multiProcessorCount);
  run<1>("FFMA2", 1, d_out, d_cycles, p.multiProcessorCount);
  run<2>("IMAD.WIDE.U32", 1, d_out, d_cycles, p.multiProcessorCount);
  run<3>("IMAD", 1, d_out, d_cycles, p.multiProcessorCount);
  run<12>("IMAD.HI", 1, d_out, d_cycles, p.multiProcessorCount);
  run<4>("LOP3", 1, d_out, d_cycles, p.multiProcessorCount);
  run<5>("SHF", 1, d_out, d_cycles, p.multiProcessorCount);
  run<11>("PRMT", 1, d_out, d_cycles, p.multiProcessorCount);
  run<14>("IADD", 2, d_out, d_cycles, p.multiProcessorCount);
  run<6>("MUFU.LG2", 1, d_out, d_cycles, p.multiProcessorCount);
  run<7>("mix IMAD.WIDE+LOP3", 2, d_out, d_cycles, p.multiProcessorCount);
  run<8>("mix FFMA2+LOP3", 2, d_out, d_cycles, p.multiProcessorCount);
  run<9>("mix FFMA+LOP3", 2, d_out, d_cycles, p.multiProcessorCount);
  run<10>("mix FFMA+IMAD.WIDE+LOP3", 3, d_out, d_cycles, p.multiProcessorCount);
  run<13>("mix FFMA2+IMAD.WIDE", 2, d_out, d_cycles, p.multiProcessorCount);
  return 0;
}
