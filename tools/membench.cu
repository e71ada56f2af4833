// Practical HBM ceilings on this B200 for the access patterns of the codec kernels: read-only stream
// (extract), write-only stream (embed), copy (the pattern MEASURED_PEAKS.json's hbm_gbs was taken with).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/membench tools/membench.cu && ./tools/membench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_read(const uint4* __restrict__ in, size_t n, uint32_t* sink) {
  uint32_t acc = 0;
  size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * 256;
  for (; i + 7 * stride < n; i += 8 * stride) {
    uint4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldcs(in + i + k * stride);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc ^= v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
  }
  for (; i < n; i += stride) { uint4 v = __ldcs(in + i); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
  if (acc == 0x12345678u) *sink = acc;
}
__global__ void __launch_bounds__(256) k_write(uint4* __restrict__ out, size_t n, uint32_t seed) {
  size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * 256;
  const uint4 v = make_uint4(seed, seed + 1, seed + 2, (uint32_t)i);
  for (; i < n; i += stride) out[i] = v;
}
__global__ void __launch_bounds__(256) k_copy(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * 256;
  for (; i + 3 * stride < n; i += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = __ldcs(in + i + k * stride);
#pragma unroll
    for (int k = 0; k < 4; ++k) out[i + k * stride] = v[k];
  }
  for (; i < n; i += stride) out[i] = in[i];
}

template <typename F>
float time_us(F f, int reps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) f();
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms / reps < best) best = ms / reps;
  }
  return best * 1e3f;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  uint32_t* sink; cudaMalloc(&sink, 4);
  for (size_t mb : {268ull, 1074ull}) {
    const size_t bytes = mb * 1000000ull / 16 * 16, n = bytes / 16;
    uint4 *a, *b; cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMemset(a, 1, bytes); cudaMemset(b, 2, bytes);
    for (int cps : {4, 8, 16}) {
      const int grid = p.multiProcessorCount * cps;
      float tr = time_us([&] { k_read<<<grid, 256>>>(a, n, sink); }, 50);
      float tw = time_us([&] { k_write<<<grid, 256>>>(b, n, 7); }, 50);
      float tc = time_us([&] { k_copy<<<grid, 256>>>(a, b, n); }, 50);
      printf("{\"MB\": %zu, \"ctas_per_sm\": %d, \"read_GBps\": %.0f, \"write_GBps\": %.0f, \"copy_GBps\": %.0f, \"read_us\": %.1f, \"write_us\": %.1f}\n",
             mb, cps, bytes / tr / 1e3, bytes / tw / 1e3, 2.0 * bytes / tc / 1e3, tr, tw);
    }
    cudaFree(a); cudaFree(b);
  }
  return 0;
}
