/* A C host of libgswm.so with nothing but the CUDA runtime underneath -- no Python, no torch: what a maintainer of a C / C++
 * pipeline links against.  Embeds a batch (gs_insert.py:8-66 for every latent), decodes it (extract.py:72-110), checks the
 * counters and the decoded message, then the host-buffer pipe on the same job.  Known answers: the ChaCha20 keystream of the
 * default key / nonce starts 61 08 48 b7, the encrypted tile starts 0d 7c 20 d2.
 *   gcc -std=c99 -Iinclude -I$CUDA/include tests/c_abi/roundtrip.c -L<libdir> -lgswm -L$CUDA/lib64 -lcudart -o roundtrip */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gswm.h"

#define CHECK(x)                                                              \
  do {                                                                        \
    int rc_ = (int)(x);                                                       \
    if (rc_ != 0) {                                                           \
      printf("FAILED %s -> %d (%s)\n", #x, rc_, gswm_strerror(rc_));          \
      return 1;                                                               \
    }                                                                         \
  } while (0)

static int hexval(char c) { return c <= '9' ? c - '0' : c - 'a' + 10; }
static void unhex(const char* s, uint8_t* out) {
  size_t i, n = strlen(s) / 2;
  for (i = 0; i < n; ++i) out[i] = (uint8_t)(hexval(s[2 * i]) * 16 + hexval(s[2 * i + 1]));
}

int main(void) {
  const int64_t B = 300, N = 4 * 64 * 64;
  const int L = 256;
  uint8_t km[32 + 16 + 32];
  uint8_t *d_km, *d_msgs, *d_flags, *d_ks;
  float* d_z;
  int64_t* d_ctr;
  int64_t ctr[GSWM_N_COUNTERS];
  uint8_t* msgs = (uint8_t*)malloc((size_t)B * 32);
  uint8_t ks[64];
  gswm_job job;
  int64_t i;

  if (gswm_abi_version() != GSWM_ABI_VERSION) return 1;
  unhex("5822ff9cce6772f714192f43863f6bad1bf54b78326973897e6b66c3186b77a7", km);       /* README.md:61 */
  unhex("05072fd1c2265f6f2e2a4080a2bfbdd8", km + 32);                                    /* README.md:67 */
  memset(km + 48, 0, 32);
  memcpy(km + 48, "lthero", 6);                                                          /* gs_insert.py:9-20 */

  CHECK(cudaSetDevice(0));
  CHECK(cudaMalloc((void**)&d_km, sizeof km));
  CHECK(cudaMalloc((void**)&d_z, (size_t)B * N * sizeof(float)));
  CHECK(cudaMalloc((void**)&d_msgs, (size_t)B * 32));
  CHECK(cudaMalloc((void**)&d_flags, (size_t)B));
  CHECK(cudaMalloc((void**)&d_ctr, sizeof ctr));
  CHECK(cudaMalloc((void**)&d_ks, 64));
  CHECK(cudaMemcpy(d_km, km, sizeof km, cudaMemcpyHostToDevice));
  CHECK(cudaMemset(d_ctr, 0, sizeof ctr));

  CHECK(gswm_chacha20_keystream(d_km, d_km + 32, 1, 64, d_ks, NULL));
  CHECK(cudaMemcpy(ks, d_ks, 64, cudaMemcpyDeviceToHost));
  if (ks[0] != 0x61 || ks[1] != 0x08 || ks[2] != 0x48 || ks[3] != 0xb7) { printf("keystream mismatch\n"); return 1; }

  job.n_latents = B; job.n_elems = N; job.msg_bits = L; job.flags = 0;
  job.d_keys = d_km; job.d_nonces = d_km + 32; job.d_msgs = d_km + 48;
  CHECK(gswm_embed(&job, 0x5EEDull, 0, 0, d_z, NULL));
  CHECK(gswm_extract(&job, d_z, GSWM_F32, d_msgs, NULL, NULL, d_flags, d_ctr, NULL));
  CHECK(cudaDeviceSynchronize());
  CHECK(cudaMemcpy(ctr, d_ctr, sizeof ctr, cudaMemcpyDeviceToHost));
  CHECK(cudaMemcpy(msgs, d_msgs, (size_t)B * 32, cudaMemcpyDeviceToHost));
  if (ctr[GSWM_CTR_MATCHED_BITS] != B * L || ctr[GSWM_CTR_TOTAL_BITS] != B * L || ctr[GSWM_CTR_EXACT_MSGS] != B ||
      ctr[GSWM_CTR_TOTAL_MSGS] != B || ctr[GSWM_CTR_NAN_LATENTS] != 0 || ctr[GSWM_CTR_RANGE_LATENTS] != 0) {
    printf("counters wrong: %lld %lld %lld %lld\n", (long long)ctr[0], (long long)ctr[1], (long long)ctr[2], (long long)ctr[3]);
    return 1;
  }
  for (i = 0; i < B; ++i)
    if (memcmp(msgs + 32 * i, km + 48, 32) != 0) { printf("latent %lld decodes to another message\n", (long long)i); return 1; }

  /* the sign pattern of a latent IS the encrypted tile (gs_insert.py:49,64): first byte 0x0d = 0000 1101 */
  {
    float z8[8];
    const int want[8] = {0, 0, 0, 0, 1, 1, 0, 1};
    CHECK(cudaMemcpy(z8, d_z, sizeof z8, cudaMemcpyDeviceToHost));
    for (i = 0; i < 8; ++i)
      if ((z8[i] >= 0.0f) != want[i]) { printf("bucket bit %lld wrong\n", (long long)i); return 1; }
  }

  /* the same job through the host-buffer layer: latents come back to host memory and decode from there */
  {
    gswm_pipe* pipe;
    gswm_host_job hj;
    float* h_z = (float*)malloc((size_t)B * N * sizeof(float));
    uint8_t* h_flags = (uint8_t*)malloc((size_t)B);
    float first[4];
    hj.n_latents = B; hj.n_elems = N; hj.msg_bits = L; hj.flags = 0;
    hj.h_keys = km; hj.h_nonces = km + 32; hj.h_msgs = km + 48;
    CHECK(gswm_pipe_create(&pipe, 0, N, 128));
    CHECK(gswm_pipe_embed(pipe, &hj, 0x5EEDull, 0, 0, h_z));
    CHECK(cudaMemcpy(first, d_z, sizeof first, cudaMemcpyDeviceToHost));
    if (memcmp(first, h_z, sizeof first) != 0) { printf("pipe and device path disagree\n"); return 1; }
    CHECK(gswm_pipe_extract(pipe, &hj, h_z, GSWM_F32, msgs, NULL, NULL, h_flags, ctr));
    if (ctr[GSWM_CTR_EXACT_MSGS] != B || ctr[GSWM_CTR_MATCHED_BITS] != B * L) { printf("pipe counters wrong\n"); return 1; }
    gswm_pipe_destroy(pipe);
    free(h_z); free(h_flags);
  }
  printf("roundtrip ok: %lld latents, %lld kernel launches\n", (long long)B, (long long)gswm_launch_count());
  free(msgs);
  return 0;
}
