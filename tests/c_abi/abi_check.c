/* Plain-C consumer of include/gswm.h: proves the header is C (not C++), that libgswm.so links from C, and that argument
 * validation happens before any CUDA call (so this runs without a GPU).  Built and run by tests/test_host_cpu.py. */
#include <stdio.h>
#include <string.h>

#include "gswm.h"

#define CHECK(cond)                                                   \
  do {                                                                \
    if (!(cond)) {                                                    \
      fprintf(stderr, "abi_check: %s failed (line %d)\n", #cond, __LINE__); \
      return 1;                                                       \
    }                                                                 \
  } while (0)

int main(void) {
  gswm_job job;
  gswm_host_job hjob;
  static unsigned char buf[64];
  float out[4];
  memset(&job, 0, sizeof job);
  memset(&hjob, 0, sizeof hjob);
  CHECK(gswm_abi_version() == GSWM_ABI_VERSION);
  CHECK(strcmp(gswm_strerror(GSWM_OK), "success") == 0);
  CHECK(strncmp(gswm_strerror(GSWM_E_MSGLEN), "gswm:", 5) == 0);
  CHECK(gswm_philox_rounds() >= 7);
  CHECK(sizeof(gswm_job) == sizeof(gswm_host_job));
  CHECK(GSWM_N_COUNTERS == 6 && GSWM_CTR_TOTAL_MSGS == 3 && GSWM_CTR_RANGE_LATENTS == 5);
  CHECK(GSWM_N_COUNTERS <= GSWM_COMM_MAX_VALUES && GSWM_JOB_PER_LATENT == 1);
  /* null / shape / message-length / dtype errors, in the order the entry points check them */
  CHECK(gswm_embed(NULL, 0, 0, 0, out, NULL) == GSWM_E_NULL);
  job.n_latents = 1; job.n_elems = 1002; job.msg_bits = 32;
  job.d_keys = buf; job.d_nonces = buf; job.d_msgs = buf;
  CHECK(gswm_embed(&job, 0, 0, 0, (float*)buf, NULL) == GSWM_E_SHAPE);
  job.n_elems = 16384; job.msg_bits = 48;
  CHECK(gswm_embed(&job, 0, 0, 0, (float*)buf, NULL) == GSWM_E_MSGLEN);   /* embed: whole 32-bit words only */
  CHECK(gswm_embed_mt19937(&job, NULL, 42u, buf, GSWM_F32, NULL) == GSWM_E_MSGLEN);
  job.msg_bits = 640;                               /* does not divide 16384 */
  CHECK(gswm_extract(&job, buf, GSWM_F32, buf, NULL, NULL, NULL, NULL, NULL) == GSWM_E_MSGLEN);
  job.msg_bits = 256;
  CHECK(gswm_extract(&job, buf, 9, buf, NULL, NULL, NULL, NULL, NULL) == GSWM_E_DTYPE);
  CHECK(gswm_embed(&job, 0, 0, -1, (float*)buf, NULL) == GSWM_E_RANGE);    /* negative global latent index */
  job.n_latents = 0;                                /* an empty batch is a no-op, not an error */
  CHECK(gswm_embed(&job, 0, 0, 0, (float*)buf, NULL) == GSWM_OK);
  job.msg_bits = 8;                                 /* extract takes any divisor of the latent size (extract.py:195) */
  CHECK(gswm_extract(&job, buf, GSWM_F16, buf, NULL, NULL, NULL, NULL, NULL) == GSWM_OK);
  /* the multi-GPU entry points validate before touching a device */
  CHECK(gswm_comm_create(NULL, 0, 0, 1, NULL) == GSWM_E_NULL);
  CHECK(gswm_comm_allreduce_counters(NULL, NULL, 4, NULL) == GSWM_E_NULL);
  CHECK(gswm_allreduce_counters(NULL, NULL, 4, NULL) == GSWM_E_NULL);
  CHECK(gswm_extract_allreduce(&job, buf, GSWM_F32, buf, NULL, NULL, NULL, NULL, NULL, NULL, NULL) == GSWM_E_NULL);
  CHECK(gswm_chacha20_keystream(buf, buf, 1, 100, buf, NULL) == GSWM_E_SHAPE);
  CHECK(gswm_pipe_embed(NULL, &hjob, 0, 0, 0, out) != GSWM_OK);
  printf("abi_check ok: ABI v%d, Philox4x32-%d\n", gswm_abi_version(), gswm_philox_rounds());
  return 0;
}
