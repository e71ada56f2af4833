/*
 * gswm.h -- C ABI of libgswm.so: the Gaussian-Shading watermark codec hot path on B200 (sm_100a).
 *
 * The reference (lthero-big/A-watermark-for-Diffusion-Models) has no FFI: its "API" is a handful of
 * Python functions built from three third-party calls.  Each entry point below names the reference
 * statement(s) it replaces (file:line relative to the reference repository).
 *
 * Conventions
 *   - plain C: pointers, sizes, no torch / C++ types.  `stream` is a cudaStream_t passed as void*
 *     (NULL = the legacy default stream).
 *   - pointers named d_* are DEVICE pointers, h_* are HOST pointers.
 *   - the device entry points never allocate, free or synchronise: work is enqueued on `stream` and
 *     the call returns; outputs are fully overwritten.  Only gswm_pipe_* (host-buffer convenience
 *     layer) owns device memory, allocated once in gswm_pipe_create.
 *   - return value: 0 = success; negative = GSWM_E_* argument error; positive = cudaError_t.
 *   - bit order: latent element e of a latent maps to keystream byte e>>3, bit 7-(e&7)
 *     (gs_insert.py:49 `format(byte,'08b')`, extract.py:86).
 *   - a latent is the C-order flattening of (4, H/8, W/8) (gs_insert.py:56,65); n_elems = 4*(H/8)*(W/8).
 */
#ifndef GSWM_H_
#define GSWM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSWM_ABI_VERSION 2

enum {
  GSWM_OK = 0,
  GSWM_E_NULL = -1,        /* a required pointer is NULL */
  GSWM_E_SHAPE = -2,       /* n_elems not a positive multiple of 4, or n_latents < 0 */
  GSWM_E_MSGLEN = -3,      /* embed: msg_bits not a positive multiple of 32 or > n_elems; extract: msg_bits not a positive divisor of n_elems */
  GSWM_E_DTYPE = -4,       /* unknown element type code */
  GSWM_E_RANGE = -5,       /* a size exceeds what the kernels index */
  GSWM_E_COMM = -6,        /* communicator: bad rank / size, peer memory cannot be mapped, NCCL not loadable, or a peer timed out */
  GSWM_E_ALIGN = -7        /* latent pointer (or row size) not 16-byte aligned, or key / nonce / message pointer not 4-byte aligned */
};

/* element type codes for latents */
enum { GSWM_F32 = 0, GSWM_F16 = 1, GSWM_BF16 = 2, GSWM_F64 = 3 };

/*
 * One batch job: n_latents latents of n_elems elements each, msg_bits-bit messages.
 *
 *   flags & GSWM_JOB_PER_LATENT == 0 : d_keys[32], d_nonces[16], d_msgs[msg_bytes] are shared by the whole batch
 *                     (the reference's usual case: one --key_hex/--nonce_hex/--message per run).
 *   flags & GSWM_JOB_PER_LATENT      : d_keys[n][32], d_nonces[n][16], d_msgs[n][msg_bytes], one row per latent.
 *   msg_bytes = (msg_bits + 7) / 8.
 *
 * key/nonce are the bytes gs_insert.py:27-42 resolves from key_hex/nonce_hex; the message bytes are
 * the padded/truncated `k` of gs_insert.py:9-20 (nodes.py:68-76, v1.5.2:29-47 for other framings).
 * The message is tiled n_elems/msg_bits times; a remainder (msg_bits not dividing n_elems, embed only)
 * carries plaintext zero, as nodes.py:79-87 does.
 *
 * msg_bits: embed needs a positive multiple of 32 (every reference embed site uses whole bytes: 256 bits in
 * gs_insert.py:13-23, a multiple of 32 from the ComfyUI widget, nodes.py:221).  Extract takes ANY positive divisor of
 * n_elems up to 8192 -- extract.py:195 `--message_length` is an arbitrary integer -- and then d_msg_out / d_msgs rows
 * are msg_bytes long with the unused low bits of the last byte zero.
 *
 * Ordering: by default d_keys / d_nonces / d_msgs must hold their final contents when the call is ENQUEUED: every gswm
 * kernel uses programmatic dependent launch and reads key material ahead of the grid dependency wait, so that the
 * ChaCha20 prologue overlaps the previous kernel's tail.  Key material written by a memcpy or by an ordinary kernel
 * earlier in the same stream is ordered as usual; key material written by a kernel that ITSELF triggers programmatic
 * launch completion early (another gswm kernel cannot: none writes key material) needs GSWM_JOB_KEYS_IN_FLIGHT, which
 * moves every key-material read behind the wait (costs the ~2.4 us prologue per launch).  Latents and all outputs
 * are only touched after the wait.
 */
enum {
  GSWM_JOB_PER_LATENT = 1,      /* one key / nonce / message row per latent */
  GSWM_JOB_KEYS_IN_FLIGHT = 2   /* key material may still be being written by the preceding kernel of the stream */
};

typedef struct gswm_job {
  int64_t n_latents;
  int64_t n_elems;
  int32_t msg_bits;
  int32_t flags;           /* GSWM_JOB_* (bit 0 was `per_latent` in ABI v1: same layout, same meaning) */
  const uint8_t* d_keys;
  const uint8_t* d_nonces;
  const uint8_t* d_msgs;   /* extract: may be NULL (no reference message to score against) */
} gswm_job;

/* counters written by gswm_extract (int64 each) -- the buffer an all-reduce over ranks sums */
enum {
  GSWM_CTR_MATCHED_BITS = 0, /* sum over latents of decoded bits equal to the reference message */
  GSWM_CTR_TOTAL_BITS = 1,   /* n_latents * msg_bits */
  GSWM_CTR_EXACT_MSGS = 2,   /* latents whose decoded message equals the reference exactly */
  GSWM_CTR_TOTAL_MSGS = 3,   /* n_latents */
  GSWM_CTR_NAN_LATENTS = 4,  /* latents holding a NaN: extract.py:83 raises ValueError (int(nan)) for them */
  GSWM_CTR_RANGE_LATENTS = 5,/* latents with no NaN but an element >= 8.292361075813597 (+inf included):
                                int(norm.cdf(z)*2) == 2 there and extract.py:86 raises on the digit '2' */
  GSWM_N_COUNTERS = 6
};

/* per-latent flags written by gswm_extract (d_flags): the inputs extract.recover_exactracted_message rejects */
enum { GSWM_FLAG_NAN = 1, GSWM_FLAG_RANGE = 2 };

int gswm_abi_version(void);
const char* gswm_strerror(int code);

/*
 * ChaCha20 keystream, original 64-bit-counter layout: state words 12,13 = LE64(nonce[0:8]) + block,
 * words 14,15 = nonce[8:16] -- what `Cipher(algorithms.ChaCha20(key, nonce)).encryptor().update(zeros)`
 * yields (gs_insert.py:45-47, nodes.py:101-103, extract.py:77-78,87).
 * d_out[s][0:n_bytes_each] for s < n_streams; n_bytes_each must be a multiple of 64.
 */
int gswm_chacha20_keystream(const uint8_t* d_keys, const uint8_t* d_nonces, int64_t n_streams,
                            int64_t n_bytes_each, uint8_t* d_out, void* stream);

/*
 * Embed with the in-kernel counter-based uniform source -- replaces the whole of
 * gs_insert.gs_watermark_init_noise's arithmetic (gs_insert.py:23-66; nodes.py:76-123) for a batch:
 * tile message, XOR ChaCha20 keystream, one uniform per element, z = Phi^-1((u + y) / 2), fp32 store.
 *
 * Uniform source ("gswm uniforms v4", csrc/gswm_math.cuh; restated in oracle/gs_oracle.py:gswm_uniforms): every
 * element gets a 23-bit integer m from Philox4x32-7 keyed by `seed`, with the counter built from the GLOBAL
 * latent index first_latent + b (so a batch sharded over ranks produces the same latents as one big batch),
 * the tile, the position and `offset` (< 2^62).  v = (m + 1/2) 2^-23; u = v for bucket bit 1, u = 1 - v for
 * bucket bit 0 (also a grid point), which makes z = +-sqrt(2) erfinv(v).  The outermost cell m = 2^23 - 1
 * (probability 2^-23 per element) is subdivided by 28 more Philox bits, v = 1 - (m2 + 1/2) 2^-51, so |z| reaches
 * 8.2095 = norm.ppf(1 - 2^-53), the largest value the reference's 53-bit uniforms produce for bucket 1 -- the
 * 23-bit grid alone would stop at 5.42.
 * d_out: [n_latents][n_elems] fp32, 16-byte aligned.  first_latent >= 0 and
 * (first_latent + n_latents) * ceil(n_elems / 16384) <= 2^52 (the counter holds 54 bits of latent / tile / position).
 */
int gswm_embed(const gswm_job* job, uint64_t seed, uint64_t offset, int64_t first_latent,
               float* d_out, void* stream);

/*
 * Embed with injected uniforms: d_u[n_latents][n_elems] float64 in [0,1) takes the place of
 * np.random.uniform(0,1) / RandomState(seed).uniform(0,1) (gs_insert.py:62; nodes.py:114-117) and
 * z = Phi^-1((u + y)/2) is evaluated in float64 (gs_insert.py:64).  u is not range-checked: u = 0 with bucket bit 0 gives
 * -inf and a u outside [0, 1] that pushes (u + y)/2 outside [0, 1] gives NaN, exactly as scipy's norm.ppf does.
 * u_per_latent == 0 reuses one row of uniforms for every latent.  out_dtype: GSWM_F32 (what every caller casts to, README.md:112)
 * or GSWM_F64 (what gs_insert.py:75 returns).
 */
int gswm_embed_injected(const gswm_job* job, const double* d_u, int32_t u_per_latent, void* d_out,
                        int32_t out_dtype, void* stream);

/*
 * Embed with the reference's OWN uniform stream generated on the device: latent b draws
 * u = np.random.RandomState(seed_b).uniform(0, 1) element by element (nodes.py:52-53,114-117; v1.5.2:27,72-75) --
 * MT19937 seeded by init_genrand(seed_b), 53-bit doubles (a >> 5, b >> 6) -- bit for bit, and
 * z = Phi^-1((u + y)/2) in float64 as gswm_embed_injected does.  seed_b = d_seeds[b], or seed0 + b (mod 2^32) when
 * d_seeds is NULL.  No uniforms cross PCIe: a seeded call site needs no 8-byte-per-element upload.
 * out_dtype: GSWM_F32 or GSWM_F64.
 */
int gswm_embed_mt19937(const gswm_job* job, const uint32_t* d_seeds, uint32_t seed0, void* d_out,
                       int32_t out_dtype, void* stream);

/* The generator alone: d_out[s][0:n_each] = RandomState(seed_s).uniform(size=n_each) (float64), seed_s as above. */
int gswm_mt19937_uniform(const uint32_t* d_seeds, uint32_t seed0, int64_t n_streams, int64_t n_each,
                         double* d_out, void* stream);

/*
 * Extract -- replaces extract.recover_exactracted_message (extract.py:72-101) and the counting half
 * of calculate_bit_accuracy (extract.py:103-110) for a batch of inverted latents d_z
 * [n_latents][n_elems] of type z_dtype (GSWM_F32 / GSWM_F16 / GSWM_BF16 / GSWM_F64):
 *   bit = int(norm.cdf(z)*2)  ==  (z >= -6.957291061679417e-17)       extract.py:82-84
 *   (the threshold is applied in the input's own type: the smallest fp32 / bf16 value >= it, +0 for fp16, itself for fp64)
 *   decrypt with the keystream, count ones per message position over the n_elems/msg_bits copies,
 *   strict majority (tie -> 0)                                         extract.py:86-99
 * Outputs (each may be NULL except d_msg_out):
 *   d_msg_out [n_latents][msg_bytes]    decoded message, MSB-first bits (bytes of the '0'/'1' string)
 *   d_counts  [n_latents][msg_bits] u16 count_1 per position (needs n_elems/msg_bits <= 65535)
 *   d_matched [n_latents] i32           bits equal to job->d_msgs (the reference message), per latent
 *   d_flags   [n_latents] u8            GSWM_FLAG_*: inputs for which the reference RAISES instead of decoding --
 *                                       a NaN (int(nan), extract.py:83) or an element >= 8.292361075813597 / +inf
 *                                       (int(norm.cdf(z)*2) == 2, extract.py:86).  Found by the kernel in the same
 *                                       pass (a NaN-propagating running maximum); such a latent is still decoded
 *                                       with bit = (z >= threshold) (NaN -> 0), and the caller decides: the
 *                                       extract.py drop-in raises ValueError, the batch front-end logs and skips.
 *   d_counters[GSWM_N_COUNTERS] i64     ACCUMULATED (caller zeroes); see GSWM_CTR_*
 * d_matched / matched counters need job->d_msgs.  Rows of d_z must be 16-byte aligned (n_elems * element size).
 */
int gswm_extract(const gswm_job* job, const void* d_z, int32_t z_dtype, uint8_t* d_msg_out,
                 uint16_t* d_counts, int32_t* d_matched, uint8_t* d_flags, int64_t* d_counters, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU: latents are independent, so ranks own disjoint batch ranges and exchange nothing but the
 * GSWM_CTR_* counter vector at the end.  A gswm_comm is a set of per-rank
 * MAILBOXES in device memory that every peer maps over NVLink (CUDA IPC between processes, peer access
 * inside one process); the all-reduce is one kernel per rank that stores this rank's counters into
 * every peer's mailbox (P2P stores + a system-scope release flag), waits for the peers' stores in its
 * own, and sums -- no host round trip, no library call.  Collective: every rank calls it the same
 * number of times, in the same order, on one stream per rank.
 *
 *   gswm_comm_create   allocates this rank's mailbox on `device` and returns its IPC handle;
 *   gswm_comm_connect  maps the peers' mailboxes from all ranks' handles ([n_ranks][64], gathered by the host
 *                      with whatever transport it has: torch.distributed, MPI, a file);
 *   gswm_comm_connect_local  the same for ranks living in ONE process (ncclCommInitAll style): comms[n_ranks].
 *
 * gswm_extract_allreduce is gswm_extract with that exchange FUSED into the kernel: the last CTA to retire
 * publishes the accumulated d_counters and writes the sum over ranks to d_reduced[GSWM_N_COUNTERS]; d_counters
 * itself keeps this rank's own totals (a rank with n_latents == 0 decodes nothing and still takes part in the exchange).
 * gswm_comm_allreduce_counters is the stand-alone form (in place,
 * n <= GSWM_COMM_MAX_VALUES).  A peer that does not show up within ~10 s makes the kernel give up:
 * gswm_comm_status() then returns GSWM_E_COMM and the reduced values are unspecified.
 *
 * gswm_allreduce_counters is the same sum through NCCL for a host that already owns a communicator
 * (ncclComm_t passed as void*; libnccl.so.2 is resolved at run time with dlopen, the copy already loaded
 * in the process if there is one): ncclAllReduce(int64, sum), in place, on `stream`.
 * ---------------------------------------------------------------------------------------------- */
typedef struct gswm_comm gswm_comm;
#define GSWM_COMM_HANDLE_BYTES 64
#define GSWM_COMM_MAX_VALUES 8
#define GSWM_COMM_MAX_RANKS 32

int gswm_comm_create(gswm_comm** out, int device, int rank, int n_ranks, uint8_t* handle_out);
int gswm_comm_connect(gswm_comm* comm, const uint8_t* all_handles);
int gswm_comm_connect_local(gswm_comm* const* comms, int n_ranks);
void gswm_comm_destroy(gswm_comm* comm);
int gswm_comm_status(gswm_comm* comm);   /* synchronises nothing: reads the status word the kernels write (mapped host memory) */

int gswm_comm_allreduce_counters(gswm_comm* comm, int64_t* d_counters, int32_t n, void* stream);
int gswm_extract_allreduce(const gswm_job* job, const void* d_z, int32_t z_dtype, uint8_t* d_msg_out,
                           uint16_t* d_counts, int32_t* d_matched, uint8_t* d_flags, int64_t* d_counters,
                           gswm_comm* comm, int64_t* d_reduced, void* stream);
int gswm_allreduce_counters(void* nccl_comm, int64_t* d_counters, int32_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Host-buffer layer: what a caller holding numpy / CPU-torch buffers uses (the reference builds its
 * latents on the CPU and `.to(device)`s them, README.md:112; extract.py:70 hands back a CPU tensor).
 * A pipe owns streams, pinned staging and device buffers sized for `max_latents_per_chunk` latents of
 * `max_elems` elements; batches of any size are streamed through it in chunks with the H2D / kernel /
 * D2H stages of consecutive chunks overlapped.  Host job fields (h_keys/h_nonces/h_msgs) follow the
 * gswm_job rules but point to host memory.
 * ---------------------------------------------------------------------------------------------- */
typedef struct gswm_pipe gswm_pipe;

typedef struct gswm_host_job {
  int64_t n_latents;
  int64_t n_elems;
  int32_t msg_bits;
  int32_t flags;           /* GSWM_JOB_PER_LATENT */
  const uint8_t* h_keys;
  const uint8_t* h_nonces;
  const uint8_t* h_msgs;
} gswm_host_job;

int gswm_pipe_create(gswm_pipe** out, int device, int64_t max_elems, int64_t max_latents_per_chunk);
void gswm_pipe_destroy(gswm_pipe* pipe);

/* h_out [n_latents][n_elems] fp32 host memory (pinned memory gives full PCIe rate). Synchronous. */
int gswm_pipe_embed(gswm_pipe* pipe, const gswm_host_job* job, uint64_t seed, uint64_t offset,
                    int64_t first_latent, float* h_out);

/* h_u [n_latents or 1][n_elems] float64, h_out fp32 or fp64 per out_dtype. Synchronous. */
int gswm_pipe_embed_injected(gswm_pipe* pipe, const gswm_host_job* job, const double* h_u,
                             int32_t u_per_latent, void* h_out, int32_t out_dtype);

/* h_z [n_latents][n_elems] of z_dtype; outputs as gswm_extract but in host memory; h_counters[GSWM_N_COUNTERS]
 * is OVERWRITTEN with this batch's totals. Synchronous. */
int gswm_pipe_extract(gswm_pipe* pipe, const gswm_host_job* job, const void* h_z, int32_t z_dtype,
                      uint8_t* h_msg_out, uint16_t* h_counts, int32_t* h_matched, uint8_t* h_flags,
                      int64_t* h_counters);

/* Test hooks (used by tests/ only): the fp32 bucket quantile of gswm_embed on caller-supplied raw
 * 32-bit words (u = ((w >> 9) + 0.5) * 2^-23, all elements in bucket `bucket_bit`; use_vec4 selects
 * the 4-wide code path the embed kernel runs), and the fp64 Phi^-1 of gswm_embed_injected. */
int gswm_debug_bucket_quantile(const uint32_t* d_words, int64_t n, int32_t bucket_bit, int32_t use_vec4,
                               float* d_out, void* stream);
int gswm_debug_norm_ppf(const double* d_p, int64_t n, double* d_out, void* stream);
/* |z| of the refined outermost grid cell (uniforms v3) for caller-supplied refinement words: m2 = w >> 4,
 * |z| = -Phi^-1((m2 + 1/2) 2^-52). */
int gswm_debug_top_cell(const uint32_t* d_words, int64_t n, float* d_out, void* stream);
/* Philox4x32-`rounds` (7 or 10) on caller-supplied inputs: d_in[n][6] = counter words 0..3, key words 0..1;
 * d_out[n][4].  rounds = 10 is the variant with published known-answer vectors (Random123 kat_vectors). */
int gswm_debug_philox4x32(const uint32_t* d_in, int64_t n, int32_t rounds, uint32_t* d_out, void* stream);

/* Issue-rate microbenchmark for one instruction class of the embed kernel (bench.py's issue-utilisation denominators,
 * measured in the same run): warp instructions per clock per SM sub-partition, and the SM clock (GHz) during the run.
 * Allocates and frees its own scratch buffers and synchronises the device: a measurement tool, not a codec entry point. */
enum {
  GSWM_ISSUE_FFMA2 = 0,      /* packed fp32x2 FMA (polynomial)           */
  GSWM_ISSUE_IMAD_WIDE = 1,  /* 32 x 32 -> 64 multiply (Philox)          */
  GSWM_ISSUE_LOP3 = 2,       /* three-input logic                        */
  GSWM_ISSUE_MUFU = 3,       /* MUFU.LG2                                 */
  GSWM_ISSUE_FFMA_IMM = 4,   /* scalar FMA, immediate operands: the one-instruction-per-clock issue peak */
  GSWM_ISSUE_PHILOX_MIX = 5, /* IMAD.WIDE + LOP3 feeding each other (a Philox round's shape) */
  GSWM_ISSUE_FFMA2_LOP3 = 6,      /* independent pairs: do the pipes overlap? (2 instructions per step) */
  GSWM_ISSUE_FFMA_IMAD_WIDE = 7,
  GSWM_ISSUE_FFMA2_IMAD_WIDE = 8,
  GSWM_ISSUE_LOP3_IMAD_WIDE = 9
};
int gswm_debug_issue_rate(int32_t kind, double* warp_inst_per_clk_per_smsp, double* sm_ghz);

/* Number of kernels the library has launched in this process (all entry points); for bench.py. */
int64_t gswm_launch_count(void);

/* Rounds of the Philox4x32 generator behind gswm_embed's uniforms (7 unless built with -DGSWM_PHILOX_ROUNDS=..). */
int gswm_philox_rounds(void);

#ifdef __cplusplus
}
#endif
#endif /* GSWM_H_ */
