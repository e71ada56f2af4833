// Device side of gswm_comm: the all-reduce of the GSWM_CTR_* counter vector over NVLink peer memory.
//
// Every rank owns a MAILBOX in its own HBM: box[parity][source rank], one 128-byte slot each.  A rank publishes by
// storing its values into slot [epoch & 1][my rank] of EVERY peer's mailbox (P2P stores through the NVSwitch) followed
// by a system-scope release store of the epoch; it collects by spinning, with acquire loads on its OWN memory, until
// every source's slot carries the epoch, and summing.  One warp does all of it, lane r talking to peer r, so the
// n_ranks exchanges are in flight together: the cost is one NVLink store round, a few microseconds, with no host in
// the loop.  Two parities because a fast rank may already be publishing call k+1 while a slow one still reads call k
// (it cannot get further ahead: call k+1 only completes once the slow rank has published k+1 too).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/gswm.h"

namespace gswm {

struct alignas(128) CommSlot {
  long long v[GSWM_COMM_MAX_VALUES];
  unsigned long long epoch;
  unsigned long long pad[16 - GSWM_COMM_MAX_VALUES - 1];
};
static_assert(sizeof(CommSlot) == 128, "one slot per 128-byte line");

// what a kernel needs of a communicator (passed by value in kernel-parameter space)
struct CommDev {
  CommSlot* box[GSWM_COMM_MAX_RANKS];   // box[r]: rank r's mailbox as mapped into THIS rank's address space, [2][GSWM_COMM_MAX_RANKS] slots
  unsigned* ticket;                     // this rank's CTA retirement counter (fused extract); self-resetting
  int* status;                          // mapped host word: set to GSWM_E_COMM when a peer does not show up
  unsigned long long epoch;             // number of this collective call, from 1
  int rank;
  int n_ranks;
};

__device__ __forceinline__ unsigned long long comm_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void comm_st_relaxed(long long* p, long long v) {
  asm volatile("st.relaxed.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void comm_st_release(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long comm_ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ long long comm_ld_relaxed(const long long* p) {
  long long v;
  asm volatile("ld.relaxed.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

constexpr unsigned long long kCommTimeoutNs = 10ull * 1000 * 1000 * 1000;

// One whole warp: sum `mine[0..n)` (every lane passes the same values) over all ranks; every lane returns the sums
// in `out`.  n <= GSWM_COMM_MAX_VALUES.
__device__ __forceinline__ void comm_allreduce_warp(const CommDev& c, const long long (&mine)[GSWM_COMM_MAX_VALUES], int n,
                                                    long long (&out)[GSWM_COMM_MAX_VALUES]) {
  const int lane = (int)(threadIdx.x & 31u);
  const unsigned parity = (unsigned)(c.epoch & 1ull);
  long long acc[GSWM_COMM_MAX_VALUES];
#pragma unroll
  for (int i = 0; i < GSWM_COMM_MAX_VALUES; ++i) acc[i] = 0;
  if (lane < c.n_ranks) {
    CommSlot* dst = c.box[lane] + parity * GSWM_COMM_MAX_RANKS + c.rank;          // my slot in peer `lane`'s mailbox
#pragma unroll
    for (int i = 0; i < GSWM_COMM_MAX_VALUES; ++i)
      if (i < n) comm_st_relaxed(&dst->v[i], mine[i]);
    comm_st_release(&dst->epoch, c.epoch);
    const CommSlot* src = c.box[c.rank] + parity * GSWM_COMM_MAX_RANKS + lane;    // peer `lane`'s slot in my mailbox
    const unsigned long long t0 = comm_globaltimer();
    bool ok = true;
    while (comm_ld_acquire(&src->epoch) != c.epoch) {
      if (comm_globaltimer() - t0 > kCommTimeoutNs) {
        *reinterpret_cast<volatile int*>(c.status) = GSWM_E_COMM;
        ok = false;
        break;
      }
      __nanosleep(64);
    }
    if (ok) {
#pragma unroll
      for (int i = 0; i < GSWM_COMM_MAX_VALUES; ++i)
        if (i < n) acc[i] = comm_ld_relaxed(&src->v[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < GSWM_COMM_MAX_VALUES; ++i) {
    long long a = acc[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xFFFFFFFFu, a, d);
    out[i] = a;
  }
}

}  // namespace gswm
