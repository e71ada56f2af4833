// Internal interfaces between the translation units of libgswm (not part of the C ABI).
#pragma once
#include <cstdint>

#include "../../include/gswm.h"
#include "gswm_comm.cuh"

namespace gswm {

void count_launch();                                          // gswm_launch_count() bookkeeping
int check_job(const gswm_job* job, bool for_extract);        // argument checks shared by every device entry point

// gswm_extract (comm == nullptr) and gswm_extract_allreduce
int extract_impl(const gswm_job* job, const void* d_z, int32_t z_dtype, uint8_t* d_msg_out, uint16_t* d_counts,
                 int32_t* d_matched, uint8_t* d_flags, int64_t* d_counters, const CommDev* comm, int64_t* d_reduced,
                 void* stream);
int comm_allreduce_launch(const CommDev& c, int64_t* d_values, int n, void* stream);

}  // namespace gswm
