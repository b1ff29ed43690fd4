// Library-internal glue: error string, launch counter, launcher prototypes.
#pragma once
#include <cuda_runtime.h>
#include "../../include/iadr1_b200.h"

namespace iadr1 {
int set_error(const char* fmt, ...);
void count_launch(int n = 1);
int launch_gemm(const iadr1_gemm_t& d, cudaStream_t stream);
int pick_block_n_public(int N, int b_mn);
void gemm_profile_enable(int on);
long long gemm_profile_collect(double* total_ms, double* total_flops, double* max_launch_ms);
}  // namespace iadr1

#define IADR1_CHECK_LAUNCH(name)                                                          \
  do {                                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess) return ::iadr1::set_error(name ": %s", cudaGetErrorString(e__)); \
    ::iadr1::count_launch();                                                              \
  } while (0)
