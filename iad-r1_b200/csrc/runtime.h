// Library-internal glue: error string, launch counter, launcher prototypes.
#pragma once
#include <cuda_runtime.h>
#include <cstring>
#include "../../include/iadr1_b200.h"

namespace iadr1 {
int set_error(const char* fmt, ...);
void count_launch(int n = 1);
int launch_gemm(const iadr1_gemm_t& d, cudaStream_t stream);
int pick_block_n_public(int N, int b_mn);
void gemm_profile_enable(int on);
long long gemm_profile_collect(double* total_ms, double* total_flops, double* max_launch_ms, const char* csv_path);
}  // namespace iadr1

namespace iadr1 {
// Programmatic dependent launch (PDL): when enabled (rollout decode chain only), kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization so that the next kernel's CTAs are scheduled - and run their
// prologue and any prefetch of data NOT produced by the immediate predecessor - while the previous kernel drains.
// Every kernel launched this way executes `pdl_wait()` before touching data its predecessor produces.
bool pdl_enabled();
void pdl_set(bool on);

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (pdl_enabled()) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
}  // namespace iadr1

#if defined(__CUDACC__)
namespace iadr1 {
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
}  // namespace iadr1
#endif

#define IADR1_CHECK_LAUNCH(name)                                                          \
  do {                                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess) return ::iadr1::set_error(name ": %s", cudaGetErrorString(e__)); \
    ::iadr1::count_launch();                                                              \
  } while (0)
