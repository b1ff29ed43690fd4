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
bool decode_chain_enabled();
int launch_decode_chain(int R, int H, int I, int QH, int D, float eps, float* h, void* xn, void* act, float* qkv, const void* attn,
                        const void* Wo, const void* ln_mid, const void* Wgu, const void* Wd, const void* ln_next, const void* Wqkv,
                        const void* qkv_bias, int with_mlp, int with_norm, int with_qkv, unsigned* counters, cudaStream_t stream);
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
void trace_install_decode(unsigned long long* p);
void trace_install_gemm(unsigned long long* p);
void trace_install_rowops(unsigned long long* p);
void trace_install_chain(unsigned long long* p);

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
// Decode-chain timeline probe (tools/decode_probe.py): when a trace buffer is installed, thread 0 of CTA 0 of every
// kernel records %globaltimer on reaching the dependency wait (= how early the CTA was resident) and on leaving it
// (= the instant its predecessor finished); successive "leave" stamps are the per-kernel critical-path times INSIDE
// the CUDA graph. buf[0] = record count, records {reach, leave} from buf[2]. One pointer copy per translation unit.
static __device__ unsigned long long* g_trace_buf = nullptr;
static inline void trace_install_tu(unsigned long long* p) { cudaMemcpyToSymbol(g_trace_buf, &p, sizeof(p)); }
__device__ __forceinline__ void pdl_wait() {
  unsigned long long* tb = nullptr;
  unsigned long long t0 = 0;
  if (threadIdx.x == 0 && (blockIdx.x | blockIdx.y | blockIdx.z) == 0) {
    tb = g_trace_buf;
    if (tb) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (tb) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    const unsigned long long i = atomicAdd(tb, 1ull);
    if (i < 8190ull) {
      tb[2 + 2 * i] = t0;
      tb[3 + 2 * i] = t1;
    }
  }
}
// Extra timeline marks inside a kernel (persistent decode chain): record {tag (< 4096), %globaltimer}; call from ONE thread.
__device__ __forceinline__ void trace_stamp(unsigned long long tag) {
  unsigned long long* tb = g_trace_buf;
  if (!tb) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  const unsigned long long i = atomicAdd(tb, 1ull);
  if (i < 8190ull) {
    tb[2 + 2 * i] = tag;
    tb[3 + 2 * i] = t;
  }
}
// Same record format with an arbitrary value instead of the timer (in-kernel cycle accounting).
__device__ __forceinline__ void trace_value(unsigned long long tag, unsigned long long v) {
  unsigned long long* tb = g_trace_buf;
  if (!tb) return;
  const unsigned long long i = atomicAdd(tb, 1ull);
  if (i < 8190ull) {
    tb[2 + 2 * i] = tag;
    tb[3 + 2 * i] = v;
  }
}
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Sigmoid / SiLU and its derivative with the SFU reciprocal (rcp.approx: <= 1 ulp, ~4 fp32 ulp for the whole expression -
// far below the bf16 rounding that follows every use). ONE definition, explicit roundings (no FMA contraction choices left to the
// compiler), used by the row kernels, the GEMM epilogues (epi 3 / 4 / 5), the decode chain and the decode fallback, so the
// fused and unfused forms of the same op stay bit-identical. An IEEE division here costs ~15 issue slots per element, which is
// what made the four epilogue warps of the fused SwiGLU GEMMs slower than the tile's MMA.
__device__ __forceinline__ float rcp_approx_f(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sigmoid_f(float g) { return rcp_approx_f(__fadd_rn(1.f, __expf(-g))); }
__device__ __forceinline__ float silu_f(float g) { return __fmul_rn(g, sigmoid_f(g)); }
// d silu(g) / dg = s (1 + g (1 - s)), s = sigmoid(g)
__device__ __forceinline__ float silu_grad_s(float g, float s) { return __fmul_rn(s, __fmaf_rn(g, __fsub_rn(1.f, s), 1.f)); }
__device__ __forceinline__ float silu_grad_f(float g) { return silu_grad_s(g, sigmoid_f(g)); }
// SwiGLU backward for one element: dgate = d u silu'(g), dup = d silu(g); _s takes s = sigmoid_f(g) computed by the caller
__device__ __forceinline__ void swiglu_bwd_s(float d, float g, float u, float s, float& dg, float& du) {
  dg = __fmul_rn(__fmul_rn(d, u), silu_grad_s(g, s));
  du = __fmul_rn(d, __fmul_rn(g, s));
}
__device__ __forceinline__ void swiglu_bwd_f(float d, float g, float u, float& dg, float& du) {
  swiglu_bwd_s(d, g, u, sigmoid_f(g), dg, du);
}
}  // namespace iadr1
#endif

#define IADR1_CHECK_LAUNCH(name)                                                          \
  do {                                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess) return ::iadr1::set_error(name ": %s", cudaGetErrorString(e__)); \
    ::iadr1::count_launch();                                                              \
  } while (0)
