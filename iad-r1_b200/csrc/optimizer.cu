// Fused multi-tensor AdamW + global gradient-norm clip over FLAT parameter / gradient / moment buffers.
// Replaces torch.optim.AdamW wrapped in DeepSpeed's ZeRO-3 optimizer (HF Trainer default `optim=adamw_torch`;
// ref: scripts/train/zero3.json has no optimizer block) with the same update rule:
//   p *= 1 - lr*wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// on fp32 master weights, emitting the bf16 working copy in the same pass and zeroing the fp32 gradient, i.e. one
// HBM pass of 4(g) + 4(p) + 4(m) + 4(v) reads and 4 + 4 + 4 + 2 + 4 writes per parameter.
// The clip coefficient min(1, max_norm / (||g|| + 1e-6)) (torch.nn.utils.clip_grad_norm_, Trainer max_grad_norm=1.0)
// is computed IN the kernel from a device-side sum of squares, so the optimizer step needs no host synchronisation.
#include "runtime.h"
#include <cuda_bf16.h>
#include <cmath>
#include <stdint.h>

namespace iadr1 {

__global__ void sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  __shared__ float red[32];
  float acc = 0.f;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    acc += g[i] * g[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) atomicAdd(out, acc);
  }
}

__global__ void adamw_kernel(float* __restrict__ p32, __nv_bfloat16* __restrict__ p16, float* __restrict__ g,
                             float* __restrict__ m, float* __restrict__ v, long long n, float lr, float b1, float b2,
                             float eps, float wd, float bc1, float bc2, float grad_scale,
                             const float* __restrict__ sumsq, float max_norm, int zero_grad) {
  float clip = grad_scale;
  if (max_norm > 0.f && sumsq != nullptr) {
    const float norm = sqrtf(*sumsq) * grad_scale;
    clip = grad_scale * fminf(1.f, max_norm / (norm + 1e-6f));
  }
  const float step = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const float decay = 1.f - lr * wd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * clip;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    float pi = p32[i] * decay;
    pi -= step * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    m[i] = mi;
    v[i] = vi;
    p32[i] = pi;
    p16[i] = __float2bfloat16(pi);
    if (zero_grad) g[i] = 0.f;
  }
}

// Same update with the two moments stored in bf16 (8 bytes per parameter less: what lets Qwen2.5-VL-7B - 8.29 B parameters,
// policy + frozen reference + fp32 gradient + fp32 master - fit one 180 GB GPU under plain data parallel).
// Round-to-nearest would freeze exp_avg_sq (beta2 = 0.999 moves it by 0.1 % per step, a bf16 ulp is 0.4 %), so both
// moments are written with STOCHASTIC rounding: the fp32 value plus 16 random low bits, truncated - unbiased, so the
// expectation of the stored moment follows the fp32 recurrence. The random bits come from a counter hash of
// (element index, step): reproducible, no state.
__device__ __forceinline__ uint32_t hash32(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return (uint32_t)x;
}
__device__ __forceinline__ __nv_bfloat16 bf16_stochastic(float x, uint32_t rnd16) {
  uint32_t u = __float_as_uint(x);
  if ((u & 0x7f800000u) != 0x7f800000u) u += (rnd16 & 0xffffu);     // finite: add random low bits, then truncate
  return __ushort_as_bfloat16((unsigned short)(u >> 16));
}

__global__ void adamw_bf16m_kernel(float* __restrict__ p32, __nv_bfloat16* __restrict__ p16, float* __restrict__ g,
                                   __nv_bfloat16* __restrict__ m, __nv_bfloat16* __restrict__ v, long long n, float lr,
                                   float b1, float b2, float eps, float wd, float bc1, float bc2, float grad_scale,
                                   const float* __restrict__ sumsq, float max_norm, int zero_grad, unsigned long long seed) {
  float clip = grad_scale;
  if (max_norm > 0.f && sumsq != nullptr) {
    const float norm = sqrtf(*sumsq) * grad_scale;
    clip = grad_scale * fminf(1.f, max_norm / (norm + 1e-6f));
  }
  const float step = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const float decay = 1.f - lr * wd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * clip;
    const float mi = b1 * __bfloat162float(m[i]) + (1.f - b1) * gi;
    const float vi = b2 * __bfloat162float(v[i]) + (1.f - b2) * gi * gi;
    float pi = p32[i] * decay;
    pi -= step * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    const uint32_t r = hash32(seed + (uint64_t)i);
    m[i] = bf16_stochastic(mi, r);
    v[i] = bf16_stochastic(vi, r >> 16);
    p32[i] = pi;
    p16[i] = __float2bfloat16(pi);
    if (zero_grad) g[i] = 0.f;
  }
}

}  // namespace iadr1

using namespace iadr1;

extern "C" {

int iadr1_sumsq_f32(const float* g, long long n, float* out, void* stream) {
  if (n <= 0) return 0;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  sumsq_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(g, n, out);
  IADR1_CHECK_LAUNCH("sumsq_f32");
  return 0;
}

int iadr1_adamw_step(float* p32, void* p16, float* g, float* m, float* v, long long n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int step, float grad_scale, const float* sumsq,
                     float max_norm, int zero_grad, void* stream) {
  if (n <= 0) return 0;
  if (step < 1) return set_error("adamw_step: step must be >= 1");
  // bias corrections in double on the host (fp32 powf loses ~6e-5 of 1 - beta2^t at t = 1)
  const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  const float bc2 = (float)(1.0 - pow((double)beta2, (double)step));
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  adamw_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p32, (__nv_bfloat16*)p16, g, m, v, n, lr, beta1, beta2,
                                                              eps, weight_decay, bc1, bc2, grad_scale, sumsq, max_norm,
                                                              zero_grad);
  IADR1_CHECK_LAUNCH("adamw_step");
  return 0;
}

int iadr1_adamw_step_bf16m(float* p32, void* p16, float* g, void* m16, void* v16, long long n, float lr, float beta1,
                           float beta2, float eps, float weight_decay, int step, float grad_scale, const float* sumsq,
                           float max_norm, int zero_grad, unsigned long long seed, void* stream) {
  if (n <= 0) return 0;
  if (step < 1) return set_error("adamw_step_bf16m: step must be >= 1");
  const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  const float bc2 = (float)(1.0 - pow((double)beta2, (double)step));
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  adamw_bf16m_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p32, (__nv_bfloat16*)p16, g, (__nv_bfloat16*)m16,
                                                                    (__nv_bfloat16*)v16, n, lr, beta1, beta2, eps, weight_decay,
                                                                    bc1, bc2, grad_scale, sumsq, max_norm, zero_grad,
                                                                    seed * 0x9E3779B97F4A7C15ULL + (unsigned long long)step * 0xD1B54A32D192ED03ULL);
  IADR1_CHECK_LAUNCH("adamw_step_bf16m");
  return 0;
}

}  // extern "C"
