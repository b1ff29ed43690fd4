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
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  adamw_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p32, (__nv_bfloat16*)p16, g, m, v, n, lr, beta1, beta2,
                                                              eps, weight_decay, bc1, bc2, grad_scale, sumsq, max_norm,
                                                              zero_grad);
  IADR1_CHECK_LAUNCH("adamw_step");
  return 0;
}

}  // extern "C"
