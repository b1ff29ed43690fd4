// HBM-bound row kernels of the VLM forward/backward: RMSNorm / LayerNorm, rotary embedding, gated activations,
// masked row softmax (+ its backward), embedding gather / scatter-add, row gather, column sums (bias gradients).
// All of them are one pass over their operands with 16-byte accesses and warp-shuffle reductions; rounding points
// follow the HF modules the reference executes (bf16 activations, fp32 statistics), cited per kernel.
#include "runtime.h"
#include <cuda_bf16.h>
#include <stdint.h>

// Eight consecutive fp32 sums into global memory as two 16-byte vector reductions (REDG.E.ADD.F32x4): the per-column weight /
// bias gradients of the norm backward kernels land on the same few cache lines from every CTA, and the L2 handles a whole
// sector per request instead of one float (8 scalar atomics per thread: the tail was a third of the kernel).
static __device__ __forceinline__ void red_add8_f32(float* p, const float (&v)[8]) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p + 4), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}

namespace iadr1 {

using bf16 = __nv_bfloat16;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Block-wide sum for blockDim.x <= 1024; `red` is a 32-float shared scratch.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  v = (l < nw) ? red[l] : 0.f;
  return warp_sum(v);
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  v = (l < nw) ? red[l] : -INFINITY;
  return warp_max(v);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

struct alignas(16) bf16x8 {
  __nv_bfloat162 h[4];
};
__device__ __forceinline__ void unpack8(const bf16x8& v, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(v.h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ bf16x8 pack8(const float (&f)[8]) {
  bf16x8 v;
#pragma unroll
  for (int i = 0; i < 4; ++i) v.h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

// ------------------------------------------------------------------------------------------------
// RMSNorm.  HF: Qwen2_5_VLRMSNorm.forward (modeling_qwen2_5_vl.py:57-71): fp32 statistics, the normalised value is
// rounded to bf16 BEFORE the bf16 weight multiply.  One CTA per row; cols % 8 == 0.
// ------------------------------------------------------------------------------------------------
template <int MAX_VEC>  // vectors of 8 held per thread
__global__ void rmsnorm_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, bf16* __restrict__ y,
                                   float* __restrict__ rstd_out, int cols, long long x_ld, long long y_ld, float eps) {
  __shared__ float red[32];
  const long long row = blockIdx.x;
  const bf16x8* xr = reinterpret_cast<const bf16x8*>(x + row * x_ld);
  const bf16x8* wr = reinterpret_cast<const bf16x8*>(w);
  bf16x8* yr = reinterpret_cast<bf16x8*>(y + row * y_ld);
  const int nvec = cols >> 3;
  bf16x8 xv[MAX_VEC];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int v = threadIdx.x + i * blockDim.x;
    if (v < nvec) {
      xv[i] = xr[v];
      float f[8];
      unpack8(xv[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
    }
  }
  ss = block_sum(ss, red);
  const float rstd = rsqrtf(ss / (float)cols + eps);
  if (threadIdx.x == 0 && rstd_out) rstd_out[row] = rstd;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int v = threadIdx.x + i * blockDim.x;
    if (v < nvec) {
      float f[8], g[8];
      unpack8(xv[i], f);
      unpack8(wr[v], g);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = g[j] * bf16_round(f[j] * rstd);
      yr[v] = pack8(f);
    }
  }
}

// dx (+)= rstd * (dy*w - xhat * mean(dy*w*xhat));  dw[c] += sum_rows dy * xhat   (fp32 atomics, one per CTA per column)
// Each CTA walks `rows_per_cta` rows so the dw partial stays in registers.
template <int MAX_VEC>
__global__ void __launch_bounds__(256, MAX_VEC == 1 ? 4 : 1)   // 4 CTAs per SM: the host sizes the grid for ONE wave of 148 x 4
rmsnorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                   const bf16* __restrict__ w, const float* __restrict__ rstd, bf16* __restrict__ dx,
                                   float* __restrict__ dw, int rows, int cols, long long ld, int rows_per_cta,
                                   int add_dx) {
  __shared__ float red[32];
  const int nvec = cols >> 3;
  float dwacc[MAX_VEC][8];
  float wv[MAX_VEC][8];
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int v = threadIdx.x + i * blockDim.x;
#pragma unroll
    for (int j = 0; j < 8; ++j) dwacc[i][j] = 0.f;
    if (v < nvec) unpack8(reinterpret_cast<const bf16x8*>(w)[v], wv[i]);
  }
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(rows, r0 + rows_per_cta);
  // Software-pipelined over the rows: the raw x / dy / dx vectors of row r + 1 are requested before row r's block
  // reduction, so the HBM latency of every row but the first hides behind the previous row's arithmetic and barriers
  // (the un-pipelined loop paid two dependent DRAM round trips per row: 83 us on [8786, 2048] against a 22 us roofline).
  bf16x8 nx[MAX_VEC], ng[MAX_VEC], nd[MAX_VEC];
  auto fetch = [&](int row) {
    const bf16x8* xr = reinterpret_cast<const bf16x8*>(x + (long long)row * ld);
    const bf16x8* gr = reinterpret_cast<const bf16x8*>(dy + (long long)row * ld);
    const bf16x8* dr = reinterpret_cast<const bf16x8*>(dx + (long long)row * ld);
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
      const int v = threadIdx.x + i * blockDim.x;
      if (v < nvec) {
        nx[i] = xr[v];
        ng[i] = gr[v];
        if (add_dx) nd[i] = dr[v];
      }
    }
  };
  if (r0 < r1) fetch(r0);
  for (int row = r0; row < r1; ++row) {
    bf16x8* dxr = reinterpret_cast<bf16x8*>(dx + (long long)row * ld);
    const float rs = rstd[row];
    float xh[MAX_VEC][8], gw[MAX_VEC][8];
    bf16x8 cd[MAX_VEC];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
      const int v = threadIdx.x + i * blockDim.x;
      if (v < nvec) {
        float g[8];
        unpack8(nx[i], xh[i]);
        unpack8(ng[i], g);
        cd[i] = nd[i];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] *= rs;
          dwacc[i][j] += g[j] * xh[i][j];
          gw[i][j] = g[j] * wv[i][j];
          dot += gw[i][j] * xh[i][j];
        }
      }
    }
    if (row + 1 < r1) fetch(row + 1);
    dot = block_sum(dot, red) / (float)cols;
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
      const int v = threadIdx.x + i * blockDim.x;
      if (v < nvec) {
        float o[8];
        if (add_dx) unpack8(cd[i], o);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = rs * (gw[i][j] - xh[i][j] * dot);
          o[j] = add_dx ? (o[j] + d) : d;
        }
        dxr[v] = pack8(o);
      }
    }
  }
  if (dw) {
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
      const int v = threadIdx.x + i * blockDim.x;
      if (v < nvec) red_add8_f32(dw + v * 8, dwacc[i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm (Qwen2-VL vision blocks, modeling_qwen2_vl.py:461-466; SigLIP): fp32 statistics, affine, bf16 out.
// ------------------------------------------------------------------------------------------------
template <int MAX_VEC>
__global__ void layernorm_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, const bf16* __restrict__ b,
                                     bf16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                     int cols, long long ld, float eps) {
  __shared__ float red[32];
  const long long row = blockIdx.x;
  const bf16x8* xr = reinterpret_cast<const bf16x8*>(x + row * ld);
  bf16x8* yr = reinterpret_cast<bf16x8*>(y + row * ld);
  const int nvec = cols >> 3;
  float xf[MAX_VEC][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int v = threadIdx.x + i * blockDim.x;
    if (v < nvec) {
      unpack8(xr[v], xf[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += xf[i][j];
    }
  }
  const float mean = block_sum(s, red) / (float)cols;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int v = threadIdx.x + i * blockDim.x;
    if (v < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = xf[i][j] - mean;
        ss += d * d;
      }
    }
  }
  const float rstd = rsqrtf(block_sum(ss, red) / (float)cols + eps);
  if (threadIdx.x == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int v = threadIdx.x + i * blockDim.x;
    if (v < nvec) {
      float g[8], bb[8], o[8];
      unpack8(reinterpret_cast<const bf16x8*>(w)[v], g);
      unpack8(reinterpret_cast<const bf16x8*>(b)[v], bb);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (xf[i][j] - mean) * rstd * g[j] + bb[j];
      yr[v] = pack8(o);
    }
  }
}

template <int MAX_VEC>
__global__ void layernorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                     const bf16* __restrict__ w, const float* __restrict__ mean,
                                     const float* __restrict__ rstd, bf16* __restrict__ dx, float* __restrict__ dw,
                                     float* __restrict__ db, int rows, int cols, long long ld, int rows_per_cta,
                                     int add_dx) {
  __shared__ float red[32];
  const int nvec = cols >> 3;
  float dwacc[MAX_VEC][8], dbacc[MAX_VEC][8], wv[MAX_VEC][8];
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int v = threadIdx.x + i * blockDim.x;
#pragma unroll
    for (int j = 0; j < 8; ++j) dwacc[i][j] = dbacc[i][j] = 0.f;
    if (v < nvec) unpack8(reinterpret_cast<const bf16x8*>(w)[v], wv[i]);
  }
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(rows, r0 + rows_per_cta);
  for (int row = r0; row < r1; ++row) {
    const bf16x8* xr = reinterpret_cast<const bf16x8*>(x + (long long)row * ld);
    const bf16x8* gr = reinterpret_cast<const bf16x8*>(dy + (long long)row * ld);
    bf16x8* dxr = reinterpret_cast<bf16x8*>(dx + (long long)row * ld);
    const float mu = mean[row], rs = rstd[row];
    float xh[MAX_VEC][8], gw[MAX_VEC][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
      const int v = threadIdx.x + i * blockDim.x;
      if (v < nvec) {
        float g[8];
        unpack8(xr[v], xh[i]);
        unpack8(gr[v], g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] = (xh[i][j] - mu) * rs;
          dwacc[i][j] += g[j] * xh[i][j];
          dbacc[i][j] += g[j];
          gw[i][j] = g[j] * wv[i][j];
          s1 += gw[i][j];
          s2 += gw[i][j] * xh[i][j];
        }
      }
    }
    s1 = block_sum(s1, red) / (float)cols;
    s2 = block_sum(s2, red) / (float)cols;
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
      const int v = threadIdx.x + i * blockDim.x;
      if (v < nvec) {
        float o[8];
        if (add_dx) unpack8(dxr[v], o);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = rs * (gw[i][j] - s1 - xh[i][j] * s2);
          o[j] = add_dx ? (o[j] + d) : d;
        }
        dxr[v] = pack8(o);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int v = threadIdx.x + i * blockDim.x;
    if (v < nvec) {
      if (dw) red_add8_f32(dw + v * 8, dwacc[i]);
      if (db) red_add8_f32(db + v * 8, dbacc[i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Rotary embedding, rotate-half form, in place on the first `heads` heads of a [tokens, heads_total, hd] buffer.
//   text  (apply_multimodal_rotary_pos_emb, modeling_qwen2_5_vl.py:627-669): bf16 cos/sin, every product and the sum
//         rounded to bf16 (bf16_ops = 1);
//   vision (apply_rotary_pos_emb_vision, :156-167): fp32 math, one rounding at the end (bf16_ops = 0).
// The table is [tokens, hd] fp32 with both halves equal, so the backward pass is the same kernel with sign = -1.
// ------------------------------------------------------------------------------------------------
__global__ void rope_kernel(bf16* __restrict__ x, const float* __restrict__ cs, const float* __restrict__ sn,
                            long long tokens, int heads, int hd, long long tok_stride, int bf16_ops, float sign) {
  const int half = hd >> 1;
  const long long total = tokens * heads * half;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % half);
    const long long th = i / half;
    const int h = (int)(th % heads);
    const long long t = th / heads;
    bf16* p = x + t * tok_stride + (long long)h * hd;
    const float x1 = __bfloat162float(p[d]), x2 = __bfloat162float(p[d + half]);
    float c1 = cs[t * hd + d], c2 = cs[t * hd + d + half];
    float s1 = sign * sn[t * hd + d], s2 = sign * sn[t * hd + d + half];
    float o1, o2;
    if (bf16_ops) {
      c1 = bf16_round(c1); c2 = bf16_round(c2); s1 = bf16_round(s1); s2 = bf16_round(s2);
      o1 = bf16_round(x1 * c1) + bf16_round(-x2 * s1);
      o2 = bf16_round(x2 * c2) + bf16_round(x1 * s2);
    } else {
      o1 = x1 * c1 - x2 * s1;
      o2 = x2 * c2 + x1 * s2;
    }
    p[d] = __float2bfloat16(o1);
    p[d + half] = __float2bfloat16(o2);
  }
}

// Vectorised form (half % 8 == 0, 16-byte aligned rows): one thread = 8 consecutive dims of both halves of one head -
// two 16-byte loads / stores of x and four float4 loads of the table (the two halves of the table are equal, so only the
// first is read). Same arithmetic, ~4x fewer memory instructions than the scalar kernel (54 -> ~15 us on [8786, 20, 128]).
__global__ void rope_vec8_kernel(bf16* __restrict__ x, const float* __restrict__ cs, const float* __restrict__ sn,
                                 long long tokens, int heads, int hd, long long tok_stride, int bf16_ops, float sign) {
  const int half = hd >> 1;
  const int vph = half >> 3;  // 8-wide vectors per half head
  const long long total = tokens * heads * vph;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vph);
    const long long th = i / vph;
    const int h = (int)(th % heads);
    const long long t = th / heads;
    bf16* p = x + t * tok_stride + (long long)h * hd + v * 8;
    float x1[8], x2[8], c[8], s_[8], o1[8], o2[8];
    unpack8(*reinterpret_cast<const bf16x8*>(p), x1);
    unpack8(*reinterpret_cast<const bf16x8*>(p + half), x2);
    const float4* cp = reinterpret_cast<const float4*>(cs + t * hd + v * 8);
    const float4* sp = reinterpret_cast<const float4*>(sn + t * hd + v * 8);
    const float4 c0 = cp[0], c1 = cp[1], s0 = sp[0], s1 = sp[1];
    c[0] = c0.x; c[1] = c0.y; c[2] = c0.z; c[3] = c0.w; c[4] = c1.x; c[5] = c1.y; c[6] = c1.z; c[7] = c1.w;
    s_[0] = s0.x; s_[1] = s0.y; s_[2] = s0.z; s_[3] = s0.w; s_[4] = s1.x; s_[5] = s1.y; s_[6] = s1.z; s_[7] = s1.w;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float cj = c[j], sj = sign * s_[j];
      if (bf16_ops) {
        cj = bf16_round(cj);
        sj = bf16_round(sj);
        o1[j] = bf16_round(x1[j] * cj) + bf16_round(-x2[j] * sj);
        o2[j] = bf16_round(x2[j] * cj) + bf16_round(x1[j] * sj);
      } else {
        o1[j] = x1[j] * cj - x2[j] * sj;
        o2[j] = x2[j] * cj + x1[j] * sj;
      }
    }
    *reinterpret_cast<bf16x8*>(p) = pack8(o1);
    *reinterpret_cast<bf16x8*>(p + half) = pack8(o2);
  }
}

// ------------------------------------------------------------------------------------------------
// Gated / plain activations.  act: 0 = SiLU-gate (silu(g) * u; Qwen2MLP :611-624, Qwen2_5_VLMLP :77-88),
// 1 = exact GELU (merger nn.GELU :139), 2 = quick-GELU x*sigmoid(1.702x) (Qwen2-VL vision MLP), 3 = tanh-GELU (SigLIP).
// HF evaluates act(g) in bf16 then multiplies in bf16: the intermediate is rounded.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_fwd(int act, float g) {
  switch (act) {
    case 0: return silu_f(g);
    case 1: return 0.5f * g * (1.f + erff(g * 0.70710678118654752f));
    case 2: return g / (1.f + __expf(-1.702f * g));
    default: {
      const float u = 0.7978845608028654f * (g + 0.044715f * g * g * g);
      return 0.5f * g * (1.f + tanhf(u));
    }
  }
}
__device__ __forceinline__ float act_grad(int act, float g) {
  switch (act) {
    case 0: return silu_grad_f(g);
    case 1: return 0.5f * (1.f + erff(g * 0.70710678118654752f)) + g * 0.3989422804014327f * __expf(-0.5f * g * g);
    case 2: {
      const float s = 1.f / (1.f + __expf(-1.702f * g));
      return s * (1.f + 1.702f * g * (1.f - s));
    }
    default: {
      const float u = 0.7978845608028654f * (g + 0.044715f * g * g * g);
      const float t = tanhf(u);
      return 0.5f * (1.f + t) + 0.5f * g * (1.f - t * t) * 0.7978845608028654f * (1.f + 3.f * 0.044715f * g * g);
    }
  }
}

// gate at gu[r*ld + c], up at gu[r*ld + up_off + c] (up_off < 0: ungated), out[r*out_ld + c]; cols % 8 == 0.
__global__ void act_mul_fwd_kernel(const bf16* __restrict__ gu, bf16* __restrict__ out, long long rows, int cols,
                                   long long ld, long long up_off, long long out_ld, int act) {
  if (threadIdx.x == 0) pdl_trigger();   // dependents may become resident (and prefetch) right away
  pdl_wait();   // no-op unless launched with the programmatic-serialization attribute (rollout decode chain)
  const int nvec = cols >> 3;
  const long long total = rows * nvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nvec;
    const int v = (int)(i % nvec);
    float g[8], u[8], o[8];
    unpack8(*reinterpret_cast<const bf16x8*>(gu + r * ld + v * 8), g);
    if (up_off >= 0) unpack8(*reinterpret_cast<const bf16x8*>(gu + r * ld + up_off + v * 8), u);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a = act_fwd(act, g[j]);
      o[j] = (up_off >= 0) ? bf16_round(a) * u[j] : a;
    }
    *reinterpret_cast<bf16x8*>(out + r * out_ld + v * 8) = pack8(o);
  }
}
// dgate = dout * up * act'(gate), dup = dout * act(gate); written to dgu with the same layout as gu (may alias gu).
__global__ void act_mul_bwd_kernel(const bf16* __restrict__ dout, const bf16* gu, bf16* dgu, long long rows, int cols,
                                   long long ld, long long up_off, long long dout_ld, int act) {
  const int nvec = cols >> 3;
  const long long total = rows * nvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nvec;
    const int v = (int)(i % nvec);
    float g[8], u[8], d[8], dg[8], du[8];
    unpack8(*reinterpret_cast<const bf16x8*>(gu + r * ld + v * 8), g);
    if (up_off >= 0) unpack8(*reinterpret_cast<const bf16x8*>(gu + r * ld + up_off + v * 8), u);
    unpack8(*reinterpret_cast<const bf16x8*>(dout + r * dout_ld + v * 8), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (up_off >= 0 && act == 0) {
        swiglu_bwd_f(d[j], g[j], u[j], dg[j], du[j]);
      } else if (up_off >= 0) {
        dg[j] = d[j] * u[j] * act_grad(act, g[j]);
        du[j] = d[j] * act_fwd(act, g[j]);
      } else {
        dg[j] = d[j] * act_grad(act, g[j]);
      }
    }
    *reinterpret_cast<bf16x8*>(dgu + r * ld + v * 8) = pack8(dg);
    if (up_off >= 0) *reinterpret_cast<bf16x8*>(dgu + r * ld + up_off + v * 8) = pack8(du);
  }
}

// ------------------------------------------------------------------------------------------------
// Masked row softmax, in place on bf16 scores S[z][q][k] (eager_attention_forward, modeling_qwen2_5_vl.py:182-206:
// fp32 softmax of bf16 scores, bf16 probabilities). Row q attends keys [lo[q], hi[q]) - causal, window and padding
// masks are all ranges; everything outside the range is written as exact zero (so skipped score tiles never leak).
// One warp per row.
// ------------------------------------------------------------------------------------------------
__global__ void softmax_rows_kernel(bf16* __restrict__ S, const int* __restrict__ lo, const int* __restrict__ hi,
                                    int Tq, int Tk, long long ld, long long z_stride, long long nrows, int hole_lo,
                                    int hole_hi) {
  const int lane = threadIdx.x & 31;
  const long long gw = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (gw >= nrows) return;
  const int q = (int)(gw % Tq);
  const long long z = gw / Tq;
  bf16* row = S + z * z_stride + (long long)q * ld;
  const int a = max(0, lo[q]), b = min(Tk, hi[q]);
  // keys in [hole_lo, hole_hi) are padding between two key segments (shared-prefix layout) and are always masked
  float mx = -INFINITY;
  for (int k = a + lane; k < b; k += 32)
    if (k < hole_lo || k >= hole_hi) mx = fmaxf(mx, __bfloat162float(row[k]));
  mx = warp_max(mx);
  float sum = 0.f;
  for (int k = a + lane; k < b; k += 32)
    if (k < hole_lo || k >= hole_hi) sum += __expf(__bfloat162float(row[k]) - mx);
  sum = warp_sum(sum);
  const float inv = (sum > 0.f) ? 1.f / sum : 0.f;
  for (int k = lane; k < Tk; k += 32) {
    float p = 0.f;
    if (k >= a && k < b && (k < hole_lo || k >= hole_hi)) p = __expf(__bfloat162float(row[k]) - mx) * inv;
    row[k] = __float2bfloat16(p);
  }
}
// dS = P * (dP - sum_k dP*P) in place on dP; exact zero outside the row's range.
__global__ void softmax_bwd_rows_kernel(const bf16* __restrict__ P, bf16* __restrict__ dP, const int* __restrict__ lo,
                                        const int* __restrict__ hi, int Tq, int Tk, long long ld, long long z_stride,
                                        long long nrows, int hole_lo, int hole_hi) {
  const int lane = threadIdx.x & 31;
  const long long gw = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (gw >= nrows) return;
  const int q = (int)(gw % Tq);
  const long long z = gw / Tq;
  const bf16* p = P + z * z_stride + (long long)q * ld;
  bf16* d = dP + z * z_stride + (long long)q * ld;
  const int a = max(0, lo[q]), b = min(Tk, hi[q]);
  float dot = 0.f;
  for (int k = a + lane; k < b; k += 32)
    if (k < hole_lo || k >= hole_hi) dot += __bfloat162float(p[k]) * __bfloat162float(d[k]);
  dot = warp_sum(dot);
  for (int k = lane; k < Tk; k += 32) {
    float o = 0.f;
    if (k >= a && k < b && (k < hole_lo || k >= hole_hi)) o = __bfloat162float(p[k]) * (__bfloat162float(d[k]) - dot);
    d[k] = __float2bfloat16(o);
  }
}

// ------------------------------------------------------------------------------------------------
// Row gather / scatter-add.  out[r] = src[index[r]] where index >= 0 addresses `table`, index < 0 addresses
// `alt` row (-1 - index) - this is `embed_tokens(ids)` followed by `masked_scatter` of the image embeddings
// (modeling_qwen2_5_vl.py:1298-1307) in one pass; with alt == nullptr it is a plain row gather (window reorder
// :478-484, :512-513).
// ------------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const bf16* __restrict__ table, const bf16* __restrict__ alt,
                                   const int* __restrict__ index, bf16* __restrict__ out, long long rows, int cols,
                                   long long table_ld, long long alt_ld, long long out_ld) {
  const int nvec = cols >> 3;
  const long long total = rows * nvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nvec;
    const int v = (int)(i % nvec);
    const int idx = index[r];
    const bf16* src = (idx >= 0) ? table + (long long)idx * table_ld : alt + (long long)(-1 - idx) * alt_ld;
    *reinterpret_cast<uint4*>(out + r * out_ld + v * 8) = *reinterpret_cast<const uint4*>(src + v * 8);
  }
}
// Backward of the above: dtable[index[r]] += d[r] (fp32 atomics; duplicate ids collide), dalt[-1-index[r]] += d[r].
__global__ void scatter_add_rows_kernel(const bf16* __restrict__ d, const int* __restrict__ index,
                                        float* __restrict__ dtable, float* __restrict__ dalt, long long rows, int cols,
                                        long long d_ld, long long table_ld, long long alt_ld) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i % cols);
    const int idx = index[r];
    const float g = __bfloat162float(d[r * d_ld + c]);
    if (idx >= 0) {
      if (dtable) atomicAdd(dtable + (long long)idx * table_ld + c, g);
    } else if (dalt) {
      atomicAdd(dalt + (long long)(-1 - idx) * alt_ld + c, g);
    }
  }
}

// out[c] += sum_r x[r][c]  (bias gradients). Grid (ceil(cols/64), row_chunks); block (64, 4).
__global__ void colsum_kernel(const bf16* __restrict__ x, float* __restrict__ out, long long rows, int cols,
                              long long ld, long long rows_per_cta) {
  __shared__ float part[4][64];
  const int c = blockIdx.x * 64 + threadIdx.x;
  const long long r0 = blockIdx.y * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  float acc = 0.f;
  if (c < cols)
    for (long long r = r0 + threadIdx.y; r < r1; r += 4) acc += __bfloat162float(x[r * ld + c]);
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols)
    atomicAdd(out + c, part[0][threadIdx.x] + part[1][threadIdx.x] + part[2][threadIdx.x] + part[3][threadIdx.x]);
}


// out[r][kvh*hd + d] = sum_{j<g} src[r][(kvh*g + j)*hd + d]: folds the per-query-head dK / dV of grouped-query
// attention back onto the shared kv heads (autograd of repeat_kv, modeling_qwen2_5_vl.py:170-179).
__global__ void group_sum_kernel(const bf16* __restrict__ src, bf16* __restrict__ out, long long rows, int nkv, int g,
                                 int hd, long long src_ld, long long out_ld, int accumulate) {
  const long long total = rows * nkv * hd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % hd);
    const long long t = i / hd;
    const int kvh = (int)(t % nkv);
    const long long r = t / nkv;
    float acc = 0.f;
    for (int j = 0; j < g; ++j) acc += __bfloat162float(src[r * src_ld + (long long)(kvh * g + j) * hd + d]);
    bf16* o = out + r * out_ld + (long long)kvh * hd + d;
    if (accumulate) acc += __bfloat162float(*o);
    *o = __float2bfloat16(acc);
  }
}

// Elementwise helpers ---------------------------------------------------------------------------
__global__ void add_bf16_kernel(const bf16* a, const bf16* b, bf16* out, long long n8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float x[8], y[8];
    unpack8(reinterpret_cast<const bf16x8*>(a)[i], x);
    unpack8(reinterpret_cast<const bf16x8*>(b)[i], y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
    reinterpret_cast<bf16x8*>(out)[i] = pack8(x);
  }
}
// dst_bf16[r][c] = src_f32[r][c] (strided rows): used to hand fp32 image-embedding grads back to the bf16 chain.
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}

// Per-row log-sum-exp finalisation of the fused lm_head (GEMM EPI_LSE partials) -> selected-token log-prob.
// Replaces `logits.log_softmax(-1)` + `gather` (sc_grpo_trainer.py:510-513) without materialising [T, V].
__global__ void lse_finalize_kernel(const float* __restrict__ pmax, const float* __restrict__ psum,
                                    const float* __restrict__ tgt, int tiles_n, int M, float* __restrict__ lse,
                                    float* __restrict__ logp) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= M) return;
  float mx = -INFINITY;
  for (int t = lane; t < tiles_n; t += 32) mx = fmaxf(mx, pmax[(long long)row * tiles_n + t]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int t = lane; t < tiles_n; t += 32)
    s += psum[(long long)row * tiles_n + t] * __expf(pmax[(long long)row * tiles_n + t] - mx);
  s = warp_sum(s);
  if (lane == 0) {
    const float l = mx + logf(s);
    lse[row] = l;
    if (logp) logp[row] = tgt[row] - l;
  }
}

static inline int grid_for(long long work, int block, int cap_mult = 8) {
  long long g = (work + block - 1) / block;
  const long long cap = 148LL * cap_mult;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

void trace_install_rowops(unsigned long long* p) { trace_install_tu(p); }
}  // namespace iadr1

using namespace iadr1;

#define NORM_DISPATCH(KERNEL, cols, ...)                                                     \
  do {                                                                                       \
    const int nvec__ = (cols) >> 3;                                                          \
    if (nvec__ <= 256) { KERNEL<1><<<grid__, 256, 0, st>>>(__VA_ARGS__); }                   \
    else if (nvec__ <= 512) { KERNEL<2><<<grid__, 256, 0, st>>>(__VA_ARGS__); }              \
    else if (nvec__ <= 1024) { KERNEL<4><<<grid__, 256, 0, st>>>(__VA_ARGS__); }             \
    else return set_error(#KERNEL ": cols %d too large (max 8192)", (int)(cols));            \
  } while (0)

extern "C" {

int iadr1_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, long long rows, int cols, long long x_ld,
                      long long y_ld, float eps, void* stream) {
  if (rows <= 0) return 0;
  if (cols % 8 || x_ld % 8 || y_ld % 8) return set_error("rmsnorm_fwd: cols and strides must be multiples of 8");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid__ = (int)rows;
  NORM_DISPATCH(rmsnorm_fwd_kernel, cols, (const bf16*)x, (const bf16*)w, (bf16*)y, rstd, cols, x_ld, y_ld, eps);
  IADR1_CHECK_LAUNCH("rmsnorm_fwd");
  return 0;
}

int iadr1_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, void* dx, float* dw,
                      long long rows, int cols, long long ld, int add_dx, void* stream) {
  if (rows <= 0) return 0;
  if (cols % 8 || ld % 8) return set_error("rmsnorm_bwd: cols and ld must be multiples of 8");
  cudaStream_t st = (cudaStream_t)stream;
  const int rows_per_cta = (int)((rows + 148 * 4 - 1) / (148 * 4));
  const int grid__ = (int)((rows + rows_per_cta - 1) / rows_per_cta);
  NORM_DISPATCH(rmsnorm_bwd_kernel, cols, (const bf16*)dy, (const bf16*)x, (const bf16*)w, rstd, (bf16*)dx, dw,
                (int)rows, cols, ld, rows_per_cta, add_dx);
  IADR1_CHECK_LAUNCH("rmsnorm_bwd");
  return 0;
}

int iadr1_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mean, float* rstd, long long rows,
                        int cols, long long ld, float eps, void* stream) {
  if (rows <= 0) return 0;
  if (cols % 8 || ld % 8) return set_error("layernorm_fwd: cols and ld must be multiples of 8");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid__ = (int)rows;
  NORM_DISPATCH(layernorm_fwd_kernel, cols, (const bf16*)x, (const bf16*)w, (const bf16*)b, (bf16*)y, mean, rstd, cols,
                ld, eps);
  IADR1_CHECK_LAUNCH("layernorm_fwd");
  return 0;
}

int iadr1_layernorm_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd, void* dx,
                        float* dw, float* db, long long rows, int cols, long long ld, int add_dx, void* stream) {
  if (rows <= 0) return 0;
  if (cols % 8 || ld % 8) return set_error("layernorm_bwd: cols and ld must be multiples of 8");
  cudaStream_t st = (cudaStream_t)stream;
  const int rows_per_cta = (int)((rows + 148 * 4 - 1) / (148 * 4));
  const int grid__ = (int)((rows + rows_per_cta - 1) / rows_per_cta);
  NORM_DISPATCH(layernorm_bwd_kernel, cols, (const bf16*)dy, (const bf16*)x, (const bf16*)w, mean, rstd, (bf16*)dx, dw,
                db, (int)rows, cols, ld, rows_per_cta, add_dx);
  IADR1_CHECK_LAUNCH("layernorm_bwd");
  return 0;
}

int iadr1_rope(void* x, const float* cos_t, const float* sin_t, long long tokens, int heads, int hd,
               long long tok_stride, int bf16_ops, int backward, void* stream) {
  if (tokens <= 0 || heads <= 0) return 0;
  if (hd % 2) return set_error("rope: head_dim must be even");
  const long long work = tokens * heads * (hd / 2);
  if ((hd / 2) % 8 == 0 && tok_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(cos_t) & 15) == 0 && (reinterpret_cast<uintptr_t>(sin_t) & 15) == 0) {
    rope_vec8_kernel<<<grid_for(work / 8, 256, 16), 256, 0, (cudaStream_t)stream>>>(
        (bf16*)x, cos_t, sin_t, tokens, heads, hd, tok_stride, bf16_ops, backward ? -1.f : 1.f);
    IADR1_CHECK_LAUNCH("rope");
    return 0;
  }
  rope_kernel<<<grid_for(work, 256, 16), 256, 0, (cudaStream_t)stream>>>((bf16*)x, cos_t, sin_t, tokens, heads, hd,
                                                                         tok_stride, bf16_ops, backward ? -1.f : 1.f);
  IADR1_CHECK_LAUNCH("rope");
  return 0;
}

int iadr1_act_mul_fwd(const void* gu, void* out, long long rows, int cols, long long ld, long long up_off,
                      long long out_ld, int act, void* stream) {
  if (rows <= 0) return 0;
  if (cols % 8 || ld % 8 || out_ld % 8 || (up_off > 0 && up_off % 8)) return set_error("act_mul_fwd: alignment");
  launch_kernel(act_mul_fwd_kernel, dim3(grid_for(rows * (cols / 8), 256, 16)), dim3(256), 0, (cudaStream_t)stream,
                (const bf16*)gu, (bf16*)out, rows, cols, ld, up_off, out_ld, act);
  IADR1_CHECK_LAUNCH("act_mul_fwd");
  return 0;
}

int iadr1_act_mul_bwd(const void* dout, const void* gu, void* dgu, long long rows, int cols, long long ld,
                      long long up_off, long long dout_ld, int act, void* stream) {
  if (rows <= 0) return 0;
  if (cols % 8 || ld % 8 || dout_ld % 8 || (up_off > 0 && up_off % 8)) return set_error("act_mul_bwd: alignment");
  act_mul_bwd_kernel<<<grid_for(rows * (cols / 8), 256, 16), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)dout, (const bf16*)gu, (bf16*)dgu, rows, cols, ld, up_off, dout_ld, act);
  IADR1_CHECK_LAUNCH("act_mul_bwd");
  return 0;
}

int iadr1_softmax_rows(void* S, const int* lo, const int* hi, int Tq, int Tk, long long ld, long long z_stride,
                       long long batch, int hole_lo, int hole_hi, void* stream) {
  const long long nrows = batch * Tq;
  if (nrows <= 0) return 0;
  const long long blocks = (nrows * 32 + 255) / 256;
  softmax_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((bf16*)S, lo, hi, Tq, Tk, ld, z_stride,
                                                                          nrows, hole_lo, hole_hi);
  IADR1_CHECK_LAUNCH("softmax_rows");
  return 0;
}

int iadr1_softmax_bwd_rows(const void* P, void* dP, const int* lo, const int* hi, int Tq, int Tk, long long ld,
                           long long z_stride, long long batch, int hole_lo, int hole_hi, void* stream) {
  const long long nrows = batch * Tq;
  if (nrows <= 0) return 0;
  const long long blocks = (nrows * 32 + 255) / 256;
  softmax_bwd_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)P, (bf16*)dP, lo, hi, Tq,
                                                                              Tk, ld, z_stride, nrows, hole_lo, hole_hi);
  IADR1_CHECK_LAUNCH("softmax_bwd_rows");
  return 0;
}

int iadr1_gather_rows(const void* table, const void* alt, const int* index, void* out, long long rows, int cols,
                      long long table_ld, long long alt_ld, long long out_ld, void* stream) {
  if (rows <= 0) return 0;
  if (cols % 8 || table_ld % 8 || out_ld % 8 || (alt && alt_ld % 8)) return set_error("gather_rows: alignment");
  gather_rows_kernel<<<grid_for(rows * (cols / 8), 256, 16), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)table, (const bf16*)alt, index, (bf16*)out, rows, cols, table_ld, alt_ld, out_ld);
  IADR1_CHECK_LAUNCH("gather_rows");
  return 0;
}

int iadr1_scatter_add_rows(const void* d, const int* index, float* dtable, float* dalt, long long rows, int cols,
                           long long d_ld, long long table_ld, long long alt_ld, void* stream) {
  if (rows <= 0) return 0;
  scatter_add_rows_kernel<<<grid_for(rows * cols, 256, 16), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)d, index, dtable, dalt, rows, cols, d_ld, table_ld, alt_ld);
  IADR1_CHECK_LAUNCH("scatter_add_rows");
  return 0;
}

int iadr1_colsum(const void* x, float* out, long long rows, int cols, long long ld, void* stream) {
  if (rows <= 0 || cols <= 0) return 0;
  const int cblocks = (cols + 63) / 64;
  int rchunks = (148 * 4 + cblocks - 1) / cblocks;
  if (rchunks > (rows + 31) / 32) rchunks = (int)((rows + 31) / 32);
  if (rchunks < 1) rchunks = 1;
  const long long rows_per_cta = (rows + rchunks - 1) / rchunks;
  colsum_kernel<<<dim3(cblocks, rchunks), dim3(64, 4), 0, (cudaStream_t)stream>>>((const bf16*)x, out, rows, cols, ld,
                                                                                  rows_per_cta);
  IADR1_CHECK_LAUNCH("colsum");
  return 0;
}

int iadr1_group_sum(const void* src, void* out, long long rows, int nkv, int g, int hd, long long src_ld,
                    long long out_ld, int accumulate, void* stream) {
  if (rows <= 0) return 0;
  group_sum_kernel<<<grid_for(rows * nkv * hd, 256, 16), 256, 0, (cudaStream_t)stream>>>((const bf16*)src, (bf16*)out,
                                                                                        rows, nkv, g, hd, src_ld, out_ld, accumulate);
  IADR1_CHECK_LAUNCH("group_sum");
  return 0;
}

int iadr1_add_bf16(const void* a, const void* b, void* out, long long n, void* stream) {
  if (n <= 0) return 0;
  if (n % 8) return set_error("add_bf16: n must be a multiple of 8");
  add_bf16_kernel<<<grid_for(n / 8, 256, 16), 256, 0, (cudaStream_t)stream>>>((const bf16*)a, (const bf16*)b,
                                                                              (bf16*)out, n / 8);
  IADR1_CHECK_LAUNCH("add_bf16");
  return 0;
}

int iadr1_cast_f32_bf16(const float* src, void* dst, long long n, void* stream) {
  if (n <= 0) return 0;
  cast_f32_bf16_kernel<<<grid_for(n, 256, 16), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, n);
  IADR1_CHECK_LAUNCH("cast_f32_bf16");
  return 0;
}

int iadr1_lse_finalize(const float* pmax, const float* psum, const float* tgt, int tiles_n, int M, float* lse,
                       float* logp, void* stream) {
  if (M <= 0) return 0;
  const int blocks = (M * 32 + 255) / 256;
  lse_finalize_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pmax, psum, tgt, tiles_n, M, lse, logp);
  IADR1_CHECK_LAUNCH("lse_finalize");
  return 0;
}

}  // extern "C"
