// Rollout (autoregressive decode) kernels: the in-rank replacement for `self.llm.generate(...)`
// (ref: train/stage_rl/trainer/sc_grpo_trainer.py:343-358, 667 - a vLLM 0.7.3 engine on a separate GPU).
//
// One decode step for R rows in lock-step is a fixed kernel sequence with NO host-visible state: the step counter,
// prompt length and per-row tokens live in device memory (`DecodeState`), so the whole step is captured once in a
// CUDA graph and replayed max_completion_length times. Dense products use the tcgen05 GEMM with operands swapped
// (weights as the 128-row operand, the R activations as the narrow one) and split-K fp32 atomics straight into the
// fp32 residual stream; this file holds everything between those GEMMs:
//   embed lookup -> [RMSNorm(f32 in) -> qkv -> rope + KV append -> split-KV attention -> o_proj -> RMSNorm -> MLP] x L
//   -> RMSNorm -> lm_head -> fused sampler (temperature -> top-k -> top-p -> multinomial, Philox).
// KV cache: the prompt's K/V are stored ONCE per group ([group][P][kv_head][hd], shared by its G rows - the prefix
// sharing vLLM gets from `enable_prefix_caching`, sc_grpo_trainer.py:351); each row appends to its own
// [row][C][kv_head][hd] slab, so decode reads are contiguous per (row, kv_head) stream.
#include "runtime.h"
#include <cuda_bf16.h>
#include <curand_kernel.h>
#include <stdint.h>
#include <cstdlib>
#include <string>

namespace iadr1 {

using bf16 = __nv_bfloat16;

// Device-resident rollout state (int32 words). Host writes it once per rollout.
enum StateWord : int {
  ST_STEP = 0,       // index of the completion token fed this step (KV slot); logits predict token ST_STEP + 1
  ST_EOS2 = 1,       // second stop token + 1 (generation_config.json lists several eos ids; 0 = none)
  ST_UNFINISHED = 2, // rows that have not produced EOS yet (host polls for early exit)
  ST_SEED_LO = 4,    // per-generate() sampling seed kept in DEVICE memory (XORed with the kernel's seed argument): a
  ST_SEED_HI = 5,    // captured CUDA graph replays with frozen arguments, so the seed must not live in them
  ST_WORDS = 8
};

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }

// N consecutive bf16 (N = 1, 2 or 4) -> fp32 with one load
template <int N>
__device__ __forceinline__ void load_bf16_vec(const bf16* p, float (&out)[N]) {
  if constexpr (N == 4) {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y));
    out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
  } else if constexpr (N == 2) {
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
    out[0] = a.x; out[1] = a.y;
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) out[i] = __bfloat162float(p[i]);
  }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// N consecutive bf16 kept as raw bits until needed (keeps many cache rows in flight with few registers)
template <int N> struct RawVec;
template <> struct RawVec<4> {
  uint2 v;
  __device__ __forceinline__ void load(const bf16* p) { v = *reinterpret_cast<const uint2*>(p); }
  __device__ __forceinline__ void unpack(float (&o)[4]) const {
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y));
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
  }
};
template <> struct RawVec<2> {
  uint32_t v;
  __device__ __forceinline__ void load(const bf16* p) { v = *reinterpret_cast<const uint32_t*>(p); }
  __device__ __forceinline__ void unpack(float (&o)[2]) const {
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v));
    o[0] = a.x; o[1] = a.y;
  }
};
template <> struct RawVec<1> {
  bf16 v;
  __device__ __forceinline__ void load(const bf16* p) { v = *p; }
  __device__ __forceinline__ void unpack(float (&o)[1]) const { o[0] = __bfloat162float(v); }
};

// h[r][:] = float(embed[tok[r]][:])
__global__ void decode_embed_kernel(const bf16* __restrict__ embed, const int* __restrict__ tok, float* __restrict__ h,
                                    int H) {
  if (threadIdx.x == 0) pdl_trigger();   // dependents may become resident (and prefetch) right away
  pdl_wait();
  const int r = blockIdx.x;
  const bf16* src = embed + (long long)tok[r] * H;
  for (int c = threadIdx.x; c < H; c += blockDim.x) h[(long long)r * H + c] = __bfloat162float(src[c]);
}

// RMSNorm with fp32 input (the decode residual stream) -> bf16, same rounding order as Qwen2RMSNorm.
__global__ void rmsnorm_f32in_kernel(const float* __restrict__ x, const bf16* __restrict__ w, bf16* __restrict__ y,
                                     int cols, float eps, float* __restrict__ zero_buf, int zero_per_row) {
  if (threadIdx.x == 0) pdl_trigger();   // dependents may become resident (and prefetch) right away
  __shared__ float red[32];
  const long long row = blockIdx.x;
  const float* xr = x + row * cols;
  constexpr int MAXE = 16;  // 256 threads x 16 = 4096 columns held in registers
  float xv[MAXE], wv[MAXE];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXE; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    if (c < cols) wv[i] = __bfloat162float(w[c]);   // static weights: requested before the PDL wait (DRAM miss hidden)
  }
  pdl_wait();
#pragma unroll
  for (int i = 0; i < MAXE; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    if (c < cols) {
      xv[i] = bf16r(xr[c]);
      ss += xv[i] * xv[i];
    }
  }
  for (int c = threadIdx.x + MAXE * blockDim.x; c < cols; c += blockDim.x) {
    const float v = bf16r(xr[c]);
    ss += v * v;
  }
  // clear this row's slice of the next split-K GEMM's fp32 accumulation target (saves a memset node per layer)
  for (int c = threadIdx.x; c < zero_per_row; c += blockDim.x) zero_buf[row * zero_per_row + c] = 0.f;
  ss = wsum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
  const float rstd = rsqrtf(tot / (float)cols + eps);
#pragma unroll
  for (int i = 0; i < MAXE; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    if (c < cols) y[row * cols + c] = __float2bfloat16(wv[i] * bf16r(xv[i] * rstd));
  }
  for (int c = threadIdx.x + MAXE * blockDim.x; c < cols; c += blockDim.x)
    y[row * cols + c] = __float2bfloat16(__bfloat162float(w[c]) * bf16r(bf16r(xr[c]) * rstd));
}

// ------------------------------------------------------------------------------------------------
// Fused decode attention: rotary on q/k (HF bf16 op order) + KV append + split-KV attention + combine in ONE launch.
// Grid (R, nkv, nsplit), 4 warps. Per CTA:
//   1. the gq rotated, pre-scaled queries of this kv head are built once into shared memory;
//   2. the chunk's K/V rows are staged in shared memory with ONE round of cp.async (one memory latency per CTA); K is
//      stored with its 16-byte pieces XOR-swizzled by the key index so that "one lane = one key" reads are conflict-free;
//      the token being decoded is rotated, appended to the row's slab and dropped into the staging buffers by warp 0;
//   3. warp w owns up to two query heads and walks the chunk in tiles of 32 keys, flash-decoding style:
//        phase A (lane = key): full-length dot products from shared memory - no per-key shuffle reductions;
//        one warp max / sum per TILE for the online softmax;
//        phase B (lane = head dims): probabilities broadcast by shuffle, P.V accumulated from the V rows;
//   4. partials go to `part`; the last CTA of each (row, kv head) to finish (atomic ticket) merges the splits.
// ------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(128) decode_attn_fused_kernel(
    const float* __restrict__ qkv, const float* __restrict__ cos_tab, const float* __restrict__ sin_tab,
    const int* __restrict__ rope_delta, const bf16* __restrict__ kp, const bf16* __restrict__ vp, bf16* __restrict__ kc,
    bf16* __restrict__ vc, const int* __restrict__ state, const int* __restrict__ row_group,
    const int* __restrict__ row_plen, float* __restrict__ part, int* __restrict__ tickets, bf16* __restrict__ out, int nq,
    int nkv, int p_max, int c_max, int chunk, int max_pos, float scale, int dbg) {
  if (threadIdx.x == 0) pdl_trigger();   // dependents may become resident (and prefetch) right away
  constexpr int DPL = HD / 32;
  constexpr int MAXG = 8;
  constexpr int HPW = 2;             // heads per warp (MAXG / 4 warps)
  constexpr int HALF = HD / 2;
  constexpr int PPR = HD / 8;        // 16-byte pieces per K/V row
  constexpr int SWZ = PPR >= 8 ? 7 : PPR - 1;
  const int r = blockIdx.x, kvh = blockIdx.y, sp = blockIdx.z, nsplit = gridDim.z;
  const int gq = nq / nkv;
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int P = row_plen[r];
  const int step = state[ST_STEP];
  const int ctx = P + step + 1;
  const int k0 = sp * chunk, k1 = min(ctx, k0 + chunk);
  const int nkeys = max(0, k1 - k0);
  const int grp = row_group[r];
  const int qkv_dim = (nq + 2 * nkv) * HD;
  const float* xrow = qkv + (long long)r * qkv_dim;
  int pos = P + step + rope_delta[r];
  pos = max(0, min(max_pos - 1, pos));
  const float* cs = cos_tab + (long long)pos * HD;
  const float* sn = sin_tab + (long long)pos * HD;

  // rotate-half with bf16 rounding of every product (apply_multimodal_rotary_pos_emb)
  auto rot = [&](const float* x, int d) -> float {
    const float xd = bf16r(x[d]);
    const float xp = bf16r(x[d < HALF ? d + HALF : d - HALF]);
    const float a = bf16r(xd * bf16r(cs[d]));
    const float b = bf16r((d < HALF ? -xp : xp) * bf16r(sn[d]));
    return bf16r(a + b);
  };

  __shared__ __align__(16) float sm_q[MAXG][HD];
  __shared__ int s_last;
  extern __shared__ __align__(128) uint8_t sm_kv_raw[];
  bf16* sK = reinterpret_cast<bf16*>(sm_kv_raw);   // [chunk][HD], 16-byte pieces swizzled by (key & SWZ)
  bf16* sV = sK + (size_t)chunk * HD;              // [chunk][HD], linear

  // ---- 1 + 2: stage K/V (async) while the queries are rotated ----
  if (!(dbg & 2) && !(dbg & 16)) {
    for (int q = threadIdx.x; q < nkeys * PPR; q += blockDim.x) {
      const int jj = q / PPR, piece = q % PPR;
      const int j = k0 + jj;
      if (j == ctx - 1) continue;
      const bf16* krow;
      const bf16* vrow;
      if (j < P) {
        krow = kp + (((long long)grp * p_max + j) * nkv + kvh) * HD;
        vrow = vp + (((long long)grp * p_max + j) * nkv + kvh) * HD;
      } else {
        krow = kc + (((long long)r * c_max + (j - P)) * nkv + kvh) * HD;
        vrow = vc + (((long long)r * c_max + (j - P)) * nkv + kvh) * HD;
      }
      cp_async16(sK + (size_t)jj * HD + ((piece ^ (jj & SWZ)) << 3), krow + piece * 8);
      cp_async16(sV + (size_t)jj * HD + (piece << 3), vrow + piece * 8);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // everything above reads only scalars / tables / cache rows written at least two kernels ago; the qkv row below is
  // produced by the immediately preceding GEMM
  pdl_wait();
  for (int i = threadIdx.x; i < gq * HD; i += blockDim.x) {
    const int h = i / HD, d = i % HD;
    sm_q[h][d] = (dbg & 1) ? 0.f : rot(xrow + (long long)(kvh * gq + h) * HD, d) * scale;
  }
  if (!(dbg & 2) && ctx - 1 >= k0 && ctx - 1 < k1 && warp == 0) {
    const int jj = ctx - 1 - k0;
    const float* knew = xrow + (long long)(nq + kvh) * HD;
    const float* vnew = xrow + (long long)(nq + nkv + kvh) * HD;
    bf16* kdst = kc + (((long long)r * c_max + step) * nkv + kvh) * HD;
    bf16* vdst = vc + (((long long)r * c_max + step) * nkv + kvh) * HD;
#pragma unroll
    for (int e = 0; e < DPL; ++e) {
      const int d = lane * DPL + e;
      const bf16 kb = __float2bfloat16(rot(knew, d));
      const bf16 vb = __float2bfloat16(vnew[d]);
      sK[(size_t)jj * HD + ((((d >> 3) ^ (jj & SWZ)) << 3) | (d & 7))] = kb;
      sV[(size_t)jj * HD + d] = vb;
      kdst[d] = kb;
      vdst[d] = vb;
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- 3: flash-decoding over tiles of 32 keys; this warp's heads are [h0, h0 + hpw) ----
  const int hpw = (gq + 3) >> 2;
  const int h0 = warp * hpw;
  float acc[HPW][DPL], mrun[HPW], lrun[HPW];
  bool on[HPW];
#pragma unroll
  for (int hh = 0; hh < HPW; ++hh) {
    mrun[hh] = -INFINITY;
    lrun[hh] = 0.f;
    on[hh] = hh < hpw && h0 + hh < gq;
#pragma unroll
    for (int e = 0; e < DPL; ++e) acc[hh][e] = 0.f;
  }
  // The tile loop is written branch-free (clamped indices, zero probabilities for out-of-range keys, clamped head
  // ids for unused head slots) so that every shuffle is provably convergent - no per-shuffle WARPSYNC/BSSY pairs.
  const int hq[HPW] = {min(h0, gq - 1), min(h0 + 1, gq - 1)};
  const int ntiles = ((dbg & 2) || (dbg & 8)) ? 0 : (nkeys + 31) >> 5;
  for (int ti = 0; ti < ntiles; ++ti) {
    const int t0 = ti << 5;
    const int jj = t0 + lane;
    const bool valid = jj < nkeys;
    const int jc = min(jj, nkeys - 1);
    // phase A: lane = key
    float sc[HPW] = {0.f, 0.f};
    {
      const uint4* krow = reinterpret_cast<const uint4*>(sK + (size_t)jc * HD);
#pragma unroll 4
      for (int pi = 0; pi < PPR; ++pi) {
        const uint4 kv = krow[pi ^ (jc & SWZ)];
        float kf[8];
        const __nv_bfloat162* kh = reinterpret_cast<const __nv_bfloat162*>(&kv);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(kh[e]);
          kf[2 * e] = f.x;
          kf[2 * e + 1] = f.y;
        }
#pragma unroll
        for (int hh = 0; hh < HPW; ++hh) {
          const float4 qa = *reinterpret_cast<const float4*>(&sm_q[hq[hh]][pi * 8]);
          const float4 qb = *reinterpret_cast<const float4*>(&sm_q[hq[hh]][pi * 8 + 4]);
          // two independent 4-term chains per head (shorter dependency chain than one 8-term sum)
          const float s0 = qa.x * kf[0] + qa.y * kf[1] + qa.z * kf[2] + qa.w * kf[3];
          const float s1 = qb.x * kf[4] + qb.y * kf[5] + qb.z * kf[6] + qb.w * kf[7];
          sc[hh] += s0 + s1;
        }
      }
    }
    float pr[HPW];
#pragma unroll
    for (int hh = 0; hh < HPW; ++hh) {
      const float sv = valid ? sc[hh] : -INFINITY;
      float mt = sv;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, o));
      const float mn = fmaxf(mrun[hh], mt);
      const float corr = (mrun[hh] == -INFINITY) ? 0.f : __expf(mrun[hh] - mn);
      pr[hh] = valid ? __expf(sv - mn) : 0.f;
      lrun[hh] = lrun[hh] * corr + wsum(pr[hh]);
#pragma unroll
      for (int e = 0; e < DPL; ++e) acc[hh][e] *= corr;
      mrun[hh] = mn;
    }
    // phase B: lane = head dims; probabilities are broadcast from the lane that owns the key (0 beyond the chunk)
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      float vf[DPL];
      load_bf16_vec<DPL>(sV + (size_t)min(t0 + k, nkeys - 1) * HD + lane * DPL, vf);
#pragma unroll
      for (int hh = 0; hh < HPW; ++hh) {
        const float pk = __shfl_sync(0xffffffffu, pr[hh], k);
#pragma unroll
        for (int e = 0; e < DPL; ++e) acc[hh][e] += pk * vf[e];
      }
    }
  }
  // each warp writes the partials of its own heads: part[r][head][split][HD + 2]
#pragma unroll
  for (int hh = 0; hh < HPW; ++hh) {
    if (on[hh]) {
      float* dst = part + (((long long)r * nq + kvh * gq + h0 + hh) * nsplit + sp) * (HD + 2);
#pragma unroll
      for (int e = 0; e < DPL; ++e) dst[lane * DPL + e] = acc[hh][e];
      if (lane == 0) {
        dst[HD] = mrun[hh];
        dst[HD + 1] = lrun[hh];
      }
    }
  }
  if (dbg & 4) return;
  // ---- 4: last CTA of this (row, kv head) merges the splits ----
  // bar.sync orders the CTA's partial writes before thread 0's (cumulative) gpu-scope fence + ticket; one fence per CTA.
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    const int t = atomicAdd(&tickets[r * nkv + kvh], 1);
    s_last = (t == nsplit - 1);
    if (s_last) {
      tickets[r * nkv + kvh] = 0;  // re-arm for the next layer / step
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
  }
  __syncthreads();
  if (!s_last) return;
  // per (head, split) rescale factors exp(m_s - m) and the normaliser 1/l, computed once into shared memory
  constexpr int MAXS = 32;
  __shared__ float s_c[MAXG][MAXS];
  __shared__ float s_invl[MAXG];
  for (int h = warp; h < gq; h += 4) {
    const float* p = part + ((long long)r * nq + kvh * gq + h) * nsplit * (HD + 2);
    const float ms = (lane < nsplit) ? __ldcg(p + lane * (HD + 2) + HD) : -INFINITY;
    const float ls = (lane < nsplit) ? __ldcg(p + lane * (HD + 2) + HD + 1) : 0.f;
    float m = ms;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float c = (ms == -INFINITY) ? 0.f : __expf(ms - m);
    s_c[h][lane] = c;
    const float l = wsum(ls * c);
    if (lane == 0) s_invl[h] = 1.f / l;
  }
  __syncthreads();
  // every thread owns OUTS outputs (element o = threadIdx.x + 128 * oo); the split loop is outermost so that OUTS
  // independent L2 loads are in flight per iteration instead of one serial chain per output.
  constexpr int OUTS = (MAXG * HD + 127) / 128;
  float a[OUTS];
#pragma unroll
  for (int oo = 0; oo < OUTS; ++oo) a[oo] = 0.f;
  const float* pbase = part + ((long long)r * nq + kvh * gq) * nsplit * (HD + 2);
#pragma unroll 2
  for (int s2 = 0; s2 < nsplit; ++s2) {
#pragma unroll
    for (int oo = 0; oo < OUTS; ++oo) {
      const int o = threadIdx.x + 128 * oo;
      if (o < gq * HD) {
        const int h = o / HD, d = o % HD;
        a[oo] += __ldcg(pbase + ((long long)h * nsplit + s2) * (HD + 2) + d) * s_c[h][s2];
      }
    }
  }
#pragma unroll
  for (int oo = 0; oo < OUTS; ++oo) {
    const int o = threadIdx.x + 128 * oo;
    if (o < gq * HD) {
      const int h = o / HD, d = o % HD;
      out[((long long)r * nq + kvh * gq + h) * HD + d] = __float2bfloat16(a[oo] * s_invl[h]);
    }
  }
}


// ------------------------------------------------------------------------------------------------
// Tensor-core fused decode attention (head_dim 64 / 128): the kernel the rollout graph runs.
// Same contract and partial/ticket protocol as decode_attn_fused_kernel; what differs:
//   * the `gq` query heads of the kv head form the M dimension (padded to 16) of mma.sync m16n8k16 tiles, so one K/V
//     row read from shared memory serves all heads at once (the scalar kernel re-read K through LDS per head);
//   * the grid is sized for ONE wave (rows x kv heads x nsplit <= 3 CTAs per SM) and each CTA walks ITS balanced share
//     of the context [sp * per, (sp + 1) * per), per = ceil16(ctx / nsplit), in 64-key chunks through a double-buffered
//     cp.async pipeline, so loads of chunk c + 1 overlap the MMAs of chunk c (the first version staged one fixed
//     128-key chunk per CTA, ran 2+ waves and idled the tensor pipe while loading);
//   * per chunk each warp owns 16 keys:  S[16 x 16] = Q K^T  ->  online softmax on the fragments (quad shuffles)  ->
//     O[16 x HD] += P V, P re-used from the S accumulators as the bf16 A operand.
// Q, K and V tiles sit in shared memory with their 16-byte pieces XOR-swizzled by (row & 7): ldmatrix (plain for Q/K,
// .trans for V) is bank-conflict free.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// Transposes an 8 x 8 matrix of b16 held one row-pair per lane (lane l: row l / 4, columns 2 (l % 4), + 1) across the warp.
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}

template <int HD, int NW, int NS, int TR>
__global__ void __launch_bounds__(NW * 32) decode_attn_mma_kernel(
    const float* __restrict__ qkv, const float* __restrict__ cos_tab, const float* __restrict__ sin_tab,
    const int* __restrict__ rope_delta, const bf16* __restrict__ kp, const bf16* __restrict__ vp, bf16* __restrict__ kc,
    bf16* __restrict__ vc, const int* __restrict__ state, const int* __restrict__ row_group,
    const int* __restrict__ row_plen, const int* __restrict__ finished, float* __restrict__ part, int* __restrict__ tickets,
    bf16* __restrict__ out, int nq, int nkv, int p_max, int c_max, int max_pos, float scale) {
  if (threadIdx.x == 0) pdl_trigger();   // dependents may become resident (and prefetch) right away
  constexpr int HALF = HD / 2, PPR = HD / 8, NT = NW * 32, CK = 16 * NW, MAXG = 8, NKS = HD / 16, NOB = HD / 8, EPL = HD / 32;
  const int r = blockIdx.x, kvh = blockIdx.y, sp = blockIdx.z, nsplit = gridDim.z;
  // a row that has produced EOS only emits padding from now on: no K/V traffic, no append (its hidden state is never used;
  // `finished` was written by the previous step's sampler, which the step's first kernel orders before everything here)
  if (finished != nullptr && finished[r]) return;
  const int gq = nq / nkv;
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int P = row_plen[r];
  const int step = state[ST_STEP];
  const int ctx = P + step + 1;
  const int per = (((ctx + nsplit - 1) / nsplit) + 15) & ~15;
  const int k0 = min(ctx, sp * per), k1 = min(ctx, k0 + per);
  const int nkeys = k1 - k0;
  const int nchunks = (nkeys + CK - 1) / CK;
  const int grp = row_group[r];
  const int qkv_dim = (nq + 2 * nkv) * HD;
  const float* xrow = qkv + (long long)r * qkv_dim;
  int pos = P + step + rope_delta[r];
  pos = max(0, min(max_pos - 1, pos));
  const float* cs = cos_tab + (long long)pos * HD;
  const float* sn = sin_tab + (long long)pos * HD;
  auto rot = [&](const float* x, int d) -> float {
    const float xd = bf16r(x[d]);
    const float xp = bf16r(x[d < HALF ? d + HALF : d - HALF]);
    const float a = bf16r(xd * bf16r(cs[d]));
    const float b = bf16r((d < HALF ? -xp : xp) * bf16r(sn[d]));
    return bf16r(a + b);
  };
  __shared__ __align__(128) bf16 sQ[16 * HD];             // 16 query rows (heads, zero-padded), swizzled pieces
  __shared__ __align__(16) bf16 sNew[2][HD];              // this step's rotated k and v (linear)
  __shared__ int s_last;
  extern __shared__ __align__(128) uint8_t sm_kv_raw[];   // NS stages x {K [CK][HD], V [CK][HD]}, swizzled
  bf16* sKV = reinterpret_cast<bf16*>(sm_kv_raw);
  constexpr int STAGE = 2 * CK * HD;                      // elements per stage

  // cache rows of chunk c -> stage buffer; the token being decoded (key ctx - 1) is patched in from sNew later
  auto stage = [&](int c) {
    bf16* sK = sKV + (c % NS) * STAGE;
    bf16* sV = sK + CK * HD;
    const int base = k0 + c * CK;
    for (int q = threadIdx.x; q < CK * PPR; q += NT) {
      const int jj = q / PPR, piece = q % PPR;
      const int j = base + jj;
      const int off = jj * HD + ((piece ^ (jj & 7)) << 3);
      if (j < k1) {
        if (j != ctx - 1) {
          const long long row = (j < P) ? (((long long)grp * p_max + j) * nkv + kvh) * HD
                                        : (((long long)r * c_max + (j - P)) * nkv + kvh) * HD;
          cp_async16(sK + off, ((j < P) ? kp : kc) + row + piece * 8);
          cp_async16(sV + off, ((j < P) ? vp : vc) + row + piece * 8);
        }
      } else {
        *reinterpret_cast<uint4*>(sK + off) = make_uint4(0, 0, 0, 0);   // beyond the range: finite zeros (masked below)
        *reinterpret_cast<uint4*>(sV + off) = make_uint4(0, 0, 0, 0);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // chunks 0 and 1 hold cache rows written by earlier decode steps / the prefill: requested before the PDL wait
#pragma unroll
  for (int c = 0; c < NS; ++c)
    if (c < nchunks) stage(c);
  pdl_wait();
  // ---- rotated, pre-scaled queries as bf16 rows (row = head within the group) ----
  {
    // thread = one head-dim column d of up to MAXG heads: all global loads are issued before the first use (one L2
    // round trip instead of one per head)
    constexpr int CPT = HD > NT ? HD / NT : 1;  // columns per thread (NT < HD: a thread owns several columns)
    constexpr int TPH = NT > HD ? NT / HD : 1;  // thread groups sharing the column range, each taking every TPH-th head
    constexpr int HPT = MAXG / TPH;             // heads per thread
    const int hsel = TPH > 1 ? threadIdx.x / HD : 0;
#pragma unroll
    for (int cc = 0; cc < CPT; ++cc) {
      const int d = (TPH > 1 ? threadIdx.x % HD : threadIdx.x) + cc * NT;
      const int dp = d < HALF ? d + HALF : d - HALF;
      const float cd = bf16r(cs[d]), sd = bf16r(sn[d]);
      float xv[HPT], xpv[HPT];
#pragma unroll
      for (int hh = 0; hh < HPT; ++hh) {
        const int h = hh * TPH + hsel;
        const float* xh = xrow + (long long)(kvh * gq + min(h, gq - 1)) * HD;
        xv[hh] = xh[d];
        xpv[hh] = xh[dp];
      }
#pragma unroll
      for (int hh = 0; hh < HPT; ++hh) {
        const int h = hh * TPH + hsel;
        const float a = bf16r(bf16r(xv[hh]) * cd);
        const float xp = bf16r(xpv[hh]);
        const float b = bf16r((d < HALF ? -xp : xp) * sd);
        const float v = (h < gq) ? bf16r(a + b) * scale : 0.f;
        sQ[h * HD + ((((d >> 3) ^ (h & 7)) << 3) | (d & 7))] = __float2bfloat16(v);
        sQ[(h + 8) * HD + ((((d >> 3) ^ (h & 7)) << 3) | (d & 7))] = __float2bfloat16(0.f);   // padding rows 8..15
      }
    }
  }
  const bool has_new = (ctx - 1 >= k0) && (ctx - 1 < k1);
  if (has_new && warp == NW - 1) {
    const float* knew = xrow + (long long)(nq + kvh) * HD;
    const float* vnew = xrow + (long long)(nq + nkv + kvh) * HD;
    bf16* kdst = kc + (((long long)r * c_max + step) * nkv + kvh) * HD;
    bf16* vdst = vc + (((long long)r * c_max + step) * nkv + kvh) * HD;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const int d = lane * EPL + e;
      const bf16 kb = __float2bfloat16(rot(knew, d));
      const bf16 vb = __float2bfloat16(vnew[d]);
      sNew[0][d] = kb;
      sNew[1][d] = vb;
      kdst[d] = kb;
      vdst[d] = vb;
    }
  }
  __syncthreads();
  // TR = 1 ("transposed"): the 16 KEYS of a warp are the M rows of the MMA tiles and the (up to 8) query heads the N = 8
  // columns - S^T = K Q^T, O^T = V^T P^T - so no tile row is padding: half the mma.sync instructions of the heads-as-M form
  // (TR = 0: 8 heads padded to m16), which is what bounds this kernel (one m16n8k16 per ~32 cycles per SM sub-core).
  constexpr int NOA = TR ? HD / 16 : HD / 8;
  uint32_t qa[TR ? 1 : NKS][4];
  uint32_t qb[TR ? NKS : 1][2];
  if constexpr (TR) {
#pragma unroll
    for (int ks = 0; ks < NKS; ks += 2) {   // B fragments of Q^T: [n = head][k = d] rows of sQ, two k-steps per ldmatrix.x4
      uint32_t t4[4];
      const int row = lane & 7;
      const int piece = ks * 2 + (lane >> 3);
      ldsm_x4(t4, sQ + row * HD + ((piece ^ (row & 7)) << 3));
      qb[ks][0] = t4[0]; qb[ks][1] = t4[1];
      qb[ks + 1][0] = t4[2]; qb[ks + 1][1] = t4[3];
    }
  } else {
#pragma unroll
    for (int ks = 0; ks < NKS; ++ks) {
      const int row = (lane & 7) + ((lane >> 3) & 1) * 8;
      const int piece = ks * 2 + (lane >> 4);
      ldsm_x4(qa[ks], sQ + row * HD + ((piece ^ (row & 7)) << 3));
    }
  }
  float oacc[NOA][4];
#pragma unroll
  for (int nb = 0; nb < NOA; ++nb)
#pragma unroll
    for (int e = 0; e < 4; ++e) oacc[nb][e] = 0.f;
  float mrun = -INFINITY, lrun = 0.f;      // TR = 0: of query row lane / 4 (rows >= 8 are padding)
  float mrun2[2] = {-INFINITY, -INFINITY}, lrun2[2] = {0.f, 0.f};   // TR = 1: of heads 2 * (lane % 4), + 1

  for (int c = 0; c < nchunks; ++c) {
    bf16* sK = sKV + (c % NS) * STAGE;
    bf16* sV = sK + CK * HD;
    const int cbase = k0 + c * CK;
    if (has_new && ctx - 1 >= cbase && ctx - 1 < cbase + CK && threadIdx.x < 2 * PPR) {
      const int jj = ctx - 1 - cbase;
      const int which = threadIdx.x / PPR, piece = threadIdx.x % PPR;
      bf16* dst = (which ? sV : sK) + jj * HD + ((piece ^ (jj & 7)) << 3);
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(&sNew[which][piece * 8]);
    }
    // chunks c + 1 .. c + NS - 1 may still be in flight
    if (NS == 3 && c + 2 < nchunks) asm volatile("cp.async.wait_group 2;" ::: "memory");
    else if (c + 1 < nchunks) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const int t0 = warp * 16;                   // this warp's 16 keys of the chunk
    const int nk = min(CK, k1 - cbase);         // valid keys in the chunk
    if constexpr (TR) {
    if (t0 < nk) {
      float sacc[4] = {0.f, 0.f, 0.f, 0.f};     // S^T tile: keys t0 + lane / 4 (+ 8) x heads 2 * (lane % 4) (+ 1)
#pragma unroll
      for (int ks = 0; ks < NKS; ++ks) {
        uint32_t a4[4];
        const int key = t0 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int piece = ks * 2 + (lane >> 4);
        ldsm_x4(a4, sK + key * HD + ((piece ^ (key & 7)) << 3));
        mma_bf16_16816(sacc, a4, qb[ks][0], qb[ks][1]);
      }
      if (t0 + (lane >> 2) >= nk) sacc[0] = sacc[1] = -INFINITY;
      if (t0 + (lane >> 2) + 8 >= nk) sacc[2] = sacc[3] = -INFINITY;
      uint32_t pb[2];
      float corr[2], pv[4];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float m = fmaxf(sacc[e], sacc[2 + e]);   // over the 16 keys: lanes with the same lane % 4
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
        const float mnew = fmaxf(mrun2[e], m);  // finite: key t0 is valid
        corr[e] = (mrun2[e] == -INFINITY) ? 0.f : __expf(mrun2[e] - mnew);
        pv[e] = __expf(sacc[e] - mnew);
        pv[2 + e] = __expf(sacc[2 + e] - mnew);
        float l = pv[e] + pv[2 + e];
        l += __shfl_xor_sync(0xffffffffu, l, 4);
        l += __shfl_xor_sync(0xffffffffu, l, 8);
        l += __shfl_xor_sync(0xffffffffu, l, 16);
        lrun2[e] = lrun2[e] * corr[e] + l;
        mrun2[e] = mnew;
      }
      // P^T as the B operand: transpose the two 8 x 8 [key][head] tiles across the warp
      pb[0] = movmatrix_trans(pack_bf16x2(pv[0], pv[1]));
      pb[1] = movmatrix_trans(pack_bf16x2(pv[2], pv[3]));
#pragma unroll
      for (int mt = 0; mt < NOA; ++mt) {
        oacc[mt][0] *= corr[0]; oacc[mt][1] *= corr[1];
        oacc[mt][2] *= corr[0]; oacc[mt][3] *= corr[1];
      }
#pragma unroll
      for (int mt = 0; mt < NOA; ++mt) {        // O^T[16 d x 8 heads] += V^T[16 d x 16 keys] P^T
        uint32_t a4[4];
        const int key = t0 + (lane & 7) + ((lane >> 4) & 1) * 8;
        const int piece = mt * 2 + ((lane >> 3) & 1);
        ldsm_x4_trans(a4, sV + key * HD + ((piece ^ (key & 7)) << 3));
        mma_bf16_16816(oacc[mt], a4, pb[0], pb[1]);
      }
    }
    }
    if constexpr (!TR) {
    if (t0 < nk) {
      float sacc[2][4];
#pragma unroll
      for (int nb = 0; nb < 2; ++nb)
#pragma unroll
        for (int e = 0; e < 4; ++e) sacc[nb][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < NKS; ++ks) {
        uint32_t b[4];
        const int key = t0 + (lane & 7) + ((lane >> 4) & 1) * 8;
        const int piece = ks * 2 + ((lane >> 3) & 1);
        ldsm_x4(b, sK + key * HD + ((piece ^ (key & 7)) << 3));
        mma_bf16_16816(sacc[0], qa[ks], b[0], b[1]);
        mma_bf16_16816(sacc[1], qa[ks], b[2], b[3]);
      }
      float mloc = -INFINITY;
#pragma unroll
      for (int nb = 0; nb < 2; ++nb)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = t0 + nb * 8 + (lane & 3) * 2 + e;
          if (key >= nk) sacc[nb][e] = -INFINITY;
          mloc = fmaxf(mloc, sacc[nb][e]);
        }
      mloc = fmaxf(mloc, __shfl_xor_sync(0xffffffffu, mloc, 1));
      mloc = fmaxf(mloc, __shfl_xor_sync(0xffffffffu, mloc, 2));
      const float mnew = fmaxf(mrun, mloc);     // finite: key t0 is valid
      const float corr = (mrun == -INFINITY) ? 0.f : __expf(mrun - mnew);
      uint32_t pa[4];
      float lloc = 0.f;
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) {
        const float p0 = __expf(sacc[nb][0] - mnew);
        const float p1 = __expf(sacc[nb][1] - mnew);
        lloc += p0 + p1;
        pa[nb * 2 + 0] = pack_bf16x2(p0, p1);
        pa[nb * 2 + 1] = 0u;                    // rows + 8: padding heads
      }
      lloc += __shfl_xor_sync(0xffffffffu, lloc, 1);
      lloc += __shfl_xor_sync(0xffffffffu, lloc, 2);
      lrun = lrun * corr + lloc;
      mrun = mnew;
#pragma unroll
      for (int nb = 0; nb < NOB; ++nb) {
        oacc[nb][0] *= corr;
        oacc[nb][1] *= corr;
      }
#pragma unroll
      for (int nb = 0; nb < NOB; nb += 2) {
        uint32_t b[4];
        const int key = t0 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int piece = nb + (lane >> 4);
        ldsm_x4_trans(b, sV + key * HD + ((piece ^ (key & 7)) << 3));
        mma_bf16_16816(oacc[nb], pa, b[0], b[1]);
        mma_bf16_16816(oacc[nb + 1], pa, b[2], b[3]);
      }
    }
    }
    __syncthreads();                            // every warp is done with this chunk's buffer:
    if (c + NS < nchunks) stage(c + NS);        // re-fill it NS chunks ahead
  }
  // ---- merge the 4 warps (rows = heads < gq live in c0/c1 of the lanes with lane/4 == head); staging memory re-used
  float (*sm_mrg)[MAXG][HD + 2] = reinterpret_cast<float (*)[MAXG][HD + 2]>(sm_kv_raw);   // NW x MAXG x (HD + 2) floats <= staging
  const int row_lo = lane >> 2;
  if constexpr (TR) {
#pragma unroll
    for (int mt = 0; mt < NOA; ++mt)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        sm_mrg[warp][(lane & 3) * 2 + (e & 1)][mt * 16 + row_lo + (e >> 1) * 8] = oacc[mt][e];
    if (row_lo == 0) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        sm_mrg[warp][(lane & 3) * 2 + e][HD] = mrun2[e];
        sm_mrg[warp][(lane & 3) * 2 + e][HD + 1] = lrun2[e];
      }
    }
  } else if (row_lo < MAXG) {
#pragma unroll
    for (int nb = 0; nb < NOA; ++nb) {
      sm_mrg[warp][row_lo][nb * 8 + (lane & 3) * 2 + 0] = oacc[nb][0];
      sm_mrg[warp][row_lo][nb * 8 + (lane & 3) * 2 + 1] = oacc[nb][1];
    }
    if ((lane & 3) == 0) {
      sm_mrg[warp][row_lo][HD] = mrun;
      sm_mrg[warp][row_lo][HD + 1] = lrun;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < gq * HD; i += NT) {
    const int h = i / HD, d = i % HD;
    float m = -INFINITY;
#pragma unroll
    for (int w = 0; w < NW; ++w) m = fmaxf(m, sm_mrg[w][h][HD]);
    float a = 0.f, l = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const float mw = sm_mrg[w][h][HD];
      const float cf = (mw == -INFINITY) ? 0.f : __expf(mw - m);
      a += sm_mrg[w][h][d] * cf;
      l += sm_mrg[w][h][HD + 1] * cf;
    }
    if (nsplit == 1) {      // the whole context in one CTA: no partials, no ticket, no second pass
      out[((long long)r * nq + kvh * gq + h) * HD + d] = __float2bfloat16(a / l);
      continue;
    }
    float* dst = part + (((long long)r * nq + kvh * gq + h) * nsplit + sp) * (HD + 2);
    dst[d] = a;
    if (d == 0) {
      dst[HD] = m;
      dst[HD + 1] = l;
    }
  }
  if (nsplit == 1) return;
  // ---- last CTA of this (row, kv head) merges the splits (same protocol as the scalar kernel) ----
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    const int t = atomicAdd(&tickets[r * nkv + kvh], 1);
    s_last = (t == nsplit - 1);
    if (s_last) {
      tickets[r * nkv + kvh] = 0;
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
  }
  __syncthreads();
  if (!s_last) return;
  constexpr int MAXS = 32;
  __shared__ float s_c[MAXG][MAXS];
  __shared__ float s_invl[MAXG];
  for (int h = warp; h < gq; h += NW) {
    const float* p = part + ((long long)r * nq + kvh * gq + h) * nsplit * (HD + 2);
    const float ms = (lane < nsplit) ? __ldcg(p + lane * (HD + 2) + HD) : -INFINITY;
    const float ls = (lane < nsplit) ? __ldcg(p + lane * (HD + 2) + HD + 1) : 0.f;
    float m = ms;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float cf = (ms == -INFINITY) ? 0.f : __expf(ms - m);
    s_c[h][lane] = cf;
    const float l = wsum(ls * cf);
    if (lane == 0) s_invl[h] = 1.f / l;
  }
  __syncthreads();
  constexpr int OUTS = (MAXG * HD + NT - 1) / NT;
  float a8[OUTS];
#pragma unroll
  for (int oo = 0; oo < OUTS; ++oo) a8[oo] = 0.f;
  const float* pbase = part + ((long long)r * nq + kvh * gq) * nsplit * (HD + 2);
#pragma unroll 2
  for (int s2 = 0; s2 < nsplit; ++s2) {
#pragma unroll
    for (int oo = 0; oo < OUTS; ++oo) {
      const int o = threadIdx.x + NT * oo;
      if (o < gq * HD) {
        const int h = o / HD, d = o % HD;
        a8[oo] += __ldcg(pbase + ((long long)h * nsplit + s2) * (HD + 2) + d) * s_c[h][s2];
      }
    }
  }
#pragma unroll
  for (int oo = 0; oo < OUTS; ++oo) {
    const int o = threadIdx.x + NT * oo;
    if (o < gq * HD) {
      const int h = o / HD, d = o % HD;
      out[((long long)r * nq + kvh * gq + h) * HD + d] = __float2bfloat16(a8[oo] * s_invl[h]);
    }
  }
}


// ------------------------------------------------------------------------------------------------
// Shared-prefix decode attention (head_dim 64 / 128): the G rows of a GRPO group attend to the SAME prompt K/V
// (ref: sc_grpo_trainer.py:351 `enable_prefix_caching`), so the prompt part of the context is processed once per group:
//   * P-kind CTAs (group, 8-row block, kv head, prompt split): the 8 rows x gq query heads form 64 query rows = 4 full
//     m16 MMA tiles (one per warp: 2 decode rows x 8 heads, no padding rows); every warp walks ALL keys of the chunk for
//     its own 16 query rows (no cross-warp merge), K/V tiles are staged once for 8 rows instead of once per row;
//   * C-kind CTAs (row, kv head, completion split): the row's own keys incl. the token being decoded (rotary + KV append),
//     the per-row kernel's structure (warps split the keys of a chunk, merged in shared memory).
// Both kinds write (o, m, l) partials per (row, head, slot); the last CTA to arrive for a (row, kv head) (atomic ticket over
// psplit + csplit slots) merges them. One launch, one wave.
// ------------------------------------------------------------------------------------------------
// Merge of the split partials of up to 8 rows of one kv head by ONE CTA (the last to arrive for those rows). Everything is
// laid out for few L2 round trips: one thread per (row, head) loads all its (m, l) pairs at once and derives the slot
// weights; then every thread owns pairs of adjacent head-dim columns and has the loads of all slots of 4 pairs in flight.
template <int HD>
__device__ __forceinline__ void attn_merge_rows(const float* __restrict__ part, bf16* __restrict__ out, const int* rows, int n_rows,
                                                int kvh, int nq, int gq, int nslots, float* s_coef /* [64][17] */) {
  constexpr int NT = 128, CS = 17;     // slot weights of an item, then 1 / l at [16]
  const int items = n_rows * gq;       // (row, head) pairs, <= 64
  if ((int)threadIdx.x < items) {
    const int ri = threadIdx.x / gq, h = threadIdx.x % gq;
    const float* p = part + ((long long)rows[ri] * nq + kvh * gq + h) * nslots * (HD + 2) + HD;
    float ms[16], ls[16];
#pragma unroll
    for (int s2 = 0; s2 < 16; ++s2) {
      ms[s2] = (s2 < nslots) ? __ldcg(p + s2 * (HD + 2)) : -INFINITY;
      ls[s2] = (s2 < nslots) ? __ldcg(p + s2 * (HD + 2) + 1) : 0.f;
    }
    float m = -INFINITY;
#pragma unroll
    for (int s2 = 0; s2 < 16; ++s2) m = fmaxf(m, ms[s2]);
    float l = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < 16; ++s2) {
      const float cf = (ms[s2] == -INFINITY) ? 0.f : __expf(ms[s2] - m);
      s_coef[threadIdx.x * CS + s2] = cf;
      l += ls[s2] * cf;
    }
    s_coef[threadIdx.x * CS + 16] = 1.f / l;
  }
  __syncthreads();
  constexpr int PAIRS = HD / 2;
  const int total = items * PAIRS;
  for (int base = threadIdx.x; base < total; base += NT * 4) {
    float2 acc[4];
    const float* src[4];
    int item[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int o = min(base + u * NT, total - 1);
      item[u] = o / PAIRS;
      const int ri = item[u] / gq, h = item[u] % gq, d = (o % PAIRS) * 2;
      src[u] = part + ((long long)rows[ri] * nq + kvh * gq + h) * nslots * (HD + 2) + d;
      acc[u] = make_float2(0.f, 0.f);
    }
    for (int s2 = 0; s2 < nslots; s2 += 4) {
      float2 v[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          v[u][k] = (s2 + k < nslots) ? __ldcg(reinterpret_cast<const float2*>(src[u] + (long long)(s2 + k) * (HD + 2))) : make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float cf = (s2 + k < nslots) ? s_coef[item[u] * CS + s2 + k] : 0.f;
          acc[u].x += v[u][k].x * cf;
          acc[u].y += v[u][k].y * cf;
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int o = base + u * NT;
      if (o < total) {
        const int ri = item[u] / gq, h = item[u] % gq, d = (o % PAIRS) * 2;
        const float inv = s_coef[item[u] * CS + 16];
        *reinterpret_cast<__nv_bfloat162*>(out + ((long long)rows[ri] * nq + kvh * gq + h) * HD + d) =
            __floats2bfloat162_rn(acc[u].x * inv, acc[u].y * inv);
      }
    }
  }
}

template <int HD>
__global__ void __launch_bounds__(128, 3) decode_attn_grouped_kernel(
    const float* __restrict__ qkv, const float* __restrict__ cos_tab, const float* __restrict__ sin_tab,
    const int* __restrict__ rope_delta, const bf16* __restrict__ kp, const bf16* __restrict__ vp, bf16* __restrict__ kc,
    bf16* __restrict__ vc, const int* __restrict__ state, const int* __restrict__ row_plen, const int* __restrict__ finished,
    float* __restrict__ part, int* __restrict__ tickets, bf16* __restrict__ out, int G, int nq, int nkv, int p_max, int c_max,
    int max_pos, float scale, int psplit, int csplit, int n_pblocks) {
  if (threadIdx.x == 0) pdl_trigger();   // dependents may become resident (and prefetch) right away
  constexpr int HALF = HD / 2, PPR = HD / 8, NT = 128, NW = 4, CK = 64, MAXG = 8, NKS = HD / 16, NOB = HD / 8, EPL = HD / 32;
  constexpr int STAGE = 2 * CK * HD;                      // bf16 elements per stage: K [CK][HD] + V [CK][HD]
  const int gq = nq / nkv, nslots = psplit + csplit;
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int step = state[ST_STEP];
  const int qkv_dim = (nq + 2 * nkv) * HD;
  __shared__ __align__(128) bf16 sQ[16 * HD];             // C-kind: 16 query rows (heads, zero-padded), swizzled pieces
  __shared__ __align__(16) bf16 sNew[2][HD];              // C-kind: this step's rotated k and v (linear)
  __shared__ int s_rows[8];
  __shared__ int s_nrows;
  extern __shared__ __align__(128) uint8_t sm_kv_raw[];   // 2 stages x {K, V}, swizzled
  bf16* sKV = reinterpret_cast<bf16*>(sm_kv_raw);
  float* s_coef = reinterpret_cast<float*>(sm_kv_raw);    // merge scratch (the staging area is idle by then)

  if ((int)blockIdx.x < n_pblocks) {
    // =========================== P-kind: the prompt keys of one group for a block of up to 8 rows ===========================
    const int rblocks = (G + 7) / 8;
    int t = blockIdx.x;
    const int ps = t % psplit; t /= psplit;
    const int kvh = t % nkv; t /= nkv;
    const int rb = t % rblocks;
    const int grp = t / rblocks;
    const int r_base = grp * G + rb * 8, nrows = min(8, G - rb * 8);
    const int P = row_plen[r_base];
    const int per = (((P + psplit - 1) / psplit) + 15) & ~15;
    const int k0 = min(P, ps * per), k1 = min(P, k0 + per);
    const int nchunks = (k1 - k0 + CK - 1) / CK;
    auto stage = [&](int c) {
      bf16* sK = sKV + (c & 1) * STAGE;
      bf16* sV = sK + CK * HD;
      const int base = k0 + c * CK;
      for (int q = threadIdx.x; q < CK * PPR; q += NT) {
        const int jj = q / PPR, piece = q % PPR;
        const int j = base + jj;
        const int off = jj * HD + ((piece ^ (jj & 7)) << 3);
        if (j < k1) {
          const long long row = (((long long)grp * p_max + j) * nkv + kvh) * HD;
          cp_async16(sK + off, kp + row + piece * 8);
          cp_async16(sV + off, vp + row + piece * 8);
        } else {
          *reinterpret_cast<uint4*>(sK + off) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(sV + off) = make_uint4(0, 0, 0, 0);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (nchunks > 0) stage(0);                            // prompt K/V: written by the prefill, independent of this step
    // rotary factors of the block's rows for this thread's head-dim column (positions are static per step)
    constexpr int TPD = NT / HD;                          // threads per column (1 at hd 128, 2 at hd 64)
    constexpr int QPT = 64 / TPD;                         // query rows per thread
    const int d = threadIdx.x % HD, sub = threadIdx.x / HD;
    const int dp = d < HALF ? d + HALF : d - HALF;
    float cs8[8], sn8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = r_base + min(i, nrows - 1);
      int pos = row_plen[r] + step + rope_delta[r];
      pos = max(0, min(max_pos - 1, pos));
      cs8[i] = bf16r(cos_tab[(long long)pos * HD + d]);
      sn8[i] = bf16r(sin_tab[(long long)pos * HD + d]);
    }
    const bool tr = blockIdx.x == 0 && threadIdx.x == 0;
    if (tr) trace_stamp(400);
    pdl_wait();
    if (tr) trace_stamp(401);
    // raw fp32 queries of the block -> stage 1 (free until chunk 1 is requested): [row i][head h][HD]
    float* raw = reinterpret_cast<float*>(sKV + STAGE);
    {
      constexpr int PCS = HD / 4;                         // 16-byte pieces per head
      for (int q = threadIdx.x; q < 64 * PCS; q += NT) {
        const int qrow = q / PCS, piece = q % PCS;
        const int i = qrow >> 3, h = qrow & 7;
        if (i < nrows && h < gq)
          cp_async16(raw + qrow * HD + piece * 4, qkv + (long long)(r_base + i) * qkv_dim + (long long)(kvh * gq + h) * HD + piece * 4);
        else
          *reinterpret_cast<float4*>(raw + qrow * HD + piece * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    __nv_bfloat16 qv[QPT];
#pragma unroll
    for (int qq = 0; qq < QPT; ++qq) {
      const int qrow = qq * TPD + sub, i = qrow >> 3;
      const float xd = bf16r(raw[qrow * HD + d]);
      const float xp = bf16r(raw[qrow * HD + dp]);
      const float a_ = bf16r(xd * cs8[i]);
      const float b_ = bf16r((d < HALF ? -xp : xp) * sn8[i]);
      qv[qq] = __float2bfloat16(bf16r(a_ + b_) * scale);
    }
    __syncthreads();                                      // every raw value has been read: the tile is rewritten in place
    bf16* sQ64 = reinterpret_cast<bf16*>(raw);            // [64][HD] bf16, 16-byte pieces swizzled by (row & 7)
#pragma unroll
    for (int qq = 0; qq < QPT; ++qq) {
      const int qrow = qq * TPD + sub;
      sQ64[qrow * HD + ((((d >> 3) ^ (qrow & 7)) << 3) | (d & 7))] = qv[qq];
    }
    __syncthreads();
    uint32_t qa[NKS][4];
#pragma unroll
    for (int ks = 0; ks < NKS; ++ks) {
      const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const int piece = ks * 2 + (lane >> 4);
      ldsm_x4(qa[ks], sQ64 + row * HD + ((piece ^ (row & 7)) << 3));
    }
    __syncthreads();                                      // stage 1 is free again
    if (tr) trace_stamp(402);
    if (nchunks > 1) stage(1);
    float oacc[NOB][4];
#pragma unroll
    for (int nb = 0; nb < NOB; ++nb)
#pragma unroll
      for (int e = 0; e < 4; ++e) oacc[nb][e] = 0.f;
    float mA = -INFINITY, mB = -INFINITY, lA = 0.f, lB = 0.f;   // query rows lane / 4 and lane / 4 + 8 of the warp's tile
    for (int c = 0; c < nchunks; ++c) {
      bf16* sK = sKV + (c & 1) * STAGE;
      bf16* sV = sK + CK * HD;
      if (c + 1 < nchunks) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      const int nk = min(CK, k1 - (k0 + c * CK));
      float sacc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) sacc[nt][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < NKS; ++ks) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t b[4];
          const int key = j * 16 + (lane & 7) + ((lane >> 4) & 1) * 8;
          const int piece = ks * 2 + ((lane >> 3) & 1);
          ldsm_x4(b, sK + key * HD + ((piece ^ (key & 7)) << 3));
          mma_bf16_16816(sacc[2 * j], qa[ks], b[0], b[1]);
          mma_bf16_16816(sacc[2 * j + 1], qa[ks], b[2], b[3]);
        }
      }
      float mlA = -INFINITY, mlB = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = nt * 8 + (lane & 3) * 2 + (e & 1);
          if (key >= nk) sacc[nt][e] = -INFINITY;
          if (e < 2) mlA = fmaxf(mlA, sacc[nt][e]);
          else mlB = fmaxf(mlB, sacc[nt][e]);
        }
      mlA = fmaxf(mlA, __shfl_xor_sync(0xffffffffu, mlA, 1));
      mlA = fmaxf(mlA, __shfl_xor_sync(0xffffffffu, mlA, 2));
      mlB = fmaxf(mlB, __shfl_xor_sync(0xffffffffu, mlB, 1));
      mlB = fmaxf(mlB, __shfl_xor_sync(0xffffffffu, mlB, 2));
      const float mnA = fmaxf(mA, mlA), mnB = fmaxf(mB, mlB);       // finite: the chunk holds at least one valid key
      const float cA = (mA == -INFINITY) ? 0.f : __expf(mA - mnA);
      const float cB = (mB == -INFINITY) ? 0.f : __expf(mB - mnB);
      float llA = 0.f, llB = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        sacc[nt][0] = __expf(sacc[nt][0] - mnA);
        sacc[nt][1] = __expf(sacc[nt][1] - mnA);
        sacc[nt][2] = __expf(sacc[nt][2] - mnB);
        sacc[nt][3] = __expf(sacc[nt][3] - mnB);
        llA += sacc[nt][0] + sacc[nt][1];
        llB += sacc[nt][2] + sacc[nt][3];
      }
      llA += __shfl_xor_sync(0xffffffffu, llA, 1);
      llA += __shfl_xor_sync(0xffffffffu, llA, 2);
      llB += __shfl_xor_sync(0xffffffffu, llB, 1);
      llB += __shfl_xor_sync(0xffffffffu, llB, 2);
      lA = lA * cA + llA; lB = lB * cB + llB;
      mA = mnA; mB = mnB;
#pragma unroll
      for (int nb = 0; nb < NOB; ++nb) {
        oacc[nb][0] *= cA; oacc[nb][1] *= cA;
        oacc[nb][2] *= cB; oacc[nb][3] *= cB;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(sacc[2 * j][0], sacc[2 * j][1]);
        pa[1] = pack_bf16x2(sacc[2 * j][2], sacc[2 * j][3]);
        pa[2] = pack_bf16x2(sacc[2 * j + 1][0], sacc[2 * j + 1][1]);
        pa[3] = pack_bf16x2(sacc[2 * j + 1][2], sacc[2 * j + 1][3]);
#pragma unroll
        for (int nb = 0; nb < NOB; nb += 2) {
          uint32_t b[4];
          const int key = j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          const int piece = nb + (lane >> 4);
          ldsm_x4_trans(b, sV + key * HD + ((piece ^ (key & 7)) << 3));
          mma_bf16_16816(oacc[nb], pa, b[0], b[1]);
          mma_bf16_16816(oacc[nb + 1], pa, b[2], b[3]);
        }
      }
      __syncthreads();
      if (c + 2 < nchunks) stage(c + 2);
    }
    if (tr) trace_stamp(403);
    // ---- partials straight from the fragments: rows A / B = decode rows 2 * warp, 2 * warp + 1; head = lane / 4
    {
      const int h = lane >> 2;
#pragma unroll
      for (int ab = 0; ab < 2; ++ab) {
        const int i = 2 * warp + ab;
        if (i < nrows && h < gq && !(finished != nullptr && finished[r_base + i])) {
          float* dst = part + (((long long)(r_base + i) * nq + kvh * gq + h) * nslots + ps) * (HD + 2);
#pragma unroll
          for (int nb = 0; nb < NOB; ++nb)
            *reinterpret_cast<float2*>(dst + nb * 8 + (lane & 3) * 2) = make_float2(oacc[nb][2 * ab], oacc[nb][2 * ab + 1]);
          if ((lane & 3) == 0) {
            dst[HD] = ab ? mB : mA;
            dst[HD + 1] = ab ? lB : lA;
          }
        }
      }
    }
    if (threadIdx.x == 0) s_nrows = 0;
    __syncthreads();
    if ((int)threadIdx.x < nrows && !(finished != nullptr && finished[r_base + threadIdx.x])) {   // finished rows: no slot, no ticket
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      const int tk = atomicAdd(&tickets[(r_base + threadIdx.x) * nkv + kvh], 1);
      if (tk == nslots - 1) {       // last arrival for this row: this CTA merges it
        tickets[(r_base + threadIdx.x) * nkv + kvh] = 0;
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        s_rows[atomicAdd(&s_nrows, 1)] = r_base + threadIdx.x;
      }
    }
    __syncthreads();
    if (tr) trace_stamp(404);
    if (s_nrows > 0) attn_merge_rows<HD>(part, out, s_rows, s_nrows, kvh, nq, gq, nslots, s_coef);
    if (tr) trace_stamp(405 + (s_nrows > 0 ? 1 : 0));
    return;
  }

  // =========================== C-kind: one row's own (completion) keys, incl. the token being decoded ===========================
  int t = blockIdx.x - n_pblocks;
  const int sp = t % csplit; t /= csplit;
  const int kvh = t % nkv;
  const int r = t / nkv;
  if (finished != nullptr && finished[r]) return;         // the row only emits padding from now on
  const int P = row_plen[r];
  const int ctx = step + 1;                               // completion keys 0 .. step (the last one is produced here)
  const int per = (((ctx + csplit - 1) / csplit) + 15) & ~15;
  const int k0 = min(ctx, sp * per), k1 = min(ctx, k0 + per);
  const int nkeys = k1 - k0;
  const int nchunks = (nkeys + CK - 1) / CK;
  const float* xrow = qkv + (long long)r * qkv_dim;
  int pos = P + step + rope_delta[r];
  pos = max(0, min(max_pos - 1, pos));
  const float* cs = cos_tab + (long long)pos * HD;
  const float* sn = sin_tab + (long long)pos * HD;
  auto rot = [&](const float* x, int d) -> float {
    const float xd = bf16r(x[d]);
    const float xp = bf16r(x[d < HALF ? d + HALF : d - HALF]);
    const float a_ = bf16r(xd * bf16r(cs[d]));
    const float b_ = bf16r((d < HALF ? -xp : xp) * bf16r(sn[d]));
    return bf16r(a_ + b_);
  };
  auto stage = [&](int c) {
    bf16* sK = sKV + (c & 1) * STAGE;
    bf16* sV = sK + CK * HD;
    const int base = k0 + c * CK;
    for (int q = threadIdx.x; q < CK * PPR; q += NT) {
      const int jj = q / PPR, piece = q % PPR;
      const int j = base + jj;
      const int off = jj * HD + ((piece ^ (jj & 7)) << 3);
      if (j < k1) {
        if (j != ctx - 1) {
          const long long row = (((long long)r * c_max + j) * nkv + kvh) * HD;
          cp_async16(sK + off, kc + row + piece * 8);
          cp_async16(sV + off, vc + row + piece * 8);
        }
      } else {
        *reinterpret_cast<uint4*>(sK + off) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(sV + off) = make_uint4(0, 0, 0, 0);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (nchunks > 0) stage(0);
  if (nchunks > 1) stage(1);
  const bool trc = (int)blockIdx.x == n_pblocks && threadIdx.x == 0;
  if (trc) trace_stamp(410);
  pdl_wait();
  if (trc) trace_stamp(411);
  {
    constexpr int CPT = HD > NT ? HD / NT : 1;
    constexpr int TPH = NT > HD ? NT / HD : 1;
    constexpr int HPT = MAXG / TPH;
    const int hsel = TPH > 1 ? threadIdx.x / HD : 0;
#pragma unroll
    for (int cc = 0; cc < CPT; ++cc) {
      const int d = (TPH > 1 ? threadIdx.x % HD : threadIdx.x) + cc * NT;
      const int dp = d < HALF ? d + HALF : d - HALF;
      const float cd = bf16r(cs[d]), sd = bf16r(sn[d]);
      float xv[HPT], xpv[HPT];
#pragma unroll
      for (int hh = 0; hh < HPT; ++hh) {
        const int h = hh * TPH + hsel;
        const float* xh = xrow + (long long)(kvh * gq + min(h, gq - 1)) * HD;
        xv[hh] = xh[d];
        xpv[hh] = xh[dp];
      }
#pragma unroll
      for (int hh = 0; hh < HPT; ++hh) {
        const int h = hh * TPH + hsel;
        const float a_ = bf16r(bf16r(xv[hh]) * cd);
        const float xp = bf16r(xpv[hh]);
        const float b_ = bf16r((d < HALF ? -xp : xp) * sd);
        const float v = (h < gq) ? bf16r(a_ + b_) * scale : 0.f;
        sQ[h * HD + ((((d >> 3) ^ (h & 7)) << 3) | (d & 7))] = __float2bfloat16(v);
        sQ[(h + 8) * HD + ((((d >> 3) ^ (h & 7)) << 3) | (d & 7))] = __float2bfloat16(0.f);
      }
    }
  }
  const bool has_new = (ctx - 1 >= k0) && (ctx - 1 < k1);
  if (has_new && warp == NW - 1) {
    const float* knew = xrow + (long long)(nq + kvh) * HD;
    const float* vnew = xrow + (long long)(nq + nkv + kvh) * HD;
    bf16* kdst = kc + (((long long)r * c_max + step) * nkv + kvh) * HD;
    bf16* vdst = vc + (((long long)r * c_max + step) * nkv + kvh) * HD;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const int d = lane * EPL + e;
      const bf16 kb = __float2bfloat16(rot(knew, d));
      const bf16 vb = __float2bfloat16(vnew[d]);
      sNew[0][d] = kb;
      sNew[1][d] = vb;
      kdst[d] = kb;
      vdst[d] = vb;
    }
  }
  __syncthreads();
  uint32_t qa[NKS][4];
#pragma unroll
  for (int ks = 0; ks < NKS; ++ks) {
    const int row = (lane & 7) + ((lane >> 3) & 1) * 8;
    const int piece = ks * 2 + (lane >> 4);
    ldsm_x4(qa[ks], sQ + row * HD + ((piece ^ (row & 7)) << 3));
  }
  float oacc[NOB][4];
#pragma unroll
  for (int nb = 0; nb < NOB; ++nb)
#pragma unroll
    for (int e = 0; e < 4; ++e) oacc[nb][e] = 0.f;
  float mrun = -INFINITY, lrun = 0.f;
  for (int c = 0; c < nchunks; ++c) {
    bf16* sK = sKV + (c & 1) * STAGE;
    bf16* sV = sK + CK * HD;
    const int cbase = k0 + c * CK;
    if (has_new && ctx - 1 >= cbase && ctx - 1 < cbase + CK && threadIdx.x < 2 * PPR) {
      const int jj = ctx - 1 - cbase;
      const int which = threadIdx.x / PPR, piece = threadIdx.x % PPR;
      bf16* dst = (which ? sV : sK) + jj * HD + ((piece ^ (jj & 7)) << 3);
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(&sNew[which][piece * 8]);
    }
    if (c + 1 < nchunks) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const int t0 = warp * 16;
    const int nk = min(CK, k1 - cbase);
    if (t0 < nk) {
      float sacc[2][4];
#pragma unroll
      for (int nb = 0; nb < 2; ++nb)
#pragma unroll
        for (int e = 0; e < 4; ++e) sacc[nb][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < NKS; ++ks) {
        uint32_t b[4];
        const int key = t0 + (lane & 7) + ((lane >> 4) & 1) * 8;
        const int piece = ks * 2 + ((lane >> 3) & 1);
        ldsm_x4(b, sK + key * HD + ((piece ^ (key & 7)) << 3));
        mma_bf16_16816(sacc[0], qa[ks], b[0], b[1]);
        mma_bf16_16816(sacc[1], qa[ks], b[2], b[3]);
      }
      float mloc = -INFINITY;
#pragma unroll
      for (int nb = 0; nb < 2; ++nb)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = t0 + nb * 8 + (lane & 3) * 2 + e;
          if (key >= nk) sacc[nb][e] = -INFINITY;
          mloc = fmaxf(mloc, sacc[nb][e]);
        }
      mloc = fmaxf(mloc, __shfl_xor_sync(0xffffffffu, mloc, 1));
      mloc = fmaxf(mloc, __shfl_xor_sync(0xffffffffu, mloc, 2));
      const float mnew = fmaxf(mrun, mloc);
      const float corr = (mrun == -INFINITY) ? 0.f : __expf(mrun - mnew);
      uint32_t pa[4];
      float lloc = 0.f;
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) {
        const float p0 = __expf(sacc[nb][0] - mnew);
        const float p1 = __expf(sacc[nb][1] - mnew);
        lloc += p0 + p1;
        pa[nb * 2 + 0] = pack_bf16x2(p0, p1);
        pa[nb * 2 + 1] = 0u;
      }
      lloc += __shfl_xor_sync(0xffffffffu, lloc, 1);
      lloc += __shfl_xor_sync(0xffffffffu, lloc, 2);
      lrun = lrun * corr + lloc;
      mrun = mnew;
#pragma unroll
      for (int nb = 0; nb < NOB; ++nb) {
        oacc[nb][0] *= corr;
        oacc[nb][1] *= corr;
      }
#pragma unroll
      for (int nb = 0; nb < NOB; nb += 2) {
        uint32_t b[4];
        const int key = t0 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int piece = nb + (lane >> 4);
        ldsm_x4_trans(b, sV + key * HD + ((piece ^ (key & 7)) << 3));
        mma_bf16_16816(oacc[nb], pa, b[0], b[1]);
        mma_bf16_16816(oacc[nb + 1], pa, b[2], b[3]);
      }
    }
    __syncthreads();
    if (c + 2 < nchunks) stage(c + 2);
  }
  if (trc) trace_stamp(412);
  float (*sm_mrg)[MAXG][HD + 2] = reinterpret_cast<float (*)[MAXG][HD + 2]>(sm_kv_raw);
  const int row_lo = lane >> 2;
  if (row_lo < MAXG) {
#pragma unroll
    for (int nb = 0; nb < NOB; ++nb) {
      sm_mrg[warp][row_lo][nb * 8 + (lane & 3) * 2 + 0] = oacc[nb][0];
      sm_mrg[warp][row_lo][nb * 8 + (lane & 3) * 2 + 1] = oacc[nb][1];
    }
    if ((lane & 3) == 0) {
      sm_mrg[warp][row_lo][HD] = mrun;
      sm_mrg[warp][row_lo][HD + 1] = lrun;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < gq * HD; i += NT) {
    const int h = i / HD, d = i % HD;
    float m = -INFINITY;
#pragma unroll
    for (int w = 0; w < NW; ++w) m = fmaxf(m, sm_mrg[w][h][HD]);
    float a_ = 0.f, l = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const float mw = sm_mrg[w][h][HD];
      const float cf = (mw == -INFINITY) ? 0.f : __expf(mw - m);
      a_ += sm_mrg[w][h][d] * cf;
      l += sm_mrg[w][h][HD + 1] * cf;
    }
    float* dst = part + (((long long)r * nq + kvh * gq + h) * nslots + psplit + sp) * (HD + 2);
    dst[d] = a_;
    if (d == 0) {
      dst[HD] = m;
      dst[HD + 1] = l;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    const int tk = atomicAdd(&tickets[r * nkv + kvh], 1);
    s_nrows = (tk == nslots - 1);
    s_rows[0] = r;
    if (s_nrows) {
      tickets[r * nkv + kvh] = 0;
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
  }
  __syncthreads();
  if (trc) trace_stamp(413);
  if (s_nrows > 0) attn_merge_rows<HD>(part, out, s_rows, 1, kvh, nq, gq, nslots, s_coef);
  if (trc) trace_stamp(414 + (s_nrows > 0 ? 1 : 0));
}


// ------------------------------------------------------------------------------------------------
// Fused sampler: one CTA per row over fp32 logits[V].
// Contract (sc_grpo_trainer.py:353-358 SamplingParams; same order as HF generate, generation/utils.py:1214-1223,
// logits_process.py:521-528 top-k "keep everything >= the k-th value", :581-586 top-p on the renormalised survivors):
//   logits / temperature -> keep the top_k largest (ties kept) -> softmax -> keep tokens whose strictly-higher-ranked
//   mass is < top_p (at least one) -> renormalise -> multinomial via Philox(seed, subsequence=row, offset=step).
// The k-th value is found with a 3-pass radix select (11+11+10 bits of the order-preserving uint key) in shared
// memory histograms, so the vocabulary is scanned 4 times from L2 and never sorted.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
constexpr int kMaxKeep = 256;  // survivors buffer (top_k <= 128 plus ties)

constexpr int kMaxCand = 2048;  // candidate pool for the top-k selection

// Applies `f(x, index)` to every logit of the row with 16-byte loads, four requests in flight per thread (the scan is
// L2-latency-bound, not bandwidth-bound).
template <typename F>
__device__ __forceinline__ void scan_row(const float* __restrict__ lg, int V, int eos_id, int eos2, int forbid_eos, F f) {
  const int V4 = ((reinterpret_cast<uintptr_t>(lg) & 15) == 0) ? (V >> 2) : 0;
  const float4* lg4 = reinterpret_cast<const float4*>(lg);
  for (int i0 = threadIdx.x; i0 < V4; i0 += 4 * blockDim.x) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < V4) v[u] = __ldcg(lg4 + i);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < V4) {
        const float xs[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int idx = 4 * i + e;
          f((forbid_eos && (idx == eos_id || idx == eos2)) ? -INFINITY : xs[e], idx);
        }
      }
    }
  }
  for (int i = 4 * V4 + threadIdx.x; i < V; i += blockDim.x) f((forbid_eos && (i == eos_id || i == eos2)) ? -INFINITY : lg[i], i);
}

// Top-k threshold search without sorting or histograms: (1) row max; (2) how many logits lie within delta_i of the max,
// for 8 nested deltas, counted in registers and block-reduced; the smallest delta holding >= k (and <= kMaxCand)
// logits defines a candidate pool; (3) the pool is gathered into shared memory and the exact k-th value is found by
// rank counting. Pathological rows (more than kMaxCand logits inside every usable delta, e.g. constant logits) fall
// back to a bisection on the threshold.
__global__ void __launch_bounds__(1024) sample_kernel(const float* __restrict__ logits, int V, float inv_temp, int top_k,
                                                      float top_p, unsigned long long seed, int* __restrict__ state,
                                                      int* __restrict__ tok, int* __restrict__ finished,
                                                      int* __restrict__ out_tokens, int c_max, int eos_id, int pad_id,
                                                      int forbid_eos, int first) {
  if (threadIdx.x == 0) pdl_trigger();   // dependents may become resident (and prefetch) right away
  __shared__ float s_red[32];
  __shared__ int s_cnt[8][32];
  __shared__ int s_count, s_nsurv;
  __shared__ float s_cval[kMaxCand];
  __shared__ int s_cidx[kMaxCand];
  __shared__ float s_val[kMaxKeep], s_sorted[kMaxKeep];
  __shared__ int s_idx[kMaxKeep], s_sidx[kMaxKeep];
  pdl_wait();
  const int r = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float* lg = logits + (long long)r * V;
  // `first`: logits come from the prefill (predict completion token 0); else they predict token step + 1.
  const int out_pos = first ? 0 : state[ST_STEP] + 1;
  const int eos2 = state[ST_EOS2] - 1;
  if (out_pos >= c_max) return;
  if (finished[r]) {
    if (threadIdx.x == 0) {
      out_tokens[(long long)r * c_max + out_pos] = pad_id;
      tok[r] = pad_id;
    }
    return;
  }
  const int k = min(max(top_k, 1), 128);
  // ---- (1) row max ----
  float mx = -INFINITY;
  scan_row(lg, V, eos_id, eos2, forbid_eos, [&](float x, int) { mx = fmaxf(mx, x); });
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = -INFINITY;
  for (int i = 0; i < nwarps; ++i) mx = fmaxf(mx, s_red[i]);
  // ---- (2) nested counts; up to three rounds of widening deltas, then bisection ----
  float thr = -INFINITY;
  bool found = false;
  float lo_thr = -INFINITY, hi_thr = mx;  // count(x >= lo_thr) >= k always; hi side may hold < k
  for (int round = 0; round < 3 && !found; ++round) {
    const float base = round == 0 ? 1.f : (round == 1 ? 16.f : 256.f);
    const float dl[8] = {base, 2 * base, 3 * base, 4 * base, 6 * base, 8 * base, 12 * base, 16 * base};
    int c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    scan_row(lg, V, eos_id, eos2, forbid_eos, [&](float x, int) {
      const float d = mx - x;
#pragma unroll
      for (int t = 0; t < 8; ++t) c[t] += (d <= dl[t]) ? 1 : 0;
    });
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      int v = c[t];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_cnt[t][warp] = v;
    }
    __syncthreads();
    int tot[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      tot[t] = 0;
      for (int i = 0; i < nwarps; ++i) tot[t] += s_cnt[t][i];
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if (!found) {
        if (tot[t] >= k && tot[t] <= kMaxCand) {
          thr = mx - dl[t];
          found = true;
        } else if (tot[t] < k) {
          hi_thr = mx - dl[t];       // still too few: the k-th value is below this
        } else if (tot[t] > kMaxCand) {
          lo_thr = mx - dl[t];       // too many inside this delta: bisect between hi_thr and lo_thr
          round = 3;
          break;
        }
      }
    }
  }
  if (!found) {
    // bisection on the threshold (flat or extremely wide rows); 40 halvings reach fp32 resolution
    if (lo_thr == -INFINITY) lo_thr = mx - 1e30f;
    for (int it = 0; it < 40 && !found; ++it) {
      const float mid = 0.5f * (lo_thr + hi_thr);
      int c0 = 0;
      scan_row(lg, V, eos_id, eos2, forbid_eos, [&](float x, int) { c0 += (x >= mid) ? 1 : 0; });
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c0 += __shfl_xor_sync(0xffffffffu, c0, o);
      if (lane == 0) s_cnt[0][warp] = c0;
      __syncthreads();
      int tot0 = 0;
      for (int i = 0; i < nwarps; ++i) tot0 += s_cnt[0][i];
      __syncthreads();
      if (tot0 >= k && tot0 <= kMaxCand) {
        thr = mid;
        found = true;
      } else if (tot0 < k) {
        hi_thr = mid;
      } else {
        lo_thr = mid;
      }
    }
    if (!found) thr = hi_thr;  // >kMaxCand exact ties at the k-th value: keep the first kMaxCand of them
  }
  // ---- (3) gather the pool, exact k-th value by rank counting (ties kept) ----
  if (threadIdx.x == 0) { s_count = 0; s_nsurv = 0; }
  __syncthreads();
  scan_row(lg, V, eos_id, eos2, forbid_eos, [&](float x, int i) {
    if (x >= thr && x > -INFINITY) {
      const int slot = atomicAdd(&s_count, 1);
      if (slot < kMaxCand) {
        s_cval[slot] = x;
        s_cidx[slot] = i;
      }
    }
  });
  __syncthreads();
  const int nc = min(s_count, kMaxCand);
  // Refine the threshold INSIDE the pool (shared memory only): bisection with __syncthreads_count until the pool above
  // the threshold holds between k and 2k values, so the exact rank counting below is O((2k)^2), not O(pool^2).
  float rlo = thr, rhi = mx;
  {
    const float e0 = ((int)threadIdx.x < nc) ? s_cval[threadIdx.x] : -INFINITY;
    const float e1 = ((int)threadIdx.x + 1024 < nc) ? s_cval[threadIdx.x + 1024] : -INFINITY;
    for (int it = 0; it < 24; ++it) {
      const float mid = 0.5f * (rlo + rhi);
      const int cnt = __syncthreads_count(e0 >= mid) + __syncthreads_count(e1 >= mid);
      if (cnt >= k) {
        rlo = mid;
        if (cnt <= 2 * k) break;
      } else {
        rhi = mid;
      }
    }
  }
  // compact the refined pool (values >= rlo: between k and ~2k of them, more only under exact ties) ...
  __shared__ int s_m;
  __shared__ float s_pv[kMaxKeep * 2];
  __shared__ int s_pi[kMaxKeep * 2];
  if (threadIdx.x == 0) s_m = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < nc; i += blockDim.x) {
    const float v = s_cval[i];
    if (v >= rlo) {
      const int slot = atomicAdd(&s_m, 1);
      if (slot < kMaxKeep * 2) {
        s_pv[slot] = v;
        s_pi[slot] = s_cidx[i];
      }
    }
  }
  __syncthreads();
  const int m = min(s_m, kMaxKeep * 2);
  // ... and rank-count inside it: anything greater than a pool member is itself in the pool, so ranks are global
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const float v = s_pv[i];
    int greater = 0;
    for (int j = 0; j < m; ++j) greater += (s_pv[j] > v) ? 1 : 0;
    if (greater < k) {  // value >= k-th largest: survives (HF TopKLogitsWarper keeps ties)
      const int slot = atomicAdd(&s_nsurv, 1);
      if (slot < kMaxKeep) {
        s_val[slot] = v * inv_temp;
        s_idx[slot] = s_pi[i];
      }
    }
  }
  __syncthreads();
  const int n = min(s_nsurv, kMaxKeep);
  // parallel rank sort: descending by value, ties by smaller token id (deterministic)
  if ((int)threadIdx.x < n) {
    const float v = s_val[threadIdx.x];
    const int id = s_idx[threadIdx.x];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const float vj = s_val[j];
      rank += (vj > v) || (vj == v && s_idx[j] < id);
    }
    s_sorted[rank] = v;
    s_sidx[rank] = id;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float m0 = s_sorted[0];
    float tot = 0.f;
    for (int i = 0; i < n; ++i) {
      s_sorted[i] = expf(s_sorted[i] - m0);
      tot += s_sorted[i];
    }
    // nucleus: keep token i while the mass of strictly higher-ranked tokens is < top_p
    float before = 0.f, kept = 0.f;
    int nkeep = 0;
    for (int i = 0; i < n; ++i) {
      if (i > 0 && before / tot >= top_p) break;
      kept += s_sorted[i];
      before += s_sorted[i];
      ++nkeep;
    }
    curandStatePhilox4_32_10_t rng;
    const unsigned long long dev_seed = (unsigned long long)(unsigned int)state[ST_SEED_LO] |
                                        ((unsigned long long)(unsigned int)state[ST_SEED_HI] << 32);
    curand_init(seed ^ dev_seed, (unsigned long long)r, (unsigned long long)out_pos, &rng);
    const float u = curand_uniform(&rng) * kept;  // (0, kept]
    float c = 0.f;
    int pick = nkeep - 1;
    for (int i = 0; i < nkeep; ++i) {
      c += s_sorted[i];
      if (u <= c) {
        pick = i;
        break;
      }
    }
    const int t = s_sidx[pick];
    out_tokens[(long long)r * c_max + out_pos] = t;
    tok[r] = t;
    if (t == eos_id || t == eos2) {
      finished[r] = 1;
      atomicSub(&state[ST_UNFINISHED], 1);
    }
  }
}

// act[r][c] = bf16(silu(gate)) * up from the fp32 gate|up accumulator of the stream-K product (values rounded to bf16
// first: the training forward stores that product in bf16), and the accumulator is cleared for the next layer.
__global__ void decode_silu_mul_f32_kernel(float* __restrict__ gu, bf16* __restrict__ out, int rows, int cols) {
  if (threadIdx.x == 0) pdl_trigger();   // dependents may become resident (and prefetch) right away
  pdl_wait();
  const int nvec = cols >> 2;
  const int total = rows * nvec;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / nvec, v = i - r * nvec;
    float4* gp = reinterpret_cast<float4*>(gu + (long long)r * 2 * cols) + v;
    float4* up = reinterpret_cast<float4*>(gu + (long long)r * 2 * cols + cols) + v;
    const float4 g = *gp, u = *up;
    *gp = make_float4(0.f, 0.f, 0.f, 0.f);
    *up = make_float4(0.f, 0.f, 0.f, 0.f);
    auto f = [](float gg, float uu) {
      gg = bf16r(gg);
      return bf16r(silu_f(gg)) * bf16r(uu);
    };
    __nv_bfloat162 o0 = __floats2bfloat162_rn(f(g.x, u.x), f(g.y, u.y));
    __nv_bfloat162 o1 = __floats2bfloat162_rn(f(g.z, u.z), f(g.w, u.w));
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&o0);
    o.y = *reinterpret_cast<uint32_t*>(&o1);
    *reinterpret_cast<uint2*>(out + (long long)r * cols + v * 4) = o;
  }
}

__global__ void decode_advance_kernel(int* state) {
  if (threadIdx.x == 0) pdl_trigger();   // dependents may become resident (and prefetch) right away
  pdl_wait();
  state[ST_STEP] += 1;
}

void trace_install_decode(unsigned long long* p) { trace_install_tu(p); }
}  // namespace iadr1

using namespace iadr1;

extern "C" {

int iadr1_decode_embed(const void* embed, const int* tok, float* h, int rows, int H, void* stream) {
  if (rows <= 0) return 0;
  // First kernel of a decode step: launched WITHOUT the programmatic-serialization attribute, i.e. it starts only after
  // the previous step (sampler, step counter) has fully completed. Every later kernel of the step may become resident
  // early and read the step counter before its own dependency wait; this launch is what makes that safe.
  decode_embed_kernel<<<dim3(rows), dim3(256), 0, (cudaStream_t)stream>>>((const bf16*)embed, tok, h, H);
  IADR1_CHECK_LAUNCH("decode_embed");
  return 0;
}

int iadr1_rmsnorm_f32in(const float* x, const void* w, void* y, int rows, int cols, float eps, float* zero_buf,
                        int zero_per_row, void* stream) {
  if (rows <= 0) return 0;
  launch_kernel(rmsnorm_f32in_kernel, dim3(rows), dim3(256), 0, (cudaStream_t)stream, x, (const bf16*)w, (bf16*)y, cols,
                eps, zero_buf, zero_buf ? zero_per_row : 0);
  IADR1_CHECK_LAUNCH("rmsnorm_f32in");
  return 0;
}

int iadr1_decode_attention_fused(const float* qkv, const float* cos_tab, const float* sin_tab, const int* rope_delta,
                                 const void* kp, const void* vp, void* kc, void* vc, const int* state,
                                 const int* row_group, const int* row_plen, const int* finished, float* part, int* tickets,
                                 void* out, int rows, int nq, int nkv, int hd, int p_max, int c_max, int nsplit, int max_pos,
                                 float scale, void* stream) {
  if (rows <= 0) return 0;
  if (nq % nkv || nq / nkv > 8) return set_error("decode_attention_fused: group size %d unsupported (max 8)", nq / nkv);
  if (nsplit == 0 || nsplit > 32 || nsplit < -32) return set_error("decode_attention_fused: nsplit %d out of range", nsplit);
  cudaStream_t st = (cudaStream_t)stream;
  static const int dbg = getenv("IADR1_ATTN_DEBUG") ? atoi(getenv("IADR1_ATTN_DEBUG")) : 0;  // phase-skipping, probes only
  static const bool use_mma = !(getenv("IADR1_DECODE_ATTN") && std::string(getenv("IADR1_DECODE_ATTN")) == "scalar");
  if ((hd == 128 || hd == 64) && use_mma) {
    // tensor-core path: any nsplit; every CTA walks its balanced share of the live context in 16 * NW-key chunks.
    // nsplit < 0 selects the 2-warp variant (32-key chunks, half the shared memory: twice the CTAs per SM).
    const int nw = nsplit < 0 ? 2 : 4;
    nsplit = nsplit < 0 ? -nsplit : nsplit;
    if (nsplit > 32) return set_error("decode_attention_fused: nsplit %d out of range (1..32)", nsplit);
    // three staging buffers (two chunks in flight behind the one being consumed) whenever the grid still fits one wave at
    // the lower residency the extra shared memory allows; IADR1_DECODE_ATTN_STAGES=2 forces the two-buffer form
    static const int forced = getenv("IADR1_DECODE_ATTN_STAGES") ? atoi(getenv("IADR1_DECODE_ATTN_STAGES")) : 0;
    const size_t stage_b = (size_t)2 * 16 * nw * hd * 2;
    const int per_sm3 = (int)((227 * 1024) / (3 * stage_b + 6 * 1024));
    const bool three = nw == 4 && forced != 2 && per_sm3 >= 1 && (long long)rows * nkv * nsplit <= (long long)per_sm3 * 148;
    const size_t smem = (three ? 3 : 2) * stage_b;
    static const bool transposed = !(getenv("IADR1_DECODE_ATTN_TR") && atoi(getenv("IADR1_DECODE_ATTN_TR")) == 0);
#define IADR1_DECODE_MMA(HD, NW, NS, TR)                                                                             \
  do {                                                                                                               \
    static bool attr = false;                                                                                        \
    if (!attr) {                                                                                                     \
      cudaFuncSetAttribute(decode_attn_mma_kernel<HD, NW, NS, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                           (int)(NS * stage_b));                                                                     \
      attr = true;                                                                                                   \
    }                                                                                                                \
    launch_kernel(decode_attn_mma_kernel<HD, NW, NS, TR>, dim3(rows, nkv, nsplit), dim3(NW * 32), smem, st, qkv, cos_tab, \
                  sin_tab, rope_delta, (const bf16*)kp, (const bf16*)vp, (bf16*)kc, (bf16*)vc, state, row_group,     \
                  row_plen, finished, part, tickets, (bf16*)out, nq, nkv, p_max, c_max, max_pos, scale);             \
  } while (0)
    if (hd == 128 && nw == 4 && three && transposed) IADR1_DECODE_MMA(128, 4, 3, 1);
    else if (hd == 128 && nw == 4 && three) IADR1_DECODE_MMA(128, 4, 3, 0);
    else if (hd == 128 && nw == 4 && transposed) IADR1_DECODE_MMA(128, 4, 2, 1);
    else if (hd == 128 && nw == 4) IADR1_DECODE_MMA(128, 4, 2, 0);
    else if (hd == 128) IADR1_DECODE_MMA(128, 2, 2, 0);
    else if (nw == 4 && three && transposed) IADR1_DECODE_MMA(64, 4, 3, 1);
    else if (nw == 4 && three) IADR1_DECODE_MMA(64, 4, 3, 0);
    else if (nw == 4 && transposed) IADR1_DECODE_MMA(64, 4, 2, 1);
    else if (nw == 4) IADR1_DECODE_MMA(64, 4, 2, 0);
    else IADR1_DECODE_MMA(64, 2, 2, 0);
#undef IADR1_DECODE_MMA
    IADR1_CHECK_LAUNCH("decode_attention_mma");
    return 0;
  }
  const int chunk = (p_max + c_max + nsplit - 1) / nsplit;
  const size_t kv_smem = (size_t)chunk * hd * 4;
  if (kv_smem > 160 * 1024) return set_error("decode_attention_fused: context too long for one staging buffer (chunk %d)", chunk);
#define IADR1_DECODE_FUSED(HD)                                                                                       \
  do {                                                                                                               \
    if (kv_smem > 48 * 1024)                                                                                         \
      cudaFuncSetAttribute(decode_attn_fused_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);   \
    launch_kernel(decode_attn_fused_kernel<HD>, dim3(rows, nkv, nsplit), dim3(128), kv_smem, st, qkv, cos_tab,       \
                  sin_tab, rope_delta, (const bf16*)kp, (const bf16*)vp, (bf16*)kc, (bf16*)vc, state, row_group,     \
                  row_plen, part, tickets, (bf16*)out, nq, nkv, p_max, c_max, chunk, max_pos, scale, dbg);           \
  } while (0)
  if (hd == 128) IADR1_DECODE_FUSED(128);
  else if (hd == 64) IADR1_DECODE_FUSED(64);
  else if (hd == 32) IADR1_DECODE_FUSED(32);
  else return set_error("decode_attention_fused: head_dim %d unsupported (32, 64 or 128)", hd);
#undef IADR1_DECODE_FUSED
  IADR1_CHECK_LAUNCH("decode_attention_fused");
  return 0;
}

int iadr1_decode_attention_grouped(const float* qkv, const float* cos_tab, const float* sin_tab, const int* rope_delta,
                                   const void* kp, const void* vp, void* kc, void* vc, const int* state, const int* row_plen,
                                   const int* finished, float* part, int* tickets, void* out, int rows, int rows_per_group, int nq,
                                   int nkv, int hd, int p_max, int c_max, int psplit, int csplit, int max_pos, float scale,
                                   void* stream) {
  if (rows <= 0) return 0;
  if (nq % nkv || nq / nkv > 8) return set_error("decode_attention_grouped: group size %d unsupported (max 8)", nq / nkv);
  if (hd != 64 && hd != 128) return set_error("decode_attention_grouped: head_dim %d unsupported (64 or 128)", hd);
  if (rows_per_group <= 0 || rows % rows_per_group) return set_error("decode_attention_grouped: rows must be a multiple of the group size");
  if (psplit < 1 || csplit < 1 || psplit + csplit > 16) return set_error("decode_attention_grouped: bad splits %d + %d (at most 16 slots)", psplit, csplit);
  const int n_groups = rows / rows_per_group, rblocks = (rows_per_group + 7) / 8;
  const int n_pblocks = n_groups * rblocks * nkv * psplit;
  const int grid = n_pblocks + rows * nkv * csplit;
  const size_t smem = (size_t)2 * 2 * 64 * hd * 2;
  cudaStream_t st = (cudaStream_t)stream;
#define IADR1_DECODE_GROUPED(HD)                                                                                     \
  do {                                                                                                               \
    static bool attr = false;                                                                                        \
    if (!attr) {                                                                                                     \
      cudaFuncSetAttribute(decode_attn_grouped_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
      attr = true;                                                                                                   \
    }                                                                                                                \
    launch_kernel(decode_attn_grouped_kernel<HD>, dim3(grid), dim3(128), smem, st, qkv, cos_tab, sin_tab, rope_delta, \
                  (const bf16*)kp, (const bf16*)vp, (bf16*)kc, (bf16*)vc, state, row_plen, finished, part, tickets,    \
                  (bf16*)out, rows_per_group, nq, nkv, p_max, c_max, max_pos, scale, psplit, csplit, n_pblocks);       \
  } while (0)
  if (hd == 128) IADR1_DECODE_GROUPED(128);
  else IADR1_DECODE_GROUPED(64);
#undef IADR1_DECODE_GROUPED
  IADR1_CHECK_LAUNCH("decode_attention_grouped");
  return 0;
}

int iadr1_sample(
const float* logits, int rows, int V, float temperature, int top_k, float top_p,
                 unsigned long long seed, int* state, int* tok, int* finished, int* out_tokens, int c_max, int eos_id,
                 int pad_id, int forbid_eos, int first, void* stream) {
  if (rows <= 0) return 0;
  if (temperature <= 0.f) return set_error("sample: temperature must be > 0");
  launch_kernel(sample_kernel, dim3(rows), dim3(1024), 0, (cudaStream_t)stream, logits, V, 1.f / temperature, top_k,
                top_p, seed, state, tok, finished, out_tokens, c_max, eos_id, pad_id, forbid_eos, first);
  IADR1_CHECK_LAUNCH("sample");
  return 0;
}

int iadr1_decode_silu_mul_f32(float* gu, void* out, int rows, int cols, void* stream) {
  if (rows <= 0) return 0;
  if (cols % 4) return set_error("decode_silu_mul_f32: cols must be a multiple of 4");
  const int total = rows * (cols / 4);
  const int grid = (total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8;
  launch_kernel(decode_silu_mul_f32_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, gu, (bf16*)out, rows, cols);
  IADR1_CHECK_LAUNCH("decode_silu_mul_f32");
  return 0;
}

int iadr1_decode_advance(int* state, void* stream) {
  launch_kernel(decode_advance_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, state);
  IADR1_CHECK_LAUNCH("decode_advance");
  return 0;
}

}  // extern "C"
