// The tcgen05 GEMM family for sm_100a:  C[z] (+)= alpha * A[z] * B[z]^T  with bf16 operands, fp32 TMEM accumulators.
//
// One persistent, warp-specialised kernel covers every dense matmul on the SC-GRPO hot path
// (K1 patch-embed, K4/K12 QKV + proj, K7/K14 MLP, K8 merger, K15 lm_head, K17 their dgrad/wgrad,
// K19 the decode "skinny" products via operand swap) and the batched QK^T / PV products of attention:
//   warp 0   : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1   : MMA issuer     (one elected lane issues tcgen05.mma, UMMA 128 x block_n x 16, commit -> mbarrier)
//   warp 2   : TMEM allocator (512 columns = two accumulator stages so tile i+1's MMAs overlap tile i's epilogue)
//   warps 4-7: epilogue       (tcgen05.ld 32x32b -> registers -> fused epilogue -> global)
// Both operands may be K-major (reduction dim contiguous: activations x[M,K], torch Linear weights W[N,K]) or
// MN-major (free dim contiguous: W[N,K] read as B^T for dgrad, dY / X read transposed for wgrad, V for P*V), which
// is what lets forward, dgrad and wgrad all run from the same row-major buffers without a transpose pass.
//
// The reference reaches these products through cuBLAS via torch (`model(**inputs).logits`,
// ref: train/stage_rl/trainer/sc_grpo_trainer.py:505); there is no reference native code to follow.
#include "gemm_sm100.cuh"
#include "ptx.cuh"
#include "runtime.h"

#include <map>
#include <vector>
#include <array>
#include <string>
#include <mutex>
#include <tuple>
#include <cstring>
#include <cstdio>

namespace iadr1 {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int kThreads = 256;
static constexpr int kMaxStages = 12;

struct TileInfo {
  int z, m0, n0, n_blk, kb_begin, kb_end, split;
  bool skip;
};

__device__ __forceinline__ TileInfo decode_tile(const GemmArgs& g, int t, int tiles_m, int tiles_n) {
  TileInfo ti;
  const int mn = tiles_m * tiles_n;
  const int per_z = mn * g.split_k;
  ti.z = t / per_z;
  int r = t - ti.z * per_z;
  const int split = r / mn;
  ti.split = split;
  r -= split * mn;
  // rasterisation: consecutive CTAs of a wave share the operand that is re-read from L2. m fastest (default) streams B
  // once and re-reads A; n fastest (raster_n) streams A once and re-reads B - the host picks the cheaper one.
  int m_blk;
  if (g.raster_n) {
    m_blk = r / tiles_n;
    ti.n_blk = r - m_blk * tiles_n;
  } else {
    ti.n_blk = r / tiles_m;
    m_blk = r - ti.n_blk * tiles_m;
  }
  ti.m0 = m_blk * BM;
  ti.n0 = ti.n_blk * g.block_n;
  int k_lo = 0, k_hi = g.K;
  if (g.kmode == 1) {
    k_hi = min(g.K, max(0, ti.m0 + BM + g.causal_off));
  } else if (g.kmode == 2) {
    k_lo = min(g.K, max(0, ti.m0 - g.causal_off));
  }
  const int kb_lo = k_lo / BK;
  const int kb_hi = (k_hi + BK - 1) / BK;
  const int nkb = max(0, kb_hi - kb_lo);
  const int per = (nkb + g.split_k - 1) / g.split_k;
  ti.kb_begin = kb_lo + split * per;
  ti.kb_end = min(kb_lo + nkb, ti.kb_begin + per);
  if (ti.kb_end < ti.kb_begin) ti.kb_end = ti.kb_begin;
  ti.skip = g.skip_mode && (ti.n0 > ti.m0 + BM - 1 + g.causal_off);
  return ti;
}

// Work iterator shared by the three warp roles. Tile mode: output tiles (x split_k x batch) round-robin over the
// persistent CTAs. Stream-K mode (g.stream_k; single problem, full K, atomic fp32 output): the tiles_m * tiles_n * nkb
// k-block units are cut into gridDim.x equal contiguous runs, so every SM streams the same number of weight bytes
// even when the tile count is not a multiple of the SM count (the decode gate_up product: 172 tiles on 148 SMs);
// a run crossing a tile boundary yields one segment per tile, each added atomically into C.
struct WorkIter {
  int cursor, end, stride;
  __device__ __forceinline__ WorkIter(const GemmArgs& g, int total_tiles, int tiles_mn) {
    if (g.stream_k) {
      const int nkb = (g.K + BK - 1) / BK;
      const int units = tiles_mn * nkb;
      const int per = (units + (int)gridDim.x - 1) / (int)gridDim.x;
      cursor = min(units, (int)blockIdx.x * per);
      end = min(units, cursor + per);
      stride = 0;
    } else {
      cursor = blockIdx.x;
      end = total_tiles;
      stride = gridDim.x;
    }
  }
  __device__ __forceinline__ bool next(const GemmArgs& g, int tiles_m, int tiles_n, TileInfo& ti) {
    if (cursor >= end) return false;
    if (!g.stream_k) {
      ti = decode_tile(g, cursor, tiles_m, tiles_n);
      cursor += stride;
      return true;
    }
    const int nkb = (g.K + BK - 1) / BK;
    const int tile = cursor / nkb;
    const int kb = cursor - tile * nkb;
    const int len = min(nkb - kb, end - cursor);
    ti.z = 0;
    ti.n_blk = tile / tiles_m;
    ti.m0 = (tile - ti.n_blk * tiles_m) * BM;
    ti.n0 = ti.n_blk * g.block_n;
    ti.kb_begin = kb;
    ti.kb_end = kb + len;
    ti.split = kb ? 1 : 0;   // bias / residual are added by the segment that starts the tile
    ti.skip = false;
    cursor += len;
    return true;
  }
};

template <int W>
__device__ __forceinline__ void epilogue_chunk(const GemmArgs& g, const TileInfo& ti, const uint32_t (&v)[W], int m,
                                               int nb, bool have_acc, bool vec_ok, float& run_max, float& run_sum,
                                               float& tgt, bool& tgt_found, int label, float row_lse,
                                               float row_g) {
  float acc[W];
#pragma unroll
  for (int j = 0; j < W; ++j) acc[j] = have_acc ? g.alpha * __uint_as_float(v[j]) : 0.f;
  const int nvalid = min(W, g.N - nb);
  if (nvalid <= 0 || m >= g.M) return;

  if (g.epi == EPI_LSE) {
    float cmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < W; ++j)
      if (j < nvalid) cmax = fmaxf(cmax, acc[j]);
    const float nmax = fmaxf(run_max, cmax);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < W; ++j)
      if (j < nvalid) s += __expf(acc[j] - nmax);
    run_sum = run_sum * __expf(run_max - nmax) + s;
    run_max = nmax;
    if (label >= nb && label < nb + nvalid) {
#pragma unroll
      for (int j = 0; j < W; ++j)
        if (nb + j == label) tgt = acc[j];
      tgt_found = true;
    }
    return;
  }
  if (g.epi == EPI_DLOGITS) {
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const float p = __expf(acc[j] - row_lse);
      acc[j] = (p - ((nb + j) == label ? 1.f : 0.f)) * row_g;
    }
  }

  const long long zoff = (long long)(ti.z % g.batch_lo) * g.c_bs_lo + (long long)(ti.z / g.batch_lo) * g.c_bs_hi;
  if (g.bias != nullptr && ti.split == 0) {
    if (g.bias_per_m) {
      const float b = __bfloat162float(g.bias[m]);
#pragma unroll
      for (int j = 0; j < W; ++j) acc[j] += b;
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j)
        if (j < nvalid) acc[j] += __bfloat162float(g.bias[nb + j]);
    }
  }
  if (g.trans_c) {
    // C^T[n][m]: lanes of a warp hold consecutive m, so each j is one coalesced row segment.
#pragma unroll
    for (int j = 0; j < W; ++j) {
      if (j < nvalid) {
        const long long idx = zoff + (long long)(nb + j) * g.ldc + m;
        float x = acc[j];
        if (g.residual && ti.split == 0) x += __bfloat162float(g.residual[idx]);
        if (g.atomic) {
          atomicAdd(reinterpret_cast<float*>(g.C) + idx, x);
        } else if (g.c_f32) {
          float* p = reinterpret_cast<float*>(g.C) + idx;
          *p = g.accumulate ? (*p + x) : x;
        } else {
          reinterpret_cast<__nv_bfloat16*>(g.C)[idx] = __float2bfloat16(x);
        }
      }
    }
    return;
  }
  const long long idx0 = zoff + (long long)m * g.ldc + nb;
  if (g.residual && ti.split == 0) {
    if (vec_ok && nvalid == W) {
      const uint4* rp = reinterpret_cast<const uint4*>(g.residual + idx0);
#pragma unroll
      for (int q = 0; q < W / 8; ++q) {
        const uint4 r = rp[q];
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(h[e]);
          acc[q * 8 + 2 * e] += f.x;
          acc[q * 8 + 2 * e + 1] += f.y;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j)
        if (j < nvalid) acc[j] += __bfloat162float(g.residual[idx0 + j]);
    }
  }
  if (g.atomic) {
    float* p = reinterpret_cast<float*>(g.C) + idx0;
#pragma unroll
    for (int j = 0; j < W; ++j)
      if (j < nvalid) atomicAdd(p + j, acc[j]);
  } else if (g.c_f32) {
    float* p = reinterpret_cast<float*>(g.C) + idx0;
    if (vec_ok && nvalid == W) {
      float4* p4 = reinterpret_cast<float4*>(p);
#pragma unroll
      for (int q = 0; q < W / 4; ++q) {
        float4 o = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        if (g.accumulate) {
          const float4 c = p4[q];
          o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
        }
        p4[q] = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j)
        if (j < nvalid) p[j] = g.accumulate ? (p[j] + acc[j]) : acc[j];
    }
  } else {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(g.C) + idx0;
    if (vec_ok && nvalid == W) {
      uint4* p4 = reinterpret_cast<uint4*>(p);
#pragma unroll
      for (int q = 0; q < W / 8; ++q) {
        uint4 o;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(acc[q * 8 + 2 * e], acc[q * 8 + 2 * e + 1]);
        p4[q] = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j)
        if (j < nvalid) p[j] = __float2bfloat16(acc[j]);
    }
  }
}

// Lean epilogue for the common store forms (row-major C, full 32-column chunk, 16-byte aligned rows): the generic
// epilogue_chunk above is ~3000 SASS instructions per chunk once every variant is inlined, which made the epilogue the
// bottleneck of small-K products (attention QK^T / PV: ~6 us per 128 x 256 tile). This path is ~100 instructions.
//   bf16:  C = bf16(alpha * acc (+ bias[n]) (+ residual))         fp32: C (+)= alpha * acc (+ bias[n])
__device__ __forceinline__ void epilogue_lean32(const GemmArgs& g, const uint32_t (&v)[32], float alpha, bool have_acc,
                                                void* crow, const __nv_bfloat16* rrow, const __nv_bfloat16* bias_n,
                                                int col) {
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = have_acc ? alpha * __uint_as_float(v[j]) : 0.f;
  if (bias_n != nullptr) {
    const uint4* bp = reinterpret_cast<const uint4*>(bias_n + col);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 b = bp[q];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        acc[q * 8 + 2 * e] += f.x;
        acc[q * 8 + 2 * e + 1] += f.y;
      }
    }
  }
  if (g.c_f32) {
    float4* p4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(crow) + col);
    if (g.accumulate) {
      float4 c[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) c[q] = p4[q];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        p4[q] = make_float4(c[q].x + acc[4 * q], c[q].y + acc[4 * q + 1], c[q].z + acc[4 * q + 2], c[q].w + acc[4 * q + 3]);
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) p4[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    }
    return;
  }
  if (rrow != nullptr) {
    const uint4* rp = reinterpret_cast<const uint4*>(rrow + col);
    uint4 r[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) r[q] = rp[q];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r[q]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        acc[q * 8 + 2 * e] += f.x;
        acc[q * 8 + 2 * e + 1] += f.y;
      }
    }
  }
  uint4* p4 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(crow) + col);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 o;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(acc[q * 8 + 2 * e], acc[q * 8 + 2 * e + 1]);
    p4[q] = o;
  }
}

// kMinCtas = 2 is the decode-chain variant: <= 128 registers, a short smem ring and a TMEM allocation sized to the two
// block_n-wide accumulator stages, so that under programmatic dependent launch the NEXT kernel's CTAs become resident
// (and request their first weight tiles) while this kernel's CTAs are still draining.
template <int kMinCtas>
__global__ void __launch_bounds__(kThreads, kMinCtas)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmC, const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array (keeps the shared address space). Launches that need
  // every byte (g.smem_tight: two co-resident CTAs with a 3-stage ring) request no slack and rely on the window base being
  // 1 KiB aligned, which holds with no static shared memory - checked here instead of assumed.
  const uint32_t smem_pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  if (g.smem_tight && smem_pad != 0) __trap();
  uint8_t* smem = smem_raw + smem_pad;

  // Programmatic dependent launch: let the next kernel's CTAs become resident as early as resources allow. Every
  // dependent still executes griddepcontrol.wait (full completion + flush of this grid) before touching our output.
  if (threadIdx.x == 0) pdl_trigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int a_bytes = BM * BK * 2;                 // 16 KiB
  const int b_bytes = g.block_n * BK * 2;          // block_n * 128 B
  const int stage_bytes = a_bytes + ((b_bytes + 1023) & ~1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)g.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int tiles_m = (g.M + BM - 1) / BM;
  const int tiles_n = (g.N + g.block_n - 1) / g.block_n;
  const int total_tiles = tiles_m * tiles_n * g.split_k * g.batch;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (g.tma_store) tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < g.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, g.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tx_bytes = a_bytes + b_bytes;
      auto load_a = [&](const TileInfo& ti, int kb, int stg, int zlo, int zhi) {
        uint8_t* sa = smem + (size_t)stg * stage_bytes;
        if (g.epi == EPI_SWIGLU) {
          // rows [f0, f0 + 64) of the gate block and of the up block form ONE 128-row MMA tile (TMEM lanes 0-63 / 64-127)
          const int f0 = ti.m0 >> 1;
          tma_load_4d(sa, &tmA, &full_bar[stg], kb * BK, f0, zlo, zhi);
          tma_load_4d(sa + 8192, &tmA, &full_bar[stg], kb * BK, g.up_row_off + f0, zlo, zhi);
        } else if (!g.a_mn) {
          tma_load_4d(sa, &tmA, &full_bar[stg], kb * BK, ti.m0, zlo, zhi);
        } else if (g.a_chunked) {
          // MN-major tile as ONE box {64 mn, BK k, 2 chunks} of the 3-D view [mn / 64][k][64]: same smem image as the two
          // 2-D boxes below, a third of the TMA instructions per k-block (the producer thread was issue-bound)
          tma_load_5d(sa, &tmA, &full_bar[stg], 0, kb * BK, ti.m0 >> 6, zlo, zhi);
        } else {
          tma_load_4d(sa, &tmA, &full_bar[stg], ti.m0, kb * BK, zlo, zhi);
          tma_load_4d(sa + 8192, &tmA, &full_bar[stg], ti.m0 + 64, kb * BK, zlo, zhi);
        }
      };
      auto load_b = [&](const TileInfo& ti, int kb, int stg, int zlo_b, int zhi) {
        uint8_t* sb = smem + (size_t)stg * stage_bytes + a_bytes;
        if (g.epi == EPI_SWIGLU_T) {
          // ONE box {64 k, 128 rows, 2 halves} of the rank-3 view [gate | up][I][K]: accumulator columns 0-127 = gate features
          // [n0 / 2, + 128), columns 128-255 = the up features of the same range
          tma_load_4d(sb, &tmB, &full_bar[stg], kb * BK, ti.n0 >> 1, 0, 0);
        } else if (!g.b_mn) {
          tma_load_4d(sb, &tmB, &full_bar[stg], kb * BK, ti.n0, zlo_b, zhi);
        } else if (g.b_chunked) {
          tma_load_5d(sb, &tmB, &full_bar[stg], 0, kb * BK, ti.n0 >> 6, zlo_b, zhi);
        } else {
          for (int j = 0; j < g.block_n / 64; ++j)
            tma_load_4d(sb + j * 8192, &tmB, &full_bar[stg], ti.n0 + 64 * j, kb * BK, zlo_b, zhi);
        }
      };
      // Programmatic dependent launch: weights (A, when a_static) of the first tile's first stages are requested
      // BEFORE waiting for the preceding kernel, so their HBM latency overlaps that kernel's tail.
      int npre = 0;
      TileInfo t0;
      WorkIter it0(g, total_tiles, tiles_m * tiles_n);
      if (g.a_static && it0.next(g, tiles_m, tiles_n, t0)) {
        if (!t0.skip) {
          npre = min(g.stages, t0.kb_end - t0.kb_begin);
          for (int i = 0; i < npre; ++i) {
            mbar_arrive_expect_tx(&full_bar[i], tx_bytes);
            load_a(t0, t0.kb_begin + i, i, t0.z % g.batch_lo, t0.z / g.batch_lo);
          }
        }
      }
      pdl_wait();
      bool first_tile = true;
      TileInfo ti;
      WorkIter it(g, total_tiles, tiles_m * tiles_n);
      while (it.next(g, tiles_m, tiles_n, ti)) {
        if (ti.skip) continue;
        const int zlo = ti.z % g.batch_lo, zhi = ti.z / g.batch_lo;
        const int zlo_b = zlo / g.b_lo_div;
        for (int kb = ti.kb_begin; kb < ti.kb_end; ++kb) {
          if (first_tile && (kb - ti.kb_begin) < npre) {
            load_b(ti, kb, s, zlo_b, zhi);   // A and the expect_tx were issued before the wait
          } else {
            mbar_wait(&empty_bar[s], ph ^ 1);
            mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
            load_a(ti, kb, s, zlo, zhi);
            load_b(ti, kb, s, zlo_b, zhi);
          }
          if (++s == g.stages) { s = 0; ph ^= 1; }
        }
        first_tile = false;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(BM, g.block_n, g.a_mn, g.b_mn);
      // K-major SW128: 8-row groups 1024 B apart (SBO), LBO unused. MN-major SW128: 64-wide MN chunks 8192 B apart
      // (LBO = BK rows * 128 B), 8-row K groups 1024 B apart (SBO).
      const uint32_t a_lbo = g.a_mn ? 8192u : 16u, b_lbo = g.b_mn ? 8192u : 16u;
      const uint32_t a_kstep = g.a_mn ? (2048u >> 4) : (32u >> 4);
      const uint32_t b_kstep = g.b_mn ? (2048u >> 4) : (32u >> 4);
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      TileInfo ti;
      WorkIter it(g, total_tiles, tiles_m * tiles_n);
      while (it.next(g, tiles_m, tiles_n, ti)) {
        if (ti.skip) continue;
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * (g.tmem_cols >> 1);
        for (int kb = ti.kb_begin; kb < ti.kb_end; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
          const uint32_t sb = sa + a_bytes;
          const uint64_t adesc = make_smem_desc_sw128(sa, a_lbo, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(sb, b_lbo, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            umma_bf16(d_tmem, adesc + (uint64_t)(k * a_kstep), bdesc + (uint64_t)(k * b_kstep), idesc,
                      (kb > ti.kb_begin || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
          if (++s == g.stages) { s = 0; ph ^= 1; }
        }
        umma_commit(&tfull_bar[as]);
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    pdl_wait();
    const int q = warp & 3;
    const bool vec_ok = ((g.ldc & 7) == 0) && ((g.c_bs_lo & 7) == 0) && ((g.c_bs_hi & 7) == 0) &&
                        ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) &&
                        (g.residual == nullptr || (reinterpret_cast<uintptr_t>(g.residual) & 15) == 0);
    int as = 0;
    int tbuf = 0;
    uint32_t aph = 0;
    TileInfo ti;
    WorkIter it(g, total_tiles, tiles_m * tiles_n);
    while (it.next(g, tiles_m, tiles_n, ti)) {
      if (ti.skip) continue;
      const bool have_acc = ti.kb_end > ti.kb_begin;
      if (kMinCtas == 1 && g.epi == EPI_SWIGLU_BWD) {
        // Ask L2 for this thread's WHOLE row segment of gate and of up (<= 512 contiguous bytes each) before waiting for the
        // accumulator: the epilogue's own reads are 64-byte pieces of 128 different rows per chunk, and issued one chunk at
        // a time every piece opens a DRAM page of its own (774 MB in 64-byte reads: the fused kernel took 500 us against 320 us
        // for the bare product, whatever the lookahead or the arithmetic cost). One bulk prefetch per row segment turns that
        // into eight times fewer, eight times longer DRAM bursts, and the loads below hit L2.
        const int mp = ti.m0 + q * 32 + lane;
        const int ncol = min(g.block_n, g.N - ti.n0);
        if (mp < g.M && ncol > 0) {
          const __nv_bfloat16* prow = g.gu_out + (long long)mp * g.gu_ld + ti.n0;
          const uint32_t nbytes = (uint32_t)ncol * 2u;              // multiple of 16: I % 8 == 0
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(prow), "r"(nbytes) : "memory");
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(prow + g.N), "r"(nbytes) : "memory");
        }
      }
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const int m = ti.m0 + q * 32 + lane;
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + as * (g.tmem_cols >> 1);
      float run_max = -INFINITY, run_sum = 0.f, tgt = 0.f, row_lse = 0.f, row_g = 0.f;
      bool tgt_found = false;
      int label = -1;
      if ((g.epi == EPI_LSE || g.epi == EPI_DLOGITS) && m < g.M) {
        label = g.labels[m];
        if (g.epi == EPI_DLOGITS) {
          row_lse = g.lse[m];
          row_g = g.gscale[m];
        }
      }
      if (g.epi == EPI_SWIGLU) {
        // out[n][f] = bf16(silu(gate)) * up for the tile's 64 features: both halves meet in the staging tile, which holds
        // 64 decode rows at a time (so that block_n = 128 still leaves room for two CTAs per SM)
        float* sC = reinterpret_cast<float*>(smem + (size_t)g.stages * stage_bytes + 512);
        const int ml = q * 32 + lane;
        const int f0 = ti.m0 >> 1;
        __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(g.C);
        const int chunk = g.swiglu_rows;         // decode rows staged at a time (64, or 32 when two CTAs share an SM)
        for (int cb = 0; cb < g.block_n; cb += chunk) {
          const int ncols = min(chunk, g.block_n - cb);
          for (int c0 = 0; c0 < ncols; c0 += 16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(taddr + cb + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) sC[(c0 + j) * BM + ml] = have_acc ? g.alpha * __uint_as_float(v[j]) : 0.f;
          }
          if (cb + chunk >= g.block_n) {        // accumulator fully read: hand the TMEM stage back to the MMA warp
            tc_fence_before();
            mbar_arrive(&tempty_bar[as]);
          }
          named_bar_sync(1, 128);
          const int nrows = min(ncols, g.N - ti.n0 - cb);
          for (int i = threadIdx.x - 128; i < nrows * 32; i += 128) {
            const int n = i >> 5, f = (i & 31) * 2;
            const float2 gg = *reinterpret_cast<const float2*>(sC + n * BM + f);
            const float2 uu = *reinterpret_cast<const float2*>(sC + n * BM + 64 + f);
            auto act = [](float gv, float uv) {
              gv = __bfloat162float(__float2bfloat16(gv));
              const float sv = __bfloat162float(__float2bfloat16(silu_f(gv)));
              return sv * __bfloat162float(__float2bfloat16(uv));
            };
            if (f0 + f < (g.M >> 1))
              *reinterpret_cast<__nv_bfloat162*>(outp + (long long)(ti.n0 + cb + n) * g.ldc + f0 + f) =
                  __floats2bfloat162_rn(act(gg.x, uu.x), act(gg.y, uu.y));
          }
          named_bar_sync(1, 128);
        }
        if (++as == 2) { as = 0; aph ^= 1; }
        continue;
      }
      if (g.epi == EPI_SWIGLU_T) {
        // thread = token row m: gate columns [c, c + 32) and up columns [128 + c, ...) of the accumulator -> act[m][f0 + c ..]
        const int f0 = ti.n0 >> 1, half_i = g.N >> 1;
        __nv_bfloat16* arow = reinterpret_cast<__nv_bfloat16*>(g.C) + (long long)m * g.ldc + f0;
        __nv_bfloat16* grow = g.gu_out ? g.gu_out + (long long)m * g.gu_ld + f0 : nullptr;
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t gv[32], uv[32];
          tmem_ld_32x32b_x32(taddr + c0, gv);
          tmem_ld_32x32b_x32(taddr + 128 + c0, uv);
          tmem_ld_wait();
          if (m < g.M) {
            uint4 og[4], ou[4], oa[4];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const __nv_bfloat162 gb = __floats2bfloat162_rn(have_acc ? __uint_as_float(gv[j]) : 0.f, have_acc ? __uint_as_float(gv[j + 1]) : 0.f);
              const __nv_bfloat162 ub = __floats2bfloat162_rn(have_acc ? __uint_as_float(uv[j]) : 0.f, have_acc ? __uint_as_float(uv[j + 1]) : 0.f);
              const float2 gf = __bfloat1622float2(gb), uf = __bfloat1622float2(ub);
              const float s0 = __bfloat162float(__float2bfloat16(silu_f(gf.x)));
              const float s1 = __bfloat162float(__float2bfloat16(silu_f(gf.y)));
              const __nv_bfloat162 ab = __floats2bfloat162_rn(s0 * uf.x, s1 * uf.y);
              reinterpret_cast<__nv_bfloat162*>(og)[j >> 1] = gb;
              reinterpret_cast<__nv_bfloat162*>(ou)[j >> 1] = ub;
              reinterpret_cast<__nv_bfloat162*>(oa)[j >> 1] = ab;
            }
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              reinterpret_cast<uint4*>(arow + c0)[q4] = oa[q4];
              if (grow) {
                reinterpret_cast<uint4*>(grow + c0)[q4] = og[q4];
                reinterpret_cast<uint4*>(grow + half_i + c0)[q4] = ou[q4];
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&tempty_bar[as]);
        if (++as == 2) { as = 0; aph ^= 1; }
        continue;
      }
      if (kMinCtas == 1 && g.epi == EPI_SWIGLU_BWD) {   // training-only epilogue: not compiled into the 128-register decode variant
        // thread = token row m; accumulator column c = feature n0 + c of dact (block_n = 256: eight 32-column chunks). gate / up
        // of the same features are read from gu two chunks ahead of the one being processed and overwritten with dgate / dup:
        // the act_mul_bwd row kernel's arithmetic on the same bf16-rounded inputs, without dact reaching HBM. Four warps doing
        // this math take ~23 us per tile (issue-bound, not latency-bound: lookahead 1 -> 2 changed nothing), so the host uses
        // this epilogue only where a tile's MMA is longer than that (K >= 3072, model.cu).
        const int I = g.N;
        __nv_bfloat16* grow = g.gu_out + (long long)m * g.gu_ld + ti.n0;
        const bool row_ok = m < g.M;
        uint4 bg[3][4], bu[3][4];
        auto fetch = [&](int c0, uint4 (&dg)[4], uint4 (&du)[4]) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const bool ok = row_ok && ti.n0 + c0 + q4 * 8 < I;      // I % 8 == 0 (host check): whole 16-byte groups
            dg[q4] = ok ? reinterpret_cast<const uint4*>(grow + c0)[q4] : make_uint4(0, 0, 0, 0);
            du[q4] = ok ? reinterpret_cast<const uint4*>(grow + I + c0)[q4] : make_uint4(0, 0, 0, 0);
          }
        };
        fetch(0, bg[0], bu[0]);
        fetch(32, bg[1], bu[1]);
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
          const int c0 = ci * 32;
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + c0, v);
          if (ci + 2 < 8) fetch(c0 + 64, bg[(ci + 2) % 3], bu[(ci + 2) % 3]);
          tmem_ld_wait();
          const uint4 (&cg)[4] = bg[ci % 3];
          const uint4 (&cu)[4] = bu[ci % 3];
          uint4 og[4], ou[4];
          // Staged over 16 elements at a time so that the SFU ops (ex2, rcp) of different elements are independent and
          // adjacent in program order: the four epilogue warps run one per scheduler, in order - written element by element
          // the dependent chain cvt -> ex2 -> add -> rcp -> mul ran at 0.13 instructions per cycle (ncu) and the epilogue took
          // 23 us per tile against 15 us of MMA.
#pragma unroll
          for (int h16 = 0; h16 < 2; ++h16) {
            float gf[16], uf[16], dd[16], sg[16];
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
              const int j = h16 * 16 + k;
              const float2 g2 = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(cg)[j >> 1]);
              const float2 u2 = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(cu)[j >> 1]);
              const float2 d2 = __bfloat1622float2(__floats2bfloat162_rn(have_acc ? g.alpha * __uint_as_float(v[j]) : 0.f,
                                                                         have_acc ? g.alpha * __uint_as_float(v[j + 1]) : 0.f));
              gf[k] = g2.x; gf[k + 1] = g2.y; uf[k] = u2.x; uf[k + 1] = u2.y; dd[k] = d2.x; dd[k + 1] = d2.y;
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) sg[k] = __expf(-gf[k]);
#pragma unroll
            for (int k = 0; k < 16; ++k) sg[k] = rcp_approx_f(__fadd_rn(1.f, sg[k]));      // sigmoid, as sigmoid_f()
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
              const int j = h16 * 16 + k;
              float dg0, dg1, du0, du1;
              swiglu_bwd_s(dd[k], gf[k], uf[k], sg[k], dg0, du0);
              swiglu_bwd_s(dd[k + 1], gf[k + 1], uf[k + 1], sg[k + 1], dg1, du1);
              reinterpret_cast<__nv_bfloat162*>(og)[j >> 1] = __floats2bfloat162_rn(dg0, dg1);
              reinterpret_cast<__nv_bfloat162*>(ou)[j >> 1] = __floats2bfloat162_rn(du0, du1);
            }
          }
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            if (row_ok && ti.n0 + c0 + q4 * 8 < I) {
              reinterpret_cast<uint4*>(grow + c0)[q4] = og[q4];
              reinterpret_cast<uint4*>(grow + I + c0)[q4] = ou[q4];
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&tempty_bar[as]);
        if (++as == 2) { as = 0; aph ^= 1; }
        continue;
      }
      if (kMinCtas == 1 && g.tma_store) {
        // Plain store / fp32 accumulate through the TMA unit. Each thread owns accumulator row r; written straight to global
        // memory that is 32 different cache lines per store instruction, and those wavefronts share the L1 / shared-memory
        // pipe with the TMA loads feeding the MMA (measured: the K = 2048 products ran 16-19 % faster with the stores
        // removed). So the row's 128 bytes of a chunk (64 bf16 / 32 fp32 columns) go into a 128B-swizzled staging box
        // (conflict-free 16-byte shared stores) and ONE thread hands the [128 rows x 128 B] box to the TMA unit, which
        // clips the M / N tails and - for the weight gradients - adds into C at L2 (no read-modify-write by the SM).
        uint8_t* sT = smem + (size_t)g.stages * stage_bytes + 1024;
        const int r = q * 32 + lane;
        const int cw = g.c_f32 ? 32 : 64;
        const bool issuer = threadIdx.x == 128;
        const __nv_bfloat16* bias_n = (g.bias != nullptr && ti.split == 0) ? g.bias + ti.n0 : nullptr;
        int buf = tbuf;      // the two staging boxes alternate ACROSS tiles too (a tile may have an odd number of chunks)
        for (int c0 = 0; c0 < g.block_n && ti.n0 + c0 < g.N; c0 += cw) {
          uint32_t v[32], v2[32];
          tmem_ld_32x32b_x32(taddr + c0, v);
          if (!g.c_f32) tmem_ld_32x32b_x32(taddr + c0 + 32, v2);
          if (issuer) bulk_wait_read1();                 // the store that last read this buffer (two chunks ago) is done with it
          named_bar_sync(1, 128);
          tmem_ld_wait();
          uint8_t* row = sT + buf * 16384 + r * 128;
          if (g.c_f32) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              float4 o;
              o.x = have_acc ? g.alpha * __uint_as_float(v[4 * k]) : 0.f;
              o.y = have_acc ? g.alpha * __uint_as_float(v[4 * k + 1]) : 0.f;
              o.z = have_acc ? g.alpha * __uint_as_float(v[4 * k + 2]) : 0.f;
              o.w = have_acc ? g.alpha * __uint_as_float(v[4 * k + 3]) : 0.f;
              *reinterpret_cast<float4*>(row + ((k ^ (r & 7)) << 4)) = o;
            }
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              float a8[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const uint32_t raw = k < 4 ? v[8 * k + e] : v2[8 * (k - 4) + e];
                a8[e] = have_acc ? g.alpha * __uint_as_float(raw) : 0.f;
              }
              if (bias_n != nullptr && ti.n0 + c0 + 8 * k < g.N) {
                const uint4 b = *reinterpret_cast<const uint4*>(bias_n + c0 + 8 * k);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __bfloat1622float2(h[e]);
                  a8[2 * e] += f.x;
                  a8[2 * e + 1] += f.y;
                }
              }
              uint4 o;
              __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int e = 0; e < 4; ++e) oh[e] = __floats2bfloat162_rn(a8[2 * e], a8[2 * e + 1]);
              *reinterpret_cast<uint4*>(row + ((k ^ (r & 7)) << 4)) = o;
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(1, 128);
          if (issuer) {
            if (g.tma_store == 2) tma_reduce_add_2d(&tmC, sT + buf * 16384, ti.n0 + c0, ti.m0);
            else tma_store_2d(&tmC, sT + buf * 16384, ti.n0 + c0, ti.m0);
            bulk_commit();
          }
          buf ^= 1;
        }
        tbuf = buf;
        tc_fence_before();
        mbar_arrive(&tempty_bar[as]);
        if (++as == 2) { as = 0; aph ^= 1; }
        continue;
      }
      if (g.bulk_red) {
        // Transposed fp32 accumulate-into-C via the TMA unit: the tile is staged as sC[n][128 m] (lanes = consecutive m:
        // conflict-free stores) and every decode row n is added to C^T[n][m0 .. m0 + 128) by ONE bulk reduction.
        float* sC = reinterpret_cast<float*>(smem + (size_t)g.stages * stage_bytes + 512);
        const int ml = q * 32 + lane;
        float bias_m = 0.f;
        if (g.bias != nullptr && ti.split == 0 && g.bias_per_m && m < g.M) bias_m = __bfloat162float(g.bias[m]);
        for (int c0 = 0; c0 < g.block_n; c0 += 16) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float x = have_acc ? g.alpha * __uint_as_float(v[j]) : 0.f;
            x += bias_m;
            if (g.bias != nullptr && ti.split == 0 && !g.bias_per_m && ti.n0 + c0 + j < g.N)
              x += __bfloat162float(g.bias[ti.n0 + c0 + j]);
            sC[(c0 + j) * BM + ml] = x;
          }
        }
        tc_fence_before();
        mbar_arrive(&tempty_bar[as]);
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        const int mw = min(BM, g.M - ti.m0);          // live features of this tile (multiple of 4, checked on the host)
        const int nrows = min(g.block_n, g.N - ti.n0);
        for (int n = threadIdx.x - 128; n < nrows; n += 128)
          bulk_reduce_add_f32(reinterpret_cast<float*>(g.C) + (long long)(ti.n0 + n) * g.ldc + ti.m0, sC + n * BM,
                              (uint32_t)mw * 4u);
        bulk_commit();
        bulk_wait_read0();                            // staging tile may be overwritten by the next segment
        named_bar_sync(1, 128);
        if (++as == 2) { as = 0; aph ^= 1; }
        continue;
      }
      int c0 = 0;
      // row base pointers for the lean path (computed once per tile)
      const bool lean = (g.epi == EPI_STORE) && !g.trans_c && !g.atomic && vec_ok && m < g.M &&
                        !(g.c_f32 && g.residual != nullptr) &&
                        (g.bias == nullptr || (!g.bias_per_m && ti.split == 0 && (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0 &&
                                               (ti.n0 & 7) == 0) || ti.split != 0);
      const long long zoff_t = (long long)(ti.z % g.batch_lo) * g.c_bs_lo + (long long)(ti.z / g.batch_lo) * g.c_bs_hi;
      const long long roff = zoff_t + (long long)m * g.ldc + ti.n0;
      void* crow = g.c_f32 ? static_cast<void*>(reinterpret_cast<float*>(g.C) + roff)
                           : static_cast<void*>(reinterpret_cast<__nv_bfloat16*>(g.C) + roff);
      const __nv_bfloat16* rrow = (g.residual != nullptr && ti.split == 0) ? g.residual + roff : nullptr;
      const __nv_bfloat16* bias_n = (g.bias != nullptr && ti.split == 0 && !g.bias_per_m) ? g.bias + ti.n0 : nullptr;
      for (; c0 + 32 <= g.block_n; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c0, v);
        tmem_ld_wait();
        if (lean && ti.n0 + c0 + 32 <= g.N)
          epilogue_lean32(g, v, g.alpha, have_acc, crow, rrow, bias_n, c0);
        else
          epilogue_chunk<32>(g, ti, v, m, ti.n0 + c0, have_acc, vec_ok, run_max, run_sum, tgt, tgt_found, label,
                             row_lse, row_g);
      }
      if (c0 < g.block_n) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(taddr + c0, v);
        tmem_ld_wait();
        epilogue_chunk<16>(g, ti, v, m, ti.n0 + c0, have_acc, vec_ok, run_max, run_sum, tgt, tgt_found, label,
                           row_lse, row_g);
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
      if (g.epi == EPI_LSE && m < g.M) {
        g.part_max[(long long)m * tiles_n + ti.n_blk] = run_max;
        g.part_sum[(long long)m * tiles_n + ti.n_blk] = run_sum;
        if (tgt_found) g.tgt_logit[m] = tgt;
      }
      if (++as == 2) { as = 0; aph ^= 1; }
    }
  }

  if ((g.bulk_red || g.tma_store) && warp >= 4) bulk_wait0();   // this thread's bulk reductions / tile stores have been performed
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, g.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// Host side: tensor-map cache + launch heuristics
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  long long d0, d1, d2, d3, s1, s2, s3;
  int b0, b1;
  bool operator<(const MapKey& o) const {
    return std::tie(ptr, d0, d1, d2, d3, s1, s2, s3, b0, b1) <
           std::tie(o.ptr, o.d0, o.d1, o.d2, o.d3, o.s1, o.s2, o.s3, o.b0, o.b1);
  }
};

static int make_map(CUtensorMap* out, const void* ptr, long long d0, long long d1, long long d2, long long d3,
                    long long s1, long long s2, long long s3, int b0, int b1, int b2 = 1) {
  static std::map<MapKey, CUtensorMap> cache;
  static std::mutex mu;
  MapKey key{ptr, d0, d1, d2, d3, s1, s2, s3, b0, b1 + (b2 << 16)};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error("cuTensorMapEncodeTiled entry point not found (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return set_error("gemm operand not 16-byte aligned");
  if (((s1 * 2) & 15) || ((s2 * 2) & 15) || ((s3 * 2) & 15))
    return set_error("gemm operand strides must be multiples of 8 elements (ld=%lld bs_lo=%lld bs_hi=%lld)", s1, s2,
                     s3);
  cuuint64_t dims[4] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2, (cuuint64_t)d3};
  cuuint64_t strides[3] = {(cuuint64_t)s1 * 2, (cuuint64_t)s2 * 2, (cuuint64_t)s3 * 2};
  cuuint32_t box[4] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error("cuTensorMapEncodeTiled failed (%d) dims=%lld,%lld,%lld,%lld strides=%lld,%lld,%lld box=%d,%d",
                     (int)r, d0, d1, d2, d3, s1, s2, s3, b0, b1);
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 8192) cache.clear();
  cache[key] = *out;
  return 0;
}

// K-major 2-D operand [rows][cols] (row stride ld elements) with a {box_cols, box_rows} box, 128-byte swizzle (decode chain).
// Output map for the TMA-store epilogue: C[rows][cols] (bf16 or f32, row stride ld elements), box = 128 rows x 128 bytes,
// 128B swizzle (the staging box the epilogue warps fill).
static int make_map_c(CUtensorMap* out, const void* ptr, long long cols, long long rows, long long ld, int f32) {
  static std::map<MapKey, CUtensorMap> cache;
  static std::mutex mu;
  MapKey key{ptr, cols, rows, ld, f32, 0, 0, 0, 0, -7};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error("cuTensorMapEncodeTiled entry point not found (no CUDA driver?)");
  const int esz = f32 ? 4 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled(C) failed (%d) cols=%lld rows=%lld ld=%lld", (int)r, cols, rows, ld);
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 8192) cache.clear();
  cache[key] = *out;
  return 0;
}

int make_tensor_map_2d(CUtensorMap* out, const void* ptr, long long cols, long long rows, long long ld, int box_cols, int box_rows) {
  return make_map(out, ptr, cols, rows, 1, 1, ld, ld, ld, box_cols, box_rows);
}

// Two row blocks [0, rows) and [rows, 2 rows) of one K-major matrix as ONE box {box_cols, box_rows, 2}: the fused gate|up
// weight of the decode chain, so that a k-block of both halves arrives with a single TMA instruction.
int make_tensor_map_pair(CUtensorMap* out, const void* ptr, long long cols, long long rows, long long ld, int box_cols, int box_rows) {
  return make_map(out, ptr, cols, rows, 2, 1, ld, rows * ld, 2 * rows * ld, box_cols, box_rows, 2);
}

// Rank-5 map of an MN-major operand through the view [batch_hi][batch_lo][mn / 64][k][64]; box {64, BK, chunks, 1, 1}.
static int make_map_chunked(CUtensorMap* out, const void* ptr, long long mn, long long k, long long n_lo, long long n_hi,
                            long long ld, long long bs_lo, long long bs_hi, int chunks) {
  static std::map<MapKey, CUtensorMap> cache;
  static std::mutex mu;
  MapKey key{ptr, mn, k, n_lo, n_hi, ld, bs_lo, bs_hi, chunks, -5};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error("cuTensorMapEncodeTiled entry point not found (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return set_error("gemm operand not 16-byte aligned");
  if (((ld * 2) & 15) || ((bs_lo * 2) & 15) || ((bs_hi * 2) & 15))
    return set_error("gemm operand strides must be multiples of 8 elements (ld=%lld bs_lo=%lld bs_hi=%lld)", ld, bs_lo, bs_hi);
  cuuint64_t dims[5] = {64, (cuuint64_t)k, (cuuint64_t)(mn / 64), (cuuint64_t)n_lo, (cuuint64_t)n_hi};
  cuuint64_t strides[4] = {(cuuint64_t)ld * 2, 128, (cuuint64_t)bs_lo * 2, (cuuint64_t)bs_hi * 2};
  cuuint32_t box[5] = {64, (cuuint32_t)BK, (cuuint32_t)chunks, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error("cuTensorMapEncodeTiled(chunked) failed (%d) mn=%lld k=%lld batch=%lld,%lld ld=%lld", (int)r, mn, k, n_lo,
                     n_hi, ld);
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 8192) cache.clear();
  cache[key] = *out;
  return 0;
}

int make_tensor_map_pair(CUtensorMap* out, const void* ptr, long long cols, long long rows, long long ld, int box_cols, int box_rows);

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

static int pick_block_n(int N, int b_mn) {
  if (b_mn) {
    // MN-major B is staged in 64-column swizzle atoms.
    if (N <= 64) return 64;
    if (N <= 128) return 128;
    if (N <= 192) return 192;
    return 256;
  }
  if (N <= 256) return (N + 15) / 16 * 16;
  // fewest tiles first, then least padding
  int best = 256, best_tiles = (N + 255) / 256;
  for (int bn = 240; bn >= 128; bn -= 16) {
    int tiles = (N + bn - 1) / bn;
    if (tiles == best_tiles) best = bn;
  }
  return best;
}

// ------------------------------------------------------------------------------------------------
// Optional per-launch timing (bench.py's live roofline): CUDA events around every GEMM launch that is not being
// captured into a graph, plus the algorithmic FLOPs of the launch. Off by default.
// ------------------------------------------------------------------------------------------------
struct GemmProf {
  bool enabled = false;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  std::vector<double> flops;
  std::vector<std::array<int, 8>> shapes;  // M, N, K, batch, a_mn, b_mn, epi, block_n
  std::mutex mu;
};
static GemmProf g_prof;

static bool prof_begin(cudaStream_t stream, cudaEvent_t* e0, cudaEvent_t* e1) {
  if (!g_prof.enabled) return false;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return false;
  std::lock_guard<std::mutex> lk(g_prof.mu);
  if (g_prof.used + 2 > g_prof.pool.size()) {
    const size_t old = g_prof.pool.size();
    g_prof.pool.resize(old + 4096);
    for (size_t i = old; i < g_prof.pool.size(); ++i) cudaEventCreate(&g_prof.pool[i]);
  }
  *e0 = g_prof.pool[g_prof.used++];
  *e1 = g_prof.pool[g_prof.used++];
  cudaEventRecord(*e0, stream);
  return true;
}

void gemm_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  g_prof.enabled = on != 0;
  g_prof.used = 0;
  g_prof.flops.clear();
  g_prof.shapes.clear();
}

// Caller must have synchronised the device. Returns the number of timed launches.
long long gemm_profile_collect(double* total_ms, double* total_flops, double* max_launch_ms, const char* csv_path) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  double ms = 0, fl = 0, mx = 0;
  const size_t n = g_prof.used / 2;
  std::map<std::array<int, 8>, std::array<double, 3>> by_shape;  // launches, ms, flops
  for (size_t i = 0; i < n; ++i) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, g_prof.pool[2 * i], g_prof.pool[2 * i + 1]) == cudaSuccess) {
      ms += t;
      if (t > mx) mx = t;
      fl += g_prof.flops[i];
      auto& e = by_shape[g_prof.shapes[i]];
      e[0] += 1;
      e[1] += t;
      e[2] += g_prof.flops[i];
    }
  }
  if (csv_path && csv_path[0]) {
    if (FILE* f = fopen(csv_path, "w")) {
      fprintf(f, "M,N,K,batch,a_mn,b_mn,epi,block_n,launches,total_ms,avg_us,tflops\n");
      for (auto& kv : by_shape) {
        const auto& k = kv.first;
        const auto& v = kv.second;
        fprintf(f, "%d,%d,%d,%d,%d,%d,%d,%d,%.0f,%.3f,%.2f,%.1f\n", k[0], k[1], k[2], k[3], k[4], k[5], k[6], k[7], v[0], v[1],
                v[1] / v[0] * 1e3, v[1] > 0 ? v[2] / (v[1] * 1e-3) / 1e12 : 0.0);
      }
      fclose(f);
    }
  }
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = fl;
  if (max_launch_ms) *max_launch_ms = mx;
  g_prof.used = 0;
  g_prof.flops.clear();
  g_prof.shapes.clear();
  return (long long)n;
}

int pick_block_n_public(int N, int b_mn) { return pick_block_n(N, b_mn); }

int launch_gemm(const iadr1_gemm_t& d, cudaStream_t stream) {
  if (d.M <= 0 || d.N <= 0 || d.batch <= 0) return 0;
  if (d.K <= 0) return set_error("gemm: K must be positive");
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M = d.M; g.N = d.N; g.K = d.K;
  g.batch = d.batch;
  g.batch_lo = d.batch_lo > 0 ? d.batch_lo : d.batch;
  g.b_lo_div = d.b_lo_div > 0 ? d.b_lo_div : 1;
  g.a_mn = d.a_mn; g.b_mn = d.b_mn;
  g.kmode = d.kmode; g.skip_mode = d.skip_mode; g.causal_off = d.causal_off;
  g.split_k = d.split_k > 0 ? d.split_k : 1;
  g.epi = d.epi;
  g.c_f32 = d.c_f32; g.trans_c = d.trans_c; g.accumulate = d.accumulate; g.atomic = d.atomic;
  g.bias_per_m = d.bias_per_m;
  g.a_static = d.a_static;
  g.alpha = d.alpha;
  g.C = d.C; g.ldc = d.ldc; g.c_bs_lo = d.c_bs_lo; g.c_bs_hi = d.c_bs_hi;
  g.bias = reinterpret_cast<const __nv_bfloat16*>(d.bias);
  g.residual = reinterpret_cast<const __nv_bfloat16*>(d.residual);
  g.labels = d.labels; g.part_max = d.part_max; g.part_sum = d.part_sum; g.tgt_logit = d.tgt_logit;
  g.lse = d.lse; g.gscale = d.gscale;
  // Tile order by a DRAM-traffic estimate: with M fastest a wave of CTAs shares B tiles and walks A, so A is re-read once
  // per group of N columns unless it stays L2-resident; with N fastest the roles swap. (ncu, MLP dgrad / wgrad with M
  // fastest: 2.2 / 3.4 GB of DRAM reads for 0.5 / 0.6 GB of operands.)
  g.raster_n = 0;
  if (d.raster == 2) {
    g.raster_n = 1;
  } else if (d.raster == 0 && g.batch == 1 && g.kmode == 0 && g.skip_mode == 0 && !d.stream_k) {
    const int bn = d.block_n > 0 ? d.block_n : pick_block_n(d.N, d.b_mn);
    const double tm = (d.M + BM - 1) / BM, tn = (d.N + bn - 1) / bn, wave = num_sms();
    const double a_bytes_tot = 2.0 * d.M * d.K, b_bytes_tot = 2.0 * d.N * d.K, l2_keep = 60e6;
    auto passes = [&](double t_other, double t_self) {   // sweeps over the re-read operand
      double cols = wave / t_self;                        // columns of the other dim covered by one wave
      if (cols < 1) cols = 1;
      return (double)(long long)((t_other + (long long)cols - 1) / (long long)cols);
    };
    const double cost_m = b_bytes_tot + a_bytes_tot * (a_bytes_tot <= l2_keep ? 1.0 : passes(tn, tm));
    const double cost_n = a_bytes_tot + b_bytes_tot * (b_bytes_tot <= l2_keep ? 1.0 : passes(tm, tn));
    g.raster_n = cost_n < 0.8 * cost_m;
  }
  g.up_row_off = d.up_row_off;
  if (g.epi == EPI_SWIGLU) {
    // d.M = 2 * I rows of the fused gate|up weight; tile t covers features [64 t, 64 t + 64) of both halves
    if (g.a_mn || g.b_mn || g.batch != 1 || g.split_k != 1 || g.atomic || g.c_f32 || d.stream_k || g.kmode || g.skip_mode ||
        (d.M % 128) || d.up_row_off * 2 != d.M || (d.ldc % 2))
      return set_error("gemm: swiglu epilogue needs K-major operands, bf16 C, M = 2 * up_row_off, M %% 128 == 0");
  }
  if (g.epi == EPI_SWIGLU_T) {
    // d.N = I features; B = fused gate|up weight [2I][K]; C = act [M][I]; d.gu_out (optional) [M][2I]
    if (g.a_mn || g.b_mn || g.batch != 1 || g.split_k != 1 || g.atomic || g.c_f32 || g.trans_c || d.stream_k || g.kmode || g.skip_mode ||
        (d.N % 128) || (d.ldc % 8) || (d.gu_out && (d.gu_ld % 8)) || d.bias || d.residual)
      return set_error("gemm: training swiglu epilogue needs K-major operands, bf16 row-major C, I %% 128 == 0, no bias / residual");
    g.N = 2 * d.N;                // accumulator columns: 128 gate + 128 up per tile
    g.gu_out = reinterpret_cast<__nv_bfloat16*>(d.gu_out);
    g.gu_ld = d.gu_ld;
  }
  if (g.epi == EPI_SWIGLU_BWD) {
    // d.N = I features of dact = dy @ W_down; d.gu_out = gate | up [M][2I] (ld gu_ld), rewritten in place with dgate | dup
    if (g.batch != 1 || g.split_k != 1 || g.atomic || g.c_f32 || g.trans_c || d.stream_k || g.kmode || g.skip_mode || (d.N % 8) ||
        !d.gu_out || (d.gu_ld % 8) || (reinterpret_cast<uintptr_t>(d.gu_out) & 15) || d.bias || d.residual || d.co_resident)
      return set_error("gemm: swiglu backward epilogue needs one plain problem, I %% 8 == 0, a 16-byte aligned gate | up buffer");
    g.gu_out = reinterpret_cast<__nv_bfloat16*>(d.gu_out);
    g.gu_ld = d.gu_ld;
  }
  // Weight-gradient form (fp32 accumulate into C, one problem): when whole 256-column tiles leave the last wave mostly idle
  // (2560 x 2048: 160 tiles on 148 SMs; the vision tower's 1280 x 1280: 50 tiles), cut the k-block units evenly over the SMs
  // instead (stream-K). The partial tiles are ADDED into C by the TMA unit (reduce-add epilogue below), so the SM issues no
  // atomics; the order of the two or three partial sums of a tile depends on which CTA gets there first.
  static const int auto_sk_on = [] { const char* e = getenv("IADR1_GEMM_WGRAD_STREAMK"); return e && e[0] >= '0' && e[0] <= '9' ? e[0] - '0' : 1; }();
  bool auto_sk = false;
  if (auto_sk_on && !d.stream_k && g.epi == EPI_STORE && g.c_f32 && g.accumulate && !g.atomic && !g.trans_c && g.batch == 1 &&
      g.split_k == 1 && g.kmode == 0 && !g.skip_mode && d.block_n <= 0 && !d.co_resident && !d.bias && !d.residual) {
    const int bn = pick_block_n(d.N, d.b_mn);
    const long long tiles = (long long)((d.M + BM - 1) / BM) * ((d.N + bn - 1) / bn), sms = num_sms();
    const long long waves = (tiles + sms - 1) / sms, nkb = (d.K + BK - 1) / BK;
    const long long pct = auto_sk_on == 1 ? 80 : (auto_sk_on == 2 ? 101 : 10 * auto_sk_on + 60);   // probes: 3..9 -> 90..150 %
    auto_sk = tiles * 100 < waves * sms * pct && tiles * nkb >= sms * 8;
  }
  g.stream_k = d.stream_k || auto_sk;
  if (g.stream_k && !((g.atomic || auto_sk) && g.c_f32 && g.batch == 1 && g.kmode == 0 && g.skip_mode == 0 && g.split_k == 1 &&
                      g.epi == EPI_STORE))
    return set_error("gemm: stream_k needs a single full-K problem with atomic f32 output and split_k == 1");
  if (g.split_k > 1 && !(g.atomic && g.c_f32)) return set_error("gemm: split_k > 1 needs atomic f32 output");
  if (g.atomic && !g.c_f32) return set_error("gemm: atomic output must be f32");
  if (g.batch % g.batch_lo) return set_error("gemm: batch must be a multiple of batch_lo");
  g.block_n = d.block_n > 0 ? d.block_n : pick_block_n(d.N, d.b_mn);
  if (g.epi == EPI_SWIGLU_T) g.block_n = 256;
  if (g.epi == EPI_SWIGLU_BWD) g.block_n = 256;   // the epilogue is unrolled over eight 32-column chunks
  static const bool wave_tiles = [] { const char* e = getenv("IADR1_GEMM_WAVE_TILES"); return !(e && e[0] == '0'); }();
  if (wave_tiles && d.block_n <= 0 && g.b_mn && g.epi == EPI_STORE && g.batch == 1 && g.split_k == 1 && !g.stream_k &&
      g.kmode == 0 && !g.skip_mode && d.N > 256) {
    // Wave quantisation of the gradient products: the persistent CTAs take tiles round-robin, so a product costs
    // ceil(tiles / SMs) tile times. Narrower tiles are NOT proportionally cheaper (A is re-staged per tile and the L2 -> SM
    // path is already near its limit at 256 columns): measured tile times relative to 256 columns are 0.92 (192) and 0.85
    // (128) (tools/gemm_tile_probe.py). 2560 x 2048 (qkv weight gradient): 160 tiles of 256 = two waves, 220 tiles of 192 =
    // two cheaper waves (1.09x measured); the vision tower's 1280 x 1280: 50 tiles on 148 SMs -> 100 tiles of 128 (1.15x);
    // 4608 x 3584 (7B qkv) stays at 256 (7 waves of 128 would lose 30 %).
    const long long tm = (d.M + BM - 1) / BM, sms = num_sms();
    auto cost = [&](int bn, int rel) { return ((tm * ((d.N + bn - 1) / bn) + sms - 1) / sms) * rel; };
    long long best_cost = cost(g.block_n, 100);
    const int cand[2][2] = {{192, 92}, {128, 85}};
    for (auto& c : cand)
      if (cost(c[0], c[1]) * 100 < best_cost * 95) {
        g.block_n = c[0];
        best_cost = cost(c[0], c[1]);
      }
  }
  if (g.block_n % 16 || g.block_n > 256 || g.block_n < 16) return set_error("gemm: bad block_n %d", g.block_n);
  if (g.b_mn && (g.block_n % 64)) return set_error("gemm: MN-major B needs block_n %% 64 == 0");
  if (g.epi == EPI_LSE && d.lse_tiles_n != (d.N + g.block_n - 1) / g.block_n)
    return set_error("gemm: lse partial buffer sized for %d tiles, kernel uses %d", d.lse_tiles_n,
                     (d.N + g.block_n - 1) / g.block_n);

  const int a_bytes = BM * BK * 2;
  const int b_bytes = g.block_n * BK * 2;
  const int stage_bytes = a_bytes + ((b_bytes + 1023) & ~1023);
  // TMEM: two accumulator stages of block_n fp32 columns; the co-resident variant allocates only what it needs
  g.tmem_cols = 512;
  bool co_resident = d.co_resident != 0;
  {
    // two CTAs per SM only when a >= 3-stage ring (plus the epilogue staging tile) fits in half the shared memory
    g.swiglu_rows = g.block_n < 64 ? g.block_n : 64;
    int epi_b = g.epi == EPI_SWIGLU ? g.swiglu_rows * BM * 4 : ((g.trans_c && g.atomic && g.c_f32) ? g.block_n * BM * 4 : 0);
    const int stage_b = BM * BK * 2 + ((g.block_n * BK * 2 + 1023) & ~1023);
    // needs a >= 3-stage ring per CTA (measured: 2 stages x 2 CTAs is slower than 6 stages x 1 CTA at block_n = 128)
    if (g.block_n > 128 || (113 * 1024 - 1024 - 512 - epi_b) / stage_b < 3) {
      // SwiGLU at 128 decode rows: 3 x 32 KiB of ring + a 32-row staging tile + barriers is 115200 bytes - it fits the
      // 115712 bytes two resident CTAs can have only without the 1 KiB alignment slack (172 tiles on 148 SMs: one CTA per
      // SM means two waves, two per SM stream all tiles concurrently)
      if (co_resident && g.epi == EPI_SWIGLU && g.block_n <= 128 && 3 * stage_b + 32 * BM * 4 + 512 <= 113 * 1024) {
        g.swiglu_rows = 32;
        g.smem_tight = 1;
      } else {
        co_resident = false;
      }
    }
  }
  if (co_resident) {
    g.tmem_cols = 32;
    while (g.tmem_cols < 2 * g.block_n) g.tmem_cols <<= 1;
  }
  // transposed fp32 atomic accumulation goes through bulk reductions (needs a [block_n][128] fp32 staging tile)
  g.bulk_red = g.trans_c && g.atomic && g.c_f32 && g.epi == EPI_STORE && g.residual == nullptr && g.batch == 1 &&
               (g.M % 4 == 0) && (g.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) && !d.no_bulk_red;
  // Plain stores and fp32 accumulation of the training products go through a tensor map of C (see the epilogue): needs a plain
  // row-major single problem, 16-byte aligned rows, tiles that split into whole 128-byte chunks, and no residual operand
  // (that one is still read row by row). IADR1_GEMM_TMA_STORE=0 restores the per-thread stores.
  static const bool tma_store_on = [] { const char* e = getenv("IADR1_GEMM_TMA_STORE"); return !(e && e[0] == '0'); }();
  g.tma_store = 0;
  if (tma_store_on && !co_resident && g.epi == EPI_STORE && !g.trans_c && !g.atomic && g.batch == 1 && g.split_k == 1 &&
      (!g.stream_k || auto_sk) && g.residual == nullptr && !g.bias_per_m && (g.bias == nullptr || (!g.c_f32 && (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0)) &&
      (reinterpret_cast<uintptr_t>(g.C) & 15) == 0 && ((g.ldc * (g.c_f32 ? 4 : 2)) % 16) == 0 && (g.block_n % (g.c_f32 ? 32 : 64)) == 0 &&
      (g.c_f32 || (d.N % 8) == 0) && d.M >= BM && d.N >= (g.c_f32 ? 32 : 64) && (g.c_f32 || !g.accumulate))
    g.tma_store = (g.c_f32 && g.accumulate) ? 2 : 1;
  if (auto_sk && g.tma_store != 2) g.stream_k = 0;   // no reduce-add epilogue available: whole tiles, read-modify-write
  const int epi_bytes = g.epi == EPI_SWIGLU ? g.swiglu_rows * BM * 4 : (g.bulk_red ? g.block_n * BM * 4 : (g.tma_store ? 512 + 2 * 16384 : 0));
  const int smem_slack = g.smem_tight ? 0 : 1024;
  const int smem_budget = (co_resident ? 113 : 227) * 1024 - smem_slack - 512 - epi_bytes;
  int stages = smem_budget / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (d.stages > 0 && d.stages < stages) stages = d.stages;
  if (stages < 2) return set_error("gemm: not enough shared memory for 2 stages");
  g.stages = stages;
  const size_t smem_bytes = (size_t)stages * stage_bytes + smem_slack + 512 + epi_bytes;

  // MN-major operands whose MN extent is a multiple of 64: one box of the chunked view per k-block
  g.a_chunked = g.a_mn && (d.M % 64 == 0) && !d.no_chunked_maps;
  g.b_chunked = g.b_mn && (d.N % 64 == 0) && !d.no_chunked_maps;
  const int nb_hi = g.batch / g.batch_lo;
  const int nb_lo_b = (g.batch_lo + g.b_lo_div - 1) / g.b_lo_div;
  CUtensorMap tmA, tmB;
  int rc;
  // dims: {contiguous, rows, batch_lo, batch_hi}
  if (!g.a_mn)
    rc = make_map(&tmA, d.A, d.K, d.M, g.batch_lo, nb_hi, d.lda, d.a_bs_lo ? d.a_bs_lo : d.lda,
                  d.a_bs_hi ? d.a_bs_hi : d.lda, BK, g.epi == EPI_SWIGLU ? 64 : BM);
  else if (g.a_chunked)
    rc = make_map_chunked(&tmA, d.A, d.M, d.K, g.batch_lo, nb_hi, d.lda, d.a_bs_lo ? d.a_bs_lo : d.lda,
                          d.a_bs_hi ? d.a_bs_hi : d.lda, BM / 64);
  else
    rc = make_map(&tmA, d.A, d.M, d.K, g.batch_lo, nb_hi, d.lda, d.a_bs_lo ? d.a_bs_lo : d.lda,
                  d.a_bs_hi ? d.a_bs_hi : d.lda, 64, BK);
  if (rc) return rc;
  if (g.epi == EPI_SWIGLU_T)
    rc = make_tensor_map_pair(&tmB, d.B, d.K, d.N, d.ldb, BK, 128);
  else if (!g.b_mn)
    rc = make_map(&tmB, d.B, d.K, d.N, nb_lo_b, nb_hi, d.ldb, d.b_bs_lo ? d.b_bs_lo : d.ldb,
                  d.b_bs_hi ? d.b_bs_hi : d.ldb, BK, g.block_n);
  else if (g.b_chunked)
    rc = make_map_chunked(&tmB, d.B, d.N, d.K, nb_lo_b, nb_hi, d.ldb, d.b_bs_lo ? d.b_bs_lo : d.ldb,
                          d.b_bs_hi ? d.b_bs_hi : d.ldb, g.block_n / 64);
  else
    rc = make_map(&tmB, d.B, d.N, d.K, nb_lo_b, nb_hi, d.ldb, d.b_bs_lo ? d.b_bs_lo : d.ldb,
                  d.b_bs_hi ? d.b_bs_hi : d.ldb, 64, BK);
  if (rc) return rc;
  CUtensorMap tmC = tmA;          // placeholder unless the epilogue stores through it
  if (g.tma_store) {
    rc = make_map_c(&tmC, d.C, d.N, d.M, d.ldc, g.c_f32);
    if (rc) return rc;
  }

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(gemm smem): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles_m = (g.M + BM - 1) / BM;
  const int tiles_n = (g.N + g.block_n - 1) / g.block_n;
  const long long total = (long long)tiles_m * tiles_n * g.split_k * g.batch;
  int grid = (int)(total < (long long)num_sms() ? total : num_sms());
  if (co_resident) grid = (int)(total < 2LL * num_sms() ? total : 2LL * num_sms());   // two CTAs per SM stream concurrently
  if (g.stream_k) grid = num_sms();   // every SM takes an equal run of k-block units (K / 64 >= 1 each tile)
  if (d.max_ctas > 0 && grid > d.max_ctas) grid = d.max_ctas;
  cudaEvent_t pe0, pe1;
  const bool timed = prof_begin(stream, &pe0, &pe1);
  if (co_resident)
    launch_kernel(gemm_bf16_tcgen05_kernel<2>, dim3(grid), dim3(kThreads), smem_bytes, stream, tmA, tmB, tmC, g);
  else
    launch_kernel(gemm_bf16_tcgen05_kernel<1>, dim3(grid), dim3(kThreads), smem_bytes, stream, tmA, tmB, tmC, g);
  if (timed) {
    cudaEventRecord(pe1, stream);
    // algorithmic FLOPs: causal products count only the unmasked half
    double fl = 2.0 * (double)g.M * (double)g.N * (double)g.K * (double)g.batch;
    if (g.kmode != 0 || g.skip_mode != 0) fl *= 0.5;
    std::lock_guard<std::mutex> lk(g_prof.mu);
    g_prof.flops.push_back(fl);
    g_prof.shapes.push_back({g.M, g.N, g.K, g.batch, g.a_mn, g.b_mn, g.epi, g.block_n});
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("gemm launch failed: %s", cudaGetErrorString(e));
  count_launch();
  return 0;
}

void trace_install_gemm(unsigned long long* p) { trace_install_tu(p); }
}  // namespace iadr1
