// Persistent decode-layer chain for the rollout (replaces the per-layer body of `self.llm.generate`,
// ref: train/stage_rl/trainer/sc_grpo_trainer.py:343-358, 667 - a vLLM engine in the reference).
//
// Between two attention kernels a decode step runs, per layer,
//     o-projection -> RMSNorm -> gate_up + SwiGLU -> down-projection -> RMSNorm (next layer) -> qkv-projection (next layer)
// which used to be six dependent launches whose fixed latencies (dependency resolution, first tile loads, epilogue, kernel
// drain) added up to twice the time the 154 MB of weights need to stream from HBM. This kernel runs the whole chain as ONE
// launch of one CTA per SM:
//   * warp 0 (one thread) is the TMA producer. It walks the CTA's share of every phase in a fixed order and keeps the
//     shared-memory ring full of WEIGHT tiles without ever waiting for a phase boundary; only the activation tile of a
//     k-block waits until the phase that produces it has finished grid-wide. HBM therefore stays busy across what used to
//     be kernel boundaries (two cursors over the same load sequence: weights run ahead, activations follow).
//   * warp 1 (one thread) issues tcgen05.mma into two TMEM accumulator stages; warps 4-7 run the epilogues and the row
//     phases (RMSNorm) and publish "phase done" on a global counter (red.release.gpu); consumers poll it (ld.acquire.gpu).
//   * the skinny products (o, down, qkv) keep the operand swap of the stand-alone decode GEMM: weights are the 128-row MMA
//     operand, the R <= 128 decode rows the N side, k-block units cut evenly over the CTAs (stream-K) and the fp32 tile
//     added into the residual stream / qkv buffer by cp.reduce.async.bulk. gate_up runs un-swapped (decode rows = M, one tile
//     of `ft` gate + `ft` up features = N <= 256 per CTA) so that its tile count fits ONE wave and silu(g) * u needs no
//     staging: a thread owns a decode row and writes ft contiguous bf16.
#include "ptx.cuh"
#include "runtime.h"

#include <cuda_bf16.h>
#include <cstdlib>

#define TRY_RC(expr)         \
  do {                       \
    int rc__ = (expr);       \
    if (rc__) return rc__;   \
  } while (0)

namespace iadr1 {

int make_tensor_map_2d(CUtensorMap* out, const void* ptr, long long cols, long long rows, long long ld, int box_cols, int box_rows);
int make_tensor_map_pair(CUtensorMap* out, const void* ptr, long long cols, long long rows, long long ld, int box_cols, int box_rows);

namespace {

using bf16 = __nv_bfloat16;
constexpr int CBM = 128, CBK = 64, CTHREADS = 256, CMAX_W = 12, CMAX_X = 4, CEPI_ROWS = 64, CTMEM_STAGE = 256, CX_BYTES = 16384;
enum ChainPhase : int { PH_O = 0, PH_NORM2 = 1, PH_GU = 2, PH_DOWN = 3, PH_NORM1 = 4, PH_QKV = 5, PH_COUNT = 6 };

struct ChainMaps {
  CUtensorMap w[PH_COUNT];   // weight operand of each GEMM phase (row-major [features][K])
  CUtensorMap x[PH_COUNT];   // activation operand of each GEMM phase (row-major [R][K])
};

struct ChainArgs {
  int R, H, I, QH, D;
  int ft;                 // gate (= up) features per SwiGLU tile, multiple of 16, <= 128
  int nw, wslot_bytes, nx;   // weight ring (deep: runs ahead across phases) / activation ring (shallow: L2 latency only)
  unsigned phase_mask;
  int zero_qkv;
  float eps;
  float* h;               // fp32 residual stream [R][H]
  bf16* xn;               // normalised operand [R][H]
  bf16* act;              // SwiGLU output [R][I]
  float* qkv;             // fp32 [R][D]
  const bf16* ln_mid;     // RMSNorm weight between o and gate_up
  const bf16* ln_next;    // RMSNorm weight after down (next layer's input norm, or the final norm)
  const bf16* qkv_bias;
  unsigned* counters;     // PH_COUNT words, zero on entry
};

__device__ __forceinline__ bool is_gemm(int p) { return p == PH_O || p == PH_GU || p == PH_DOWN || p == PH_QKV; }
__device__ __forceinline__ bool enabled(const ChainArgs& a, int p) { return (a.phase_mask >> p) & 1u; }
__device__ __forceinline__ int dep_of(const ChainArgs& a, int p) {
  for (int q = p - 1; q >= 0; --q)
    if (enabled(a, q)) return q;
  return -1;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }

struct Seg { int m0, kb0, kb1; };

// This CTA's segments of one GEMM phase. Swapped products: every 128-feature tile is cut into `splits` k-ranges, one
// (tile, k-range) per CTA - exactly ONE epilogue (staging + bulk reduction + publish) per CTA; equal contiguous runs of
// k-block units (stream-K) balance the bytes a little better but make most CTAs cross a tile boundary and pay two. Only when
// there are more tiles than CTAs do tiles go round-robin with the whole K. gate_up: whole-K tiles, round-robin.
struct PhaseIter {
  int swiglu, nkb, cur, end, step, ft, per, splits;
  __device__ __forceinline__ void init(const ChainArgs& a, int p) {
    swiglu = (p == PH_GU);
    const int K = (p == PH_O) ? a.QH : (p == PH_DOWN ? a.I : a.H);
    nkb = (K + CBK - 1) / CBK;
    ft = a.ft;
    if (!swiglu) {
      const int F = (p == PH_QKV) ? a.D : a.H;
      const int tiles = (F + CBM - 1) / CBM;
      splits = max(1, min((int)gridDim.x / tiles, nkb));
      per = (nkb + splits - 1) / splits;
      splits = (nkb + per - 1) / per;               // drop empty k-ranges
      cur = blockIdx.x;
      end = tiles * splits;
      step = gridDim.x;
    } else {
      cur = blockIdx.x;
      end = (a.I + a.ft - 1) / a.ft;
      step = gridDim.x;
    }
  }
  __device__ __forceinline__ bool next(Seg& s) {
    if (cur >= end) return false;
    if (!swiglu) {
      const int tile = cur / splits, sp = cur - tile * splits;
      s.m0 = tile * CBM; s.kb0 = sp * per; s.kb1 = min(nkb, s.kb0 + per);
    } else {
      s.m0 = cur * ft; s.kb0 = 0; s.kb1 = nkb;
    }
    cur += step;
    return true;
  }
};

// Position in the CTA's flattened sequence of k-block loads (all enabled GEMM phases, in order).
struct LoadCursor {
  int p, kb;
  bool done;
  PhaseIter it;
  Seg s;
  __device__ __forceinline__ void next_segment(const ChainArgs& a) {
    while (true) {
      if (p >= 0 && it.next(s)) { kb = s.kb0; return; }
      do { ++p; } while (p < PH_COUNT && !(enabled(a, p) && is_gemm(p)));
      if (p >= PH_COUNT) { done = true; return; }
      it.init(a, p);
    }
  }
  __device__ __forceinline__ void start(const ChainArgs& a) { p = -1; kb = 0; done = false; next_segment(a); }
  __device__ __forceinline__ void advance(const ChainArgs& a) { if (++kb >= s.kb1) next_segment(a); }
};

// spin with a bound: a chain that cannot make progress (a CTA that never became resident) traps instead of hanging the GPU
__device__ __forceinline__ void wait_counter(const unsigned* c, unsigned target) {
  unsigned long long spins = 0;
  while (ld_acquire(c) < target) {
    if (++spins > (1ull << 21)) __trap();   // ~1 s of polling: a legitimate wait is < 1 ms
  }
}

__global__ void __launch_bounds__(CTHREADS, 1)
decode_chain_kernel(const __grid_constant__ ChainMaps maps, const ChainArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  if (threadIdx.x == 0) pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = a.nw, nx = a.nx, wslot = a.wslot_bytes;
  uint8_t* ring_w = smem;
  uint8_t* ring_x = smem + (size_t)nw * wslot;
  float* sC = reinterpret_cast<float*>(ring_x + (size_t)nx * CX_BYTES);                  // [CEPI_ROWS][128] fp32 staging
  uint64_t* wfull = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sC) + CEPI_ROWS * CBM * 4);
  uint64_t* wempty = wfull + CMAX_W;
  uint64_t* xfull = wempty + CMAX_W;
  uint64_t* xempty = xfull + CMAX_X;
  uint64_t* tfull_bar = xempty + CMAX_X;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_red = reinterpret_cast<float*>(tmem_slot + 4);                                 // 4 floats (norm reduction)
  const unsigned nctas = gridDim.x;

  if (warp == 0 && lane == 0) {
    for (int p = 0; p < PH_COUNT; ++p)
      if (enabled(a, p) && is_gemm(p)) {
        tma_prefetch_desc(&maps.w[p]);
        tma_prefetch_desc(&maps.x[p]);
      }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < nw; ++i) {
      mbar_init(&wfull[i], 1);
      mbar_init(&wempty[i], 1);
    }
    for (int i = 0; i < nx; ++i) {
      mbar_init(&xfull[i], 1);
      mbar_init(&xempty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 2 * CTMEM_STAGE);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 3) {
    // ===================== TMA producers: two independent streams, one thread each (warps 3 and 0) =====================
    // Weight tiles go through a deep ring and are requested as soon as a slot is free - across phase boundaries, so HBM keeps
    // streaming while a phase drains; activation tiles go through a shallow ring and wait for the phase that produces them.
    // A lone thread issues ~300-450 cycles of mbarrier + TMA instructions per tile (measured with clock64), i.e. one thread
    // feeding both rings cannot keep up with HBM: each stream has its own thread; slots / parities are carried incrementally.
    if (lane == 0) {
      // ---- weight stream (warp 3): never waits for a phase
      LoadCursor cw;
      cw.start(a);
      int sw = 0;
      uint32_t pw = 1;                     // parity of the "slot free" phase the next tile waits for
      const uint32_t gu_bytes = 2u * (uint32_t)a.ft * 128u;
      while (!cw.done) {
        mbar_wait(&wempty[sw], pw);
        uint8_t* st = ring_w + (size_t)sw * wslot;
        mbar_arrive_expect_tx(&wfull[sw], cw.p == PH_GU ? gu_bytes : 16384u);
        tma_load_4d(st, &maps.w[cw.p], &wfull[sw], cw.kb * CBK, cw.s.m0, 0, 0);
        cw.advance(a);
        if (++sw == nw) { sw = 0; pw ^= 1u; }
      }
    }
  } else if (warp == 0) {
    if (lane == 0) {
      // ---- activation stream (warp 0): a tile is requested once the phase producing it has finished grid-wide
      LoadCursor cx;
      cx.start(a);
      int sx = 0, ready_upto = -1;
      uint32_t px = 1;
      if (blockIdx.x == 0) trace_stamp(100);
      pdl_wait();                          // every activation of the chain follows the preceding kernel
      while (!cx.done) {
        if (cx.p > ready_upto) {
          const int d = dep_of(a, cx.p);
          if (d >= 0) {
            wait_counter(a.counters + d, nctas);
            fence_proxy_async_all();
          }
          if (blockIdx.x == 0) trace_stamp(110 + cx.p);
          ready_upto = cx.p;
        }
        mbar_wait(&xempty[sx], px);
        mbar_arrive_expect_tx(&xfull[sx], (uint32_t)CX_BYTES);
        tma_load_4d(ring_x + (size_t)sx * CX_BYTES, &maps.x[cx.p], &xfull[sx], cx.kb * CBK, 0, 0, 0);
        cx.advance(a);
        if (++sx == nx) { sx = 0; px ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int sw = 0, sx = 0, as = 0;
      uint32_t aph = 0, wph = 0, xph = 0;
      for (int p = 0; p < PH_COUNT; ++p) {
        if (!(enabled(a, p) && is_gemm(p))) continue;
        const bool gu = p == PH_GU;
        const uint32_t idesc = make_idesc_bf16(CBM, gu ? 2 * a.ft : CBM, 0, 0);
        PhaseIter it;
        it.init(a, p);
        Seg sg;
        while (it.next(sg)) {
          mbar_wait(&tempty_bar[as], aph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * CTMEM_STAGE;
          for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
            mbar_wait(&wfull[sw], wph);
            mbar_wait(&xfull[sx], xph);
            tc_fence_after();
            const uint32_t wa = smem_u32(ring_w + (size_t)sw * wslot), xa = smem_u32(ring_x + (size_t)sx * CX_BYTES);
            // swapped products: weights are the 128-row A operand, decode rows the N side; gate_up: decode rows are A
            const uint64_t adesc = make_smem_desc_sw128(gu ? xa : wa, 16, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(gu ? wa : xa, 16, 1024);
#pragma unroll
            for (int k = 0; k < CBK / 16; ++k)
              umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb > sg.kb0 || k > 0) ? 1u : 0u);
            umma_commit(&wempty[sw]);
            umma_commit(&xempty[sx]);
            if (++sw == nw) { sw = 0; wph ^= 1u; }
            if (++sx == nx) { sx = 0; xph ^= 1u; }
          }
          umma_commit(&tfull_bar[as]);
          if (++as == 2) { as = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogues + row phases =====================
    const int q = warp & 3, et = threadIdx.x - 128, ml = q * 32 + lane;
    int as = 0;
    uint32_t aph = 0;
    bool waited = false;
    for (int p = 0; p < PH_COUNT; ++p) {
      if (!enabled(a, p)) continue;
      if (!is_gemm(p)) {
        // ---- RMSNorm of the fp32 residual stream -> bf16 operand (Qwen2RMSNorm rounding order), rows round-robin over CTAs
        const bf16* w = (p == PH_NORM2) ? a.ln_mid : a.ln_next;
        const int H = a.H;
        constexpr int MAXV = 8;                   // 128 threads x 8 x float4 = 4096 columns in registers
        uint2 wr[MAXV];                           // static weights: requested before the dependency wait
        if ((int)blockIdx.x < a.R) {
#pragma unroll
          for (int i = 0; i < MAXV; ++i) {
            const int c = (et + i * 128) * 4;
            if (c < H) wr[i] = *reinterpret_cast<const uint2*>(w + c);
          }
        }
        const int d = dep_of(a, p);
        if (d < 0) {
          if (!waited) { pdl_wait(); waited = true; }
        } else {
          if (et == 0) {
            wait_counter(a.counters + d, nctas);
            if (blockIdx.x == 0) trace_stamp(110 + p);
          }
          named_bar_sync(1, 128);
        }
        for (int r = blockIdx.x; r < a.R; r += gridDim.x) {
          const float* xr = a.h + (long long)r * H;
          float4 xv[MAXV];
          float ss = 0.f;
#pragma unroll
          for (int i = 0; i < MAXV; ++i) {
            const int c = (et + i * 128) * 4;
            if (c < H) {
              float4 v = __ldcg(reinterpret_cast<const float4*>(xr + c));
              v.x = bf16r(v.x); v.y = bf16r(v.y); v.z = bf16r(v.z); v.w = bf16r(v.w);
              xv[i] = v;
              ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
          if (lane == 0) s_red[q] = ss;
          named_bar_sync(1, 128);
          const float rstd = rsqrtf((s_red[0] + s_red[1] + s_red[2] + s_red[3]) / (float)H + a.eps);
          named_bar_sync(1, 128);                 // s_red may be rewritten by the next row
#pragma unroll
          for (int i = 0; i < MAXV; ++i) {
            const int c = (et + i * 128) * 4;
            if (c < H) {
              const float2 w0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wr[i].x));
              const float2 w1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wr[i].y));
              uint2 o;
              *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(w0.x * bf16r(xv[i].x * rstd), w0.y * bf16r(xv[i].y * rstd));
              *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(w1.x * bf16r(xv[i].z * rstd), w1.y * bf16r(xv[i].w * rstd));
              *reinterpret_cast<uint2*>(a.xn + (long long)r * H + c) = o;
            }
          }
          if (p == PH_NORM1 && a.zero_qkv)        // the qkv product adds its split-K partials into a cleared buffer
            for (int c = et * 4; c < a.D; c += 512)
              *reinterpret_cast<float4*>(a.qkv + (long long)r * a.D + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
        if (!waited) { pdl_wait(); waited = true; }   // first global write of this CTA must follow the preceding kernel
        PhaseIter it;
        it.init(a, p);
        Seg sg;
        while (it.next(sg)) {
          // one lane polls (with back-off), the rest of the warp waits on it: 128 threads spinning on try_wait for the whole
          // k-loop of a tile slow the TMA / mbarrier traffic of the producer down
          if (lane == 0) {
            while (!mbar_try_wait(&tfull_bar[as], aph)) __nanosleep(256);
          }
          __syncwarp();
          mbar_wait(&tfull_bar[as], aph);
          tc_fence_after();
          const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + as * CTMEM_STAGE;
          if (p == PH_GU) {
            // D[decode row][gate ft | up ft]: the thread owns row `ml`
            const int f0 = sg.m0;
            bf16* orow = a.act + (long long)ml * a.I + f0;
            for (int c0 = 0; c0 < a.ft; c0 += 16) {
              uint32_t gv[16], uv[16];
              tmem_ld_32x32b_x16(taddr + c0, gv);
              tmem_ld_32x32b_x16(taddr + a.ft + c0, uv);
              tmem_ld_wait();
              uint32_t o[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float r2[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const float g = bf16r(__uint_as_float(gv[2 * j + e]));
                  const float sv = bf16r(silu_f(g));
                  r2[e] = sv * bf16r(__uint_as_float(uv[2 * j + e]));
                }
                const __nv_bfloat162 pk = __floats2bfloat162_rn(r2[0], r2[1]);
                o[j] = *reinterpret_cast<const uint32_t*>(&pk);
              }
              if (ml < a.R) {
                if (f0 + c0 + 16 <= a.I) {
                  *reinterpret_cast<uint4*>(orow + c0) = make_uint4(o[0], o[1], o[2], o[3]);
                  *reinterpret_cast<uint4*>(orow + c0 + 8) = make_uint4(o[4], o[5], o[6], o[7]);
                } else {
                  for (int j = 0; j < 16; ++j)
                    if (f0 + c0 + j < a.I)
                      orow[c0 + j] = __ushort_as_bfloat16((unsigned short)((o[j >> 1] >> ((j & 1) * 16)) & 0xffffu));
                }
              }
            }
            tc_fence_before();
            mbar_arrive(&tempty_bar[as]);
          } else {
            // D[feature][decode row] -> out[row][m0 .. m0 + 128) += tile, one bulk reduction per decode row
            const int F = (p == PH_QKV) ? a.D : a.H;
            float* out = (p == PH_QKV) ? a.qkv : a.h;
            const int m = sg.m0 + ml;
            float bias_m = 0.f;
            if (p == PH_QKV && a.qkv_bias != nullptr && sg.kb0 == 0 && m < F) bias_m = __bfloat162float(a.qkv_bias[m]);
            const int mw = min(CBM, F - sg.m0);
            for (int r0 = 0; r0 < a.R; r0 += CEPI_ROWS) {
              const int nrows = min(CEPI_ROWS, a.R - r0);
              for (int c0 = 0; c0 < nrows; c0 += 16) {
                uint32_t v[16];
                tmem_ld_32x32b_x16(taddr + r0 + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) sC[(c0 + j) * CBM + ml] = __uint_as_float(v[j]) + bias_m;
              }
              if (r0 + CEPI_ROWS >= a.R) {        // accumulator fully read: hand the TMEM stage back
                tc_fence_before();
                mbar_arrive(&tempty_bar[as]);
              }
              fence_proxy_async_smem();
              named_bar_sync(1, 128);
              for (int n = et; n < nrows; n += 128)
                bulk_reduce_add_f32(out + (long long)(r0 + n) * F + sg.m0, sC + n * CBM, (uint32_t)mw * 4u);
              bulk_commit();
              bulk_wait_read0();
              named_bar_sync(1, 128);
            }
          }
          if (++as == 2) { as = 0; aph ^= 1; }
        }
      }
      // ---- publish: this CTA's part of phase p is in global memory
      if (is_gemm(p) && p != PH_GU) bulk_wait0();   // this thread's bulk reductions have been performed
      else fence_proxy_async_all();                 // this thread's generic stores (xn / act) before other CTAs' TMA reads
      named_bar_sync(1, 128);
      if (et == 0) {
        red_release_add(a.counters + p, 1u);
        if (blockIdx.x == 0) trace_stamp(120 + p);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * CTMEM_STAGE);
  }
}

int g_chain_sms = 0;

}  // namespace

bool decode_chain_enabled() {
  const char* v = getenv("IADR1_DECODE_CHAIN");     // read per call: tests compare both forms in one process
  return !(v && atoi(v) == 0);
}

// One chain launch. Phases: with_mlp = o -> norm -> gate_up -> down on (attn, Wo, ln_mid, Wgu, Wd); with_norm = the RMSNorm
// after it with ln_next; with_qkv = the qkv product of the next layer (needs with_norm). counters: PH_COUNT zeroed words.
int launch_decode_chain(int R, int H, int I, int QH, int D, float eps, float* h, void* xn, void* act, float* qkv, const void* attn,
                        const void* Wo, const void* ln_mid, const void* Wgu, const void* Wd, const void* ln_next, const void* Wqkv,
                        const void* qkv_bias, int with_mlp, int with_norm, int with_qkv, unsigned* counters, cudaStream_t stream) {
  if (R <= 0) return 0;
  if (R > 128) return set_error("decode_chain: at most 128 rows (got %d)", R);
  if ((H % 8) || (I % 8) || (QH % 8) || (D % 4) || H > 4096) return set_error("decode_chain: unsupported widths H=%d I=%d QH=%d D=%d", H, I, QH, D);
  if (with_qkv && !with_norm) return set_error("decode_chain: the qkv phase needs the norm phase before it");
  if (g_chain_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_chain_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_chain_sms <= 0) g_chain_sms = 148;
  }
  const int grid = g_chain_sms;
  ChainArgs a;
  memset(&a, 0, sizeof(a));
  a.R = R; a.H = H; a.I = I; a.QH = QH; a.D = D;
  int ft = ((I + grid - 1) / grid + 15) & ~15;
  if (ft > 128) ft = 128;
  a.ft = ft;
  a.eps = eps;
  a.h = h; a.xn = static_cast<bf16*>(xn); a.act = static_cast<bf16*>(act); a.qkv = qkv;
  a.ln_mid = static_cast<const bf16*>(ln_mid); a.ln_next = static_cast<const bf16*>(ln_next);
  a.qkv_bias = static_cast<const bf16*>(qkv_bias);
  a.counters = counters;
  a.zero_qkv = with_qkv;
  a.phase_mask = (with_mlp ? 0xFu : 0u) | (with_norm ? (1u << PH_NORM1) : 0u) | (with_qkv ? (1u << PH_QKV) : 0u);
  if (getenv("IADR1_CHAIN_DEBUG_NOX")) a.phase_mask |= 1u << 29;
  if (getenv("IADR1_CHAIN_DEBUG_NOW")) a.phase_mask |= 1u << 28;
  if (getenv("IADR1_CHAIN_DEBUG_NOMMA")) a.phase_mask |= 1u << 31;   // timing experiment only (wrong results)
  a.wslot_bytes = with_mlp ? ((2 * ft * 128 + 1023) & ~1023) : 16384;
  if (a.wslot_bytes < 16384) a.wslot_bytes = 16384;
  a.nx = 3;
  const int fixed = 1024 + CEPI_ROWS * CBM * 4 + 512 + a.nx * CX_BYTES;
  int nw = (227 * 1024 - fixed) / a.wslot_bytes;
  if (nw > CMAX_W) nw = CMAX_W;
  if (getenv("IADR1_CHAIN_STAGES") && atoi(getenv("IADR1_CHAIN_STAGES")) >= 2 && atoi(getenv("IADR1_CHAIN_STAGES")) < nw)
    nw = atoi(getenv("IADR1_CHAIN_STAGES"));      // experiments only
  if (nw < 2) return set_error("decode_chain: not enough shared memory");
  a.nw = nw;
  const size_t smem_bytes = (size_t)nw * a.wslot_bytes + fixed;
  ChainMaps maps;
  memset(&maps, 0, sizeof(maps));
  if (with_mlp) {
    TRY_RC(make_tensor_map_2d(&maps.w[PH_O], Wo, QH, H, QH, CBK, CBM));
    TRY_RC(make_tensor_map_2d(&maps.x[PH_O], attn, QH, R, QH, CBK, CBM));
    TRY_RC(make_tensor_map_pair(&maps.w[PH_GU], Wgu, H, I, H, CBK, ft));
    TRY_RC(make_tensor_map_2d(&maps.x[PH_GU], xn, H, R, H, CBK, CBM));
    TRY_RC(make_tensor_map_2d(&maps.w[PH_DOWN], Wd, I, H, I, CBK, CBM));
    TRY_RC(make_tensor_map_2d(&maps.x[PH_DOWN], act, I, R, I, CBK, CBM));
  }
  if (with_qkv) {
    TRY_RC(make_tensor_map_2d(&maps.w[PH_QKV], Wqkv, H, D, H, CBK, CBM));
    TRY_RC(make_tensor_map_2d(&maps.x[PH_QKV], xn, H, R, H, CBK, CBM));
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(decode_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(decode_chain smem): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  launch_kernel(decode_chain_kernel, dim3(grid), dim3(CTHREADS), smem_bytes, stream, maps, a);
  IADR1_CHECK_LAUNCH("decode_chain");
  return 0;
}

void trace_install_chain(unsigned long long* p) { trace_install_tu(p); }

}  // namespace iadr1
