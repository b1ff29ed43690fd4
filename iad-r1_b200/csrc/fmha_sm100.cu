// Fused attention for sm_100a: forward + backward with the score matrix kept on chip.
//
//   forward :  S = Q K^T (tcgen05.mma -> TMEM)  ->  online softmax in registers (tcgen05.ld, exp2, running max / sum,
//              lazy rescale of the TMEM-resident O accumulator)  ->  P (bf16) through 128B-swizzled shared memory as the A
//              operand of  O += P V (tcgen05.mma, V read MN-major straight from the fused qkv buffer).
//              Stores O (bf16) and the per-row log-sum-exp (log2 domain) only.
//   backward:  two kernels that recompute P tile by tile from (Q, K, lse):
//              dQ kernel  (one CTA per 128-query tile x head, 64-key tiles):  S, dP = dO V^T in TMEM, dS = P o (dP - delta)
//                         in registers, dQ += dS K accumulated in TMEM;
//              dKV kernel (one CTA per 128-key tile x kv head, 64-query tiles, the g query heads of the kv head looped
//                         inside): S^T = K Q^T, dP^T = V dO^T, dV += P^T dO, dK += dS^T Q accumulated in TMEM over ALL
//                         query tiles and heads of the group (GQA reduction in the CTA); partial key tiles of long
//                         query ranges (the shared prompt of a GRPO group) are split over CTAs and added with
//                         cp.reduce.async.bulk into an fp32 buffer.
// Masking is general: every query row carries two key ranges [lo, hi) u [plo, phi) (token indices into the same qkv
// buffer): causal rows, the shared-prefix GRPO layout (prompt once + G completion rows), vision windows / crops / full
// images and several packed groups are all the same kernel. Work items (query tiles / key tiles) are built on the host.
//
// Replaces the flash-attn call the reference selects with `--attn_implementation flash_attention_2`
// (ref: scripts/train/SC_GRPO/SC_GRPO_Qwen_Instruct_2_5_VL_3B.sh:58; HF modeling_qwen2_5_vl.py:214-286 vision attention,
// :704-760 decoder attention) and its autograd backward. The reference holds no native code to follow.
#include "ptx.cuh"
#include "runtime.h"

#include <map>
#include <mutex>
#include <tuple>

namespace iadr1 {

typedef __nv_bfloat16 bf16;

struct FmhaArgs {
  const int4* rng;       // [>= N + 64] per query row: keys [x, y) u [z, w)
  const int* items;      // forward / dQ: 6 ints per item {q0, nrows, kv0, kv1, p0, p1}; dKV: 4 ints {k0, nkeys, q0, q1}
  const int* sched;      // [n_cta + 1] offsets, then the unit ids (item * heads + head) each CTA walks
  int n_items;
  int nq, nkv, hd;
  int ksteps;            // ceil(hd / 16): UMMA K steps over the head dimension (columns >= hd are TMA zero fill)
  int nch;               // 64-column chunks per operand tile
  int pv_n;              // N of the products whose free dimension is the head dimension (multiple of 16, >= hd)
  float sl2;             // softmax scale * log2(e)
  float scale;
  bf16* out;             // forward: O [N][out_ld]
  long long out_ld;
  float* lse2;           // [nq][npad] (head-major) log2-domain log-sum-exp of the scaled scores
  const float* delta;    // backward: [nq][npad] rowsum(dO o O)
  long long npad;        // row stride of lse2 / delta: a multiple of 4 >= N + 64 (64-query bulk copies stay in bounds)
  bf16* dq;              // backward: dQ written at dq[tok * dq_ld + head * hd + d]
  long long dq_ld;
  float* dkv32;          // backward: [N][2 * nkv * hd] fp32, dK at column kvh * hd, dV at (nkv + kvh) * hd
  long long* trace;      // optional timeline probe (tools/fmha_probe.py --trace): CTA 0 stamps clock64 per role and iteration
};

// trace record: buf[(role * 64 + iteration) * 8 + slot]; roles 0 producer, 1 MMA, 2 compute (warp 4 lane 0)
#define FTRACE(role, iter, slot)                                                                   \
  do {                                                                                             \
    if (a.trace != nullptr && blockIdx.x == 0 && (iter) < 64u) a.trace[((role) * 64 + (iter)) * 8 + (slot)] = clock64(); \
  } while (0)

static constexpr int kFThreads = 384;     // warp 0 TMA, warp 1 MMA, warp 2 TMEM allocator, warps 4-11 two compute warpgroups
static constexpr int kCompute = 256;      // threads of the two compute warpgroups (arrival count of their barriers)
static constexpr uint32_t kSpinLimit = 1u << 28;   // a lost barrier traps (visible error) instead of hanging the GPU

__device__ __forceinline__ void mbar_wait_safe(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > kSpinLimit) __trap();
  }
}

__device__ __forceinline__ uint4 pack8(const float* x) {
  uint4 o;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(x[2 * e], x[2 * e + 1]);
  return o;
}

// Branch-free masking: predicates are scarce (7 per thread) and serialise the unrolled column loops, integer masks do not.
// Returns v when lo <= i < hi (as integers below 2^30), -inf otherwise.
__device__ __forceinline__ float keep_if_inside(float v, int i, int lo, int hi) {
  const int outside = ((i - lo) | (hi - 1 - i)) >> 31;
  return __int_as_float((__float_as_int(v) & ~outside) | (int)(0xff800000u & (unsigned)outside));
}

// One 16-byte unit `u` (8 bf16) of row `row` inside a [rows][64] bf16 chunk stored in the 128-byte-swizzled layout TMA
// writes and tcgen05.mma reads (chunk base 1024-byte aligned).
__device__ __forceinline__ uint32_t sw128_off(int row, int u) { return (uint32_t)row * 128u + (uint32_t)((u ^ (row & 7)) << 4); }

// Dynamic shared memory rounded up to 1024 bytes WITHOUT leaving the shared address space (pointer arithmetic on the
// __shared__ array keeps LDS / STS; rounding the generic address made every access a generic LD / ST).
#define IADR1_SMEM_BASE()                                                 \
  extern __shared__ uint8_t smem_raw[];                                   \
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u)

// Host-built schedule: sched[0 .. n_cta] = offsets into the unit list that follows; every CTA walks its own units
// (longest-processing-time assignment, the units' costs are known on the host).
struct UnitWalk {
  const int* list;
  int i, end;
  __device__ __forceinline__ UnitWalk(const int* sched) {
    const int n_cta = gridDim.x;
    i = sched[blockIdx.x];
    end = sched[blockIdx.x + 1];
    list = sched + n_cta + 1;
  }
  __device__ __forceinline__ bool next(int& u) {
    if (i >= end) return false;
    u = list[i++];
    return true;
  }
};

// D[tmem] = A * B^T over the head dimension (both operands K-major, `ksteps` x 16 columns, chunks of 64 columns ACH / BCH
// bytes apart). Descriptors are built once per kernel; the k-step offsets are compile-time constants (the MMA thread is a
// single lane: rebuilding descriptors per instruction made ISSUING the dominant cost of a tile).
template <int ACH, int BCH>
__device__ __forceinline__ void umma_headdim(uint32_t tm, uint64_t ad, uint64_t bd, uint32_t idesc, int ksteps) {
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    if (ks < ksteps) {
      const uint32_t offa = ((ks >> 2) * ACH + (ks & 3) * 32) >> 4, offb = ((ks >> 2) * BCH + (ks & 3) * 32) >> 4;
      umma_bf16(tm, ad + offa, bd + offb, idesc, ks ? 1u : 0u);
    }
  }
}
// D[tmem] (+)= A * B with A K-major [128][KS * 16] (one or two 64-column chunks ACH bytes apart) and B MN-major (KS * 16
// rows of 128 bytes per 64-column chunk): the P V / dS K / P^T dO / dS^T Q products.
template <int KS, int ACH>
__device__ __forceinline__ void umma_rows(uint32_t tm, uint64_t ad, uint64_t bd, uint32_t idesc, bool first) {
#pragma unroll
  for (int kk = 0; kk < KS; ++kk) {
    const uint32_t offa = ((kk >> 2) * ACH + (kk & 3) * 32) >> 4, offb = (kk * 2048) >> 4;
    umma_bf16(tm, ad + offa, bd + offb, idesc, (first && kk == 0) ? 0u : 1u);
  }
}

struct QItem {
  int q0, nrows, kv0, kv1, p0, p1, n_pre, nt;
};
template <int BN>
__device__ __forceinline__ QItem load_qitem(const int* items, int i) {
  QItem it;
  const int* p = items + 6 * i;
  it.q0 = p[0]; it.nrows = p[1]; it.kv0 = p[2]; it.kv1 = p[3]; it.p0 = p[4]; it.p1 = p[5];
  it.n_pre = it.p1 > it.p0 ? (it.p1 - it.p0 + BN - 1) / BN : 0;
  const int n_rng = it.kv1 > it.kv0 ? (it.kv1 - it.kv0 + BN - 1) / BN : 0;
  it.nt = it.n_pre + n_rng;
  return it;
}
template <int BN>
__device__ __forceinline__ int tile_k0(const QItem& it, int j) {
  return j < it.n_pre ? it.p0 + BN * j : it.kv0 + BN * (j - it.n_pre);
}

// =====================================================================================================================
// forward
// =====================================================================================================================
// smem: Q [2 chunks][128][64] | K [2 stages][2 chunks][128][64] | V same | P [2 chunks][128][64]   (bf16, 16 KiB per chunk)
//       | row-max / row-sum exchange between the two compute warpgroups [2 buffers][2 warpgroups][128] fp32
// TMEM: S0 [0,128)  S1 [128,256)  O [256, 256 + pv_n)
// The two compute warpgroups share every query row: warpgroup h owns score columns [64 h, 64 h + 64) of each key tile
// (= chunk h of the P tile) and the 16-column pieces i = h (mod 2) of O.
__global__ void __launch_bounds__(kFThreads, 1)
fmha_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const FmhaArgs a) {
  IADR1_SMEM_BASE();
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 32768;
  uint8_t* sV = smem + 32768 + 65536;
  uint8_t* sP = smem + 32768 + 131072;
  float* sX = reinterpret_cast<float*>(smem + 196608);                 // 2 KiB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 196608 + 3072);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;     // [2]
  uint64_t* v_full = bars + 4;     // [2]
  uint64_t* k_empty = bars + 6;    // [2] released by the QK^T product (the next-but-one K tile streams in during the softmax)
  uint64_t* s_full = bars + 8;     // [2]
  uint64_t* s_free = bars + 10;    // [2]
  uint64_t* p_full = bars + 12;
  uint64_t* pv_done = bars + 13;
  uint64_t* v_empty = bars + 14;   // [2] released by the P V product
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmQKV);
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&s_free[i], kCompute);
    }
    mbar_init(p_full, kCompute); mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int g = a.nq / a.nkv;
  const uint32_t tile_tx = (uint32_t)a.nch * 16384u;

  if (warp == 0) {
    // ===================== TMA producer =====================
    {   // the whole warp walks the schedule (uniform control flow keeps operands in uniform registers); one elected lane issues
      const bool leader = elect_one();
      uint32_t un = 0, n = 0;
      UnitWalk walk(a.sched);
      int u;
      while (walk.next(u)) {
        const int head = u % a.nq;
        const QItem it = load_qitem<128>(a.items, u / a.nq);
        if (it.nt == 0) continue;            // empty items are skipped by every role before any barrier traffic
        const int kvh = head / g;
        mbar_wait_safe(q_empty, (un & 1) ^ 1);
        if (leader) mbar_arrive_expect_tx(q_full, tile_tx);
        for (int c = 0; c < a.nch && leader; ++c) tma_load_3d(sQ + c * 16384, &tmQKV, q_full, c * 64, head, it.q0);
        for (int j = 0; j < it.nt; ++j, ++n) {
          const int s = n & 1;
          const int k0 = tile_k0<128>(it, j);
          mbar_wait_safe(&k_empty[s], ((n >> 1) & 1) ^ 1);
          if (leader) mbar_arrive_expect_tx(&k_full[s], tile_tx);
          for (int c = 0; c < a.nch && leader; ++c)
            tma_load_3d(sK + s * 32768 + c * 16384, &tmQKV, &k_full[s], c * 64, a.nq + kvh, k0);
          mbar_wait_safe(&v_empty[s], ((n >> 1) & 1) ^ 1);
          if (leader) mbar_arrive_expect_tx(&v_full[s], tile_tx);
          for (int c = 0; c < a.nch && leader; ++c)
            tma_load_3d(sV + s * 32768 + c * 16384, &tmQKV, &v_full[s], c * 64, a.nq + a.nkv + kvh, k0);
        }
        ++un;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer 1: S = Q K^T (runs ahead of the softmax by up to two key tiles) =====================
    {   // the whole warp walks the schedule (uniform control flow keeps operands in uniform registers); one elected lane issues
      const bool leader = elect_one();
      const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      const uint64_t dQ_ = make_smem_desc_sw128(smem_u32(sQ), 16, 1024), dK_ = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
      uint32_t un = 0, n = 0;
      UnitWalk walk(a.sched);
      int u;
      while (walk.next(u)) {
        const QItem it = load_qitem<128>(a.items, u / a.nq);
        if (it.nt == 0) continue;
        mbar_wait_safe(q_full, un & 1);
        ++un;
        for (int j = 0; j < it.nt; ++j, ++n) {
          const int s = n & 1;
          mbar_wait_safe(&k_full[s], (n >> 1) & 1);
          if (n >= 2) mbar_wait_safe(&s_free[s], ((n >> 1) - 1) & 1);
          tc_fence_after();
          if (leader) umma_headdim<16384, 16384>(tmem_base + s * 128, dQ_, dK_ + (uint64_t)(s * (32768 >> 4)), idesc_s, a.ksteps);
          if (leader) umma_commit(&s_full[s]);
          if (leader) umma_commit(&k_empty[s]);
          if (leader && j == it.nt - 1) umma_commit(q_empty);
        }
      }
    }
  } else if (warp == 3) {
    // ===================== MMA issuer 2: O += P V (its barrier waits overlap issuer 1's) =====================
    {
      const bool leader = elect_one();
      const uint32_t idesc_pv = make_idesc_bf16(128, a.pv_n, 0, 1);
      const uint32_t tO = tmem_base + 256;
      const uint64_t dP_ = make_smem_desc_sw128(smem_u32(sP), 16, 1024), dV_ = make_smem_desc_sw128(smem_u32(sV), 16384, 1024);
      uint32_t m = 0;
      UnitWalk walk(a.sched);
      int u;
      while (walk.next(u)) {
        const QItem it = load_qitem<128>(a.items, u / a.nq);
        for (int j = 0; j < it.nt; ++j, ++m) {
          const int s = m & 1;
          mbar_wait_safe(&v_full[s], (m >> 1) & 1);
          mbar_wait_safe(p_full, m & 1);
          tc_fence_after();
          if (leader) umma_rows<8, 16384>(tO, dP_, dV_ + (uint64_t)(s * (32768 >> 4)), idesc_pv, j == 0);
          if (leader) umma_commit(&v_empty[s]);
          if (leader) umma_commit(pv_done);
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax + epilogue: row = query, warpgroup h = half of the key columns =====================
    const int h = (warp - 4) >> 2;
    const int wq = (warp - 4) & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_addr = tmem_base + (uint32_t(wq * 32) << 16);
    uint32_t n = 0;
    UnitWalk walk(a.sched);
    int u;
    while (walk.next(u)) {
      const int head = u % a.nq;
      const QItem it = load_qitem<128>(a.items, u / a.nq);
      if (it.nt == 0) continue;
      const bool mine = row < it.nrows;
      const int tok = it.q0 + row;
      int4 r = make_int4(0, 0, 0, 0);
      if (mine) r = a.rng[tok];
      float m_ref = -INFINITY, l = 0.f;     // l: this thread's share (its columns) of the row sum
      for (int j = 0; j < it.nt; ++j, ++n) {
        const int s = n & 1;
        const int k0 = tile_k0<128>(it, j) + 64 * h;
        const int lo = j < it.n_pre ? r.z : r.x, hi = j < it.n_pre ? r.w : r.y;
        const int c_lo = max(0, lo - k0), c_hi = min(64, hi - k0);
        mbar_wait_safe(&s_full[s], (n >> 1) & 1);
        tc_fence_after();
        uint32_t v[2][32];
        tmem_ld_32x32b_x32(lane_addr + s * 128 + h * 64, v[0]);
        tmem_ld_32x32b_x32(lane_addr + s * 128 + h * 64 + 32, v[1]);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_free[s]);
        float x[64];
        float mt = -INFINITY;
        const bool full = c_lo <= 0 && c_hi >= 64;
        if (__all_sync(0xffffffffu, full)) {
#pragma unroll
          for (int c = 0; c < 64; ++c) {
            x[c] = __uint_as_float(v[c >> 5][c & 31]);
            mt = fmaxf(mt, x[c]);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 64; ++c) {
            x[c] = keep_if_inside(__uint_as_float(v[c >> 5][c & 31]), c, c_lo, c_hi);
            mt = fmaxf(mt, x[c]);
          }
        }
        // row maximum over both halves (raw scores; the positive scale commutes with max)
        float* xb = sX + (n & 1) * 256;
        xb[h * 128 + row] = mt;
        named_bar_sync(2, kCompute);
        mt = fmaxf(mt, xb[(h ^ 1) * 128 + row]) * a.sl2;
        const float m_new = fmaxf(m_ref, mt);
        bool waited = false;
        if (j == 0) {
          m_ref = m_new;
        } else {
          // lazy rescale: the accumulator keeps its old reference maximum unless the new one is > 2^8 larger
          const bool need = m_new > m_ref + 8.f;
          if (__any_sync(0xffffffffu, need)) {
            mbar_wait_safe(pv_done, (n - 1) & 1);
            tc_fence_after();
            waited = true;
            const float alpha = need ? exp2f(m_ref - m_new) : 1.f;
            for (int c0 = 16 * h; c0 < a.pv_n; c0 += 32) {
              uint32_t o[16];
              tmem_ld_32x32b_x16(lane_addr + 256 + c0, o);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
              tmem_st_32x32b_x16(lane_addr + 256 + c0, o);
            }
            tmem_st_wait();
            l *= alpha;
            if (need) m_ref = m_new;
          }
        }
        const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int c = 0; c < 64; c += 2) {
          x[c] = exp2f(fmaf(x[c], a.sl2, -m_use));
          x[c + 1] = exp2f(fmaf(x[c + 1], a.sl2, -m_use));
          rs0 += x[c];
          rs1 += x[c + 1];
        }
        l += rs0 + rs1;
        if (n >= 1 && !waited) mbar_wait_safe(pv_done, (n - 1) & 1);   // the previous P V product has consumed the P tile
#pragma unroll
        for (int uu = 0; uu < 8; ++uu)
          *reinterpret_cast<uint4*>(sP + h * 16384 + sw128_off(row, uu)) = pack8(&x[uu * 8]);
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(p_full);
      }
      // ---- epilogue: O / l -> bf16, log-sum-exp ----
      float* xb = sX + (n & 1) * 256;
      xb[h * 128 + row] = l;
      named_bar_sync(2, kCompute);
      l += xb[(h ^ 1) * 128 + row];
      mbar_wait_safe(pv_done, (n - 1) & 1);
      tc_fence_after();
      const float inv = l > 0.f ? 1.f / l : 0.f;
      bf16* orow = a.out + (long long)tok * a.out_ld + (long long)head * a.hd;
      for (int c0 = 16 * h; c0 < a.hd; c0 += 32) {
        uint32_t o[16];
        tmem_ld_32x32b_x16(lane_addr + 256 + c0, o);
        tmem_ld_wait();
        float f[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(o[e]) * inv;
        if (mine) {
          *reinterpret_cast<uint4*>(orow + c0) = pack8(f);
          if (c0 + 8 < a.hd) *reinterpret_cast<uint4*>(orow + c0 + 8) = pack8(f + 8);
        }
      }
      if (mine && h == 0) a.lse2[(long long)head * a.npad + tok] = l > 0.f ? m_ref + log2f(l) : 0.f;
      tc_fence_before();
      named_bar_sync(2, kCompute);   // the exchange buffer of parity n & 1 is rewritten by the next unit's first tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =====================================================================================================================
// backward, dQ:  dQ[q] = scale * sum_k dS[q][k] K[k],  dS = P o (dO V^T - delta),  P = exp2(scale' Q K^T - lse2)
// =====================================================================================================================
// smem: Q [2][128][64] | dO [2][128][64] | 4 stages x { K [2][64][64], V [2][64][64] } | dS [128][64]
// TMEM: S[b] at 64 b, dP[b] at 128 + 64 b, dQ at 256
// Compute warpgroup h owns key columns [32 h, 32 h + 32) of every 64-key tile and the 16-column pieces i = h (mod 2) of dQ.
__global__ void __launch_bounds__(kFThreads, 1)
fmha_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ128, const __grid_constant__ CUtensorMap tmKV64,
                   const __grid_constant__ CUtensorMap tmDO128, const FmhaArgs a) {
  IADR1_SMEM_BASE();
  uint8_t* sQ = smem;
  uint8_t* sdO = smem + 32768;
  uint8_t* sKV = smem + 65536;             // stage st: K at st * 32768, V at st * 32768 + 16384 (chunks 8192 apart)
  uint8_t* sdS = smem + 65536 + 131072;    // 16 KiB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 212992);
  uint64_t* qdo_full = bars + 0;
  uint64_t* qdo_empty = bars + 1;
  uint64_t* kv_full = bars + 2;    // [4]
  uint64_t* kv_empty = bars + 6;   // [4]
  uint64_t* sdp_full = bars + 10;  // [2]
  uint64_t* sdp_free = bars + 12;  // [2]
  uint64_t* ds_full = bars + 14;
  uint64_t* ds_free = bars + 15;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ128);
    tma_prefetch_desc(&tmKV64);
    tma_prefetch_desc(&tmDO128);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(qdo_full, 1); mbar_init(qdo_empty, 1);
    for (int i = 0; i < 4; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sdp_full[i], 1); mbar_init(&sdp_free[i], kCompute); }
    mbar_init(ds_full, kCompute); mbar_init(ds_free, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int g = a.nq / a.nkv;

  if (warp == 0) {
    {   // the whole warp walks the schedule (uniform control flow keeps operands in uniform registers); one elected lane issues
      const bool leader = elect_one();
      uint32_t un = 0, n = 0;
      UnitWalk walk(a.sched);
      int u;
      while (walk.next(u)) {
        const int head = u % a.nq;
        const QItem it = load_qitem<64>(a.items, u / a.nq);
        if (it.nt == 0) continue;
        const int kvh = head / g;
        mbar_wait_safe(qdo_empty, (un & 1) ^ 1);
        if (leader) mbar_arrive_expect_tx(qdo_full, 2u * a.nch * 16384u);
        for (int c = 0; c < a.nch && leader; ++c) {
          tma_load_3d(sQ + c * 16384, &tmQ128, qdo_full, c * 64, head, it.q0);
          tma_load_3d(sdO + c * 16384, &tmDO128, qdo_full, c * 64, head, it.q0);
        }
        for (int j = 0; j < it.nt; ++j, ++n) {
          const int st = n & 3;
          const int k0 = tile_k0<64>(it, j);
          mbar_wait_safe(&kv_empty[st], ((n >> 2) & 1) ^ 1);
          if (leader) mbar_arrive_expect_tx(&kv_full[st], 2u * a.nch * 8192u);
          for (int c = 0; c < a.nch && leader; ++c) {
            tma_load_3d(sKV + st * 32768 + c * 8192, &tmKV64, &kv_full[st], c * 64, a.nq + kvh, k0);
            tma_load_3d(sKV + st * 32768 + 16384 + c * 8192, &tmKV64, &kv_full[st], c * 64, a.nq + a.nkv + kvh, k0);
          }
        }
        ++un;
      }
    }
  } else if (warp == 1) {
    {   // issuer 1: S = Q K^T and dP = dO V^T of every 64-key tile
      const bool leader = elect_one();
      const uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
      const uint64_t dQ_ = make_smem_desc_sw128(smem_u32(sQ), 16, 1024), dO_ = make_smem_desc_sw128(smem_u32(sdO), 16, 1024);
      const uint64_t dK_ = make_smem_desc_sw128(smem_u32(sKV), 16, 1024);
      uint32_t un = 0, n = 0;
      UnitWalk walk(a.sched);
      int u;
      while (walk.next(u)) {
        const QItem it = load_qitem<64>(a.items, u / a.nq);
        if (it.nt == 0) continue;
        mbar_wait_safe(qdo_full, un & 1);
        ++un;
        for (int j = 0; j < it.nt; ++j, ++n) {
          const int st = n & 3, b = n & 1;
          mbar_wait_safe(&kv_full[st], (n >> 2) & 1);
          if (n >= 2) mbar_wait_safe(&sdp_free[b], ((n >> 1) - 1) & 1);
          tc_fence_after();
          const uint64_t kst = dK_ + (uint64_t)(st * (32768 >> 4));
          if (leader) umma_headdim<16384, 8192>(tmem_base + b * 64, dQ_, kst, idesc_s, a.ksteps);
          if (leader) umma_headdim<16384, 8192>(tmem_base + 128 + b * 64, dO_, kst + (16384 >> 4), idesc_s, a.ksteps);
          if (leader) umma_commit(&sdp_full[b]);
          if (leader && j == it.nt - 1) umma_commit(qdo_empty);
        }
      }
    }
  } else if (warp == 3) {
    {   // issuer 2: dQ += dS K (K tile read MN-major from the same stage)
      const bool leader = elect_one();
      const uint32_t idesc_dq = make_idesc_bf16(128, a.pv_n, 0, 1);
      const uint64_t dKmn = make_smem_desc_sw128(smem_u32(sKV), 8192, 1024), dS_ = make_smem_desc_sw128(smem_u32(sdS), 16, 1024);
      uint32_t m = 0;
      UnitWalk walk(a.sched);
      int u;
      while (walk.next(u)) {
        const QItem it = load_qitem<64>(a.items, u / a.nq);
        for (int j = 0; j < it.nt; ++j, ++m) {
          const int st = m & 3;
          mbar_wait_safe(&kv_full[st], (m >> 2) & 1);
          mbar_wait_safe(ds_full, m & 1);
          tc_fence_after();
          if (leader) umma_rows<4, 16384>(tmem_base + 256, dS_, dKmn + (uint64_t)(st * (32768 >> 4)), idesc_dq, j == 0);
          if (leader) umma_commit(&kv_empty[st]);
          if (leader) umma_commit(ds_free);
        }
      }
    }
  } else if (warp >= 4) {
    const int h = (warp - 4) >> 2;
    const int wq = (warp - 4) & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_addr = tmem_base + (uint32_t(wq * 32) << 16);
    uint32_t n = 0;
    UnitWalk walk(a.sched);
    int u;
    while (walk.next(u)) {
      const int head = u % a.nq;
      const QItem it = load_qitem<64>(a.items, u / a.nq);
      if (it.nt == 0) continue;
      const bool mine = row < it.nrows;
      const int tok = it.q0 + row;
      int4 r = make_int4(0, 0, 0, 0);
      float L2 = 0.f, dl = 0.f;
      if (mine) {
        r = a.rng[tok];
        L2 = a.lse2[(long long)head * a.npad + tok];
        dl = a.delta[(long long)head * a.npad + tok];
      }
      for (int j = 0; j < it.nt; ++j, ++n) {
        const int b = n & 1;
        const int k0 = tile_k0<64>(it, j) + 32 * h;
        const int lo = j < it.n_pre ? r.z : r.x, hi = j < it.n_pre ? r.w : r.y;
        const int c_lo = max(0, lo - k0), c_hi = min(32, hi - k0);
        mbar_wait_safe(&sdp_full[b], (n >> 1) & 1);
        tc_fence_after();
        uint32_t sv[32], dv[32];
        tmem_ld_32x32b_x32(lane_addr + b * 64 + 32 * h, sv);
        tmem_ld_32x32b_x32(lane_addr + 128 + b * 64 + 32 * h, dv);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&sdp_free[b]);
        float ds[32];
        const bool full = c_lo <= 0 && c_hi >= 32;
        if (__all_sync(0xffffffffu, full)) {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            ds[c] = exp2f(fmaf(__uint_as_float(sv[c]), a.sl2, -L2)) * (__uint_as_float(dv[c]) - dl);
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float e = keep_if_inside(fmaf(__uint_as_float(sv[c]), a.sl2, -L2), c, c_lo, c_hi);
            ds[c] = exp2f(e) * (__uint_as_float(dv[c]) - dl);
          }
        }
        if (n >= 1) mbar_wait_safe(ds_free, (n - 1) & 1);   // the previous dQ product has consumed the dS tile
#pragma unroll
        for (int uu = 0; uu < 4; ++uu)
          *reinterpret_cast<uint4*>(sdS + sw128_off(row, 4 * h + uu)) = pack8(&ds[uu * 8]);
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(ds_full);
      }
      mbar_wait_safe(ds_free, (n - 1) & 1);   // the last dQ product of this tile has completed
      tc_fence_after();
      bf16* drow = a.dq + (long long)tok * a.dq_ld + (long long)head * a.hd;
      for (int c0 = 16 * h; c0 < a.hd; c0 += 32) {
        uint32_t o[16];
        tmem_ld_32x32b_x16(lane_addr + 256 + c0, o);
        tmem_ld_wait();
        float f[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(o[e]) * a.scale;
        if (mine) {
          *reinterpret_cast<uint4*>(drow + c0) = pack8(f);
          if (c0 + 8 < a.hd) *reinterpret_cast<uint4*>(drow + c0 + 8) = pack8(f + 8);
        }
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =====================================================================================================================
// backward, dK / dV: one (128-key tile, kv head) per unit; inner loop over the g query heads of the kv head and the
// 64-query tiles of the item's query range. S^T = K Q^T and dP^T = V dO^T put the KEY on the TMEM lane, so P^T / dS^T
// come out directly as the K-major A operands of dV += P^T dO and dK += dS^T Q (dO / Q tiles read MN-major).
// =====================================================================================================================
// smem: K [2][128][64] | V [2][128][64] | 3 stages x { Q [2][64][64], dO [2][64][64] } | P^T [128][64] | dS^T [128][64] |
//       per-query vectors (2 buffers x {lse2 [64], delta [64], ranges [64] int4, flags})
// TMEM: S^T[b] at 64 b, dP^T[b] at 128 + 64 b, dK at 256, dV at 384
// Compute warpgroup h owns query columns [32 h, 32 h + 32) of every 64-query tile; in the epilogue h = 0 adds dK, h = 1 dV.
struct KIter {      // position in the (unit, query head, query tile) iteration space of one CTA
  int u, hq, t;     // unit index, head within the kv group, 64-query tile index
  int k0, nkeys, q0, q1, f0, f1, nqt;
};
__device__ __forceinline__ void kiter_load(KIter& it, const FmhaArgs& a) {
  const int* p = a.items + 6 * (it.u / a.nkv);
  it.k0 = p[0]; it.nkeys = p[1]; it.q0 = p[2]; it.q1 = p[3]; it.f0 = p[4]; it.f1 = p[5];
  it.nqt = (it.q1 - it.q0 + 63) / 64;
}
// advances to the next (hq, t); returns false when the unit is finished
__device__ __forceinline__ bool kiter_next_in_unit(KIter& it, int g) {
  if (++it.t < it.nqt) return true;
  it.t = 0;
  return ++it.hq < g;
}
static constexpr int kKvStage = 34816;   // Q 16 KiB | dO 16 KiB | lse2 [64] | delta [64] | ranges [64] int4 (+ pad to 1 KiB)

__global__ void __launch_bounds__(kFThreads, 1)
fmha_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmKV128, const __grid_constant__ CUtensorMap tmQ64,
                    const __grid_constant__ CUtensorMap tmDO64, const FmhaArgs a) {
  IADR1_SMEM_BASE();
  uint8_t* sK = smem;
  uint8_t* sV = smem + 32768;
  uint8_t* sQD = smem + 65536;                  // stage st at st * kKvStage: Q, dO (chunks 8192 apart), per-query vectors
  uint8_t* sPT = smem + 65536 + 3 * kKvStage;   // 16 KiB
  uint8_t* sdST = sPT + 16384;                  // 16 KiB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdST + 16384);
  uint64_t* kv_full = bars + 0;
  uint64_t* kv_empty = bars + 1;
  uint64_t* qd_full = bars + 2;    // [3]
  uint64_t* qd_empty = bars + 5;   // [3]
  uint64_t* st_full = bars + 8;    // [2]
  uint64_t* st_free = bars + 10;   // [2]
  uint64_t* pds_full = bars + 12;
  uint64_t* pds_free = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmKV128);
    tma_prefetch_desc(&tmQ64);
    tma_prefetch_desc(&tmDO64);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1); mbar_init(kv_empty, 1);
    for (int i = 0; i < 3; ++i) { mbar_init(&qd_full[i], 1); mbar_init(&qd_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&st_full[i], 1); mbar_init(&st_free[i], kCompute); }
    mbar_init(pds_full, kCompute); mbar_init(pds_free, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int g = a.nq / a.nkv;

  if (warp == 0) {
    {   // the whole warp walks the schedule (uniform control flow keeps operands in uniform registers); one elected lane issues
      const bool leader = elect_one();
      uint32_t un = 0, n = 0;
      KIter it;
      UnitWalk walk(a.sched);
      while (walk.next(it.u)) {
        kiter_load(it, a);
        if (it.nqt == 0) continue;
        const int kvh = it.u % a.nkv;
        mbar_wait_safe(kv_empty, (un & 1) ^ 1);
        if (leader) mbar_arrive_expect_tx(kv_full, 2u * a.nch * 16384u);
        for (int c = 0; c < a.nch && leader; ++c) {
          tma_load_3d(sK + c * 16384, &tmKV128, kv_full, c * 64, a.nq + kvh, it.k0);
          tma_load_3d(sV + c * 16384, &tmKV128, kv_full, c * 64, a.nq + a.nkv + kvh, it.k0);
        }
        it.hq = 0; it.t = 0;
        do {
          const int st = n % 3;
          const int head = kvh * g + it.hq, qs = it.q0 + 64 * it.t;
          uint8_t* sb = sQD + st * kKvStage;
          if (leader) FTRACE(0, n, 0);
          mbar_wait_safe(&qd_empty[st], ((n / 3) & 1) ^ 1);
          if (leader) FTRACE(0, n, 1);
          if (leader) mbar_arrive_expect_tx(&qd_full[st], 2u * a.nch * 8192u + 1536u);
          for (int c = 0; c < a.nch && leader; ++c) {
            tma_load_3d(sb + c * 8192, &tmQ64, &qd_full[st], c * 64, head, qs);
            tma_load_3d(sb + 16384 + c * 8192, &tmDO64, &qd_full[st], c * 64, head, qs);
          }
          // per-query vectors of the tile ride on the same barrier (qs % 4 == 0: the host aligns the items' q0)
          if (leader) bulk_copy_g2s(sb + 32768, a.lse2 + (long long)head * a.npad + qs, 256, &qd_full[st]);
          if (leader) bulk_copy_g2s(sb + 32768 + 256, a.delta + (long long)head * a.npad + qs, 256, &qd_full[st]);
          if (leader) bulk_copy_g2s(sb + 32768 + 512, a.rng + qs, 1024, &qd_full[st]);
          ++n;
        } while (kiter_next_in_unit(it, g));
        ++un;
      }
    }
  } else if (warp == 1) {
    {   // issuer 1: S^T = K Q^T and dP^T = V dO^T of every (head, 64-query tile)
      const bool leader = elect_one();
      const uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
      const uint64_t dK_ = make_smem_desc_sw128(smem_u32(sK), 16, 1024), dV_ = make_smem_desc_sw128(smem_u32(sV), 16, 1024);
      const uint64_t dQs = make_smem_desc_sw128(smem_u32(sQD), 16, 1024);
      uint32_t un = 0, n = 0;
      KIter it;
      UnitWalk walk(a.sched);
      while (walk.next(it.u)) {
        kiter_load(it, a);
        const int nit = it.nqt * g;
        if (nit == 0) continue;
        mbar_wait_safe(kv_full, un & 1);
        ++un;
        for (int i = 0; i < nit; ++i, ++n) {
          const int st = n % 3, b = n & 1;
          if (leader) FTRACE(1, n, 0);
          mbar_wait_safe(&qd_full[st], (n / 3) & 1);
          if (leader) FTRACE(1, n, 1);
          if (n >= 2) mbar_wait_safe(&st_free[b], ((n >> 1) - 1) & 1);
          if (leader) FTRACE(1, n, 2);
          tc_fence_after();
          const uint64_t qst = dQs + (uint64_t)(st * (kKvStage >> 4));
          if (leader) umma_headdim<16384, 8192>(tmem_base + b * 64, dK_, qst, idesc_s, a.ksteps);
          if (leader) umma_headdim<16384, 8192>(tmem_base + 128 + b * 64, dV_, qst + (16384 >> 4), idesc_s, a.ksteps);
          if (leader) umma_commit(&st_full[b]);
          if (leader) FTRACE(1, n, 3);
          if (leader && i == nit - 1) umma_commit(kv_empty);
        }
      }
    }
  } else if (warp == 3) {
    {   // issuer 2: dV += P^T dO and dK += dS^T Q (dO / Q tiles read MN-major from the same stage)
      const bool leader = elect_one();
      const uint32_t idesc_kv = make_idesc_bf16(128, a.pv_n, 0, 1);
      const uint64_t dQmn = make_smem_desc_sw128(smem_u32(sQD), 8192, 1024);
      const uint64_t dPT = make_smem_desc_sw128(smem_u32(sPT), 16, 1024), dDS = make_smem_desc_sw128(smem_u32(sdST), 16, 1024);
      uint32_t m = 0;
      KIter it;
      UnitWalk walk(a.sched);
      while (walk.next(it.u)) {
        kiter_load(it, a);
        const int nit = it.nqt * g;
        for (int i = 0; i < nit; ++i, ++m) {
          const int st = m % 3;
          if (leader) FTRACE(1, m, 4);
          mbar_wait_safe(&qd_full[st], (m / 3) & 1);
          mbar_wait_safe(pds_full, m & 1);
          if (leader) FTRACE(1, m, 5);
          tc_fence_after();
          const uint64_t qmn = dQmn + (uint64_t)(st * (kKvStage >> 4));
          if (leader) umma_rows<4, 16384>(tmem_base + 384, dPT, qmn + (16384 >> 4), idesc_kv, i == 0);   // dV += P^T dO
          if (leader) umma_rows<4, 16384>(tmem_base + 256, dDS, qmn, idesc_kv, i == 0);                  // dK += dS^T Q
          if (leader) umma_commit(&qd_empty[st]);
          if (leader) umma_commit(pds_free);
          if (leader) FTRACE(1, m, 6);
        }
      }
    }
  } else if (warp >= 4) {
    const int h = (warp - 4) >> 2;
    const int wq = (warp - 4) & 3;
    const int row = wq * 32 + lane;              // key row of the tile
    const uint32_t lane_addr = tmem_base + (uint32_t(wq * 32) << 16);
    uint32_t n = 0;
    KIter it;
    UnitWalk walk(a.sched);
    while (walk.next(it.u)) {
      kiter_load(it, a);
      if (it.nqt == 0) continue;
      const int kvh = it.u % a.nkv;
      const int key = it.k0 + row;
      const bool kvalid = row < it.nkeys;
      it.hq = 0; it.t = 0;
      do {
        const int b = n & 1, st = n % 3;
        const int nqr = min(64, it.q1 - (it.q0 + 64 * it.t)) - 32 * h;   // valid columns among this warpgroup's 32
        const bool tr = threadIdx.x == 128;
        if (tr) FTRACE(2, n, 0);
        mbar_wait_safe(&qd_full[st], (n / 3) & 1);                       // the tile's per-query vectors have landed
        mbar_wait_safe(&st_full[b], (n >> 1) & 1);
        if (tr) FTRACE(2, n, 1);
        tc_fence_after();
        uint32_t sv[32], dv[32];
        tmem_ld_32x32b_x32(lane_addr + b * 64 + 32 * h, sv);
        tmem_ld_32x32b_x32(lane_addr + 128 + b * 64 + 32 * h, dv);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&st_free[b]);
        if (tr) FTRACE(2, n, 2);
        const uint8_t* vb = sQD + st * kKvStage + 32768;
        const float4* vL = reinterpret_cast<const float4*>(vb) + 8 * h;
        const float4* vD = reinterpret_cast<const float4*>(vb + 256) + 8 * h;
        const int4* vR = reinterpret_cast<const int4*>(vb + 512) + 32 * h;
        float pp[32], ds[32];
        if (it.t >= it.f0 && it.t < it.f1) {   // every (key, query) pair of the tile is allowed (host-computed): no masks
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 L = vL[c4], Dl = vD[c4];
            const float Ls[4] = {L.x, L.y, L.z, L.w}, Ds[4] = {Dl.x, Dl.y, Dl.z, Dl.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = 4 * c4 + e;
              pp[c] = exp2f(fmaf(__uint_as_float(sv[c]), a.sl2, -Ls[e]));
              ds[c] = pp[c] * (__uint_as_float(dv[c]) - Ds[e]);
            }
          }
        } else if (nqr >= 32) {
          // masked tile, all of this warpgroup's 32 queries valid: integer masks (no predicates -> the columns pipeline)
          const int keym = kvalid ? key : -(1 << 30);
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 L = vL[c4], Dl = vD[c4];
            const float Ls[4] = {L.x, L.y, L.z, L.w}, Ds[4] = {Dl.x, Dl.y, Dl.z, Dl.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = 4 * c4 + e;
              const int4 rr = vR[c];
              const int outside = (((keym - rr.x) | (rr.y - 1 - keym)) & ((keym - rr.z) | (rr.w - 1 - keym))) >> 31;
              const float ex = fmaf(__uint_as_float(sv[c]), a.sl2, -Ls[e]);
              pp[c] = exp2f(__int_as_float((__float_as_int(ex) & ~outside) | (int)(0xff800000u & (unsigned)outside)));
              ds[c] = pp[c] * (__uint_as_float(dv[c]) - Ds[e]);
            }
          }
        } else {
          // ragged last tile of a query range: per-column validity as well
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 L = vL[c4], Dl = vD[c4];
            const float Ls[4] = {L.x, L.y, L.z, L.w}, Ds[4] = {Dl.x, Dl.y, Dl.z, Dl.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = 4 * c4 + e;
              const int4 rr = vR[c];
              const bool ok = kvalid && c < nqr && ((key >= rr.x && key < rr.y) || (key >= rr.z && key < rr.w));
              const float ex = ok ? fmaf(__uint_as_float(sv[c]), a.sl2, -Ls[e]) : -INFINITY;
              pp[c] = exp2f(ex);
              ds[c] = ok ? pp[c] * (__uint_as_float(dv[c]) - Ds[e]) : 0.f;
            }
          }
        }
        if (tr) FTRACE(2, n, 3);
        if (n >= 1) mbar_wait_safe(pds_free, (n - 1) & 1);   // the previous dV / dK products have consumed P^T / dS^T
        if (tr) FTRACE(2, n, 4);
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) {
          *reinterpret_cast<uint4*>(sPT + sw128_off(row, 4 * h + uu)) = pack8(&pp[uu * 8]);
          *reinterpret_cast<uint4*>(sdST + sw128_off(row, 4 * h + uu)) = pack8(&ds[uu * 8]);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(pds_full);
        if (tr) FTRACE(2, n, 5);
        ++n;
      } while (kiter_next_in_unit(it, g));
      // ---- epilogue: warpgroup 0 adds dK (x scale), warpgroup 1 adds dV to the fp32 gradient rows (bulk reductions) ----
      mbar_wait_safe(pds_free, (n - 1) & 1);
      tc_fence_after();
      float* stage = reinterpret_cast<float*>(sPT) + h * (128 * 20) + row * 20;   // [2][128][20] fp32 over P^T / dS^T
      const long long ldkv = 2LL * a.nkv * a.hd;
      float* grow = a.dkv32 + (long long)key * ldkv + (long long)(h * a.nkv + kvh) * a.hd;
      const float mul = h == 0 ? a.scale : 1.f;
      for (int c0 = 0; c0 < a.hd; c0 += 16) {
        uint32_t o[16];
        tmem_ld_32x32b_x16(lane_addr + 256 + h * 128 + c0, o);
        tmem_ld_wait();
        bulk_wait_read0();             // the previous piece's reduction has read the staging row
#pragma unroll
        for (int e = 0; e < 16; e += 4)
          *reinterpret_cast<float4*>(stage + e) =
              make_float4(__uint_as_float(o[e]) * mul, __uint_as_float(o[e + 1]) * mul, __uint_as_float(o[e + 2]) * mul,
                          __uint_as_float(o[e + 3]) * mul);
        fence_proxy_async_smem();
        if (kvalid) bulk_reduce_add_f32(grow + c0, stage, (uint32_t)min(16, a.hd - c0) * 4u);
        bulk_commit();
      }
      bulk_wait_read0();
      tc_fence_before();
      named_bar_sync(1, kCompute);   // the staging rows alias the P^T / dS^T tiles of the next unit
    }
    bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// delta[head][tok] = sum_d dO[tok][head][d] * O[tok][head][d]   (one warp per (tok, head); head-major output)
__global__ void fmha_delta_kernel(const bf16* __restrict__ dO, const bf16* __restrict__ O, float* __restrict__ delta,
                                  long long n_pairs, int nq, int hd, long long ld, long long npad) {
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= n_pairs) return;
  const int lane = threadIdx.x & 31;
  const long long tok = wid / nq;
  const int head = (int)(wid - tok * nq);
  const bf16* a = dO + tok * ld + (long long)head * hd;
  const bf16* b = O + tok * ld + (long long)head * hd;
  float s = 0.f;
  for (int d = lane * 8; d < hd; d += 256) {
    const uint4 x = *reinterpret_cast<const uint4*>(a + d), y = *reinterpret_cast<const uint4*>(b + d);
    const __nv_bfloat162* xh = reinterpret_cast<const __nv_bfloat162*>(&x);
    const __nv_bfloat162* yh = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fx = __bfloat1622float2(xh[e]), fy = __bfloat1622float2(yh[e]);
      s += fx.x * fy.x + fx.y * fy.y;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) delta[(long long)head * npad + tok] = s;
}

// dst[r][c] (bf16, row stride ld) = src[r][c] (fp32, dense [rows][cols]); cols % 4 == 0
__global__ void fmha_cast_rows_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long rows, int cols,
                                      long long ld) {
  const long long nvec = rows * (cols >> 2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (cols >> 2);
    const int c = (int)(i - r * (cols >> 2)) * 4;
    const float4 v = *reinterpret_cast<const float4*>(src + r * cols + c);
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&lo);
    o.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(dst + r * ld + c) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn3 fmha_encode_fn() {
  static EncodeTiledFn3 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn3>(p);
  });
  return fn;
}

// [tokens][heads][hd] view of a row-major activation buffer (row stride ld elements); box {64, 1, box_rows}
static int make_head_map(CUtensorMap* out, const void* ptr, long long tokens, int heads, int hd, long long ld, int box_rows) {
  typedef std::tuple<const void*, long long, int, int, long long, int> Key;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  Key key(ptr, tokens, heads, hd, ld, box_rows);
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn3 fn = fmha_encode_fn();
  if (!fn) return set_error("cuTensorMapEncodeTiled entry point not found (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (hd & 7) || (ld & 7))
    return set_error("fmha: buffers must be 16-byte aligned with head_dim and row stride multiples of 8 elements");
  cuuint64_t dims[3] = {(cuuint64_t)hd, (cuuint64_t)heads, (cuuint64_t)tokens};
  cuuint64_t strides[2] = {(cuuint64_t)hd * 2, (cuuint64_t)ld * 2};
  cuuint32_t box[3] = {64, 1, (cuuint32_t)box_rows};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error("fmha: cuTensorMapEncodeTiled failed (%d) tokens=%lld heads=%d hd=%d ld=%lld", (int)r, tokens, heads, hd, ld);
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 4096) cache.clear();
  cache[key] = *out;
  return 0;
}

static long long* g_fmha_trace = nullptr;
void fmha_set_trace(long long* p) { g_fmha_trace = p; }

static int fmha_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static int fmha_fill_args(FmhaArgs& a, int nq, int nkv, int hd, float scale, int pv_n_override) {
  if (nq <= 0 || nkv <= 0 || nq % nkv) return set_error("fmha: nq must be a positive multiple of nkv");
  if (hd < 8 || hd > 128 || (hd & 7)) return set_error("fmha: head_dim must be a multiple of 8 in [8, 128] (got %d)", hd);
  a.nq = nq; a.nkv = nkv; a.hd = hd;
  a.ksteps = (hd + 15) / 16;
  a.nch = (a.ksteps * 16 + 63) / 64;
  a.pv_n = pv_n_override > 0 ? pv_n_override : a.ksteps * 16;
  if (a.pv_n < hd || a.pv_n > 128 || (a.pv_n & 15)) return set_error("fmha: bad pv_n %d", a.pv_n);
  a.scale = scale;
  a.sl2 = scale * 1.4426950408889634f;
  a.trace = g_fmha_trace;
  return 0;
}

static constexpr size_t kFmhaSmem = 212992 + 512 + 1024;
static int fmha_set_attrs() {
  static bool done = false;
  if (done) return 0;
  cudaError_t e = cudaFuncSetAttribute(fmha_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFmhaSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(fmha_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFmhaSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(fmha_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFmhaSmem);
  if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(fmha smem): %s", cudaGetErrorString(e));
  done = true;
  return 0;
}

int launch_fmha_fwd(const void* qkv, long long n_tokens, int nq, int nkv, int hd, const int* rng, const int* items, int n_items,
                    const int* sched, int n_cta, void* out, float* lse2, long long npad, float scale, int pv_n,
                    cudaStream_t stream) {
  if (n_items <= 0 || n_tokens <= 0) return 0;
  if (npad < n_tokens + 64 || (npad & 3)) return set_error("fmha: npad must be a multiple of 4 >= n_tokens + 64");
  if (n_cta <= 0 || !sched) return set_error("fmha: missing schedule");
  FmhaArgs a;
  memset(&a, 0, sizeof(a));
  int rc = fmha_fill_args(a, nq, nkv, hd, scale, pv_n);
  if (rc) return rc;
  a.rng = reinterpret_cast<const int4*>(rng);
  a.items = items;
  a.sched = sched;
  a.n_items = n_items;
  a.out = reinterpret_cast<bf16*>(out);
  a.out_ld = (long long)nq * hd;
  a.lse2 = lse2;
  a.npad = npad;
  if (reinterpret_cast<uintptr_t>(out) & 15) return set_error("fmha: out must be 16-byte aligned");
  const long long D = (long long)(nq + 2 * nkv) * hd;
  CUtensorMap tm;
  if ((rc = make_head_map(&tm, qkv, n_tokens, nq + 2 * nkv, hd, D, 128))) return rc;
  if ((rc = fmha_set_attrs())) return rc;
  launch_kernel(fmha_fwd_kernel, dim3(n_cta), dim3(kFThreads), kFmhaSmem, stream, tm, a);
  IADR1_CHECK_LAUNCH("fmha_fwd");
  return 0;
}

int launch_fmha_bwd(const void* qkv, const void* dout, const void* out, const float* lse2, long long n_tokens, int nq, int nkv,
                    int hd, const int* rng, const int* q_items, int n_q_items, const int* q_sched, int n_q_cta,
                    const int* k_items, int n_k_items, const int* k_sched, int n_k_cta, void* dqkv, float* delta, float* dkv32,
                    long long npad, float scale, int pv_n, cudaStream_t stream) {
  if (n_tokens <= 0) return 0;
  if (npad < n_tokens + 64 || (npad & 3)) return set_error("fmha: npad must be a multiple of 4 >= n_tokens + 64");
  FmhaArgs a;
  memset(&a, 0, sizeof(a));
  int rc = fmha_fill_args(a, nq, nkv, hd, scale, pv_n);
  if (rc) return rc;
  const long long D = (long long)(nq + 2 * nkv) * hd, QH = (long long)nq * hd;
  a.rng = reinterpret_cast<const int4*>(rng);
  a.lse2 = const_cast<float*>(lse2);
  a.delta = delta;
  a.npad = npad;
  a.dq = reinterpret_cast<bf16*>(dqkv);
  a.dq_ld = D;
  a.dkv32 = dkv32;
  if ((reinterpret_cast<uintptr_t>(dqkv) & 15) || (reinterpret_cast<uintptr_t>(dkv32) & 15))
    return set_error("fmha: gradient buffers must be 16-byte aligned");
  if ((rc = fmha_set_attrs())) return rc;
  CUtensorMap tq128, tq64, to128, to64;
  if ((rc = make_head_map(&tq128, qkv, n_tokens, nq + 2 * nkv, hd, D, 128))) return rc;
  if ((rc = make_head_map(&tq64, qkv, n_tokens, nq + 2 * nkv, hd, D, 64))) return rc;
  if ((rc = make_head_map(&to128, dout, n_tokens, nq, hd, QH, 128))) return rc;
  if ((rc = make_head_map(&to64, dout, n_tokens, nq, hd, QH, 64))) return rc;
  {
    const long long pairs = n_tokens * nq;
    const int wpb = 8;
    fmha_delta_kernel<<<(unsigned)((pairs + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
        reinterpret_cast<const bf16*>(dout), reinterpret_cast<const bf16*>(out), delta, pairs, nq, hd, QH, npad);
    IADR1_CHECK_LAUNCH("fmha_delta");
  }
  const long long kv_cols = 2LL * nkv * hd;
  if (cudaMemsetAsync(dkv32, 0, (size_t)(n_tokens * kv_cols) * sizeof(float), stream) != cudaSuccess)
    return set_error("fmha: cudaMemsetAsync failed");
  if (n_q_items > 0) {
    if (n_q_cta <= 0 || !q_sched) return set_error("fmha: missing dQ schedule");
    a.items = q_items;
    a.sched = q_sched;
    a.n_items = n_q_items;
    launch_kernel(fmha_bwd_dq_kernel, dim3(n_q_cta), dim3(kFThreads), kFmhaSmem, stream, tq128, tq64, to128, a);
    IADR1_CHECK_LAUNCH("fmha_bwd_dq");
  }
  if (n_k_items > 0) {
    if (n_k_cta <= 0 || !k_sched) return set_error("fmha: missing dK/dV schedule");
    a.items = k_items;
    a.sched = k_sched;
    a.n_items = n_k_items;
    launch_kernel(fmha_bwd_dkv_kernel, dim3(n_k_cta), dim3(kFThreads), kFmhaSmem, stream, tq128, tq64, to64, a);
    IADR1_CHECK_LAUNCH("fmha_bwd_dkv");
  }
  {
    const long long nvec = n_tokens * (kv_cols >> 2);
    const int blocks = (int)((nvec + 255) / 256 < 4096 ? (nvec + 255) / 256 : 4096);
    fmha_cast_rows_kernel<<<blocks, 256, 0, stream>>>(dkv32, reinterpret_cast<bf16*>(dqkv) + QH, n_tokens, (int)kv_cols, D);
    IADR1_CHECK_LAUNCH("fmha_cast_dkv");
  }
  return 0;
}

}  // namespace iadr1

extern "C" {
int iadr1_fmha_set_trace(long long* device_buf) {
  iadr1::fmha_set_trace(device_buf);
  return 0;
}
int iadr1_fmha_fwd(const void* qkv, long long n_tokens, int nq, int nkv, int hd, const int* ranges, const int* items,
                   int n_items, const int* sched, int n_cta, void* out, float* lse2, long long npad, float scale, int pv_n,
                   void* stream) {
  return iadr1::launch_fmha_fwd(qkv, n_tokens, nq, nkv, hd, ranges, items, n_items, sched, n_cta, out, lse2, npad, scale, pv_n,
                                static_cast<cudaStream_t>(stream));
}
int iadr1_fmha_bwd(const void* qkv, const void* dout, const void* out, const float* lse2, long long n_tokens, int nq, int nkv,
                   int hd, const int* ranges, const int* q_items, int n_q_items, const int* q_sched, int n_q_cta,
                   const int* k_items, int n_k_items, const int* k_sched, int n_k_cta, void* dqkv, float* delta, float* dkv32,
                   long long npad, float scale, int pv_n, void* stream) {
  return iadr1::launch_fmha_bwd(qkv, dout, out, lse2, n_tokens, nq, nkv, hd, ranges, q_items, n_q_items, q_sched, n_q_cta,
                                k_items, n_k_items, k_sched, n_k_cta, dqkv, delta, dkv32, npad, scale, pv_n,
                                static_cast<cudaStream_t>(stream));
}
}
