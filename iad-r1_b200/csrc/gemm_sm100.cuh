// Argument block of the one tcgen05 GEMM kernel family (see gemm_sm100.cu).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace iadr1 {

enum GemmEpilogue : int {
  EPI_STORE = 0,     // C = alpha * acc (+ bias) (+ residual), bf16 or f32, optional f32 accumulate / atomic / transposed
  EPI_LSE = 1,       // per-row partial (max, sum exp) over this tile's columns + target-logit pick  (fused lm_head)
  EPI_SWIGLU = 3,    // A = [gate rows; up rows] (M = 2I): C^T[n][f] = bf16(silu(gate_f . b_n)) * (up_f . b_n), bf16 (decode MLP)
  EPI_DLOGITS = 2,   // C = (exp(alpha*acc - lse[m]) - [n == label[m]]) * gscale[m]  as bf16         (lm_head bwd)
  EPI_SWIGLU_T = 4,  // training MLP: B = [gate rows; up rows] (N = 2I accumulator columns per 2 x 128-feature tile):
                     // C[m][f] = bf16(silu(bf16(x_m . gate_f))) * bf16(x_m . up_f), optionally gate | up kept for the backward
  EPI_SWIGLU_BWD = 5,  // down-projection input gradient with the SwiGLU backward in the epilogue: d = bf16(acc) (= dact[m][f]),
                       // gu[m][f] <- bf16(d * up * silu'(gate)), gu[m][I + f] <- bf16(d * silu(gate)) in place (gu = gu_out, N = I)
};

struct GemmArgs {
  int M, N, K;
  int batch;      // number of independent problems; z -> (z % batch_lo, z / batch_lo)
  int batch_lo;
  int b_lo_div;   // B's lo batch coordinate is (z % batch_lo) / b_lo_div (grouped-query sharing)
  int block_n;    // multiple of 16, <= 256
  int stages;
  int tmem_cols;  // TMEM columns allocated (power of two >= 2 * block_n); accumulator stage s lives at s * tmem_cols / 2
  int a_mn, b_mn; // operand majorness: 0 = K-major (reduction dim contiguous), 1 = MN-major
  int a_chunked, b_chunked;  // MN-major operand addressed through the 3-D view [mn / 64][k][64]: one TMA box per k-block
  int kmode;      // 0 full K; 1: k < m0 + 128 + causal_off (rows attend to keys <= row + off); 2: k >= m0 - causal_off
  int skip_mode;  // 1: skip output tiles with n0 > m0 + 127 + causal_off (fully masked)
  int causal_off;
  int split_k;    // >= 1; > 1 requires atomic f32 output
  int up_row_off; // EPI_SWIGLU: row offset of the up block inside A (= I)
  int raster_n;   // 1: consecutive tiles walk N first (A streamed once), 0: M first (B streamed once)
  int bulk_red;   // transposed fp32 atomic output via cp.reduce.async.bulk from a staged tile (decode products)
  int stream_k;   // 1: k-block units split evenly over the CTAs (see WorkIter); atomic f32 output, split_k == 1
  int swiglu_rows; // EPI_SWIGLU: decode rows staged per epilogue pass (64, or 32 for the two-CTAs-per-SM form at 128 rows)
  int smem_tight;  // dynamic shared memory requested without alignment slack (kernel traps unless the base is 1 KiB aligned)
  int epi;
  int c_f32, trans_c, accumulate, atomic;
  int bias_per_m;
  int a_static;   // A does not depend on the preceding kernel (weights): its first TMA loads may precede pdl_wait()
  float alpha;
  void* C;
  long long ldc, c_bs_lo, c_bs_hi;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* residual;  // same indexing as C
  // EPI_LSE / EPI_DLOGITS
  const int* labels;     // [M]
  float* part_max;       // [M, tiles_n]
  float* part_sum;       // [M, tiles_n]
  float* tgt_logit;      // [M]
  const float* lse;      // [M]
  const float* gscale;   // [M]
  __nv_bfloat16* gu_out; // EPI_SWIGLU_T: bf16 [M][2I] gate | up pre-activations kept for the backward (or nullptr);
                         // EPI_SWIGLU_BWD: the same buffer, read and overwritten with dgate | dup
  long long gu_ld;
  int tma_store;         // EPI_STORE through a tensor map of C: 1 = tile store (bf16 / f32), 2 = f32 reduce-add (accumulate)
};

}  // namespace iadr1
