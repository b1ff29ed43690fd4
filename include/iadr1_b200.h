/* iadr1_b200.h — C ABI of libiadr1_b200.so, the B200-native (sm_100a) replacement for the native work that sits under
 * IAD-R1's SC-GRPO / PA-SFT hot path.
 *
 * The reference has no FFI of its own: its hot path is three Python call sites into third-party wheels,
 *   - `model(**inputs).logits`            ref: train/stage_rl/trainer/sc_grpo_trainer.py:505   (HF VLM forward, cuBLAS + flash-attn)
 *   - `self.llm.generate(...)`            ref: train/stage_rl/trainer/sc_grpo_trainer.py:667   (vLLM rollout)
 *   - `Trainer.training_step` backward/optimizer (accelerate + DeepSpeed) around `compute_loss` (ibid. :586-819)
 * Each entry point below names the reference call it stands in for. Conventions: every pointer is a raw DEVICE pointer
 * owned by the caller (PyTorch owns all buffers; this library never allocates caller-visible memory), sizes are in
 * elements unless noted, `stream` is a cudaStream_t passed as void*, return value 0 = ok, negative = error with the
 * message available from iadr1_last_error() (thread-local). No torch types cross this boundary.
 */
#ifndef IADR1_B200_H_
#define IADR1_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library state ------------------------------------------------------------------------------------------- */
const char* iadr1_last_error(void);
int iadr1_version(void);
/* Number of kernels this library has launched since load / since the last reset (bench.py's `gpu_launches`). */
long long iadr1_launch_count(void);
void iadr1_reset_launch_count(void);

/* ---- dense products: replaces torch.nn.functional.linear / torch.matmul (cuBLAS) under the HF modules ---------
 * C[z] (+)= alpha * A[z] * B[z]^T, bf16 operands, fp32 accumulation on tcgen05 tensor cores.
 *   A is M x K, B is N x K. `*_mn = 0`: the operand is stored with K contiguous (x[M,K], torch Linear weight W[N,K]);
 *   `*_mn = 1`: stored with the M (resp. N) index contiguous, i.e. as a [K, M] (resp. [K, N]) row-major buffer.
 *   lda/ldb: elements between consecutive rows of the stored buffer. Batch index z = z_hi * batch_lo + z_lo addresses
 *   A + z_lo*a_bs_lo + z_hi*a_bs_hi, B + (z_lo / b_lo_div)*b_bs_lo + z_hi*b_bs_hi, C + z_lo*c_bs_lo + z_hi*c_bs_hi.
 * HF call sites covered: modeling_qwen2_5_vl.py:91-114 (patch embed), :214-286 (vision qkv/proj), :77-88, :611-624 (MLPs),
 *   :133-146 (merger), :704-707 (decoder q/k/v/o), :1519 (lm_head) and autograd's dgrad/wgrad of each.            */
typedef struct iadr1_gemm_t {
  int M, N, K;
  int batch, batch_lo, b_lo_div;
  const void* A; long long lda, a_bs_lo, a_bs_hi; int a_mn;
  const void* B; long long ldb, b_bs_lo, b_bs_hi; int b_mn;
  void* C; long long ldc, c_bs_lo, c_bs_hi;
  int c_f32;       /* 0: C is bf16, 1: C is fp32 */
  int trans_c;     /* store C^T (element (m,n) at C[n*ldc + m]) */
  int accumulate;  /* fp32 C += result (gradient accumulation across micro-steps) */
  int atomic;      /* fp32 C atomically += result (split-K) */
  int split_k;
  float alpha;
  const void* bias; int bias_per_m;  /* bf16 [N] (or [M] when bias_per_m) */
  const void* residual;              /* bf16, indexed like C */
  int kmode, skip_mode, causal_off;  /* causal trimming for attention products, see gemm_sm100.cuh */
  int epi;                           /* 0 store, 1 row log-sum-exp partials, 2 softmax-gradient (lm_head backward) */
  const int* labels; float* part_max; float* part_sum; float* tgt_logit; int lse_tiles_n;
  const float* lse; const float* gscale;
  int block_n, stages, max_ctas;     /* 0 = library heuristics */
} iadr1_gemm_t;
int iadr1_gemm_bf16(const iadr1_gemm_t* desc, void* stream);
/* block_n the library would choose for an N-wide product (sizes the EPI_LSE partial buffers). */
int iadr1_gemm_pick_block_n(int N, int b_mn);

#ifdef __cplusplus
}
#endif
#endif /* IADR1_B200_H_ */
