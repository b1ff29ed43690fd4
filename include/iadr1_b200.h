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
  int epi;                           /* 0 store, 1 row log-sum-exp partials, 2 softmax-gradient (lm_head backward), 3 SwiGLU (decode),
                                        4 SwiGLU (training: N = I features, B = fused gate|up weight [2I][K], C = act bf16 [M][I]),
                                        5 SwiGLU backward (A = dy [M][K = H], B = down weight (MN-major [H][I]), N = I: the product
                                          dact never leaves the SM - gu_out [M][2I] (gate | up) is rewritten with dgate | dup; C unused) */
  const int* labels; float* part_max; float* part_sum; float* tgt_logit; int lse_tiles_n;
  const float* lse; const float* gscale;
  int block_n, stages, max_ctas;     /* 0 = library heuristics */
  int a_static;                      /* A (weights) is not produced by the preceding kernel: prefetchable under PDL */
  int stream_k;                      /* split the K-block units evenly over all SMs (atomic fp32 C, split_k = 1) */
  int up_row_off;                    /* epi 3 (decode SwiGLU): A holds gate rows [0, I) then up rows [I, 2I); M = 2I, up_row_off = I;
                                        C is bf16 [N][ldc] with C[n][f] = silu(gate_f . b_n) * (up_f . b_n)            */
  int raster;                        /* tile order: 0 library heuristic, 1 M fastest (B streamed once), 2 N fastest (A streamed once) */
  int no_chunked_maps;               /* probes: MN-major operands as 64-column 2-D boxes (several TMA instructions per k-block) */
  int no_bulk_red;                   /* probes: force per-lane atomics instead of bulk reductions for transposed fp32 atomic C */
  int co_resident;                   /* decode chain: <= 113 KB smem, <= 128 regs, minimal TMEM so two CTAs fit per SM */
  void* gu_out; long long gu_ld;     /* epi 4: optional bf16 [M][2I] gate | up pre-activations kept for the backward;
                                        epi 5: the same buffer, read and overwritten in place with dgate | dup */
} iadr1_gemm_t;
int iadr1_gemm_bf16(const iadr1_gemm_t* desc, void* stream);
/* block_n the library would choose for an N-wide product (sizes the EPI_LSE partial buffers). */
int iadr1_gemm_pick_block_n(int N, int b_mn);
/* Programmatic dependent launch for the kernels launched while enabled (the rollout decode chain): each kernel's
 * launch + prologue + predecessor-independent prefetch overlaps the previous kernel's tail.                          */
int iadr1_set_pdl(int on);
/* Live roofline support: time every (non-graph-captured) GEMM launch with CUDA events on its own stream and count its
 * algorithmic FLOPs. Collect after a device synchronise.                                                            */
/* Decode-chain timeline probe: op 1 installs + clears a device trace buffer (thread 0 of CTA 0 of every PDL-aware
 * kernel then stamps %globaltimer before / after its dependency wait), op 0 removes it, op 2 copies it to `out`
 * (word 0 = record count, records {reach, leave} from word 2). Diagnostics only; off by default.                   */
int iadr1_trace(int op, unsigned long long* out, int max_words);
int iadr1_gemm_profile_enable(int on);
int iadr1_gemm_profile_collect(double* total_ms, double* total_flops, double* max_launch_ms, long long* launches,
                               const char* csv_path /* optional: per-shape breakdown */);

/* ---- row kernels (HBM-bound; bf16 activations, fp32 statistics) ------------------------------------------------
 * RMSNorm: HF Qwen2_5_VLRMSNorm.forward, modeling_qwen2_5_vl.py:57-71 (decoder :775-776,:849; vision blocks; merger ln_q).
 * rstd (optional, [rows] fp32) is saved for the backward. cols % 8 == 0.                                           */
int iadr1_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, long long rows, int cols, long long x_ld,
                      long long y_ld, float eps, void* stream);
/* dx (+)= d(rmsnorm)/dx, dw (fp32, may be NULL) += d/dw. add_dx = 1 accumulates into dx (residual-stream gradient). */
int iadr1_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, void* dx, float* dw,
                      long long rows, int cols, long long ld, int add_dx, void* stream);
/* LayerNorm: Qwen2-VL vision blocks (modeling_qwen2_vl.py:461-466) and SigLIP (LLaVA-OneVision tower).            */
int iadr1_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mean, float* rstd, long long rows,
                        int cols, long long ld, float eps, void* stream);
int iadr1_layernorm_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd, void* dx,
                        float* dw, float* db, long long rows, int cols, long long ld, int add_dx, void* stream);
/* Rotate-half rotary embedding in place on the first `heads` heads of x[tokens][heads_total][hd] (tok_stride elements
 * per token). cos/sin: fp32 [tokens][hd]. bf16_ops = 1 reproduces apply_multimodal_rotary_pos_emb (:627-669, bf16
 * products), 0 reproduces apply_rotary_pos_emb_vision (:156-167, fp32). backward = 1 applies the transpose.         */
int iadr1_rope(void* x, const float* cos_t, const float* sin_t, long long tokens, int heads, int hd,
               long long tok_stride, int bf16_ops, int backward, void* stream);
/* out = act(gate) * up with gate at gu[r*ld + c], up at gu[r*ld + up_off + c]; up_off < 0: out = act(gate).
 * act: 0 SiLU (Qwen2MLP :611-624, vision SwiGLU :77-88), 1 exact GELU (merger :139), 2 quick-GELU, 3 tanh-GELU.       */
int iadr1_act_mul_fwd(const void* gu, void* out, long long rows, int cols, long long ld, long long up_off,
                      long long out_ld, int act, void* stream);
int iadr1_act_mul_bwd(const void* dout, const void* gu, void* dgu, long long rows, int cols, long long ld,
                      long long up_off, long long dout_ld, int act, void* stream);
/* In-place masked softmax over bf16 scores S[z][q][k]; row q keeps keys [lo[q], hi[q]) (causal / window / padding
 * masks as ranges), exact zeros elsewhere. eager_attention_forward, modeling_qwen2_5_vl.py:182-206.                 */
/* Keys in [hole_lo, hole_hi) are always masked (padding between the shared-prefix and per-row key segments).          */
int iadr1_softmax_rows(void* S, const int* lo, const int* hi, int Tq, int Tk, long long ld, long long z_stride,
                       long long batch, int hole_lo, int hole_hi, void* stream);
int iadr1_softmax_bwd_rows(const void* P, void* dP, const int* lo, const int* hi, int Tq, int Tk, long long ld,
                           long long z_stride, long long batch, int hole_lo, int hole_hi, void* stream);
/* out[r] = index[r] >= 0 ? table[index[r]] : alt[-1 - index[r]]: embed_tokens + masked_scatter of image embeddings
 * (:1298-1307) in one pass; with alt = NULL a plain row gather (vision window reorder :478-484, :512-513).          */
int iadr1_gather_rows(const void* table, const void* alt, const int* index, void* out, long long rows, int cols,
                      long long table_ld, long long alt_ld, long long out_ld, void* stream);
int iadr1_scatter_add_rows(const void* d, const int* index, float* dtable, float* dalt, long long rows, int cols,
                           long long d_ld, long long table_ld, long long alt_ld, void* stream);
/* out[c] += sum_r x[r][c] (bias gradients).                                                                       */
int iadr1_colsum(const void* x, float* out, long long rows, int cols, long long ld, void* stream);
/* Sum the g query-head gradients of each kv head (autograd of repeat_kv, modeling_qwen2_5_vl.py:170-179).          */
int iadr1_group_sum(const void* src, void* out, long long rows, int nkv, int g, int hd, long long src_ld,
                    long long out_ld, int accumulate, void* stream);
int iadr1_add_bf16(const void* a, const void* b, void* out, long long n, void* stream);
int iadr1_cast_f32_bf16(const float* src, void* dst, long long n, void* stream);
/* Finishes the fused lm_head -> log-softmax -> gather (replaces sc_grpo_trainer.py:505-514 / trl selective_log_softmax,
 * ref: trl/trl/trainer/utils.py:1683-1715) from the GEMM's per-tile (max, sum-exp) partials.                       */
int iadr1_lse_finalize(const float* pmax, const float* psum, const float* tgt, int tiles_n, int M, float* lse,
                       float* logp, void* stream);

/* ---- image preprocessing on the GPU (Qwen2-VL / Qwen2.5-VL): replaces the CPU image half of the processor call
 * (ref: train/stage_rl/trainer/sc_grpo_trainer.py:614-621; HF image_processing_qwen2_vl.py:148-232). rgb_u8: device uint8
 * [in_h][in_w][3]; the image is resized to out_h x out_w (multiples of patch * merge, from smart_resize on the host) with
 * Pillow's antialiased bicubic, rescaled by 1/255, normalised with mean3 / std3 (HOST pointers to 3 floats) and written as
 * bf16 patch rows [out_h/patch * out_w/patch][3 * tps * patch * patch] in merge-block-major order. scratch_u8: device
 * uint8 [in_h * out_w * 3 + out_h * out_w * 3], needed only when the size changes.                                    */
int iadr1_image_preprocess_qwen(const void* rgb_u8, int in_h, int in_w, int out_h, int out_w, int patch, int merge, int tps,
                                const float* mean3, const float* std3, void* scratch_u8, void* out_bf16, void* stream);

/* ---- the one data-path collective: gradient all-reduce over NVLink ---------------------------
 * Replaces ZeRO-3's per-parameter reduce-scatter / all-gather (ref: scripts/train/zero3.json:14-33). NCCL is bound at run time
 * (the libnccl.so.2 the process carries); `comm` is an ncclComm_t created here from a 128-byte ncclUniqueId that rank 0
 * generates and the host distributes. iadr1_grad_allreduce sums grad[0, n) in place, cut into buckets of bucket_elems
 * (0 = one call), ordered on `stream` - the trainer enqueues each decoder layer's range on a side stream as soon as the
 * backward sweep retires the layer, so the transfer overlaps the rest of the backward.                               */
int iadr1_comm_unique_id(void* out128);
int iadr1_comm_create(const void* unique_id128, int rank, int world, void** comm_out);
int iadr1_comm_destroy(void* comm);
int iadr1_grad_allreduce(void* comm, float* grad, long long n, long long bucket_elems, void* stream);

/* ---- optimizer: replaces torch.optim.AdamW inside DeepSpeed ZeRO-3 + clip_grad_norm_ ------------ */
int iadr1_sumsq_f32(const float* g, long long n, float* out, void* stream);
int iadr1_adamw_step(float* p32, void* p16, float* g, float* m, float* v, long long n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int step, float grad_scale, const float* sumsq,
                     float max_norm, int zero_grad, void* stream);
/* The same step with exp_avg / exp_avg_sq stored in bf16 under stochastic rounding (seeded counter hash): 8 bytes per
 * parameter less optimizer state - the form that fits Qwen2.5-VL-7B on one 180 GB GPU without sharding (the reference shards
 * with ZeRO-3, ref: scripts/train/zero3.json:14-33).                                                                   */
int iadr1_adamw_step_bf16m(float* p32, void* p16, float* g, void* m16, void* v16, long long n, float lr, float beta1,
                           float beta2, float eps, float weight_decay, int step, float grad_scale, const float* sumsq,
                           float max_norm, int zero_grad, unsigned long long seed, void* stream);

/* ---- rollout: replaces vLLM `LLM.generate` (sc_grpo_trainer.py:343-358, 667). State words: [0] step,
 * [2] unfinished rows; per-row prompt lengths in row_plen. All per-step inputs live on the device so one step is CUDA-graph replayable.        */
int iadr1_decode_embed(const void* embed, const int* tok, float* h, int rows, int H, void* stream);
/* zero_buf (optional): rows x zero_per_row fp32 cleared in the same launch (the next split-K GEMM's target).        */
int iadr1_rmsnorm_f32in(const float* x, const void* w, void* y, int rows, int cols, float eps, float* zero_buf,
                        int zero_per_row, void* stream);
/* One launch: rotary on q/k + KV append + split-KV attention over (shared prompt prefix, row slab) + split merge.
 * head_dim 64 / 128: tensor-core kernel, nsplit > 0 = 4-warp CTAs (64-key chunks), nsplit < 0 = |nsplit| splits with 2-warp
 * CTAs (32-key chunks); other head sizes: scalar kernel, nsplit = ceil((p_max + c_max) / 128).
 * part: fp32 [rows][nq][|nsplit|][hd + 2] scratch; tickets: int32 [rows * nkv], zero before the first call.           */
int iadr1_decode_attention_fused(const float* qkv, const float* cos_tab, const float* sin_tab, const int* rope_delta,
                                 const void* kp, const void* vp, void* kc, void* vc, const int* state,
                                 const int* row_group, const int* row_plen, const int* finished, float* part, int* tickets,
                                 void* out, int rows, int nq, int nkv, int hd, int p_max, int c_max, int nsplit, int max_pos,
                                 float scale, void* stream);   /* finished (may be NULL): rows that have produced EOS are skipped */
/* Shared-prefix form of the same step: the prompt keys are processed ONCE per group (8 rows x gq heads = full tensor-core
 * tiles; K/V staged once for the group's rows), each row's own keys per row; `part` holds psplit + csplit slots per (row, head),
 * `tickets` one counter per (row, kv head). Rows of a group are consecutive and share the prompt (vLLM prefix caching,
 * ref: train/stage_rl/trainer/sc_grpo_trainer.py:351). head_dim 64 or 128.                                              */
int iadr1_decode_attention_grouped(const float* qkv, const float* cos_tab, const float* sin_tab, const int* rope_delta,
                                   const void* kp, const void* vp, void* kc, void* vc, const int* state, const int* row_plen,
                                   const int* finished, float* part, int* tickets, void* out, int rows, int rows_per_group, int nq,
                                   int nkv, int hd, int p_max, int c_max, int psplit, int csplit, int max_pos, float scale,
                                   void* stream);
/* temperature -> top-k (ties kept) -> top-p -> multinomial; SamplingParams at sc_grpo_trainer.py:353-358.
 * The Philox seed is `seed ^ (state[4] | state[5] << 32)`: callers that replay a captured graph keep the per-call seed
 * in the device-resident state words and pass seed = 0 (graph arguments are frozen at capture).                     */
int iadr1_sample(const float* logits, int rows, int V, float temperature, int top_k, float top_p,
                 unsigned long long seed, int* state, int* tok, int* finished, int* out_tokens, int c_max, int eos_id,
                 int pad_id, int forbid_eos, int first, void* stream);
/* Decode-step SwiGLU on the fp32 gate|up accumulator [rows][2*cols] of the stream-K gate_up product (Qwen2MLP,
 * modeling_qwen2_5_vl.py:611-624): out bf16 [rows][cols]; the accumulator is zeroed for the next layer.              */
int iadr1_decode_silu_mul_f32(float* gu, void* out, int rows, int cols, void* stream);
int iadr1_decode_advance(int* state, void* stream);

/* ---- fused attention (tcgen05, scores never leave the SM): replaces the flash-attn forward / backward the reference selects
 * with `--attn_implementation flash_attention_2` (ref: scripts/train/SC_GRPO/SC_GRPO_Qwen_Instruct_2_5_VL_3B.sh:58; HF
 * modeling_qwen2_5_vl.py:214-286 vision attention, :704-760 decoder attention).
 *   qkv     bf16 [n_tokens][(nq + 2 nkv) * hd]: q heads, then k heads, then v heads of every token (post-rotary)
 *   ranges  int32 [>= n_tokens + 64][4] (tail rows zero): query row t attends keys [lo, hi) u [plo, phi) (token indices into the same buffer; the two
 *           ranges must be disjoint): causal rows, the shared-prefix GRPO layout, vision windows / crops, packed groups
 *   items   forward / dQ work items, int32 [n][6] = {q0, nrows <= 128, kv0, kv1, p0, p1}: rows [q0, q0 + nrows) walk the
 *           key tiles of [p0, p1) (masked by plo / phi) and of [kv0, kv1) (masked by lo / hi); built by the host
 *   sched   int32 [n_cta + 1 + n_units]: offsets of every CTA into the unit list that follows; unit = item * nq + head
 *           (the host balances the CTAs: the cost of every unit is known from its item)
 *   out     bf16 [n_tokens][nq * hd];  lse2 fp32 [nq][npad] (head-major): log2-domain log-sum-exp of the scaled scores;
 *           npad = a multiple of 4 >= n_tokens + 64 (the backward bulk-copies 64 consecutive rows of a head)
 *   pv_n    0 = library choice; N of the P.V-type products (multiple of 16 in [hd, 128])                                   */
int iadr1_fmha_fwd(const void* qkv, long long n_tokens, int nq, int nkv, int hd, const int* ranges, const int* items,
                   int n_items, const int* sched, int n_cta, void* out, float* lse2, long long npad, float scale, int pv_n,
                   void* stream);
/* Timeline probe (tools/fmha_probe.py --trace): device buffer of 3 * 64 * 8 int64 that CTA 0 of the dK/dV kernel fills with
 * clock64() stamps per role and iteration; NULL switches it off.                                                        */
int iadr1_fmha_set_trace(long long* device_buf);
/* Backward: dqkv bf16 [n_tokens][(nq + 2 nkv) * hd] receives dQ | dK | dV (GQA heads and shared-prefix rows reduced inside).
 *   k_items int32 [n][6] = {k0, nkeys <= 128, q0, q1, f0, f1}: key tile [k0, k0 + nkeys) and the query rows [q0, q1) that may
 *           attend to it (q0 % 4 == 0; a key tile may appear in several items with disjoint query ranges; their results are
 *           added); 64-query tiles f0 <= t < f1 of the range are allowed in full (no mask arithmetic)
 *   q_sched / k_sched: CTA schedules of the dQ kernel (unit = item * nq + head) and the dK/dV kernel (unit = item * nkv + kv head)
 *   delta   fp32 [nq][npad] scratch;  dkv32 fp32 [n_tokens][2 * nkv * hd] scratch (zeroed and filled here)                    */
int iadr1_fmha_bwd(const void* qkv, const void* dout, const void* out, const float* lse2, long long n_tokens, int nq, int nkv,
                   int hd, const int* ranges, const int* q_items, int n_q_items, const int* q_sched, int n_q_cta,
                   const int* k_items, int n_k_items, const int* k_sched, int n_k_cta, void* dqkv, float* delta, float* dkv32,
                   long long npad, float scale, int pv_n, void* stream);

/* ==== model-level entry points ("B-inner"): one call per reference call site =======================
 * Weights are registered ONCE by pointer under the names of the flat parameter store ("embed_tokens.weight",
 * "layers.<i>.{qkv,o,gate_up,down,ln1,ln2}.weight", "layers.<i>.qkv.bias", "norm.weight", optional "lm_head.weight");
 * grad_f32 (may be NULL for a frozen model) is the fp32 gradient view the backward accumulates into.
 * Workspaces are caller-owned; *_workspace_bytes reports the size for a token count. One handle per process / GPU, calls
 * serialised on the passed stream, no global state besides the thread-local error string.                            */
typedef struct iadr1_model_cfg_t {
  int vocab, hidden, inter, layers, nq, nkv, hd;   /* decoder geometry (HF config.json: vocab_size, hidden_size, ...)   */
  float rms_eps;
  /* vision tower: v_kind -1 none, 0 qwen2_5_vl (RMSNorm, SwiGLU + bias, windowed / full attention), 1 qwen2_vl (LayerNorm,
   * quick-GELU MLP), 2 siglip (LLaVA-OneVision: learned positions, LayerNorm, tanh-GELU MLP, projector + anyres packing),
 * 3 clip (LLaVA-1.5: class token as a unit-vector pixel row, learned positions, pre-LayerNorm, quick-GELU MLP; v_depth = the
 * blocks actually run for vision_feature_layer; the packing gather drops the class token).
   * v_inter / v_patch_dim are the 8-element padded widths of the store; v_fullatt_mask bit i = block i attends the whole image */
  int v_kind, v_depth, v_hidden, v_heads, v_inter, v_out_hidden, v_patch_dim, v_merge_unit, v_tokens_per_crop;
  unsigned int v_fullatt_mask;
  float v_eps;
} iadr1_model_cfg_t;
/* Device-resident tables of one token layout for the fused attention (built by the host, see iadr1_fmha_fwd / _bwd). */
typedef struct iadr1_attn_plan_t {
  const int* ranges; const int* q_items; int n_q; const int* sched_fwd; int n_cta_fwd; const int* sched_dq; int n_cta_dq;
  const int* k_items; int n_k; const int* sched_kv; int n_cta_kv; long long npad; long long n_tokens;
} iadr1_attn_plan_t;
/* Rollout prefill: post-rotary K / V of prompt g (rows [g * p_len, (g + 1) * p_len) of the token stream) go to
 * kp / vp [layer][group][p_max][nkv][hd]; layer_stride = elements between consecutive layers.                        */
typedef struct iadr1_kv_sink_t { void* kp; void* vp; int n_groups; int p_len; int p_max; long long layer_stride; } iadr1_kv_sink_t;
/* State of the in-rank rollout engine (all device pointers; R rows decode in lock-step, see the decode-step kernels). */
typedef struct iadr1_decode_t {
  int R, n_groups, p_max, c_max, nsplit, max_pos, block_n;
  const void* kp; const void* vp; void* kc; void* vc;       /* [layers][n_groups][p_max][nkv][hd], [layers][R][c_max][nkv][hd] */
  int* state; int* tok; int* finished; int* out_tokens; const int* rope_delta; const int* row_plen; const int* row_group;
  float* h; void* xn; float* qkv; void* attn; float* part; int* tickets; void* act; float* logits;
  const float* cos_tab; const float* sin_tab;
  float temperature; int top_k; float top_p; int eos_id, pad_id, forbid_eos;
  unsigned int* chain_counters;   /* (layers + 1) * 8 words for the persistent decode-layer chain (R <= 128); NULL: one kernel per op */
  int attn_psplit;                /* > 0: shared-prefix decode attention with this many prompt splits; nsplit - attn_psplit completion splits */
} iadr1_decode_t;
typedef void (*iadr1_layer_cb)(int layer, void* user);
/* What the vision tower derives from the image grids (host-computed, device-resident; HF get_window_index / rot_pos_emb,
 * modeling_qwen2_5_vl.py:382-453; pack_image_features, modeling_llava_onevision.py:292-355).                          */
typedef struct iadr1_vision_geom_t {
  long long n_patches;                 /* rows of pixel_values                                                       */
  long long n_out;                     /* rows of the output: merged tokens (Qwen) / packed image tokens (LLaVA-OV)  */
  const float* cos; const float* sin;  /* 2-D rotary tables fp32 [n_patches][head_dim]; NULL for SigLIP              */
  const int* window_index;             /* Qwen2.5: merge units in window order [n_patches / unit]; else NULL         */
  const int* reverse_index;            /* Qwen2.5: its inverse; else NULL                                            */
  const iadr1_attn_plan_t* plan_full;  /* attention inside whole images / crops                                      */
  const iadr1_attn_plan_t* plan_win;   /* attention inside windows (Qwen2.5); else NULL                              */
  const int* pos_index;                /* SigLIP: row of the learned position table per patch                        */
  const int* pack_index;               /* SigLIP: packed token -> projected feature row, or -1 = image_newline       */
} iadr1_vision_geom_t;

int iadr1_model_create(const iadr1_model_cfg_t* cfg, void** handle);
int iadr1_model_destroy(void* handle);
int iadr1_bind_weights(void* handle, const char* name, const void* param_bf16, float* grad_f32);
/* Decoder stack over a flat token stream (replaces the decoder half of `model(**inputs).logits`,
 * ref: train/stage_rl/trainer/sc_grpo_trainer.py:505; HF modeling_qwen2_5_vl.py:778-827 per layer).
 * src_index[t] >= 0: token id, < 0: row -1 - src_index[t] of image_embeds. mode 0: forward only; 1: every activation the
 * backward needs stays in the workspace; 2: only layer inputs stay (the backward recomputes each layer).
 * *h_last points (inside the workspace) at the last hidden states [n_tokens][hidden] (pre final norm).               */
int iadr1_decoder_workspace_bytes(void* handle, long long n_tokens, long long npad, int mode, long long* bytes);
int iadr1_decoder_fwd(void* handle, const int* src_index, const void* image_embeds, long long n_tokens, const float* cos_t,
                      const float* sin_t, const iadr1_attn_plan_t* plan, void* workspace, int mode,
                      const iadr1_kv_sink_t* sink, void** h_last, void* stream);
/* dh bf16 [n_tokens][hidden] is consumed in place; weight gradients accumulate into the bound fp32 views; dimg32 (may be
 * NULL) receives d(image_embeds); on_layer_done(layer, user) fires on the host after layer `layer`'s kernels are enqueued
 * (its weight gradients are then final in stream order: the trainer hands the range to iadr1_grad_allreduce).          */
int iadr1_decoder_bwd(void* handle, void* dh, const int* src_index, const float* cos_t, const float* sin_t,
                      const iadr1_attn_plan_t* plan, void* workspace, int mode, long long n_tokens, float* dimg32,
                      iadr1_layer_cb on_layer_done, void* cb_user, void* stream);
/* Vision tower + merger / projector (replaces `model.visual(pixel_values, grid_thw)` and the LLaVA-OneVision
 * `get_image_features` + `pack_image_features`; HF modeling_qwen2_5_vl.py:455-518, modeling_llava_onevision.py:292-417).
 * pixel_values bf16 [n_patches][v_patch_dim]; out bf16 [n_out][v_out_hidden]. save = 1 keeps what iadr1_vision_bwd needs.  */
int iadr1_vision_workspace_bytes(void* handle, long long n_patches, long long n_out, long long npad, int save, long long* bytes);
int iadr1_vision_fwd(void* handle, const void* pixel_values, const iadr1_vision_geom_t* geo, void* workspace, int save, void* out,
                     void* stream);
int iadr1_vision_bwd(void* handle, const void* d_out, const void* pixel_values, const iadr1_vision_geom_t* geo, void* workspace,
                     void* stream);
/* hidden -> selected-token log-prob through final norm + fused lm_head (never materialising [rows, vocab] logits in the
 * forward): replaces `_get_per_token_logps` (ref: sc_grpo_trainer.py:505-514). The backward runs in the forward's workspace. */
int iadr1_logprob_workspace_bytes(void* handle, long long m_rows, int backward, long long* bytes);
int iadr1_logprob_fwd(void* handle, const void* h, const int* sel_index, const int* labels, long long m_rows, float temperature,
                      void* workspace, float* logp, void* stream);
int iadr1_logprob_bwd(void* handle, const float* dlogp, const int* sel_index, const int* labels, long long m_rows,
                      float temperature, void* workspace, float* dh32, void* stream);
/* SFT head: per-token cross-entropy nll = -log p(label) (HF ForCausalLMLoss; ref: llamafactory/train/sft/trainer.py:92-107). */
int iadr1_ce_fwd(void* handle, const void* h, const int* sel_index, const int* labels, long long m_rows, void* workspace,
                 float* nll, void* stream);
int iadr1_ce_bwd(void* handle, const float* dnll, const int* sel_index, const int* labels, long long m_rows, void* workspace,
                 float* dh32, void* stream);
/* Rollout (replaces `self.llm.generate`, ref: sc_grpo_trainer.py:343-358, 667): prefill = forward-only decoder pass over the
 * right-padded prompts that fills the shared-prefix KV cache; decode_head = final norm + lm_head + sampler on e->h;
 * decode_step = one lock-step token for all R rows (the body the engine captures into a CUDA graph).                   */
int iadr1_prefill(void* handle, const int* src_index, const void* image_embeds, long long n_tokens, const float* cos_t,
                  const float* sin_t, const iadr1_attn_plan_t* plan, void* workspace, const iadr1_kv_sink_t* sink, void** h_last,
                  void* stream);
int iadr1_decode_head(void* handle, const iadr1_decode_t* e, int first, void* stream);
int iadr1_decode_step(void* handle, const iadr1_decode_t* e, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IADR1_B200_H_ */
