// Model-level entry points of the C ABI (SURVEY.md §8b "B-inner"): the layer loops of the decoder forward / backward,
// the fused lm_head log-prob / cross-entropy head and the rollout's decode step as SINGLE calls over weights registered
// once by pointer. They stand in for the three Python call sites of the reference's hot path:
//   `model(**inputs).logits`     ref: train/stage_rl/trainer/sc_grpo_trainer.py:505   -> iadr1_decoder_fwd + iadr1_logprob_fwd
//   `loss.backward()`            (HF Trainer.training_step)                           -> iadr1_logprob_bwd + iadr1_decoder_bwd
//   `self.llm.generate(...)`     ref: train/stage_rl/trainer/sc_grpo_trainer.py:667   -> iadr1_prefill + iadr1_decode_step
// Conventions as everywhere in this library: raw device pointers owned by the caller, a caller-provided workspace whose
// size the library reports, nothing allocated here, work ordered on the passed stream.
#include "runtime.h"

#include <cstdio>
#include <string>
#include <unordered_map>
#include <vector>

namespace iadr1 {

typedef unsigned short bf16_t;   // storage only

struct Weight {
  const void* p = nullptr;
  float* g = nullptr;
};

struct Model {
  iadr1_model_cfg_t c;
  std::unordered_map<std::string, Weight> w;
  const Weight* find(const std::string& n) const {
    auto it = w.find(n);
    return it == w.end() ? nullptr : &it->second;
  }
};

struct Arena {      // bump allocator over the caller's workspace (base == nullptr: size computation only)
  char* base;
  size_t off = 0;
  explicit Arena(void* b) : base(static_cast<char*>(b)) {}
  void* take(size_t bytes) {
    void* p = base ? base + off : nullptr;
    off += (bytes + 255) & ~size_t(255);
    return p;
  }
};

// ---- dense products over row-major buffers (the forms ops.py uses) --------------------------------------------------
static int linear_fwd(const void* x, const void* W, void* out, long long M, int N, int K, const void* bias, const void* residual,
                      cudaStream_t s) {
  iadr1_gemm_t d;
  memset(&d, 0, sizeof(d));
  d.M = (int)M; d.N = N; d.K = K;
  d.batch = d.batch_lo = d.b_lo_div = 1;
  d.A = x; d.lda = K;
  d.B = W; d.ldb = K;
  d.C = out; d.ldc = N;
  d.split_k = 1; d.alpha = 1.f;
  d.bias = bias; d.residual = residual;
  return launch_gemm(d, s);
}
// dx[M, K] = dy[M, N] @ W[N, K]
static int linear_dgrad(const void* dy, const void* W, void* dx, long long M, int N, int K, cudaStream_t s) {
  iadr1_gemm_t d;
  memset(&d, 0, sizeof(d));
  d.M = (int)M; d.N = K; d.K = N;
  d.batch = d.batch_lo = d.b_lo_div = 1;
  d.A = dy; d.lda = N;
  d.B = W; d.ldb = K; d.b_mn = 1;
  d.C = dx; d.ldc = K;
  d.split_k = 1; d.alpha = 1.f;
  return launch_gemm(d, s);
}
// dW32[N, K] += dy[M, N]^T @ x[M, K]
static int linear_wgrad(const void* dy, const void* x, float* dW, long long M, int N, int K, cudaStream_t s) {
  if (!dW) return 0;
  iadr1_gemm_t d;
  memset(&d, 0, sizeof(d));
  d.M = N; d.N = K; d.K = (int)M;
  d.batch = d.batch_lo = d.b_lo_div = 1;
  d.A = dy; d.lda = N; d.a_mn = 1;
  d.B = x; d.ldb = K; d.b_mn = 1;
  d.C = dW; d.ldc = K; d.c_f32 = 1; d.accumulate = 1;
  d.split_k = 1; d.alpha = 1.f;
  return launch_gemm(d, s);
}

// gu[M, 2I] (gate | up) <- (dgate | dup) for dact = dy[M, H] @ Wdown[H, I]: the down-projection input gradient with the SwiGLU
// backward in its epilogue (epi 5) - replaces linear_dgrad + iadr1_act_mul_bwd, dact is never written.
// Measured (tools/gemm_tile_probe.py swiglu_bwd): the four epilogue warps need ~23 us per 128 x 256 tile for the SwiGLU backward
// arithmetic, the MMA of a tile takes 15 us at K = H = 2048 (3B: fused 498 us vs 458 us for GEMM + row kernel) and 27 us at
// H = 3584 (7B: 1133 vs 1279 us) - so the fused form is used where the reduction is long enough to hide it.
// IADR1_SWIGLU_BWD_EPILOGUE: 0 never, 1 (default) H >= 3072, 2 always.
static bool swiglu_bwd_epilogue_enabled(int H) {
  static const int mode = [] { const char* e = getenv("IADR1_SWIGLU_BWD_EPILOGUE"); return e && e[0] >= '0' && e[0] <= '2' ? e[0] - '0' : 1; }();
  return mode == 2 || (mode == 1 && H >= 3072);
}
static int linear_dgrad_swiglu_bwd(const void* dy, const void* Wdown, void* gu, long long M, int H, int I, cudaStream_t s) {
  iadr1_gemm_t d;
  memset(&d, 0, sizeof(d));
  d.M = (int)M; d.N = I; d.K = H;
  d.batch = d.batch_lo = d.b_lo_div = 1;
  d.A = dy; d.lda = H;
  d.B = Wdown; d.ldb = I; d.b_mn = 1;
  d.split_k = 1; d.alpha = 1.f; d.epi = 5;
  d.gu_out = gu; d.gu_ld = 2LL * I;
  return launch_gemm(d, s);
}

__global__ void scale_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n, float scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i] * scale;
}
static int scale_f32(const float* src, float* dst, long long n, float scale, cudaStream_t s) {
  if (n <= 0) return 0;
  long long blocks = (n + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  scale_f32_kernel<<<(int)blocks, 256, 0, s>>>(src, dst, n, scale);
  IADR1_CHECK_LAUNCH("scale_f32");
  return 0;
}

#define TRY(expr)            \
  do {                       \
    int rc__ = (expr);       \
    if (rc__) return rc__;   \
  } while (0)

// ---- decoder ---------------------------------------------------------------------------------------------------------
struct LayerBuf {
  float* r1; void* xn; void* qkv; float* lse2; void* a_out; void* h_mid; float* r2; void* xn2; void* gu; void* act;
};
struct DecoderLayout {
  std::vector<void*> h;          // h[i] = input of layer i, h[L] = output
  std::vector<LayerBuf> lb;      // one per layer (mode 1) or a single shared one (modes 0, 2)
  void* dact; void* dx; void* dattn; void* dqkv; float* delta; float* dkv32;   // backward scratch (modes 1, 2)
  size_t bytes;
};

static LayerBuf take_layer(Arena& a, const iadr1_model_cfg_t& c, long long N, long long npad) {
  const long long H = c.hidden, I = c.inter, D = (long long)(c.nq + 2 * c.nkv) * c.hd, QH = (long long)c.nq * c.hd;
  LayerBuf b;
  b.r1 = static_cast<float*>(a.take(N * 4));
  b.xn = a.take(N * H * 2);
  b.qkv = a.take(N * D * 2);
  b.lse2 = static_cast<float*>(a.take((size_t)c.nq * npad * 4));
  b.a_out = a.take(N * QH * 2);
  b.h_mid = a.take(N * H * 2);
  b.r2 = static_cast<float*>(a.take(N * 4));
  b.xn2 = a.take(N * H * 2);
  b.gu = a.take(N * 2 * I * 2);
  b.act = a.take(N * I * 2);
  return b;
}

// mode 0: forward only (two ping-pong hidden buffers, one layer scratch); 1: every activation resident; 2: layer inputs
// resident + one layer scratch (recompute in the backward)
static DecoderLayout decoder_layout(const iadr1_model_cfg_t& c, long long N, long long npad, int mode, void* base) {
  Arena a(base);
  DecoderLayout L;
  const long long H = c.hidden, I = c.inter, D = (long long)(c.nq + 2 * c.nkv) * c.hd, QH = (long long)c.nq * c.hd;
  L.h.resize(c.layers + 1);
  if (mode == 0) {
    void* p0 = a.take(N * H * 2);
    void* p1 = a.take(N * H * 2);
    for (int i = 0; i <= c.layers; ++i) L.h[i] = (i & 1) ? p1 : p0;
  } else {
    for (int i = 0; i <= c.layers; ++i) L.h[i] = a.take(N * H * 2);
  }
  if (mode == 1) {
    for (int i = 0; i < c.layers; ++i) L.lb.push_back(take_layer(a, c, N, npad));
  } else {
    L.lb.push_back(take_layer(a, c, N, npad));
  }
  L.dact = L.dx = L.dattn = L.dqkv = nullptr;
  L.delta = L.dkv32 = nullptr;
  if (mode != 0) {
    L.dact = a.take(N * I * 2);
    L.dx = a.take(N * H * 2);
    L.dattn = a.take(N * QH * 2);
    L.dqkv = a.take(N * D * 2);
    L.delta = static_cast<float*>(a.take((size_t)c.nq * npad * 4));
    L.dkv32 = static_cast<float*>(a.take((size_t)N * 2 * c.nkv * c.hd * 4));
  }
  L.bytes = a.off;
  return L;
}

struct LayerW {
  const Weight *ln1, *qkv, *qkv_b, *o, *ln2, *gu, *down;
};
static int layer_weights(const Model& m, int i, LayerW& w) {
  const std::string b = "layers." + std::to_string(i) + ".";
  w.ln1 = m.find(b + "ln1.weight"); w.qkv = m.find(b + "qkv.weight"); w.qkv_b = m.find(b + "qkv.bias");
  w.o = m.find(b + "o.weight"); w.ln2 = m.find(b + "ln2.weight"); w.gu = m.find(b + "gate_up.weight");
  w.down = m.find(b + "down.weight");
  if (!w.ln1 || !w.qkv || !w.o || !w.ln2 || !w.gu || !w.down)      // qkv.bias is optional (LLaMA / Vicuna have none)
    return set_error("decoder layer %d: weights not bound (iadr1_bind_weights)", i);
  return 0;
}

static int layer_fwd(const Model& m, int i, const void* h_in, void* h_out, const LayerBuf& b, bool save, long long N,
                     const float* cos, const float* sin, const iadr1_attn_plan_t* plan, const iadr1_kv_sink_t* sink,
                     cudaStream_t s) {
  const iadr1_model_cfg_t& c = m.c;
  const int H = c.hidden, I = c.inter, D = (c.nq + 2 * c.nkv) * c.hd, QH = c.nq * c.hd;
  LayerW w;
  TRY(layer_weights(m, i, w));
  TRY(iadr1_rmsnorm_fwd(h_in, w.ln1->p, b.xn, save ? b.r1 : nullptr, N, H, H, H, c.rms_eps, s));
  TRY(linear_fwd(b.xn, w.qkv->p, b.qkv, N, D, H, w.qkv_b ? w.qkv_b->p : nullptr, nullptr, s));
  TRY(iadr1_rope(b.qkv, cos, sin, N, c.nq + c.nkv, c.hd, D, 1, 0, s));
  if (sink && sink->kp) {
    // rollout prefill: post-rotary K / V rows of every prompt go to the shared-prefix cache [layer][group][p_max][nkv][hd]
    const long long kvw = (long long)c.nkv * c.hd;
    for (int gi = 0; gi < sink->n_groups; ++gi) {
      const char* src = static_cast<const char*>(b.qkv) + ((long long)gi * sink->p_len * D + QH) * 2;
      char* kd = static_cast<char*>(sink->kp) + ((long long)i * sink->layer_stride + (long long)gi * sink->p_max * kvw) * 2;
      char* vd = static_cast<char*>(sink->vp) + ((long long)i * sink->layer_stride + (long long)gi * sink->p_max * kvw) * 2;
      if (cudaMemcpy2DAsync(kd, kvw * 2, src, (size_t)D * 2, kvw * 2, sink->p_len, cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
          cudaMemcpy2DAsync(vd, kvw * 2, src + kvw * 2, (size_t)D * 2, kvw * 2, sink->p_len, cudaMemcpyDeviceToDevice, s) != cudaSuccess)
        return set_error("prefill: KV copy failed");
    }
  }
  const float scale = 1.f / sqrtf((float)c.hd);
  TRY(iadr1_fmha_fwd(b.qkv, N, c.nq, c.nkv, c.hd, plan->ranges, plan->q_items, plan->n_q, plan->sched_fwd, plan->n_cta_fwd,
                     b.a_out, b.lse2, plan->npad, scale, 0, s));
  TRY(linear_fwd(b.a_out, w.o->p, b.h_mid, N, H, QH, nullptr, h_in, s));
  TRY(iadr1_rmsnorm_fwd(b.h_mid, w.ln2->p, b.xn2, save ? b.r2 : nullptr, N, H, H, H, c.rms_eps, s));
  static const int swiglu_mode = getenv("IADR1_SWIGLU_EPILOGUE") ? atoi(getenv("IADR1_SWIGLU_EPILOGUE")) : 1;   // 0 never, 1 no-grad passes, 2 always
  if (I % 128 == 0 && (swiglu_mode == 2 || (swiglu_mode == 1 && !save))) {
    // gate_up product with the SwiGLU epilogue: act straight from the accumulator, no activation kernel, no [N, 2I] round trip.
    // Default: the passes that keep nothing for a backward (reference model, rollout prefill) - there it removes a 387 MB write
    // and a 580 MB activation pass per layer (ref_fwd 458 -> 410 ms per step); with the gate | up pre-activations ALSO stored for
    // the backward the epilogue writes 2.25x the bytes and the pass comes out even (measured), so the policy pass keeps the
    // plain GEMM + activation kernel
    iadr1_gemm_t d;
    memset(&d, 0, sizeof(d));
    d.M = (int)N; d.N = I; d.K = H;
    d.batch = d.batch_lo = d.b_lo_div = 1;
    d.A = b.xn2; d.lda = H;
    d.B = w.gu->p; d.ldb = H;
    d.C = b.act; d.ldc = I;
    d.split_k = 1; d.alpha = 1.f; d.epi = 4;
    if (save) { d.gu_out = b.gu; d.gu_ld = 2 * I; }
    TRY(launch_gemm(d, s));
  } else {
    TRY(linear_fwd(b.xn2, w.gu->p, b.gu, N, 2 * I, H, nullptr, nullptr, s));
    TRY(iadr1_act_mul_fwd(b.gu, b.act, N, I, 2 * I, I, I, 0, s));
  }
  TRY(linear_fwd(b.act, w.down->p, h_out, N, H, I, nullptr, b.h_mid, s));
  return 0;
}

}  // namespace iadr1

using namespace iadr1;

extern "C" {

int iadr1_model_create(const iadr1_model_cfg_t* cfg, void** handle) {
  if (!cfg || !handle) return set_error("model_create: null argument");
  if (cfg->hidden <= 0 || cfg->layers <= 0 || cfg->nq <= 0 || cfg->nkv <= 0 || cfg->nq % cfg->nkv || cfg->hd % 8)
    return set_error("model_create: bad text geometry");
  Model* m = new Model();
  m->c = *cfg;
  *handle = m;
  return 0;
}

int iadr1_model_destroy(void* handle) {
  delete static_cast<Model*>(handle);
  return 0;
}

int iadr1_bind_weights(void* handle, const char* name, const void* param_bf16, float* grad_f32) {
  if (!handle || !name || !param_bf16) return set_error("bind_weights: null argument");
  Weight w;
  w.p = param_bf16;
  w.g = grad_f32;
  static_cast<Model*>(handle)->w[name] = w;
  return 0;
}

int iadr1_decoder_workspace_bytes(void* handle, long long n_tokens, long long npad, int mode, long long* bytes) {
  if (!handle || !bytes) return set_error("decoder_workspace_bytes: null argument");
  *bytes = (long long)decoder_layout(static_cast<Model*>(handle)->c, n_tokens, npad, mode, nullptr).bytes;
  return 0;
}

int iadr1_decoder_fwd(void* handle, const int* src_index, const void* image_embeds, long long n_tokens, const float* cos_t,
                      const float* sin_t, const iadr1_attn_plan_t* plan, void* workspace, int mode,
                      const iadr1_kv_sink_t* sink, void** h_last, void* stream) {
  if (!handle || !plan || !workspace) return set_error("decoder_fwd: null argument");
  Model& m = *static_cast<Model*>(handle);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long N = n_tokens;
  if (plan->n_tokens != N) return set_error("decoder_fwd: attention plan covers %lld tokens, called with %lld", plan->n_tokens, N);
  DecoderLayout L = decoder_layout(m.c, N, plan->npad, mode, workspace);
  const Weight* emb = m.find("embed_tokens.weight");
  if (!emb) return set_error("decoder_fwd: embed_tokens.weight not bound");
  const int H = m.c.hidden;
  TRY(iadr1_gather_rows(emb->p, image_embeds, src_index, L.h[0], N, H, H, H, H, s));
  for (int i = 0; i < m.c.layers; ++i) {
    const LayerBuf& b = mode == 1 ? L.lb[i] : L.lb[0];
    TRY(layer_fwd(m, i, L.h[i], L.h[i + 1], b, mode == 1, N, cos_t, sin_t, plan, sink, s));
  }
  if (h_last) *h_last = L.h[m.c.layers];
  return 0;
}

int iadr1_decoder_bwd(void* handle, void* dh, const int* src_index, const float* cos_t, const float* sin_t,
                      const iadr1_attn_plan_t* plan, void* workspace, int mode, long long n_tokens, float* dimg32,
                      iadr1_layer_cb on_layer_done, void* cb_user, void* stream) {
  if (!handle || !plan || !workspace || !dh) return set_error("decoder_bwd: null argument");
  if (mode != 1 && mode != 2) return set_error("decoder_bwd: the forward ran without saving (mode %d)", mode);
  Model& m = *static_cast<Model*>(handle);
  const iadr1_model_cfg_t& c = m.c;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long N = n_tokens;
  const int H = c.hidden, I = c.inter, D = (c.nq + 2 * c.nkv) * c.hd, QH = c.nq * c.hd;
  DecoderLayout L = decoder_layout(c, N, plan->npad, mode, workspace);
  const float scale = 1.f / sqrtf((float)c.hd);
  for (int i = c.layers - 1; i >= 0; --i) {
    const LayerBuf& b = mode == 1 ? L.lb[i] : L.lb[0];
    LayerW w;
    TRY(layer_weights(m, i, w));
    if (mode == 2)      // recompute this layer's activations from its saved input (h[i + 1] is rewritten with the same values)
      TRY(layer_fwd(m, i, L.h[i], L.h[i + 1], b, true, N, cos_t, sin_t, plan, nullptr, s));
    // MLP
    TRY(linear_wgrad(dh, b.act, w.down->g, N, H, I, s));
    if (swiglu_bwd_epilogue_enabled(H) && I % 8 == 0) {
      TRY(linear_dgrad_swiglu_bwd(dh, w.down->p, b.gu, N, H, I, s));
    } else {
      TRY(linear_dgrad(dh, w.down->p, L.dact, N, H, I, s));
      TRY(iadr1_act_mul_bwd(L.dact, b.gu, b.gu, N, I, 2 * I, I, I, 0, s));
    }
    TRY(linear_dgrad(b.gu, w.gu->p, L.dx, N, 2 * I, H, s));
    TRY(linear_wgrad(b.gu, b.xn2, w.gu->g, N, 2 * I, H, s));
    TRY(iadr1_rmsnorm_bwd(L.dx, b.h_mid, w.ln2->p, b.r2, dh, w.ln2->g, N, H, H, 1, s));
    // attention
    TRY(linear_dgrad(dh, w.o->p, L.dattn, N, H, QH, s));
    TRY(linear_wgrad(dh, b.a_out, w.o->g, N, H, QH, s));
    TRY(iadr1_fmha_bwd(b.qkv, L.dattn, b.a_out, b.lse2, N, c.nq, c.nkv, c.hd, plan->ranges, plan->q_items, plan->n_q,
                       plan->sched_dq, plan->n_cta_dq, plan->k_items, plan->n_k, plan->sched_kv, plan->n_cta_kv, L.dqkv, L.delta,
                       L.dkv32, plan->npad, scale, 0, s));
    TRY(iadr1_rope(L.dqkv, cos_t, sin_t, N, c.nq + c.nkv, c.hd, D, 0, 1, s));
    TRY(linear_dgrad(L.dqkv, w.qkv->p, L.dx, N, D, H, s));
    TRY(linear_wgrad(L.dqkv, b.xn, w.qkv->g, N, D, H, s));
    if (w.qkv_b && w.qkv_b->g) TRY(iadr1_colsum(L.dqkv, w.qkv_b->g, N, D, D, s));
    TRY(iadr1_rmsnorm_bwd(L.dx, L.h[i], w.ln1->p, b.r1, dh, w.ln1->g, N, H, H, 1, s));
    if (on_layer_done) on_layer_done(i, cb_user);
  }
  const Weight* emb = m.find("embed_tokens.weight");
  if (!emb) return set_error("decoder_bwd: embed_tokens.weight not bound");
  TRY(iadr1_scatter_add_rows(dh, src_index, emb->g, dimg32, N, H, H, H, H, s));
  return 0;
}

// ---- fused lm_head -> log-softmax -> gather (and its backward); the SFT cross-entropy is the same head ------------------
}  // extern "C"

struct HeadLayout {
  void* hsel; float* rf; void* hn; float* pmax; float* psum; float* tgt; float* lse; float* gscale; void* dlogits; void* dhn; void* dhsel;
  long long tiles; int bn; size_t bytes;
};
static HeadLayout head_layout(const iadr1_model_cfg_t& c, long long M, void* base, bool backward) {
  HeadLayout L;
  L.bn = pick_block_n_public(c.vocab, 0);
  L.tiles = (c.vocab + L.bn - 1) / L.bn;
  Arena a(base);
  L.hsel = a.take(M * c.hidden * 2);
  L.rf = static_cast<float*>(a.take(M * 4));
  L.hn = a.take(M * c.hidden * 2);
  L.pmax = static_cast<float*>(a.take(M * L.tiles * 4));
  L.psum = static_cast<float*>(a.take(M * L.tiles * 4));
  L.tgt = static_cast<float*>(a.take(M * 4));
  L.lse = static_cast<float*>(a.take(M * 4));
  L.gscale = nullptr; L.dlogits = L.dhn = L.dhsel = nullptr;
  if (backward) {
    L.gscale = static_cast<float*>(a.take(M * 4));
    L.dlogits = a.take((size_t)M * c.vocab * 2);
    L.dhn = a.take(M * c.hidden * 2);
    L.dhsel = a.take(M * c.hidden * 2);
  }
  L.bytes = a.off;
  return L;
}

static const Weight* head_weight(const Model& m) {
  const Weight* w = m.find("lm_head.weight");
  return w ? w : m.find("embed_tokens.weight");      // tied embeddings
}

extern "C" {

// the backward's workspace (backward = 1) must be the one the forward ran in: it holds hsel / hn / lse
int iadr1_logprob_workspace_bytes(void* handle, long long m_rows, int backward, long long* bytes) {
  if (!handle || !bytes) return set_error("logprob_workspace_bytes: null argument");
  *bytes = (long long)head_layout(static_cast<Model*>(handle)->c, m_rows, nullptr, backward != 0).bytes;
  return 0;
}

// logp[j] = log_softmax(norm(h[sel[j]]) @ E^T / temperature)[labels[j]]; lse is kept in the workspace for the backward
int iadr1_logprob_fwd(void* handle, const void* h, const int* sel_index, const int* labels, long long m_rows, float temperature,
                      void* workspace, float* logp, void* stream) {
  if (!handle || !workspace) return set_error("logprob_fwd: null argument");
  Model& m = *static_cast<Model*>(handle);
  const iadr1_model_cfg_t& c = m.c;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  HeadLayout L = head_layout(c, m_rows, workspace, false);
  const Weight* E = head_weight(m);
  const Weight* nw = m.find("norm.weight");
  if (!E || !nw) return set_error("logprob_fwd: lm_head / norm.weight not bound");
  const int H = c.hidden;
  TRY(iadr1_gather_rows(h, nullptr, sel_index, L.hsel, m_rows, H, H, 0, H, s));
  TRY(iadr1_rmsnorm_fwd(L.hsel, nw->p, L.hn, L.rf, m_rows, H, H, H, c.rms_eps, s));
  if (cudaMemsetAsync(L.tgt, 0, m_rows * 4, s) != cudaSuccess) return set_error("logprob_fwd: memset failed");
  iadr1_gemm_t d;
  memset(&d, 0, sizeof(d));
  d.M = (int)m_rows; d.N = c.vocab; d.K = H;
  d.batch = d.batch_lo = d.b_lo_div = 1;
  d.A = L.hn; d.lda = H;
  d.B = E->p; d.ldb = H;
  d.split_k = 1; d.alpha = 1.f / temperature;
  d.epi = 1;
  d.labels = labels; d.part_max = L.pmax; d.part_sum = L.psum; d.tgt_logit = L.tgt; d.lse_tiles_n = (int)L.tiles;
  d.block_n = L.bn;
  TRY(launch_gemm(d, s));
  TRY(iadr1_lse_finalize(L.pmax, L.psum, L.tgt, (int)L.tiles, (int)m_rows, L.lse, logp, s));
  return 0;
}

}  // extern "C"

// dh32 [n_tokens][H] fp32 (zeroed by the caller) receives the scatter-ADD of d(hidden) at rows sel_index; `sign` = +1 when
// dsrc is d(loss)/d(logp), -1 when it is d(loss)/d(nll)
static int head_bwd(void* handle, const float* dlogp, float sign, const int* sel_index, const int* labels, long long m_rows,
                    float temperature, void* workspace, float* dh32, void* stream) {
  if (!handle || !workspace) return set_error("logprob_bwd: null argument");
  Model& m = *static_cast<Model*>(handle);
  const iadr1_model_cfg_t& c = m.c;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  HeadLayout L = head_layout(c, m_rows, workspace, true);
  const Weight* E = head_weight(m);
  const Weight* nw = m.find("norm.weight");
  if (!E || !nw) return set_error("logprob_bwd: lm_head / norm.weight not bound");
  const int H = c.hidden, V = c.vocab;
  TRY(scale_f32(dlogp, L.gscale, m_rows, -sign / temperature, s));          // kernel forms (softmax - onehot) * gscale
  iadr1_gemm_t d;
  memset(&d, 0, sizeof(d));
  d.M = (int)m_rows; d.N = V; d.K = H;
  d.batch = d.batch_lo = d.b_lo_div = 1;
  d.A = L.hn; d.lda = H;
  d.B = E->p; d.ldb = H;
  d.C = L.dlogits; d.ldc = V;
  d.split_k = 1; d.alpha = 1.f / temperature;
  d.epi = 2;
  d.labels = labels; d.lse = L.lse; d.gscale = L.gscale;
  TRY(launch_gemm(d, s));
  TRY(linear_dgrad(L.dlogits, E->p, L.dhn, m_rows, V, H, s));
  TRY(linear_wgrad(L.dlogits, L.hn, E->g, m_rows, V, H, s));
  TRY(iadr1_rmsnorm_bwd(L.dhn, L.hsel, nw->p, L.rf, L.dhsel, nw->g, m_rows, H, H, 0, s));
  TRY(iadr1_scatter_add_rows(L.dhsel, sel_index, dh32, nullptr, m_rows, H, H, H, 0, s));
  return 0;
}

extern "C" {

int iadr1_logprob_bwd(void* handle, const float* dlogp, const int* sel_index, const int* labels, long long m_rows,
                      float temperature, void* workspace, float* dh32, void* stream) {
  return head_bwd(handle, dlogp, 1.f, sel_index, labels, m_rows, temperature, workspace, dh32, stream);
}

// SFT head (HF ForCausalLMLoss, ref: train/stage_sft/llamafactory/train/sft/trainer.py:92-107): token cross-entropy is the
// negated log-prob of the label; the two entry points share the kernels with the log-prob head.
int iadr1_ce_fwd(void* handle, const void* h, const int* sel_index, const int* labels, long long m_rows, void* workspace,
                 float* nll, void* stream) {
  TRY(iadr1_logprob_fwd(handle, h, sel_index, labels, m_rows, 1.f, workspace, nll, stream));
  return scale_f32(nll, nll, m_rows, -1.f, static_cast<cudaStream_t>(stream));
}
int iadr1_ce_bwd(void* handle, const float* dnll, const int* sel_index, const int* labels, long long m_rows, void* workspace,
                 float* dh32, void* stream) {
  return head_bwd(handle, dnll, -1.f, sel_index, labels, m_rows, 1.f, workspace, dh32, stream);   // d(nll) = -d(logp)
}

// ---- rollout: prefill = forward-only decoder pass that also fills the shared-prefix KV cache ------------------------------
int iadr1_prefill(void* handle, const int* src_index, const void* image_embeds, long long n_tokens, const float* cos_t,
                  const float* sin_t, const iadr1_attn_plan_t* plan, void* workspace, const iadr1_kv_sink_t* sink, void** h_last,
                  void* stream) {
  if (!sink || !sink->kp || !sink->vp) return set_error("prefill: KV sink required");
  if ((long long)sink->n_groups * sink->p_len != n_tokens) return set_error("prefill: n_groups * p_len must equal n_tokens");
  return iadr1_decoder_fwd(handle, src_index, image_embeds, n_tokens, cos_t, sin_t, plan, workspace, 0, sink, h_last, stream);
}

// One decode step of all rows in flight (the body the rollout captures into a CUDA graph): embed -> L x [rmsnorm, qkv,
// rotary + KV append + attention, o, rmsnorm, gate_up + SwiGLU, down] -> rmsnorm -> lm_head -> sampler -> advance.
static int skinny(const void* W, const void* x, void* out, int F, int K, int R, int split_k, const void* bias, int block_n,
                  cudaStream_t s) {
  iadr1_gemm_t d;
  memset(&d, 0, sizeof(d));
  d.M = F; d.N = R; d.K = K;
  d.batch = d.batch_lo = d.b_lo_div = 1;
  d.A = W; d.lda = K;
  d.B = x; d.ldb = K;
  d.C = out; d.ldc = F; d.c_f32 = 1; d.trans_c = 1;
  d.atomic = split_k > 0 ? 1 : 0;
  d.split_k = split_k > 0 ? split_k : 1;
  d.alpha = 1.f;
  d.bias = bias; d.bias_per_m = 1;
  d.block_n = block_n; d.a_static = 1;
  return launch_gemm(d, s);
}
static int split_for(int m_feat, int k) {
  const int tiles = (m_feat + 127) / 128, nkb = (k + 63) / 64;
  int sp = 148 / (tiles > 0 ? tiles : 1);
  if (sp > nkb) sp = nkb;
  return sp < 1 ? 1 : sp;
}

int iadr1_decode_head(void* handle, const iadr1_decode_t* e, int first, void* stream) {
  Model& m = *static_cast<Model*>(handle);
  const iadr1_model_cfg_t& c = m.c;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const Weight* nw = m.find("norm.weight");
  const Weight* E = head_weight(m);
  if (!nw || !E) return set_error("decode_head: weights not bound");
  TRY(iadr1_rmsnorm_f32in(e->h, nw->p, e->xn, e->R, c.hidden, c.rms_eps, nullptr, 0, s));
  TRY(skinny(E->p, e->xn, e->logits, c.vocab, c.hidden, e->R, 0, nullptr, e->block_n, s));
  return iadr1_sample(e->logits, e->R, c.vocab, e->temperature, e->top_k, e->top_p, 0ull, e->state, e->tok, e->finished,
                      e->out_tokens, e->c_max, e->eos_id, e->pad_id, e->forbid_eos, first, s);
}

// rotary + KV append + attention of one layer for all rows in flight (shared-prefix form when the engine asks for it)
static int decode_attention(const iadr1_model_cfg_t& c, const iadr1_decode_t* e, int i, cudaStream_t s) {
  const long long kvw = (long long)c.nkv * c.hd;
  const int R = e->R;
  const char* kp = static_cast<const char*>(e->kp) + (long long)i * e->n_groups * e->p_max * kvw * 2;
  const char* vp = static_cast<const char*>(e->vp) + (long long)i * e->n_groups * e->p_max * kvw * 2;
  char* kc = static_cast<char*>(e->kc) + (long long)i * R * e->c_max * kvw * 2;
  char* vc = static_cast<char*>(e->vc) + (long long)i * R * e->c_max * kvw * 2;
  const float scale = 1.f / sqrtf((float)c.hd);
  if (e->attn_psplit > 0 && e->nsplit > e->attn_psplit && (c.hd == 64 || c.hd == 128) && e->n_groups > 0 && R % e->n_groups == 0)
    return iadr1_decode_attention_grouped(e->qkv, e->cos_tab, e->sin_tab, e->rope_delta, kp, vp, kc, vc, e->state, e->row_plen,
                                          e->finished, e->part, e->tickets, e->attn, R, R / e->n_groups, c.nq, c.nkv, c.hd, e->p_max, e->c_max,
                                          e->attn_psplit, e->nsplit - e->attn_psplit, e->max_pos, scale, s);
  return iadr1_decode_attention_fused(e->qkv, e->cos_tab, e->sin_tab, e->rope_delta, kp, vp, kc, vc, e->state, e->row_group,
                                      e->row_plen, e->finished, e->part, e->tickets, e->attn, R, c.nq, c.nkv, c.hd, e->p_max, e->c_max,
                                      e->nsplit, e->max_pos, scale, s);
}

// lm_head + sampler on the already normalised rows in e->xn
static int decode_head_normed(Model& m, const iadr1_decode_t* e, int first, cudaStream_t s) {
  const iadr1_model_cfg_t& c = m.c;
  const Weight* E = head_weight(m);
  if (!E) return set_error("decode_head: weights not bound");
  TRY(skinny(E->p, e->xn, e->logits, c.vocab, c.hidden, e->R, 0, nullptr, e->block_n, s));
  return iadr1_sample(e->logits, e->R, c.vocab, e->temperature, e->top_k, e->top_p, 0ull, e->state, e->tok, e->finished,
                      e->out_tokens, e->c_max, e->eos_id, e->pad_id, e->forbid_eos, first, s);
}

// The decode step as 2 launches per layer: attention + the persistent chain kernel (csrc/decode_chain.cu), which runs
// o -> norm -> gate_up + SwiGLU -> down -> next layer's norm -> next layer's qkv with the weights streaming across the phases.
static int decode_step_chained(Model& m, const iadr1_decode_t* e, cudaStream_t s) {
  const iadr1_model_cfg_t& c = m.c;
  const int H = c.hidden, I = c.inter, D = (c.nq + 2 * c.nkv) * c.hd, QH = c.nq * c.hd, R = e->R;
  const Weight* emb = m.find("embed_tokens.weight");
  const Weight* nw = m.find("norm.weight");
  if (!emb || !nw) return set_error("decode_step: embed_tokens.weight / norm.weight not bound");
  if (cudaMemsetAsync(e->chain_counters, 0, (size_t)(c.layers + 1) * 8 * sizeof(unsigned), s) != cudaSuccess)
    return set_error("decode_step: memset failed");
  TRY(iadr1_decode_embed(emb->p, e->tok, e->h, R, H, s));
  const long long kvw = (long long)c.nkv * c.hd;
  LayerW w, wn;
  TRY(layer_weights(m, 0, w));
  TRY(launch_decode_chain(R, H, I, QH, D, c.rms_eps, e->h, e->xn, e->act, e->qkv, nullptr, nullptr, nullptr, nullptr, nullptr,
                          w.ln1->p, w.qkv->p, w.qkv_b ? w.qkv_b->p : nullptr, 0, 1, 1, e->chain_counters, s));
  for (int i = 0; i < c.layers; ++i) {
    const bool last = i + 1 == c.layers;
    if (!last) TRY(layer_weights(m, i + 1, wn));
    TRY(decode_attention(c, e, i, s));
    TRY(launch_decode_chain(R, H, I, QH, D, c.rms_eps, e->h, e->xn, e->act, e->qkv, e->attn, w.o->p, w.ln2->p, w.gu->p, w.down->p,
                            last ? nw->p : wn.ln1->p, last ? nullptr : wn.qkv->p, (last || !wn.qkv_b) ? nullptr : wn.qkv_b->p, 1, 1, last ? 0 : 1,
                            e->chain_counters + (i + 1) * 8, s));
    if (!last) w = wn;
  }
  TRY(decode_head_normed(m, e, 0, s));
  return iadr1_decode_advance(e->state, s);
}

int iadr1_decode_step(void* handle, const iadr1_decode_t* e, void* stream) {
  if (!handle || !e) return set_error("decode_step: null argument");
  Model& m = *static_cast<Model*>(handle);
  const iadr1_model_cfg_t& c = m.c;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int H = c.hidden, I = c.inter, D = (c.nq + 2 * c.nkv) * c.hd, QH = c.nq * c.hd, R = e->R;
  if (e->chain_counters && R <= 128 && H <= 4096 && decode_chain_enabled()) return decode_step_chained(m, e, s);
  const Weight* emb = m.find("embed_tokens.weight");
  if (!emb) return set_error("decode_step: embed_tokens.weight not bound");
  TRY(iadr1_decode_embed(emb->p, e->tok, e->h, R, H, s));
  const int sk_qkv = split_for(D, H), sk_o = split_for(H, QH), sk_d = split_for(H, I);
  const long long kvw = (long long)c.nkv * c.hd;
  for (int i = 0; i < c.layers; ++i) {
    LayerW w;
    TRY(layer_weights(m, i, w));
    TRY(iadr1_rmsnorm_f32in(e->h, w.ln1->p, e->xn, R, H, c.rms_eps, e->qkv, D, s));
    TRY(skinny(w.qkv->p, e->xn, e->qkv, D, H, R, sk_qkv, w.qkv_b ? w.qkv_b->p : nullptr, e->block_n, s));
    TRY(decode_attention(c, e, i, s));
    TRY(skinny(w.o->p, e->attn, e->h, H, QH, R, sk_o, nullptr, e->block_n, s));
    TRY(iadr1_rmsnorm_f32in(e->h, w.ln2->p, e->xn, R, H, c.rms_eps, nullptr, 0, s));
    {
      iadr1_gemm_t d;
      memset(&d, 0, sizeof(d));
      d.M = 2 * I; d.N = R; d.K = H;
      d.batch = d.batch_lo = d.b_lo_div = 1;
      d.A = w.gu->p; d.lda = H;
      d.B = e->xn; d.ldb = H;
      d.C = e->act; d.ldc = I;
      d.split_k = 1; d.alpha = 1.f; d.epi = 3; d.up_row_off = I;
      d.block_n = e->block_n; d.a_static = 1; d.co_resident = 1;
      TRY(launch_gemm(d, s));
    }
    TRY(skinny(w.down->p, e->act, e->h, H, I, R, sk_d, nullptr, e->block_n, s));
  }
  TRY(iadr1_decode_head(handle, e, 0, stream));
  return iadr1_decode_advance(e->state, s);
}

}  // extern "C"

// =====================================================================================================================
// vision towers (Qwen2.5-VL: RMSNorm + SwiGLU + windowed / full attention; Qwen2-VL: LayerNorm + quick-GELU MLP, full
// attention; SigLIP under LLaVA-OneVision: learned positions, LayerNorm, tanh-GELU MLP, projector + anyres packing):
// patch embedding -> blocks -> merger / projector as one call each way. HF: modeling_qwen2_5_vl.py:455-518,
// modeling_qwen2_vl.py (VisionTransformer), modeling_siglip.py:116-186 + modeling_llava_onevision.py:137-156, 292-355.
// =====================================================================================================================
namespace iadr1 {

struct VBlockBuf {
  float* st1; void* xn; void* qkv; float* lse2; void* attn; void* x_mid; float* st2; void* xn2; void* gu; void* act;
};
struct VisionLayout {
  std::vector<void*> x;            // x[i] = input of block i, x[depth] = tower output
  std::vector<VBlockBuf> bb;       // per block (save) or one shared
  void* pos; void* x_pre; float* st0; float* stq; void* xq; void* m1; void* a1; void* feat;       // embedding residual, merger / projector
  void* dx; void* dact; void* dxn; void* dattn; void* dqkv; float* delta; float* dkv32; void* dm1; void* dfeat; float* dfeat32;
  size_t bytes;
};

static VisionLayout vision_layout(const iadr1_model_cfg_t& c, long long N, long long n_out, long long npad, int save, void* base) {
  Arena a(base);
  VisionLayout L;
  const long long E = c.v_hidden, Ip = c.v_inter, W1 = c.v_kind == 0 ? 2 * Ip : Ip, Ho = c.v_out_hidden, unit = c.v_merge_unit;
  L.x.resize(c.v_depth + 1);
  if (save) {
    for (int i = 0; i <= c.v_depth; ++i) L.x[i] = a.take(N * E * 2);
  } else {
    void* p0 = a.take(N * E * 2);
    void* p1 = a.take(N * E * 2);
    for (int i = 0; i <= c.v_depth; ++i) L.x[i] = (i & 1) ? p1 : p0;
  }
  const int nb = save ? c.v_depth : 1;
  for (int i = 0; i < nb; ++i) {
    VBlockBuf b;
    b.st1 = static_cast<float*>(a.take(2 * N * 4));
    b.xn = a.take(N * E * 2);
    b.qkv = a.take(N * 3 * E * 2);
    b.lse2 = static_cast<float*>(a.take((size_t)c.v_heads * npad * 4));
    b.attn = a.take(N * E * 2);
    b.x_mid = a.take(N * E * 2);
    b.st2 = static_cast<float*>(a.take(2 * N * 4));
    b.xn2 = a.take(N * E * 2);
    b.gu = a.take(N * W1 * 2);
    b.act = a.take(N * Ip * 2);
    L.bb.push_back(b);
  }
  L.pos = c.v_kind >= 2 ? a.take(N * E * 2) : nullptr;
  L.x_pre = c.v_kind == 3 ? a.take(N * E * 2) : nullptr;            // CLIP: embeddings before pre_layrnorm
  L.st0 = c.v_kind == 3 ? static_cast<float*>(a.take(2 * N * 4)) : nullptr;
  L.stq = static_cast<float*>(a.take(2 * N * 4));
  L.xq = a.take(N * E * 2);
  const long long Mm = c.v_kind >= 2 ? N : N / unit, Km = c.v_kind >= 2 ? Ho : unit * E;      // merger rows / fc1 width
  L.m1 = a.take(Mm * Km * 2);
  L.a1 = a.take(Mm * Km * 2);
  L.feat = a.take(Mm * Ho * 2);
  L.dx = L.dact = L.dxn = L.dattn = L.dqkv = L.dm1 = L.dfeat = nullptr;
  L.delta = L.dkv32 = L.dfeat32 = nullptr;
  if (save) {
    L.dx = a.take(N * E * 2);
    L.dact = a.take(N * Ip * 2);
    L.dxn = a.take(N * E * 2);
    L.dattn = a.take(N * E * 2);
    L.dqkv = a.take(N * 3 * E * 2);
    L.delta = static_cast<float*>(a.take((size_t)c.v_heads * npad * 4));
    L.dkv32 = static_cast<float*>(a.take((size_t)N * 2 * E * 4));
    L.dm1 = a.take(Mm * Km * 2);
    L.dfeat = a.take(Mm * Ho * 2);
    L.dfeat32 = c.v_kind >= 2 ? static_cast<float*>(a.take((size_t)N * Ho * 4)) : nullptr;
  }
  (void)n_out;
  L.bytes = a.off;
  return L;
}

struct VBlockW {
  const Weight *n1, *n1b, *qkv, *qkvb, *proj, *projb, *n2, *n2b, *w1, *w1b, *w2, *w2b;
};
static int vblock_weights(const Model& m, int i, VBlockW& w) {
  const std::string b = "visual.blocks." + std::to_string(i) + ".";
  const bool q25 = m.c.v_kind == 0;
  w.n1 = m.find(b + "norm1.weight"); w.n1b = m.find(b + "norm1.bias");
  w.qkv = m.find(b + "qkv.weight"); w.qkvb = m.find(b + "qkv.bias");
  w.proj = m.find(b + "proj.weight"); w.projb = m.find(b + "proj.bias");
  w.n2 = m.find(b + "norm2.weight"); w.n2b = m.find(b + "norm2.bias");
  w.w1 = m.find(b + (q25 ? "gate_up.weight" : "fc1.weight")); w.w1b = m.find(b + (q25 ? "gate_up.bias" : "fc1.bias"));
  w.w2 = m.find(b + (q25 ? "down.weight" : "fc2.weight")); w.w2b = m.find(b + (q25 ? "down.bias" : "fc2.bias"));
  if (!w.n1 || !w.qkv || !w.qkvb || !w.proj || !w.projb || !w.n2 || !w.w1 || !w.w1b || !w.w2 || !w.w2b || (!q25 && (!w.n1b || !w.n2b)))
    return set_error("vision block %d: weights not bound", i);
  return 0;
}

static int vnorm_fwd(const iadr1_model_cfg_t& c, const void* x, const Weight* w, const Weight* b, void* y, float* st, long long N,
                     cudaStream_t s) {
  const int E = c.v_hidden;
  if (c.v_kind == 0) return iadr1_rmsnorm_fwd(x, w->p, y, st, N, E, E, E, c.v_eps, s);
  return iadr1_layernorm_fwd(x, w->p, b->p, y, st, st + N, N, E, E, c.v_eps, s);
}
static int vnorm_bwd(const iadr1_model_cfg_t& c, const void* dy, const void* x, const Weight* w, const Weight* b, const float* st,
                     void* dx, long long N, int add, cudaStream_t s) {
  const int E = c.v_hidden;
  if (c.v_kind == 0) return iadr1_rmsnorm_bwd(dy, x, w->p, st, dx, w->g, N, E, E, add, s);
  return iadr1_layernorm_bwd(dy, x, w->p, st, st + N, dx, w->g, b->g, N, E, E, add, s);
}
// y = x @ W^T + b (+ residual) with the bias gradient as a column sum in the backward
static int vlinear_bwd(const void* dy, const void* x, const Weight* W, const Weight* b, void* dx, long long M, int N, int K,
                       cudaStream_t s) {
  if (dx) TRY(linear_dgrad(dy, W->p, dx, M, N, K, s));
  TRY(linear_wgrad(dy, x, W->g, M, N, K, s));
  if (b && b->g) TRY(iadr1_colsum(dy, b->g, M, N, N, s));
  return 0;
}

}  // namespace iadr1

extern "C" {

int iadr1_vision_workspace_bytes(void* handle, long long n_patches, long long n_out, long long npad, int save, long long* bytes) {
  if (!handle || !bytes) return set_error("vision_workspace_bytes: null argument");
  const iadr1_model_cfg_t& c = static_cast<Model*>(handle)->c;
  if (c.v_kind < 0) return set_error("vision: the model was created without a vision tower");
  *bytes = (long long)vision_layout(c, n_patches, n_out, npad, save, nullptr).bytes;
  return 0;
}

int iadr1_vision_fwd(void* handle, const void* pixel_values, const iadr1_vision_geom_t* geo, void* workspace, int save, void* out,
                     void* stream) {
  if (!handle || !geo || !workspace || !out) return set_error("vision_fwd: null argument");
  Model& m = *static_cast<Model*>(handle);
  const iadr1_model_cfg_t& c = m.c;
  if (c.v_kind < 0) return set_error("vision: the model was created without a vision tower");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long N = geo->n_patches;
  const int E = c.v_hidden, nh = c.v_heads, hd = E / nh, Ip = c.v_inter, W1 = c.v_kind == 0 ? 2 * Ip : Ip, Ho = c.v_out_hidden;
  const int unit = c.v_merge_unit, Kp = c.v_patch_dim;
  const iadr1_attn_plan_t* pf = geo->plan_full;
  VisionLayout L = vision_layout(c, N, geo->n_out, pf->npad, save, workspace);
  const Weight* pe = m.find("visual.patch_embed.weight");
  if (!pe) return set_error("vision_fwd: visual.patch_embed.weight not bound");
  const float scale = 1.f / sqrtf((float)hd);
  // ---- patch embedding (Conv as one GEMM over patch rows)
  if (c.v_kind == 2) {
    const Weight *pb = m.find("visual.patch_embed.bias"), *pt = m.find("visual.pos_embed.weight");
    if (!pb || !pt) return set_error("vision_fwd: SigLIP embedding weights not bound");
    TRY(iadr1_gather_rows(pt->p, nullptr, geo->pos_index, L.pos, N, E, E, 0, E, s));       // position table tiled over the crops
    TRY(linear_fwd(pixel_values, pe->p, L.x[0], N, E, Kp, pb->p, L.pos, s));
  } else if (c.v_kind == 3) {
    // CLIP (HF modeling_clip.py CLIPVisionEmbeddings + pre_layrnorm): no conv bias; the class token is the row whose pixel
    // vector is the unit vector of the class-embedding column of the fused patch weight; LayerNorm before the first block
    const Weight *pt = m.find("visual.pos_embed.weight"), *pl = m.find("visual.pre_ln.weight"), *plb = m.find("visual.pre_ln.bias");
    if (!pt || !pl || !plb) return set_error("vision_fwd: CLIP embedding weights not bound");
    TRY(iadr1_gather_rows(pt->p, nullptr, geo->pos_index, L.pos, N, E, E, 0, E, s));
    TRY(linear_fwd(pixel_values, pe->p, L.x_pre, N, E, Kp, nullptr, L.pos, s));
    TRY(iadr1_layernorm_fwd(L.x_pre, pl->p, plb->p, L.x[0], L.st0, L.st0 + N, N, E, E, c.v_eps, s));
  } else if (geo->window_index) {
    // block 0's normalised-input buffer is free until its first norm: scratch for the un-permuted patch embedding (x[depth]
    // would alias x[0] in the two-buffer layout of a forward-only call with an even depth)
    void* pe_out = L.bb[0].xn;
    TRY(linear_fwd(pixel_values, pe->p, pe_out, N, E, Kp, nullptr, nullptr, s));
    TRY(iadr1_gather_rows(pe_out, nullptr, geo->window_index, L.x[0], N / unit, unit * E, unit * E, 0, unit * E, s));
  } else {
    TRY(linear_fwd(pixel_values, pe->p, L.x[0], N, E, Kp, nullptr, nullptr, s));
  }
  // ---- blocks
  for (int i = 0; i < c.v_depth; ++i) {
    const VBlockBuf& b = save ? L.bb[i] : L.bb[0];
    VBlockW w;
    TRY(vblock_weights(m, i, w));
    const bool full = c.v_kind != 0 || ((c.v_fullatt_mask >> i) & 1u);
    const iadr1_attn_plan_t* plan = full ? pf : geo->plan_win;
    if (!plan) return set_error("vision_fwd: missing attention plan for block %d", i);
    TRY(vnorm_fwd(c, L.x[i], w.n1, w.n1b, b.xn, b.st1, N, s));
    TRY(linear_fwd(b.xn, w.qkv->p, b.qkv, N, 3 * E, E, w.qkvb->p, nullptr, s));
    if (geo->cos) TRY(iadr1_rope(b.qkv, geo->cos, geo->sin, N, 2 * nh, hd, 3 * E, 0, 0, s));
    TRY(iadr1_fmha_fwd(b.qkv, N, nh, nh, hd, plan->ranges, plan->q_items, plan->n_q, plan->sched_fwd, plan->n_cta_fwd, b.attn,
                       b.lse2, plan->npad, scale, 0, s));
    TRY(linear_fwd(b.attn, w.proj->p, b.x_mid, N, E, E, w.projb->p, L.x[i], s));
    TRY(vnorm_fwd(c, b.x_mid, w.n2, w.n2b, b.xn2, b.st2, N, s));
    TRY(linear_fwd(b.xn2, w.w1->p, b.gu, N, W1, E, w.w1b->p, nullptr, s));
    if (c.v_kind == 0) TRY(iadr1_act_mul_fwd(b.gu, b.act, N, Ip, 2 * Ip, Ip, Ip, 0, s));
    else TRY(iadr1_act_mul_fwd(b.gu, b.act, N, Ip, Ip, -1, Ip, c.v_kind == 2 ? 3 : 2, s));      // quick-GELU (Qwen2-VL, CLIP) / tanh-GELU
    TRY(linear_fwd(b.act, w.w2->p, L.x[i + 1], N, E, Ip, w.w2b->p, b.x_mid, s));
  }
  const void* xl = L.x[c.v_depth];
  const Weight *f1 = m.find("visual.merger.fc1.weight"), *f1b = m.find("visual.merger.fc1.bias");
  const Weight *f2 = m.find("visual.merger.fc2.weight"), *f2b = m.find("visual.merger.fc2.bias");
  if (!f1 || !f1b || !f2 || !f2b) return set_error("vision_fwd: merger / projector weights not bound");
  if (c.v_kind >= 2) {
    // Llava(Onevision)MultiModalProjector on the selected encoder hidden state, then the packing gather: anyres re-tiling
    // with image_newline (OneVision) or simply the patch tokens without the class token (LLaVA-1.5)
    const Weight* nl = m.find("image_newline");
    if (c.v_kind == 2 && !nl) return set_error("vision_fwd: image_newline not bound");
    TRY(linear_fwd(xl, f1->p, L.m1, N, Ho, E, f1b->p, nullptr, s));
    TRY(iadr1_act_mul_fwd(L.m1, L.a1, N, Ho, Ho, -1, Ho, 1, s));
    TRY(linear_fwd(L.a1, f2->p, L.feat, N, Ho, Ho, f2b->p, nullptr, s));
    TRY(iadr1_gather_rows(L.feat, nl ? nl->p : nullptr, geo->pack_index, out, geo->n_out, Ho, Ho, nl ? Ho : 0, Ho, s));
  } else {
    const Weight *lq = m.find("visual.merger.ln_q.weight"), *lqb = m.find("visual.merger.ln_q.bias");
    if (!lq || (c.v_kind == 1 && !lqb)) return set_error("vision_fwd: merger ln_q not bound");
    TRY(vnorm_fwd(c, xl, lq, lqb, L.xq, L.stq, N, s));
    const long long Mm = N / unit;
    const int Km = unit * E;
    TRY(linear_fwd(L.xq, f1->p, L.m1, Mm, Km, Km, f1b->p, nullptr, s));
    TRY(iadr1_act_mul_fwd(L.m1, L.a1, Mm, Km, Km, -1, Km, 1, s));
    if (geo->reverse_index) {
      TRY(linear_fwd(L.a1, f2->p, L.feat, Mm, Ho, Km, f2b->p, nullptr, s));
      TRY(iadr1_gather_rows(L.feat, nullptr, geo->reverse_index, out, Mm, Ho, Ho, 0, Ho, s));
    } else {
      TRY(linear_fwd(L.a1, f2->p, out, Mm, Ho, Km, f2b->p, nullptr, s));
    }
  }
  return 0;
}

// d_out bf16 [n_out][out_hidden]; pixel_values as in the forward (the patch-embedding weight gradient needs them)
int iadr1_vision_bwd(void* handle, const void* d_out, const void* pixel_values, const iadr1_vision_geom_t* geo, void* workspace,
                     void* stream) {
  if (!handle || !geo || !workspace || !d_out) return set_error("vision_bwd: null argument");
  Model& m = *static_cast<Model*>(handle);
  const iadr1_model_cfg_t& c = m.c;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long N = geo->n_patches;
  const int E = c.v_hidden, nh = c.v_heads, hd = E / nh, Ip = c.v_inter, W1 = c.v_kind == 0 ? 2 * Ip : Ip, Ho = c.v_out_hidden;
  const int unit = c.v_merge_unit, Kp = c.v_patch_dim;
  const iadr1_attn_plan_t* pf = geo->plan_full;
  VisionLayout L = vision_layout(c, N, geo->n_out, pf->npad, 1, workspace);
  const float scale = 1.f / sqrtf((float)hd);
  const Weight *f1 = m.find("visual.merger.fc1.weight"), *f1b = m.find("visual.merger.fc1.bias");
  const Weight *f2 = m.find("visual.merger.fc2.weight"), *f2b = m.find("visual.merger.fc2.bias");
  const void* xl = L.x[c.v_depth];
  if (c.v_kind >= 2) {
    const Weight* nl = m.find("image_newline");
    // un-pack: dropped features (and CLIP's class token) get zero gradient, image_newline collects one row per feature-map row
    if (cudaMemsetAsync(L.dfeat32, 0, (size_t)N * Ho * 4, s) != cudaSuccess) return set_error("vision_bwd: memset failed");
    TRY(iadr1_scatter_add_rows(d_out, geo->pack_index, L.dfeat32, nl ? nl->g : nullptr, geo->n_out, Ho, Ho, Ho, nl ? Ho : 0, s));
    TRY(iadr1_cast_f32_bf16(L.dfeat32, L.dfeat, N * Ho, s));
    TRY(vlinear_bwd(L.dfeat, L.a1, f2, f2b, L.dm1, N, Ho, Ho, s));
    TRY(iadr1_act_mul_bwd(L.dm1, L.m1, L.m1, N, Ho, Ho, -1, Ho, 1, s));
    TRY(vlinear_bwd(L.m1, xl, f1, f1b, L.dx, N, Ho, E, s));
  } else {
    const Weight *lq = m.find("visual.merger.ln_q.weight"), *lqb = m.find("visual.merger.ln_q.bias");
    const long long Mm = N / unit;
    const int Km = unit * E;
    const void* dfe = d_out;
    if (geo->window_index) {      // inverse of the final un-permute
      TRY(iadr1_gather_rows(d_out, nullptr, geo->window_index, L.dfeat, Mm, Ho, Ho, 0, Ho, s));
      dfe = L.dfeat;
    }
    TRY(vlinear_bwd(dfe, L.a1, f2, f2b, L.dm1, Mm, Ho, Km, s));
    TRY(iadr1_act_mul_bwd(L.dm1, L.m1, L.m1, Mm, Km, Km, -1, Km, 1, s));
    TRY(vlinear_bwd(L.m1, L.xq, f1, f1b, L.dxn, Mm, Km, Km, s));      // dxn viewed [Mm, unit * E] = [N, E]
    TRY(vnorm_bwd(c, L.dxn, xl, lq, lqb, L.stq, L.dx, N, 0, s));
  }
  for (int i = c.v_depth - 1; i >= 0; --i) {
    const VBlockBuf& b = L.bb[i];
    VBlockW w;
    TRY(vblock_weights(m, i, w));
    const bool full = c.v_kind != 0 || ((c.v_fullatt_mask >> i) & 1u);
    const iadr1_attn_plan_t* plan = full ? pf : geo->plan_win;
    TRY(vlinear_bwd(L.dx, b.act, w.w2, w.w2b, L.dact, N, E, Ip, s));
    if (c.v_kind == 0) TRY(iadr1_act_mul_bwd(L.dact, b.gu, b.gu, N, Ip, 2 * Ip, Ip, Ip, 0, s));
    else TRY(iadr1_act_mul_bwd(L.dact, b.gu, b.gu, N, Ip, Ip, -1, Ip, c.v_kind == 2 ? 3 : 2, s));
    TRY(vlinear_bwd(b.gu, b.xn2, w.w1, w.w1b, L.dxn, N, W1, E, s));
    TRY(vnorm_bwd(c, L.dxn, b.x_mid, w.n2, w.n2b, b.st2, L.dx, N, 1, s));
    TRY(vlinear_bwd(L.dx, b.attn, w.proj, w.projb, L.dattn, N, E, E, s));
    TRY(iadr1_fmha_bwd(b.qkv, L.dattn, b.attn, b.lse2, N, nh, nh, hd, plan->ranges, plan->q_items, plan->n_q, plan->sched_dq,
                       plan->n_cta_dq, plan->k_items, plan->n_k, plan->sched_kv, plan->n_cta_kv, L.dqkv, L.delta, L.dkv32,
                       plan->npad, scale, 0, s));
    if (geo->cos) TRY(iadr1_rope(L.dqkv, geo->cos, geo->sin, N, 2 * nh, hd, 3 * E, 0, 1, s));
    TRY(vlinear_bwd(L.dqkv, b.xn, w.qkv, w.qkvb, L.dxn, N, 3 * E, E, s));
    TRY(vnorm_bwd(c, L.dxn, L.x[i], w.n1, w.n1b, b.st1, L.dx, N, 1, s));
  }
  const Weight* pe = m.find("visual.patch_embed.weight");
  if (c.v_kind == 3) {
    // through pre_layrnorm to the embeddings; the class embedding's gradient is a column of the patch-weight gradient
    const Weight *pt = m.find("visual.pos_embed.weight"), *pl = m.find("visual.pre_ln.weight"), *plb = m.find("visual.pre_ln.bias");
    TRY(iadr1_layernorm_bwd(L.dx, L.x_pre, pl->p, L.st0, L.st0 + N, L.dxn, pl->g, plb->g, N, E, E, 0, s));
    if (pt->g) TRY(iadr1_colsum(L.dxn, pt->g, N / c.v_tokens_per_crop, c.v_tokens_per_crop * E, c.v_tokens_per_crop * E, s));
    TRY(vlinear_bwd(L.dxn, pixel_values, pe, nullptr, nullptr, N, E, Kp, s));
  } else if (c.v_kind == 2) {
    const Weight *pb = m.find("visual.patch_embed.bias"), *pt = m.find("visual.pos_embed.weight");
    // x0 = patch_embed(px) + bias + pos[tile]: the table's gradient is the sum over crops
    if (pt->g) TRY(iadr1_colsum(L.dx, pt->g, N / c.v_tokens_per_crop, c.v_tokens_per_crop * E, c.v_tokens_per_crop * E, s));
    TRY(vlinear_bwd(L.dx, pixel_values, pe, pb, nullptr, N, E, Kp, s));
  } else {
    const void* dxe = L.dx;
    if (geo->reverse_index) {
      TRY(iadr1_gather_rows(L.dx, nullptr, geo->reverse_index, L.dxn, N / unit, unit * E, unit * E, 0, unit * E, s));
      dxe = L.dxn;
    }
    TRY(linear_wgrad(dxe, pixel_values, pe->g, N, E, Kp, s));
  }
  return 0;
}

}  // extern "C"
