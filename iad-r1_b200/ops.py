"""Tensor-level wrappers over the C ABI (include/iadr1_b200.h) plus the three compositions the model code uses:
linear (fwd / dgrad / wgrad), attention (QK^T -> masked softmax -> PV on the tcgen05 GEMM) and the fused
lm_head -> log-softmax -> gather. PyTorch only supplies device buffers and the current stream.

Every function here launches library kernels; none computes with torch ops (DESIGN.md "no fallback").
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import fmha as F
from . import lib as L

bf16 = torch.bfloat16
f32 = torch.float32

ACT_SILU, ACT_GELU, ACT_QUICK_GELU, ACT_GELU_TANH = 0, 1, 2, 3
EPI_STORE, EPI_LSE, EPI_DLOGITS = 0, 1, 2


def _s():
    return L.stream_ptr()


def _p(t):
    return None if t is None else t.data_ptr()


def ceil_to(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# ---- norms -------------------------------------------------------------------------------------------------------
def rmsnorm_fwd(x, w, eps, out=None, save_rstd=True):
    rows, cols = x.shape
    y = torch.empty_like(x) if out is None else out
    rstd = torch.empty(rows, dtype=f32, device=x.device) if save_rstd else None
    L.check(L.lib().iadr1_rmsnorm_fwd(_p(x), _p(w), _p(y), _p(rstd), rows, cols, x.stride(0), y.stride(0), eps, _s()),
            "rmsnorm_fwd")
    return y, rstd


def rmsnorm_bwd(dy, x, w, rstd, dx, dw32, add_dx):
    rows, cols = x.shape
    assert dy.stride(0) == x.stride(0) == dx.stride(0)
    L.check(L.lib().iadr1_rmsnorm_bwd(_p(dy), _p(x), _p(w), _p(rstd), _p(dx), _p(dw32), rows, cols, x.stride(0),
                                      int(add_dx), _s()), "rmsnorm_bwd")
    return dx


def layernorm_fwd(x, w, b, eps, out=None):
    rows, cols = x.shape
    y = torch.empty_like(x) if out is None else out
    mean = torch.empty(rows, dtype=f32, device=x.device)
    rstd = torch.empty(rows, dtype=f32, device=x.device)
    L.check(L.lib().iadr1_layernorm_fwd(_p(x), _p(w), _p(b), _p(y), _p(mean), _p(rstd), rows, cols, x.stride(0), eps,
                                        _s()), "layernorm_fwd")
    return y, mean, rstd


def layernorm_bwd(dy, x, w, mean, rstd, dx, dw32, db32, add_dx):
    rows, cols = x.shape
    L.check(L.lib().iadr1_layernorm_bwd(_p(dy), _p(x), _p(w), _p(mean), _p(rstd), _p(dx), _p(dw32), _p(db32), rows,
                                        cols, x.stride(0), int(add_dx), _s()), "layernorm_bwd")
    return dx


# ---- rotary / activations ------------------------------------------------------------------------------------------
def rope_(x, cos, sin, heads, hd, bf16_ops, backward=False):
    """In place on the first `heads` heads of x[tokens, heads_total*hd]; cos/sin fp32 [tokens, hd]."""
    tokens = x.shape[0]
    assert cos.dtype == f32 and cos.shape == (tokens, hd) and cos.is_contiguous() and sin.is_contiguous()
    L.check(L.lib().iadr1_rope(_p(x), _p(cos), _p(sin), tokens, heads, hd, x.stride(0), int(bf16_ops), int(backward),
                               _s()), "rope")
    return x


def act_mul_fwd(gu, cols, act, gated=True, out=None):
    rows = gu.shape[0]
    if out is None:
        out = torch.empty(rows, cols, dtype=bf16, device=gu.device)
    L.check(L.lib().iadr1_act_mul_fwd(_p(gu), _p(out), rows, cols, gu.stride(0), cols if gated else -1, out.stride(0),
                                      act, _s()), "act_mul_fwd")
    return out


def act_mul_bwd(dout, gu, cols, act, gated=True, dgu=None):
    rows = gu.shape[0]
    if dgu is None:
        dgu = torch.empty_like(gu)
    assert dgu.stride(0) == gu.stride(0)
    L.check(L.lib().iadr1_act_mul_bwd(_p(dout), _p(gu), _p(dgu), rows, cols, gu.stride(0), cols if gated else -1,
                                      dout.stride(0), act, _s()), "act_mul_bwd")
    return dgu


# ---- softmax / gathers / reductions -----------------------------------------------------------------------------------
def softmax_rows_(S, lo, hi, Tq, Tk, ld, z_stride, batch, hole=(0, 0)):
    L.check(L.lib().iadr1_softmax_rows(_p(S), _p(lo), _p(hi), Tq, Tk, ld, z_stride, batch, hole[0], hole[1], _s()),
            "softmax_rows")


def softmax_bwd_rows_(P, dP, lo, hi, Tq, Tk, ld, z_stride, batch, hole=(0, 0)):
    L.check(L.lib().iadr1_softmax_bwd_rows(_p(P), _p(dP), _p(lo), _p(hi), Tq, Tk, ld, z_stride, batch, hole[0], hole[1],
                                           _s()), "softmax_bwd_rows")


def gather_rows(table, index, alt=None, out=None):
    rows = index.shape[0]
    cols = table.shape[1]
    if out is None:
        out = torch.empty(rows, cols, dtype=bf16, device=table.device)
    assert index.dtype == torch.int32
    L.check(L.lib().iadr1_gather_rows(_p(table), _p(alt), _p(index), _p(out), rows, cols, table.stride(0),
                                      alt.stride(0) if alt is not None else 0, out.stride(0), _s()), "gather_rows")
    return out


def scatter_add_rows(d, index, dtable32, dalt32):
    rows, cols = d.shape
    L.check(L.lib().iadr1_scatter_add_rows(_p(d), _p(index), _p(dtable32), _p(dalt32), rows, cols, d.stride(0),
                                           dtable32.stride(0) if dtable32 is not None else 0,
                                           dalt32.stride(0) if dalt32 is not None else 0, _s()), "scatter_add_rows")


def colsum(x, out32):
    rows, cols = x.shape
    L.check(L.lib().iadr1_colsum(_p(x), _p(out32), rows, cols, x.stride(0), _s()), "colsum")


def group_sum(src, out, rows, nkv, g, hd, src_ld, out_ld, accumulate=False):
    L.check(L.lib().iadr1_group_sum(_p(src), _p(out), rows, nkv, g, hd, src_ld, out_ld, int(accumulate), _s()), "group_sum")


def add_bf16(a, b, out=None):
    out = torch.empty_like(a) if out is None else out
    L.check(L.lib().iadr1_add_bf16(_p(a), _p(b), _p(out), a.numel(), _s()), "add_bf16")
    return out


def cast_f32_bf16(src, dst=None):
    dst = torch.empty(src.shape, dtype=bf16, device=src.device) if dst is None else dst
    L.check(L.lib().iadr1_cast_f32_bf16(_p(src), _p(dst), src.numel(), _s()), "cast_f32_bf16")
    return dst


# ---- linear ---------------------------------------------------------------------------------------------------------
def linear_fwd(x, W, bias=None, residual=None, out=None):
    """x[M,K] @ W[N,K]^T (+bias) (+residual) -> bf16 [M,N]   (torch.nn.functional.linear)"""
    return L.gemm(x, W, out=out, bias=bias, residual=residual)


def linear_bwd(dy, x, W, dW32, db32=None, need_dx=True, dx_out=None):
    """dx = dy @ W ; dW32 += dy^T @ x ; db32 += colsum(dy). All from the row-major buffers (MN-major operands)."""
    dx = None
    if need_dx:
        dx = L.gemm(dy, W.t(), out=dx_out)
    if dW32 is not None:
        L.gemm(dy.t(), x.t(), out=dW32, accumulate=True)
    if db32 is not None:
        colsum(dy, db32)
    return dx


# ---- attention ------------------------------------------------------------------------------------------------------
class AttnShape:
    """Geometry of one attention call over a fused qkv buffer [B*T, (nq + 2 nkv) * hd]."""

    def __init__(self, B, T, nq, nkv, hd, causal):
        self.B, self.T, self.nq, self.nkv, self.hd, self.causal = B, T, nq, nkv, hd, bool(causal)
        self.g = nq // nkv
        self.D = (nq + 2 * nkv) * hd
        self.Tp = ceil_to(T, 8)
        self.scale = float(hd) ** -0.5


def attention_fwd(qkv, sh: AttnShape, lo, hi, out=None, P=None):
    """Returns (attn [B*T, nq*hd] bf16, P [B, nq, T, Tp] bf16 probabilities kept for the backward)."""
    B, T, nq, hd, D, Tp, g = sh.B, sh.T, sh.nq, sh.hd, sh.D, sh.Tp, sh.g
    if P is None:
        P = torch.empty(B, nq, T, Tp, dtype=bf16, device=qkv.device)
    c = int(sh.causal)
    # S = scale * Q K^T  (per (row, head); kv head shared by g query heads)
    L.gemm_batched(qkv, qkv, P, M=T, N=T, K=hd, batch=B * nq, batch_lo=nq, b_lo_div=g,
                   lda=D, a_bs_lo=hd, a_bs_hi=T * D, ldb=D, b_bs_lo=hd, b_bs_hi=T * D, b_off=nq * hd,
                   ldc=Tp, c_bs_lo=T * Tp, c_bs_hi=nq * T * Tp, alpha=sh.scale, skip_mode=c)
    softmax_rows_(P, lo, hi, T, Tp, Tp, T * Tp, B * nq)
    if out is None:
        out = torch.empty(B * T, nq * hd, dtype=bf16, device=qkv.device)
    # O = P V   (V is read MN-major straight from the qkv buffer)
    L.gemm_batched(P, qkv, out, M=T, N=hd, K=T, batch=B * nq, batch_lo=nq, b_lo_div=g,
                   lda=Tp, a_bs_lo=T * Tp, a_bs_hi=nq * T * Tp,
                   ldb=D, b_bs_lo=hd, b_bs_hi=T * D, b_mn=1, b_off=(nq + sh.nkv) * hd,
                   ldc=nq * hd, c_bs_lo=hd, c_bs_hi=T * nq * hd, kmode=c)
    return out, P


def attention_bwd(dattn, qkv, P, sh: AttnShape, lo, hi, dqkv=None):
    """Gradient wrt the (post-rotary) fused qkv buffer. `P` is consumed (dP/dS are formed in a scratch of its size)."""
    B, T, nq, nkv, hd, D, Tp, g = sh.B, sh.T, sh.nq, sh.nkv, sh.hd, sh.D, sh.Tp, sh.g
    c = int(sh.causal)
    if dqkv is None:
        dqkv = torch.empty(B * T, D, dtype=bf16, device=qkv.device)
    dP = torch.empty_like(P)
    v_off, k_off = (nq + nkv) * hd, nq * hd
    # dP = dO V^T
    L.gemm_batched(dattn, qkv, dP, M=T, N=T, K=hd, batch=B * nq, batch_lo=nq, b_lo_div=g,
                   lda=nq * hd, a_bs_lo=hd, a_bs_hi=T * nq * hd, ldb=D, b_bs_lo=hd, b_bs_hi=T * D, b_off=v_off,
                   ldc=Tp, c_bs_lo=T * Tp, c_bs_hi=nq * T * Tp, skip_mode=c)
    softmax_bwd_rows_(P, dP, lo, hi, T, Tp, Tp, T * Tp, B * nq)  # dP now holds dS
    # dQ = scale * dS K
    L.gemm_batched(dP, qkv, dqkv, M=T, N=hd, K=T, batch=B * nq, batch_lo=nq, b_lo_div=g,
                   lda=Tp, a_bs_lo=T * Tp, a_bs_hi=nq * T * Tp, ldb=D, b_bs_lo=hd, b_bs_hi=T * D, b_mn=1, b_off=k_off,
                   ldc=D, c_bs_lo=hd, c_bs_hi=T * D, alpha=sh.scale, kmode=c)
    if g == 1:
        dk_buf, dk_ld, dk_off, dv_buf, dv_off = dqkv, D, k_off, dqkv, v_off
    else:
        tmp = torch.empty(2, B * T, nq * hd, dtype=bf16, device=qkv.device)
        dk_buf, dk_ld, dk_off, dv_buf, dv_off = tmp[0], nq * hd, 0, tmp[1], 0
    # dK_h = scale * dS^T Q_h ; dV_h = P^T dO_h   (A and B both MN-major; reduction over queries)
    L.gemm_batched(dP, qkv, dk_buf, M=T, N=hd, K=T, batch=B * nq, batch_lo=nq,
                   lda=Tp, a_bs_lo=T * Tp, a_bs_hi=nq * T * Tp, a_mn=1, ldb=D, b_bs_lo=hd, b_bs_hi=T * D, b_mn=1,
                   ldc=dk_ld, c_bs_lo=hd, c_bs_hi=T * dk_ld, c_off=dk_off, alpha=sh.scale, kmode=2 * c)
    L.gemm_batched(P, dattn, dv_buf, M=T, N=hd, K=T, batch=B * nq, batch_lo=nq,
                   lda=Tp, a_bs_lo=T * Tp, a_bs_hi=nq * T * Tp, a_mn=1,
                   ldb=nq * hd, b_bs_lo=hd, b_bs_hi=T * nq * hd, b_mn=1,
                   ldc=dk_ld, c_bs_lo=hd, c_bs_hi=T * dk_ld, c_off=dv_off, kmode=2 * c)
    if g > 1:
        group_sum(dk_buf, dqkv[:, k_off:], B * T, nkv, g, hd, nq * hd, D)
        group_sum(dv_buf, dqkv[:, v_off:], B * T, nkv, g, hd, nq * hd, D)
    return dqkv


# ---- fused attention plans (csrc/fmha_sm100.cu): built once per token layout, cached ----------------------------------
_PLAN_CACHE: dict = {}


def _cached_plan(key, build):
    p = _PLAN_CACHE.get(key)
    if p is None:
        if len(_PLAN_CACHE) > 256:
            _PLAN_CACHE.clear()
        p = _PLAN_CACHE[key] = build()
    return p


def range_attention(owner, tag, lo, hi, nq, nkv, hd):
    """Fused attention over per-row key ranges [lo[t], hi[t]) (vision windows / whole images / SigLIP crops; absolute token
    indices, any number of images in one stream). The plan is cached on `owner` (the geometry object the ranges belong to)."""
    cache = owner.__dict__.setdefault("_fused_attn", {})
    key = (tag, nq, nkv, hd)
    fa = cache.get(key)
    if fa is None:
        rng = F.range_rows(lo.cpu().numpy().astype(np.int64), hi.cpu().numpy().astype(np.int64))
        fa = cache[key] = F.FusedAttention(F.FmhaPlan(rng, lo.device, nkv=nkv, nq=nq), nq, nkv, hd)
    return fa


class FullAttention:
    """B sequences of T tokens, causal attention inside each sequence. Fused tcgen05 kernels when the head size allows
    (fmha.supported), else the composed QK^T -> softmax -> PV products."""

    def __init__(self, B, T, nq, nkv, hd, lo, hi, causal=True):
        self.sh = AttnShape(B, T, nq, nkv, hd, causal)
        self.lo, self.hi = lo, hi
        self.n_tokens = B * T
        self.fused = None
        if causal and F.supported(hd):
            dev = lo.device
            seqs = [(b * T, (b + 1) * T) for b in range(B)]
            plan = _cached_plan(("causal", B, T, nq, nkv, str(dev)),
                                lambda: F.FmhaPlan(F.causal_rows(B, T), dev, seqs, seqs, nkv=nkv, nq=nq))
            self.fused = F.FusedAttention(plan, nq, nkv, hd)

    def forward(self, qkv):
        if self.fused is not None:
            return self.fused.forward(qkv)
        return attention_fwd(qkv, self.sh, self.lo, self.hi)

    def backward(self, dattn, qkv, saved):
        if self.fused is not None:
            return self.fused.backward(dattn, qkv, saved)
        return attention_bwd(dattn, qkv, saved, self.sh, self.lo, self.hi)


def _plan_args(geom, device):
    rng, probs, segs = geom
    return rng, device, probs, segs


class SharedPrefixAttention:
    """One prompt of P tokens followed by G completions of C tokens, laid out [prefix | row 0 | row 1 | ...].

    The G rows of a GRPO group share their prompt (ref: sc_grpo_trainer.py:624-628 repeats the prompt tensors G times and
    runs the full [G, P + C] batch); causal attention makes the prompt's hidden states identical in every row, so they
    are computed once: prefix tokens attend causally to the prefix, completion token c of row g attends to the whole
    prefix plus tokens <= c of its own row. Scores live in S[nq, G*C, Pp + Cp] with the prefix keys in columns [0, P),
    a masked alignment gap [P, Pp) and the row's own keys from column Pp. Gradients into the prefix K/V sum over rows
    inside the GEMM reduction (the K dimension runs over all G*C queries)."""

    def __init__(self, P, G, C, nq, nkv, hd, device):
        self.P, self.G, self.C, self.nq, self.nkv, self.hd = P, G, C, nq, nkv, hd
        self.g = nq // nkv
        self.D = (nq + 2 * nkv) * hd
        self.Pp, self.Cp = ceil_to(P, 8), ceil_to(C, 8)
        self.Tk = self.Pp + self.Cp
        self.scale = float(hd) ** -0.5
        self.n_tokens = P + G * C
        self.prefix = AttnShape(1, P, nq, nkv, hd, True)
        self.p_lo = torch.zeros(P, dtype=torch.int32, device=device)
        self.p_hi = torch.arange(1, P + 1, dtype=torch.int32, device=device)
        self.c_lo = torch.zeros(G * C, dtype=torch.int32, device=device)
        self.c_hi = (self.Pp + (torch.arange(G * C, device=device) % C) + 1).to(torch.int32)
        self.fused = None
        if F.supported(hd):
            plan = _cached_plan(("shared", P, G, C, nq, nkv, str(device)),
                                lambda: F.FmhaPlan(*_plan_args(F.shared_prefix_geometry(P, G, C), device), nkv=nkv, nq=nq))
            self.fused = F.FusedAttention(plan, nq, nkv, hd)

    def forward(self, qkv, out=None):
        if self.fused is not None:
            return self.fused.forward(qkv, out=out)
        P, G, C, nq, nkv, hd, g, D, Pp, Tk = self.P, self.G, self.C, self.nq, self.nkv, self.hd, self.g, self.D, self.Pp, self.Tk
        GC = G * C
        attn = torch.empty(self.n_tokens, nq * hd, dtype=bf16, device=qkv.device) if out is None else out
        _, Pm = attention_fwd(qkv[:P], self.prefix, self.p_lo, self.p_hi, out=attn[:P])
        S = torch.empty(nq, GC, Tk, dtype=bf16, device=qkv.device)
        k_off, v_off = nq * hd, (nq + nkv) * hd
        L.gemm_batched(qkv, qkv, S, M=GC, N=P, K=hd, batch=nq, batch_lo=nq, b_lo_div=g, lda=D, a_bs_lo=hd, a_off=P * D,
                       ldb=D, b_bs_lo=hd, b_off=k_off, ldc=Tk, c_bs_lo=GC * Tk, alpha=self.scale)
        L.gemm_batched(qkv, qkv, S, M=C, N=C, K=hd, batch=G * nq, batch_lo=nq, b_lo_div=g, lda=D, a_bs_lo=hd,
                       a_bs_hi=C * D, a_off=P * D, ldb=D, b_bs_lo=hd, b_bs_hi=C * D, b_off=P * D + k_off, ldc=Tk,
                       c_bs_lo=GC * Tk, c_bs_hi=C * Tk, c_off=Pp, alpha=self.scale, skip_mode=1)
        softmax_rows_(S, self.c_lo, self.c_hi, GC, Tk, Tk, GC * Tk, nq, hole=(P, Pp))
        out = attn[P:]
        L.gemm_batched(S, qkv, out, M=GC, N=hd, K=P, batch=nq, batch_lo=nq, b_lo_div=g, lda=Tk, a_bs_lo=GC * Tk,
                       ldb=D, b_bs_lo=hd, b_mn=1, b_off=v_off, ldc=nq * hd, c_bs_lo=hd)
        L.gemm_batched(S, qkv, out, M=C, N=hd, K=C, batch=G * nq, batch_lo=nq, b_lo_div=g, lda=Tk, a_bs_lo=GC * Tk,
                       a_bs_hi=C * Tk, a_off=Pp, ldb=D, b_bs_lo=hd, b_bs_hi=C * D, b_mn=1, b_off=P * D + v_off,
                       ldc=nq * hd, c_bs_lo=hd, c_bs_hi=C * nq * hd, kmode=1, residual=out)
        return attn, (Pm, S)

    def backward(self, dattn, qkv, saved, out=None):
        if self.fused is not None:
            return self.fused.backward(dattn, qkv, saved, out=out)
        P, G, C, nq, nkv, hd, g, D, Pp, Tk = self.P, self.G, self.C, self.nq, self.nkv, self.hd, self.g, self.D, self.Pp, self.Tk
        GC, QH = G * C, nq * hd
        Pm, S = saved
        dev = qkv.device
        dqkv = torch.empty(self.n_tokens, D, dtype=bf16, device=dev) if out is None else out
        attention_bwd(dattn[:P], qkv[:P], Pm, self.prefix, self.p_lo, self.p_hi, dqkv=dqkv[:P])
        k_off, v_off = nq * hd, (nq + nkv) * hd
        dP = torch.empty_like(S)
        # dP = dO V^T over both key segments
        L.gemm_batched(dattn, qkv, dP, M=GC, N=P, K=hd, batch=nq, batch_lo=nq, b_lo_div=g, lda=QH, a_bs_lo=hd,
                       a_off=P * QH, ldb=D, b_bs_lo=hd, b_off=v_off, ldc=Tk, c_bs_lo=GC * Tk)
        L.gemm_batched(dattn, qkv, dP, M=C, N=C, K=hd, batch=G * nq, batch_lo=nq, b_lo_div=g, lda=QH, a_bs_lo=hd,
                       a_bs_hi=C * QH, a_off=P * QH, ldb=D, b_bs_lo=hd, b_bs_hi=C * D, b_off=P * D + v_off, ldc=Tk,
                       c_bs_lo=GC * Tk, c_bs_hi=C * Tk, c_off=Pp, skip_mode=1)
        softmax_bwd_rows_(S, dP, self.c_lo, self.c_hi, GC, Tk, Tk, GC * Tk, nq, hole=(P, Pp))  # dP now holds dS
        # dQ_c = scale * (dS[:, :P] K_p + dS[:, Pp:] K_g)
        L.gemm_batched(dP, qkv, dqkv, M=GC, N=hd, K=P, batch=nq, batch_lo=nq, b_lo_div=g, lda=Tk, a_bs_lo=GC * Tk,
                       ldb=D, b_bs_lo=hd, b_mn=1, b_off=k_off, ldc=D, c_bs_lo=hd, c_off=P * D, alpha=self.scale)
        L.gemm_batched(dP, qkv, dqkv, M=C, N=hd, K=C, batch=G * nq, batch_lo=nq, b_lo_div=g, lda=Tk, a_bs_lo=GC * Tk,
                       a_bs_hi=C * Tk, a_off=Pp, ldb=D, b_bs_lo=hd, b_bs_hi=C * D, b_mn=1, b_off=P * D + k_off,
                       ldc=D, c_bs_lo=hd, c_bs_hi=C * D, c_off=P * D, alpha=self.scale, kmode=1, residual=dqkv)
        # prefix keys / values: reduction over ALL G*C completion queries in one product per head
        tp = torch.empty(2, P, QH, dtype=bf16, device=dev)
        L.gemm_batched(dP, qkv, tp[0], M=P, N=hd, K=GC, batch=nq, batch_lo=nq, lda=Tk, a_bs_lo=GC * Tk, a_mn=1,
                       ldb=D, b_bs_lo=hd, b_mn=1, b_off=P * D, ldc=QH, c_bs_lo=hd, alpha=self.scale)
        L.gemm_batched(S, dattn, tp[1], M=P, N=hd, K=GC, batch=nq, batch_lo=nq, lda=Tk, a_bs_lo=GC * Tk, a_mn=1,
                       ldb=QH, b_bs_lo=hd, b_mn=1, b_off=P * QH, ldc=QH, c_bs_lo=hd)
        group_sum(tp[0], dqkv[:, k_off:], P, nkv, g, hd, QH, D, accumulate=True)
        group_sum(tp[1], dqkv[:, v_off:], P, nkv, g, hd, QH, D, accumulate=True)
        # each row's own keys / values
        tc = torch.empty(2, GC, QH, dtype=bf16, device=dev)
        L.gemm_batched(dP, qkv, tc[0], M=C, N=hd, K=C, batch=G * nq, batch_lo=nq, lda=Tk, a_bs_lo=GC * Tk,
                       a_bs_hi=C * Tk, a_off=Pp, a_mn=1, ldb=D, b_bs_lo=hd, b_bs_hi=C * D, b_mn=1, b_off=P * D,
                       ldc=QH, c_bs_lo=hd, c_bs_hi=C * QH, alpha=self.scale, kmode=2)
        L.gemm_batched(S, dattn, tc[1], M=C, N=hd, K=C, batch=G * nq, batch_lo=nq, lda=Tk, a_bs_lo=GC * Tk,
                       a_bs_hi=C * Tk, a_off=Pp, a_mn=1, ldb=QH, b_bs_lo=hd, b_bs_hi=C * QH, b_mn=1, b_off=P * QH,
                       ldc=QH, c_bs_lo=hd, c_bs_hi=C * QH, kmode=2)
        group_sum(tc[0], dqkv[P:, k_off:], GC, nkv, g, hd, QH, D)
        group_sum(tc[1], dqkv[P:, v_off:], GC, nkv, g, hd, QH, D)
        return dqkv


class MultiGroupAttention:
    """Several independent groups packed back to back in one token stream ([group 0 layout | group 1 layout | ...]):
    every row-wise kernel and dense product then runs once over all of them (bigger M, one weight-gradient
    read-modify-write per pass instead of one per group), attention runs group by group on slices."""

    def __init__(self, groups):
        self.groups = groups
        self.offsets = [0]
        for g_ in groups:
            self.offsets.append(self.offsets[-1] + g_.n_tokens)
        self.n_tokens = self.offsets[-1]
        # fused kernels: ONE plan over all packed groups (one launch per pass instead of one per group)
        self.fused = None
        if all(getattr(g_, "fused", None) is not None and isinstance(g_, SharedPrefixAttention) for g_ in groups):
            t0 = groups[0]
            dev = t0.p_lo.device

            def build():
                rngs, probs, segs = [], [], []
                for g_, base in zip(groups, self.offsets):
                    r, p_, s_ = F.shared_prefix_geometry(g_.P, g_.G, g_.C, base)
                    rngs.append(r)
                    probs += p_
                    segs += s_
                return F.FmhaPlan(np.concatenate(rngs), dev, probs, segs, nkv=t0.nkv, nq=t0.nq)

            plan = _cached_plan(("multi", tuple((g_.P, g_.G, g_.C) for g_ in groups), t0.nq, t0.nkv, str(dev)), build)
            self.fused = F.FusedAttention(plan, t0.nq, t0.nkv, t0.hd)

    def forward(self, qkv):
        if self.fused is not None:
            return self.fused.forward(qkv)
        t0 = self.groups[0]
        out = torch.empty(self.n_tokens, t0.nq * t0.hd, dtype=bf16, device=qkv.device)
        saved = []
        for g_, a, b in zip(self.groups, self.offsets[:-1], self.offsets[1:]):
            _, s_ = g_.forward(qkv[a:b], out=out[a:b])
            saved.append(s_)
        return out, saved

    def backward(self, dattn, qkv, saved):
        if self.fused is not None:
            return self.fused.backward(dattn, qkv, saved)
        dqkv = torch.empty(self.n_tokens, self.groups[0].D, dtype=bf16, device=qkv.device)
        for g_, s_, a, b in zip(self.groups, saved, self.offsets[:-1], self.offsets[1:]):
            g_.backward(dattn[a:b], qkv[a:b], s_, out=dqkv[a:b])
        return dqkv


# ---- fused lm_head -> log-softmax -> gather -----------------------------------------------------------------------------
def logprob_fwd(h, E, labels, temperature: float = 1.0):
    """logp[m] = log_softmax(h[m] @ E^T / temperature)[labels[m]] in fp32 without materialising the logits.

    Replaces `model(**inputs).logits` + row-wise `log_softmax` + `gather` (sc_grpo_trainer.py:505-514).
    Returns (logp [M] fp32, lse [M] fp32)."""
    M, H = h.shape
    V = E.shape[0]
    bn = L.lib().iadr1_gemm_pick_block_n(V, 0)
    tiles = (V + bn - 1) // bn
    dev = h.device
    pmax = torch.empty(M, tiles, dtype=f32, device=dev)
    psum = torch.empty(M, tiles, dtype=f32, device=dev)
    tgt = torch.zeros(M, dtype=f32, device=dev)
    d = L.GemmDesc()
    d.M, d.N, d.K = M, V, H
    d.batch = d.batch_lo = d.b_lo_div = 1
    d.A, d.lda, d.a_mn = h.data_ptr(), h.stride(0), 0
    d.B, d.ldb, d.b_mn = E.data_ptr(), E.stride(0), 0
    d.split_k = 1
    d.alpha = 1.0 / temperature
    d.epi = EPI_LSE
    d.labels, d.part_max, d.part_sum, d.tgt_logit, d.lse_tiles_n = _p(labels), _p(pmax), _p(psum), _p(tgt), tiles
    d.block_n = bn
    L.check(L.lib().iadr1_gemm_bf16(C.byref(d), _s()), "logprob_fwd gemm")
    lse = torch.empty(M, dtype=f32, device=dev)
    logp = torch.empty(M, dtype=f32, device=dev)
    L.check(L.lib().iadr1_lse_finalize(_p(pmax), _p(psum), _p(tgt), tiles, M, _p(lse), _p(logp), _s()), "lse_finalize")
    return logp, lse


def logprob_bwd(dlogp, h, E, labels, lse, dE32, temperature: float = 1.0, need_dh=True):
    """dh = dlogits @ E, dE32 += dlogits^T @ h with dlogits = (onehot - softmax) * dlogp / temperature, recomputing
    the logits tile by tile (bf16 dlogits [M, V] is the only large intermediate)."""
    M, H = h.shape
    V = E.shape[0]
    dlogits = torch.empty(M, V, dtype=bf16, device=h.device)
    gscale = (-dlogp.to(f32) / temperature).contiguous()  # kernel forms (softmax - onehot) * gscale
    d = L.GemmDesc()
    d.M, d.N, d.K = M, V, H
    d.batch = d.batch_lo = d.b_lo_div = 1
    d.A, d.lda, d.a_mn = h.data_ptr(), h.stride(0), 0
    d.B, d.ldb, d.b_mn = E.data_ptr(), E.stride(0), 0
    d.C, d.ldc = dlogits.data_ptr(), V
    d.split_k = 1
    d.alpha = 1.0 / temperature
    d.epi = EPI_DLOGITS
    d.labels, d.lse, d.gscale = _p(labels), _p(lse), _p(gscale)
    L.check(L.lib().iadr1_gemm_bf16(C.byref(d), _s()), "logprob_bwd gemm")
    dh = L.gemm(dlogits, E.t()) if need_dh else None
    if dE32 is not None:
        L.gemm(dlogits.t(), h.t(), out=dE32, accumulate=True)
    return dh
