"""GPU parity of every library kernel (called through the C ABI via ctypes) against plain torch fp32 math on the
same seeded inputs. Tolerances: bf16 outputs are compared at 2 bf16 ulps of the reference magnitude (stated per
test); fp32 outputs at 1e-4 relative. Integer / index work is bit-exact."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

bf16, f32 = torch.bfloat16, torch.float32


@pytest.fixture(scope="module")
def ops(cuda):
    from iad_r1_b200 import ops
    from iad_r1_b200 import lib
    lib.lib()  # must load: no fallback
    return ops


def rnd(*shape, scale=1.0, dev="cuda"):
    return (torch.randn(*shape, device=dev) * scale).to(bf16)


def close(got, want, tol=2 ** -7, what=""):
    got, want = got.float(), want.float()
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    err = (got - want).abs().max().item()
    ref = want.abs().max().item() + 1e-6
    assert err <= tol * ref, f"{what}: max abs err {err:.4g} vs ref max {ref:.4g} (tol {tol:.3g} rel)"


# ---------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (200, 72, 1176), (1024, 1280, 1176), (333, 3424, 1280), (8, 2048, 896)])
@pytest.mark.parametrize("amn,bmn", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_layouts(ops, M, N, K, amn, bmn):
    from iad_r1_b200 import lib as L
    torch.manual_seed(M + N + K + amn * 2 + bmn)
    if (amn and M % 8) or (bmn and N % 8):
        pytest.skip("MN-major operand needs a 16-byte aligned leading stride")
    a = rnd(K, M).t() if amn else rnd(M, K)
    b = rnd(K, N).t() if bmn else rnd(N, K)
    out = L.gemm(a, b)
    close(out, a.float() @ b.float().t(), 2 ** -7, "gemm")


def test_gemm_epilogues(ops):
    from iad_r1_b200 import lib as L
    torch.manual_seed(1)
    M, N, K = 384, 520, 264
    a, b = rnd(M, K), rnd(N, K)
    want = a.float() @ b.float().t()
    bias, res = rnd(N), rnd(M, N)
    close(L.gemm(a, b, bias=bias, residual=res), want + bias.float() + res.float(), 2 ** -7, "bias+residual")
    close(L.gemm(a, b, alpha=0.25, out_dtype=f32), 0.25 * want, 1e-5, "alpha f32")
    acc = torch.full((M, N), 2.0, device="cuda")
    L.gemm(a, b, out=acc, accumulate=True)
    close(acc, want + 2.0, 1e-5, "accumulate")
    close(L.gemm(a, b, split_k=3, out_dtype=f32, bias=bias), want + bias.float(), 1e-5, "split-k atomic + bias once")
    close(L.gemm(a, b, trans_out=True), want.t(), 2 ** -7, "transposed store")
    close(L.gemm(a, b, raster=2), want, 2 ** -7, "N-fastest tile order")
    close(L.gemm(a, b, raster=1, bias=bias), want + bias.float(), 2 ** -7, "M-fastest tile order")
    w, x = rnd(2048, 1024), rnd(8, 1024)
    bm = rnd(2048)
    close(L.gemm(w, x, trans_out=True, split_k=4, out_dtype=f32, block_n=16, bias=bm, bias_per_m=True),
          x.float() @ w.float().t() + bm.float(), 1e-5, "decode swap-AB")


@pytest.mark.parametrize("N_out,K_in,T", [(2560, 2048, 1111), (1280, 1280, 1024), (3840, 1280, 520), (2048, 11008, 300)])
def test_gemm_weight_gradient_wave_quantised_tiles(ops, N_out, K_in, T):
    """dW[N, K] += dy[T, N]^T x[T, K], fp32 accumulate: the shapes whose 256-column tiling leaves a ragged last wave take 192- or
    128-column tiles (gemm_sm100.cu: wave quantisation); values must not depend on the tile width."""
    from iad_r1_b200 import lib as L
    torch.manual_seed(N_out + T)
    dy, x = rnd(T, N_out), rnd(T, K_in)
    base = torch.randn(N_out, K_in, device="cuda")
    want = base + dy.float().t() @ x.float()
    got = base.clone()
    L.gemm(dy.t(), x.t(), out=got, accumulate=True, out_dtype=f32)
    close(got, want, 2 ** -9, "wgrad")
    ref = base.clone()
    L.gemm(dy.t(), x.t(), out=ref, accumulate=True, out_dtype=f32, block_n=256)
    # stream-K shapes accumulate a tile's K range in two or three fp32 partial sums added in arrival order: equal up to the
    # fp32 rounding of ~K products (measured 2e-4 on values up to 180), far inside the bf16-product tolerance above
    close(got, ref, 2 ** -17, "tile width / stream-K changed the accumulated values")


@pytest.mark.parametrize("F,K,R", [(22016, 2048, 64), (2560, 2048, 64), (2048, 11008, 16), (1000, 200, 8), (128, 64, 24)])
def test_gemm_stream_k(ops, F, K, R):
    """Stream-K schedule of the decode products (k-block units spread evenly over the SMs, fp32 atomics, bias added once
    per tile) and the fp32 SwiGLU that consumes + clears the accumulator."""
    from iad_r1_b200 import lib as L
    torch.manual_seed(F + K)
    w, x, bm = rnd(F, K, scale=0.05), rnd(R, K), rnd(F)
    out = torch.zeros(R, F, device="cuda")
    L.gemm(w, x, out=out, trans_out=True, atomic=True, stream_k=True, block_n=max(16, (R + 15) // 16 * 16), bias=bm,
           bias_per_m=True, a_static=True, co_resident=(F % 256 == 0))   # + the two-CTAs-per-SM decode variant
    close(out, x.float() @ w.float().t() + bm.float(), 1e-4, "stream-k")
    with pytest.raises(L.NativeLibraryError):
        L.gemm(w, x, trans_out=True, stream_k=True)       # needs atomic fp32 output
    if F % 8 == 0:
        I = F // 2
        gu = out.clone()
        act = torch.empty(R, I, dtype=bf16, device="cuda")
        L.check(L.lib().iadr1_decode_silu_mul_f32(gu.data_ptr(), act.data_ptr(), R, I, L.stream_ptr()))
        g, u = out[:, :I].to(bf16).float(), out[:, I:].to(bf16).float()
        want = (torch.nn.functional.silu(g).to(bf16).float() * u).to(bf16)
        close(act, want, 2 ** -7, "silu_mul_f32")
        assert (gu == 0).all(), "accumulator must be cleared"


@pytest.mark.parametrize("I,K,R", [(11008, 2048, 64), (4864, 896, 56), (256, 128, 8), (11008, 2048, 128), (1024, 512, 100)])
def test_gemm_swiglu(ops, I, K, R):
    """Decode MLP front half in one launch (gate and up rows of the fused weight share a 128-row MMA tile, SwiGLU in the
    epilogue) vs the two-step torch form with the same bf16 rounding points as Qwen2MLP."""
    from iad_r1_b200 import lib as L
    torch.manual_seed(I + K)
    w, x = rnd(2 * I, K, scale=0.05), rnd(R, K)
    out = torch.empty(R, I, dtype=bf16, device="cuda")
    L.gemm_swiglu(w, x, out, block_n=max(16, (R + 15) // 16 * 16))
    gu = (x.float() @ w.float().t()).to(bf16).float()
    want = (torch.nn.functional.silu(gu[:, :I]).to(bf16).float() * gu[:, I:]).to(bf16)
    close(out, want, 2 ** -6, "swiglu")


def test_gemm_empty_and_errors(ops):
    from iad_r1_b200 import lib as L
    a, b = rnd(0, 64), rnd(16, 64)
    assert L.gemm(a, b).shape == (0, 16)
    with pytest.raises(L.NativeLibraryError):
        L.gemm(rnd(8, 60)[:, :59], rnd(8, 59))  # leading stride not a multiple of 8 elements


# ---------------------------------------------------------------------------------------------- norms
@pytest.mark.parametrize("rows,cols", [(1, 64), (37, 1280), (300, 2048), (16, 3584)])
def test_rmsnorm(ops, rows, cols):
    torch.manual_seed(rows)
    x, w, dy = rnd(rows, cols), (1 + 0.1 * torch.randn(cols, device="cuda")).to(bf16), rnd(rows, cols)
    eps = 1e-6
    y, rstd = ops.rmsnorm_fwd(x, w, eps)
    xf = x.float().requires_grad_(True)
    wf = w.float().requires_grad_(True)
    var = xf.pow(2).mean(-1, keepdim=True)
    ref = wf * (xf * torch.rsqrt(var + eps))
    close(y, ref, 2 ** -7, "rmsnorm fwd")
    close(rstd, torch.rsqrt(var + eps).squeeze(-1), 1e-4, "rstd")
    ref.backward(dy.float())
    dx = rnd(rows, cols)
    dx0 = dx.clone()
    dw = torch.zeros(cols, device="cuda")
    ops.rmsnorm_bwd(dy, x, w, rstd, dx, dw, add_dx=True)
    close(dx, dx0.float() + xf.grad, 2 ** -6, "rmsnorm dx (+=)")
    close(dw, wf.grad, 2e-3, "rmsnorm dw")


@pytest.mark.parametrize("rows,cols", [(5, 1280), (129, 1152)])
def test_layernorm(ops, rows, cols):
    torch.manual_seed(rows)
    x, dy = rnd(rows, cols), rnd(rows, cols)
    w, b = (1 + 0.1 * torch.randn(cols, device="cuda")).to(bf16), rnd(cols, scale=0.1)
    y, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-6)
    xf, wf, bf = x.float().requires_grad_(True), w.float().requires_grad_(True), b.float().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xf, (cols,), wf, bf, 1e-6)
    close(y, ref, 2 ** -7, "layernorm fwd")
    ref.backward(dy.float())
    dx, dw, db = torch.empty_like(x), torch.zeros(cols, device="cuda"), torch.zeros(cols, device="cuda")
    ops.layernorm_bwd(dy, x, w, mean, rstd, dx, dw, db, add_dx=False)
    close(dx, xf.grad, 2 ** -6, "layernorm dx")
    close(dw, wf.grad, 2e-3, "layernorm dw")
    close(db, bf.grad, 2e-3, "layernorm db")


# ---------------------------------------------------------------------------------------------- rope / activations
def _rot_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), -1)


@pytest.mark.parametrize("hd,bf16_ops", [(128, 1), (80, 0), (64, 1)])
def test_rope(ops, hd, bf16_ops):
    torch.manual_seed(hd)
    tokens, heads, extra = 77, 6, 2
    x = rnd(tokens, (heads + extra) * hd)
    ang = torch.rand(tokens, hd // 2, device="cuda") * 50
    emb = torch.cat((ang, ang), -1)
    cos, sin = emb.cos().contiguous(), emb.sin().contiguous()
    x0 = x.clone()
    ops.rope_(x, cos, sin, heads, hd, bf16_ops)
    xv = x0.view(tokens, heads + extra, hd)[:, :heads]
    if bf16_ops:
        c, s = cos.to(bf16)[:, None], sin.to(bf16)[:, None]
        ref = (xv * c) + (_rot_half(xv) * s)  # bf16 arithmetic, as HF
        assert torch.equal(x.view(tokens, heads + extra, hd)[:, :heads], ref), "text rope must be bit-exact vs bf16 torch"
    else:
        ref = (xv.float() * cos[:, None] + _rot_half(xv.float()) * sin[:, None]).to(bf16)
        close(x.view(tokens, heads + extra, hd)[:, :heads], ref, 2 ** -8, "vision rope")
    assert torch.equal(x.view(tokens, heads + extra, hd)[:, heads:], x0.view(tokens, heads + extra, hd)[:, heads:])
    # backward = transpose of the rotation: <R x, y> == <x, R^T y>
    y = rnd(tokens, (heads + extra) * hd)
    yb = y.clone()
    ops.rope_(yb, cos, sin, heads, hd, 0, backward=True)
    xf = x0.clone()
    ops.rope_(xf, cos, sin, heads, hd, 0)
    sl = slice(0, heads * hd)
    lhs = (xf[:, sl].float() * y[:, sl].float()).sum()
    rhs = (x0[:, sl].float() * yb[:, sl].float()).sum()
    assert abs(lhs - rhs) <= 2e-2 * abs(lhs) + 1.0


@pytest.mark.parametrize("act,gated", [(0, True), (1, False), (2, False), (3, False), (0, False)])
def test_act_mul(ops, act, gated):
    torch.manual_seed(act)
    rows, cols = 45, 1712
    gu = rnd(rows, cols * (2 if gated else 1))
    dout = rnd(rows, cols)
    out = ops.act_mul_fwd(gu, cols, act, gated)
    g = gu[:, :cols].float().requires_grad_(True)
    u = gu[:, cols:].float().requires_grad_(True) if gated else None
    fn = {0: torch.nn.functional.silu, 1: torch.nn.functional.gelu, 2: lambda t: t * torch.sigmoid(1.702 * t),
          3: lambda t: torch.nn.functional.gelu(t, approximate="tanh")}[act]
    ref = fn(g) * u if gated else fn(g)
    close(out, ref, 2 ** -6, "act fwd")
    ref.backward(dout.float())
    dgu = ops.act_mul_bwd(dout, gu, cols, act, gated)
    close(dgu[:, :cols], g.grad, 2 ** -6, "act dgate")
    if gated:
        close(dgu[:, cols:], u.grad, 2 ** -6, "act dup")


# ---------------------------------------------------------------------------------------------- softmax / gathers
def test_softmax_rows_and_bwd(ops):
    torch.manual_seed(3)
    Z, Tq, Tk = 6, 50, 56
    S = rnd(Z, Tq, Tk, scale=3.0)
    S[:, :, 40:] = float("nan")  # never-written tiles must not leak
    lo = torch.randint(0, 5, (Tq,), device="cuda", dtype=torch.int32)
    hi = (lo + torch.randint(1, 35, (Tq,), device="cuda", dtype=torch.int32)).to(torch.int32)
    S0 = S.clone()
    ops.softmax_rows_(S, lo, hi, Tq, Tk, Tk, Tq * Tk, Z)
    k = torch.arange(Tk, device="cuda")
    mask = (k[None, :] >= lo[:, None]) & (k[None, :] < hi[:, None])
    ref = torch.softmax(torch.nan_to_num(S0.float()).masked_fill(~mask, float("-inf")), -1)
    close(S, ref, 2 ** -7, "softmax")
    assert (S.masked_select(~mask.expand(Z, -1, -1)) == 0).all()
    dP = rnd(Z, Tq, Tk)
    dP[:, :, 40:] = float("inf")
    dP0 = dP.clone()
    P = S.clone()
    ops.softmax_bwd_rows_(P, dP, lo, hi, Tq, Tk, Tk, Tq * Tk, Z)
    assert torch.isfinite(dP).all()
    # values under the ranged mask: dS = P * (dP - rowsum(P * dP)) inside [lo, hi), exact zeros outside
    Pf = S.float()
    dPf = torch.nan_to_num(dP0.float(), posinf=0.0).masked_fill(~mask, 0.0)
    ref_ds = Pf * (dPf - (Pf * dPf).sum(-1, keepdim=True))
    close(dP, ref_ds, 2 ** -6, "softmax bwd (ranged mask)")
    assert (dP.masked_select(~mask.expand(Z, -1, -1)) == 0).all()


def test_softmax_hole_mask_fwd_bwd(ops):
    """The shared-prefix score layout [prefix keys | alignment gap (hole) | own-row keys]: the hole columns are masked in
    the forward and carry exact zeros through the backward (ops.SharedPrefixAttention)."""
    torch.manual_seed(6)
    Z, Tq, P, Pp, Cc = 4, 48, 21, 24, 48
    Tk = Pp + Cc
    S = rnd(Z, Tq, Tk, scale=2.0)
    S0 = S.clone()
    lo = torch.zeros(Tq, device="cuda", dtype=torch.int32)
    hi = (Pp + torch.arange(Tq, device="cuda") % Cc + 1).to(torch.int32)
    ops.softmax_rows_(S, lo, hi, Tq, Tk, Tk, Tq * Tk, Z, hole=(P, Pp))
    k = torch.arange(Tk, device="cuda")
    mask = (k[None, :] < hi[:, None]) & ~((k[None, :] >= P) & (k[None, :] < Pp))
    Sf = S0.float().requires_grad_(True)
    ref = torch.softmax(Sf.masked_fill(~mask, float("-inf")), -1)
    close(S, ref, 2 ** -7, "softmax (hole)")
    assert (S.masked_select(~mask.expand(Z, -1, -1)) == 0).all()
    dP = rnd(Z, Tq, Tk)
    ref.backward(dP.float())
    d = dP.clone()
    ops.softmax_bwd_rows_(S.clone(), d, lo, hi, Tq, Tk, Tk, Tq * Tk, Z, hole=(P, Pp))
    close(d, Sf.grad, 2 ** -5, "softmax bwd (hole)")
    assert (d.masked_select(~mask.expand(Z, -1, -1)) == 0).all()


def test_softmax_bwd_values(ops):
    torch.manual_seed(4)
    Z, T = 3, 40
    S = rnd(Z, T, T, scale=2.0)
    lo = torch.zeros(T, device="cuda", dtype=torch.int32)
    hi = torch.arange(1, T + 1, device="cuda", dtype=torch.int32)
    Sf = S.float().requires_grad_(True)
    mask = torch.tril(torch.ones(T, T, device="cuda", dtype=torch.bool))
    Pf = torch.softmax(Sf.masked_fill(~mask, float("-inf")), -1)
    dP = rnd(Z, T, T)
    Pf.backward(dP.float())
    P = S.clone()
    ops.softmax_rows_(P, lo, hi, T, T, T, T * T, Z)
    d = dP.clone()
    ops.softmax_bwd_rows_(P, d, lo, hi, T, T, T, T * T, Z)
    close(d, Sf.grad, 2 ** -5, "softmax bwd")


def test_gather_scatter_colsum_groupsum(ops):
    torch.manual_seed(5)
    table, alt = rnd(100, 64), rnd(7, 64)
    index = torch.tensor([3, 99, -1, -7, 0, 3, 3, -2], device="cuda", dtype=torch.int32)
    out = ops.gather_rows(table, index, alt)
    ref = torch.stack([table[i] if i >= 0 else alt[-1 - i] for i in index.tolist()])
    assert torch.equal(out, ref)
    d = rnd(8, 64)
    dt, da = torch.zeros(100, 64, device="cuda"), torch.zeros(7, 64, device="cuda")
    ops.scatter_add_rows(d, index, dt, da)
    rt, ra = torch.zeros_like(dt), torch.zeros_like(da)
    for r, i in enumerate(index.tolist()):
        (rt if i >= 0 else ra)[i if i >= 0 else -1 - i] += d[r].float()
    close(dt, rt, 1e-5, "scatter table")
    close(da, ra, 1e-5, "scatter alt")
    x = rnd(1000, 200)
    cs = torch.ones(200, device="cuda")
    ops.colsum(x, cs)
    close(cs, x.float().sum(0) + 1, 1e-4, "colsum")
    src = rnd(33, 4 * 3 * 16)
    dst = torch.zeros(33, 200, dtype=bf16, device="cuda")
    ops.group_sum(src, dst[:, 8:], 33, 4, 3, 16, src.stride(0), dst.stride(0))
    close(dst[:, 8:8 + 64], src.view(33, 4, 3, 16).float().sum(2).reshape(33, 64), 2 ** -7, "group_sum")


# ---------------------------------------------------------------------------------------------- attention
@pytest.mark.parametrize("B,T,nq,nkv,hd,causal", [(2, 200, 4, 2, 128, True), (1, 256, 4, 4, 80, False),
                                                  (3, 72, 6, 2, 16, True), (2, 832, 16, 2, 128, True)])
def test_attention_fwd_bwd(ops, B, T, nq, nkv, hd, causal):
    torch.manual_seed(T)
    sh = ops.AttnShape(B, T, nq, nkv, hd, causal)
    qkv = rnd(B * T, sh.D, scale=0.7)
    dattn = rnd(B * T, nq * hd)
    if causal:
        lo = torch.zeros(T, device="cuda", dtype=torch.int32)
        hi = torch.arange(1, T + 1, device="cuda", dtype=torch.int32)
    else:  # windows of 64 tokens
        lo = (torch.arange(T, device="cuda") // 64 * 64).to(torch.int32)
        hi = torch.clamp(lo + 64, max=T).to(torch.int32)
    attn, P = ops.attention_fwd(qkv, sh, lo, hi)
    x = qkv.float().view(B, T, nq + 2 * nkv, hd).requires_grad_(True)
    q, k, v = x[:, :, :nq], x[:, :, nq:nq + nkv], x[:, :, nq + nkv:]
    kk, vv = k.repeat_interleave(sh.g, 2), v.repeat_interleave(sh.g, 2)
    s = torch.einsum("bqhd,bkhd->bhqk", q, kk) * sh.scale
    kidx = torch.arange(T, device="cuda")
    mask = (kidx[None, :] >= lo[:, None]) & (kidx[None, :] < hi[:, None])
    p = torch.softmax(s.masked_fill(~mask, float("-inf")), -1)
    ref = torch.einsum("bhqk,bkhd->bqhd", p, vv).reshape(B * T, nq * hd)
    close(attn, ref, 2 ** -6, "attention fwd")
    ref.backward(dattn.float())
    dqkv = ops.attention_bwd(dattn, qkv, P, sh, lo, hi)
    close(dqkv, x.grad.reshape(B * T, sh.D), 2 ** -5, "attention bwd")


# ---------------------------------------------------------------------------------------------- fused lm_head
@pytest.mark.parametrize("M,H,V,temp", [(70, 64, 1000, 1.0), (512, 256, 151936, 1.0), (33, 128, 4099 // 8 * 8, 0.9)])
def test_logprob_fwd_bwd(ops, M, H, V, temp):
    torch.manual_seed(V)
    h, E = rnd(M, H), rnd(V, H, scale=0.3)
    labels = torch.randint(0, V, (M,), device="cuda", dtype=torch.int32)
    logp, lse = ops.logprob_fwd(h, E, labels, temp)
    hf, Ef = h.float().requires_grad_(True), E.float().requires_grad_(True)
    logits = hf @ Ef.t() / temp
    ref = torch.log_softmax(logits, -1).gather(1, labels.long()[:, None]).squeeze(1)
    close(lse, torch.logsumexp(logits, -1), 1e-4, "lse")
    assert (logp - ref).abs().max().item() < 2e-3, "fp32 fused log-prob vs fp32 torch (tolerance 2e-3 abs)"
    dlogp = torch.randn(M, device="cuda")
    ref.backward(dlogp)
    dE = torch.zeros(V, H, device="cuda")
    dh = ops.logprob_bwd(dlogp, h, E, labels, lse, dE, temp)
    close(dh, hf.grad, 2 ** -5, "lm_head dh")
    close(dE, Ef.grad, 2 ** -5, "lm_head dE")


# ---------------------------------------------------------------------------------------------- optimizer
def test_adamw_matches_torch(ops):
    from iad_r1_b200 import lib as L
    torch.manual_seed(7)
    n = 100003
    p = torch.randn(n, device="cuda")
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.1)
    p32, p16 = p.clone(), p.to(bf16)
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in range(1, 4):
        g = torch.randn(n, device="cuda") * 3
        ref.grad = g.clone()
        torch.nn.utils.clip_grad_norm_([ref], 1.0)
        opt.step()
        gbuf = g.clone()
        ss = torch.zeros(1, device="cuda")
        L.check(L.lib().iadr1_sumsq_f32(gbuf.data_ptr(), n, ss.data_ptr(), L.stream_ptr()))
        L.check(L.lib().iadr1_adamw_step(p32.data_ptr(), p16.data_ptr(), gbuf.data_ptr(), m.data_ptr(), v.data_ptr(), n,
                                         1e-3, 0.9, 0.999, 1e-8, 0.1, step, 1.0, ss.data_ptr(), 1.0, 1, L.stream_ptr()))
        assert (gbuf == 0).all()
    close(p32, ref.data, 1e-5, "adamw master weights")
    assert torch.equal(p16, p32.to(bf16))


# ---------------------------------------------------------------------------------------------- sampler
def test_adamw_bf16_moments_track_fp32(ops):
    """AdamW with bf16 moments under stochastic rounding (the 7B memory plan) follows the fp32-moment kernel: after 200 steps
    of the same gradient stream the parameters agree to a small fraction of the distance travelled, and the bf16 second
    moment has not frozen (round-to-nearest would: beta2 = 0.999 moves it by a quarter of a bf16 ulp per step)."""
    from iad_r1_b200 import lib as L
    torch.manual_seed(9)
    n = 1 << 16
    p0 = torch.randn(n, device="cuda")
    st = {}
    for kind in ("fp32", "bf16"):
        dt = torch.float32 if kind == "fp32" else bf16
        st[kind] = dict(p32=p0.clone(), p16=p0.to(bf16), m=torch.zeros(n, device="cuda", dtype=dt), v=torch.zeros(n, device="cuda", dtype=dt))
    gen = torch.Generator(device="cuda").manual_seed(1)
    base = torch.randn(n, device="cuda", generator=gen) * 0.5
    for step in range(1, 201):
        g = base + torch.randn(n, device="cuda", generator=gen)
        for kind, s_ in st.items():
            gg = g.clone()
            if kind == "fp32":
                L.check(L.lib().iadr1_adamw_step(s_["p32"].data_ptr(), s_["p16"].data_ptr(), gg.data_ptr(), s_["m"].data_ptr(),
                                                 s_["v"].data_ptr(), n, 1e-3, 0.9, 0.999, 1e-8, 0.01, step, 1.0, None, 0.0, 1,
                                                 L.stream_ptr()), "adamw")
            else:
                L.check(L.lib().iadr1_adamw_step_bf16m(s_["p32"].data_ptr(), s_["p16"].data_ptr(), gg.data_ptr(), s_["m"].data_ptr(),
                                                       s_["v"].data_ptr(), n, 1e-3, 0.9, 0.999, 1e-8, 0.01, step, 1.0, None, 0.0, 1,
                                                       7, L.stream_ptr()), "adamw_bf16m")
            assert (gg == 0).all()
    torch.cuda.synchronize()
    moved = (st["fp32"]["p32"] - p0).abs().mean().item()
    diff = (st["fp32"]["p32"] - st["bf16"]["p32"]).abs()
    print(f"\nadamw bf16 moments: mean |dp| {moved:.4f}, mean |p_bf16m - p_fp32m| {diff.mean().item():.2e}, max {diff.max().item():.2e}")
    assert diff.mean().item() < 0.01 * moved and diff.max().item() < 0.1 * moved
    rel_v = (st["bf16"]["v"].float() - st["fp32"]["v"]) / st["fp32"]["v"]
    # noise at the bf16 rounding level (a random walk damped by beta2), and UNBIASED: no systematic drift of exp_avg_sq
    assert rel_v.abs().mean().item() < 0.03 and abs(rel_v.mean().item()) < 2e-3, (rel_v.abs().mean().item(), rel_v.mean().item())
    assert (st["bf16"]["p16"].float() - st["bf16"]["p32"]).abs().max().item() <= 2 ** -7 * st["bf16"]["p32"].abs().max().item()


def test_sampler_support_and_distribution(ops):
    from iad_r1_b200 import lib as L
    torch.manual_seed(8)
    V, rows, c_max = 5000, 64, 4
    base = torch.randn(V, device="cuda") * 2
    logits = base.repeat(rows, 1).contiguous()
    temp, top_k, top_p = 0.9, 50, 0.9
    # reference support: HF warper order (temperature -> top-k -> top-p)
    sc = base / temp
    kth = torch.topk(sc, top_k)[0][-1]
    sc = sc.masked_fill(sc < kth, float("-inf"))
    sp, si = torch.sort(sc, descending=False)
    cum = sp.softmax(-1).cumsum(-1)
    remove = cum <= (1 - top_p)
    remove[-1:] = False
    keep_ids = si[~remove]
    probs = torch.zeros(V, device="cuda")
    probs[keep_ids] = torch.softmax(sc[keep_ids], -1)
    counts = torch.zeros(V, device="cuda")
    n_draws = 0
    for seed in range(40):
        state = torch.tensor([0, 10, rows, 0, 0, 0, 0, 0], device="cuda", dtype=torch.int32)
        tok = torch.zeros(rows, device="cuda", dtype=torch.int32)
        fin = torch.zeros(rows, device="cuda", dtype=torch.int32)
        out = torch.full((rows, c_max), -1, device="cuda", dtype=torch.int32)
        L.check(L.lib().iadr1_sample(logits.data_ptr(), rows, V, temp, top_k, top_p, seed, state.data_ptr(), tok.data_ptr(),
                                     fin.data_ptr(), out.data_ptr(), c_max, V + 5, 0, 0, 1, L.stream_ptr()))
        assert torch.equal(out[:, 0], tok)
        assert (out[:, 1:] == -1).all()
        counts += torch.bincount(tok.long(), minlength=V).float()
        n_draws += rows
    assert (counts[probs == 0] == 0).all(), "sampled a token outside the top-k/top-p support"
    emp = counts / n_draws
    assert (emp - probs).abs().max().item() < 0.03, "empirical frequencies deviate from the warped distribution"
    # determinism for a fixed (seed, row, step)
    state = torch.tensor([0, 10, rows, 0, 0, 0, 0, 0], device="cuda", dtype=torch.int32)
    t1, t2 = torch.zeros(rows, device="cuda", dtype=torch.int32), torch.zeros(rows, device="cuda", dtype=torch.int32)
    fin = torch.zeros(rows, device="cuda", dtype=torch.int32)
    out = torch.zeros(rows, c_max, device="cuda", dtype=torch.int32)
    for t in (t1, t2):
        fin.zero_()
        L.check(L.lib().iadr1_sample(logits.data_ptr(), rows, V, temp, top_k, top_p, 123, state.data_ptr(), t.data_ptr(),
                                     fin.data_ptr(), out.data_ptr(), c_max, V + 5, 0, 0, 1, L.stream_ptr()))
    assert torch.equal(t1, t2)
    assert len(set(t1.tolist())) > 1, "rows must draw independently"


# ---------------------------------------------------------------------------------------------- decode attention
def _rot_half_f(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), -1)


@pytest.mark.parametrize("nq,nkv,hd,nsplit", [(8, 1, 128, 3), (12, 2, 128, 1), (14, 2, 128, 13), (16, 2, 128, 5),
                                              (4, 2, 64, 3), (14, 2, 64, 2), (4, 2, 32, 0),
                                              (16, 2, 128, -2), (14, 2, 128, -7), (14, 2, 64, -3), (4, 2, 64, -1)])
def test_decode_attention_fused(ops, nq, nkv, hd, nsplit):
    """The one-launch decode attention (rotary + KV append + split-KV attention + merge; tensor-core path for hd=64/128,
    scalar path otherwise) against a torch restatement: per row, keys = shared prompt prefix of its group + its own slab +
    the token being decoded."""
    from iad_r1_b200 import lib as L
    torch.manual_seed(nq * 7 + hd)
    dev = "cuda"
    R, p_max, c_max, step = 5, 200, 100, 57
    plen = torch.tensor([150, 200, 37, 37, 1], dtype=torch.int32, device=dev)
    grp = torch.tensor([0, 0, 1, 1, 2], dtype=torch.int32, device=dev)
    delta = torch.tensor([-20, -20, 0, 0, 3], dtype=torch.int32, device=dev)
    D = (nq + 2 * nkv) * hd
    qkv = torch.randn(R, D, device=dev)
    kp, vp = rnd(3, p_max, nkv, hd), rnd(3, p_max, nkv, hd)
    kc, vc = rnd(R, c_max, nkv, hd), rnd(R, c_max, nkv, hd)
    kc0, vc0 = kc.clone(), vc.clone()
    max_pos = p_max + c_max + 8
    ang = torch.arange(max_pos, device=dev, dtype=torch.float32)[:, None] * (1.0 / (1e6 ** (torch.arange(0, hd, 2, device=dev).float() / hd)))
    emb = torch.cat((ang, ang), -1)
    cos_t, sin_t = emb.cos().contiguous(), emb.sin().contiguous()
    state = torch.tensor([step, 0, R, 0, 0, 0, 0, 0], dtype=torch.int32, device=dev)
    nsplit = nsplit or (p_max + c_max + 127) // 128     # tensor-core path: any split count (negative: the 2-warp
    part = torch.zeros(R, nq, abs(nsplit), hd + 2, device=dev)   # variant); scalar path: 128-key chunks
    tickets = torch.zeros(R * nkv, dtype=torch.int32, device=dev)
    out = torch.zeros(R, nq * hd, dtype=bf16, device=dev)
    scale = hd ** -0.5
    for _ in range(2):  # second call checks the ticket re-arm
        kc.copy_(kc0); vc.copy_(vc0)
        L.check(L.lib().iadr1_decode_attention_fused(
            qkv.data_ptr(), cos_t.data_ptr(), sin_t.data_ptr(), delta.data_ptr(), kp.data_ptr(), vp.data_ptr(), kc.data_ptr(),
            vc.data_ptr(), state.data_ptr(), grp.data_ptr(), plen.data_ptr(), None, part.data_ptr(), tickets.data_ptr(), out.data_ptr(),
            R, nq, nkv, hd, p_max, c_max, nsplit, max_pos, scale, L.stream_ptr()))
    torch.cuda.synchronize()
    assert (tickets == 0).all()
    g = nq // nkv
    for r in range(R):
        P = int(plen[r]); pos = P + step + int(delta[r])
        c, s_ = cos_t[pos].to(bf16), sin_t[pos].to(bf16)
        x = qkv[r].to(bf16).view(nq + 2 * nkv, hd)
        qk = (x[:nq + nkv] * c) + (_rot_half_f(x[:nq + nkv]) * s_)          # bf16 arithmetic, as HF
        q, knew, vnew = qk[:nq].float(), qk[nq:], x[nq + nkv:]
        assert torch.equal(kc[r, step], knew) and torch.equal(vc[r, step], vnew), "KV append"
        assert torch.equal(kc[r, :step], kc0[r, :step]) and torch.equal(kc[r, step + 1:], kc0[r, step + 1:])
        K = torch.cat([kp[grp[r], :P], kc0[r, :step], knew[None]], 0).float()   # [ctx, nkv, hd]
        V = torch.cat([vp[grp[r], :P], vc0[r, :step], vnew[None]], 0).float()
        for h in range(nq):
            kvh = h // g
            p_ = torch.softmax((K[:, kvh] @ q[h]) * scale, 0)
            ref = p_ @ V[:, kvh]
            close(out[r].view(nq, hd)[h], ref, 2 ** -6, f"decode attention row {r} head {h}")


@pytest.mark.parametrize("nq,nkv,hd,G,n_groups,psplit,csplit,step", [
    (16, 2, 128, 8, 2, 3, 1, 57), (16, 2, 128, 8, 2, 1, 2, 0), (14, 2, 128, 16, 1, 4, 2, 90), (12, 2, 128, 4, 3, 2, 1, 33),
    (14, 2, 64, 8, 2, 2, 2, 64), (4, 2, 64, 3, 2, 5, 3, 17), (8, 1, 128, 1, 4, 2, 1, 5)])
def test_decode_attention_grouped(ops, nq, nkv, hd, G, n_groups, psplit, csplit, step):
    """Shared-prefix decode attention (prompt keys once per group in full m16 tiles of 2 rows x 8 heads, each row's own keys
    per row, one launch, ticket merge over psplit + csplit slots) against the same torch restatement as the per-row kernel:
    keys of a row = the prompt prefix of its group + its own slab + the token being decoded (rotary + KV append)."""
    from iad_r1_b200 import lib as L
    torch.manual_seed(nq * 7 + hd + G)
    dev = "cuda"
    R, p_max, c_max = G * n_groups, 200, 100
    plens = [150, 200, 37, 1][:n_groups] if n_groups <= 4 else None
    plen = torch.tensor([plens[r // G] for r in range(R)], dtype=torch.int32, device=dev)
    delta = torch.tensor([(-20, 0, 3, 7)[r // G] for r in range(R)], dtype=torch.int32, device=dev)
    D = (nq + 2 * nkv) * hd
    qkv = torch.randn(R, D, device=dev)
    kp, vp = rnd(n_groups, p_max, nkv, hd), rnd(n_groups, p_max, nkv, hd)
    kc, vc = rnd(R, c_max, nkv, hd), rnd(R, c_max, nkv, hd)
    kc0, vc0 = kc.clone(), vc.clone()
    max_pos = p_max + c_max + 8
    ang = torch.arange(max_pos, device=dev, dtype=torch.float32)[:, None] * (1.0 / (1e6 ** (torch.arange(0, hd, 2, device=dev).float() / hd)))
    emb = torch.cat((ang, ang), -1)
    cos_t, sin_t = emb.cos().contiguous(), emb.sin().contiguous()
    state = torch.tensor([step, 0, R, 0, 0, 0, 0, 0], dtype=torch.int32, device=dev)
    part = torch.zeros(R, nq, psplit + csplit, hd + 2, device=dev)
    tickets = torch.zeros(R * nkv, dtype=torch.int32, device=dev)
    out = torch.zeros(R, nq * hd, dtype=bf16, device=dev)
    scale = hd ** -0.5
    fin = torch.zeros(R, dtype=torch.int32, device=dev)       # rows that already produced EOS are skipped entirely
    if R > 2:
        fin[1] = 1
    for _ in range(2):  # second call checks the ticket re-arm
        kc.copy_(kc0); vc.copy_(vc0)
        L.check(L.lib().iadr1_decode_attention_grouped(
            qkv.data_ptr(), cos_t.data_ptr(), sin_t.data_ptr(), delta.data_ptr(), kp.data_ptr(), vp.data_ptr(), kc.data_ptr(),
            vc.data_ptr(), state.data_ptr(), plen.data_ptr(), fin.data_ptr(), part.data_ptr(), tickets.data_ptr(), out.data_ptr(),
            R, G, nq, nkv, hd, p_max, c_max, psplit, csplit, max_pos, scale, L.stream_ptr()))
    torch.cuda.synchronize()
    assert (tickets == 0).all()
    g = nq // nkv
    for r in range(R):
        if int(fin[r]):
            assert torch.equal(kc[r], kc0[r]) and torch.equal(vc[r], vc0[r]) and (out[r] == 0).all(), "finished row must be untouched"
            continue
        P = int(plen[r]); pos = P + step + int(delta[r])
        c, s_ = cos_t[pos].to(bf16), sin_t[pos].to(bf16)
        x = qkv[r].to(bf16).view(nq + 2 * nkv, hd)
        qk = (x[:nq + nkv] * c) + (_rot_half_f(x[:nq + nkv]) * s_)          # bf16 arithmetic, as HF
        q, knew, vnew = qk[:nq].float(), qk[nq:], x[nq + nkv:]
        assert torch.equal(kc[r, step], knew) and torch.equal(vc[r, step], vnew), "KV append"
        assert torch.equal(kc[r, :step], kc0[r, :step]) and torch.equal(kc[r, step + 1:], kc0[r, step + 1:])
        K = torch.cat([kp[r // G, :P], kc0[r, :step], knew[None]], 0).float()   # [ctx, nkv, hd]
        V = torch.cat([vp[r // G, :P], vc0[r, :step], vnew[None]], 0).float()
        for h in range(nq):
            kvh = h // g
            p_ = torch.softmax((K[:, kvh] @ q[h]) * scale, 0)
            ref = p_ @ V[:, kvh]
            close(out[r].view(nq, hd)[h], ref, 2 ** -6, f"grouped decode attention row {r} head {h}")


# ---------------------------------------------------------------------------------------------- image preprocessing
@pytest.mark.parametrize("hw,max_pixels", [((448, 448), 480000), ((300, 500), 480000), ((1000, 700), 480000), ((90, 120), 12845056)])
def test_image_preprocess_matches_hf_processor(ops, hw, max_pixels):
    """GPU resize + normalise + patchify (csrc/preprocess.cu) against HF `Qwen2VLImageProcessor` (PIL bicubic) on the same
    uint8 image: same grid, identical values when no resize is needed (up to the bf16 store), within 2 grey levels
    (2 / 255 / std ~ 0.03) when Pillow's fixed-point resampling rounds differently from the fp32 kernel."""
    import numpy as np
    from PIL import Image
    from transformers import Qwen2VLImageProcessor
    from iad_r1_b200.config import tiny_config
    from iad_r1_b200.preprocess import qwen_preprocess_gpu, smart_resize
    rng = np.random.RandomState(hw[0])
    # smooth + noisy content: low-frequency gradient plus noise exercises the antialiasing window
    yy, xx = np.mgrid[0:hw[0], 0:hw[1]]
    img = (127 + 80 * np.sin(yy / 23.0)[..., None] * np.cos(xx / 17.0)[..., None] + rng.randint(-40, 40, (hw[0], hw[1], 3))).clip(0, 255).astype(np.uint8)
    v = tiny_config("qwen2_5_vl").vision
    hf = Qwen2VLImageProcessor(min_pixels=3136, max_pixels=max_pixels, patch_size=v.patch_size,
                               temporal_patch_size=v.temporal_patch_size, merge_size=v.spatial_merge_size)
    ref = hf(images=[Image.fromarray(img)], return_tensors="pt")
    pv, grid = qwen_preprocess_gpu([Image.fromarray(img)], v, torch.device("cuda"), 3136, max_pixels)
    assert [list(g) for g in ref["image_grid_thw"].tolist()] == grid
    th, tw = smart_resize(hw[0], hw[1], 28, 3136, max_pixels)
    assert (grid[0][1] * 14, grid[0][2] * 14) == (th, tw)
    want = ref["pixel_values"].float().cuda()
    got = pv.float()
    assert got.shape == want.shape
    err = (got - want).abs()
    if (th, tw) == hw:
        assert err.max().item() <= 2 ** -7 * want.abs().max().item()
    else:
        assert err.max().item() <= 0.035 and err.mean().item() <= 0.004, (err.max().item(), err.mean().item())


@pytest.mark.parametrize("M,I,K,keep", [(300, 256, 192, True), (1000, 1408, 256, False), (129, 128, 64, True)])
def test_gemm_swiglu_training_epilogue(ops, M, I, K, keep):
    """Training MLP front half in one launch (epi 4): the gate_up product with the SwiGLU epilogue against the two-kernel form
    (GEMM -> act_mul): bit-identical act and gate | up pre-activations (same bf16 rounding points)."""
    from iad_r1_b200 import lib as L
    torch.manual_seed(M + I)
    x = rnd(M, K)
    w = rnd(2 * I, K, scale=K ** -0.5)
    gu_ref = ops.linear_fwd(x, w)
    act_ref = ops.act_mul_fwd(gu_ref, I, ops.ACT_SILU, gated=True)
    act = torch.full((M, I), float("nan"), dtype=bf16, device="cuda")
    gu = torch.full((M, 2 * I), float("nan"), dtype=bf16, device="cuda") if keep else None
    L.gemm_swiglu_train(x, w, act, gu)
    torch.cuda.synchronize()
    assert torch.equal(act, act_ref)
    if keep:
        assert torch.equal(gu, gu_ref)


@pytest.mark.parametrize("M,I,H", [(300, 384, 256), (1111, 3424, 1280), (1000, 11008, 2048), (64, 520, 128)])
def test_gemm_swiglu_backward_epilogue(ops, M, I, H):
    """epi 5: the down-projection input gradient with the SwiGLU backward in the epilogue rewrites gate | up with dgate | dup
    bit-identically to GEMM (bf16 dact) -> act_mul_bwd (what it replaces in iadr1_decoder_bwd), and matches fp32 autograd."""
    from iad_r1_b200 import lib as L
    torch.manual_seed(M + I)
    dy, w_down, gu = rnd(M, H, scale=0.5), rnd(H, I, scale=0.05), rnd(M, 2 * I)
    want = gu.clone()
    dact = L.gemm(dy, w_down.t())                                   # dy @ w_down, bf16
    ops.act_mul_bwd(dact, want, I, 0, True, dgu=want)
    got = gu.clone()
    L.gemm_swiglu_bwd(dy, w_down, got)
    assert torch.equal(got, want), f"max diff {(got.float() - want.float()).abs().max().item():.4g}"
    g32 = gu[:, :I].float().requires_grad_()
    u32 = gu[:, I:].float().requires_grad_()
    (torch.nn.functional.silu(g32) * u32).backward(dy.float() @ w_down.float())
    close(got[:, :I], g32.grad, 2 ** -6, "dgate")
    close(got[:, I:], u32.grad, 2 ** -6, "dup")
