"""GPU parity of the fused attention kernels (csrc/fmha_sm100.cu: tcgen05 forward, dQ and dK/dV backward) against dense
fp32 torch attention with the same per-row key ranges, through the C ABI (iadr1_fmha_fwd / iadr1_fmha_bwd).

Layouts: causal batches, the shared-prefix GRPO layout (prompt once + G completion rows, rows straddling 128-row tiles),
two groups packed in one stream, vision windows (head_dim 80), SigLIP crops (head_dim 72, 729 tokens), GQA 8:1 / 7:1,
the tiny twins' head dims (32 / 16 / 24). Tolerances: outputs and gradients are bf16; compared at 2^-7 (forward) / 2^-5
(backward, as the composed-attention test) of the reference magnitude; the log-sum-exp is fp32, 2e-3 absolute."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def _close(got, want, tol, what):
    got, want = got.float(), want.float()
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    err = (got - want).abs().max().item()
    ref = want.abs().max().item() + 1e-6
    assert err <= tol * ref, f"{what}: max abs err {err:.4g} vs ref max {ref:.4g} (tol {tol:.3g} rel)"
    return err / ref


def _layout(name):
    from iad_r1_b200 import fmha
    if name == "causal":
        B, T = 2, 200
        return fmha.causal_rows(B, T), [(b * T, (b + 1) * T) for b in range(B)], [(b * T, (b + 1) * T) for b in range(B)]
    if name == "causal_long":
        return fmha.causal_rows(1, 700), None, None
    if name == "shared":
        return fmha.shared_prefix_geometry(297, 3, 150)
    if name == "shared_small":
        return fmha.shared_prefix_geometry(29, 4, 12)
    if name == "two_groups":
        r1, p1, s1 = fmha.shared_prefix_geometry(140, 2, 70, 0)
        n1 = 140 + 2 * 70
        r2, p2, s2 = fmha.shared_prefix_geometry(90, 2, 130, n1)
        return np.concatenate([r1, r2]), p1 + p2, s1 + s2
    if name == "windows":     # 64-token vision windows, one ragged window at the end
        N = 512 + 40
        lo = np.arange(N) // 64 * 64
        hi = np.minimum(lo + 64, N)
        return fmha.range_rows(lo, hi), None, None
    if name == "crops":       # SigLIP: full attention inside crops of 729 tokens
        N = 2 * 729
        lo = np.arange(N) // 729 * 729
        return fmha.range_rows(lo, lo + 729), None, None
    if name == "full_image":
        N = 1024
        return fmha.range_rows(np.zeros(N, dtype=np.int64), np.full(N, N)), None, None
    raise KeyError(name)


CASES = [
    ("causal", 4, 2, 128), ("causal_long", 2, 1, 128), ("shared", 8, 1, 128), ("shared", 16, 2, 128), ("shared_small", 4, 2, 32),
    ("two_groups", 14, 2, 64), ("windows", 4, 4, 80), ("crops", 4, 4, 72), ("full_image", 2, 2, 80), ("shared_small", 4, 4, 16),
    ("windows", 4, 4, 24), ("causal", 6, 2, 64),
]


@pytest.mark.parametrize("layout,nq,nkv,hd", CASES)
def test_fmha_fwd_bwd(cuda, layout, nq, nkv, hd):
    from iad_r1_b200 import fmha
    rng, probs, segs = _layout(layout)
    N = rng.shape[0]
    plan = fmha.FmhaPlan(rng, cuda, probs, segs, nkv=nkv, nq=nq)
    torch.manual_seed(N + hd)
    D = (nq + 2 * nkv) * hd
    qkv = (torch.randn(N, D, device=cuda) * 0.7).to(bf16)
    dout = torch.randn(N, nq * hd, device=cuda).to(bf16)
    scale = hd ** -0.5
    out, lse2 = fmha.fmha_fwd(qkv, plan, nq, nkv, hd, scale)
    dqkv = fmha.fmha_bwd(dout, qkv, out, lse2, plan, nq, nkv, hd, scale)
    torch.cuda.synchronize()
    # dense fp32 reference
    g = nq // nkv
    x = qkv.float().view(N, nq + 2 * nkv, hd).requires_grad_(True)
    q, k, v = x[:, :nq], x[:, nq:nq + nkv], x[:, nq + nkv:]
    kk, vv = k.repeat_interleave(g, 1), v.repeat_interleave(g, 1)
    s = torch.einsum("qhd,khd->hqk", q, kk) * scale
    mask = torch.from_numpy(fmha.reference_mask(rng)).to(cuda)
    s = s.masked_fill(~mask, float("-inf"))
    p = torch.softmax(s, -1)
    ref = torch.einsum("hqk,khd->qhd", p, vv).reshape(N, nq * hd)
    ref.backward(dout.float())
    e_out = _close(out, ref, 2 ** -7, f"{layout} out")
    lse_ref = torch.logsumexp(s, -1).t()                               # [N, nq], natural log
    lse_err = (lse2[:, :N].t() * math.log(2.0) - lse_ref).abs().max().item()
    assert lse_err < 2e-3, f"{layout}: lse err {lse_err}"
    gref = x.grad.reshape(N, D)
    QH, KH = nq * hd, nkv * hd
    e_q = _close(dqkv[:, :QH], gref[:, :QH], 2 ** -5, f"{layout} dQ")
    e_k = _close(dqkv[:, QH:QH + KH], gref[:, QH:QH + KH], 2 ** -5, f"{layout} dK")
    e_v = _close(dqkv[:, QH + KH:], gref[:, QH + KH:], 2 ** -5, f"{layout} dV")
    print(f"\n[{layout} nq={nq} nkv={nkv} hd={hd} N={N}] items q={plan.n_q} k={plan.n_k}: rel err out {e_out:.2e} "
          f"dQ {e_q:.2e} dK {e_k:.2e} dV {e_v:.2e} lse {lse_err:.1e}")
