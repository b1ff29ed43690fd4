"""Times the fused attention kernels on the benchmark geometries (CUDA events, L2 flushed between iterations by the
working set of several distinct buffers) and prints algorithmic TFLOP/s: forward 4 * pairs * hd per head, backward
10 * pairs * hd per head (5 products; the two-kernel backward executes 7).   python tools/fmha_probe.py [--ncu]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iad_r1_b200 import fmha  # noqa: E402

dev = torch.device("cuda:0")
bf16 = torch.bfloat16


def case(name):
    if name == "3b_2groups":      # Qwen2.5-VL-3B scoring pass: two packed groups, P = 297, G = 8, C = 512
        rs, ps, ss, base = [], [], [], 0
        for _ in range(2):
            r, p, s = fmha.shared_prefix_geometry(297, 8, 512, base)
            rs.append(r); ps += p; ss += s
            base += 297 + 8 * 512
        return np.concatenate(rs), ps, ss, 16, 2, 128
    if name == "3b_prefill":      # rollout prefill: 16 prompts of 297 tokens, causal
        B, T = 16, 297
        seq = [(b * T, (b + 1) * T) for b in range(B)]
        return fmha.causal_rows(B, T), seq, seq, 16, 2, 128
    if name == "vit_full":        # 16 images x 1024 patches, full attention inside each image, hd = 80
        N = 16 * 1024
        lo = np.arange(N) // 1024 * 1024
        return fmha.range_rows(lo, lo + 1024), None, None, 16, 16, 80
    if name == "vit_win":         # 64-patch windows
        N = 16 * 1024
        lo = np.arange(N) // 64 * 64
        return fmha.range_rows(lo, lo + 64), None, None, 16, 16, 80
    if name == "siglip":          # LLaVA-OneVision: 2 images x 5 crops x 729 tokens, hd = 72
        N = 10 * 729
        lo = np.arange(N) // 729 * 729
        return fmha.range_rows(lo, lo + 729), None, None, 16, 16, 72
    if name == "ov_2groups":      # LLaVA-OV-0.5B scoring pass: P = 3738, G = 8, C = 512, hd = 64, 14:2 heads
        rs, ps, ss, base = [], [], [], 0
        for _ in range(2):
            r, p, s = fmha.shared_prefix_geometry(3738, 8, 512, base)
            rs.append(r); ps += p; ss += s
            base += 3738 + 8 * 512
        return np.concatenate(rs), ps, ss, 14, 2, 64
    raise KeyError(name)


def main():
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["3b_2groups", "3b_prefill", "vit_full", "vit_win", "siglip", "ov_2groups"]
    iters = 3 if "--ncu" in sys.argv else 10
    for name in names:
        rng, probs, segs, nq, nkv, hd = case(name)
        N = rng.shape[0]
        plan = fmha.FmhaPlan(rng, dev, probs, segs, nkv=nkv, nq=nq)
        D = (nq + 2 * nkv) * hd
        nbuf = 4
        qkvs = [(torch.randn(N, D, device=dev) * 0.5).to(bf16) for _ in range(nbuf)]
        douts = [torch.randn(N, nq * hd, device=dev).to(bf16) for _ in range(nbuf)]
        scale = hd ** -0.5
        outs = [fmha.fmha_fwd(q, plan, nq, nkv, hd, scale) for q in qkvs]
        for q, d, (o, l) in zip(qkvs, douts, outs):
            fmha.fmha_bwd(d, q, o, l, plan, nq, nkv, hd, scale)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        for i in range(iters):
            fmha.fmha_fwd(qkvs[i % nbuf], plan, nq, nkv, hd, scale)
        e[1].record()
        for i in range(iters):
            o, l = outs[i % nbuf]
            fmha.fmha_bwd(douts[i % nbuf], qkvs[i % nbuf], o, l, plan, nq, nkv, hd, scale)
        e[2].record()
        torch.cuda.synchronize()
        t_f, t_b = e[0].elapsed_time(e[1]) / iters, e[1].elapsed_time(e[2]) / iters
        fl_f = 4.0 * plan.pairs * hd * nq
        print(f"{name}: N={N} nq={nq} nkv={nkv} hd={hd} q_items={plan.n_q} k_items={plan.n_k} pairs={plan.pairs / 1e6:.2f}M | "
              f"fwd {t_f * 1e3:.1f} us ({fl_f / t_f / 1e9:.0f} TFLOP/s) | bwd {t_b * 1e3:.1f} us ({2.5 * fl_f / t_b / 1e9:.0f} TFLOP/s algorithmic)",
              flush=True)


def trace():
    """Per-iteration timeline of CTA 0 of the dK/dV kernel (clock64 stamps; cycles relative to the first stamp)."""
    from iad_r1_b200 import lib as L
    rng, probs, segs, nq, nkv, hd = case("3b_2groups")
    N = rng.shape[0]
    plan = fmha.FmhaPlan(rng, dev, probs, segs, nkv=nkv, nq=nq)
    D = (nq + 2 * nkv) * hd
    qkv = (torch.randn(N, D, device=dev) * 0.5).to(bf16)
    dout = torch.randn(N, nq * hd, device=dev).to(bf16)
    scale = hd ** -0.5
    o, l = fmha.fmha_fwd(qkv, plan, nq, nkv, hd, scale)
    fmha.fmha_bwd(dout, qkv, o, l, plan, nq, nkv, hd, scale)
    buf = torch.zeros(3 * 64 * 8, dtype=torch.int64, device=dev)
    L.lib().iadr1_fmha_set_trace(buf.data_ptr())
    fmha.fmha_bwd(dout, qkv, o, l, plan, nq, nkv, hd, scale)
    torch.cuda.synchronize()
    L.lib().iadr1_fmha_set_trace(None)
    t = buf.cpu().view(3, 64, 8)
    t0 = int(t[t > 0].min())
    rel = torch.where(t > 0, t - t0, torch.zeros_like(t))
    print("iter | producer: wait_empty got_empty | mma: top qd_full st_free issued_ST | kv: wait_pds got_pds issued_KV | "
          "compute: top got_ST ld_done compute_done got_pds_free arrived")
    for i in range(40):
        p_, m_, c_ = rel[0, i].tolist(), rel[1, i].tolist(), rel[2, i].tolist()
        print(f"{i:3d} | {p_[0]:7d} {p_[1]:7d} | {m_[0]:7d} {m_[1]:7d} {m_[2]:7d} {m_[3]:7d} | {m_[4]:7d} {m_[5]:7d} {m_[6]:7d} | "
              f"{c_[0]:7d} {c_[1]:7d} {c_[2]:7d} {c_[3]:7d} {c_[4]:7d} {c_[5]:7d}")


if __name__ == "__main__":
    if "--trace" in sys.argv:
        trace()
    else:
        main()
