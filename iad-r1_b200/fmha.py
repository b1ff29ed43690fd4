"""Host side of the fused attention kernels (csrc/fmha_sm100.cu): per-row key ranges and the work lists the kernels walk.

Every token layout the trainers use is described the same way: query row t attends keys [lo, hi) u [plo, phi) (token
indices into the same fused qkv buffer). Causal sequences, the shared-prefix GRPO layout (prompt once + G completion rows,
ref: sc_grpo_trainer.py:624-628 tiles the prompt G times instead), vision windows / full images / SigLIP crops and several
groups packed into one pass are all instances. The plan (device-resident int32 tables) is built once per layout.

Replaces the flash-attn varlen call sites of the HF modules the reference runs (`--attn_implementation flash_attention_2`,
ref: scripts/train/SC_GRPO/SC_GRPO_Qwen_Instruct_2_5_VL_3B.sh:58)."""
from __future__ import annotations

import os

import numpy as np
import torch

from . import lib as L

bf16, f32 = torch.bfloat16, torch.float32
NUM_SMS = 148


def supported(hd: int) -> bool:
    return hd % 8 == 0 and 8 <= hd <= 128 and os.environ.get("IADR1_ATTN", "fused") != "composed"


def build_items(rng: np.ndarray, q_problems=None, k_segments=None, nkv: int = 1, n_sms: int = NUM_SMS):
    """rng int32 [N, 4] = (lo, hi, plo, phi) per query row. Returns (q_items [n, 6], k_items [m, 4]) int32.

    q_items: 128-row query tiles {q0, nrows, kv0, kv1, p0, p1} (tiles never straddle two `q_problems`), heaviest first.
    k_items: {k0, nkeys, q0, q1, f0, f1}: 128-key tiles of every key segment with the contiguous range of query rows that may
    attend to them (q0 rounded down to a multiple of 4), long query ranges split so that no unit dominates the backward's
    critical path, heaviest first; 64-query tiles f0 <= t < f1 of the range are allowed in full."""
    N = rng.shape[0]
    lo, hi, plo, phi = (rng[:, i].astype(np.int64) for i in range(4))
    has1, has2 = hi > lo, phi > plo
    if not np.all(has1 | has2):
        raise ValueError("fmha: every query row needs at least one key")
    if np.any(has1 & has2 & (np.maximum(lo, plo) < np.minimum(hi, phi))):
        raise ValueError("fmha: the two key ranges of a row must be disjoint")
    q_problems = [(0, N)] if q_problems is None else q_problems
    k_segments = [(0, N)] if k_segments is None else k_segments
    big = np.iinfo(np.int64).max
    q_items = []
    for qs, qe in q_problems:
        for q0 in range(qs, qe, 128):
            q1 = min(q0 + 128, qe)
            m1, m2 = has1[q0:q1], has2[q0:q1]
            kv0 = int(np.where(m1, lo[q0:q1], big).min()) if m1.any() else 0
            kv1 = int(np.where(m1, hi[q0:q1], 0).max()) if m1.any() else 0
            p0 = int(np.where(m2, plo[q0:q1], big).min()) if m2.any() else 0
            p1 = int(np.where(m2, phi[q0:q1], 0).max()) if m2.any() else 0
            cost = (kv1 - kv0 + 127) // 128 + (p1 - p0 + 127) // 128
            q_items.append((cost, q0, q1 - q0, kv0, kv1, p0, p1))
    q_items.sort(key=lambda x: -x[0])
    # key tiles -> attending query range (vectorised over rows)
    tiles = []
    for ks, ke in k_segments:
        for k0 in range(ks, ke, 128):
            k1 = min(k0 + 128, ke)
            att = (has1 & (lo < k1) & (hi > k0)) | (has2 & (plo < k1) & (phi > k0))
            idx = np.flatnonzero(att)
            if len(idx):
                tiles.append((k0, k1 - k0, int(idx[0]), int(idx[-1]) + 1))
    tiles = [(k0, nk, q0 & ~3, q1) for k0, nk, q0, q1 in tiles]     # q0 % 4 == 0: 16-byte aligned 64-query vector copies
    total_qt = sum((q1 - q0 + 63) // 64 for _, _, q0, q1 in tiles)
    max_qt = max(2, -(-total_qt * max(1, nkv) // max(1, n_sms)))       # 64-query tiles per unit
    k_items = []
    for k0, nk, q0, q1 in tiles:
        step = max_qt * 64
        for a in range(q0, q1, step):
            b = min(a + step, q1)
            nqt = (b - a + 63) // 64
            # 64-query tiles whose every (query, key) pair is allowed: the kernel skips the mask arithmetic there
            f0 = f1 = 0
            if nk == 128:
                qq = np.arange(a, a + nqt * 64)
                inb = qq < b
                qc = np.minimum(qq, N - 1)
                ok = inb & (((lo[qc] <= k0) & (hi[qc] >= k0 + nk) & has1[qc]) | ((plo[qc] <= k0) & (phi[qc] >= k0 + nk) & has2[qc]))
                full = ok.reshape(nqt, 64).all(1)
                best, cur = (0, 0), None
                for t in range(nqt + 1):          # longest run of full tiles
                    if t < nqt and full[t]:
                        cur = t if cur is None else cur
                    elif cur is not None:
                        if t - cur > best[1] - best[0]:
                            best = (cur, t)
                        cur = None
                f0, f1 = best
            k_items.append((nqt, k0, nk, a, b, f0, f1))
    k_items.sort(key=lambda x: -x[0])
    qi = np.array([x[1:] for x in q_items], dtype=np.int32).reshape(-1, 6)
    ki = np.array([x[1:] for x in k_items], dtype=np.int32).reshape(-1, 6)
    return qi, ki


def lpt_schedule(costs: np.ndarray, n_cta: int = NUM_SMS) -> np.ndarray:
    """Longest-processing-time assignment of units (cost[u] known on the host) to persistent CTAs.
    Returns int32 [n + 1 + n_units]: n = number of CTAs used, offsets [n + 1], then the unit ids CTA by CTA."""
    import heapq
    n_units = len(costs)
    n = max(1, min(n_cta, n_units))
    order = np.argsort(-np.asarray(costs, dtype=np.int64), kind="stable")
    heap = [(0, c) for c in range(n)]
    lists = [[] for _ in range(n)]
    for u in order:
        load, c = heapq.heappop(heap)
        lists[c].append(int(u))
        heapq.heappush(heap, (load + int(costs[u]), c))
    offs = np.zeros(n + 1, dtype=np.int32)
    offs[1:] = np.cumsum([len(x) for x in lists])
    return np.concatenate([offs, np.array([u for x in lists for u in x], dtype=np.int32)]).astype(np.int32)


class FmhaPlan:
    """Device-resident tables of one token layout for a given head configuration (nq query heads, nkv kv heads)."""

    def __init__(self, rng: np.ndarray, device, q_problems=None, k_segments=None, nkv: int = 1, nq: int | None = None):
        rng = np.ascontiguousarray(rng, dtype=np.int32)
        self.n_tokens = N = rng.shape[0]
        nq = nkv if nq is None else nq
        self.nq, self.nkv = nq, nkv
        qi, ki = build_items(rng, q_problems, k_segments, nkv)
        self.npad = (N + 3) // 4 * 4 + 64      # row stride of the head-major lse2 / delta; the range table is padded alike
        self.rng = torch.from_numpy(np.concatenate([rng, np.zeros((self.npad - N, 4), dtype=np.int32)])).to(device)
        self.q_items = torch.from_numpy(qi).to(device)
        self.k_items = torch.from_numpy(ki).to(device)
        self.n_q, self.n_k = qi.shape[0], ki.shape[0]
        # CTA schedules: unit = item * heads + head; cost = inner iterations + a fixed per-unit overhead
        span = lambda a, b, t: np.where(b > a, (b - a + t - 1) // t, 0)
        c_fwd = span(qi[:, 2], qi[:, 3], 128) + span(qi[:, 4], qi[:, 5], 128) + 1
        c_dq = span(qi[:, 2], qi[:, 3], 64) + span(qi[:, 4], qi[:, 5], 64) + 2
        nqt = (ki[:, 3] - ki[:, 2] + 63) // 64
        c_kv = (2 * nqt - (ki[:, 5] - ki[:, 4])) * (nq // nkv) + 6      # masked tiles cost about twice a full one
        s_fwd, s_dq, s_kv = lpt_schedule(np.repeat(c_fwd, nq)), lpt_schedule(np.repeat(c_dq, nq)), lpt_schedule(np.repeat(c_kv, nkv))
        self.n_cta_fwd, self.n_cta_dq, self.n_cta_kv = (int(min(NUM_SMS, len(qi) * nq)), int(min(NUM_SMS, len(qi) * nq)),
                                                        int(min(NUM_SMS, len(ki) * nkv)))
        self.sched_fwd = torch.from_numpy(s_fwd).to(device)
        self.sched_dq = torch.from_numpy(s_dq).to(device)
        self.sched_kv = torch.from_numpy(s_kv).to(device)
        # algorithmic FLOPs of one forward per (head, head_dim): 4 * (#unmasked score entries) * hd
        r = rng.astype(np.int64)
        self.pairs = int(np.maximum(r[:, 1] - r[:, 0], 0).sum() + np.maximum(r[:, 3] - r[:, 2], 0).sum())


# ---- range tables of the layouts -------------------------------------------------------------------------------------
def causal_rows(B: int, T: int, base: int = 0) -> np.ndarray:
    t = np.arange(B * T)
    start = t // T * T
    z = np.zeros_like(t)
    return np.stack([start + base, t + 1 + base, z, z], 1).astype(np.int32)


def shared_prefix_rows(P: int, G: int, C: int, base: int = 0) -> np.ndarray:
    """[prompt (P) | completion row 0 (C) | ... | row G-1 (C)]: prompt rows causal; completion token c of row g attends the
    whole prompt plus tokens <= c of its own row."""
    rp = causal_rows(1, P, base)
    c = np.arange(G * C)
    row_start = P + c // C * C
    rc = np.stack([row_start + base, P + c + 1 + base, np.full_like(c, base), np.full_like(c, base + P)], 1)
    return np.concatenate([rp, rc.astype(np.int32)], 0)


def shared_prefix_geometry(P: int, G: int, C: int, base: int = 0):
    """(rng, q_problems, k_segments) of one group at token offset `base`."""
    segs = [(base, base + P)] + [(base + P + g * C, base + P + (g + 1) * C) for g in range(G)]
    probs = [(base, base + P), (base + P, base + P + G * C)]
    return shared_prefix_rows(P, G, C, base), probs, segs


def range_rows(lo: np.ndarray, hi: np.ndarray) -> np.ndarray:
    z = np.zeros_like(lo)
    return np.stack([lo, hi, z, z], 1).astype(np.int32)


# ---- launches ----------------------------------------------------------------------------------------------------------
_PV_N = int(os.environ.get("IADR1_FMHA_PVN", "0"))


def fmha_fwd(qkv: torch.Tensor, plan: FmhaPlan, nq: int, nkv: int, hd: int, scale: float, out: torch.Tensor | None = None):
    """qkv bf16 [N, (nq + 2 nkv) * hd] (contiguous) -> (out bf16 [N, nq * hd], lse2 fp32 [nq, npad] head-major)."""
    N = plan.n_tokens
    if qkv.shape != (N, (nq + 2 * nkv) * hd) or qkv.dtype != bf16 or not qkv.is_contiguous():
        raise ValueError(f"fmha_fwd: qkv must be contiguous bf16 [{N}, {(nq + 2 * nkv) * hd}], got {tuple(qkv.shape)} {qkv.dtype}")
    if out is None:
        out = torch.empty(N, nq * hd, dtype=bf16, device=qkv.device)
    elif out.shape != (N, nq * hd) or not out.is_contiguous():
        raise ValueError("fmha_fwd: out must be contiguous [N, nq * hd]")
    lse2 = torch.zeros(nq, plan.npad, dtype=f32, device=qkv.device)
    if (plan.nq, plan.nkv) != (nq, nkv):
        raise ValueError(f"fmha: plan built for {plan.nq}/{plan.nkv} heads, called with {nq}/{nkv}")
    L.check(L.lib().iadr1_fmha_fwd(qkv.data_ptr(), N, nq, nkv, hd, plan.rng.data_ptr(), plan.q_items.data_ptr(), plan.n_q,
                                   plan.sched_fwd.data_ptr(), plan.n_cta_fwd, out.data_ptr(), lse2.data_ptr(), plan.npad, scale, _PV_N,
                                   L.stream_ptr()), "fmha_fwd")
    return out, lse2


def fmha_bwd(dout: torch.Tensor, qkv: torch.Tensor, out: torch.Tensor, lse2: torch.Tensor, plan: FmhaPlan, nq: int, nkv: int,
             hd: int, scale: float, dqkv: torch.Tensor | None = None):
    N = plan.n_tokens
    D = (nq + 2 * nkv) * hd
    if dout.shape != (N, nq * hd) or not dout.is_contiguous() or not out.is_contiguous() or not qkv.is_contiguous():
        raise ValueError("fmha_bwd: dout / out / qkv must be contiguous")
    if dqkv is None:
        dqkv = torch.empty(N, D, dtype=bf16, device=qkv.device)
    elif dqkv.shape != (N, D) or not dqkv.is_contiguous():
        raise ValueError("fmha_bwd: dqkv must be contiguous [N, D]")
    delta = torch.zeros(nq, plan.npad, dtype=f32, device=qkv.device)
    dkv32 = torch.empty(N, 2 * nkv * hd, dtype=f32, device=qkv.device)
    L.check(L.lib().iadr1_fmha_bwd(qkv.data_ptr(), dout.data_ptr(), out.data_ptr(), lse2.data_ptr(), N, nq, nkv, hd,
                                   plan.rng.data_ptr(), plan.q_items.data_ptr(), plan.n_q, plan.sched_dq.data_ptr(),
                                   plan.n_cta_dq, plan.k_items.data_ptr(), plan.n_k, plan.sched_kv.data_ptr(), plan.n_cta_kv,
                                   dqkv.data_ptr(), delta.data_ptr(), dkv32.data_ptr(), plan.npad, scale, _PV_N,
                                   L.stream_ptr()), "fmha_bwd")
    return dqkv


class FusedAttention:
    """Attention strategy object (same interface as the composed ops.FullAttention / ops.SharedPrefixAttention):
    forward(qkv) -> (out, saved), backward(dattn, qkv, saved) -> dqkv."""

    def __init__(self, plan: FmhaPlan, nq: int, nkv: int, hd: int):
        self.plan, self.nq, self.nkv, self.hd = plan, nq, nkv, hd
        self.D = (nq + 2 * nkv) * hd
        self.scale = float(hd) ** -0.5
        self.n_tokens = plan.n_tokens

    def forward(self, qkv, out=None):
        o, lse2 = fmha_fwd(qkv, self.plan, self.nq, self.nkv, self.hd, self.scale, out=out)
        return o, (o, lse2)

    def backward(self, dattn, qkv, saved, out=None):
        o, lse2 = saved
        return fmha_bwd(dattn, qkv, o, lse2, self.plan, self.nq, self.nkv, self.hd, self.scale, dqkv=out)


def reference_mask(rng: np.ndarray) -> np.ndarray:
    """Dense boolean [N, N] mask of a range table (tests)."""
    k = np.arange(rng.shape[0])[None, :]
    r = rng.astype(np.int64)
    return ((k >= r[:, 0:1]) & (k < r[:, 1:2])) | ((k >= r[:, 2:3]) & (k < r[:, 3:4]))
