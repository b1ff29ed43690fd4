"""Times the rollout decode step and each of its kernels in isolation (CUDA events, warm, 3B widths, few layers).
Usage (GPU box): python tools/decode_probe.py [groups] [layers]"""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iad_r1_b200 import lib as L, ops
from iad_r1_b200.config import PRESETS
from iad_r1_b200.params import ParamStore
from iad_r1_b200.model import VLM
from iad_r1_b200.rollout import RolloutEngine, _split_for
from iad_r1_b200.synthetic import SyntheticProcessor, synthetic_dataset

groups = int(sys.argv[1]) if len(sys.argv) > 1 else 2
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda:0")
cfg = PRESETS["qwen2.5-vl-3b"]()
cfg.text.num_layers = layers
cfg.vision.depth = 2
cfg.vision.fullatt_block_indexes = (1,)
ps = ParamStore(cfg, dev)
ps.init_random(0)
vlm = VLM(cfg, ps)
proc = SyntheticProcessor(cfg, max_pixels=480000)
G, C = 8, 512
encs = []
for ex in synthetic_dataset(groups, 448):
    e = proc(text=[proc.apply_chat_template(ex["prompt"])], images=ex["image"])
    encs.append(dict(input_ids=e["input_ids"][0].numpy(), pixel_values=e["pixel_values"].to(dev), grid_thw=e["image_grid_thw"].tolist()))


def timeit(fn, n=20, reps=10):
    """GPU-side time per launch: n launches captured in a CUDA graph (no Python / launch overhead), replayed reps times."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * reps) * 1000  # us


NEW = int(os.environ.get("PROBE_NEW_TOKENS", "64"))
for graph in (False, True):
    eng = RolloutEngine(vlm, groups, G, 320, C, forbid_eos=True, use_cuda_graph=graph)
    out, st = eng.generate(encs, seed=1, max_new_tokens=NEW)
    out, st = eng.generate(encs, seed=2, max_new_tokens=NEW)
    print(f"R={eng.R} layers={layers} graph={graph}: {st['decode_ms'] / st['steps'] * 1000:.1f} us per decode step "
          f"({st['decode_ms'] / st['steps'] * 1000 / layers:.1f} us per layer incl. head)", flush=True)

if os.environ.get("PROBE_TRACE"):
    # in-graph timeline: per-kernel critical-path time = gap between successive dependency-resolved stamps
    import ctypes, numpy as np
    lib = L.lib()
    nwords = 2 + 2 * 8190
    buf = (ctypes.c_ulonglong * nwords)()
    nk = eng.kernels_per_step
    per_layer = ["rms1", "qkv", "attn", "o", "rms2", "gate_up", "silu", "down"]
    if eng.gu_mode.startswith("fused"):
        per_layer.remove("silu")
    names = ["embed"] + per_layer * layers + ["rms_f", "lm_head", "sample", "advance"]
    if os.environ.get("IADR1_DECODE_CHAIN", "1") != "0" and eng.R <= 128 and eng._native_ok:
        # persistent decode-layer chain: 2 launches per layer (csrc/decode_chain.cu)
        names = ["embed", "chain0"] + ["attn", "chain"] * layers + ["lm_head", "sample", "advance"]
    assert nk == len(names), (nk, len(names))
    eng.state[0] = int(os.environ.get("PROBE_CTX_STEP", "256"))
    torch.cuda.synchronize()
    L.check(lib.iadr1_trace(1, None, 0))
    reps = 20
    for _ in range(reps):
        eng._graph.replay()
    torch.cuda.synchronize()
    L.check(lib.iadr1_trace(2, buf, nwords))
    L.check(lib.iadr1_trace(0, None, 0))
    a = np.frombuffer(buf, dtype=np.uint64)
    n = int(a[0])
    rec = a[2:2 + 2 * n].reshape(n, 2).astype(np.int64)
    marks = rec[rec[:, 0] < 4096]                            # in-kernel phase marks of the persistent chain (tag, time)
    rec = rec[rec[:, 0] >= 4096]
    n = len(rec)
    assert n == reps * nk, (n, reps, nk)
    vals = marks[marks[:, 0] >= 1000]
    marks = marks[marks[:, 0] < 1000]
    if len(vals):
        nm = {1000: "producer loop cycles", 1001: "  cycles issuing activation tiles", 1002: "  activation tiles", 1003: "  cycles issuing weight tiles",
              1004: "  weight tiles", 1005: "  cycles polling dependencies", 1006: "  loop iterations"}
        print("CTA 0 producer thread, one chain launch (clock64):")
        for tag in sorted(nm):
            v = vals[vals[:, 0] == tag][:, 1]
            if len(v):
                print(f"    {nm[tag]:40s} {np.median(v):10.0f}")
    am = marks[(marks[:, 0] >= 400) & (marks[:, 0] < 420)]
    marks = marks[(marks[:, 0] < 400) | (marks[:, 0] >= 420)]
    if len(am):
        # grouped decode attention: first P-kind CTA (tags 400..406) and first C-kind CTA (410..415) of the last launches
        nm = {400: "P-kind CTA 0 reaches wait", 401: "  wait returns", 402: "  Q tile built (64 rows)", 403: "  chunk loop done",
              404: "  partials + tickets done", 405: "  end (no merge)", 406: "  end (merged rows)",
              410: "C-kind CTA 0 reaches wait", 411: "  wait returns", 412: "  chunk loop done", 413: "  partials + ticket done",
              414: "  end (no merge)", 415: "  end (merged)"}
        for first, lo, hi in ((400, 400, 410), (410, 410, 420)):
            sel = am[(am[:, 0] >= lo) & (am[:, 0] < hi)]
            st_ = np.nonzero(sel[:, 0] == first)[0]
            if len(st_) >= 3:
                seg = sel[st_[-2]:st_[-1]]
                t0 = seg[seg[:, 0] == first + 1][0, 1] if (seg[:, 0] == first + 1).any() else seg[0, 1]
                for tag, tm in seg:
                    print(f"    {nm.get(int(tag), int(tag)):32s} {(tm - t0) / 1e3:8.2f}")
    if len(marks):
        # last replay, one mid-stack chain launch: times relative to the kernel's dependency wait returning
        mk = marks[-(len(marks) // reps):]
        starts = np.nonzero(mk[:, 0] == 100)[0]
        if len(starts) > 3:
            seg = mk[starts[2]:starts[3]]
            t_wait = rec.reshape(reps, nk, 2)[-1, 2 + 2 * 1 + 1, 1]   # leave stamp of the 2nd chain kernel (after attention 1)
            tagname = {100: "producer reaches wait", 110: "o acts ready", 111: "norm2 sees o done", 112: "gate_up acts ready",
                       113: "down acts ready", 114: "norm1 sees down done", 115: "qkv acts ready", 120: "CTA0 published o",
                       121: "CTA0 published norm2", 122: "CTA0 published gate_up", 123: "CTA0 published down",
                       124: "CTA0 published norm1", 125: "CTA0 published qkv"}
            print("chain kernel (layer 1) phase marks, us after its dependency wait returned:")
            for ph, nm in ((0, "o"), (2, "gate_up"), (3, "down"), (5, "qkv")):
                for j in range(8):
                    tagname[200 + ph * 10 + j] = f"  mma {nm} k-block {8 * j} landed"
            for j in range(8, 16):
                tagname[300 + j] = f"      gate_up kb {j}: weights issued"
                tagname[320 + j] = f"      gate_up kb {j}: acts issued"
                tagname[340 + j] = f"      gate_up kb {j}: landed, mma issued"
            for tag, tm in seg[np.argsort(seg[:, 1])]:
                print(f"    {tagname.get(int(tag), int(tag)):40s} {(tm - t_wait) / 1e3:8.2f}")
    rec = rec.reshape(reps, nk, 2)[2:]                      # drop warm-up replays
    leave = rec[:, :, 1]
    dur = np.diff(np.concatenate([leave, leave[:, -1:] ], 1), axis=1)[:, :-1]     # kernel k: leave[k+1] - leave[k]
    early = (rec[:, :, 1] - rec[:, :, 0])                    # how long CTA 0 sat resident before its dependency resolved
    agg = {}
    for k, nm in enumerate(names[:-1]):
        agg.setdefault(nm, []).append((dur[:, k].mean() / 1e3, early[:, k].mean() / 1e3))
    print(f"in-graph timeline at completion step {int(eng.state[0])} (R={eng.R}, {layers} layers), us: critical-path time | resident-before-ready")
    tot = 0
    for nm, v in agg.items():
        d, e = np.mean([x[0] for x in v]), np.mean([x[1] for x in v])
        tot += d * len(v)
        print(f"  {nm:8s} x{len(v):3d}  {d:7.2f} | {e:6.2f}")
    print(f"  step total {(leave[:, -1] - leave[:, 0]).mean() / 1e3:.1f} us (sum of parts {tot:.1f})")
    sys.exit(0)
if os.environ.get("PROBE_STEP_ONLY"):
    sys.exit(0)
t, p, lib = cfg.text, vlm.p, L.lib()


class _S:
    def __index__(self):
        return L.stream_ptr()


s = None
R, H, I, nq, nkv, hd = eng.R, t.hidden_size, t.intermediate_size, t.num_heads, t.num_kv_heads, t.head_dim
eng.state[0] = 300  # mid-rollout context
b = "layers.0."
sk_qkv, sk_o, sk_d = _split_for(t.qkv_dim, H), _split_for(H, nq * hd), _split_for(H, I)
gu32 = torch.zeros(R, 2 * I, dtype=torch.float32, device=dev)
items = {
    "decode_embed": lambda: lib.iadr1_decode_embed(p["embed_tokens.weight"].data_ptr(), eng.tok.data_ptr(), eng.h.data_ptr(), R, H, L.stream_ptr()),
    "rmsnorm_f32in(+zero)": lambda: lib.iadr1_rmsnorm_f32in(eng.h.data_ptr(), p[b + "ln1.weight"].data_ptr(), eng.xn.data_ptr(), R, H, 1e-6, eng.qkv.data_ptr(), t.qkv_dim, L.stream_ptr()),
    f"gemm qkv split{sk_qkv}": lambda: eng._skinny(p[b + "qkv.weight"], eng.xn, eng.qkv, split_k=sk_qkv, atomic=True, bias=p[b + "qkv.bias"]),
    "gemm qkv split1": lambda: eng._skinny(p[b + "qkv.weight"], eng.xn, eng.qkv, split_k=1, atomic=True, bias=p[b + "qkv.bias"]),
    "gemm qkv split7 stages4": lambda: L.gemm(p[b + "qkv.weight"], eng.xn, out=eng.qkv, trans_out=True, split_k=sk_qkv, atomic=True, bias=p[b + "qkv.bias"], bias_per_m=True, block_n=eng.block_n, stages=4),
    "attn_fused": lambda: lib.iadr1_decode_attention_fused(
        eng.qkv.data_ptr(), eng.cos_tab.data_ptr(), eng.sin_tab.data_ptr(), eng.rope_delta.data_ptr(), eng.kp[0].data_ptr(),
        eng.vp[0].data_ptr(), eng.kc[0].data_ptr(), eng.vc[0].data_ptr(), eng.state.data_ptr(), eng.row_group.data_ptr(),
        eng.row_plen.data_ptr(), None, eng.part.data_ptr(), eng.tickets.data_ptr(), eng.attn.data_ptr(), R, nq, nkv, hd, eng.p_max,
        eng.c_max, (-eng.nsplit if eng.attn_nw == 2 else eng.nsplit), eng.max_pos, hd ** -0.5, L.stream_ptr()),
    f"gemm o split{sk_o}": lambda: eng._skinny(p[b + "o.weight"], eng.attn, eng.h, split_k=sk_o, atomic=True),
    "gemm gate_up stream-k f32 atomic": lambda: eng._skinny(p[b + "gate_up.weight"], eng.xn, eng.gu, atomic=True, stream_k=True),
    "gemm qkv stream-k": lambda: eng._skinny(p[b + "qkv.weight"], eng.xn, eng.qkv, stream_k=True, bias=p[b + "qkv.bias"]),
    "gemm o stream-k": lambda: eng._skinny(p[b + "o.weight"], eng.attn, eng.h, stream_k=True),
    "gemm down stream-k": lambda: eng._skinny(p[b + "down.weight"], eng.act, eng.h, stream_k=True),
    "silu_mul_f32": lambda: lib.iadr1_decode_silu_mul_f32(eng.gu.data_ptr(), eng.act.data_ptr(), R, I, L.stream_ptr()),
    "gemm gate_up split2 f32 atomic": lambda: L.gemm(p[b + "gate_up.weight"], eng.xn, out=gu32, trans_out=True, split_k=2, atomic=True, block_n=eng.block_n, a_static=True),
    "gemm gate_up split3 f32 atomic": lambda: L.gemm(p[b + "gate_up.weight"], eng.xn, out=gu32, trans_out=True, split_k=3, atomic=True, block_n=eng.block_n, a_static=True),
    "gemm gate_up f32 no split": lambda: L.gemm(p[b + "gate_up.weight"], eng.xn, out=gu32, trans_out=True, block_n=eng.block_n, a_static=True),
    "gemm qkv split3": lambda: eng._skinny(p[b + "qkv.weight"], eng.xn, eng.qkv, split_k=3, atomic=True, bias=p[b + "qkv.bias"]),
    "gemm down split4": lambda: eng._skinny(p[b + "down.weight"], eng.act, eng.h, split_k=4, atomic=True),
    "gemm down split18": lambda: eng._skinny(p[b + "down.weight"], eng.act, eng.h, split_k=18, atomic=True),
    f"gemm down split{sk_d}": lambda: eng._skinny(p[b + "down.weight"], eng.act, eng.h, split_k=sk_d, atomic=True),
    "gemm lm_head": lambda: eng._skinny(vlm.params.lm_head, eng.xn, eng.logits),
    "sample": lambda: lib.iadr1_sample(eng.logits.data_ptr(), R, t.vocab_size, 0.9, 50, 0.9, 1, eng.state.data_ptr(), eng.tok.data_ptr(),
                                       eng.finished.data_ptr(), eng.out_tokens.data_ptr(), eng.c_max, cfg.eos_token_id, cfg.pad_token_id, 1, 0, L.stream_ptr()),
}
eng.finished.zero_()
tot = 0
only = sys.argv[3] if len(sys.argv) > 3 else None
if only:
    for name, fn in items.items():
        if only in name:
            for _ in range(8):
                fn()
            torch.cuda.synchronize()
            print('ran', name)
    sys.exit(0)
for name, fn in items.items():
    us = timeit(fn)
    print(f"{us:9.2f} us  {name}", flush=True)
print("note: back-to-back launches of ONE kernel inside a CUDA graph (that layer's weights stay L2-resident for the small ones)")
