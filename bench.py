#!/usr/bin/env python
"""bench.py - the SC-GRPO hot path on N x B200 (one process per GPU), BASELINE.json's metric:
GRPO groups/sec, G=8, Qwen2.5-VL-3B geometry (random init, synthetic 448x448 images), bf16.

One "step" = one optimizer step of the trainer = world x per_device_batch x grad_accum groups, each taken through
rollout (G completions x C tokens) -> policy log-prob forward -> reference log-prob forward -> rewards / advantages /
SC-GRPO loss (Python) -> backward -> gradient all-reduce -> fused AdamW.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nproc-per-node N ... bench.py --gpus N ...

`value`  : groups/sec with prompts pre-encoded and resident in HBM (pixel_values, token ids on the device).
`e2e`    : the same through the public trainer API from HOST inputs (PIL images -> HF image processor -> pinned host
           buffers -> H2D, completions D2H for the Python reward callbacks) inside the timed region.
`roofline`: the tcgen05 GEMM (dominant kernel): algorithmic FLOPs / CUDA-event time of every non-graph GEMM launch in
           the timed region, against MEASURED_PEAKS.json's sustained bf16 peak.
`cpu_baseline` / `--impl reference`: the reference-form HF/torch CPU step (oracle/cpu_reference.py) on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

# stdout carries exactly ONE JSON line: NCCL's own version / debug banner goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="qwen2.5-vl-3b")
    ap.add_argument("--ga", type=int, default=16,
                    help="groups per optimizer step per rank = groups rolled out together (G x ga rows decode in lock-step; "
                         "the reference scripts use per_device_batch 1 x grad_accum 2)")
    ap.add_argument("--groups-per-pass", type=int, default=2,
                    help="per_device_train_batch_size: groups packed into one forward/backward pass")
    ap.add_argument("--completion-len", type=int, default=512)
    ap.add_argument("--image-size", type=int, default=448)
    ap.add_argument("--num-generations", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only. cpu (default, the contract's reference arm): bounded depth-reduced sample on "
                         "the host cores. cuda: BASELINE.md §4.4's optional arm - the same HF modules + HF generate + restated "
                         "loss, FULL size, bf16, on one B200 (context for the >= 10x target; not the contract's reference line)")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    def __init__(self):
        self.proc, self.lines, self.thread = None, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", os.environ.get("LOCAL_RANK", "0"),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(args, cfg_full, steps=1, warmup=0):
    """Reference-form CPU step on a bounded, depth-reduced twin; returns (groups_per_sec_scaled, dict)."""
    import copy
    import torch
    from iad_r1_b200.synthetic import SyntheticProcessor, synthetic_dataset
    from oracle.cpu_reference import CPUReference, reference_form_flops
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    twin = copy.deepcopy(cfg_full)
    twin.text.num_layers = 2
    twin.vision.depth = 2
    twin.vision.fullatt_block_indexes = (1,) if twin.vision.kind == "qwen2_5_vl" else ()
    G, C_s = args.num_generations, 16
    proc = SyntheticProcessor(twin, max_pixels=480000)
    ex = synthetic_dataset(1, args.image_size)[0]
    enc = proc(text=[proc.apply_chat_template(ex["prompt"])], images=ex["image"])
    ids = enc["input_ids"][0]
    if twin.family == "llava_onevision":      # HF layout: crops + image sizes
        px, grid = enc["pixel_values"], enc["image_sizes"].tolist()
        Np = px.shape[1] * twin.vision.tokens_per_crop
    else:
        px, grid = enc["pixel_values"], enc["image_grid_thw"].tolist()
        Np = px.shape[0]
    ref = CPUReference(twin, seed=0, threads=cores)
    P = ids.shape[0]
    times = []
    for i in range(warmup + steps):
        dt, _ = ref.group_step(ids, px, grid, G, C_s, lambda comp: torch.randn(comp.shape[0]).tolist(), seed=i)
        if i >= warmup:
            times.append(dt)
    t_s = sum(times) / len(times)
    f_s = reference_form_flops(twin, G, P, C_s, Np)
    f_full = reference_form_flops(cfg_full, G, P, args.completion_len, Np)
    gps = (1.0 / t_s) * (f_s / f_full)
    info = {"value": gps, "unit": "groups/s", "cores": cores, "kind": "port",
            "sample": (f"measured {t_s:.1f} s per group on a depth-reduced twin (true widths, {twin.text.num_layers} decoder "
                       f"layers, {twin.vision.depth} vision blocks, G={G}, P={P}, C={C_s}; HF generate + 2 log-prob forwards "
                       f"+ backward + AdamW, fp32, {cores} threads, {f_s / t_s / 1e9:.0f} GFLOP/s); scaled by reference-form FLOPs "
                       f"{f_s / 1e12:.2f} -> {f_full / 1e12:.1f} TFLOP per group")}
    return gps, info, t_s


def hf_on_gpu(args, cfg):
    """full-size HF model + HF generate + restated compute_loss + autograd + torch AdamW on cuda:0, bf16
    weights, sdpa attention - the reference's arithmetic stack (minus vLLM / DeepSpeed, absent here) on the same GPU."""
    import torch
    from iad_r1_b200.synthetic import SyntheticProcessor, synthetic_dataset
    from oracle.cpu_reference import CPUReference
    G, C = args.num_generations, args.completion_len
    proc = SyntheticProcessor(cfg, max_pixels=480000)
    ex = synthetic_dataset(1, args.image_size)[0]
    enc = proc(text=[proc.apply_chat_template(ex["prompt"])], images=ex["image"])
    ids = enc["input_ids"][0]
    px = enc["pixel_values"]
    grid = (enc["image_sizes"] if cfg.family == "llava_onevision" else enc["image_grid_thw"]).tolist()
    ref = CPUReference(cfg, seed=0, device="cuda:0", dtype=torch.bfloat16, attn_implementation="sdpa")
    times = []
    W, K = min(args.warmup, 1), max(1, min(args.steps, 3))
    for i in range(W + K):
        dt, _ = ref.group_step(ids, px, grid, G, C, lambda comp: torch.randn(comp.shape[0]).tolist(), seed=i)
        if i >= W:
            times.append(dt)
    t_s = sum(times) / len(times)
    info = {"value": 1.0 / t_s, "unit": "groups/s", "cores": 0, "kind": "port",
            "sample": (f"HF-on-B200 (BASELINE.md §4.4): FULL-size {args.model}, bf16, sdpa, one group per step (G={G}, "
                       f"P={ids.shape[0]}, C={C}): HF generate + policy / reference log-prob forwards + autograd backward + "
                       f"torch AdamW, {t_s:.2f} s per group over {K} groups, peak {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")}
    return 1.0 / t_s, info, t_s


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from iad_r1_b200.config import PRESETS
    cfg = PRESETS[args.model]()
    t0 = time.time()
    if args.ref_device == "cuda":
        gps, info, t_s = hf_on_gpu(args, cfg)
        line = {"metric": f"GRPO groups/sec (G={args.num_generations})", "value": gps, "unit": "groups/s", "impl": "reference",
                "ref_device": "cuda", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1000.0 * args.ga / gps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic", "config": workload_config(args, cfg, 1), "cpu_baseline": info,
                "e2e": {"value": gps, "unit": "groups/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "wall_s": time.time() - t0}
        print(json.dumps(line), flush=True)
        return
    gps, info, t_s = cpu_reference_sample(args, cfg, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    line = {"metric": f"GRPO groups/sec (G={args.num_generations})", "value": gps, "unit": "groups/s", "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * args.ga / gps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, cfg, 1), "cpu_baseline": info,
            "e2e": {"value": gps, "unit": "groups/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.time() - t0}
    print(json.dumps(line), flush=True)


def workload_config(args, cfg, world):
    return {"workload": f"{args.model} SC-GRPO G={args.num_generations}, one {args.image_size}x{args.image_size} synthetic "
                        f"image per prompt, C={args.completion_len} (fixed length), {args.ga} groups per optimizer step per GPU "
                        f"(per_device_batch={min(args.groups_per_pass, args.ga)} x grad_accum={args.ga // max(1, min(args.groups_per_pass, args.ga))}), "
                        f"beta=0.04 (reference model on), random-init weights",
            "groups_per_step": world * args.ga, "parallelism": f"dp{world}",
            "l2": "working set (bf16 weights of policy + reference and GBs of activations per pass) >> 126 MB L2; "
                  "no explicit flush needed"}


# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from iad_r1_b200 import lib as L
    from iad_r1_b200.config import PRESETS
    from iad_r1_b200.grpo_config import GRPOConfig
    from iad_r1_b200.synthetic import SyntheticProcessor, format_reward, make_noise_reward, synthetic_dataset
    from iad_r1_b200.trainer import SCGRPOTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    cfg = PRESETS[args.model]()
    G, GA, Cl = args.num_generations, args.ga, args.completion_len
    bs = max(1, min(args.groups_per_pass, GA))
    if GA % bs:
        raise SystemExit("--ga must be a multiple of --groups-per-pass")
    targs = GRPOConfig(output_dir="/tmp/iadr1_bench", per_device_train_batch_size=bs, gradient_accumulation_steps=GA // bs,
                       num_generations=G, max_prompt_length=4096, max_completion_length=Cl, bf16=True, beta=0.04,
                       logging_steps=0, save_strategy="no", rollout_forbid_eos=True, temperature=0.9, seed=42)
    proc = SyntheticProcessor(cfg, max_pixels=480000)
    trainer = SCGRPOTrainer(model=cfg, reward_funcs=[format_reward, make_noise_reward(rank)], args=targs,
                            processing_class=proc, max_pixels=480000)
    dev = trainer.device
    n_steps = args.warmup + args.steps
    data = synthetic_dataset(GA * n_steps + GA * (1 + args.steps), args.image_size)
    trainer.state.max_steps = 10 ** 6  # keep the LR schedule flat-ish over the benchmark

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- leg 1: device-resident inputs -------------------------------------------------------------------------------
    encoded = []
    for ex in data[: GA * n_steps]:
        e = trainer._encode_prompt(ex)
        e["pixel_values"] = e["pixel_values"].to(dev, dtype=torch.bfloat16)
        encoded.append((ex, e))

    def step_resident(i):
        win = encoded[i * GA:(i + 1) * GA]
        trainer.prepare_window([ex for ex, _ in win], encoded=[e for _, e in win])   # pre-encoded prompts, HBM-resident pixels
        for j in range(0, GA, bs):
            trainer.training_step([ex for ex, _ in win[j:j + bs]])
        trainer.optimizer_step()

    for i in range(args.warmup):
        step_resident(i)
    trainer.flush_timers()
    trainer.phase_ms.clear()
    L.reset_launch_count()
    eng = trainer._engine
    replay0 = eng.replays if eng is not None else 0
    tok0 = trainer.total_rollout_tokens
    clocks = ClockSampler()
    barrier()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.warmup, n_steps):
        step_resident(i)
    e1.record()
    barrier()
    clk = clocks.stop()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    eng = trainer._engine
    launches = L.launch_count() + (eng.replays - replay0) * eng.kernels_per_step
    trainer.flush_timers()
    phases = {k: v / args.steps for k, v in trainer.phase_ms.items()}
    rollout_tokens = trainer.total_rollout_tokens - tok0
    ms_per_step = ms_total / args.steps
    groups_per_step = world * GA
    value = groups_per_step / (ms_per_step / 1e3)

    # ---- profile leg (NOT in the timed region): one more step with a CUDA-event pair around every GEMM launch -> live roofline
    L.check(L.lib().iadr1_gemm_profile_enable(1))
    step_resident(n_steps - 1)
    torch.cuda.synchronize()
    tms, tfl, tmax, nl = C.c_double(), C.c_double(), C.c_double(), C.c_longlong()
    L.lib().iadr1_gemm_profile_collect.argtypes = [C.POINTER(C.c_double)] * 3 + [C.POINTER(C.c_longlong), C.c_char_p]
    shape_csv = os.environ.get("IADR1_GEMM_SHAPES_CSV", "") if rank == 0 else ""
    L.check(L.lib().iadr1_gemm_profile_collect(C.byref(tms), C.byref(tfl), C.byref(tmax), C.byref(nl), shape_csv.encode()))
    L.check(L.lib().iadr1_gemm_profile_enable(0))
    trainer.flush_timers()
    trainer.phase_ms.clear()

    # ---- leg 2: end to end from host inputs through the public API ----------------------------------------------------
    e2e = None
    if not args.no_e2e:
        host = data[GA * n_steps:]

        def step_e2e(i):
            win = host[i * GA:(i + 1) * GA]
            trainer.prepare_window(win)            # PIL -> HF image processor -> pinned host -> H2D -> rollout
            for j in range(0, GA, bs):
                trainer.training_step(win[j:j + bs])   # completions D2H for the Python reward callbacks
            trainer.optimizer_step()

        step_e2e(0)  # warm the host path once (PIL / processor caches)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(1, 1 + args.steps):
            step_e2e(i)
        f1.record()
        barrier()
        ms_e2e = max_over_ranks(f0.elapsed_time(f1)) / args.steps
        px_bytes = int(encoded[0][1]["pixel_values"].numel()) * 4  # processor emits fp32 on the host
        ids_bytes = len(encoded[0][1]["input_ids"]) * 8
        e2e = {"value": groups_per_step / (ms_e2e / 1e3), "unit": "groups/s",
               "h2d_bytes_per_step": GA * (px_bytes + ids_bytes + G * (ids_bytes + Cl * 8)),
               "d2h_bytes_per_step": GA * (G * Cl * 4 + 4), "ms_per_step": ms_e2e}

    if rank == 0:
        hbm, tf_peak, src = peaks()
        roof = {"bound": "tensor", "kernel": "gemm_bf16_tcgen05_kernel", "achieved": tfl.value / (tms.value / 1e3) / 1e12
                if tms.value > 0 else None, "peak": tf_peak, "unit": "TFLOP/s", "peak_source": f"bf16_tflops_sustained, {src}",
                "launches_timed": nl.value, "gemm_ms_per_step": tms.value,
                "measured_in": "separate profile step (per-launch CUDA events), not in the timed region",
                # DRAM bytes of ONE launch of the training shape with the WORST traffic-to-operand ratio in the committed
                # `ncu --set full` capture profiles/r02_gemm_full_ncu_summary.txt: decoder gate_up dgrad, M = 8786 tokens of two
                # packed groups, [M, 22016] x [22016, 2048]: dram__bytes_read.sum + dram__bytes_write.sum = 792 MB against
                # 513 MB of operands + output (1.54x; gate_up fwd 1.00x, down fwd 1.49x, gate_up wgrad 1.30x)
                "traffic": 758.670592e6 + 33.501184e6, "traffic_algorithmic": 513.4e6,
                "traffic_unit": "bytes per launch (gate_up dgrad, worst shape of the ncu capture)"}
        roof["frac"] = roof["achieved"] / tf_peak if roof["achieved"] else None
        # second roofline: the rollout decode step is HBM-bound - every decoder weight + the lm_head once per step, plus the
        # KV cache of every row (prompt K/V once per group); algorithmic bytes per step / measured time per step
        t_ = cfg.text
        wd = t_.num_layers * (t_.qkv_dim * t_.hidden_size + t_.num_heads * t_.head_dim * t_.hidden_size
                              + 3 * t_.hidden_size * t_.intermediate_size) * 2
        wh = t_.vocab_size * t_.hidden_size * 2
        kv_tok = 2 * t_.num_layers * t_.num_kv_heads * t_.head_dim * 2
        p_len = len(encoded[0][1]["input_ids"])
        kv_avg = (GA * p_len + GA * G * (Cl / 2.0)) * kv_tok
        step_bytes = wd + wh + kv_avg
        dec_ms = phases.get("rollout", float("nan")) / max(1, Cl)
        dec = {"bound": "hbm", "kernel": f"decode step (CUDA graph: {t_.num_layers} x [attention, persistent chain: o -> rmsnorm -> gate_up+SwiGLU -> down -> rmsnorm -> qkv] + lm_head + sampler)",
               "achieved": step_bytes / (dec_ms / 1e3) / 1e9, "peak": hbm, "unit": "GB/s", "peak_source": f"hbm_gbs, {src}",
               "bytes_per_step": step_bytes, "ms_per_decode_step": dec_ms}
        dec["frac"] = dec["achieved"] / hbm
        line = {"metric": f"GRPO groups/sec (G={G})", "value": value, "unit": "groups/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, cfg, world),
                "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof,
                "rollout": {"tok_per_s": world * rollout_tokens / args.steps / (phases.get("rollout", float("nan")) / 1e3),
                            "rows_in_flight": GA * G, "ms_per_step": phases.get("rollout"), "roofline": dec},
                "phase_ms_per_step": phases}
        if world == 1 and not args.no_cpu_baseline:
            try:
                _, info, _ = cpu_reference_sample(args, cfg, steps=1, warmup=0)
                line["cpu_baseline"] = info
            except Exception as ex:  # the baseline is reported, never required for the GPU numbers
                line["cpu_baseline"] = {"value": None, "unit": "groups/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {type(ex).__name__}: {ex}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
