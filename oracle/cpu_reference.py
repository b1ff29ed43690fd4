"""ORACLE / CPU baseline (test infrastructure, never on the product path): one reference-form SC-GRPO group step on the
host cores - HF model, HF `generate` standing in for vLLM, `compute_loss` restated (oracle/grpo_ref.py), autograd
backward, torch.optim.AdamW - i.e. what ref: train/stage_rl/trainer/sc_grpo_trainer.py:586-819 + Trainer.training_step
execute, minus DeepSpeed. The reference trainer itself cannot be imported here (accelerate / deepspeed / vllm /
sentence_transformers missing), hence kind = "port" in bench.py's cpu_baseline.

Full-size models do not fit a bounded CPU sample (3B: ~290 TFLOP per group, > 60 GB fp32 state), so the sample is a
DEPTH-REDUCED twin (true widths, few layers) with a short completion; `reference_form_flops` gives the analytic FLOPs of
sample and full workload so the caller can scale (bench.py states both in `cpu_baseline.sample`).
"""
from __future__ import annotations

import copy
import time

import torch

from . import grpo_ref
from .hf_oracle import build_hf_model, per_token_logps


def reference_form_flops(cfg, G, P, C, n_patches):
    """FLOPs of one group in the REFERENCE's form: vision tower run G x, lm_head on all T positions,
    policy fwd + ref fwd + 2x backward (no recompute term: the CPU port does not checkpoint), + rollout prefill/decode."""
    t, v = cfg.text, cfg.vision
    T = P + C
    Wd = t.num_layers * (2 * t.hidden_size * t.num_heads * t.head_dim + 2 * t.hidden_size * t.num_kv_heads * t.head_dim
                         + 3 * t.hidden_size * t.intermediate_size)
    Wh = t.vocab_size * t.hidden_size
    E = v.hidden_size
    mlp = (3 if v.kind == "qwen2_5_vl" else 2) * E * v.intermediate_size
    Wblk = 4 * E * E + mlp
    m = v.spatial_merge_size ** 2 * E
    if v.kind == "siglip":   # projector: Linear(E -> H) + Linear(H -> H) per patch token
        head = 2 * n_patches * (E * v.out_hidden_size + v.out_hidden_size ** 2)
    else:                    # merger: two Linears over 4-patch units
        head = 2 * (n_patches // v.spatial_merge_size ** 2) * (m * m + m * v.out_hidden_size)
    Fv = 2 * n_patches * (v.patch_dim * E + v.depth * Wblk) + head
    attn = t.num_layers * 2 * T * T * t.num_heads * t.head_dim
    fwd = G * (2 * T * (Wd + Wh) + attn + Fv)
    train = 4 * fwd
    rollout = (2 * P * Wd + Fv) * G + G * C * 2 * (Wd + Wh)
    return float(train + rollout)


class CPUReference:
    """`device="cuda"` + `dtype=torch.bfloat16` runs the SAME reference-form step with the HF modules on the GPU
    (BASELINE.md §4.4 "optional GPU reference": context for the >= 10x target; bench.py --impl reference --ref-device cuda)."""

    def __init__(self, cfg, seed=0, threads=None, device="cpu", dtype=torch.float32, attn_implementation="eager"):
        if threads:
            torch.set_num_threads(threads)
        self.cfg = cfg
        self.device = torch.device(device)
        self.policy = build_hf_model(cfg, seed=seed, dtype=dtype, attn_implementation=attn_implementation).to(self.device)
        self.policy.train(False)
        self.ref = copy.deepcopy(self.policy).requires_grad_(False)
        self.opt = torch.optim.AdamW(self.policy.parameters(), lr=1e-6, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)

    def group_step(self, prompt_ids, pixel_values, grid_thw, G, C, reward_fn, beta=0.04, seed=0):
        """prompt_ids [P] LongTensor; Qwen families: pixel_values [Np, patch_dim] float + grid_thw [[t,h,w]];
        LLaVA-OneVision: pixel_values [1, n_crops, 3, S, S] + grid_thw = image_sizes [[H, W]]. Returns (seconds, loss)."""
        cfg, dev = self.cfg, self.device
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        torch.manual_seed(seed)
        prompt_ids = prompt_ids.to(dev)
        pixel_values = pixel_values.to(dev, next(self.policy.parameters()).dtype)
        ids = prompt_ids[None, :].repeat(G, 1)                                   # sc_grpo_trainer.py:624-628
        if cfg.family == "llava_onevision":
            mm = dict(pixel_values=pixel_values.repeat(G, 1, 1, 1, 1), image_sizes=torch.tensor(grid_thw * G, device=dev))
        else:
            mm = dict(pixel_values=pixel_values.repeat(G, 1), image_grid_thw=torch.tensor(grid_thw * G, device=dev))
        P = ids.shape[1]
        with torch.no_grad():                                                      # stands in for vLLM, :343-358, :667
            out = self.policy.generate(input_ids=ids, attention_mask=torch.ones_like(ids), do_sample=True, temperature=0.9,
                                       top_k=50, top_p=0.9, max_new_tokens=C, min_new_tokens=C,
                                       pad_token_id=cfg.pad_token_id, eos_token_id=cfg.eos_token_id, **mm)
        comp = out[:, P:]
        mask = grpo_ref.completion_mask_ref(comp, cfg.eos_token_id)
        attn = torch.cat([torch.ones(G, P, dtype=torch.long, device=dev), mask.long()], 1)
        kw = dict(input_ids=out, attention_mask=attn, use_cache=False, **mm)
        logps = per_token_logps(self.policy(**kw).logits, out)[:, P - 1:]          # :733-735
        with torch.no_grad():
            ref_logps = per_token_logps(self.ref(**kw).logits, out)[:, P - 1:]     # :737-743
        rewards = torch.tensor(reward_fn(comp), dtype=torch.float32, device=dev).view(G, -1)
        adv, _, _ = grpo_ref.advantages_ref(rewards, G)
        loss, _ = grpo_ref.sc_grpo_loss_ref(logps, ref_logps, adv, mask, beta)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.policy.parameters(), 1.0)
        self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        loss_v = float(loss.detach())                                            # device sync on the GPU arm
        return time.perf_counter() - t0, loss_v
