"""`SCGRPOTrainer`: the reference's trainer surface (constructor signature, `.train()`, `.save_model()`, `.log()`,
`.state`, `._metrics`, reward-callback convention) over the B200-native hot path.

Mirrors ref: train/stage_rl/trainer/sc_grpo_trainer.py:72-373 (ctor), :586-819 (`compute_loss`), :821-827 (`log`) and
the parts of `transformers.Trainer.training_step` / DeepSpeed it relied on (backward, 1/GA loss scaling because
`model_accepts_loss_kwargs=False` :292-295, clip at `max_grad_norm`, AdamW, linear LR decay, seed-42 shuffling sharded
per rank). What changed underneath: rollout = in-rank RolloutEngine instead of a vLLM GPU (K19, K20, C3),
log-probs = fused kernels instead of HF modules + [G,T,V] logits (K1-K15), backward = hand-written kernels (K17),
optimizer = fused flat AdamW (K18), ZeRO-3 traffic = ONE gradient all-reduce per optimizer step (C1 -> C2).
The group-relative advantage, KL and loss stay in Python (grpo_loss.py).
"""
from __future__ import annotations

import json
import math
import os
import sys
import time
import warnings
from collections import defaultdict
from typing import Any, Callable, Optional, Union

import numpy as np
import torch

from . import grpo_loss
from . import lib as L
from .checkpoint import load_pretrained, save_pretrained
from .config import VLMConfig
from .grpo_config import GRPOConfig
from .model import VLM
from .geometry import vision_inputs_from_processor
from .params import ParamStore
from .rollout import RolloutEngine
from .trainer_base import TrainerCore, TrainerState

RewardFunc = Union[str, Callable[[list, list], list]]


def is_conversational(example: dict) -> bool:
    """ref: trl/trl/data_utils.py:30-68."""
    for key in ("prompt", "chosen", "rejected", "completion", "messages"):
        v = example.get(key)
        if isinstance(v, list) and v and isinstance(v[0], dict) and "role" in v[0] and "content" in v[0]:
            return True
    return False


def maybe_apply_chat_template(example: dict, processing_class) -> dict:
    """Prompt-only branch of ref: trl/trl/data_utils.py:71-200."""
    if is_conversational(example):
        return {"prompt": processing_class.apply_chat_template(example["prompt"], tokenize=False, add_generation_prompt=True)}
    return {"prompt": example["prompt"]}


class _LogProbFn(torch.autograd.Function):
    """Bridges torch autograd (the Python loss) and the CUDA forward/backward: `loss.backward()` in unchanged Python
    drives model.logprobs_backward."""

    @staticmethod
    def forward(ctx, anchor, vlm, batch, rows, labels, temperature, image_embeds=None, dimg_sink=None):
        logp, saved = vlm.logprobs_forward(batch, rows, labels, temperature, save=True, image_embeds=image_embeds,
                                           dimg_sink=dimg_sink)
        ctx.vlm, ctx.saved = vlm, saved
        return logp

    @staticmethod
    def backward(ctx, dlogp):
        ctx.vlm.logprobs_backward(dlogp.contiguous(), ctx.saved)
        ctx.saved = None
        return None, None, None, None, None, None, None, None


def _is_reward_model(rf) -> bool:
    """A reward *model* in the reference's sense (sc_grpo_trainer.py:232-236): a sequence-classification network, not a callback."""
    return isinstance(rf, torch.nn.Module)


def _reward_model_scores(model, tokenizer, example: dict, completions, conv: bool) -> torch.Tensor:
    """ref: sc_grpo_trainer.py:759-772 - prompt + completion rendered with the reward tokenizer's chat template, right-padded,
    `logits[:, 0]` of an `AutoModelForSequenceClassification(num_labels=1)` under inference mode. This branch is library code
    (Transformers / torch) exactly as in the reference: reward scoring is not on the hot path this repository replaces."""
    if conv:
        texts = [tokenizer.apply_chat_template(list(example["prompt"]) + c, tokenize=False) for c in completions]
    else:
        texts = [example["prompt"] + c for c in completions]
    enc = tokenizer(texts, return_tensors="pt", padding=True, padding_side="right", add_special_tokens=False)
    dev = next(model.parameters()).device
    enc = {k: v.to(dev) for k, v in enc.items()}
    with torch.inference_mode():
        return model(**enc).logits[:, 0].float().cpu()


def call_reward_funcs(reward_funcs, example: dict, texts: list, current_step: int, reward_processing_classes=None) -> torch.Tensor:
    """Calls every reward callback with the reference's convention (ref: train/stage_rl/trainer/sc_grpo_trainer.py:749-781):
    `reward_func(prompts=[prompt] * G, completions=..., current_step=global_step, **{column: [value] * G})` where
    completions are `[[{"role": "assistant", "content": text}]]` for conversational prompts (plain strings otherwise) and
    every dataset column except `prompt` / `completion` is repeated G times (Q10). Returns fp32 [G, n_funcs] on the CPU.
    Callbacks are called verbatim, not wrapped: a callback that returns the wrong number of rewards (Q9) raises here, as
    the tensor assignment does in the reference."""
    G = len(texts)
    conv = is_conversational(example)
    completions = [[{"role": "assistant", "content": c}] for c in texts] if conv else list(texts)
    prompts = [example["prompt"] for _ in range(G)]
    reward_kwargs = {k: [example[k]] * G for k in example.keys() if k not in ("prompt", "completion")}
    out = torch.zeros(G, len(reward_funcs), dtype=torch.float32)
    for i, rf in enumerate(reward_funcs):
        if _is_reward_model(rf):
            out[:, i] = _reward_model_scores(rf, reward_processing_classes[i], example, completions, conv)
            continue
        r = rf(prompts=prompts, completions=completions, current_step=current_step, **reward_kwargs)
        out[:, i] = torch.tensor(r, dtype=torch.float32)
    return out


class SCGRPOTrainer(TrainerCore):
    def __init__(self, model, reward_funcs, args: GRPOConfig = None, train_dataset=None, eval_dataset=None,
                 processing_class=None, reward_processing_classes=None, callbacks=None, optimizers=(None, None),
                 peft_config=None, max_pixels: Optional[int] = 12845056, min_pixels: Optional[int] = 3136,
                 attn_implementation: str = "flash_attention_2", use_vllm_for_gen: bool = True):
        if args is None:
            name = model if isinstance(model, str) else getattr(model, "family", "model")
            args = GRPOConfig(f"{str(name).split('/')[-1]}-GRPO")
        if peft_config is not None:
            raise NotImplementedError("peft_config: LoRA is outside the B200 hot path (SURVEY.md §8f)")
        if optimizers != (None, None):
            raise ValueError("custom optimizers are not supported: the fused flat AdamW is part of the hot path")
        self.args = args
        self.state = TrainerState()
        self._metrics = defaultdict(list)
        self._setup_distributed()
        torch.manual_seed(args.seed)
        np.random.seed(args.seed)

        # ---- model / reference ----------------------------------------------------------------------------------------
        if isinstance(model, str):
            self.model_id = model
            # The family comes from config.json's model_type (the reference's own PA-SFT output directories, e.g.
            # `Qwen_Intruct_2_5_VL_3B_Expert_AD_PA_SFT`, match none of its substring keys). The reference picks the HF class
            # by substring of the id (sc_grpo_trainer.py:116-137) and falls through to AutoModelForCausalLM otherwise; the
            # substring test survives here only as the error message for a path WITHOUT a readable config.json.
            if not os.path.isfile(os.path.join(model, "config.json")):
                families = ("qwen2-vl", "qwen2_vl", "qwen2vl", "qwen2.5-vl", "qwen2.5_vl", "qwen2.5vl", "qwen2_5_vl",
                            "llava-ov", "llava_ov", "llava_si", "llava-onevision", "llava_onevision", "llava-1_5", "llava_1_5", "llava-next", "llava_next", "llava_1_6")
                hint = "" if any(k in model.lower() for k in families) else \
                    " (and the id names none of the supported families: Qwen2-VL / Qwen2.5-VL / LLaVA-OneVision / LLaVA-1.5 / LLaVA-Next)"
                raise ValueError(f"Unsupported model: {model}: no config.json under that path{hint}; there is no hub access "
                                 f"on the training box, pass a local checkpoint directory")
            from .checkpoint import load_config
            mdt = ParamStore.plan_moment_dtype(load_config(model), self.device, args.beta != 0.0, args.optimizer_moments)
            self.cfg, self.params = load_pretrained(model, self.device, moment_dtype=mdt)
        elif isinstance(model, VLMConfig):
            self.model_id = model.family
            self.cfg = model
            mdt = ParamStore.plan_moment_dtype(model, self.device, args.beta != 0.0, args.optimizer_moments)
            self.params = ParamStore(model, self.device, with_grads=True, with_optimizer=True, moment_dtype=mdt)
            self.params.init_random(seed=args.seed)
        elif isinstance(model, ParamStore):
            self.model_id, self.cfg, self.params = model.cfg.family, model.cfg, model
        else:
            raise ValueError("model must be a checkpoint path, a VLMConfig (random init) or a ParamStore")
        if self.params.grad_flat is None or self.params.master is None:
            raise ValueError("the policy ParamStore needs with_grads=True, with_optimizer=True")
        self.model = VLM(self.cfg, self.params)
        # per-layer activation recompute (what --gradient_checkpointing asks the reference for): "auto" switches it on only
        # when the moments had to drop to bf16, i.e. the unsharded state already takes most of the device
        rc = args.activation_recompute
        self.model.recompute = rc == "on" or (rc == "auto" and self.params.exp_avg.dtype == torch.bfloat16)
        self.beta = args.beta
        # Q11: the reference always builds a frozen copy of the initial policy (sc_grpo_trainer.py:152-182). With
        # beta == 0 its only use (beta * KL) vanishes, so it is skipped and `kl` is logged as 0.
        if self.beta != 0.0:
            ref_store = ParamStore(self.cfg, self.device)
            ref_store.copy_from(self.params)
            self.ref_model = VLM(self.cfg, ref_store)
        else:
            self.ref_model = None

        # ---- processor ----------------------------------------------------------------------------------------------------
        if processing_class is None:
            if not isinstance(model, str):
                raise ValueError("processing_class is required when the model is not a checkpoint path")
            from transformers import AutoProcessor
            processing_class = AutoProcessor.from_pretrained(model)
            tok = getattr(processing_class, "tokenizer", processing_class)
            processing_class.pad_token_id = tok.pad_token_id
            processing_class.eos_token_id = tok.eos_token_id
            if hasattr(processing_class, "image_processor"):
                processing_class.image_processor.max_pixels = max_pixels
                processing_class.image_processor.min_pixels = min_pixels
        self.processing_class = processing_class

        # ---- rewards ------------------------------------------------------------------------------------------------------
        if not isinstance(reward_funcs, list):
            reward_funcs = [reward_funcs]
        reward_funcs = list(reward_funcs)
        for i, rf in enumerate(reward_funcs):
            if isinstance(rf, str):       # a reward MODEL id / directory (sc_grpo_trainer.py:232-236)
                from transformers import AutoModelForSequenceClassification
                reward_funcs[i] = AutoModelForSequenceClassification.from_pretrained(rf, num_labels=1, torch_dtype=torch.bfloat16)
            elif not callable(rf):
                raise ValueError(f"reward_funcs[{i}] must be a callable, a model id or a torch module, got {type(rf)}")
        self.reward_funcs = reward_funcs
        if reward_processing_classes is None:
            reward_processing_classes = [None] * len(reward_funcs)
        elif not isinstance(reward_processing_classes, list):
            reward_processing_classes = [reward_processing_classes]
        elif len(reward_processing_classes) != len(reward_funcs):
            raise ValueError("The number of reward processing classes must match the number of reward functions.")
        for i, (rpc, rf) in enumerate(zip(reward_processing_classes, reward_funcs)):
            if _is_reward_model(rf):      # :248-258: tokenizer of the reward model, pad = eos when missing, pad id into its config
                if rpc is None:
                    from transformers import AutoTokenizer
                    rpc = AutoTokenizer.from_pretrained(rf.config._name_or_path)
                if rpc.pad_token_id is None:
                    rpc.pad_token = rpc.eos_token
                rf.config.pad_token_id = rpc.pad_token_id
                rf.to(self.device).eval()
                reward_processing_classes[i] = rpc
        self.reward_processing_classes = reward_processing_classes

        self.max_pixels, self.min_pixels = max_pixels or 12845056, min_pixels or 3136
        self.max_prompt_length = args.max_prompt_length
        self.max_completion_length = args.max_completion_length
        self.num_generations = args.num_generations
        self.use_vllm = use_vllm_for_gen  # kept for surface compatibility; generation always runs in-rank
        self.train_dataset, self.eval_dataset = train_dataset, eval_dataset
        self.callbacks = callbacks or []
        self._engine: Optional[RolloutEngine] = None
        self._rollout_cache: dict = {}
        self._old_logps: dict = {}
        if args.num_iterations > 1 and args.loss_mode != "clip":
            raise ValueError("num_iterations > 1 needs loss_mode='clip' (the SC loss has no old-policy ratio)")
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._opt_step = 0
        self.phase_ms = defaultdict(float)
        self._timers = []
        self.total_rollout_tokens = 0
        if args.deepspeed and self.is_main:
            print("[iadr1-b200] note: --deepspeed is accepted for script compatibility and ignored (plain data parallel)",
                  file=sys.stderr)
        if self.is_main and (args.gradient_checkpointing or self.model.recompute):
            print(f"[iadr1-b200] activation recompute: {'on' if self.model.recompute else 'off (every activation stays resident)'}"
                  f"; Adam moments: {str(self.params.exp_avg.dtype).replace('torch.', '')}", file=sys.stderr)

    # ---------------------------------------------------------------------------------------------------------------
    # prompt encoding + rollout
    # ---------------------------------------------------------------------------------------------------------------
    def _encode_prompt(self, example: dict) -> dict:
        from PIL import Image
        text = maybe_apply_chat_template(example, self.processing_class)["prompt"]
        images = []
        img = example.get("image")
        if img is not None:
            for im in (img if isinstance(img, list) else [img]):
                images.append(Image.open(im) if isinstance(im, str) else im)
        gpu_pv = None
        if images and self.args.gpu_image_preprocess and self.cfg.family in ("qwen2_vl", "qwen2_5_vl"):
            # image half of the processor call on the GPU: the host decodes to uint8 and sizes the grid; the
            # text half (placeholder expansion + tokenisation) is what the processor does with `images=None`
            from .preprocess import OPENAI_CLIP_MEAN, OPENAI_CLIP_STD, qwen_preprocess_gpu
            ip = getattr(self.processing_class, "image_processor", None)
            size = getattr(ip, "size", None) or {}
            minp = getattr(ip, "min_pixels", None) or size.get("shortest_edge") or self.min_pixels
            maxp = getattr(ip, "max_pixels", None) or size.get("longest_edge") or self.max_pixels
            gpu_pv, gpu_grid = qwen_preprocess_gpu(images, self.cfg.vision, self.device, int(minp), int(maxp),
                                                   tuple(getattr(ip, "image_mean", None) or OPENAI_CLIP_MEAN),
                                                   tuple(getattr(ip, "image_std", None) or OPENAI_CLIP_STD))
            tok = getattr(self.processing_class, "image_token", "<|image_pad|>")
            merge2 = self.cfg.vision.spatial_merge_size ** 2
            for g_ in gpu_grid:
                text = text.replace(tok, "<|iadr1_placeholder|>" * (g_[1] * g_[2] // merge2), 1)
            text = text.replace("<|iadr1_placeholder|>", tok)
            images = []
        enc = self.processing_class(text=[text], images=images if images else None, return_tensors="pt", padding=True,
                                    padding_side="left", add_special_tokens=False)
        ids = enc["input_ids"][0]
        if "attention_mask" in enc:
            ids = ids[enc["attention_mask"][0].bool()]
        if self.max_prompt_length is not None:
            ids = ids[-self.max_prompt_length:]  # Q14: ids only; cutting into image tokens raises downstream
        pv, grid = (gpu_pv, gpu_grid) if gpu_pv is not None else vision_inputs_from_processor(self.cfg, enc)
        if pv is not None and not pv.is_cuda:
            pv = pv.pin_memory().to(self.device, non_blocking=True)   # pinned staging -> async H2D
        return dict(input_ids=ids.numpy().astype(np.int64), pixel_values=pv, grid_thw=grid, text=text)

    def _engine_for(self, n_groups: int, p_len: int) -> RolloutEngine:
        e = self._engine
        p_need = (p_len + 63) // 64 * 64
        if e is None or e.n_groups < n_groups or e.p_max < p_need:
            a = self.args
            self._engine = e = RolloutEngine(self.model, n_groups, self.num_generations, max(p_need, e.p_max if e else 0),
                                             self.max_completion_length, temperature=a.temperature, top_k=a.rollout_top_k,
                                             top_p=a.rollout_top_p, forbid_eos=a.rollout_forbid_eos)
        return e

    def _rollout(self, encoded: list, image_embeds=None) -> list:
        """Sample G completions for each encoded prompt in ONE decode batch. Returns a list of [G, C] int32 tensors."""
        eng = self._engine_for(len(encoded), max(len(e["input_ids"]) for e in encoded))
        seed = self.args.rollout_seed if self.args.rollout_seed is not None else self.args.seed
        seed = (seed * 1000003 + self.state.global_step * 8191 + self.rank * 131 + self._rollout_calls) & 0x7FFFFFFF
        self._rollout_calls += 1
        with self._phase("rollout"):
            out, stats = eng.generate(encoded, seed=seed, image_embeds=image_embeds)
        G = self.num_generations
        self.total_rollout_tokens += out.numel()
        return [out[i * G:(i + 1) * G] for i in range(len(encoded))]

    _rollout_calls = 0
    _window = None           # per-window vision state (features, saved activations, gradient accumulator)
    _iteration = 0           # policy iteration inside the current generated window (num_iterations > 1)

    def prepare_window(self, examples: list, encoded: Optional[list] = None):
        """Batched-rollout mode: roll out every group of the coming accumulation window together (the weights they are
        sampled from are the weights at the start of the optimizer step, as in the reference: Q12).

        The vision tower also runs ONCE per window here: policy features of all the window's images in one pass (saved for
        the backward; the same features feed the rollout prefill and every scoring pass - the reference runs the tower
        G times per forward on tiled pixels, :624-628), reference-model features in one no-grad pass. The gradient of the
        policy features is accumulated over the window's passes and pushed through the tower once, before the optimizer
        step (`_flush_window_vision`)."""
        enc = [self._encode_prompt(ex) for ex in examples] if encoded is None else encoded
        self._window = None
        img = None
        if self.args.window_vision and all(e["pixel_values"] is not None for e in enc):
            from .geometry import image_token_count
            H = self.cfg.text.hidden_size
            slices, off = {}, 0
            for ex, e in zip(examples, enc):
                n = sum(image_token_count(self.cfg, g_) for g_ in e["grid_thw"])
                slices[id(ex)] = (off, off + n)
                off += n
            img = torch.empty(off, H, dtype=torch.bfloat16, device=self.device)
            ref_img = torch.empty_like(img) if self.ref_model is not None else None
            # one tower pass per set of prompts with IDENTICAL image grids: equal-length attention segments keep the
            # batched (un-masked) attention path, and mixed image sizes never build one [Np_total x Np_total] score matrix
            by_grid = {}
            for i, e in enumerate(enc):
                by_grid.setdefault(tuple(tuple(int(x) for x in g_) for g_ in e["grid_thw"]), []).append(i)
            parts = []
            with self._phase("vision_fwd"):
                for idxs in by_grid.values():
                    pv = torch.cat([enc[i]["pixel_values"].to(self.device) for i in idxs], 0)
                    grids = [g_ for i in idxs for g_ in enc[i]["grid_thw"]]
                    rng = [slices[id(examples[i])] for i in idxs]
                    f, vctx = self.model.vision_forward(pv, grids, save=True)
                    o = 0
                    for lo, hi in rng:
                        img[lo:hi] = f[o:o + hi - lo]
                        o += hi - lo
                    if ref_img is not None:
                        with torch.no_grad():
                            rf, _ = self.ref_model.vision_forward(pv, grids, save=False)
                        o = 0
                        for lo, hi in rng:
                            ref_img[lo:hi] = rf[o:o + hi - lo]
                            o += hi - lo
                    parts.append((vctx, rng))
            self._window = dict(img=img, ref_img=ref_img, vctx=parts, slices=slices,
                                dimg=torch.zeros(off, H, dtype=torch.float32, device=self.device))
        comps = self._rollout(enc, image_embeds=img)
        for ex, e, c in zip(examples, enc, comps):
            self._rollout_cache[id(ex)] = (e, c)
            self._host_completions(ex, c)      # the rollout has just synchronised: a free host copy for the reward callbacks

    def _window_features(self, examples: list):
        """(policy image embeds, reference image embeds, gradient sink) for a micro-batch whose groups all belong to the
        current window, else (None, None, None) -> the pass runs its own vision tower."""
        w = self._window
        if w is None or w["vctx"] is None or any(id(ex) not in w["slices"] for ex in examples):
            return None, None, None
        rng = [w["slices"][id(ex)] for ex in examples]
        img = torch.cat([w["img"][lo:hi] for lo, hi in rng], 0)
        ref = torch.cat([w["ref_img"][lo:hi] for lo, hi in rng], 0) if w["ref_img"] is not None else None

        def sink(dimg32):
            o = 0
            for lo, hi in rng:
                w["dimg"][lo:hi] += dimg32[o:o + hi - lo]
                o += hi - lo
        return img, ref, sink

    def _flush_window_vision(self):
        """Vision-tower backward of the whole window (once), then the window's features are stale (weights change)."""
        w = self._window
        if w is not None and w["vctx"] is not None:
            from . import ops
            with self._phase("backward"):
                d = ops.cast_f32_bf16(w["dimg"])
                for vctx, rng in w["vctx"]:
                    dpart = d if len(w["vctx"]) == 1 and len(rng) == len(w["slices"]) else torch.cat([d[lo:hi] for lo, hi in rng], 0)
                    self.model.vision_backward(dpart, vctx)
        self._window = None

    def optimizer_step(self):
        self._flush_window_vision()
        super().optimizer_step()

    # ---------------------------------------------------------------------------------------------------------------
    # the hot path: one micro-step (ref: sc_grpo_trainer.py:586-819)
    # ---------------------------------------------------------------------------------------------------------------
    def compute_loss(self, model=None, inputs=None, return_outputs=False, num_items_in_batch=None):
        """`inputs`: the micro-batch, a list of `per_device_train_batch_size` examples (the reference uses 1). Each example
        is its own GRPO group (advantages are normalised inside the group); with the shared-prefix layout all groups of
        the micro-batch are packed into ONE forward/backward pass. Returns the mean of the group losses."""
        if return_outputs:
            raise ValueError("The GRPOTrainer does not support returning outputs")
        G, dev, a = self.num_generations, self.device, self.args
        items = []
        keep = self._iteration + 1 < max(1, a.num_iterations)   # the window is re-used for the next policy iteration
        for example in inputs:
            cached = self._rollout_cache.get(id(example)) if keep else self._rollout_cache.pop(id(example), None)
            if cached is None:
                enc = self._encode_prompt(example)
                completion_ids = self._rollout([enc])[0]
                cached = (enc, completion_ids)
                if keep:
                    self._rollout_cache[id(example)] = cached
            enc, completion_ids = cached[0], cached[1]
            items.append((example, enc, completion_ids))
        # host copies of the completions BEFORE the scoring passes are enqueued: the reward callbacks (CPU Python) then run while
        # the GPU executes the policy / reference forwards instead of waiting for them
        host_ids = {id(ex): self._host_completions(ex, c) for ex, _, c in items}
        temp = a.temperature if a.loss_mode == "clip" else 1.0   # Q1: SC mode does not temperature-scale the logits
        if a.shared_prefix:
            batch = self.model.prepare_groups([dict(prompt_ids=enc["input_ids"], completion_ids=c, pixel_values=enc["pixel_values"],
                                                    grid_thw=enc["grid_thw"]) for _, enc, c in items])
            rows, labels, slices = batch["sel_index"], batch["labels"], batch["group_slices"]
            logps_all, ref_all = self._score(batch, rows, labels, temp, self._window_features([ex for ex, _, _ in items]))
            per_group = [(logps_all[lo:hi], None if ref_all is None else ref_all[lo:hi]) for lo, hi in slices]
        else:
            per_group = []
            for _, enc, c in items:          # reference layout: the full [G, P + C] batch, one group per pass
                P, C = len(enc["input_ids"]), c.shape[1]
                T = P + C
                ids = torch.cat([torch.from_numpy(enc["input_ids"]).to(dev)[None, :].expand(G, -1), c.long()], 1)  # :681-683
                batch = self.model.prepare_batch(ids, enc["pixel_values"], enc["grid_thw"], prompt_len=P)
                rows = (torch.arange(G, device=dev)[:, None] * T + (P - 1) + torch.arange(C, device=dev)[None, :]).reshape(-1).to(torch.int32)
                per_group.append(self._score(batch, rows, c.reshape(-1).to(torch.int32).contiguous(), temp))
        losses = [self._group_loss(ex, c, lp.view(G, -1), None if rf is None else rf.view(G, -1), host_ids[id(ex)])
                  for (ex, _, c), (lp, rf) in zip(items, per_group)]
        return torch.stack(losses).mean()

    def _score(self, batch, rows, labels, temp, window_feats=(None, None, None)):
        """Policy log-probs (with the CUDA backward attached) and reference log-probs (no grad), :733-743."""
        img, ref_img, sink = window_feats
        anchor = torch.zeros((), device=self.device, requires_grad=True)
        with self._phase("policy_fwd"):
            logps = _LogProbFn.apply(anchor, self.model, batch, rows, labels, temp, img, sink)
        ref_logps = None
        if self.ref_model is not None:
            with torch.no_grad(), self._phase("ref_fwd"):
                ref_logps, _ = self.ref_model.logprobs_forward(batch, rows, labels, temp, save=False, image_embeds=ref_img)
        return logps, ref_logps

    def _host_completions(self, example, completion_ids: torch.Tensor) -> torch.Tensor:
        """CPU copy of a group's completion ids, made once per rollout (cached with it)."""
        cache = self.__dict__.setdefault("_host_ids", {})
        key = (id(example), completion_ids.data_ptr())
        if key not in cache:
            if len(cache) > 256:
                cache.clear()
            cache[key] = completion_ids.cpu()
        return cache[key]

    def _group_loss(self, example: dict, completion_ids: torch.Tensor, logps: torch.Tensor, ref_logps, host_ids=None) -> torch.Tensor:
        """Mask, rewards, advantages and loss of ONE group from its [G, C] log-probs (ref: :722-726, :746-798)."""
        G, dev, a = self.num_generations, self.device, self.args
        mask = grpo_loss.completion_mask(completion_ids.long(), self.processing_class.eos_token_id)          # :722-726
        if a.mask_truncated_completions and a.loss_mode == "clip":
            mask = grpo_loss.mask_truncated(mask, completion_ids.long(), self.processing_class.eos_token_id)
        # multi-iteration GRPO (trl grpo_trainer.py:872-901, 1066-1089): the first pass over a window records the
        # log-probs the completions were sampled under; later passes clip the ratio against them
        old_logps = None
        if a.loss_mode == "clip" and a.num_iterations > 1:
            if self._iteration == 0:
                self._old_logps[id(example)] = logps.detach().clone()
            old_logps = self._old_logps[id(example)]
            if self._iteration + 1 >= a.num_iterations:
                self._old_logps.pop(id(example), None)
        # ---- rewards on decoded text (CPU Python callbacks, verbatim convention :749-781) ----
        with self._phase("rewards"):
            texts = self.processing_class.batch_decode(completion_ids.cpu() if host_ids is None else host_ids, skip_special_tokens=True)
            rewards_per_func = call_reward_funcs(self.reward_funcs, example, texts, self.state.global_step,
                                                  getattr(self, "reward_processing_classes", None)).to(dev)
        rw = torch.tensor(a.reward_weights, device=dev) if (a.reward_weights and a.loss_mode == "clip") else None
        adv, rewards, std = grpo_loss.group_advantages(rewards_per_func, G, a.scale_rewards or a.loss_mode == "sc", rw)  # :784-793
        if a.loss_mode == "sc":
            loss, mean_kl = grpo_loss.sc_grpo_loss(logps, ref_logps, adv, mask, self.beta)       # :796-798
        else:
            loss, mean_kl = grpo_loss.clip_grpo_loss(logps, old_logps, ref_logps, adv, mask, self.beta, a.epsilon,
                                                     a.epsilon_high if a.epsilon_high is not None else a.epsilon,
                                                     a.loss_type, self.max_completion_length)
        # ---- metrics (:801-817); kept as device scalars, reduced across ranks at log time ----
        m = self._metrics
        m["completion_length"].append(mask.sum(1).float().mean().detach())
        for i, rf in enumerate(self.reward_funcs):
            # :806-810: a reward model is named after its checkpoint directory, a callback after the function
            name = rf.config._name_or_path.split("/")[-1] if _is_reward_model(rf) else rf.__name__
            m[f"rewards/{name}"].append(rewards_per_func[:, i].mean().detach())
        m["reward"].append(rewards.mean().detach())
        m["reward_std"].append(std.mean().detach())
        m["kl"].append(mean_kl.detach())
        return loss

    def training_step(self, inputs: list) -> torch.Tensor:
        loss = self.compute_loss(self.model, inputs)
        GA = self.args.gradient_accumulation_steps
        self.arm_overlap(self._micro_idx == GA - 1)      # last micro-step: layers' gradients go to NCCL as they retire
        self._micro_idx += 1
        with self._phase("backward"):
            (loss / GA).backward()
        return loss.detach()

    def train(self, resume_from_checkpoint=None):
        a = self.args
        bs, GA = a.per_device_train_batch_size, a.gradient_accumulation_steps
        per_rank = len(self.train_dataset) // self.world
        steps_per_epoch = max(1, per_rank // (bs * GA))
        self.state.max_steps = a.max_steps if a.max_steps > 0 else int(math.ceil(a.num_train_epochs * steps_per_epoch))
        if per_rank < bs * GA:
            raise ValueError(f"this rank's shard holds {per_rank} examples, one optimizer step needs "
                             f"per_device_train_batch_size * gradient_accumulation_steps = {bs * GA}")
        t_start = time.time()
        epoch = 0
        tr_loss = []
        skip_windows = 0
        resume = resume_from_checkpoint or a.resume_from_checkpoint
        if resume:
            # continue where the checkpoint stopped: same epoch order (seeded), the windows already consumed are skipped
            self.load_checkpoint(resume)
            n_iter0 = max(1, a.num_iterations) if a.loss_mode == "clip" else 1
            windows_done = self.state.global_step // n_iter0
            epoch, skip_windows = windows_done // steps_per_epoch, windows_done % steps_per_epoch
        while self.state.global_step < self.state.max_steps:
            order = self._epoch_order(epoch)
            micro = [order[i:i + bs] for i in range(0, len(order) - bs + 1, bs)]
            for w in range(skip_windows * GA, len(micro) - GA + 1, GA):
                window = [[self.train_dataset[j] for j in mb] for mb in micro[w:w + GA]]
                if a.batched_rollout:
                    self.prepare_window([ex for mb in window for ex in mb])
                # num_iterations optimizer steps per generated window (vendored TRL `num_iterations`, clip mode only)
                n_iter = max(1, a.num_iterations) if a.loss_mode == "clip" else 1
                for it in range(n_iter):
                    self._iteration = it
                    for mb in window:
                        tr_loss.append(self.training_step(mb))
                    self.optimizer_step()
                    if self.state.global_step >= self.state.max_steps:
                        break
                self._iteration = 0
                self._rollout_cache.clear()
                self._old_logps.clear()
                self.state.epoch = epoch + (w + GA) / max(1, len(micro))
                if a.logging_steps and self.state.global_step % max(1, int(a.logging_steps)) == 0:
                    loss_val = torch.stack(tr_loss).mean().item()
                    tr_loss = []
                    self.log({"loss": loss_val, "grad_norm": float(self._grad_norm_dev.item()),
                              "learning_rate": self._last_lr, "epoch": round(self.state.epoch, 4)})
                if a.save_strategy == "steps" and a.save_steps and self.state.global_step % max(1, int(a.save_steps)) == 0:
                    self.save_checkpoint(os.path.join(a.output_dir, f"checkpoint-{self.state.global_step}"))
                if self.state.global_step >= self.state.max_steps:
                    break
            epoch += 1
            skip_windows = 0
        self.flush_timers()
        runtime = time.time() - t_start
        self.state.log_history.append({"train_runtime": runtime, "step": self.state.global_step})
        return {"global_step": self.state.global_step, "train_runtime": runtime}

