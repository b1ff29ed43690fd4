"""PA-SFT on the B200 path: teacher-forced cross-entropy on assistant tokens with the SAME kernels as SC-GRPO scoring
(vision tower, decoder, fused lm_head log-softmax, hand-written backward, flat AdamW).

Stands in for ref: train/stage_sft/llamafactory/train/sft/workflow.py:40-112 (`run_sft`) +
sft/trainer.py:46-107 (`CustomSeq2SeqTrainer`) + HF `ForCausalLMLoss` ($HF/loss/loss_utils.py:45-67): logits in fp32,
labels shifted by one, ignore_index = -100, token-mean over the accumulation window. Only the model forward/backward of
the SFT stage is in scope; the LLaMA-Factory data pipeline is restated minimally here
(sharegpt `messages` + `images`, chat template from the processor, labels on assistant turns only,
`cutoff_len`, pre-shrink above 512^2 px with NEAREST - quirk Q16).

Freeze quirk Q17 is preserved: with the LLaMA-Factory defaults the vision tower and merger are frozen ONLY for
`model_type == "qwen2_vl"`; Qwen2.5-VL trains everything.
"""
from __future__ import annotations

import json
import math
import os
import time
from collections import defaultdict
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch

from . import lib as L
from .checkpoint import load_pretrained
from .config import VLMConfig
from .model import VLM
from .geometry import vision_inputs_from_processor
from .params import ParamStore
from .trainer_base import TrainerCore, TrainerState

IGNORE_INDEX = -100


@dataclass
class SFTArguments:
    """The flags `scripts/train/PA_SFT/*.sh` pass to `train/stage_sft/train.py` (:25-50); unknown flags raise, as in
    LLaMA-Factory (hparams/parser.py:81)."""
    model_name_or_path: Optional[str] = None
    output_dir: str = "sft_output"
    stage: str = "sft"
    do_train: bool = False
    dataset: Optional[str] = None
    dataset_dir: str = "data"
    image_dir: Optional[str] = None
    template: Optional[str] = None
    finetuning_type: str = "full"
    overwrite_cache: bool = False
    overwrite_output_dir: bool = False
    deepspeed: Optional[str] = None            # ignored (plain data parallel)
    warmup_steps: int = 0
    warmup_ratio: float = 0.0
    weight_decay: float = 0.0
    per_device_train_batch_size: int = 1
    gradient_accumulation_steps: int = 1
    ddp_timeout: int = 1800
    learning_rate: float = 5e-5
    lr_scheduler_type: str = "linear"
    logging_steps: float = 500
    cutoff_len: int = 2048
    save_steps: float = 500
    save_strategy: str = "steps"
    plot_loss: bool = False
    num_train_epochs: float = 3.0
    max_steps: int = -1
    bf16: bool = False
    fp16: bool = False
    seed: int = 42
    data_seed: Optional[int] = None
    max_grad_norm: float = 1.0
    adam_beta1: float = 0.9
    adam_beta2: float = 0.999
    adam_epsilon: float = 1e-8
    freeze_vision_tower: bool = True           # finetuning_args.py:416-423 defaults
    freeze_multi_modal_projector: bool = True
    image_resolution: int = 512 * 512          # hparams/model_args.py:61-64
    shuffle_dataset: bool = True
    resume_from_checkpoint: Optional[str] = None
    report_to: Optional[str] = "none"

    def __post_init__(self):
        if self.stage != "sft":
            raise ValueError("only --stage sft is on the B200 path (pt/rm/ppo/dpo/kto are out of scope)")
        if self.finetuning_type != "full":
            raise ValueError("only --finetuning_type full is supported")
        if self.fp16:
            raise ValueError("fp16 is not supported (bf16 parameters, fp32 master weights)")


def load_sharegpt_dataset(name: str, dataset_dir: str = "data") -> list:
    """`--dataset NAME` is looked up in <dataset_dir>/dataset_info.json relative to the CWD (data_args.py:40-41,
    data/parser.py:87); the entry's `file_name` holds sharegpt rows {messages, images} (README.md:71-100)."""
    info_path = os.path.join(dataset_dir, "dataset_info.json")
    with open(info_path) as f:
        info = json.load(f)
    if name not in info:
        raise ValueError(f"Undefined dataset {name} in {info_path}.")
    fn = info[name]["file_name"]
    if not os.path.isabs(fn):
        fn = os.path.join(dataset_dir, fn)
    with open(fn) as f:
        return json.load(f)


from .sft_data import infer_seqlen  # noqa: E402,F401  (re-exported: tests pin it with the reference function's outputs)


def encode_supervised_example(example: dict, processor, cutoff_len: int, image_dir: Optional[str], image_resolution: int,
                              cfg=None, template: Optional[str] = None, device=None, gpu_preprocess: bool = False):
    """One sharegpt row -> (input_ids [T], labels [T] with IGNORE_INDEX outside assistant answers, pixel_values, grid_thw)
    through the native data path (sft_data.py): template -> `<image>` expansion -> per-turn token pairs -> truncation.
    Multi-turn and multi-image rows are supported; the template defaults to the model family's
    (ref: scripts/train/PA_SFT/*.sh:33)."""
    from . import sft_data as D
    from .geometry import image_token_count
    family = cfg.family if cfg is not None else ("llava_onevision" if getattr(processor, "llava", False) else "qwen2_5_vl")
    if cfg is None and getattr(processor, "llava15", False):
        family = "llava"
    tpl = D.get_template(template, family)
    images = D.load_images(example, image_dir, image_resolution, tpl.plugin)
    pv, grid, seqlens = None, None, []
    if images:
        ip = processor.image_processor
        if tpl.plugin == "qwen2_vl":
            merge = getattr(ip, "merge_size", 2)
            if gpu_preprocess and cfg is not None and device is not None:
                from .preprocess import OPENAI_CLIP_MEAN, OPENAI_CLIP_STD, qwen_preprocess_gpu
                size = getattr(ip, "size", None) or {}
                pv, grid = qwen_preprocess_gpu(images, cfg.vision, device,
                                               int(getattr(ip, "min_pixels", None) or size.get("shortest_edge") or 3136),
                                               int(getattr(ip, "max_pixels", None) or size.get("longest_edge") or 12845056),
                                               tuple(getattr(ip, "image_mean", None) or OPENAI_CLIP_MEAN),
                                               tuple(getattr(ip, "image_std", None) or OPENAI_CLIP_STD))
            else:
                im = ip(images=images, return_tensors="pt")
                pv, grid = im["pixel_values"], im["image_grid_thw"].tolist()
            seqlens = [g[0] * g[1] * g[2] // merge ** 2 for g in grid]
        else:
            if cfg is None:
                raise ValueError("the llava_next data path needs the model config (anyres token counts)")
            im = ip(images=images, return_tensors="pt")
            pv, grid = vision_inputs_from_processor(cfg, im)
            seqlens = [image_token_count(cfg, g) for g in grid]
    msgs = D.expand_image_placeholders(example["messages"], seqlens, tpl)
    eos = getattr(getattr(processor, "tokenizer", processor), "eos_token", None) or "</s>"
    pair_ids = [(D.tokenize(processor, p), D.tokenize(processor, r)) for p, r in D.render_pairs(msgs, tpl, eos)]
    ids, labels = D.encode_pairs(pair_ids, cutoff_len)
    return dict(input_ids=np.asarray(ids, dtype=np.int64), labels=np.asarray(labels, dtype=np.int64), pixel_values=pv, grid_thw=grid)


class PASFTTrainer(TrainerCore):
    def __init__(self, model, args: SFTArguments, train_dataset=None, processing_class=None):
        self.args = args
        self.state = TrainerState()
        self._metrics = defaultdict(list)
        self._setup_distributed()
        torch.manual_seed(args.seed)
        if isinstance(model, str):
            from .checkpoint import load_config
            mdt = ParamStore.plan_moment_dtype(load_config(model), self.device, False, "auto")
            self.cfg, self.params = load_pretrained(model, self.device, moment_dtype=mdt)
        elif isinstance(model, VLMConfig):
            self.cfg = model
            self.params = ParamStore(model, self.device, with_grads=True, with_optimizer=True,
                                     moment_dtype=ParamStore.plan_moment_dtype(model, self.device, False, "auto"))
            self.params.init_random(seed=args.seed)
        else:
            self.cfg, self.params = model.cfg, model
        self.model = VLM(self.cfg, self.params)
        self.model.recompute = self.params.exp_avg.dtype == torch.bfloat16   # models whose state fills the device (7B)
        if processing_class is None:
            from transformers import AutoProcessor
            processing_class = AutoProcessor.from_pretrained(model)
        self.processing_class = processing_class
        self.train_dataset = train_dataset
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._opt_step = 0
        self.phase_ms = defaultdict(float)
        self._timers = []
        # Q17: LLaMA-Factory registers only "qwen2_vl" as a composite model, so only there the freeze flags bite
        self.freeze_vision = self.cfg.family == "qwen2_vl" and (args.freeze_vision_tower or args.freeze_multi_modal_projector)

    def _trainable_ranges(self):
        ps, a = self.params, self.args
        if not self.freeze_vision:
            return [(0, ps.n_decay, a.weight_decay), (ps.n_decay, ps.numel, 0.0)]
        # the flat buffer is [decayed: visual..., text...] [non-decayed: visual..., text...]; skip both visual runs
        def end_of_visual(lo, hi):
            offs = sorted((o, n) for n, o in ps.offsets.items() if lo <= o < hi)
            e = lo
            for o, n in offs:
                if n.startswith("visual."):
                    numel = int(np.prod(ps.shapes[n]))
                    e = max(e, o + (numel + 7) // 8 * 8)
            return e
        return [(end_of_visual(0, ps.n_decay), ps.n_decay, a.weight_decay),
                (end_of_visual(ps.n_decay, ps.numel), ps.numel, 0.0)]

    # ---------------------------------------------------------------------------------------------------------------
    def compute_loss_and_backward(self, enc: dict, n_items_window: int) -> torch.Tensor:
        """One sequence: CE summed over its labelled tokens / n_items_window, gradient accumulated in fp32."""
        ids, labels = enc["input_ids"], enc["labels"]
        T = len(ids)
        batch = self.model.prepare_batch(ids[None, :], enc["pixel_values"], enc["grid_thw"])
        tgt = np.nonzero(labels[1:] != IGNORE_INDEX)[0]           # logits at t predict token t + 1
        if len(tgt) == 0:
            return torch.zeros((), device=self.device)
        rows = torch.from_numpy(tgt.astype(np.int32)).to(self.device)
        lab = torch.from_numpy(labels[1:][tgt].astype(np.int32)).to(self.device)
        with self._phase("forward"):
            logp, ctx = self.model.logprobs_forward(batch, rows, lab, 1.0, save=True)
        loss = -logp.sum() / n_items_window
        with self._phase("backward"):
            dlogp = torch.full_like(logp, -1.0 / n_items_window)
            if self.freeze_vision:
                ctx["vctx"] = None      # no gradient flows into the frozen tower
            self.model.logprobs_backward(dlogp, ctx)
        self.state.num_input_tokens_seen += T
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
        t0, epoch, losses = time.time(), 0, []
        skip = 0
        resume = resume_from_checkpoint or a.resume_from_checkpoint
        if resume:
            self.load_checkpoint(resume)
            epoch, skip = self.state.global_step // steps_per_epoch, self.state.global_step % steps_per_epoch
        while self.state.global_step < self.state.max_steps:
            order = self._epoch_order(epoch)
            win = bs * GA
            for w in range(skip * win, len(order) - win + 1, win):
                encs = [encode_supervised_example(self.train_dataset[j], self.processing_class, a.cutoff_len, a.image_dir,
                                                  a.image_resolution, cfg=self.cfg, template=a.template, device=self.device,
                                                  gpu_preprocess=True) for j in order[w:w + win]]
                n_items = max(1, sum(int((e["labels"][1:] != IGNORE_INDEX).sum()) for e in encs))
                step_loss = torch.zeros((), device=self.device)
                for e in encs:
                    step_loss = step_loss + self.compute_loss_and_backward(e, n_items)
                self.optimizer_step()
                losses.append(step_loss)
                self.state.epoch = epoch + (w + win) / max(1, len(order))
                if a.logging_steps and self.state.global_step % max(1, int(a.logging_steps)) == 0:
                    self.log({"loss": torch.stack(losses).mean().item(), "grad_norm": float(self._grad_norm_dev.item()),
                              "learning_rate": self._last_lr, "epoch": round(self.state.epoch, 4)})
                    losses = []
                if a.save_strategy == "steps" and a.save_steps and self.state.global_step % max(1, int(a.save_steps)) == 0:
                    self.save_checkpoint(os.path.join(a.output_dir, f"checkpoint-{self.state.global_step}"))
                if self.state.global_step >= self.state.max_steps:
                    break
            epoch += 1
            skip = 0
        self.flush_timers()
        return {"global_step": self.state.global_step, "train_runtime": time.time() - t0}


def run_sft(argv=None):
    """Entry used by train/stage_sft/train.py (ref: llamafactory/train/tuner.py:79 `run_exp`)."""
    from transformers import HfArgumentParser
    (args,) = HfArgumentParser(SFTArguments).parse_args_into_dataclasses(args=argv)   # unknown flags raise ValueError
    if not args.do_train:
        raise ValueError("only --do_train is supported")
    data = load_sharegpt_dataset(args.dataset, args.dataset_dir)
    trainer = PASFTTrainer(args.model_name_or_path, args, train_dataset=data)
    trainer.train()
    trainer.save_model(args.output_dir)
    return trainer
