"""Shared machinery of the B200 trainers (SC-GRPO and PA-SFT): one process per GPU, flat-buffer optimizer step
(gradient all-reduce = the only data-path collective -> device-side global-norm clip -> fused AdamW over the trainable
ranges), LR schedules of the HF Trainer (`linear`, `cosine`, `constant`), seed-42 shuffling sharded per rank, metric
averaging, HF-layout checkpoints. Replaces what the reference gets from transformers.Trainer + accelerate + DeepSpeed
(ref: train/stage_rl/trainer/sc_grpo_trainer.py:281; scripts/train/zero3.json)."""
from __future__ import annotations

import json
import math
import os
import warnings
from collections import defaultdict
from typing import Optional

import torch

from . import lib as L
from .checkpoint import save_pretrained


class TrainerState:
    def __init__(self):
        self.global_step = 0
        self.epoch = 0.0
        self.max_steps = 0
        self.log_history: list = []
        self.num_input_tokens_seen = 0


class TrainerCore:
    # ---------------------------------------------------------------------------------------------------------------
    def _setup_distributed(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        local = int(os.environ.get("LOCAL_RANK", "0"))
        self.is_main = self.rank == 0
        if not torch.cuda.is_available():
            raise L.NativeLibraryError("SCGRPOTrainer needs a CUDA device: the hot path has no CPU fallback")
        torch.cuda.set_device(local)
        self.device = torch.device("cuda", local)
        if self.world > 1 and not torch.distributed.is_initialized():
            import datetime
            torch.distributed.init_process_group("nccl", timeout=datetime.timedelta(seconds=self.args.ddp_timeout),
                                                 device_id=self.device)
        L.lib()

    def _phase(self, name):
        trainer = self

        class _T:
            def __enter__(self_inner):
                self_inner.e0 = torch.cuda.Event(enable_timing=True)
                self_inner.e1 = torch.cuda.Event(enable_timing=True)
                self_inner.e0.record()

            def __exit__(self_inner, *a):
                self_inner.e1.record()
                trainer._timers.append((name, self_inner.e0, self_inner.e1))

        return _T()

    def flush_timers(self):
        torch.cuda.synchronize()
        for name, e0, e1 in self._timers:
            self.phase_ms[name] += e0.elapsed_time(e1)
        self._timers = []

    # ---------------------------------------------------------------------------------------------------------------
    # optimizer step (backward already accumulated fp32 grads)
    # ---------------------------------------------------------------------------------------------------------------
    def _lr_at(self, step: int) -> float:
        a, total = self.args, max(1, self.state.max_steps)
        warm = a.warmup_steps if a.warmup_steps > 0 else int(math.ceil(total * a.warmup_ratio))
        if step < warm:
            return a.learning_rate * step / max(1, warm)
        prog = (step - warm) / max(1, total - warm)
        if a.lr_scheduler_type == "linear":
            return a.learning_rate * max(0.0, 1.0 - prog)
        if a.lr_scheduler_type == "cosine":
            return a.learning_rate * max(0.0, 0.5 * (1.0 + math.cos(math.pi * min(1.0, prog))))
        if a.lr_scheduler_type in ("constant", "constant_with_warmup"):
            return a.learning_rate
        raise ValueError(f"unsupported lr_scheduler_type {a.lr_scheduler_type}")

    def _trainable_ranges(self):
        """[(lo, hi, weight_decay)] over the flat buffers; subclasses exclude frozen tensors."""
        ps, a = self.params, self.args
        return [(0, ps.n_decay, a.weight_decay), (ps.n_decay, ps.numel, 0.0)]

    def optimizer_step(self):
        ps, a, lib = self.params, self.args, L.lib()
        s = L.stream_ptr()
        with self._phase("allreduce"):
            if self.world > 1:
                # the ONLY data-path collective: sum of the flat fp32 gradient over NVLink
                torch.distributed.all_reduce(ps.grad_flat)
        with self._phase("optimizer"):
            scale = 1.0 / self.world
            self._sumsq.zero_()
            L.check(lib.iadr1_sumsq_f32(ps.grad_flat.data_ptr(), ps.numel, self._sumsq.data_ptr(), s), "sumsq")
            self._opt_step += 1
            lr = self._lr_at(self.state.global_step)
            self._last_lr = lr
            for lo, hi, wd in self._trainable_ranges():
                if hi <= lo:
                    continue
                if ps.exp_avg.dtype == torch.bfloat16:
                    L.check(lib.iadr1_adamw_step_bf16m(
                        ps.master[lo:].data_ptr(), ps.flat[lo:].data_ptr(), ps.grad_flat[lo:].data_ptr(),
                        ps.exp_avg[lo:].data_ptr(), ps.exp_avg_sq[lo:].data_ptr(), hi - lo, lr, a.adam_beta1, a.adam_beta2,
                        a.adam_epsilon, wd, self._opt_step, scale, self._sumsq.data_ptr(),
                        a.max_grad_norm if a.max_grad_norm else 0.0, 1, (a.seed << 20) + lo, s), "adamw_step_bf16m")
                    continue
                L.check(lib.iadr1_adamw_step(ps.master[lo:].data_ptr(), ps.flat[lo:].data_ptr(), ps.grad_flat[lo:].data_ptr(),
                                             ps.exp_avg[lo:].data_ptr(), ps.exp_avg_sq[lo:].data_ptr(), hi - lo, lr,
                                             a.adam_beta1, a.adam_beta2, a.adam_epsilon, wd, self._opt_step, scale,
                                             self._sumsq.data_ptr(), a.max_grad_norm if a.max_grad_norm else 0.0, 1, s),
                        "adamw_step")
            self._grad_norm_dev = self._sumsq.sqrt() * scale
        self.state.global_step += 1

    # ---------------------------------------------------------------------------------------------------------------
    def _epoch_order(self, epoch: int):
        n = len(self.train_dataset)
        if self.args.shuffle_dataset:
            g = torch.Generator().manual_seed((self.args.data_seed if self.args.data_seed is not None else self.args.seed) + epoch)
            order = torch.randperm(n, generator=g).tolist()
        else:
            order = list(range(n))
        return order[self.rank::self.world] if self.world > 1 else order

    def log(self, logs: dict, start_time: Optional[float] = None) -> None:
        """Average the per-micro-step metrics (and across ranks, as `gather_for_metrics(...).mean()` did), :821-827."""
        keys = sorted(self._metrics.keys())
        if keys:
            vec = torch.stack([torch.stack([torch.as_tensor(v, device=self.device, dtype=torch.float32) for v in self._metrics[k]]).mean()
                               for k in keys])
            if self.world > 1:
                torch.distributed.all_reduce(vec)
                vec /= self.world
            metrics = dict(zip(keys, vec.tolist()))
        else:
            metrics = {}
        logs = {**logs, **metrics, "step": self.state.global_step}
        self._metrics.clear()
        self.state.log_history.append(logs)
        if self.is_main:
            print(json.dumps({k: (round(v, 6) if isinstance(v, float) else v) for k, v in logs.items()}), flush=True)

    def save_model(self, output_dir: Optional[str] = None, _internal_call: bool = False):
        output_dir = output_dir or self.args.output_dir
        if self.is_main:
            save_pretrained(self.params, output_dir)
            if hasattr(self.processing_class, "save_pretrained"):
                self.processing_class.save_pretrained(output_dir)
        if self.world > 1:
            torch.distributed.barrier()

    def push_to_hub(self, **kwargs):
        warnings.warn("push_to_hub: no network access from the training box; skipped")
