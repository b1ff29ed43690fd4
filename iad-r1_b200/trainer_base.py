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


def complement_ranges(done: list, n: int) -> list:
    """Sub-ranges of [0, n) not covered by the (disjoint) ranges in `done`, in ascending order."""
    out, cur = [], 0
    for lo, hi in sorted(done):
        if lo > cur:
            out.append((cur, lo))
        cur = max(cur, hi)
    if cur < n:
        out.append((cur, n))
    return out


GRAD_BUCKET_ELEMS = 64 << 20      # 256 MB of fp32 per NCCL call


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
        self._comm, self._comm_stream, self._reduced, self._micro_idx = None, None, [], 0
        if self.world > 1:
            self._setup_grad_comm()

    def _setup_grad_comm(self):
        """Own NCCL communicator for the gradient (C ABI: iadr1_comm_create / iadr1_grad_allreduce): rank 0 draws the unique
        id, torch.distributed carries its 128 bytes to the other ranks."""
        import ctypes as C
        lib = L.lib()
        uid = torch.zeros(128, dtype=torch.uint8, device=self.device)
        if self.rank == 0:
            buf = (C.c_ubyte * 128)()
            L.check(lib.iadr1_comm_unique_id(buf), "comm_unique_id")
            uid.copy_(torch.tensor(list(buf), dtype=torch.uint8))
        torch.distributed.broadcast(uid, src=0)
        raw = (C.c_ubyte * 128)(*uid.cpu().tolist())
        comm = C.c_void_p()
        L.check(lib.iadr1_comm_create(raw, self.rank, self.world, C.byref(comm)), "comm_create")
        self._comm, self._comm_stream = comm, torch.cuda.Stream(device=self.device)

    # ---- overlapped gradient all-reduce -------------------------------------------------------------------------------
    def _enqueue_grad_range(self, lo: int, hi: int):
        """Sum grad[lo, hi) over the ranks on the communication stream, ordered after everything the compute stream has
        enqueued so far (the range must be final: no later kernel of this optimizer step writes it)."""
        if self._comm is None or hi <= lo:
            return
        ev = torch.cuda.Event()
        ev.record()
        self._comm_stream.wait_event(ev)
        g = self.params.grad_flat
        L.check(L.lib().iadr1_grad_allreduce(self._comm, g.data_ptr() + 4 * lo, hi - lo, GRAD_BUCKET_ELEMS,
                                             self._comm_stream.cuda_stream), "grad_allreduce")
        self._reduced.append((lo, hi))

    def _layer_grad_ready(self, i: int):
        lo, hi = self.params.layer_matrix_range(i)
        self._enqueue_grad_range(lo, hi)

    def arm_overlap(self, last_micro_step: bool):
        """During the LAST micro-step of an accumulation window every decoder layer's gradient is final as soon as the
        backward sweep leaves the layer: its range goes to NCCL right away (layers retire last -> first)."""
        self.model.on_layer_grad_ready = self._layer_grad_ready if (last_micro_step and self._comm is not None) else None

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
        self.model.on_layer_grad_ready = None
        self._micro_idx = 0
        with self._phase("allreduce"):
            if self.world > 1:
                # the ONLY data-path collective: sum of the flat fp32 gradient over NVLink. Decoder-layer
                # ranges were enqueued during the last backward (arm_overlap); what is left - vision tower, embeddings,
                # norms and biases - goes now; the phase timer therefore shows the EXPOSED communication time
                for lo, hi in complement_ranges(self._reduced, ps.numel):
                    self._enqueue_grad_range(lo, hi)
                torch.cuda.current_stream().wait_stream(self._comm_stream)
                self._reduced = []
        with self._phase("optimizer"):
            scale = 1.0 / self.world
            self._sumsq.zero_()
            for lo, hi, _ in self._trainable_ranges():          # frozen ranges carry no gradient and are never zeroed
                if hi > lo:
                    L.check(lib.iadr1_sumsq_f32(ps.grad_flat.data_ptr() + 4 * lo, hi - lo, self._sumsq.data_ptr(), s), "sumsq")
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

    # ---- checkpoints with optimizer state (resume) ----------------------------------------------------------------------
    def save_checkpoint(self, ckpt_dir: str):
        """`checkpoint-<step>` directory: HF-layout weights (save_model) + what a resume needs - fp32 master weights, both
        Adam moments, the optimizer / scheduler counters and the trainer state (what the reference gets from the HF Trainer +
        DeepSpeed checkpoints behind `--save_steps`, ref: scripts/train/SC_GRPO/SC_GRPO_Qwen_Instruct_2_5_VL_3B.sh:60). Every
        rank holds identical optimizer state under plain data parallel, so rank 0 writes it."""
        self.save_model(ckpt_dir)
        if self.is_main:
            ps = self.params
            torch.save({"master": ps.master, "exp_avg": ps.exp_avg, "exp_avg_sq": ps.exp_avg_sq}, os.path.join(ckpt_dir, "optimizer.pt"))
            state = {"global_step": self.state.global_step, "epoch": self.state.epoch, "opt_step": self._opt_step,
                     "rollout_calls": getattr(self, "_rollout_calls", 0), "log_history": self.state.log_history,
                     "max_steps": self.state.max_steps, "numel": ps.numel, "moment_dtype": str(ps.exp_avg.dtype)}
            with open(os.path.join(ckpt_dir, "trainer_state.json"), "w") as f:
                json.dump(state, f)
            limit = getattr(self.args, "save_total_limit", None)
            if limit:      # keep the newest `save_total_limit` checkpoints (HF Trainer semantics)
                import re
                import shutil
                root = os.path.dirname(os.path.abspath(ckpt_dir))
                found = sorted((int(m.group(1)), d) for d in os.listdir(root) if (m := re.fullmatch(r"checkpoint-(\d+)", d)))
                for _, d in found[:-int(limit)]:
                    shutil.rmtree(os.path.join(root, d), ignore_errors=True)
        if self.world > 1:
            torch.distributed.barrier()

    def load_checkpoint(self, ckpt_dir: str):
        """Restore weights are NOT re-read here (construct the trainer on the checkpoint directory or on the same initial
        model): this restores master weights (and from them the bf16 working copy), moments, counters and state."""
        path = os.path.join(ckpt_dir, "optimizer.pt")
        if not os.path.isfile(path):
            raise FileNotFoundError(f"{ckpt_dir} holds no optimizer.pt: not a resumable checkpoint (save_strategy='steps' writes them)")
        with open(os.path.join(ckpt_dir, "trainer_state.json")) as f:
            st = json.load(f)
        ps = self.params
        if st["numel"] != ps.numel:
            raise ValueError(f"checkpoint has {st['numel']} parameters, the model {ps.numel}")
        blob = torch.load(path, map_location=self.device, weights_only=True)
        ps.master.copy_(blob["master"])
        ps.exp_avg.copy_(blob["exp_avg"].to(ps.exp_avg.dtype))
        ps.exp_avg_sq.copy_(blob["exp_avg_sq"].to(ps.exp_avg_sq.dtype))
        ps.flat.copy_(ps.master.to(torch.bfloat16))
        self.state.global_step, self.state.epoch = int(st["global_step"]), float(st["epoch"])
        self.state.log_history = list(st.get("log_history", []))
        self._opt_step = int(st["opt_step"])
        if hasattr(self, "_rollout_calls"):
            self._rollout_calls = int(st.get("rollout_calls", 0))
        return st

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
