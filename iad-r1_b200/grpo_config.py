"""`GRPOConfig` / `ModelConfig` / `ScriptArguments` / `TrlParser`: the configuration surface the reference's entry point
parses (ref: train/stage_rl/grpo_ad.py:211 `TrlParser((GRPOScriptArguments, GRPOConfig, ModelConfig))`).

Field names and defaults follow ref: trl/trl/trainer/grpo_config.py:176-413 and ref: train/stage_rl/configs.py:24-42,
plus the subset of `transformers.TrainingArguments` the launch scripts pass or the loop reads
(ref: scripts/train/SC_GRPO/*.sh:40-63). `transformers.TrainingArguments` itself cannot be constructed in this
environment (it demands `accelerate`), and the B200 path replaces the HF Trainer / DeepSpeed loop anyway, so this is a
plain dataclass; flags that only configured the replaced machinery (`--deepspeed`, `--gradient_checkpointing`,
`--attn_implementation`) are accepted and ignored with a note.
"""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass, field
from typing import Optional, Union


@dataclass
class GRPOConfig:
    output_dir: str = field(default="trainer_output")
    # ---- TrainingArguments subset -------------------------------------------------------------------------------
    per_device_train_batch_size: int = 1
    per_device_eval_batch_size: int = 1
    gradient_accumulation_steps: int = 1
    learning_rate: float = 1e-6                # grpo_config.py:300
    weight_decay: float = 0.0
    adam_beta1: float = 0.9
    adam_beta2: float = 0.999
    adam_epsilon: float = 1e-8
    max_grad_norm: float = 1.0
    num_train_epochs: float = 3.0
    max_steps: int = -1
    lr_scheduler_type: str = "linear"
    warmup_steps: int = 0
    warmup_ratio: float = 0.0
    logging_steps: float = 500
    save_steps: float = 500
    save_strategy: str = "steps"
    save_total_limit: Optional[int] = None
    seed: int = 42
    data_seed: Optional[int] = None
    bf16: bool = False
    fp16: bool = False
    tf32: Optional[bool] = None
    gradient_checkpointing: bool = False      # honoured through `activation_recompute` (below): "auto" recomputes only when needed
    deepspeed: Optional[str] = None           # ignored: plain data parallel replaces ZeRO-3
    ddp_timeout: int = 1800
    report_to: Union[None, str, list[str]] = "none"
    run_name: Optional[str] = None
    eval_strategy: str = "no"
    do_train: bool = False
    do_eval: bool = False
    overwrite_output_dir: bool = False
    resume_from_checkpoint: Optional[str] = None
    dataloader_num_workers: int = 0
    push_to_hub: bool = False
    hub_model_id: Optional[str] = None
    log_level: str = "passive"
    disable_tqdm: Optional[bool] = None
    optim: str = "adamw_torch"
    # ---- GRPO (trl/trl/trainer/grpo_config.py) ---------------------------------------------------------------------------
    model_init_kwargs: Optional[Union[dict, str]] = None
    disable_dropout: bool = False
    remove_unused_columns: Optional[bool] = False
    max_prompt_length: Optional[int] = 512
    num_generations: Optional[int] = 8
    max_completion_length: Optional[int] = 256
    ds3_gather_for_generation: bool = True
    shuffle_dataset: Optional[bool] = True
    temperature: float = 0.9
    top_p: float = 1.0
    top_k: Optional[int] = 50
    min_p: Optional[float] = None
    repetition_penalty: float = 1.0
    cache_implementation: Optional[str] = None
    use_vllm: bool = False
    vllm_server_host: str = "0.0.0.0"
    vllm_server_port: int = 8000
    vllm_server_timeout: float = 240.0
    vllm_guided_decoding_regex: Optional[str] = None
    beta: float = 0.04
    num_iterations: int = 1
    epsilon: float = 0.2
    epsilon_high: Optional[float] = None
    reward_weights: Optional[list[float]] = None
    scale_rewards: bool = True
    loss_type: str = "bnpo"
    mask_truncated_completions: bool = False
    sync_ref_model: bool = False
    ref_model_mixup_alpha: float = 0.6
    ref_model_sync_steps: int = 512
    use_liger_loss: bool = False
    log_completions: bool = False
    num_completions_to_print: Optional[int] = None
    wandb_log_unique_prompts: Optional[bool] = False
    # ---- IAD-R1 extension (train/stage_rl/configs.py:29-42) ----------------------------------------------------------------
    benchmarks: list[str] = field(default_factory=list)
    callbacks: list[str] = field(default_factory=list)
    system_prompt: Optional[str] = None
    hub_model_revision: Optional[str] = "main"
    overwrite_hub_revision: bool = False
    push_to_hub_revision: bool = False
    # ---- B200 path knobs (new; defaults reproduce the reference's semantics) -----------------------------------------------
    loss_mode: str = "sc"                      # "sc": sc_grpo_trainer.py:796-798; "clip": trl grpo_trainer.py:1182-1219
    rollout_top_p: float = 0.9                 # SCGRPOTrainer hard-codes top_p=0.9, top_k=50 (sc_grpo_trainer.py:353-358)
    rollout_top_k: int = 50
    batched_rollout: bool = True               # roll out every group of an accumulation window in one decode batch
    shared_prefix: bool = True                 # score a group as [prompt | G completions]: the prompt is computed once
    window_vision: bool = True                 # vision tower fwd/bwd once per accumulation window (batched_rollout only)
    rollout_forbid_eos: bool = False           # benchmarking only: fixed-length completions
    gpu_image_preprocess: bool = True          # Qwen families: resize / normalise / patchify on the GPU (csrc/preprocess.cu)
    activation_recompute: str = "auto"         # per-layer recompute in the backward ("on" | "off" | "auto": only for models
                                               # whose unsharded state leaves too little HBM, e.g. Qwen2.5-VL-7B; 3B keeps
                                               # every activation resident even when --gradient_checkpointing is passed)
    optimizer_moments: str = "auto"            # "fp32" | "bf16" (stochastic rounding) | "auto" (bf16 only when fp32 cannot fit)
    rollout_seed: Optional[int] = None

    def __post_init__(self):
        if isinstance(self.report_to, str):
            self.report_to = [] if self.report_to in ("none", "") else [self.report_to]
        if self.report_to is None:
            self.report_to = []
        if self.num_generations is not None and self.num_generations < 1:
            raise ValueError("num_generations must be >= 1")
        if self.loss_mode not in ("sc", "clip"):
            raise ValueError("loss_mode must be 'sc' or 'clip'")
        if self.loss_type not in ("grpo", "bnpo", "dr_grpo"):
            raise ValueError(f"Unknown loss type: {self.loss_type}")
        if self.fp16:
            raise ValueError("fp16 is not supported by the B200 path (bf16 parameters, fp32 master weights)")

    @property
    def world_size(self) -> int:
        return int(os.environ.get("WORLD_SIZE", "1"))

    @property
    def process_index(self) -> int:
        return int(os.environ.get("RANK", "0"))

    @property
    def local_process_index(self) -> int:
        return int(os.environ.get("LOCAL_RANK", "0"))


@dataclass
class ModelConfig:
    """Subset of trl.ModelConfig read by grpo_ad.py (:189, :195-196)."""
    model_name_or_path: Optional[str] = None
    model_revision: str = "main"
    torch_dtype: Optional[str] = None
    trust_remote_code: bool = False
    attn_implementation: Optional[str] = None  # accepted; attention always runs on the library's tcgen05 kernels
    use_peft: bool = False
    lora_r: int = 16
    lora_alpha: int = 32
    lora_dropout: float = 0.05
    lora_target_modules: Optional[list[str]] = None
    load_in_8bit: bool = False
    load_in_4bit: bool = False


@dataclass
class ScriptArguments:
    """ref: trl/trl/scripts/utils.py:34-76."""
    dataset_name: Optional[str] = None
    dataset_config: Optional[str] = None
    dataset_train_split: str = "train"
    dataset_test_split: str = "test"
    gradient_checkpointing_use_reentrant: bool = False
    ignore_bias_buffers: bool = False


def get_peft_config(model_args: ModelConfig):
    if model_args.use_peft:
        raise NotImplementedError("LoRA/PEFT training is outside the B200 hot path (SURVEY.md §8f item 4)")
    return None


class TrlParser:
    """HfArgumentParser + `--config file.yaml` (+ `env:` block exported to os.environ), ref: trl/trl/scripts/utils.py:98-224."""

    def __init__(self, dataclass_types):
        from transformers import HfArgumentParser
        if not isinstance(dataclass_types, (list, tuple)):
            dataclass_types = [dataclass_types]
        self.dataclass_types = list(dataclass_types)
        self._parser = HfArgumentParser(self.dataclass_types)

    def parse_args_and_config(self, args=None, return_remaining_strings: bool = False):
        args = list(args) if args is not None else sys.argv[1:]
        if "--config" in args:
            import yaml
            i = args.index("--config")
            path = args[i + 1]
            del args[i:i + 2]
            with open(path) as f:
                conf = yaml.safe_load(f) or {}
            env = conf.pop("env", {})
            if not isinstance(env, dict):
                raise ValueError("`env` field should be a dict in the YAML file.")
            for k, v in env.items():
                os.environ[k] = str(v)
            defaults = []
            for k, v in conf.items():
                if f"--{k}" in args:
                    continue  # command line wins
                if isinstance(v, list):
                    defaults += [f"--{k}"] + [str(x) for x in v]
                else:
                    defaults += [f"--{k}", str(v)]
            args = defaults + args
        out = self._parser.parse_args_into_dataclasses(args=args, return_remaining_strings=return_remaining_strings)
        if return_remaining_strings:
            return out
        return tuple(out)
