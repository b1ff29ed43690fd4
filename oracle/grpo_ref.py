"""ORACLE (test infrastructure, never on the product path): CPU restatement of the reference's GRPO arithmetic.

Each function follows the reference lines it cites and nothing else:
  completion_mask / attention mask   ref: train/stage_rl/trainer/sc_grpo_trainer.py:722-728
  KL (k3 estimator)                  ref: ...sc_grpo_trainer.py:746
  advantages                         ref: ...sc_grpo_trainer.py:784-793
  SC-GRPO loss                       ref: ...sc_grpo_trainer.py:796-798
  clip-mode loss                     ref: trl/trl/trainer/grpo_trainer.py:1182-1219
  pad                                ref: trl/trl/trainer/utils.py:418-479   (KATs: trl/tests/test_utils.py:47-130)
  selective_log_softmax              ref: trl/trl/trainer/utils.py:1683-1715 (KATs: trl/tests/test_utils.py:494-512)
  RepeatSampler                      ref: trl/trl/trainer/grpo_trainer.py:78-172 (tests: trl/tests/test_grpo_trainer.py:36-142)
Pinned by tests/test_oracle.py against those KATs. Only tests/, smoke() and bench.py's CPU legs import this.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def completion_mask_ref(completion_ids: torch.Tensor, eos_token_id: int) -> torch.Tensor:
    is_eos = completion_ids == eos_token_id
    eos_idx = torch.full((is_eos.size(0),), is_eos.size(1), dtype=torch.long, device=completion_ids.device)
    eos_idx[is_eos.any(dim=1)] = is_eos.int().argmax(dim=1)[is_eos.any(dim=1)]
    sequence_indices = torch.arange(is_eos.size(1), device=completion_ids.device).expand(is_eos.size(0), -1)
    return (sequence_indices <= eos_idx.unsqueeze(1)).int()


def kl_ref(ref_logps, logps):
    return torch.exp(ref_logps - logps) - (ref_logps - logps) - 1


def advantages_ref(rewards_per_func: torch.Tensor, num_generations: int):
    rewards = rewards_per_func.sum(dim=1)
    mean_g = rewards.view(-1, num_generations).mean(dim=1)
    std_g = rewards.view(-1, num_generations).std(dim=1)
    mean_g = mean_g.repeat_interleave(num_generations, dim=0)
    std_g = std_g.repeat_interleave(num_generations, dim=0)
    return (rewards - mean_g) / (std_g + 1e-4), rewards, std_g


def sc_grpo_loss_ref(logps, ref_logps, advantages, completion_mask, beta: float):
    per_token_kl = kl_ref(ref_logps, logps)
    per_token_loss = torch.exp(logps - logps.detach()) * advantages.unsqueeze(1)
    per_token_loss = -(per_token_loss - beta * per_token_kl)
    loss = ((per_token_loss * completion_mask).sum(dim=1) / completion_mask.sum(dim=1)).mean()
    mean_kl = ((per_token_kl * completion_mask).sum(dim=1) / completion_mask.sum(dim=1)).mean()
    return loss, mean_kl


def clip_loss_ref(logps, old_logps, ref_logps, advantages, completion_mask, beta, eps_low, eps_high, loss_type,
                  max_completion_length):
    old = logps.detach() if old_logps is None else old_logps
    coef_1 = torch.exp(logps - old)
    coef_2 = torch.clamp(coef_1, 1 - eps_low, 1 + eps_high)
    per_token_loss = -torch.min(coef_1 * advantages.unsqueeze(1), coef_2 * advantages.unsqueeze(1))
    if beta != 0.0:
        per_token_loss = per_token_loss + beta * kl_ref(ref_logps, logps)
    if loss_type == "grpo":
        return ((per_token_loss * completion_mask).sum(-1) / completion_mask.sum(-1).clamp(min=1.0)).mean()
    if loss_type == "bnpo":
        return (per_token_loss * completion_mask).sum() / completion_mask.sum().clamp(min=1.0)
    if loss_type == "dr_grpo":
        return (per_token_loss * completion_mask).sum() / (per_token_loss.size(0) * max_completion_length)
    raise ValueError(f"Unknown loss type: {loss_type}")


def pad_ref(tensors, padding_value=0, padding_side="right", pad_to_multiple_of=None):
    output_shape = np.max([t.shape for t in tensors], 0).tolist()
    if pad_to_multiple_of is not None:
        remainder = output_shape[0] % pad_to_multiple_of
        if remainder != 0:
            output_shape[0] += pad_to_multiple_of - remainder
    output = torch.full((len(tensors), *output_shape), padding_value, dtype=tensors[0].dtype)
    for i, t in enumerate(tensors):
        if padding_side == "left":
            seq_start = output_shape[0] - t.shape[0]
        elif padding_side == "right":
            seq_start = 0
        else:
            raise ValueError("padding_side must be 'left' or 'right'")
        slices = (slice(seq_start, seq_start + t.shape[0]),) + tuple(slice(0, s) for s in t.shape[1:])
        output[i][slices] = t
    return output


def selective_log_softmax_ref(logits, index):
    if logits.dtype in [torch.float32, torch.float64]:
        selected = torch.gather(logits, dim=-1, index=index.unsqueeze(-1)).squeeze(-1)
        lse = torch.stack([torch.logsumexp(lg, dim=-1) for lg in logits])
        return selected - lse
    out = []
    for row_logits, row_labels in zip(logits, index):
        row_logps = F.log_softmax(row_logits, dim=-1)
        out.append(row_logps.gather(dim=-1, index=row_labels.unsqueeze(-1)).squeeze(-1))
    return torch.stack(out)


def repeat_sampler_ref(num_samples, mini_repeat_count, batch_size=1, repeat_count=1, shuffle=True, seed=None):
    if shuffle:
        gen = torch.Generator()
        if seed is not None:
            gen.manual_seed(seed)
        indexes = torch.randperm(num_samples, generator=gen).tolist()
    else:
        indexes = list(range(num_samples))
    chunks = [indexes[i:i + batch_size] for i in range(0, len(indexes), batch_size)]
    chunks = [c for c in chunks if len(c) == batch_size]
    out = []
    for chunk in chunks:
        for _ in range(repeat_count):
            for index in chunk:
                out.extend([index] * mini_repeat_count)
    return out
