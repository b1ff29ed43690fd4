"""The part of SC-GRPO that stays in Python (BASELINE.json north_star): EOS masking, the k3 KL-to-reference, group-relative
advantages and the SC / clip losses, as torch ops on [G, C] tensors. Its gradient with respect to the per-token
log-probs is what seeds the CUDA backward (model.logprobs_backward).

Behaviour follows ref: train/stage_rl/trainer/sc_grpo_trainer.py:722-728 (mask), :746 (KL), :784-793 (advantages,
unbiased std + 1e-4), :796-798 (loss) and, for `loss_mode="clip"`, ref: trl/trl/trainer/grpo_trainer.py:1182-1219.
Quirks kept on purpose: Q1 no ratio clip / no temperature in SC mode, Q8 the first EOS is inside
the mask and the denominator is never clamped, Q13 G=1 gives NaN advantages.
"""
from __future__ import annotations

import torch


def completion_mask(completion_ids: torch.Tensor, eos_token_id: int) -> torch.Tensor:
    """1 up to and including the first EOS of each row (all ones when a row has none)."""
    is_eos = completion_ids == eos_token_id
    C = completion_ids.size(1)
    first = torch.where(is_eos.any(1), is_eos.int().argmax(1), torch.full_like(is_eos[:, 0], C, dtype=torch.long))
    return (torch.arange(C, device=completion_ids.device)[None, :] <= first[:, None]).int()


def mask_truncated(mask: torch.Tensor, completion_ids: torch.Tensor, eos_token_id: int) -> torch.Tensor:
    """`mask_truncated_completions` (ref: trl/trl/trainer/grpo_trainer.py:976-989): rows that never produced EOS are
    dropped from the loss entirely."""
    truncated = ~(completion_ids == eos_token_id).any(1)
    return mask * (~truncated).int()[:, None]


def group_advantages(rewards_per_func: torch.Tensor, num_generations: int, scale_rewards: bool = True,
                     reward_weights: torch.Tensor | None = None):
    """rewards_per_func [B*G, n_funcs] -> (advantages [B*G], rewards [B*G], per-row group std [B*G])."""
    if reward_weights is not None:
        rewards = (rewards_per_func * reward_weights[None, :]).nansum(1)
    else:
        rewards = rewards_per_func.sum(1)
    grouped = rewards.view(-1, num_generations)
    mean = grouped.mean(1).repeat_interleave(num_generations)
    std = grouped.std(1).repeat_interleave(num_generations)
    adv = rewards - mean
    if scale_rewards:
        adv = adv / (std + 1e-4)
    return adv, rewards, std


def per_token_kl(ref_logps: torch.Tensor, logps: torch.Tensor) -> torch.Tensor:
    d = ref_logps - logps
    return torch.exp(d) - d - 1


def sc_grpo_loss(logps, ref_logps, advantages, mask, beta: float):
    """-(exp(lp - sg(lp)) * A - beta * KL), masked mean per row, mean over rows. Returns (loss, mean_kl)."""
    kl = per_token_kl(ref_logps, logps) if ref_logps is not None else torch.zeros_like(logps)
    ptl = -(torch.exp(logps - logps.detach()) * advantages[:, None] - beta * kl)
    denom = mask.sum(1)
    loss = ((ptl * mask).sum(1) / denom).mean()
    mean_kl = ((kl * mask).sum(1) / denom).mean()
    return loss, mean_kl


def clip_grpo_loss(logps, old_logps, ref_logps, advantages, mask, beta, eps_low, eps_high, loss_type, max_completion_length):
    old = logps.detach() if old_logps is None else old_logps
    c1 = torch.exp(logps - old)
    c2 = torch.clamp(c1, 1 - eps_low, 1 + eps_high)
    ptl = -torch.min(c1 * advantages[:, None], c2 * advantages[:, None])
    kl = None
    if beta != 0.0:
        kl = per_token_kl(ref_logps, logps)
        ptl = ptl + beta * kl
    if loss_type == "grpo":
        loss = ((ptl * mask).sum(-1) / mask.sum(-1).clamp(min=1.0)).mean()
    elif loss_type == "bnpo":
        loss = (ptl * mask).sum() / mask.sum().clamp(min=1.0)
    elif loss_type == "dr_grpo":
        loss = (ptl * mask).sum() / (ptl.size(0) * max_completion_length)
    else:
        raise ValueError(f"Unknown loss type: {loss_type}")
    mean_kl = (kl * mask).sum() / mask.sum() if kl is not None else torch.zeros((), device=logps.device)
    return loss, mean_kl
