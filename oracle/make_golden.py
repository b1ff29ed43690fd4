"""Generates tests/golden/tiny_<family>.pt by running the HF implementation (oracle/hf_oracle.py) and the reference loss
restatement (oracle/grpo_ref.py) HERE on CPU. Run:  python oracle/make_golden.py
The fixtures are what the GPU box compares against (it has neither /root/reference nor a need to run HF)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iad_r1_b200.config import tiny_config  # noqa: E402
from iad_r1_b200.geometry import mrope_position_ids, position_ids as family_position_ids, image_token_count, patchify_crops, clip_pixel_rows  # noqa: E402
from oracle import grpo_ref  # noqa: E402
from oracle.hf_oracle import build_hf_model, hf_logits, hf_logits_llava, hf_logits_llava15, per_token_logps  # noqa: E402


def synthetic_batch(cfg, G=4, C=12, grid=(1, 8, 8), seed=0):
    """One prompt (text + one image) x G completions, completions right-padded after EOS (sc_grpo_trainer.py:679-683)."""
    rng = np.random.RandomState(seed)
    n_img = grid[0] * grid[1] * grid[2] // cfg.vision.spatial_merge_size ** 2
    prompt = list(rng.randint(10, 900, size=6)) + [cfg.vision_start_token_id] + [cfg.image_token_id] * n_img + \
        [cfg.vision_end_token_id] + list(rng.randint(10, 900, size=5))
    P = len(prompt)
    comp = rng.randint(10, 900, size=(G, C))
    eos_at = [None, 7, C - 1, 2][:G] + [None] * max(0, G - 4)
    for g, e in enumerate(eos_at):
        if e is not None:
            comp[g, e] = cfg.eos_token_id
            comp[g, e + 1:] = cfg.pad_token_id
    ids = np.concatenate([np.tile(np.array(prompt)[None], (G, 1)), comp], 1).astype(np.int64)
    np_patches = grid[0] * grid[1] * grid[2]
    gen = torch.Generator().manual_seed(seed + 7)
    px = torch.randn(np_patches, cfg.vision.patch_dim, generator=gen).to(torch.bfloat16).float()
    return ids, P, px


def synthetic_batch_llava(cfg, G=4, C=12, image_hw=(80, 100), seed=0):
    """LLaVA-OneVision twin: one (H, W) image -> anyres crops [n_crops, 3, S, S] (bf16-valued), prompt with the packed
    number of <image> placeholders (unpadding drops feature rows for this aspect ratio), G completions."""
    rng = np.random.RandomState(seed)
    n_crops = 5
    n_img = image_token_count(cfg, (n_crops, image_hw[0], image_hw[1]))
    prompt = list(rng.randint(10, 900, size=6)) + [cfg.image_token_id] * n_img + list(rng.randint(10, 900, size=5))
    P = len(prompt)
    comp = rng.randint(10, 900, size=(G, C))
    for g, e in enumerate([None, 7, C - 1, 2][:G]):
        if e is not None:
            comp[g, e] = cfg.eos_token_id
            comp[g, e + 1:] = cfg.pad_token_id
    ids = np.concatenate([np.tile(np.array(prompt)[None], (G, 1)), comp], 1).astype(np.int64)
    gen = torch.Generator().manual_seed(seed + 7)
    S = cfg.vision.image_size
    crops = torch.randn(n_crops, cfg.vision.in_channels, S, S, generator=gen).to(torch.bfloat16).float()
    return ids, P, crops, (n_crops, image_hw[0], image_hw[1])


def synthetic_batch_llava15(cfg, G=4, C=12, seed=0):
    """LLaVA-1.5 twin: one S x S crop [1, 3, S, S] (bf16-valued), prompt with tokens_per_crop - 1 <image> placeholders."""
    rng = np.random.RandomState(seed)
    n_img = cfg.vision.tokens_per_crop - 1
    prompt = list(rng.randint(10, 900, size=6)) + [cfg.image_token_id] * n_img + list(rng.randint(10, 900, size=5))
    P = len(prompt)
    comp = rng.randint(10, 900, size=(G, C))
    for g, e in enumerate([None, 7, C - 1, 2][:G]):
        if e is not None:
            comp[g, e] = cfg.eos_token_id
            comp[g, e + 1:] = cfg.pad_token_id
    ids = np.concatenate([np.tile(np.array(prompt)[None], (G, 1)), comp], 1).astype(np.int64)
    gen = torch.Generator().manual_seed(seed + 7)
    S = cfg.vision.image_size
    img = torch.randn(1, cfg.vision.in_channels, S, S, generator=gen).to(torch.bfloat16).float()
    return ids, P, img, (1, S, S)


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    only = sys.argv[1:]
    for family in ("qwen2_5_vl", "qwen2_vl", "llava_onevision", "llava", "llava_next"):
        if only and family not in only:
            continue
        cfg = tiny_config(family)
        G, C, grid = 4, 12, (1, 8, 8)
        if family in ("llava_onevision", "llava_next"):
            ids, P, crops, grid = synthetic_batch_llava(cfg, G, C)
            px = patchify_crops(crops, cfg.vision.patch_size) if family == "llava_onevision" else clip_pixel_rows(crops, cfg.vision)
        elif family == "llava":
            ids, P, crops, grid = synthetic_batch_llava15(cfg, G, C)
            px = clip_pixel_rows(crops, cfg.vision)
        else:
            ids, P, px = synthetic_batch(cfg, G, C, grid)
        pos, deltas = family_position_ids(ids, [grid] * G, cfg)
        ids_t, pos_t = torch.from_numpy(ids), torch.from_numpy(pos)
        grid_t = torch.tensor([grid] * G)
        comp = ids_t[:, P:]
        mask = grpo_ref.completion_mask_ref(comp, cfg.eos_token_id)
        attn_mask = torch.cat([torch.ones(G, P, dtype=torch.int64), mask.long()], 1)

        model = build_hf_model(cfg, seed=0, dtype=torch.float32)
        sd = {k: v.detach().to(torch.bfloat16).clone() for k, v in model.state_dict().items()}
        if family in ("llava_onevision", "llava_next"):
            sizes = torch.tensor([[grid[1], grid[2]]] * G)

            def fwd(m):
                return hf_logits_llava(m, ids_t, crops[None].repeat(G, 1, 1, 1, 1), sizes, pos_t[0], attn_mask)
        elif family == "llava":
            def fwd(m):
                return hf_logits_llava15(m, ids_t, crops.repeat(G, 1, 1, 1), pos_t[0], attn_mask)
        else:
            def fwd(m):
                return hf_logits(m, ids_t, px.repeat(G, 1), grid_t, pos_t, attn_mask)
        logits = fwd(model)
        logp = per_token_logps(logits, ids_t)[:, P - 1:]            # [G, C] fp32 oracle
        # reference-form bf16 pass (what the reference actually runs: bf16 model, bf16 log_softmax)
        m16 = build_hf_model(cfg, seed=0, dtype=torch.bfloat16)
        with torch.no_grad():
            logp16 = per_token_logps(fwd(m16), ids_t)[:, P - 1:]
        del m16
        gen = torch.Generator().manual_seed(11)
        ref_logp = (logp.detach() + 0.3 * torch.randn(G, C, generator=gen)).contiguous()
        rewards = torch.tensor([[1.0, 1.0], [0.0, 1.0], [2.0, 0.0], [0.5, 1.0]])[:G]
        adv, rew, std = grpo_ref.advantages_ref(rewards, G)
        beta = 0.04
        loss, mean_kl = grpo_ref.sc_grpo_loss_ref(logp, ref_logp, adv, mask, beta)
        logp.retain_grad()
        loss.backward()
        grads = {k: p.grad.detach().to(torch.bfloat16).clone() for k, p in model.named_parameters() if p.grad is not None}
        # second loss form (clip mode, bnpo) on the same forward, for the Python-side loss parity
        old = (logp.detach() + 0.05 * torch.randn(G, C, generator=gen))
        clip = grpo_ref.clip_loss_ref(logp.detach(), old, ref_logp, adv, mask, beta, 0.2, 0.2, "bnpo", C)
        fix = dict(family=family, G=G, C=C, P=P, grid=list(grid), input_ids=ids_t, position_ids=pos_t, rope_deltas=torch.from_numpy(deltas),
                   pixel_values=px.to(torch.bfloat16), attention_mask=attn_mask, completion_mask=mask,
                   state_dict=sd, logp_fp32=logp.detach().clone(), logp_bf16_ref=logp16.float(),
                   dlogp=logp.grad.detach().clone(), ref_logp=ref_logp, rewards_per_func=rewards, advantages=adv,
                   beta=beta, loss=loss.detach(), mean_kl=mean_kl.detach(), old_logp=old, clip_bnpo_loss=clip.detach(),
                   grads=grads, hidden_last=None)
        path = os.path.join(out_dir, f"tiny_{family}.pt")
        torch.save(fix, path)
        err16 = (logp16.float() - logp.detach()).abs().max().item()
        print(f"{family}: P={P} loss={loss.item():.6f} kl={mean_kl.item():.6f} |logp16-logp32|max={err16:.4f} "
              f"params={sum(v.numel() for v in sd.values())} file={os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
