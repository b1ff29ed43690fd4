"""HF-layout checkpoint IO (config.json + *.safetensors) so a PA-SFT output directory loads as the SC-GRPO input
(ref: scripts/train/SC_GRPO/*.sh:25 points MODEL_NAME_OR_PATH at the PA-SFT dir; `trainer.save_model`, grpo_ad.py:203)."""
from __future__ import annotations

import glob
import json
import os

import torch

from .config import VLMConfig
from .params import ParamStore


def load_config(path: str) -> VLMConfig:
    """config.json -> VLMConfig. The ORIGINAL dict (and generation_config.json when present) is carried in `cfg.extra` so that a
    later `save_pretrained` writes it back unchanged (bos / sliding-window / max_position_embeddings / the full eos list ...:
    HF and vLLM consumers of the saved directory get the source checkpoint's defaults), and the rollout stops on every id of the
    generation config's eos list, as the reference's vLLM engine does."""
    with open(os.path.join(path, "config.json")) as f:
        raw = json.load(f)
    cfg = VLMConfig.from_hf_dict(raw)
    cfg.extra["hf_config"] = raw
    gpath = os.path.join(path, "generation_config.json")
    if os.path.exists(gpath):
        with open(gpath) as f:
            gen = json.load(f)
        cfg.extra["generation_config"] = gen
        eos = gen.get("eos_token_id")
        if isinstance(eos, list):
            cfg.extra["eos_token_ids"] = [int(x) for x in eos]
        elif eos is not None:
            cfg.extra["eos_token_ids"] = [int(eos)]
    return cfg


def load_state_dict(path: str) -> dict:
    from safetensors.torch import load_file
    files = sorted(glob.glob(os.path.join(path, "*.safetensors")))
    if not files:
        bins = sorted(glob.glob(os.path.join(path, "pytorch_model*.bin")))
        if not bins:
            raise FileNotFoundError(f"no *.safetensors or pytorch_model*.bin under {path}")
        sd = {}
        for b in bins:
            sd.update(torch.load(b, map_location="cpu", weights_only=True))
        return sd
    sd = {}
    for f in files:
        sd.update(load_file(f, device="cpu"))
    return sd


def load_pretrained(path: str, device, with_grads=True, with_optimizer=True, moment_dtype=torch.float32):
    cfg = load_config(path)
    ps = ParamStore(cfg, device, with_grads=with_grads, with_optimizer=with_optimizer, moment_dtype=moment_dtype)
    ps.load_hf_state_dict(load_state_dict(path))
    return cfg, ps


def save_pretrained(ps: ParamStore, path: str, max_shard_bytes: int = 5 * 2 ** 30):
    """Write config.json + sharded safetensors with the 4.51-era key names the reference's `from_pretrained` reads."""
    from safetensors.torch import save_file
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(ps.cfg.extra.get("hf_config") or ps.cfg.to_hf_dict(), f, indent=2)   # the loaded config travels unchanged
    if ps.cfg.extra.get("generation_config"):
        with open(os.path.join(path, "generation_config.json"), "w") as f:
            json.dump(ps.cfg.extra["generation_config"], f, indent=2)
    sd = {k: v.cpu() for k, v in ps.hf_state_dict().items()}
    shards, cur, cur_bytes = [], {}, 0
    for k, v in sd.items():
        nb = v.numel() * v.element_size()
        if cur and cur_bytes + nb > max_shard_bytes:
            shards.append(cur)
            cur, cur_bytes = {}, 0
        cur[k] = v
        cur_bytes += nb
    shards.append(cur)
    if len(shards) == 1:
        save_file(shards[0], os.path.join(path, "model.safetensors"), metadata={"format": "pt"})
        return
    index = {"metadata": {"total_size": sum(v.numel() * v.element_size() for v in sd.values())}, "weight_map": {}}
    for i, sh in enumerate(shards):
        name = f"model-{i + 1:05d}-of-{len(shards):05d}.safetensors"
        save_file(sh, os.path.join(path, name), metadata={"format": "pt"})
        for k in sh:
            index["weight_map"][k] = name
    with open(os.path.join(path, "model.safetensors.index.json"), "w") as f:
        json.dump(index, f, indent=2)
