"""Synthetic stand-ins for the assets that cannot exist offline (tokenizer files, datasets, checkpoints): a processor with
the `AutoProcessor` call surface the trainer uses (ref: train/stage_rl/trainer/sc_grpo_trainer.py:600-621, :749), built on
the REAL HF Qwen2-VL image processor (smart-resize, normalise, merge-block-major patchify) plus a hashing word tokenizer;
and the seeded image + question dataset of SURVEY.md §8d. Used by bench.py, smoke() and the tests only."""
from __future__ import annotations

import re
import zlib

import numpy as np
import torch

from .config import VLMConfig

_SPECIAL = ("<|im_start|>", "<|im_end|>", "<|vision_start|>", "<|vision_end|>", "<|image_pad|>", "<image>")
_WORDS = ["<think>", "</think>", "<answer>", "</answer>", "<location>", "</location>", "<type>", "</type>", "yes", "no",
          "scratch", "top", "left", "center", "the", "defect", "surface", "image", "is", "a"]


class _Encoding(dict):
    def to(self, device):
        return _Encoding({k: (v.to(device) if torch.is_tensor(v) else v) for k, v in self.items()})

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class SyntheticProcessor:
    def __init__(self, cfg: VLMConfig, min_pixels: int = 3136, max_pixels: int = 12845056):
        self.cfg = cfg
        self.llava = cfg.family in ("llava_onevision", "llava", "llava_next")     # bare <image> placeholder in the chat template
        self.llava15 = cfg.family == "llava"
        if cfg.family == "llava_next":
            # the REAL HF anyres image processor of LLaVA-1.6 (best-resolution select, resize + pad, S-pixel crops + base crop)
            from transformers import LlavaNextImageProcessor
            S = cfg.vision.image_size
            self.image_processor = LlavaNextImageProcessor(image_grid_pinpoints=cfg.extra["image_grid_pinpoints"],
                                                           size={"shortest_edge": S}, crop_size={"height": S, "width": S})
        elif self.llava15:
            # the REAL HF CLIP image processor (shortest edge -> S bicubic, centre crop S x S, CLIP mean / std)
            from transformers import CLIPImageProcessor
            S = cfg.vision.image_size
            self.image_processor = CLIPImageProcessor(size={"shortest_edge": S}, crop_size={"height": S, "width": S})
        elif self.llava:
            # the REAL HF anyres image processor (best-resolution select, resize + pad, 384-pixel crops + base crop)
            from transformers import LlavaOnevisionImageProcessor
            S = cfg.vision.image_size
            self.image_processor = LlavaOnevisionImageProcessor(image_grid_pinpoints=cfg.extra["image_grid_pinpoints"],
                                                                size={"height": S, "width": S})
        else:
            from transformers import Qwen2VLImageProcessor
            self.image_processor = Qwen2VLImageProcessor(min_pixels=min_pixels, max_pixels=max_pixels,
                                                         patch_size=cfg.vision.patch_size,
                                                         temporal_patch_size=cfg.vision.temporal_patch_size,
                                                         merge_size=cfg.vision.spatial_merge_size)
        self.pad_token_id, self.eos_token_id = cfg.pad_token_id, cfg.eos_token_id
        self.pad_token, self.eos_token = "<|endoftext|>", "<|im_end|>"
        self.tokenizer = self
        v = cfg.text.vocab_size
        self._lo, self._hi = min(1000, v // 4), min(100000, v - 64) if v > 2000 else v // 2
        self._special = {"<|vision_start|>": cfg.vision_start_token_id, "<|vision_end|>": cfg.vision_end_token_id,
                         "<|image_pad|>": cfg.image_token_id, "<image>": cfg.image_token_id,   # llava templates spell it <image>
                         "<|im_end|>": cfg.eos_token_id, "<|im_start|>": self._lo - 1}
        self._word_ids = {w: self._lo + i for i, w in enumerate(_WORDS)}
        self._id_words = {i: w for w, i in self._word_ids.items()}

    # -- chat template (Qwen2-VL style) ----------------------------------------------------------------------------
    def apply_chat_template(self, messages, tokenize=False, add_generation_prompt=True, **_):
        out = []
        for m in messages:
            out.append(f"<|im_start|>{m['role']}\n")
            c = m["content"]
            if isinstance(c, str):
                out.append(c)
            else:
                for part in c:
                    if part.get("type") == "image":
                        # llava-onevision chat template: a bare <image> placeholder followed by a newline
                        out.append("<|image_pad|>\n" if self.llava else "<|vision_start|><|image_pad|><|vision_end|>")
                    elif part.get("type") == "text":
                        out.append(part["text"])
            out.append("<|im_end|>\n")
        if add_generation_prompt:
            out.append("<|im_start|>assistant\n")
        return "".join(out)

    def _tok(self, text: str):
        ids = []
        for piece in re.split("(" + "|".join(re.escape(s) for s in _SPECIAL) + ")", text):
            if not piece:
                continue
            if piece in self._special:
                ids.append(self._special[piece])
                continue
            for w in re.findall(r"</?\w+>|\w+|[^\w\s]", piece):
                ids.append(self._word_ids.get(w, self._lo + len(_WORDS) + zlib.crc32(w.encode()) % (self._hi - self._lo - len(_WORDS))))
        return ids

    def __call__(self, text=None, images=None, return_tensors="pt", padding=True, padding_side="left",
                 add_special_tokens=False, **_):
        texts = [text] if isinstance(text, str) else list(text)
        enc = {}
        grids = []
        counts = []
        if images is not None and len(images) > 0:
            im = self.image_processor(images=list(images), return_tensors="pt")
            if self.llava15:
                enc["pixel_values"] = im["pixel_values"]                      # [n_images, 3, S, S], one crop per image
                counts = [self.cfg.vision.tokens_per_crop - 1] * im["pixel_values"].shape[0]
            elif self.llava:
                from .geometry import image_token_count, llava_image_layout
                enc["pixel_values"], enc["image_sizes"] = im["pixel_values"], im["image_sizes"]
                for h, w in im["image_sizes"].tolist():
                    n_crops = llava_image_layout(self.cfg, (h, w))[0]
                    counts.append(image_token_count(self.cfg, (n_crops, h, w)))
            else:
                enc["pixel_values"], enc["image_grid_thw"] = im["pixel_values"], im["image_grid_thw"]
                merge = self.cfg.vision.spatial_merge_size ** 2
                counts = [g[0] * g[1] * g[2] // merge for g in im["image_grid_thw"].tolist()]
        rows, cursor = [], 0
        for t in texts:
            while "<|image_pad|>" in t and cursor < len(counts):
                n_tok = counts[cursor]
                cursor += 1
                t = t.replace("<|image_pad|>", "<|placeholder|>" * n_tok, 1)
            rows.append(self._tok(t.replace("<|placeholder|>", "<|image_pad|>")))
        P = max(len(r) for r in rows)
        ids = np.full((len(rows), P), self.pad_token_id, dtype=np.int64)
        am = np.zeros((len(rows), P), dtype=np.int64)
        for i, r in enumerate(rows):
            if padding_side == "left":
                ids[i, P - len(r):], am[i, P - len(r):] = r, 1
            else:
                ids[i, :len(r)], am[i, :len(r)] = r, 1
        enc["input_ids"], enc["attention_mask"] = torch.from_numpy(ids), torch.from_numpy(am)
        return _Encoding(enc)

    def batch_decode(self, ids, skip_special_tokens=True, **_):
        if torch.is_tensor(ids):
            ids = ids.tolist()
        special = set(self._special.values()) | {self.pad_token_id, self.eos_token_id}
        out = []
        for row in ids:
            words = []
            for t in row:
                if skip_special_tokens and t in special:
                    continue
                words.append(self._id_words.get(t, f"w{t}"))
            out.append(" ".join(words))
        return out

    def save_pretrained(self, path):
        import json, os
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "synthetic_processor.json"), "w") as f:
            json.dump({"kind": "SyntheticProcessor", "vocab_size": self.cfg.text.vocab_size}, f)


def synthetic_image(i: int, size: int = 448):
    """torch.Generator().manual_seed(1234+i), uint8 uniform, size x size x 3, as a PIL RGB image."""
    from PIL import Image
    g = torch.Generator().manual_seed(1234 + i)
    arr = torch.randint(0, 256, (size, size, 3), generator=g, dtype=torch.uint8).numpy()
    return Image.fromarray(arr, "RGB")


QUESTION_PROMPT = ("You are an expert in detecting defects in image. Your task is to detect if there are any defects in the "
                   "test image.{Question}")  # ref: train/stage_rl/grpo_ad.py:88-91


def synthetic_dataset(n: int, image_size: int = 448):
    """Rows shaped like the reference's mapped dataset (grpo_ad.py:135-181 + README.md:104-119)."""
    rows = []
    for i in range(n):
        yes = i % 2 == 0
        sol = ("<answer>yes</answer><location>top left</location><type>scratch</type>" if yes else "<answer>no</answer>")
        rows.append({
            "id": i, "problem": "Are there any defects in the query image?", "solution": sol,
            "image": [synthetic_image(i, image_size)],
            "prompt": [{"role": "user", "content": [{"type": "image"}, {"type": "text", "text": QUESTION_PROMPT.format(
                Question="Are there any defects in the query image?")}]}],
        })
    return rows


def format_reward(prompts, completions, current_step=0, **kwargs):
    """Structured-format check in the spirit of consistency_reward (ref: train/stage_rl/reward.py:13-30)."""
    out = []
    for c in completions:
        text = c[0]["content"] if isinstance(c, list) else c
        out.append(1.0 if re.search(r"<think>.*</think>.*<answer>.*</answer>", text, re.S) else 0.0)
    return out


def make_noise_reward(seed: int = 0):
    """Seeded N(0,1) reward so group advantages are non-degenerate on random-init models."""
    def noise_reward(prompts, completions, current_step=0, **kwargs):
        out = []
        for j, c in enumerate(completions):
            text = c[0]["content"] if isinstance(c, list) else c
            g = np.random.RandomState((zlib.crc32(text.encode()) + seed + 7919 * j + current_step) % (2 ** 31))
            out.append(float(g.randn()))
        return out
    return noise_reward
