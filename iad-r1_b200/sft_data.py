"""PA-SFT data path, natively: sharegpt rows -> chat template -> `<image>` expansion -> per-turn
token pairs -> truncation -> labels -> pad-to-8 collation, without importing LLaMA-Factory.

Restates, for the templates the reference's PA-SFT scripts select (`--template qwen2_vl` for the Qwen families,
`--template llava_next_qwen` for LLaVA-OneVision, `--template llava` for LLaVA-1.5; ref: scripts/train/PA_SFT/*.sh:33):
  * the template tables            ref: train/stage_sft/llamafactory/data/template.py:832-841, 901-913, 1121-1133
  * `<image>` expansion            ref: data/mm_plugin.py:287-311 (llava), :327-367 (llava_next), :808-823, 850-897 (qwen2_vl)
  * multi-turn pair encoding       ref: data/template.py `Template._encode` / `encode_multiturn`
  * labels + per-turn truncation   ref: data/processors/supervised.py:34-88, processor_utils.py:51-65
  * collation                      ref: data/collator.py:79-161 (pad to a multiple of 8, IGNORE_INDEX labels)
  * image pre-shrink (Q16)         ref: data/mm_plugin.py:108-123 (NEAREST above image_resolution pixels)
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

IGNORE_INDEX = -100
IMAGE_PLACEHOLDER = "<image>"


@dataclass(frozen=True)
class Template:
    name: str
    system: str            # "{content}" slots
    user: str              # includes the assistant header, as in the reference's format_user
    assistant: str
    default_system: str
    plugin: str            # "qwen2_vl" | "llava_next" | "llava"
    image_token: str
    prefix: str = ""       # in front of the first prompt ("{bos}" for the Mistral format)


_CHATML = dict(system="<|im_start|>system\n{content}<|im_end|>\n",
               user="<|im_start|>user\n{content}<|im_end|>\n<|im_start|>assistant\n",
               assistant="{content}<|im_end|>\n", default_system="You are a helpful assistant.")
TEMPLATES = {
    "qwen2_vl": Template("qwen2_vl", plugin="qwen2_vl", image_token="<|image_pad|>", **_CHATML),
    "llava_next_qwen": Template("llava_next_qwen", plugin="llava_next", image_token="<image>", **_CHATML),
    # the vicuna format (template.py:832-841): the system text runs straight into the first "USER:" (default system
    # formatter, no separator), an answer ends with the tokenizer's eos token, every <image> becomes image_seqlen tokens
    "llava": Template("llava", plugin="llava", image_token="<image>", system="{content}", user="USER: {content} ASSISTANT:",
                      assistant="{content}{eos}",
                      default_system="A chat between a curious user and an artificial intelligence assistant. "
                                     "The assistant gives helpful, detailed, and polite answers to the user's questions."),
    # LLaVA-1.6 (template.py:844-853, 868-880): the vicuna format again / Mistral's [INST] format, with the anyres plugin
    "llava_next": Template("llava_next", plugin="llava_next", image_token="<image>", system="{content}",
                           user="USER: {content} ASSISTANT:", assistant="{content}{eos}",
                           default_system="A chat between a curious user and an artificial intelligence assistant. "
                                          "The assistant gives helpful, detailed, and polite answers to the user's questions."),
    "llava_next_mistral": Template("llava_next_mistral", plugin="llava_next", image_token="<image>", system="{content}\n\n",
                                   user="[INST] {content}[/INST]", assistant=" {content}{eos}", default_system="", prefix="{bos}"),
}
FAMILY_TEMPLATE = {"qwen2_vl": "qwen2_vl", "qwen2_5_vl": "qwen2_vl", "llava_onevision": "llava_next_qwen", "llava": "llava",
                   "llava_next": "llava_next_mistral"}      # what PA_SFT_LLaVA_1_6.sh passes; `--template llava_next` for the Vicuna variant


def get_template(name: Optional[str], family: Optional[str] = None) -> Template:
    if name is None:
        name = FAMILY_TEMPLATE.get(family)
    if name not in TEMPLATES:
        raise ValueError(f"Template {name} does not exist on the B200 path (supported: {sorted(TEMPLATES)}; the LLaVA-Next "
                         f"templates belong to a model family outside it)")
    return TEMPLATES[name]


def preshrink_image(im, image_resolution: int, plugin: str):
    """Q16: any image above `image_resolution` pixels is shrunk with NEAREST before the HF processor sees it
    (mm_plugin.py:108-123); the qwen2_vl plugin also lifts sides below 28 px and caps the aspect ratio at 200 (:810-823)."""
    from PIL import Image
    if im.width * im.height > image_resolution:
        f = math.sqrt(image_resolution / (im.width * im.height))
        im = im.resize((int(im.width * f), int(im.height * f)), resample=Image.NEAREST)
    if im.mode != "RGB":
        im = im.convert("RGB")
    if plugin == "qwen2_vl":
        if min(im.width, im.height) < 28:
            im = im.resize((max(im.width, 28), max(im.height, 28)), resample=Image.NEAREST)
        if im.width / im.height > 200:
            im = im.resize((im.height * 180, im.height), resample=Image.NEAREST)
        if im.height / im.width > 200:
            im = im.resize((im.width, im.width * 180), resample=Image.NEAREST)
    return im


def load_images(example: dict, image_dir: Optional[str], image_resolution: int, plugin: str) -> list:
    from PIL import Image
    out = []
    for im in example.get("images", []) or []:
        if isinstance(im, str):
            # data/aligner.py:52-53: the path is joined with image_dir when the joined file exists
            path = os.path.join(image_dir, im) if image_dir and os.path.isfile(os.path.join(image_dir, im)) else im
            im = Image.open(path)
        out.append(preshrink_image(im, image_resolution, plugin))
    return out


def expand_image_placeholders(messages: list, seqlens: list, tpl: Template) -> list:
    """Every `<image>` in the message contents becomes its token run: `<|vision_start|>` + image_token x n + `<|vision_end|>`
    (qwen2_vl) or image_token x n (llava_next). Raises, as the reference does, when images and placeholders disagree."""
    out, used = [], 0
    for m in messages:
        content = m["content"]
        while IMAGE_PLACEHOLDER in content:
            if used >= len(seqlens):
                raise ValueError(f"`len(images)` is less than the number of {IMAGE_PLACEHOLDER} tokens.")
            run = "{{image}}" * int(seqlens[used])
            if tpl.plugin == "qwen2_vl":
                run = f"<|vision_start|>{run}<|vision_end|>"
            content = content.replace(IMAGE_PLACEHOLDER, run, 1)
            used += 1
        out.append({"role": m["role"], "content": content.replace("{{image}}", tpl.image_token)})
    if used != len(seqlens):
        raise ValueError(f"The number of images does not match the number of {IMAGE_PLACEHOLDER} tokens.")
    return out


def render_pairs(messages: list, tpl: Template, eos: str = "</s>", bos: str = "<s>") -> list:
    """[(prompt text, response text)] per (user, assistant) turn: the first prompt carries the system block (the dataset's
    system message, else the template default), every prompt ends with the assistant header, every response with the
    end-of-turn token - Template._encode with the chatml slots."""
    msgs = list(messages)
    system = tpl.default_system
    if msgs and msgs[0]["role"] == "system":
        system = msgs[0]["content"]
        msgs = msgs[1:]
    if len(msgs) % 2 != 0:
        raise ValueError("a supervised example needs alternating user / assistant messages")
    pairs = []
    for i in range(0, len(msgs), 2):
        u, a = msgs[i], msgs[i + 1]
        if u["role"] != "user" or a["role"] != "assistant":
            raise ValueError(f"expected a (user, assistant) turn at message {i}, got ({u['role']}, {a['role']})")
        prompt = ((tpl.prefix.format(bos=bos) if i == 0 else "") + (tpl.system.format(content=system) if i == 0 and system else "")
                  + tpl.user.format(content=u["content"]))
        pairs.append((prompt, tpl.assistant.format(content=a["content"], eos=eos)))
    return pairs


def infer_seqlen(source_len: int, target_len: int, cutoff_len: int):
    """Lengths of one (prompt, answer) turn after truncation to `cutoff_len` tokens (processor_utils.py:51-65)."""
    if target_len * 2 < cutoff_len:
        max_target = cutoff_len
    elif source_len * 2 < cutoff_len:
        max_target = cutoff_len - source_len
    else:
        max_target = int(cutoff_len * (target_len / (source_len + target_len)))
    new_target = min(max_target, target_len)
    new_source = min(max(cutoff_len - new_target, 0), source_len)
    return new_source, new_target


def encode_pairs(pairs_ids: list, cutoff_len: int):
    """[(source ids, target ids)] -> (input_ids, labels): each turn gets what is left of cutoff_len, split by infer_seqlen;
    prompts are masked with IGNORE_INDEX, answers are their own labels (supervised.py:50-74, train_on_prompt = mask_history = False)."""
    ids, labels, total = [], [], 0
    for src, tgt in pairs_ids:
        if total >= cutoff_len:
            break
        s_len, t_len = infer_seqlen(len(src), len(tgt), cutoff_len - total)
        src, tgt = list(src[:s_len]), list(tgt[:t_len])
        total += s_len + t_len
        ids += src + tgt
        labels += [IGNORE_INDEX] * len(src) + tgt
    return ids, labels


def collate(features: list, pad_token_id: int, pad_to_multiple_of: int = 8):
    """Right-pad a list of {input_ids, labels} to the batch maximum rounded up to a multiple of 8 (collator.py:79-161:
    DataCollatorForSeq2Seq with pad_to_multiple_of=8, label_pad_token_id=IGNORE_INDEX). Returns int64 arrays
    input_ids / labels / attention_mask of shape [B, T]."""
    T = max(len(f["input_ids"]) for f in features)
    T = (T + pad_to_multiple_of - 1) // pad_to_multiple_of * pad_to_multiple_of
    B = len(features)
    ids = np.full((B, T), pad_token_id, dtype=np.int64)
    labels = np.full((B, T), IGNORE_INDEX, dtype=np.int64)
    mask = np.zeros((B, T), dtype=np.int64)
    for i, f in enumerate(features):
        n = len(f["input_ids"])
        ids[i, :n], labels[i, :n], mask[i, :n] = f["input_ids"], f["labels"], 1
    return dict(input_ids=ids, labels=labels, attention_mask=mask)


def tokenize(processor, text: str) -> list:
    """Token ids of `text` with no special tokens added (the reference tokenises every template element on its own)."""
    if hasattr(processor, "_tok"):                     # synthetic processor of the tests / benchmarks
        return list(processor._tok(text))
    tok = getattr(processor, "tokenizer", processor)
    return list(tok(text, add_special_tokens=False)["input_ids"])
