"""GPU image preprocessing for the Qwen2-VL / Qwen2.5-VL families: the host only decodes the image to
uint8 RGB and computes the target size; resize, normalisation and the patch layout run in csrc/preprocess.cu and land
in HBM as the bf16 `pixel_values` the vision tower reads.

Replaces the image half of the reference's processor call (ref: train/stage_rl/trainer/sc_grpo_trainer.py:614-621, which runs
HF `Qwen2VLImageProcessor` on the CPU and then tiles the pixels G times, :624-628). LLaVA-OneVision's anyres crop / pad
pipeline stays on its HF image processor."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import lib as L

OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def smart_resize(height: int, width: int, factor: int = 28, min_pixels: int = 56 * 56, max_pixels: int = 14 * 14 * 4 * 1280):
    """Target size of the Qwen2-VL processors: both sides multiples of `factor`, area within [min_pixels, max_pixels], aspect
    ratio kept as far as the rounding allows (HF image_processing_qwen2_vl.py:62-88; checked against it on CPU)."""
    if max(height, width) / min(height, width) > 200:
        raise ValueError(f"absolute aspect ratio must be smaller than 200, got {max(height, width) / min(height, width)}")
    h_bar = round(height / factor) * factor
    w_bar = round(width / factor) * factor
    if h_bar * w_bar > max_pixels:
        beta = math.sqrt((height * width) / max_pixels)
        h_bar = max(factor, math.floor(height / beta / factor) * factor)
        w_bar = max(factor, math.floor(width / beta / factor) * factor)
    elif h_bar * w_bar < min_pixels:
        beta = math.sqrt(min_pixels / (height * width))
        h_bar = math.ceil(height * beta / factor) * factor
        w_bar = math.ceil(width * beta / factor) * factor
    return h_bar, w_bar


def to_rgb_u8(image) -> np.ndarray:
    """PIL image / HWC array -> contiguous uint8 [H, W, 3]."""
    if hasattr(image, "convert"):
        image = np.asarray(image.convert("RGB"))
    a = np.asarray(image)
    if a.ndim != 3 or a.shape[2] != 3 or a.dtype != np.uint8:
        raise ValueError(f"expected an RGB uint8 image, got shape {a.shape} dtype {a.dtype}")
    return np.ascontiguousarray(a) if a.flags.writeable else np.array(a, order="C")


def qwen_preprocess_gpu(images: list, vision_cfg, device, min_pixels: int, max_pixels: int,
                        mean=OPENAI_CLIP_MEAN, std=OPENAI_CLIP_STD):
    """images (PIL / uint8 HWC) -> (pixel_values bf16 [sum Np, C * tps * ps * ps] on `device`, image_grid_thw list)."""
    ps, m, tps = vision_cfg.patch_size, vision_cfg.spatial_merge_size, vision_cfg.temporal_patch_size
    outs, grids = [], []
    mean_c, std_c = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    for im in images:
        a = to_rgb_u8(im)
        h, w = a.shape[:2]
        th, tw = smart_resize(h, w, ps * m, min_pixels, max_pixels)
        src = torch.from_numpy(a).pin_memory().to(device, non_blocking=True)
        gh, gw = th // ps, tw // ps
        out = torch.empty(gh * gw, 3 * tps * ps * ps, dtype=torch.bfloat16, device=device)
        scratch = None
        if (th, tw) != (h, w):
            scratch = torch.empty(h * tw * 3 + th * tw * 3, dtype=torch.uint8, device=device)
        L.check(L.lib().iadr1_image_preprocess_qwen(src.data_ptr(), h, w, th, tw, ps, m, tps, mean_c, std_c,
                                                    None if scratch is None else scratch.data_ptr(), out.data_ptr(),
                                                    L.stream_ptr()), "image_preprocess_qwen")
        outs.append(out)
        grids.append([1, gh, gw])
    return (torch.cat(outs, 0) if len(outs) > 1 else outs[0]), grids
