"""Host-side integer bookkeeping of the VLM forward: M-RoPE position ids, rotary tables, the vision tower's window
permutation / attention segments / 2-D rotary positions, and the image-token -> image-row index map.

All of this is tiny, data-independent (a function of token ids and `image_grid_thw` only) and runs once per batch;
results are cached per grid. Semantics follow transformers 4.51.3, the version the reference pins
(ref: requirements.txt:205), re-stated from the HF sources cited per function.
"""
from __future__ import annotations

import functools

import math

import numpy as np
import torch

from .config import TextConfig, VisionConfig, VLMConfig


# ---- M-RoPE position ids (Qwen2_5_VLForConditionalGeneration.get_rope_index, 4.51.3 semantics) ------------------------
def mrope_position_ids(input_ids: np.ndarray, grid_thw: list, cfg: VLMConfig, attention_mask: np.ndarray | None = None,
                       prompt_len: int | None = None):
    """input_ids [B, T] int -> (position_ids [3, B, T] int64, rope_deltas [B] int64).

    Text runs count up on all three axes; a still image occupying t*(h/m)*(w/m) tokens from base b gets
    (b, b + row, b + col); text resumes at max + 1. Masked (pad) positions get 1 (4.51.3 filler, masked anyway).
    `prompt_len`: image placeholders are only recognised in the first `prompt_len` columns - a SAMPLED completion token
    that happens to equal `image_token_id` is plain text (HF would raise on the feature/token count mismatch).
    """
    B, T = input_ids.shape
    m = cfg.vision.spatial_merge_size
    pos = np.ones((3, B, T), dtype=np.int64)
    deltas = np.zeros((B,), dtype=np.int64)
    img_cursor = 0
    for b in range(B):
        keep = np.ones(T, dtype=bool) if attention_mask is None else attention_mask[b].astype(bool)
        ids = input_ids[b][keep]
        in_prompt = (np.arange(T) < (T if prompt_len is None else prompt_len))[keep]
        out = np.zeros((3, len(ids)), dtype=np.int64)
        i, nxt = 0, 0
        while i < len(ids):
            if ids[i] == cfg.image_token_id and in_prompt[i]:
                t, h, w = grid_thw[img_cursor]
                img_cursor += 1
                gh, gw = h // m, w // m
                n = t * gh * gw
                if i + n > len(ids) or not (ids[i:i + n] == cfg.image_token_id).all():
                    raise ValueError("image token span does not match image_grid_thw (prompt truncated into an image?)")
                tt = np.repeat(np.arange(t), gh * gw) * 0  # still images: temporal index 0 for every token
                hh = np.tile(np.repeat(np.arange(gh), gw), t)
                ww = np.tile(np.tile(np.arange(gw), gh), t)
                out[0, i:i + n] = nxt + tt
                out[1, i:i + n] = nxt + hh
                out[2, i:i + n] = nxt + ww
                nxt = nxt + max(gh, gw, 1 if t > 0 else 0)
                i += n
            else:
                out[:, i] = nxt
                nxt += 1
                i += 1
        pos[:, b, keep] = out
        deltas[b] = (out.max() + 1 - T) if len(ids) else 0
    return pos, deltas


def text_rope_tables(position_ids: torch.Tensor, t: TextConfig, device):
    """cos/sin fp32 [B*T, hd] for the decoder. Qwen2_5_VLRotaryEmbedding.forward (modeling_qwen2_5_vl.py:595-608) +
    the section select of apply_multimodal_rotary_pos_emb (:659-665): band i of the duplicated `mrope_section` list
    takes its angle from axis i % 3."""
    pos = position_ids.to(device=device, dtype=torch.float32)  # [3, B, T]
    hd = t.head_dim
    inv_freq = 1.0 / (t.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64, device=device).float() / hd))
    freqs = pos[..., None] * inv_freq  # [3, B, T, hd/2]
    if t.mrope_section:
        sel = torch.cat([torch.full((s,), i % 3, dtype=torch.long, device=device)
                         for i, s in enumerate(t.mrope_section)])  # [hd/2] axis per frequency band
        f = torch.gather(freqs.permute(1, 2, 3, 0), 3, sel.view(1, 1, -1, 1).expand(freqs.shape[1], freqs.shape[2], -1, 1))
        f = f.squeeze(-1)  # [B, T, hd/2]
    else:
        f = freqs[0]
    emb = torch.cat((f, f), dim=-1).reshape(-1, hd)
    return emb.cos().contiguous(), emb.sin().contiguous()


def decode_rope_table(t: TextConfig, max_pos: int, device):
    """[max_pos, hd] fp32 tables for the rollout (generated tokens have equal positions on the three axes)."""
    hd = t.head_dim
    inv_freq = 1.0 / (t.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64, device=device).float() / hd))
    f = torch.arange(max_pos, device=device, dtype=torch.float32)[:, None] * inv_freq
    emb = torch.cat((f, f), dim=-1)
    return emb.cos().contiguous(), emb.sin().contiguous()


# ---- vision tower geometry -------------------------------------------------------------------------------------------------
class VisionGeometry:
    """Everything the vision tower derives from `image_grid_thw` (Qwen2_5_VisionTransformerPretrainedModel.forward,
    modeling_qwen2_5_vl.py:455-518; rot_pos_emb :382-409; get_window_index :411-453)."""

    def __init__(self, v: VisionConfig, grid_thw: list, device):
        m = v.spatial_merge_size
        unit = m * m
        self.n_patches = int(sum(t * h * w for t, h, w in grid_thw))
        self.n_tokens = self.n_patches // unit
        # 2-D rotary position of each patch, in merge-block-major patch order
        pos = []
        for t, h, w in grid_thw:
            hp = np.arange(h)[:, None].repeat(w, 1).reshape(h // m, m, w // m, m).transpose(0, 2, 1, 3).reshape(-1)
            wp = np.arange(w)[None, :].repeat(h, 0).reshape(h // m, m, w // m, m).transpose(0, 2, 1, 3).reshape(-1)
            pos.append(np.tile(np.stack([hp, wp], -1), (t, 1)))
        pos = np.concatenate(pos, 0)  # [Np, 2]
        hd = v.head_dim
        dim = hd // 2
        inv_freq = 1.0 / (10000.0 ** (np.arange(0, dim, 2, dtype=np.float32) / dim))
        fr = np.concatenate([pos[:, 0:1] * inv_freq[None], pos[:, 1:2] * inv_freq[None]], -1).astype(np.float32)  # [Np, hd/2]
        # full-attention segments: one per (image, frame)
        seg = []
        for t, h, w in grid_thw:
            seg += [h * w] * t
        cu_full = np.concatenate([[0], np.cumsum(seg)]).astype(np.int64)
        if v.kind == "qwen2_5_vl" and v.window_size > 0:
            win = v.window_size // m // v.patch_size
            index, cu_win, base = [], [0], 0
            for t, h, w in grid_thw:
                gh, gw = h // m, w // m
                idx = np.arange(t * gh * gw).reshape(t, gh, gw)
                ph, pw = win - gh % win, win - gw % win
                nh, nw = (gh + ph) // win, (gw + pw) // win
                padded = np.full((t, gh + ph, gw + pw), -100, dtype=np.int64)
                padded[:, :gh, :gw] = idx
                padded = padded.reshape(t, nh, win, nw, win).transpose(0, 1, 3, 2, 4).reshape(t, nh * nw, win, win)
                lens = (padded != -100).sum((2, 3)).reshape(-1)
                flat = padded.reshape(-1)
                index.append(flat[flat != -100] + base)
                cu_win += (np.cumsum(lens) * unit + cu_win[-1]).tolist()
                base += t * gh * gw
            window_index = np.concatenate(index)
            cu_win = np.array(sorted(set(cu_win)), dtype=np.int64)  # unique_consecutive on a non-decreasing list
            # reorder rotary angles the same way as the hidden states (units of `unit` patches)
            fr = fr.reshape(self.n_tokens, unit, -1)[window_index].reshape(self.n_patches, -1)
        else:
            window_index = None
            cu_win = cu_full
        emb = np.concatenate([fr, fr], -1)
        self.cos = torch.from_numpy(np.cos(emb)).to(device).contiguous()
        self.sin = torch.from_numpy(np.sin(emb)).to(device).contiguous()

        def ranges(cu):
            lo = np.zeros(self.n_patches, dtype=np.int32)
            hi = np.zeros(self.n_patches, dtype=np.int32)
            for a, b in zip(cu[:-1], cu[1:]):
                lo[a:b], hi[a:b] = a, b
            return torch.from_numpy(lo).to(device), torch.from_numpy(hi).to(device)

        self.full_lo, self.full_hi = ranges(cu_full)
        self.win_lo, self.win_hi = ranges(cu_win)

        # equal-length segments (every image the same grid / every window full): attention runs as a BATCH of
        # [segment x segment] problems instead of one masked [Np x Np] product (448 x 448: 16 windows of 64 patches)
        def uniform(cu):
            d = np.diff(cu)
            return int(d[0]) if len(d) and (d == d[0]).all() and d[0] % 8 == 0 else 0

        self.full_seg, self.win_seg = uniform(cu_full), uniform(cu_win)
        # ragged case: attention runs image by image (patch range + key ranges relative to the image), never as one masked
        # [Np_total x Np_total] product over all images of a batch
        self.image_ranges, a = [], 0
        for t, h, w in grid_thw:
            self.image_ranges.append((a, a + t * h * w))
            a += t * h * w
        self._rel = {}
        self.seg_ranges = {}
        for L_ in {self.full_seg, self.win_seg} - {0}:
            self.seg_ranges[L_] = (torch.zeros(L_, dtype=torch.int32, device=device),
                                   torch.full((L_,), L_, dtype=torch.int32, device=device))
        self._abs = {"full": (self.full_lo, self.full_hi), "win": (self.win_lo, self.win_hi)}
        if window_index is not None:
            self.window_index = torch.from_numpy(window_index.astype(np.int32)).to(device)
            self.reverse_index = torch.from_numpy(np.argsort(window_index).astype(np.int32)).to(device)
        else:
            self.window_index = self.reverse_index = None


def _relative_ranges(self, kind: str, j: int):
    """(lo, hi) key ranges of image j's patches, relative to the image's first patch."""
    key = (kind, j)
    r = self._rel.get(key)
    if r is None:
        a, b = self.image_ranges[j]
        lo, hi = self._abs[kind]
        r = self._rel[key] = ((lo[a:b] - a).contiguous(), (hi[a:b] - a).contiguous())
    return r


VisionGeometry.relative_ranges = _relative_ranges

_GEOM_CACHE: dict = {}


def vision_geometry(v: VisionConfig, grid_thw, device) -> VisionGeometry:
    key = (v.kind, v.window_size, v.spatial_merge_size, v.patch_size, v.head_dim, tuple(map(tuple, grid_thw)), str(device))
    g = _GEOM_CACHE.get(key)
    if g is None:
        if len(_GEOM_CACHE) > 64:
            _GEOM_CACHE.clear()
        g = _GEOM_CACHE[key] = VisionGeometry(v, [tuple(int(x) for x in r) for r in grid_thw], device)
    return g


def embed_source_index(input_ids: np.ndarray, image_token_id: int, rows_share_images: bool, n_image_tokens: int,
                       prompt_len: int | None = None):
    """[B*T] int32: token id for text, -1 - (image embedding row) for image placeholders (`masked_scatter`,
    modeling_qwen2_5_vl.py:1298-1307). With `rows_share_images` every row consumes the same image rows from 0 (the G
    completions of one prompt: the reference tiles pixel_values G times, sc_grpo_trainer.py:624-628, we don't)."""
    B, T = input_ids.shape
    out = input_ids.astype(np.int64).copy()
    cursor = 0
    for b in range(B):
        is_img = input_ids[b] == image_token_id
        if prompt_len is not None:
            is_img &= np.arange(T) < prompt_len
        n = int(is_img.sum())
        if rows_share_images:
            cursor = 0
        if cursor + n > n_image_tokens:
            raise ValueError(f"image features and image tokens do not match: tokens {cursor + n}, features {n_image_tokens}")
        out[b, is_img] = -1 - (cursor + np.arange(n))
        cursor += n
    return out.reshape(-1).astype(np.int32)


# ---- LLaVA-OneVision: anyres geometry (HF image_processing_utils.select_best_resolution, modeling_llava_onevision.py
# :159-187 get_anyres_image_grid_shape, :227-267 unpad_image, :292-355 pack_image_features) ---------------------------------
def select_best_resolution(original_size, possible_resolutions):
    """original_size (h, w); returns the (h, w) pinpoint with the largest effective resolution, ties -> least waste."""
    oh, ow = original_size
    best, best_eff, best_waste = None, -1, float("inf")
    for h, w in possible_resolutions:
        scale = min(w / ow, h / oh)
        dw, dh = int(ow * scale), int(oh * scale)
        eff = min(dw * dh, ow * oh)
        waste = w * h - eff
        if eff > best_eff or (eff == best_eff and waste < best_waste):
            best, best_eff, best_waste = (h, w), eff, waste
    return best


def llava_image_layout(cfg: VLMConfig, image_hw):
    """For one image of size (H, W): (n_crops incl. the base crop, crop grid (gh, gw), kept feature-grid rows / cols after
    unpadding as slices). Whether the kept feature map is then shrunk (anyres_max_N) is `llava_shrink`."""
    v = cfg.vision
    side = v.image_size // v.patch_size                       # 27 feature rows / cols per crop
    bh, bw = select_best_resolution(image_hw, cfg.extra["image_grid_pinpoints"])
    gh, gw = bh // v.image_size, bw // v.image_size
    cur_h, cur_w = gh * side, gw * side
    oh, ow = image_hw
    if ow / oh > cur_w / cur_h:
        new_h = int(round(oh * (cur_w / ow), 7))
        pad = (cur_h - new_h) // 2
        rows, cols = (pad, cur_h - pad), (0, cur_w)
    else:
        new_w = int(round(ow * (cur_h / oh), 7))
        pad = (cur_w - new_w) // 2
        rows, cols = (0, cur_h), (pad, cur_w - pad)
    return gh * gw + 1, (gh, gw), rows, cols


def llava_shrink(cfg: VLMConfig, image_hw):
    """LLaVA-OneVision `anyres_max_N` (HF modeling_llava_onevision.py:328-335): when the kept feature map of an image holds more
    than 1.1^2 x N crops' worth of tokens it is bilinearly resized to (kh // ratio, kw // ratio), ratio = sqrt(kh kw / (N side^2)).
    Returns (kh, kw, kh2, kw2) or None (LLaVA-Next never shrinks)."""
    if cfg.family != "llava_onevision":
        return None
    v = cfg.vision
    side = v.image_size // v.patch_size
    _, _, rows, cols = llava_image_layout(cfg, image_hw)
    kh, kw = rows[1] - rows[0], cols[1] - cols[0]
    ratio = math.sqrt(kh * kw / (int(cfg.extra.get("anyres_max", 9)) * side * side))
    if ratio <= 1.1:
        return None
    return kh, kw, int(kh // ratio), int(kw // ratio)


def llava_pack_index(cfg: VLMConfig, image_hw):
    """Row map of pack_image_features for one image: packed token i takes projected-feature row idx[i] (relative to the
    image's first feature row; crop 0 = base crop, then the grid crops row-major), or the `image_newline` vector where
    idx[i] == -1: [base crop tokens | for every kept feature-grid row: its kept columns, newline]."""
    v = cfg.vision
    side = v.image_size // v.patch_size
    tpc = side * side
    stride, off = v.tokens_per_crop, v.tokens_per_crop - tpc     # CLIP (LLaVA-Next): a crop's row 0 is its class token, never packed
    n_crops, (gh, gw), rows, cols = llava_image_layout(cfg, image_hw)
    idx = [np.arange(tpc, dtype=np.int64) + off]
    ys = np.arange(rows[0], rows[1])
    xs = np.arange(cols[0], cols[1])
    crop = 1 + (ys[:, None] // side) * gw + (xs[None, :] // side)
    body = crop * stride + off + (ys[:, None] % side) * side + (xs[None, :] % side)
    body = np.concatenate([body, np.full((len(ys), 1), -1, dtype=np.int64)], 1)
    idx.append(body.reshape(-1))
    return np.concatenate(idx).astype(np.int32), n_crops


def image_token_count(cfg: VLMConfig, grid_entry) -> int:
    """Number of placeholder tokens one image expands to. grid_entry: (t, h, w) patch grid for the Qwen families;
    (n_crops, H, W) with the ORIGINAL image size in pixels for LLaVA-OneVision."""
    if cfg.family in ("llava_onevision", "llava_next"):
        hw = (int(grid_entry[1]), int(grid_entry[2]))
        sh = llava_shrink(cfg, hw)
        if sh is not None:      # base crop + the resized feature map with one image_newline per row
            side = cfg.vision.image_size // cfg.vision.patch_size
            return side * side + sh[2] * (sh[3] + 1)
        return len(llava_pack_index(cfg, hw)[0])
    if cfg.family == "llava":          # LLaVA-1.5: one 336-pixel crop, class token dropped (select strategy "default")
        return cfg.vision.tokens_per_crop - 1
    t, h, w = grid_entry
    return int(t * h * w) // cfg.vision.spatial_merge_size ** 2


def position_ids(input_ids: np.ndarray, grid_thw: list, cfg: VLMConfig, attention_mask: np.ndarray | None = None,
                 prompt_len: int | None = None):
    """Family dispatch: M-RoPE ids for Qwen2(.5)-VL; plain 1-D positions (cumulative count of unmasked tokens, the
    Qwen2 text model under LLaVA-OneVision) replicated on the three axes otherwise. Returns ([3, B, T], deltas [B])."""
    if cfg.family not in ("llava_onevision", "llava", "llava_next"):
        return mrope_position_ids(input_ids, grid_thw, cfg, attention_mask, prompt_len)
    B, T = input_ids.shape
    am = np.ones((B, T), dtype=np.int64) if attention_mask is None else attention_mask.astype(np.int64)
    p = np.cumsum(am, 1) - 1
    p[am == 0] = 1
    pos = np.broadcast_to(p[None], (3, B, T)).astype(np.int64).copy()
    deltas = (am.sum(1) - T).astype(np.int64)
    return pos, deltas


class SiglipGeometry:
    """What the LLaVA-OneVision vision path derives from the image sizes: per-crop attention segments (SigLIP attends
    within one 384 x 384 crop), the tiled position-table index, and the anyres packing index over all images."""

    def __init__(self, cfg: VLMConfig, grid: list, device):
        tpc = cfg.vision.tokens_per_crop
        packs, base, n_crops_total = [], 0, 0
        # anyres_max shrink (LLaVA-OneVision, rare: images spanning more than N crops): the kernels pack the UNSHRUNK map; the
        # images listed here are resized afterwards (model.VLM._apply_shrink): (offset in the packed stream, packed length,
        # offset in the final stream, final length, base tokens, kh, kw, kh2, kw2)
        self.shrinks, u_off, f_off = [], 0, 0
        for n_crops, h, w in grid:
            if cfg.family == "llava":      # LLaVA-1.5: every patch token of the single crop, class token (row 0) dropped
                idx, n = np.arange(1, tpc, dtype=np.int64), 1
            else:
                idx, n = llava_pack_index(cfg, (int(h), int(w)))
            if n != int(n_crops):
                raise ValueError(f"image of size {(h, w)} needs {n} crops, processor supplied {n_crops}")
            packs.append(np.where(idx >= 0, idx + base, -1))
            sh = llava_shrink(cfg, (int(h), int(w))) if cfg.family != "llava" else None
            n_f = image_token_count(cfg, (n, h, w))
            self.shrinks.append((u_off, len(idx), f_off, n_f, (cfg.vision.image_size // cfg.vision.patch_size) ** 2, sh))
            u_off += len(idx)
            f_off += n_f
            base += n * tpc
            n_crops_total += n
        self.n_final = f_off
        self.has_shrink = any(s_[5] is not None for s_ in self.shrinks)
        self.n_crops = n_crops_total
        self.n_patches = n_crops_total * tpc
        pack = np.concatenate(packs) if packs else np.zeros(0, dtype=np.int64)
        self.n_tokens = len(pack)
        self.pack_index = torch.from_numpy(pack.astype(np.int32)).to(device)
        lo = (np.arange(self.n_patches) // tpc * tpc).astype(np.int32)
        self.full_lo = torch.from_numpy(lo).to(device)
        self.full_hi = torch.from_numpy(lo + tpc).to(device)
        self.pos_index = torch.from_numpy((np.arange(self.n_patches) % tpc).astype(np.int32)).to(device)   # learned-position row


def siglip_geometry(cfg: VLMConfig, grid, device) -> SiglipGeometry:
    key = ("siglip", cfg.family, cfg.vision.kind, cfg.vision.image_size, cfg.vision.patch_size, tuple(map(tuple, grid)), str(device),
           str(cfg.extra.get("image_grid_pinpoints")), cfg.extra.get("anyres_max"), cfg.vision.feature_layer)
    g = _GEOM_CACHE.get(key)
    if g is None:
        if len(_GEOM_CACHE) > 64:
            _GEOM_CACHE.clear()
        g = _GEOM_CACHE[key] = SiglipGeometry(cfg, [tuple(int(x) for x in r) for r in grid], device)
    return g


def patchify_crops(pixel_values: torch.Tensor, patch: int) -> torch.Tensor:
    """[n_crops, C, S, S] -> [n_crops * (S/patch)^2, C * patch * patch]: the rows the SigLIP Conv2d(k = stride = patch)
    patch embedding multiplies, in (crop, grid row, grid col) order with (channel, py, px) columns - so that the
    convolution is ONE GEMM against the flattened kernel (modeling_siglip.py:116-186)."""
    n, c, s, _ = pixel_values.shape
    g = s // patch                 # a VALID convolution: 384 = 27 * 14 + 6, the last 6 pixel rows / columns are unused
    x = pixel_values[:, :, :g * patch, :g * patch].reshape(n, c, g, patch, g, patch).permute(0, 2, 4, 1, 3, 5)
    return x.reshape(n * g * g, c * patch * patch).contiguous()


def clip_pixel_rows(pixel_values: torch.Tensor, v: VisionConfig) -> torch.Tensor:
    """[n_images, C, S, S] (CLIPImageProcessor output) -> the rows of the CLIP patch-embedding GEMM, `tokens_per_crop` per image:
    row 0 of every image is the class token = the unit vector of column `patch_dim` (that column of the fused patch weight holds
    the class embedding, HF modeling_clip.py CLIPVisionEmbeddings), rows 1.. are the patches in (grid row, grid col) order with
    (channel, py, px) columns; width `patch_dim_padded`."""
    n = pixel_values.shape[0]
    patches = patchify_crops(pixel_values, v.patch_size).view(n, v.tokens_per_crop - 1, v.patch_dim)
    rows = torch.zeros(n, v.tokens_per_crop, v.patch_dim_padded, dtype=patches.dtype, device=patches.device)
    rows[:, 1:, :v.patch_dim] = patches
    rows[:, 0, v.patch_dim] = 1
    return rows.view(n * v.tokens_per_crop, v.patch_dim_padded)


def vision_inputs_from_processor(cfg: VLMConfig, enc) -> tuple:
    """(pixel_values, grid) in the layout the vision kernels take, from a processor's output mapping.
    Qwen processors already emit patch rows + `image_grid_thw`. LLaVA-OneVision processors emit crops
    `pixel_values [n_images, n_crops_max, C, S, S]` (padded over images) + `image_sizes [n_images, 2]`: each image keeps
    its own crops, which are cut into patch rows; grid entries become (n_crops, H, W)."""
    pv = enc.get("pixel_values") if hasattr(enc, "get") else None
    if pv is None:
        return None, None
    if "image_grid_thw" in enc:
        g = enc["image_grid_thw"]
        return pv, (g.tolist() if torch.is_tensor(g) else g)
    if cfg.family == "llava":              # LlavaProcessor: pixel_values [n_images, C, S, S], one crop per image
        if pv.dim() == 3:
            pv = pv[None]
        S = cfg.vision.image_size
        return clip_pixel_rows(pv, cfg.vision), [(1, S, S)] * pv.shape[0]
    if cfg.family not in ("llava_onevision", "llava_next") or "image_sizes" not in enc:
        raise ValueError("processor output has pixel_values but neither image_grid_thw nor image_sizes")
    sizes = enc["image_sizes"].tolist() if torch.is_tensor(enc["image_sizes"]) else enc["image_sizes"]
    if pv.dim() == 4:
        pv = pv[None]
    rows, grid = [], []
    for i, (h, w) in enumerate(sizes):
        n, _, _, _ = llava_image_layout(cfg, (int(h), int(w)))
        rows.append(clip_pixel_rows(pv[i, :n], cfg.vision) if cfg.vision.kind == "clip"
                    else patchify_crops(pv[i, :n], cfg.vision.patch_size))
        grid.append((n, int(h), int(w)))
    return torch.cat(rows, 0), grid
