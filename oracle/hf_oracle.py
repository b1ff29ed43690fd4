"""ORACLE (test infrastructure, never on the product path): the HF Transformers implementation the reference executes
for `model(**inputs).logits` (ref: train/stage_rl/trainer/sc_grpo_trainer.py:117-137 `from_pretrained`, :505 forward),
built from an explicit config with seeded random weights and run on CPU.

The model arithmetic is NOT in /root/reference: it lives in the third-party dependency `transformers==4.51.3`
(ref: requirements.txt:205), absent from the repo; this container has transformers 5.5.0, whose Qwen2-VL / Qwen2.5-VL
modules compute the same function given explicit `position_ids` (SURVEY.md §8c lists the version-skew traps; we
always pass 4.51.3-semantics position ids). Parity status: the reference's own tests hold no golden vectors for this
path ("parity unpinned") - the pins are (a) this HF implementation run here, frozen into
tests/golden/*.pt by oracle/make_golden.py, and (b) the KATs of the vendored TRL helpers (oracle/grpo_ref.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import torch


def hf_config(cfg, attn_implementation="eager"):
    """iad_r1_b200.config.VLMConfig -> transformers config object (5.x nested schema)."""
    t, v = cfg.text, cfg.vision
    text = dict(vocab_size=t.vocab_size, hidden_size=t.hidden_size, intermediate_size=t.intermediate_size,
                num_hidden_layers=t.num_layers, num_attention_heads=t.num_heads, num_key_value_heads=t.num_kv_heads,
                max_position_embeddings=32768, rms_norm_eps=t.rms_norm_eps, tie_word_embeddings=t.tie_word_embeddings,
                rope_parameters=dict(rope_type="default", rope_theta=t.rope_theta, mrope_section=list(t.mrope_section)))
    common = dict(image_token_id=cfg.image_token_id, video_token_id=cfg.video_token_id,
                  vision_start_token_id=cfg.vision_start_token_id, vision_end_token_id=cfg.vision_end_token_id,
                  tie_word_embeddings=t.tie_word_embeddings)
    if cfg.family == "llava_onevision":
        from transformers import LlavaOnevisionConfig
        c = LlavaOnevisionConfig(
            text_config=dict(model_type="qwen2", vocab_size=t.vocab_size, hidden_size=t.hidden_size,
                             intermediate_size=t.intermediate_size, num_hidden_layers=t.num_layers,
                             num_attention_heads=t.num_heads, num_key_value_heads=t.num_kv_heads,
                             max_position_embeddings=32768, rms_norm_eps=t.rms_norm_eps, rope_theta=t.rope_theta,
                             tie_word_embeddings=t.tie_word_embeddings, use_sliding_window=False),
            vision_config=dict(model_type="siglip_vision_model", hidden_size=v.hidden_size, intermediate_size=v.intermediate_size,
                               num_hidden_layers=v.depth, num_attention_heads=v.num_heads, patch_size=v.patch_size,
                               image_size=v.image_size, num_channels=v.in_channels, hidden_act="gelu_pytorch_tanh",
                               layer_norm_eps=cfg.extra.get("vision_layer_norm_eps", 1e-6), vision_use_head=False),
            image_token_index=cfg.image_token_id, video_token_index=cfg.video_token_id,
            image_grid_pinpoints=cfg.extra["image_grid_pinpoints"], vision_feature_layer=-1,
            vision_feature_select_strategy="full", vision_aspect_ratio=f"anyres_max_{cfg.extra.get('anyres_max', 9)}",
            projector_hidden_act="gelu",
            multimodal_projector_bias=True, tie_word_embeddings=t.tie_word_embeddings)
        c._attn_implementation = attn_implementation
        return c
    if cfg.family in ("llava", "llava_next"):
        from transformers import LlavaConfig, LlavaNextConfig
        nxt = cfg.family == "llava_next"
        extra_kw = dict(image_grid_pinpoints=cfg.extra["image_grid_pinpoints"], use_image_newline_parameter=True) if nxt else {}
        c = (LlavaNextConfig if nxt else LlavaConfig)(
            **extra_kw,
            text_config=dict(model_type=cfg.extra.get("text_model_type", "llama"), vocab_size=t.vocab_size, hidden_size=t.hidden_size,
                             intermediate_size=t.intermediate_size, num_hidden_layers=t.num_layers,
                             num_attention_heads=t.num_heads, num_key_value_heads=t.num_kv_heads, head_dim=t.head_dim,
                             max_position_embeddings=4096, rms_norm_eps=t.rms_norm_eps, rope_theta=t.rope_theta,
                             attention_bias=t.qkv_bias, mlp_bias=False, tie_word_embeddings=t.tie_word_embeddings),
            vision_config=dict(model_type="clip_vision_model", hidden_size=v.hidden_size, intermediate_size=v.intermediate_size,
                               num_hidden_layers=v.depth, num_attention_heads=v.num_heads, patch_size=v.patch_size,
                               image_size=v.image_size, num_channels=v.in_channels, hidden_act="quick_gelu",
                               projection_dim=v.hidden_size, layer_norm_eps=cfg.extra.get("vision_layer_norm_eps", 1e-5)),
            image_token_index=cfg.image_token_id, vision_feature_layer=v.feature_layer,
            vision_feature_select_strategy="default", projector_hidden_act="gelu", multimodal_projector_bias=True,
            image_seq_length=v.tokens_per_crop - 1, tie_word_embeddings=t.tie_word_embeddings)
        c._attn_implementation = attn_implementation
        return c
    if cfg.family == "qwen2_5_vl":
        from transformers import Qwen2_5_VLConfig
        vis = dict(depth=v.depth, hidden_size=v.hidden_size, intermediate_size=v.intermediate_size, num_heads=v.num_heads,
                   out_hidden_size=v.out_hidden_size, patch_size=v.patch_size, spatial_merge_size=v.spatial_merge_size,
                   temporal_patch_size=v.temporal_patch_size, window_size=v.window_size,
                   fullatt_block_indexes=list(v.fullatt_block_indexes), in_channels=v.in_channels)
        c = Qwen2_5_VLConfig(text_config=text, vision_config=vis, **common)
    else:
        from transformers import Qwen2VLConfig
        vis = dict(depth=v.depth, embed_dim=v.hidden_size, hidden_size=v.out_hidden_size, num_heads=v.num_heads,
                   mlp_ratio=v.intermediate_size // v.hidden_size, patch_size=v.patch_size,
                   spatial_merge_size=v.spatial_merge_size, temporal_patch_size=v.temporal_patch_size,
                   in_channels=v.in_channels)
        c = Qwen2VLConfig(text_config=text, vision_config=vis, **common)
    c._attn_implementation = attn_implementation
    return c


def build_hf_model(cfg, seed=0, dtype=torch.float32, attn_implementation="eager"):
    """Random-init HF model (initializer_range 0.02, HF `_init_weights`), weights rounded to bf16 values so that the
    fp32 oracle and the bf16 product path hold bit-identical parameters."""
    if cfg.family == "llava_onevision":
        from transformers import LlavaOnevisionForConditionalGeneration as Cls
    elif cfg.family == "llava":
        from transformers import LlavaForConditionalGeneration as Cls
    elif cfg.family == "llava_next":
        from transformers import LlavaNextForConditionalGeneration as Cls
    elif cfg.family == "qwen2_5_vl":
        from transformers import Qwen2_5_VLForConditionalGeneration as Cls
    else:
        from transformers import Qwen2VLForConditionalGeneration as Cls
    torch.manual_seed(seed)
    m = Cls(hf_config(cfg, attn_implementation))
    with torch.no_grad():
        gen = torch.Generator().manual_seed(seed + 1)
        for n, p in m.named_parameters():
            if n.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=gen) * 0.02)       # non-zero biases exercise the bias paths
            elif p.dim() == 1:
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=gen))  # norm gains away from 1
            p.copy_(p.to(torch.bfloat16).to(p.dtype))
    return m.to(dtype).eval()


def hf_logits(model, input_ids, pixel_values, grid_thw, position_ids, attention_mask=None, logits_to_keep=0):
    """`model(**inputs).logits` with explicit 4.51.3-semantics position ids. `logits_to_keep=n` keeps only the last n
    positions' logits (same values; used at true vocabulary width, where [G, T, 151936] fp32 would not fit the test box)."""
    kw = dict(input_ids=input_ids, position_ids=position_ids, use_cache=False)
    if logits_to_keep:
        kw["logits_to_keep"] = int(logits_to_keep)
    if attention_mask is not None:
        kw["attention_mask"] = attention_mask
    if pixel_values is not None:
        kw["pixel_values"] = pixel_values.to(next(model.parameters()).dtype)
        kw["image_grid_thw"] = grid_thw
    return model(**kw).logits


def hf_logits_llava(model, input_ids, pixel_values, image_sizes, position_ids, attention_mask=None, logits_to_keep=0):
    """LLaVA-OneVision: pixel_values [B, n_crops, 3, S, S] (crop 0 = base crop), image_sizes [B, 2] (H, W)."""
    kw = dict(input_ids=input_ids, position_ids=position_ids, use_cache=False,
              pixel_values=pixel_values.to(next(model.parameters()).dtype), image_sizes=image_sizes)
    if logits_to_keep:
        kw["logits_to_keep"] = int(logits_to_keep)
    if attention_mask is not None:
        kw["attention_mask"] = attention_mask
    return model(**kw).logits


def hf_logits_llava15(model, input_ids, pixel_values, position_ids, attention_mask=None, logits_to_keep=0):
    """LLaVA-1.5: pixel_values [n_images, 3, S, S] (one crop per image, one image per row of input_ids)."""
    kw = dict(input_ids=input_ids, position_ids=position_ids, use_cache=False,
              pixel_values=pixel_values.to(next(model.parameters()).dtype))
    if logits_to_keep:
        kw["logits_to_keep"] = int(logits_to_keep)
    if attention_mask is not None:
        kw["attention_mask"] = attention_mask
    return model(**kw).logits


def completion_logps(tail_logits, input_ids, C):
    """Log-probs of the last C tokens of every row from the logits of the last C + 1 positions (`logits_to_keep=C + 1`):
    the same slice `_get_per_token_logps` + `[:, P - 1:]` keeps (sc_grpo_trainer.py:505-514, 734)."""
    lp = tail_logits[:, :-1, :].float().log_softmax(-1)
    return torch.gather(lp, 2, input_ids[:, -C:].unsqueeze(-1)).squeeze(-1)


def per_token_logps(logits, input_ids):
    """Restates `_get_per_token_logps` (sc_grpo_trainer.py:505-514): drop the last logit, row-wise log_softmax + gather,
    in the dtype of `logits` (the reference runs it in bf16; call with fp32 logits for the tight oracle)."""
    logits = logits[:, :-1, :]
    ids = input_ids[:, 1:]
    out = []
    for lr, ir in zip(logits, ids):
        lp = lr.log_softmax(dim=-1)
        out.append(torch.gather(lp, 1, ir.unsqueeze(1)).squeeze(1))
    return torch.stack(out)
