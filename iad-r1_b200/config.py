"""Model geometry for the VLM families the reference trains. Parsed from an HF `config.json`
dict (both the transformers 4.51 flat schema the reference pins and the 5.x nested schema), never hard-coded."""
from __future__ import annotations

from dataclasses import dataclass, field, asdict


@dataclass
class TextConfig:
    vocab_size: int
    hidden_size: int
    intermediate_size: int
    num_layers: int
    num_heads: int
    num_kv_heads: int
    head_dim: int
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1e6
    mrope_section: tuple = (16, 24, 24)   # () -> plain 1-D rotary (Qwen2 under LLaVA-OneVision, LLaMA under LLaVA-1.5)
    tie_word_embeddings: bool = True
    qkv_bias: bool = True                 # Qwen2 projections carry a bias, LLaMA / Vicuna ones do not

    @property
    def qkv_dim(self) -> int:
        return (self.num_heads + 2 * self.num_kv_heads) * self.head_dim


@dataclass
class VisionConfig:
    kind: str                 # "qwen2_5_vl" (RMSNorm, SwiGLU+bias, windowed attention) | "qwen2_vl" (LayerNorm, quick-GELU MLP)
                              # | "siglip" (LLaVA-OneVision tower: LayerNorm, tanh-GELU MLP, learned position table, no rotary)
                              # | "clip" (LLaVA-1.5 tower: class token + learned positions, pre-LayerNorm, quick-GELU MLP;
                              #   features = hidden state `feature_layer` with the class token dropped)
    depth: int
    hidden_size: int
    num_heads: int
    intermediate_size: int
    out_hidden_size: int
    patch_size: int = 14
    spatial_merge_size: int = 2
    temporal_patch_size: int = 2
    in_channels: int = 3
    window_size: int = 112
    fullatt_block_indexes: tuple = (7, 15, 23, 31)
    image_size: int = 0       # siglip / clip: side of one crop in pixels (384 -> 27 x 27 = 729 tokens per crop)
    feature_layer: int = -1   # clip: index into HF's hidden_states tuple (-2 for LLaVA-1.5: the last block is not run)

    @property
    def tokens_per_crop(self) -> int:
        """Tower tokens per crop (CLIP: the class token counts, it is dropped only after the tower)."""
        return (self.image_size // self.patch_size) ** 2 + (1 if self.kind == "clip" else 0)

    @property
    def run_depth(self) -> int:
        """Blocks actually executed: hidden_states[feature_layer] of a `depth`-block encoder (tuple of depth + 1 entries)."""
        return self.depth + 1 + self.feature_layer if self.feature_layer < 0 else self.feature_layer

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_heads

    @property
    def patch_dim(self) -> int:
        return self.in_channels * self.temporal_patch_size * self.patch_size * self.patch_size

    @property
    def patch_dim_padded(self) -> int:
        """K of the patch-embedding GEMM rounded up to 8 elements (SigLIP: 3*14*14 = 588 -> 592; Qwen: 1176). CLIP keeps one
        extra column: the class embedding is column `patch_dim` of the fused patch weight and the class-token row of the
        pixel matrix is the unit vector of that column, so the class token comes out of the same GEMM."""
        return (self.patch_dim + (1 if self.kind == "clip" else 0) + 7) // 8 * 8

    @property
    def intermediate_padded(self) -> int:
        """MLP width rounded up to 8 elements so every row / half-row starts 16-byte aligned (3420 -> 3424)."""
        return (self.intermediate_size + 7) // 8 * 8


@dataclass
class VLMConfig:
    family: str               # "qwen2_5_vl" | "qwen2_vl" | "llava_onevision" | "llava" (LLaVA-1.5) | "llava_next" (LLaVA-1.6)
    text: TextConfig
    vision: VisionConfig
    image_token_id: int = 151655
    video_token_id: int = 151656
    vision_start_token_id: int = 151652
    vision_end_token_id: int = 151653
    eos_token_id: int = 151645
    pad_token_id: int = 151643
    extra: dict = field(default_factory=dict)

    def to_dict(self):
        return asdict(self)

    # ------------------------------------------------------------------------------------------------------------
    @staticmethod
    def from_hf_dict(d: dict) -> "VLMConfig":
        mt = d.get("model_type", "")
        if mt == "llava_onevision":
            return VLMConfig._from_llava_onevision(d)
        if mt in ("llava", "llava_next"):
            return VLMConfig._from_llava(d)
        if mt not in ("qwen2_5_vl", "qwen2_vl"):
            raise ValueError(f"unsupported model_type {mt!r} (supported: qwen2_5_vl, qwen2_vl, llava_onevision, llava, llava_next)")
        t = d.get("text_config") or d
        rope = t.get("rope_parameters") or t.get("rope_scaling") or d.get("rope_scaling") or {}
        theta = rope.get("rope_theta", t.get("rope_theta", d.get("rope_theta", 1e6)))
        nh = t["num_attention_heads"]
        text = TextConfig(
            vocab_size=t["vocab_size"], hidden_size=t["hidden_size"], intermediate_size=t["intermediate_size"],
            num_layers=t["num_hidden_layers"], num_heads=nh, num_kv_heads=t.get("num_key_value_heads", nh),
            head_dim=t.get("head_dim") or t["hidden_size"] // nh, rms_norm_eps=t.get("rms_norm_eps", 1e-6),
            rope_theta=float(theta), mrope_section=tuple(rope.get("mrope_section", (16, 24, 24))),
            tie_word_embeddings=bool(d.get("tie_word_embeddings", t.get("tie_word_embeddings", False))))
        v = d["vision_config"]
        if mt == "qwen2_5_vl":
            vision = VisionConfig(
                kind="qwen2_5_vl", depth=v["depth"], hidden_size=v["hidden_size"], num_heads=v["num_heads"],
                intermediate_size=v["intermediate_size"], out_hidden_size=v.get("out_hidden_size", text.hidden_size),
                patch_size=v.get("patch_size", 14), spatial_merge_size=v.get("spatial_merge_size", 2),
                temporal_patch_size=v.get("temporal_patch_size", 2), in_channels=v.get("in_channels", v.get("in_chans", 3)),
                window_size=v.get("window_size", 112), fullatt_block_indexes=tuple(v.get("fullatt_block_indexes", (7, 15, 23, 31))))
        else:
            e = v["embed_dim"]
            vision = VisionConfig(
                kind="qwen2_vl", depth=v["depth"], hidden_size=e, num_heads=v["num_heads"],
                intermediate_size=int(e * v.get("mlp_ratio", 4)), out_hidden_size=v.get("hidden_size", text.hidden_size),
                patch_size=v.get("patch_size", 14), spatial_merge_size=v.get("spatial_merge_size", 2),
                temporal_patch_size=v.get("temporal_patch_size", 2), in_channels=v.get("in_channels", v.get("in_chans", 3)),
                window_size=0, fullatt_block_indexes=())
        return VLMConfig(family=mt, text=text, vision=vision,
                         image_token_id=d.get("image_token_id", 151655), video_token_id=d.get("video_token_id", 151656),
                         vision_start_token_id=d.get("vision_start_token_id", 151652),
                         vision_end_token_id=d.get("vision_end_token_id", 151653),
                         eos_token_id=d.get("eos_token_id", 151645) if not isinstance(d.get("eos_token_id"), list) else d["eos_token_id"][0],
                         pad_token_id=d.get("pad_token_id") or 151643)

    @staticmethod
    def _from_llava_onevision(d: dict) -> "VLMConfig":
        """LlavaOnevisionConfig: Qwen2 text model (1-D rotary) + SigLIP tower + 2-layer GELU projector + anyres packing
        (HF modeling_llava_onevision.py:137-156, 292-355)."""
        t, v = d["text_config"], d["vision_config"]
        if d.get("vision_feature_select_strategy", "full") != "full" or d.get("vision_feature_layer", -1) != -1:
            raise ValueError("llava_onevision: only vision_feature_layer=-1 / select_strategy='full' is supported")
        ar = str(d.get("vision_aspect_ratio", "anyres_max_9"))
        if not ar.startswith("anyres_max_") or not ar[len("anyres_max_"):].isdigit():
            raise ValueError("llava_onevision: vision_aspect_ratio must be 'anyres_max_<N>'")
        nh = t["num_attention_heads"]
        rope = t.get("rope_parameters") or {}
        text = TextConfig(
            vocab_size=t["vocab_size"], hidden_size=t["hidden_size"], intermediate_size=t["intermediate_size"],
            num_layers=t["num_hidden_layers"], num_heads=nh, num_kv_heads=t.get("num_key_value_heads", nh),
            head_dim=t.get("head_dim") or t["hidden_size"] // nh, rms_norm_eps=t.get("rms_norm_eps", 1e-6),
            rope_theta=float(rope.get("rope_theta", t.get("rope_theta", 1e6))), mrope_section=(),
            tie_word_embeddings=bool(d.get("tie_word_embeddings", t.get("tie_word_embeddings", False))))
        vision = VisionConfig(
            kind="siglip", depth=v["num_hidden_layers"], hidden_size=v["hidden_size"], num_heads=v["num_attention_heads"],
            intermediate_size=v["intermediate_size"], out_hidden_size=text.hidden_size, patch_size=v.get("patch_size", 14),
            spatial_merge_size=1, temporal_patch_size=1, in_channels=v.get("num_channels", 3), window_size=0,
            fullatt_block_indexes=(), image_size=v.get("image_size", 384))
        eos = d.get("eos_token_id", t.get("eos_token_id", 151645))
        return VLMConfig(family="llava_onevision", text=text, vision=vision,
                         image_token_id=d.get("image_token_index", d.get("image_token_id", 151646)),
                         video_token_id=d.get("video_token_index", d.get("video_token_id", 151647)),
                         vision_start_token_id=-1, vision_end_token_id=-1,
                         eos_token_id=eos[0] if isinstance(eos, list) else eos,
                         pad_token_id=d.get("pad_token_id") or t.get("pad_token_id") or 151643,
                         extra={"image_grid_pinpoints": [list(x) for x in d["image_grid_pinpoints"]],
                                "vision_layer_norm_eps": v.get("layer_norm_eps", 1e-6),
                                "anyres_max": int(ar[len("anyres_max_"):])})

    @staticmethod
    def _from_llava(d: dict) -> "VLMConfig":
        """LlavaConfig (LLaVA-1.5, ref: sc_grpo_trainer.py:134-136) and LlavaNextConfig (LLaVA-1.6, :131-133): CLIP ViT tower
        (class token, pre-LayerNorm, quick-GELU), features = hidden_states[vision_feature_layer] without the class token,
        2-layer GELU projector, LLaMA / Vicuna / Mistral decoder (no projection biases, 1-D rotary); LLaVA-Next adds anyres crops
        packed with `image_newline` (no feature shrink). HF modeling_llava.py, modeling_llava_next.py, modeling_clip.py."""
        t, v = d["text_config"], d["vision_config"]
        family = d.get("model_type", "llava")
        if d.get("vision_feature_select_strategy", "default") != "default":
            raise ValueError("llava: only vision_feature_select_strategy='default' (class token dropped) is supported")
        if isinstance(d.get("vision_feature_layer", -2), (list, tuple)):
            raise ValueError("llava: a single vision_feature_layer is supported")
        if d.get("projector_hidden_act", "gelu") != "gelu" or d.get("multimodal_projector_bias", True) is not True:
            raise ValueError("llava: the projector must be Linear - GELU - Linear with biases")
        nh = t.get("num_attention_heads", 32)
        hidden = t.get("hidden_size", 4096)
        rope = t.get("rope_parameters") or {}
        text = TextConfig(
            vocab_size=t.get("vocab_size", 32064), hidden_size=hidden, intermediate_size=t.get("intermediate_size", 11008),
            num_layers=t.get("num_hidden_layers", 32), num_heads=nh, num_kv_heads=t.get("num_key_value_heads", nh),
            head_dim=t.get("head_dim") or hidden // nh, rms_norm_eps=t.get("rms_norm_eps", 1e-5),
            rope_theta=float(rope.get("rope_theta", t.get("rope_theta", 10000.0))), mrope_section=(),
            tie_word_embeddings=bool(d.get("tie_word_embeddings", t.get("tie_word_embeddings", False))),
            qkv_bias=bool(t.get("attention_bias", False)))
        if t.get("mlp_bias", False):
            raise ValueError("llava: MLP biases in the language model are not supported")
        vision = VisionConfig(
            kind="clip", depth=v.get("num_hidden_layers", 24), hidden_size=v.get("hidden_size", 1024),
            num_heads=v.get("num_attention_heads", 16), intermediate_size=v.get("intermediate_size", 4096),
            out_hidden_size=text.hidden_size, patch_size=v.get("patch_size", 14), spatial_merge_size=1, temporal_patch_size=1,
            in_channels=v.get("num_channels", 3), window_size=0, fullatt_block_indexes=(), image_size=v.get("image_size", 336),
            feature_layer=int(d.get("vision_feature_layer", -2)))
        if v.get("hidden_act", "quick_gelu") != "quick_gelu":
            raise ValueError("llava: the CLIP tower must use quick_gelu")
        eos = d.get("eos_token_id", t.get("eos_token_id", 2))
        extra = {"vision_layer_norm_eps": v.get("layer_norm_eps", 1e-5), "text_model_type": t.get("model_type", "llama")}
        if family == "llava_next":
            if not d.get("use_image_newline_parameter", True):
                raise ValueError("llava_next: only use_image_newline_parameter=True is supported")
            extra["image_grid_pinpoints"] = [list(x) for x in d.get("image_grid_pinpoints") or
                                             [[336, 672], [672, 336], [672, 672], [1008, 336], [336, 1008]]]
        return VLMConfig(family=family, text=text, vision=vision,
                         image_token_id=d.get("image_token_index", d.get("image_token_id", 32000)), video_token_id=-1,
                         vision_start_token_id=-1, vision_end_token_id=-1,
                         eos_token_id=eos[0] if isinstance(eos, list) else eos,
                         pad_token_id=d.get("pad_token_id") if d.get("pad_token_id") is not None else t.get("pad_token_id", 32001),
                         extra=extra)

    def to_hf_dict(self) -> dict:
        """config.json in the 4.51-era flat schema the reference's `from_pretrained` expects."""
        t, v = self.text, self.vision
        if self.family in ("llava", "llava_next"):
            nxt = self.family == "llava_next"
            return {
                "model_type": self.family,
                "architectures": ["LlavaNextForConditionalGeneration" if nxt else "LlavaForConditionalGeneration"],
                **({"image_grid_pinpoints": self.extra["image_grid_pinpoints"], "use_image_newline_parameter": True} if nxt else {}),
                "image_token_index": self.image_token_id, "vision_feature_layer": v.feature_layer,
                "vision_feature_select_strategy": "default", "projector_hidden_act": "gelu", "multimodal_projector_bias": True,
                "image_seq_length": v.tokens_per_crop - 1, "tie_word_embeddings": t.tie_word_embeddings,
                "pad_token_id": self.pad_token_id, "eos_token_id": self.eos_token_id,
                "vision_config": {"model_type": "clip_vision_model", "hidden_size": v.hidden_size,
                                  "intermediate_size": v.intermediate_size, "num_hidden_layers": v.depth,
                                  "num_attention_heads": v.num_heads, "image_size": v.image_size, "patch_size": v.patch_size,
                                  "num_channels": v.in_channels, "hidden_act": "quick_gelu", "projection_dim": v.hidden_size,
                                  "layer_norm_eps": self.extra.get("vision_layer_norm_eps", 1e-5)},
                "text_config": {"model_type": self.extra.get("text_model_type", "llama"), "vocab_size": t.vocab_size,
                                "hidden_size": t.hidden_size,
                                "intermediate_size": t.intermediate_size, "num_hidden_layers": t.num_layers,
                                "num_attention_heads": t.num_heads, "num_key_value_heads": t.num_kv_heads,
                                "head_dim": t.head_dim, "rms_norm_eps": t.rms_norm_eps, "rope_theta": t.rope_theta,
                                "max_position_embeddings": 4096, "hidden_act": "silu", "attention_bias": t.qkv_bias,
                                "mlp_bias": False, "tie_word_embeddings": t.tie_word_embeddings,
                                "eos_token_id": self.eos_token_id, "pad_token_id": self.pad_token_id}}
        if self.family == "llava_onevision":
            return {
                "model_type": "llava_onevision", "architectures": ["LlavaOnevisionForConditionalGeneration"],
                "image_token_index": self.image_token_id, "video_token_index": self.video_token_id,
                "image_grid_pinpoints": self.extra["image_grid_pinpoints"], "vision_feature_layer": -1,
                "vision_feature_select_strategy": "full", "vision_aspect_ratio": f"anyres_max_{self.extra.get('anyres_max', 9)}",
                "projector_hidden_act": "gelu", "multimodal_projector_bias": True,
                "tie_word_embeddings": t.tie_word_embeddings, "torch_dtype": "bfloat16",
                "text_config": {"model_type": "qwen2", "vocab_size": t.vocab_size, "hidden_size": t.hidden_size,
                                "intermediate_size": t.intermediate_size, "num_hidden_layers": t.num_layers,
                                "num_attention_heads": t.num_heads, "num_key_value_heads": t.num_kv_heads,
                                "rms_norm_eps": t.rms_norm_eps, "rope_theta": t.rope_theta, "hidden_act": "silu",
                                "max_position_embeddings": 32768, "tie_word_embeddings": t.tie_word_embeddings,
                                "eos_token_id": self.eos_token_id, "pad_token_id": self.pad_token_id},
                "vision_config": {"model_type": "siglip_vision_model", "hidden_size": v.hidden_size,
                                  "intermediate_size": v.intermediate_size, "num_hidden_layers": v.depth,
                                  "num_attention_heads": v.num_heads, "patch_size": v.patch_size, "image_size": v.image_size,
                                  "num_channels": v.in_channels, "hidden_act": "gelu_pytorch_tanh",
                                  "layer_norm_eps": self.extra.get("vision_layer_norm_eps", 1e-6), "vision_use_head": False},
            }
        d = {
            "model_type": self.family,
            "architectures": ["Qwen2_5_VLForConditionalGeneration" if self.family == "qwen2_5_vl" else "Qwen2VLForConditionalGeneration"],
            "vocab_size": t.vocab_size, "hidden_size": t.hidden_size, "intermediate_size": t.intermediate_size,
            "num_hidden_layers": t.num_layers, "num_attention_heads": t.num_heads, "num_key_value_heads": t.num_kv_heads,
            "rms_norm_eps": t.rms_norm_eps, "rope_theta": t.rope_theta,
            "rope_scaling": {"type": "mrope", "mrope_section": list(t.mrope_section)},
            "tie_word_embeddings": t.tie_word_embeddings, "hidden_act": "silu", "torch_dtype": "bfloat16",
            "image_token_id": self.image_token_id, "video_token_id": self.video_token_id,
            "vision_start_token_id": self.vision_start_token_id, "vision_end_token_id": self.vision_end_token_id,
            "eos_token_id": self.eos_token_id, "pad_token_id": self.pad_token_id, "max_position_embeddings": 128000,
        }
        if self.family == "qwen2_5_vl":
            d["vision_config"] = {
                "model_type": "qwen2_5_vl", "depth": v.depth, "hidden_size": v.hidden_size, "num_heads": v.num_heads,
                "intermediate_size": v.intermediate_size, "out_hidden_size": v.out_hidden_size, "patch_size": v.patch_size,
                "spatial_merge_size": v.spatial_merge_size, "temporal_patch_size": v.temporal_patch_size,
                "in_chans": v.in_channels, "window_size": v.window_size, "fullatt_block_indexes": list(v.fullatt_block_indexes),
                "hidden_act": "silu", "tokens_per_second": 2}
        else:
            d["vision_config"] = {
                "model_type": "qwen2_vl", "depth": v.depth, "embed_dim": v.hidden_size, "num_heads": v.num_heads,
                "mlp_ratio": v.intermediate_size // v.hidden_size, "hidden_size": v.out_hidden_size,
                "patch_size": v.patch_size, "spatial_merge_size": v.spatial_merge_size,
                "temporal_patch_size": v.temporal_patch_size, "in_chans": v.in_channels, "hidden_act": "quick_gelu"}
        return d


# ---- the public checkpoints' geometry (used for synthetic random-init benchmarks) ---------------
def _qwen25_vision(out_hidden):
    return VisionConfig(kind="qwen2_5_vl", depth=32, hidden_size=1280, num_heads=16, intermediate_size=3420,
                        out_hidden_size=out_hidden)


PRESETS = {
    "qwen2.5-vl-3b": lambda: VLMConfig("qwen2_5_vl", TextConfig(151936, 2048, 11008, 36, 16, 2, 128, tie_word_embeddings=True),
                                       _qwen25_vision(2048)),
    "qwen2.5-vl-7b": lambda: VLMConfig("qwen2_5_vl", TextConfig(152064, 3584, 18944, 28, 28, 4, 128, tie_word_embeddings=False),
                                       _qwen25_vision(3584)),
    "qwen2-vl-2b": lambda: VLMConfig("qwen2_vl", TextConfig(151936, 1536, 8960, 28, 12, 2, 128, tie_word_embeddings=True),
                                     VisionConfig(kind="qwen2_vl", depth=32, hidden_size=1280, num_heads=16,
                                                  intermediate_size=5120, out_hidden_size=1536, window_size=0,
                                                  fullatt_block_indexes=())),
    # LLaVA-OneVision-Qwen2-0.5B-SI: Qwen2-0.5B + SigLIP-so400m/14-384 (26 encoder layers used), anyres_max_9
    "llava-ov-0.5b": lambda: VLMConfig(
        "llava_onevision", TextConfig(151936, 896, 4864, 24, 14, 2, 64, mrope_section=(), tie_word_embeddings=True),
        VisionConfig(kind="siglip", depth=26, hidden_size=1152, num_heads=16, intermediate_size=4304, out_hidden_size=896,
                     patch_size=14, spatial_merge_size=1, temporal_patch_size=1, window_size=0, fullatt_block_indexes=(),
                     image_size=384),
        image_token_id=151646, video_token_id=151647, vision_start_token_id=-1, vision_end_token_id=-1,
        extra={"image_grid_pinpoints": [[384 * a, 384 * b] for a in range(1, 7) for b in range(1, 7)],
               "vision_layer_norm_eps": 1e-6}),
    # llava-1.5-7b-hf: Vicuna-7B (LLaMA: MHA 32 x 128, no projection biases, untied head) + CLIP ViT-L/14-336 (24 blocks,
    # hidden_states[-2] = 23 blocks run, class token dropped: 576 image tokens)
    "llava-1.5-7b": lambda: VLMConfig(
        "llava", TextConfig(32064, 4096, 11008, 32, 32, 32, 128, rms_norm_eps=1e-5, rope_theta=10000.0, mrope_section=(),
                            tie_word_embeddings=False, qkv_bias=False),
        VisionConfig(kind="clip", depth=24, hidden_size=1024, num_heads=16, intermediate_size=4096, out_hidden_size=4096,
                     patch_size=14, spatial_merge_size=1, temporal_patch_size=1, window_size=0, fullatt_block_indexes=(),
                     image_size=336, feature_layer=-2),
        image_token_id=32000, video_token_id=-1, vision_start_token_id=-1, vision_end_token_id=-1, eos_token_id=2,
        pad_token_id=32001, extra={"vision_layer_norm_eps": 1e-5}),
    # llava-v1.6-mistral-7b-hf: Mistral-7B (GQA 32 : 8, hd 128, I 14336, no biases) + the same CLIP tower, anyres crops of 336
    "llava-1.6-mistral-7b": lambda: VLMConfig(
        "llava_next", TextConfig(32064, 4096, 14336, 32, 32, 8, 128, rms_norm_eps=1e-5, rope_theta=1000000.0, mrope_section=(),
                                 tie_word_embeddings=False, qkv_bias=False),
        VisionConfig(kind="clip", depth=24, hidden_size=1024, num_heads=16, intermediate_size=4096, out_hidden_size=4096,
                     patch_size=14, spatial_merge_size=1, temporal_patch_size=1, window_size=0, fullatt_block_indexes=(),
                     image_size=336, feature_layer=-2),
        image_token_id=32000, video_token_id=-1, vision_start_token_id=-1, vision_end_token_id=-1, eos_token_id=2,
        pad_token_id=32001,
        extra={"vision_layer_norm_eps": 1e-5, "text_model_type": "mistral",
               "image_grid_pinpoints": [[336, 672], [672, 336], [672, 672], [1008, 336], [336, 1008]]}),
}


def depth_reduced(cfg: VLMConfig, text_layers: int = 2, vision_depth: int = 2) -> VLMConfig:
    """TRUE widths (hidden sizes, head dims, GQA ratios, MLP widths, vocabulary) at reduced depth: the geometry of the
    true-width parity tests and of the CPU baseline's bounded sample. Qwen2.5-VL keeps one windowed and one full-attention
    vision block."""
    import copy
    c = copy.deepcopy(cfg)
    c.text.num_layers = text_layers
    c.vision.depth = vision_depth
    c.vision.fullatt_block_indexes = (vision_depth - 1,) if c.vision.kind == "qwen2_5_vl" else ()
    return c


def tiny_config(family: str = "qwen2_5_vl") -> VLMConfig:
    """A few-hundred-k-parameter twin with every structural feature of the real model (GQA, M-RoPE sections, windowed +
    full vision blocks, ragged MLP width, tied head) - the parity-test geometry (tests/golden)."""
    text = TextConfig(vocab_size=1024, hidden_size=128, intermediate_size=256, num_layers=2, num_heads=4, num_kv_heads=2,
                      head_dim=32, mrope_section=(4, 6, 6), tie_word_embeddings=True)
    if family == "llava_onevision":
        # head_dim 24 (not a power of two, like SigLIP's 72), 4 x 4 tokens per 56-pixel crop, 2 x 2 anyres grid
        # text head_dim 64 (as Qwen2-0.5B): the rollout takes the tensor-core decode-attention path on the twin too
        text = TextConfig(vocab_size=1024, hidden_size=128, intermediate_size=256, num_layers=2, num_heads=2,
                          num_kv_heads=1, head_dim=64, mrope_section=(), tie_word_embeddings=True)
        vis = VisionConfig(kind="siglip", depth=2, hidden_size=96, num_heads=4, intermediate_size=200, out_hidden_size=128,
                           patch_size=14, spatial_merge_size=1, temporal_patch_size=1, window_size=0,
                           fullatt_block_indexes=(), image_size=56)
        return VLMConfig(family, text, vis, image_token_id=1001, video_token_id=1002, vision_start_token_id=-1,
                         vision_end_token_id=-1, eos_token_id=1005, pad_token_id=1006,
                         extra={"image_grid_pinpoints": [[56, 56], [56, 112], [112, 56], [112, 112], [168, 56], [56, 168]],
                                "vision_layer_norm_eps": 1e-6})
    if family == "llava_next":
        # LLaVA-1.6 twin: CLIP tower as the 1.5 twin, anyres crops of 56 pixels (4 x 4 patch tokens + class token each) packed with
        # image_newline, grouped-query decoder without biases (Mistral-like), head_dim 64
        text = TextConfig(vocab_size=1024, hidden_size=128, intermediate_size=256, num_layers=2, num_heads=2,
                          num_kv_heads=1, head_dim=64, rms_norm_eps=1e-5, rope_theta=10000.0, mrope_section=(),
                          tie_word_embeddings=False, qkv_bias=False)
        vis = VisionConfig(kind="clip", depth=3, hidden_size=64, num_heads=4, intermediate_size=160, out_hidden_size=128,
                           patch_size=14, spatial_merge_size=1, temporal_patch_size=1, window_size=0,
                           fullatt_block_indexes=(), image_size=56, feature_layer=-2)
        return VLMConfig(family, text, vis, image_token_id=1001, video_token_id=-1, vision_start_token_id=-1,
                         vision_end_token_id=-1, eos_token_id=1005, pad_token_id=1006,
                         extra={"vision_layer_norm_eps": 1e-5, "text_model_type": "mistral",
                                "image_grid_pinpoints": [[56, 112], [112, 56], [112, 112], [168, 56], [56, 168]]})
    if family == "llava":
        # LLaVA-1.5 twin: multi-head attention without projection biases (LLaMA), CLIP tower of 3 blocks whose last one is
        # skipped (vision_feature_layer -2), class token + 4 x 4 patches of a 56-pixel image
        text = TextConfig(vocab_size=1024, hidden_size=128, intermediate_size=256, num_layers=2, num_heads=2,
                          num_kv_heads=2, head_dim=64, rms_norm_eps=1e-5, rope_theta=10000.0, mrope_section=(),
                          tie_word_embeddings=False, qkv_bias=False)
        vis = VisionConfig(kind="clip", depth=3, hidden_size=64, num_heads=4, intermediate_size=160, out_hidden_size=128,
                           patch_size=14, spatial_merge_size=1, temporal_patch_size=1, window_size=0,
                           fullatt_block_indexes=(), image_size=56, feature_layer=-2)
        return VLMConfig(family, text, vis, image_token_id=1001, video_token_id=-1, vision_start_token_id=-1,
                         vision_end_token_id=-1, eos_token_id=1005, pad_token_id=1006, extra={"vision_layer_norm_eps": 1e-5})
    if family == "qwen2_5_vl":
        vis = VisionConfig(kind="qwen2_5_vl", depth=2, hidden_size=64, num_heads=4, intermediate_size=108,
                           out_hidden_size=128, window_size=56, fullatt_block_indexes=(1,))
    else:
        vis = VisionConfig(kind="qwen2_vl", depth=2, hidden_size=64, num_heads=4, intermediate_size=256,
                           out_hidden_size=128, window_size=0, fullatt_block_indexes=())
    return VLMConfig(family, text, vis, image_token_id=1001, video_token_id=1002, vision_start_token_id=1003,
                     vision_end_token_id=1004, eos_token_id=1005, pad_token_id=1006)
