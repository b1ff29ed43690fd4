"""Flat parameter storage for the VLM: ONE bf16 buffer holds every weight (16-byte aligned sub-tensors, fused
q|k|v and gate|up matrices), with matching flat fp32 buffers for gradients, master weights and Adam moments.

Why flat: the gradient all-reduce (the only NCCL call), the global-norm reduction and the fused AdamW
(csrc/optimizer.cu) each run over one contiguous range instead of ~800 tensors; 180 GB of HBM3e holds policy +
reference + fp32 master/moments/gradients of the 3B and 7B models without sharding (DESIGN.md "HBM layout").

Names follow the HF checkpoints (`model.language_model.layers.N.self_attn.q_proj.weight` ...), which are exposed
as *views* of the fused tensors by `hf_named_tensors`, so `from_pretrained`-style directories round-trip.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from .config import VLMConfig


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


class ParamStore:
    @staticmethod
    def numel_of(cfg: VLMConfig) -> int:
        """Elements of the flat parameter buffer for `cfg` without allocating anything (memory planning)."""
        return ParamStore(cfg, "meta").numel

    @staticmethod
    def plan_moment_dtype(cfg: VLMConfig, device, with_reference: bool, requested: str = "auto") -> torch.dtype:
        """fp32 Adam moments (the reference's regime) unless the unsharded state - bf16 policy (+ reference) + fp32 gradient,
        master, exp_avg, exp_avg_sq - would not leave a fifth of the device for activations and the KV cache; then bf16 moments
        with stochastic rounding (Qwen2.5-VL-7B: 166 GB -> 133 GB on a 180 GB B200)."""
        if requested in ("fp32", "float32"):
            return torch.float32
        if requested in ("bf16", "bfloat16"):
            return torch.bfloat16
        n = ParamStore.numel_of(cfg)
        total = torch.cuda.get_device_properties(device).total_memory if torch.cuda.is_available() else 180e9
        return torch.bfloat16 if n * (2 + (2 if with_reference else 0) + 16) > 0.8 * total else torch.float32

    def __init__(self, cfg: VLMConfig, device, with_grads: bool = False, with_optimizer: bool = False,
                 moment_dtype: torch.dtype = torch.float32):
        self.cfg = cfg
        self.device = torch.device(device)
        self.shapes: "OrderedDict[str, tuple]" = OrderedDict()
        self.decay: dict = {}
        self._declare()
        # decayed (matrices) first, then non-decayed (norms, biases): two contiguous AdamW ranges.
        order = [n for n in self.shapes if self.decay[n]] + [n for n in self.shapes if not self.decay[n]]
        self.offsets = {}
        off = 0
        for n in order:
            self.offsets[n] = off
            numel = 1
            for s in self.shapes[n]:
                numel *= s
            off += _pad8(numel)
            if self.decay[n]:
                self.n_decay = off
        self.numel = off
        self.flat = torch.zeros(self.numel, dtype=torch.bfloat16, device=self.device)
        self.p = self._views(self.flat)
        self.grad_flat = self.master = self.exp_avg = self.exp_avg_sq = None
        self.g = None
        if with_grads:
            self.grad_flat = torch.zeros(self.numel, dtype=torch.float32, device=self.device)
            self.g = self._views(self.grad_flat)
        if with_optimizer:
            self.master = torch.zeros(self.numel, dtype=torch.float32, device=self.device)
            # fp32 moments (the reference's regime, Q18) or bf16 under stochastic rounding (optimizer.cu: 8 B/param less)
            self.exp_avg = torch.zeros(self.numel, dtype=moment_dtype, device=self.device)
            self.exp_avg_sq = torch.zeros(self.numel, dtype=moment_dtype, device=self.device)

    # ---------------------------------------------------------------------------------------------------------------
    def _add(self, name, shape, decay):
        self.shapes[name] = tuple(shape)
        self.decay[name] = decay

    def _declare(self):
        t, v = self.cfg.text, self.cfg.vision
        E, Ip = v.hidden_size, v.intermediate_padded
        self._add("visual.patch_embed.weight", (E, v.patch_dim_padded), True)   # pad columns stay exactly zero
        if v.kind == "siglip":
            self._add("visual.patch_embed.bias", (E,), False)
            # HF Trainer's decay rule excludes only biases and norm weights (get_decay_parameter_names): the nn.Embedding position
            # table, class_embedding and image_newline all decay
            self._add("visual.pos_embed.weight", (v.tokens_per_crop, E), True)
        if v.kind == "clip":
            # column `patch_dim` of the patch weight IS the class embedding (see VisionConfig.patch_dim_padded)
            self._add("visual.pos_embed.weight", (v.tokens_per_crop, E), True)
            self._add("visual.pre_ln.weight", (E,), False)
            self._add("visual.pre_ln.bias", (E,), False)
        for i in range(v.depth):
            b = f"visual.blocks.{i}."
            self._add(b + "qkv.weight", (3 * E, E), True)
            self._add(b + "qkv.bias", (3 * E,), False)
            self._add(b + "proj.weight", (E, E), True)
            self._add(b + "proj.bias", (E,), False)
            self._add(b + "norm1.weight", (E,), False)
            self._add(b + "norm2.weight", (E,), False)
            if v.kind == "qwen2_5_vl":
                self._add(b + "gate_up.weight", (2 * Ip, E), True)
                self._add(b + "gate_up.bias", (2 * Ip,), False)
                self._add(b + "down.weight", (E, Ip), True)
                self._add(b + "down.bias", (E,), False)
            else:
                self._add(b + "norm1.bias", (E,), False)
                self._add(b + "norm2.bias", (E,), False)
                self._add(b + "fc1.weight", (Ip, E), True)
                self._add(b + "fc1.bias", (Ip,), False)
                self._add(b + "fc2.weight", (E, Ip), True)
                self._add(b + "fc2.bias", (E,), False)
        m = v.spatial_merge_size ** 2 * E
        if v.kind == "clip":
            # post_layernorm is not on LLaVA's path (features come from an encoder hidden state); kept for round trips
            self._add("visual.post_layernorm.weight", (E,), False)
            self._add("visual.post_layernorm.bias", (E,), False)
            self._add("visual.merger.fc1.weight", (v.out_hidden_size, E), True)
            self._add("visual.merger.fc1.bias", (v.out_hidden_size,), False)
            self._add("visual.merger.fc2.weight", (v.out_hidden_size, v.out_hidden_size), True)
            self._add("visual.merger.fc2.bias", (v.out_hidden_size,), False)
            if self.cfg.family == "llava_next":
                self._add("image_newline", (v.out_hidden_size,), True)       # the anyres row separator (a bare nn.Parameter: decays)
        elif v.kind == "siglip":
            # SigLIP's post_layernorm is NOT on the path (features = last encoder layer output); kept for round trips.
            self._add("visual.post_layernorm.weight", (E,), False)
            self._add("visual.post_layernorm.bias", (E,), False)
            # LlavaOnevisionMultiModalProjector: Linear(E -> H) GELU Linear(H -> H); + the anyres row separator
            self._add("visual.merger.fc1.weight", (v.out_hidden_size, E), True)
            self._add("visual.merger.fc1.bias", (v.out_hidden_size,), False)
            self._add("visual.merger.fc2.weight", (v.out_hidden_size, v.out_hidden_size), True)
            self._add("visual.merger.fc2.bias", (v.out_hidden_size,), False)
            self._add("image_newline", (v.out_hidden_size,), True)
        else:
            self._add("visual.merger.ln_q.weight", (E,), False)
            if v.kind == "qwen2_vl":
                self._add("visual.merger.ln_q.bias", (E,), False)
            self._add("visual.merger.fc1.weight", (m, m), True)
            self._add("visual.merger.fc1.bias", (m,), False)
            self._add("visual.merger.fc2.weight", (v.out_hidden_size, m), True)
            self._add("visual.merger.fc2.bias", (v.out_hidden_size,), False)
        self._add("embed_tokens.weight", (t.vocab_size, t.hidden_size), True)
        for i in range(t.num_layers):
            b = f"layers.{i}."
            self._add(b + "qkv.weight", (t.qkv_dim, t.hidden_size), True)
            if t.qkv_bias:
                self._add(b + "qkv.bias", (t.qkv_dim,), False)
            self._add(b + "o.weight", (t.hidden_size, t.num_heads * t.head_dim), True)
            self._add(b + "gate_up.weight", (2 * t.intermediate_size, t.hidden_size), True)
            self._add(b + "down.weight", (t.hidden_size, t.intermediate_size), True)
            self._add(b + "ln1.weight", (t.hidden_size,), False)
            self._add(b + "ln2.weight", (t.hidden_size,), False)
        self._add("norm.weight", (t.hidden_size,), False)
        if not t.tie_word_embeddings:
            self._add("lm_head.weight", (t.vocab_size, t.hidden_size), True)

    def layer_matrix_range(self, i: int):
        """[lo, hi) of decoder layer i's weight matrices in the flat buffers (contiguous: qkv, o, gate_up, down)."""
        t = self.cfg.text
        lo = self.offsets[f"layers.{i}.qkv.weight"]
        if i + 1 < t.num_layers:
            hi = self.offsets[f"layers.{i + 1}.qkv.weight"]
        else:
            hi = self.offsets["lm_head.weight"] if "lm_head.weight" in self.offsets else self.n_decay
        return lo, hi

    def _views(self, flat):
        out = {}
        for n, shape in self.shapes.items():
            numel = 1
            for s in shape:
                numel *= s
            out[n] = flat[self.offsets[n]: self.offsets[n] + numel].view(shape)
        return out

    @property
    def lm_head(self):
        return self.p["embed_tokens.weight"] if self.cfg.text.tie_word_embeddings else self.p["lm_head.weight"]

    @property
    def lm_head_grad(self):
        return self.g["embed_tokens.weight"] if self.cfg.text.tie_word_embeddings else self.g["lm_head.weight"]

    # ---------------------------------------------------------------------------------------------------------------
    def hf_named_tensors(self, which: str = "p"):
        """Yield (HF checkpoint name, view) pairs: fused tensors are split back into q/k/v, gate/up; MLP padding dropped.
        Names use the 4.51-era layout (`visual.*`, `model.*`, `lm_head.weight`)."""
        src = {"p": self.p, "g": self.g}[which]
        t, v = self.cfg.text, self.cfg.vision
        E, I, Ip = v.hidden_size, v.intermediate_size, v.intermediate_padded
        if v.kind in ("siglip", "clip"):
            yield from self._hf_named_siglip(src)
            yield from self._hf_named_text(src, "language_model.model.", "language_model.lm_head.weight")
            return
        yield "visual.patch_embed.proj.weight", src["visual.patch_embed.weight"]
        for i in range(v.depth):
            b, hb = f"visual.blocks.{i}.", f"visual.blocks.{i}."
            yield hb + "attn.qkv.weight", src[b + "qkv.weight"]
            yield hb + "attn.qkv.bias", src[b + "qkv.bias"]
            yield hb + "attn.proj.weight", src[b + "proj.weight"]
            yield hb + "attn.proj.bias", src[b + "proj.bias"]
            yield hb + "norm1.weight", src[b + "norm1.weight"]
            yield hb + "norm2.weight", src[b + "norm2.weight"]
            if v.kind == "qwen2_5_vl":
                yield hb + "mlp.gate_proj.weight", src[b + "gate_up.weight"][:I]
                yield hb + "mlp.up_proj.weight", src[b + "gate_up.weight"][Ip:Ip + I]
                yield hb + "mlp.gate_proj.bias", src[b + "gate_up.bias"][:I]
                yield hb + "mlp.up_proj.bias", src[b + "gate_up.bias"][Ip:Ip + I]
                yield hb + "mlp.down_proj.weight", src[b + "down.weight"][:, :I]
                yield hb + "mlp.down_proj.bias", src[b + "down.bias"]
            else:
                yield hb + "norm1.bias", src[b + "norm1.bias"]
                yield hb + "norm2.bias", src[b + "norm2.bias"]
                yield hb + "mlp.fc1.weight", src[b + "fc1.weight"][:I]
                yield hb + "mlp.fc1.bias", src[b + "fc1.bias"][:I]
                yield hb + "mlp.fc2.weight", src[b + "fc2.weight"][:, :I]
                yield hb + "mlp.fc2.bias", src[b + "fc2.bias"]
        yield "visual.merger.ln_q.weight", src["visual.merger.ln_q.weight"]
        if v.kind == "qwen2_vl":
            yield "visual.merger.ln_q.bias", src["visual.merger.ln_q.bias"]
        yield "visual.merger.mlp.0.weight", src["visual.merger.fc1.weight"]
        yield "visual.merger.mlp.0.bias", src["visual.merger.fc1.bias"]
        yield "visual.merger.mlp.2.weight", src["visual.merger.fc2.weight"]
        yield "visual.merger.mlp.2.bias", src["visual.merger.fc2.bias"]
        yield from self._hf_named_text(src, "model.", "lm_head.weight")

    def _hf_named_siglip(self, src):
        """LLaVA-OneVision 4.51-era names: vision_tower.vision_model.*, multi_modal_projector.*, image_newline."""
        v = self.cfg.vision
        E = v.hidden_size
        vb = "vision_tower.vision_model."
        yield vb + "embeddings.patch_embedding.weight", src["visual.patch_embed.weight"][:, :v.patch_dim]
        if v.kind == "clip":
            yield vb + "embeddings.class_embedding", src["visual.patch_embed.weight"][:, v.patch_dim]
            yield vb + "pre_layrnorm.weight", src["visual.pre_ln.weight"]        # (sic: HF's CLIP spells it this way)
            yield vb + "pre_layrnorm.bias", src["visual.pre_ln.bias"]
        else:
            yield vb + "embeddings.patch_embedding.bias", src["visual.patch_embed.bias"]
        yield vb + "embeddings.position_embedding.weight", src["visual.pos_embed.weight"]
        for i in range(v.depth):
            b, hb = f"visual.blocks.{i}.", f"{vb}encoder.layers.{i}."
            w, bi = src[b + "qkv.weight"], src[b + "qkv.bias"]
            for j, nm in enumerate(("q_proj", "k_proj", "v_proj")):
                yield hb + f"self_attn.{nm}.weight", w[j * E:(j + 1) * E]
                yield hb + f"self_attn.{nm}.bias", bi[j * E:(j + 1) * E]
            yield hb + "self_attn.out_proj.weight", src[b + "proj.weight"]
            yield hb + "self_attn.out_proj.bias", src[b + "proj.bias"]
            yield hb + "layer_norm1.weight", src[b + "norm1.weight"]
            yield hb + "layer_norm1.bias", src[b + "norm1.bias"]
            yield hb + "layer_norm2.weight", src[b + "norm2.weight"]
            yield hb + "layer_norm2.bias", src[b + "norm2.bias"]
            yield hb + "mlp.fc1.weight", src[b + "fc1.weight"]
            yield hb + "mlp.fc1.bias", src[b + "fc1.bias"]
            yield hb + "mlp.fc2.weight", src[b + "fc2.weight"]
            yield hb + "mlp.fc2.bias", src[b + "fc2.bias"]
        yield vb + "post_layernorm.weight", src["visual.post_layernorm.weight"]
        yield vb + "post_layernorm.bias", src["visual.post_layernorm.bias"]
        yield "multi_modal_projector.linear_1.weight", src["visual.merger.fc1.weight"]
        yield "multi_modal_projector.linear_1.bias", src["visual.merger.fc1.bias"]
        yield "multi_modal_projector.linear_2.weight", src["visual.merger.fc2.weight"]
        yield "multi_modal_projector.linear_2.bias", src["visual.merger.fc2.bias"]
        if "image_newline" in src:
            yield "image_newline", src["image_newline"]

    def _hf_named_text(self, src, prefix, head_name):
        t = self.cfg.text
        yield prefix + "embed_tokens.weight", src["embed_tokens.weight"]
        nq, nkv, hd = t.num_heads * t.head_dim, t.num_kv_heads * t.head_dim, t.head_dim
        Ti = t.intermediate_size
        for i in range(t.num_layers):
            b, hb = f"layers.{i}.", f"{prefix}layers.{i}."
            w = src[b + "qkv.weight"]
            yield hb + "self_attn.q_proj.weight", w[:nq]
            yield hb + "self_attn.k_proj.weight", w[nq:nq + nkv]
            yield hb + "self_attn.v_proj.weight", w[nq + nkv:]
            if t.qkv_bias:
                bi = src[b + "qkv.bias"]
                yield hb + "self_attn.q_proj.bias", bi[:nq]
                yield hb + "self_attn.k_proj.bias", bi[nq:nq + nkv]
                yield hb + "self_attn.v_proj.bias", bi[nq + nkv:]
            yield hb + "self_attn.o_proj.weight", src[b + "o.weight"]
            yield hb + "mlp.gate_proj.weight", src[b + "gate_up.weight"][:Ti]
            yield hb + "mlp.up_proj.weight", src[b + "gate_up.weight"][Ti:]
            yield hb + "mlp.down_proj.weight", src[b + "down.weight"]
            yield hb + "input_layernorm.weight", src[b + "ln1.weight"]
            yield hb + "post_attention_layernorm.weight", src[b + "ln2.weight"]
        yield prefix + "norm.weight", src["norm.weight"]
        if not t.tie_word_embeddings:
            yield head_name, src["lm_head.weight"]

    def canonical_name(self, name: str) -> str:
        """Map transformers-5.x key names (`model.visual.*`, `model.language_model.*`) onto the 4.51 layout."""
        if self.cfg.family in ("llava_onevision", "llava", "llava_next"):
            # 5.x: model.vision_tower.*, model.multi_modal_projector.*, model.image_newline, model.language_model.*, lm_head.*
            if name.startswith("model.language_model."):
                return "language_model.model." + name[len("model.language_model."):]
            if name == "lm_head.weight":
                return "language_model.lm_head.weight"
            if name.startswith("model.") and not name.startswith("model.layers") and not name.startswith("model.embed") \
                    and not name.startswith("model.norm"):
                return name[len("model."):]
            return name
        if name.startswith("model.visual."):
            return name[len("model."):]
        if name.startswith("model.language_model."):
            return "model." + name[len("model.language_model."):]
        return name

    def load_hf_state_dict(self, sd: dict, strict: bool = True):
        """Copy an HF state dict (either key layout, any float dtype, any device) into the flat bf16 buffer."""
        sd = {self.canonical_name(k): v for k, v in sd.items()}
        missing = []
        with torch.no_grad():
            for name, view in self.hf_named_tensors("p"):
                if name not in sd:
                    if name.endswith("lm_head.weight") or not strict:
                        continue
                    missing.append(name)
                    continue
                src = sd[name]
                if name in ("visual.patch_embed.proj.weight", "vision_tower.vision_model.embeddings.patch_embedding.weight"):
                    src = src.reshape(src.shape[0], -1)
                if tuple(src.shape) != tuple(view.shape):
                    raise ValueError(f"{name}: checkpoint shape {tuple(src.shape)} != model shape {tuple(view.shape)}")
                view.copy_(src.to(device=self.device, dtype=torch.bfloat16))
        if missing:
            raise KeyError(f"checkpoint is missing {len(missing)} tensors, e.g. {missing[:4]}")
        self.sync_master()

    def hf_state_dict(self) -> dict:
        sd = {}
        for name, view in self.hf_named_tensors("p"):
            tns = view.detach().clone().contiguous()
            if name == "visual.patch_embed.proj.weight":
                v = self.cfg.vision
                tns = tns.view(v.hidden_size, v.in_channels, v.temporal_patch_size, v.patch_size, v.patch_size)
            elif name == "vision_tower.vision_model.embeddings.patch_embedding.weight":
                v = self.cfg.vision
                tns = tns.view(v.hidden_size, v.in_channels, v.patch_size, v.patch_size)
            sd[name] = tns
        return sd

    def init_random(self, seed: int = 0, std: float = 0.02):
        """HF `_init_weights` convention: N(0, std) matrices/embeddings, unit norms, zero biases."""
        gen = torch.Generator(device=self.device).manual_seed(seed)
        with torch.no_grad():
            self.flat.zero_()
            for name, view in self.hf_named_tensors("p"):
                if name.endswith("bias"):
                    continue
                if ("norm" in name and "weight" in name) or "ln_q" in name:
                    view.fill_(1.0)
                else:
                    view.copy_(torch.empty(view.shape, dtype=torch.float32, device=self.device)
                               .normal_(0.0, std, generator=gen).to(torch.bfloat16))
        self.sync_master()

    def sync_master(self):
        if self.master is not None:
            self.master.copy_(self.flat.float())

    def copy_from(self, other: "ParamStore"):
        self.flat.copy_(other.flat)
        self.sync_master()

    def zero_grad(self):
        if self.grad_flat is not None:
            self.grad_flat.zero_()
