"""The VLM forward / backward, written as explicit kernel sequences over the C ABI (ops.py) - no torch.nn modules,
no autograd graph. By default every activation the backward needs is kept in HBM (180 GB makes
`--gradient_checkpointing` unnecessary for the 0.5-3 B models); `VLM.recompute` switches the decoder to per-layer
recompute (only layer inputs stay resident) for models whose state fills the device (Qwen2.5-VL-7B; DESIGN.md "HBM layout").

This is the B200-native stand-in for `model(**inputs).logits` and autograd's backward over it
(ref: train/stage_rl/trainer/sc_grpo_trainer.py:505; HF modeling_qwen2_5_vl.py:455-518 vision tower, :778-827 decoder
layer, :1270-1345 model forward). What it computes per call is the per-token log-prob of given labels (the only thing
`_get_per_token_logps` keeps, sc_grpo_trainer.py:505-514), never the [B, T, V] logits.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .config import VLMConfig
from .geometry import (VisionGeometry, embed_source_index, image_token_count, position_ids as family_position_ids,
                       siglip_geometry, text_rope_tables, vision_geometry)
from .params import ParamStore

bf16, f32 = torch.bfloat16, torch.float32


class VisionCtx:
    pass


class DecoderCtx:
    pass


class VLM:
    """One set of weights (policy or frozen reference) + the kernel schedules that run over them."""

    def __init__(self, cfg: VLMConfig, params: ParamStore):
        self.cfg = cfg
        self.params = params
        self.p = params.p
        self.device = params.device
        self.recompute = False      # per-layer activation recompute in the decoder backward (set by the trainers)
        self.on_layer_grad_ready = None   # callback(layer) once a decoder layer's weight gradients are complete (all-reduce overlap)
        self._native = None               # model-level C ABI handle (native.NativeModel), created on first use
        self._zero_row = torch.zeros(1, max(cfg.text.hidden_size, cfg.vision.hidden_size), dtype=bf16, device=self.device)
        self._causal_cache = {}

    @property
    def g(self):
        return self.params.g

    @property
    def native(self):
        """Handle of the model-level C ABI (decoder / head / rollout layer loops in C++, csrc/model.cu)."""
        if self._native is None:
            from .native import NativeModel
            self._native = NativeModel(self.cfg, self.params)
        return self._native

    def _causal_ranges(self, T):
        r = self._causal_cache.get(T)
        if r is None:
            lo = torch.zeros(T, dtype=torch.int32, device=self.device)
            hi = torch.arange(1, T + 1, dtype=torch.int32, device=self.device)
            r = self._causal_cache[T] = (lo, hi)
        return r

    # =============================================================================================================
    # vision tower
    # =============================================================================================================
    def vision_forward(self, pixel_values: torch.Tensor, grid_thw, save: bool = True):
        """pixel_values [Np, C*tp*ps*ps] (any float dtype) -> image embeddings [Np/merge^2, H_text] bf16."""
        v, p = self.cfg.vision, self.p
        if v.kind in ("siglip", "clip"):
            return self._siglip_forward(pixel_values, grid_thw, save)
        geo = vision_geometry(v, grid_thw, self.device)
        E, nh, hd, Ip = v.hidden_size, v.num_heads, v.head_dim, v.intermediate_padded
        unit = v.spatial_merge_size ** 2
        Np = geo.n_patches
        q25_ = v.kind == "qwen2_5_vl"
        if pixel_values.shape[0] != Np:
            raise ValueError(f"pixel_values has {pixel_values.shape[0]} patches, image_grid_thw implies {Np}")
        px = pixel_values.to(device=self.device, dtype=bf16).contiguous()
        if self._native_vision(hd):
            # ONE call into the model-level C ABI (iadr1_vision_fwd): patch embedding, every block, merger
            fa_full = ops.range_attention(geo, "full", geo.full_lo, geo.full_hi, nh, nh, hd)
            fa_win = ops.range_attention(geo, "win", geo.win_lo, geo.win_hi, nh, nh, hd) if q25_ else None
            return self._vision_native(px, geo, fa_full, fa_win, geo.n_tokens, save)
        x = ops.linear_fwd(px, p["visual.patch_embed.weight"])
        if geo.window_index is not None:
            x = ops.gather_rows(x.view(Np // unit, unit * E), geo.window_index).view(Np, E)
        sh_masked = ops.AttnShape(1, Np, nh, nh, hd, causal=False)
        ctx = VisionCtx()
        ctx.geo, ctx.px, ctx.blocks = geo, px, []
        q25 = v.kind == "qwen2_5_vl"
        for i in range(v.depth):
            b = f"visual.blocks.{i}."
            full = (not q25) or (i in v.fullatt_block_indexes)
            seg = geo.full_seg if full else geo.win_seg
            per_image = False
            fused = None
            if ops.F.supported(hd):
                # fused tcgen05 attention over per-patch key ranges (windows or whole images; any number of images)
                lo, hi = (geo.full_lo, geo.full_hi) if full else (geo.win_lo, geo.win_hi)
                fused = sh = ops.range_attention(geo, "full" if full else "win", lo, hi, nh, nh, hd)
            elif seg:    # equal-length segments: a batch of [seg x seg] attention problems (no masked-out work)
                sh = ops.AttnShape(Np // seg, seg, nh, nh, hd, causal=False)
                lo, hi = geo.seg_ranges[seg]
            elif len(geo.image_ranges) > 1:   # ragged segments, several images: masked attention image by image
                per_image, sh, lo, hi = True, None, None, None
            else:        # ragged segments: one [Np x Np] product with per-row key ranges
                sh = sh_masked
                lo, hi = (geo.full_lo, geo.full_hi) if full else (geo.win_lo, geo.win_hi)
            if q25:
                xn, st1 = ops.rmsnorm_fwd(x, p[b + "norm1.weight"], 1e-6)
            else:
                xn, m1, r1 = ops.layernorm_fwd(x, p[b + "norm1.weight"], p[b + "norm1.bias"], 1e-6)
                st1 = (m1, r1)
            qkv = ops.linear_fwd(xn, p[b + "qkv.weight"], bias=p[b + "qkv.bias"])
            ops.rope_(qkv, geo.cos, geo.sin, 2 * nh, hd, bf16_ops=0)
            if fused is not None:
                attn, P = fused.forward(qkv)
            elif per_image:
                attn = torch.empty(Np, nh * hd, dtype=bf16, device=self.device)
                P, sh = [], ("full" if full else "win")
                for j, (a_, b_) in enumerate(geo.image_ranges):
                    lo_j, hi_j = geo.relative_ranges(sh, j)
                    _, P_j = ops.attention_fwd(qkv[a_:b_], ops.AttnShape(1, b_ - a_, nh, nh, hd, causal=False), lo_j, hi_j,
                                               out=attn[a_:b_])
                    P.append(P_j if save else None)
            else:
                attn, P = ops.attention_fwd(qkv, sh, lo, hi)
            x_mid = ops.linear_fwd(attn, p[b + "proj.weight"], bias=p[b + "proj.bias"], residual=x)
            if q25:
                xn2, st2 = ops.rmsnorm_fwd(x_mid, p[b + "norm2.weight"], 1e-6)
                gu = ops.linear_fwd(xn2, p[b + "gate_up.weight"], bias=p[b + "gate_up.bias"])
                act = ops.act_mul_fwd(gu, Ip, ops.ACT_SILU, gated=True)
                x_out = ops.linear_fwd(act, p[b + "down.weight"], bias=p[b + "down.bias"], residual=x_mid)
            else:
                xn2, m2, r2 = ops.layernorm_fwd(x_mid, p[b + "norm2.weight"], p[b + "norm2.bias"], 1e-6)
                st2 = (m2, r2)
                gu = ops.linear_fwd(xn2, p[b + "fc1.weight"], bias=p[b + "fc1.bias"])
                act = ops.act_mul_fwd(gu, Ip, ops.ACT_QUICK_GELU, gated=False)
                x_out = ops.linear_fwd(act, p[b + "fc2.weight"], bias=p[b + "fc2.bias"], residual=x_mid)
            if save:
                ctx.blocks.append((x, st1, xn, qkv, P, attn, x_mid, st2, xn2, gu, act, lo, hi, sh))
            x = x_out
        if q25:
            xq, stq = ops.rmsnorm_fwd(x, p["visual.merger.ln_q.weight"], 1e-6)
        else:
            xq, mq, rq = ops.layernorm_fwd(x, p["visual.merger.ln_q.weight"], p["visual.merger.ln_q.bias"], 1e-6)
            stq = (mq, rq)
        xm = xq.view(Np // unit, unit * E)
        m1 = ops.linear_fwd(xm, p["visual.merger.fc1.weight"], bias=p["visual.merger.fc1.bias"])
        a1 = ops.act_mul_fwd(m1, unit * E, ops.ACT_GELU, gated=False)
        out = ops.linear_fwd(a1, p["visual.merger.fc2.weight"], bias=p["visual.merger.fc2.bias"])
        if geo.reverse_index is not None:
            out = ops.gather_rows(out, geo.reverse_index)
        if save:
            ctx.x_last, ctx.stq, ctx.xq, ctx.m1, ctx.a1 = x, stq, xq, m1, a1
        return out, (ctx if save else None)

    def _native_vision(self, hd: int) -> bool:
        import os
        return ops.F.supported(hd) and os.environ.get("IADR1_VISION", "native") != "python"

    def _vision_native(self, px, geo, fa_full, fa_win, n_out, save):
        cg = self.native.vision_geom(geo, fa_full.plan, fa_win.plan if fa_win is not None else None, n_out)
        out, ws = self.native.vision_fwd(px, cg, fa_full.plan.npad, save)
        if not save:
            return out, None
        ctx = VisionCtx()
        ctx.native, ctx.geo, ctx.px, ctx.cgeom, ctx.ws = True, geo, px, cg, ws
        return out, ctx

    def vision_backward(self, d_out: torch.Tensor, ctx: VisionCtx):
        """Accumulates vision-tower gradients into the fp32 grad buffer given d(image embeddings) bf16."""
        v, p, g, geo = self.cfg.vision, self.p, self.g, ctx.geo
        if getattr(geo, "has_shrink", False):          # anyres_max: back through the feature-map resize first
            d_out = self._unshrink_grad(d_out, geo)
        if getattr(ctx, "native", False):
            self.native.vision_bwd(d_out.contiguous(), ctx.px, ctx.cgeom, ctx.ws)
            ctx.ws = None
            return
        if v.kind in ("siglip", "clip"):
            return self._siglip_backward(d_out, ctx)
        E, nh, hd, Ip = v.hidden_size, v.num_heads, v.head_dim, v.intermediate_padded
        unit = v.spatial_merge_size ** 2
        Np = geo.n_patches
        q25 = v.kind == "qwen2_5_vl"
        if geo.window_index is not None:
            d_out = ops.gather_rows(d_out, geo.window_index)  # inverse of the final un-permute
        da1 = ops.linear_bwd(d_out, ctx.a1, p["visual.merger.fc2.weight"], g["visual.merger.fc2.weight"],
                             g["visual.merger.fc2.bias"])
        dm1 = ops.act_mul_bwd(da1, ctx.m1, unit * E, ops.ACT_GELU, gated=False)
        dxm = ops.linear_bwd(dm1, ctx.xq.view(Np // unit, unit * E), p["visual.merger.fc1.weight"],
                             g["visual.merger.fc1.weight"], g["visual.merger.fc1.bias"])
        dx = torch.empty(Np, E, dtype=bf16, device=self.device)
        if q25:
            ops.rmsnorm_bwd(dxm.view(Np, E), ctx.x_last, p["visual.merger.ln_q.weight"], ctx.stq, dx,
                            g["visual.merger.ln_q.weight"], add_dx=False)
        else:
            ops.layernorm_bwd(dxm.view(Np, E), ctx.x_last, p["visual.merger.ln_q.weight"], ctx.stq[0], ctx.stq[1], dx,
                              g["visual.merger.ln_q.weight"], g["visual.merger.ln_q.bias"], add_dx=False)
        for i in reversed(range(v.depth)):
            b = f"visual.blocks.{i}."
            x, st1, xn, qkv, P, attn, x_mid, st2, xn2, gu, act, lo, hi, sh = ctx.blocks[i]
            if q25:
                dact = ops.linear_bwd(dx, act, p[b + "down.weight"], g[b + "down.weight"], g[b + "down.bias"])
                dgu = ops.act_mul_bwd(dact, gu, Ip, ops.ACT_SILU, gated=True, dgu=gu)
                dxn2 = ops.linear_bwd(dgu, xn2, p[b + "gate_up.weight"], g[b + "gate_up.weight"], g[b + "gate_up.bias"])
                ops.rmsnorm_bwd(dxn2, x_mid, p[b + "norm2.weight"], st2, dx, g[b + "norm2.weight"], add_dx=True)
            else:
                dact = ops.linear_bwd(dx, act, p[b + "fc2.weight"], g[b + "fc2.weight"], g[b + "fc2.bias"])
                dgu = ops.act_mul_bwd(dact, gu, Ip, ops.ACT_QUICK_GELU, gated=False, dgu=gu)
                dxn2 = ops.linear_bwd(dgu, xn2, p[b + "fc1.weight"], g[b + "fc1.weight"], g[b + "fc1.bias"])
                ops.layernorm_bwd(dxn2, x_mid, p[b + "norm2.weight"], st2[0], st2[1], dx, g[b + "norm2.weight"],
                                  g[b + "norm2.bias"], add_dx=True)
            dattn = ops.linear_bwd(dx, attn, p[b + "proj.weight"], g[b + "proj.weight"], g[b + "proj.bias"])
            if isinstance(sh, ops.F.FusedAttention):
                dqkv = sh.backward(dattn, qkv, P)
            elif isinstance(P, list):     # per-image masked attention (ragged segments, several images)
                dqkv = torch.empty_like(qkv)
                for j, (a_, b_) in enumerate(geo.image_ranges):
                    lo_j, hi_j = geo.relative_ranges(sh, j)
                    ops.attention_bwd(dattn[a_:b_], qkv[a_:b_], P[j], ops.AttnShape(1, b_ - a_, nh, nh, hd, causal=False),
                                      lo_j, hi_j, dqkv=dqkv[a_:b_])
            else:
                dqkv = ops.attention_bwd(dattn, qkv, P, sh, lo, hi)
            ops.rope_(dqkv, geo.cos, geo.sin, 2 * nh, hd, bf16_ops=0, backward=True)
            dxn = ops.linear_bwd(dqkv, xn, p[b + "qkv.weight"], g[b + "qkv.weight"], g[b + "qkv.bias"])
            if q25:
                ops.rmsnorm_bwd(dxn, x, p[b + "norm1.weight"], st1, dx, g[b + "norm1.weight"], add_dx=True)
            else:
                ops.layernorm_bwd(dxn, x, p[b + "norm1.weight"], st1[0], st1[1], dx, g[b + "norm1.weight"],
                                  g[b + "norm1.bias"], add_dx=True)
            ctx.blocks[i] = None
        if geo.reverse_index is not None:
            dx = ops.gather_rows(dx.view(Np // unit, unit * E), geo.reverse_index).view(Np, E)
        ops.linear_bwd(dx, ctx.px, p["visual.patch_embed.weight"], g["visual.patch_embed.weight"], need_dx=False)

    # ---- LLaVA-OneVision: SigLIP tower -> 2-layer GELU projector -> anyres packing -----------------------------------
    def _siglip_forward(self, pixel_values: torch.Tensor, grid, save: bool):
        """Tower + projector + packing, then the anyres_max feature-map shrink for the (rare) images that need it."""
        out, ctx = self._siglip_forward_core(pixel_values, grid, save)
        geo = siglip_geometry(self.cfg, grid, self.device)
        if geo.has_shrink:
            out = self._apply_shrink(out, geo)
        return out, ctx

    def _apply_shrink(self, out_u: torch.Tensor, geo) -> torch.Tensor:
        """LLaVA-OneVision `anyres_max_N` (HF modeling_llava_onevision.py:328-347): the kernels packed every image as
        [base crop | kept feature rows, each followed by image_newline]; an image whose kept map exceeds N crops' worth of tokens
        has that map bilinearly resized (torch's `interpolate`, exactly the reference's call - a rare branch outside the kernels)
        and re-packed with one newline per resized row."""
        Hd = out_u.shape[1]
        nl = self.p["image_newline"].view(1, 1, Hd)
        parts = []
        for u_off, n_u, _, _, tpc, sh in geo.shrinks:
            if sh is None:
                parts.append(out_u[u_off:u_off + n_u])
                continue
            kh, kw, kh2, kw2 = sh
            parts.append(out_u[u_off:u_off + tpc])
            body = out_u[u_off + tpc:u_off + n_u].view(kh, kw + 1, Hd)[:, :kw]                # drop the newline column
            small = torch.nn.functional.interpolate(body.permute(2, 0, 1)[None], [kh2, kw2], mode="bilinear")[0]
            parts.append(torch.cat([small.permute(1, 2, 0), nl.expand(kh2, 1, Hd).to(small.dtype)], 1).reshape(kh2 * (kw2 + 1), Hd))
        return torch.cat(parts, 0).contiguous()

    def _unshrink_grad(self, d_out: torch.Tensor, geo) -> torch.Tensor:
        """Adjoint of `_apply_shrink`: gradient of the final packed tokens -> gradient of the kernels' (unshrunk) packed tokens;
        the newline columns of resized images feed `image_newline`'s gradient directly."""
        Hd = d_out.shape[1]
        parts = []
        for _, n_u, f_off, n_f, tpc, sh in geo.shrinks:
            if sh is None:
                parts.append(d_out[f_off:f_off + n_f])
                continue
            kh, kw, kh2, kw2 = sh
            parts.append(d_out[f_off:f_off + tpc])
            dbody = d_out[f_off + tpc:f_off + n_f].view(kh2, kw2 + 1, Hd)
            if self.g is not None and "image_newline" in self.g:
                self.g["image_newline"].add_(dbody[:, kw2].float().sum(0))
            x = torch.zeros(1, Hd, kh, kw, dtype=torch.float32, device=d_out.device, requires_grad=True)
            y = torch.nn.functional.interpolate(x, [kh2, kw2], mode="bilinear")
            (gx,) = torch.autograd.grad(y, x, dbody[:, :kw2].permute(2, 0, 1)[None].float())          # the resize is linear
            du = torch.zeros(kh, kw + 1, Hd, dtype=d_out.dtype, device=d_out.device)
            du[:, :kw] = gx[0].permute(1, 2, 0).to(d_out.dtype)
            parts.append(du.view(kh * (kw + 1), Hd))
        return torch.cat(parts, 0).contiguous()

    def _siglip_forward_core(self, pixel_values: torch.Tensor, grid, save: bool):
        """pixel_values [n_crops * 729, 3*14*14] (crop 0 of every image = the base crop) -> packed image embeddings
        [n_image_tokens, H_text] bf16. HF: SiglipVisionEmbeddings (modeling_siglip.py:116-186: Conv2d patch embed with bias +
        learned position table), SiglipEncoderLayer x depth (pre-LN attention + tanh-GELU MLP), hidden_states[-1] WITHOUT
        post_layernorm (modeling_llava_onevision.py:405-417), LlavaOnevisionMultiModalProjector (:137-156), pack_image_features
        (:292-355: crops re-tiled to one feature map, unpadded, `image_newline` appended to every row)."""
        v, p = self.cfg.vision, self.p
        geo = siglip_geometry(self.cfg, grid, self.device)
        E, nh, hd, Ip, tpc = v.hidden_size, v.num_heads, v.head_dim, v.intermediate_padded, v.tokens_per_crop
        eps = float(self.cfg.extra.get("vision_layer_norm_eps", 1e-6))
        Np, nc = geo.n_patches, geo.n_crops
        if pixel_values.shape[0] != Np:
            raise ValueError(f"pixel_values has {pixel_values.shape[0]} patches, the image sizes imply {Np}")
        px = pixel_values.to(device=self.device, dtype=bf16)
        if px.shape[1] != v.patch_dim_padded:                      # K of the patch GEMM: 588 -> 592 (zero columns)
            px = torch.nn.functional.pad(px, (0, v.patch_dim_padded - px.shape[1]))
        px = px.contiguous()
        if self._native_vision(hd):
            fa = ops.range_attention(geo, "crops", geo.full_lo, geo.full_hi, nh, nh, hd)
            return self._vision_native(px, geo, fa, None, geo.n_tokens, save)
        if v.kind == "clip":
            raise NotImplementedError("the CLIP tower (LLaVA-1.5) runs through iadr1_vision_fwd only (head_dim must be a multiple of 8)")
        pos = p["visual.pos_embed.weight"].repeat(nc, 1)          # position table tiled over the crops (residual operand)
        x = ops.linear_fwd(px, p["visual.patch_embed.weight"], bias=p["visual.patch_embed.bias"], residual=pos)
        sh = ops.AttnShape(nc, tpc, nh, nh, hd, causal=False)     # attention within one crop
        lo = torch.zeros(tpc, dtype=torch.int32, device=self.device)
        hi = torch.full((tpc,), tpc, dtype=torch.int32, device=self.device)
        fused = ops.range_attention(geo, "crops", geo.full_lo, geo.full_hi, nh, nh, hd) if ops.F.supported(hd) else None
        ctx = VisionCtx()
        ctx.geo, ctx.px, ctx.blocks, ctx.sh, ctx.lo, ctx.hi, ctx.fused = geo, px, [], sh, lo, hi, fused
        for i in range(v.depth):
            b = f"visual.blocks.{i}."
            xn, m1, r1 = ops.layernorm_fwd(x, p[b + "norm1.weight"], p[b + "norm1.bias"], eps)
            qkv = ops.linear_fwd(xn, p[b + "qkv.weight"], bias=p[b + "qkv.bias"])
            attn, P = fused.forward(qkv) if fused is not None else ops.attention_fwd(qkv, sh, lo, hi)
            x_mid = ops.linear_fwd(attn, p[b + "proj.weight"], bias=p[b + "proj.bias"], residual=x)
            xn2, m2, r2 = ops.layernorm_fwd(x_mid, p[b + "norm2.weight"], p[b + "norm2.bias"], eps)
            h1 = ops.linear_fwd(xn2, p[b + "fc1.weight"], bias=p[b + "fc1.bias"])
            act = ops.act_mul_fwd(h1, Ip, ops.ACT_GELU_TANH, gated=False)
            x_out = ops.linear_fwd(act, p[b + "fc2.weight"], bias=p[b + "fc2.bias"], residual=x_mid)
            if save:
                ctx.blocks.append((x, (m1, r1), xn, qkv, P, attn, x_mid, (m2, r2), xn2, h1, act))
            x = x_out
        H = v.out_hidden_size
        m1_ = ops.linear_fwd(x, p["visual.merger.fc1.weight"], bias=p["visual.merger.fc1.bias"])
        a1 = ops.act_mul_fwd(m1_, H, ops.ACT_GELU, gated=False)
        feat = ops.linear_fwd(a1, p["visual.merger.fc2.weight"], bias=p["visual.merger.fc2.bias"])
        out = ops.gather_rows(feat, geo.pack_index, alt=p["image_newline"].view(1, H))
        if save:
            ctx.x_last, ctx.m1, ctx.a1 = x, m1_, a1
        return out, (ctx if save else None)

    def _siglip_backward(self, d_out: torch.Tensor, ctx: VisionCtx):
        v, p, g, geo = self.cfg.vision, self.p, self.g, ctx.geo
        E, Ip, tpc, H = v.hidden_size, v.intermediate_padded, v.tokens_per_crop, v.out_hidden_size
        Np, nc = geo.n_patches, geo.n_crops
        # un-pack: features dropped by the unpadding get zero gradient, image_newline collects one row per feature-map row
        dfeat32 = torch.zeros(Np, H, dtype=f32, device=self.device)
        ops.scatter_add_rows(d_out, geo.pack_index, dfeat32, g["image_newline"].view(1, H))
        dfeat = ops.cast_f32_bf16(dfeat32)
        da1 = ops.linear_bwd(dfeat, ctx.a1, p["visual.merger.fc2.weight"], g["visual.merger.fc2.weight"],
                             g["visual.merger.fc2.bias"])
        dm1 = ops.act_mul_bwd(da1, ctx.m1, H, ops.ACT_GELU, gated=False)
        dx = ops.linear_bwd(dm1, ctx.x_last, p["visual.merger.fc1.weight"], g["visual.merger.fc1.weight"],
                            g["visual.merger.fc1.bias"])
        for i in reversed(range(v.depth)):
            b = f"visual.blocks.{i}."
            x, st1, xn, qkv, P, attn, x_mid, st2, xn2, h1, act = ctx.blocks[i]
            dact = ops.linear_bwd(dx, act, p[b + "fc2.weight"], g[b + "fc2.weight"], g[b + "fc2.bias"])
            dh1 = ops.act_mul_bwd(dact, h1, Ip, ops.ACT_GELU_TANH, gated=False, dgu=h1)
            dxn2 = ops.linear_bwd(dh1, xn2, p[b + "fc1.weight"], g[b + "fc1.weight"], g[b + "fc1.bias"])
            ops.layernorm_bwd(dxn2, x_mid, p[b + "norm2.weight"], st2[0], st2[1], dx, g[b + "norm2.weight"],
                              g[b + "norm2.bias"], add_dx=True)
            dattn = ops.linear_bwd(dx, attn, p[b + "proj.weight"], g[b + "proj.weight"], g[b + "proj.bias"])
            dqkv = ctx.fused.backward(dattn, qkv, P) if ctx.fused is not None else \
                ops.attention_bwd(dattn, qkv, P, ctx.sh, ctx.lo, ctx.hi)
            dxn = ops.linear_bwd(dqkv, xn, p[b + "qkv.weight"], g[b + "qkv.weight"], g[b + "qkv.bias"])
            ops.layernorm_bwd(dxn, x, p[b + "norm1.weight"], st1[0], st1[1], dx, g[b + "norm1.weight"],
                              g[b + "norm1.bias"], add_dx=True)
            ctx.blocks[i] = None
        # x0 = patch_embed(px) + bias + pos[tile]: the table's gradient is the sum over crops
        ops.colsum(dx.view(nc, tpc * E), g["visual.pos_embed.weight"].view(tpc * E))
        ops.linear_bwd(dx, ctx.px, p["visual.patch_embed.weight"], g["visual.patch_embed.weight"],
                       g["visual.patch_embed.bias"], need_dx=False)

    # =============================================================================================================
    # decoder
    # =============================================================================================================
    def full_attention(self, B: int, T: int):
        t = self.cfg.text
        lo, hi = self._causal_ranges(T)
        return ops.FullAttention(B, T, t.num_heads, t.num_kv_heads, t.head_dim, lo, hi, causal=True)

    def decoder_forward(self, src_index: torch.Tensor, image_embeds, attn, cos, sin, save: bool = True, kv_sink=None):
        """src_index [N] int32 (token id, or -1-row into image_embeds) -> last hidden states [N, H] (pre final norm).
        `attn` is the attention strategy for the token layout (ops.FullAttention: B x T rows; ops.SharedPrefixAttention:
        one prompt + G completions). kv_sink(layer, qkv) sees each layer's post-rotary fused qkv buffer (rollout prefill)."""
        t, p = self.cfg.text, self.p
        H, I, nq, nkv, hd = t.hidden_size, t.intermediate_size, t.num_heads, t.num_kv_heads, t.head_dim
        fused = getattr(attn, "fused", None)
        if fused is not None:
            # ONE call into the model-level C ABI: embedding gather + every decoder layer (iadr1_decoder_fwd / iadr1_prefill);
            # activations live in one workspace tensor that the context keeps until the backward
            from .native import KvSink
            mode = 0 if not save else (2 if self.recompute else 1)
            sink = None
            if kv_sink is not None:
                kp, vp = kv_sink["kp"], kv_sink["vp"]
                sink = KvSink(kp.data_ptr(), vp.data_ptr(), kv_sink["n"], kv_sink["p_len"], kp.shape[2], kp.stride(0))
            h, ws = self.native.decoder_fwd(src_index, image_embeds, cos, sin, fused.plan, mode, sink, prefill=sink is not None)
            ctx = DecoderCtx()
            ctx.native, ctx.ws, ctx.mode, ctx.plan = True, ws, mode, fused.plan
            ctx.cos, ctx.sin, ctx.src_index, ctx.layers = cos, sin, src_index, None
            return h, (ctx if save else None)
        if isinstance(kv_sink, dict):     # composed-attention path: the sink as per-layer copies
            kd, n_, pl_ = kv_sink, kv_sink["n"], kv_sink["p_len"]

            def kv_sink(layer, qkv):
                kv = qkv.view(n_, pl_, nq + 2 * nkv, hd)
                kd["kp"][layer, :n_, :pl_].copy_(kv[:, :, nq:nq + nkv])
                kd["vp"][layer, :n_, :pl_].copy_(kv[:, :, nq + nkv:])
        h = ops.gather_rows(p["embed_tokens.weight"], src_index, alt=image_embeds)
        ctx = DecoderCtx()
        ctx.native = False
        ctx.layers, ctx.attn, ctx.cos, ctx.sin, ctx.src_index = [], attn, cos, sin, src_index
        for i in range(t.num_layers):
            if save and self.recompute:
                # activation recompute (what `--gradient_checkpointing` asks of the reference, SC_GRPO_*.sh:57): only the
                # layer INPUT stays resident; the backward re-runs this layer's forward before differentiating it
                ctx.layers.append((h,))
                h, _ = self._layer_forward(i, h, attn, cos, sin, False, kv_sink)
            else:
                h, saved = self._layer_forward(i, h, attn, cos, sin, save, kv_sink)
                if save:
                    ctx.layers.append(saved)
        return h, (ctx if save else None)

    def _layer_forward(self, i, h, attn, cos, sin, save, kv_sink=None):
        """One decoder layer (HF Qwen2_5_VLDecoderLayer, modeling_qwen2_5_vl.py:778-827). Returns (h_out, saved tuple)."""
        t, p = self.cfg.text, self.p
        I, nq, nkv, hd = t.intermediate_size, t.num_heads, t.num_kv_heads, t.head_dim
        b = f"layers.{i}."
        xn, r1 = ops.rmsnorm_fwd(h, p[b + "ln1.weight"], t.rms_norm_eps, save_rstd=save)
        qkv = ops.linear_fwd(xn, p[b + "qkv.weight"], bias=p.get(b + "qkv.bias"))      # no bias under LLaMA (LLaVA-1.5)
        ops.rope_(qkv, cos, sin, nq + nkv, hd, bf16_ops=1)
        if kv_sink is not None:
            kv_sink(i, qkv)
        a_out, a_saved = attn.forward(qkv)
        h_mid = ops.linear_fwd(a_out, p[b + "o.weight"], residual=h)
        xn2, r2 = ops.rmsnorm_fwd(h_mid, p[b + "ln2.weight"], t.rms_norm_eps, save_rstd=save)
        gu = ops.linear_fwd(xn2, p[b + "gate_up.weight"])
        act = ops.act_mul_fwd(gu, I, ops.ACT_SILU, gated=True)
        h_out = ops.linear_fwd(act, p[b + "down.weight"], residual=h_mid)
        return h_out, ((h, r1, xn, qkv, a_saved, a_out, h_mid, r2, xn2, gu, act) if save else None)

    def decoder_backward(self, dh: torch.Tensor, ctx: DecoderCtx, n_image_rows: int):
        """dh [B*T, H] bf16 (consumed in place). Returns d(image embeddings) fp32 [n_image_rows, H] or None."""
        t, p, g = self.cfg.text, self.p, self.g
        I, nq, nkv, hd = t.intermediate_size, t.num_heads, t.num_kv_heads, t.head_dim
        if getattr(ctx, "native", False):
            dimg = torch.zeros(n_image_rows, t.hidden_size, dtype=f32, device=self.device) if n_image_rows > 0 else None
            self.native.decoder_bwd(dh, ctx.src_index, ctx.cos, ctx.sin, ctx.plan, ctx.ws, ctx.mode, dimg, self.on_layer_grad_ready)
            ctx.ws = None
            return dimg
        for i in reversed(range(t.num_layers)):
            b = f"layers.{i}."
            if len(ctx.layers[i]) == 1:       # recompute this layer's activations from its saved input
                _, ctx.layers[i] = self._layer_forward(i, ctx.layers[i][0], ctx.attn, ctx.cos, ctx.sin, True)
            h, r1, xn, qkv, P, attn, h_mid, r2, xn2, gu, act = ctx.layers[i]
            dact = ops.linear_bwd(dh, act, p[b + "down.weight"], g[b + "down.weight"])
            dgu = ops.act_mul_bwd(dact, gu, I, ops.ACT_SILU, gated=True, dgu=gu)
            dxn2 = ops.linear_bwd(dgu, xn2, p[b + "gate_up.weight"], g[b + "gate_up.weight"])
            ops.rmsnorm_bwd(dxn2, h_mid, p[b + "ln2.weight"], r2, dh, g[b + "ln2.weight"], add_dx=True)
            dattn = ops.linear_bwd(dh, attn, p[b + "o.weight"], g[b + "o.weight"])
            dqkv = ctx.attn.backward(dattn, qkv, P)
            ops.rope_(dqkv, ctx.cos, ctx.sin, nq + nkv, hd, bf16_ops=0, backward=True)
            dxn = ops.linear_bwd(dqkv, xn, p[b + "qkv.weight"], g[b + "qkv.weight"], g.get(b + "qkv.bias"))
            ops.rmsnorm_bwd(dxn, h, p[b + "ln1.weight"], r1, dh, g[b + "ln1.weight"], add_dx=True)
            ctx.layers[i] = None  # free this layer's activations as the sweep passes
            if self.on_layer_grad_ready is not None:
                self.on_layer_grad_ready(i)
        dimg = None
        if n_image_rows > 0:
            dimg = torch.zeros(n_image_rows, t.hidden_size, dtype=f32, device=self.device)
        ops.scatter_add_rows(dh, ctx.src_index, g["embed_tokens.weight"], dimg)
        return dimg

    # =============================================================================================================
    # per-token log-probs (the whole `_get_per_token_logps`, sc_grpo_trainer.py:502-514)
    # =============================================================================================================
    def prepare_batch(self, input_ids, pixel_values, grid_thw, position_ids=None, attention_mask=None, prompt_len=None):
        """Host-side prep shared by every forward: M-RoPE positions (4.51.3 semantics), rotary tables, embed index."""
        ids_np = input_ids.detach().cpu().numpy() if torch.is_tensor(input_ids) else np.asarray(input_ids)
        B, T = ids_np.shape
        grid = [tuple(int(x) for x in r) for r in (grid_thw.tolist() if torch.is_tensor(grid_thw) else grid_thw)] \
            if grid_thw is not None else []
        n_img_tokens = sum(image_token_count(self.cfg, g_) for g_ in grid)
        pl = T if prompt_len is None else prompt_len
        per_row = int((ids_np[0, :pl] == self.cfg.image_token_id).sum()) if B else 0
        share = B > 1 and per_row == n_img_tokens
        if position_ids is None:
            am = attention_mask.detach().cpu().numpy() if torch.is_tensor(attention_mask) else attention_mask
            g_rows = grid * B if share else grid
            position_ids, _ = family_position_ids(ids_np, g_rows, self.cfg, am, prompt_len)
            position_ids = torch.from_numpy(position_ids)
        cos, sin = text_rope_tables(position_ids, self.cfg.text, self.device)
        src = torch.from_numpy(embed_source_index(ids_np, self.cfg.image_token_id, share, n_img_tokens, prompt_len)).to(self.device)
        return dict(B=B, T=T, grid=grid, cos=cos, sin=sin, src_index=src, n_img_tokens=n_img_tokens,
                    pixel_values=pixel_values)

    def prepare_group(self, prompt_ids, completion_ids, pixel_values, grid_thw):
        """Shared-prefix layout for ONE GRPO group: tokens [prompt (P) | completion row 0 (C) | ... | row G-1 (C)].
        Returns the batch dict plus `sel_index` / flattened labels for the G*C completion log-probs: the logit predicting
        completion token 0 of every row comes from the (single) last prompt position."""
        pr = np.asarray(prompt_ids, dtype=np.int64).reshape(1, -1)
        comp = completion_ids.detach().cpu().numpy() if torch.is_tensor(completion_ids) else np.asarray(completion_ids)
        G, C = comp.shape
        P = pr.shape[1]
        grid = [tuple(int(x) for x in r) for r in (grid_thw.tolist() if torch.is_tensor(grid_thw) else grid_thw)] \
            if grid_thw is not None else []
        n_img_tokens = sum(image_token_count(self.cfg, g_) for g_ in grid)
        pos_p, delta = family_position_ids(pr, grid, self.cfg)
        nxt = int(pos_p.max()) + 1 if P else 0
        pos_c = np.broadcast_to((nxt + np.arange(C))[None, None, :], (3, G, C)).reshape(3, 1, G * C)
        pos = torch.from_numpy(np.concatenate([pos_p, pos_c], axis=2))
        cos, sin = text_rope_tables(pos, self.cfg.text, self.device)
        src_p = embed_source_index(pr, self.cfg.image_token_id, False, n_img_tokens)
        src = torch.from_numpy(np.concatenate([src_p, comp.reshape(-1).astype(np.int32)])).to(self.device)
        rows = np.empty((G, C), dtype=np.int32)
        rows[:, 0] = P - 1
        rows[:, 1:] = P + np.arange(G)[:, None] * C + np.arange(C - 1)[None, :]
        t = self.cfg.text
        attn = ops.SharedPrefixAttention(P, G, C, t.num_heads, t.num_kv_heads, t.head_dim, self.device)
        return dict(shared=True, P=P, G=G, C=C, N=P + G * C, attn=attn, grid=grid, cos=cos, sin=sin, src_index=src,
                    n_img_tokens=n_img_tokens, pixel_values=pixel_values,
                    sel_index=torch.from_numpy(rows.reshape(-1)).to(self.device),
                    labels=torch.from_numpy(comp.reshape(-1).astype(np.int32)).to(self.device))

    def prepare_groups(self, groups: list):
        """Pack several shared-prefix groups (each a dict prompt_ids / completion_ids / pixel_values / grid_thw) into one
        token stream. Returns the merged batch; `group_slices` gives each group's range in the flattened log-probs."""
        parts = [self.prepare_group(g_["prompt_ids"], g_["completion_ids"], g_["pixel_values"], g_["grid_thw"]) for g_ in groups]
        if len(parts) == 1:
            parts[0]["group_slices"] = [(0, parts[0]["G"] * parts[0]["C"])]
            return parts[0]
        tok_off, img_off, lp_off = 0, 0, 0
        src, sel, slices = [], [], []
        for b in parts:
            s_ = b["src_index"].clone()
            neg = s_ < 0
            s_[neg] -= img_off          # -1 - (row + img_off)
            src.append(s_)
            sel.append(b["sel_index"] + tok_off)
            n_lp = b["G"] * b["C"]
            slices.append((lp_off, lp_off + n_lp))
            tok_off += b["N"]
            img_off += b["n_img_tokens"]
            lp_off += n_lp
        pvs = [b["pixel_values"] for b in parts if b["pixel_values"] is not None]
        return dict(shared=True, N=tok_off, attn=ops.MultiGroupAttention([b["attn"] for b in parts]),
                    grid=[g_ for b in parts for g_ in b["grid"]], cos=torch.cat([b["cos"] for b in parts]),
                    sin=torch.cat([b["sin"] for b in parts]), src_index=torch.cat(src), n_img_tokens=img_off,
                    pixel_values=torch.cat([x.to(self.device) for x in pvs]) if pvs else None,
                    sel_index=torch.cat(sel), labels=torch.cat([b["labels"] for b in parts]), group_slices=slices)

    def logprobs_forward(self, batch: dict, sel_index: torch.Tensor, labels: torch.Tensor, temperature: float = 1.0,
                         save: bool = True, image_embeds: torch.Tensor | None = None, dimg_sink=None):
        """log p(labels[j] | prefix) at hidden-state rows sel_index[j] (flattened b*T + t). Returns (logp fp32, ctx).
        `image_embeds` [n_img_tokens, H]: features of this batch's images computed elsewhere (the trainer runs the vision
        tower ONCE per accumulation window); their gradient is then handed to `dimg_sink(fp32 [n_img_tokens, H])` instead
        of being pushed through the tower here."""
        img, vctx = (None, None)
        if batch["n_img_tokens"] > 0:
            if image_embeds is not None:
                if image_embeds.shape[0] != batch["n_img_tokens"]:
                    raise ValueError(f"image_embeds has {image_embeds.shape[0]} rows, the batch needs {batch['n_img_tokens']}")
                img = image_embeds
            else:
                img, vctx = self.vision_forward(batch["pixel_values"], batch["grid"], save=save)
        attn = batch["attn"] if batch.get("shared") else self.full_attention(batch["B"], batch["T"])
        h, dctx = self.decoder_forward(batch["src_index"], img, attn, batch["cos"], batch["sin"], save=save)
        # final norm + fused lm_head -> log-softmax -> gather in one C call (iadr1_logprob_fwd); its workspace keeps what the
        # backward needs (selected rows, their norm, the row log-sum-exp)
        logp, hws = self.native.logprob_fwd(h, sel_index, labels, temperature, for_backward=save)
        ctx = None
        if save:
            ctx = dict(vctx=vctx, dctx=dctx, hws=hws, labels=labels, sel_index=sel_index,
                       temperature=temperature, N=attn.n_tokens, n_img=batch["n_img_tokens"],
                       dimg_sink=dimg_sink if image_embeds is not None else None)
        return logp, ctx

    def logprobs_backward(self, dlogp: torch.Tensor, ctx: dict):
        """Back-propagates d(loss)/d(logp) [M] through lm_head, decoder and vision tower into the fp32 grad buffer."""
        t = self.cfg.text
        # lm_head / final-norm backward + scatter-ADD back to token rows (the last prompt position is selected once per row in
        # the shared layout) in one C call (iadr1_logprob_bwd)
        dh32 = self.native.logprob_bwd(dlogp, ctx["sel_index"], ctx["labels"], ctx["temperature"], ctx["hws"], ctx["N"])
        ctx["hws"] = None
        dh = ops.cast_f32_bf16(dh32)
        dimg32 = self.decoder_backward(dh, ctx["dctx"], ctx["n_img"])
        if dimg32 is not None and ctx.get("dimg_sink") is not None:
            ctx["dimg_sink"](dimg32)
        elif dimg32 is not None and ctx["vctx"] is not None:
            self.vision_backward(ops.cast_f32_bf16(dimg32), ctx["vctx"])
