"""Python handle of the model-level C ABI (include/iadr1_b200.h "B-inner"): weights registered once by pointer, then ONE
ctypes call per decoder forward / backward, log-prob head, prefill and decode step - the layer loops run in C++
(csrc/model.cu). These calls stand in for `model(**inputs).logits` (ref: train/stage_rl/trainer/sc_grpo_trainer.py:505),
autograd's backward and `llm.generate` (:667)."""
from __future__ import annotations

import ctypes as C

import torch

from . import lib as L


class ModelCfg(C.Structure):
    _fields_ = [("vocab", C.c_int), ("hidden", C.c_int), ("inter", C.c_int), ("layers", C.c_int), ("nq", C.c_int),
                ("nkv", C.c_int), ("hd", C.c_int), ("rms_eps", C.c_float),
                ("v_kind", C.c_int), ("v_depth", C.c_int), ("v_hidden", C.c_int), ("v_heads", C.c_int), ("v_inter", C.c_int),
                ("v_out_hidden", C.c_int), ("v_patch_dim", C.c_int), ("v_merge_unit", C.c_int), ("v_tokens_per_crop", C.c_int),
                ("v_fullatt_mask", C.c_uint), ("v_eps", C.c_float)]


class AttnPlan(C.Structure):
    _fields_ = [("ranges", C.c_void_p), ("q_items", C.c_void_p), ("n_q", C.c_int), ("sched_fwd", C.c_void_p),
                ("n_cta_fwd", C.c_int), ("sched_dq", C.c_void_p), ("n_cta_dq", C.c_int), ("k_items", C.c_void_p),
                ("n_k", C.c_int), ("sched_kv", C.c_void_p), ("n_cta_kv", C.c_int), ("npad", C.c_longlong),
                ("n_tokens", C.c_longlong)]


class KvSink(C.Structure):
    _fields_ = [("kp", C.c_void_p), ("vp", C.c_void_p), ("n_groups", C.c_int), ("p_len", C.c_int), ("p_max", C.c_int),
                ("layer_stride", C.c_longlong)]


class DecodeState(C.Structure):
    _fields_ = [("R", C.c_int), ("n_groups", C.c_int), ("p_max", C.c_int), ("c_max", C.c_int), ("nsplit", C.c_int),
                ("max_pos", C.c_int), ("block_n", C.c_int),
                ("kp", C.c_void_p), ("vp", C.c_void_p), ("kc", C.c_void_p), ("vc", C.c_void_p),
                ("state", C.c_void_p), ("tok", C.c_void_p), ("finished", C.c_void_p), ("out_tokens", C.c_void_p),
                ("rope_delta", C.c_void_p), ("row_plen", C.c_void_p), ("row_group", C.c_void_p),
                ("h", C.c_void_p), ("xn", C.c_void_p), ("qkv", C.c_void_p), ("attn", C.c_void_p), ("part", C.c_void_p),
                ("tickets", C.c_void_p), ("act", C.c_void_p), ("logits", C.c_void_p),
                ("cos_tab", C.c_void_p), ("sin_tab", C.c_void_p),
                ("temperature", C.c_float), ("top_k", C.c_int), ("top_p", C.c_float), ("eos_id", C.c_int),
                ("pad_id", C.c_int), ("forbid_eos", C.c_int), ("chain_counters", C.c_void_p), ("attn_psplit", C.c_int)]


class VisionGeom(C.Structure):
    _fields_ = [("n_patches", C.c_longlong), ("n_out", C.c_longlong), ("cos", C.c_void_p), ("sin", C.c_void_p),
                ("window_index", C.c_void_p), ("reverse_index", C.c_void_p), ("plan_full", C.POINTER(AttnPlan)),
                ("plan_win", C.POINTER(AttnPlan)), ("pos_index", C.c_void_p), ("pack_index", C.c_void_p)]


V_KINDS = {"qwen2_5_vl": 0, "qwen2_vl": 1, "siglip": 2, "clip": 3}

LAYER_CB = C.CFUNCTYPE(None, C.c_int, C.c_void_p)


def plan_struct(plan) -> AttnPlan:
    """ctypes view of an fmha.FmhaPlan (cached on the plan; the device tensors stay owned by it)."""
    s = getattr(plan, "_cstruct", None)
    if s is None:
        s = AttnPlan(plan.rng.data_ptr(), plan.q_items.data_ptr(), plan.n_q, plan.sched_fwd.data_ptr(), plan.n_cta_fwd,
                     plan.sched_dq.data_ptr(), plan.n_cta_dq, plan.k_items.data_ptr(), plan.n_k, plan.sched_kv.data_ptr(),
                     plan.n_cta_kv, plan.npad, plan.n_tokens)
        plan._cstruct = s
    return s


class NativeModel:
    def __init__(self, cfg, params):
        t = cfg.text
        self.cfg, self.params = cfg, params
        self.handle = C.c_void_p()
        v = cfg.vision
        mask = 0
        for i in v.fullatt_block_indexes:
            mask |= 1 << i
        mc = ModelCfg(t.vocab_size, t.hidden_size, t.intermediate_size, t.num_layers, t.num_heads, t.num_kv_heads, t.head_dim,
                      t.rms_norm_eps, V_KINDS[v.kind], v.run_depth if v.kind == "clip" else v.depth, v.hidden_size, v.num_heads,
                      v.intermediate_padded, v.out_hidden_size, v.patch_dim_padded, v.spatial_merge_size ** 2,
                      v.tokens_per_crop if v.kind in ("siglip", "clip") else 0, mask,
                      float(cfg.extra.get("vision_layer_norm_eps", 1e-6)))
        L.check(L.lib().iadr1_model_create(C.byref(mc), C.byref(self.handle)), "model_create")
        for name, w in params.p.items():
            g = params.g[name] if params.g is not None else None
            L.check(L.lib().iadr1_bind_weights(self.handle, name.encode(), w.data_ptr(), None if g is None else g.data_ptr()),
                    "bind_weights")

    def __del__(self):
        try:
            if self.handle:
                L.lib().iadr1_model_destroy(self.handle)
        except Exception:
            pass

    # ---- decoder ----------------------------------------------------------------------------------------------------
    def decoder_workspace(self, n_tokens: int, npad: int, mode: int) -> torch.Tensor:
        nbytes = C.c_longlong()
        L.check(L.lib().iadr1_decoder_workspace_bytes(self.handle, n_tokens, npad, mode, C.byref(nbytes)), "decoder_workspace_bytes")
        return torch.empty(nbytes.value, dtype=torch.uint8, device=self.params.device)

    def decoder_fwd(self, src_index, image_embeds, cos, sin, plan, mode: int, sink: KvSink | None = None, prefill: bool = False):
        N, H = int(src_index.shape[0]), self.cfg.text.hidden_size
        ws = self.decoder_workspace(N, plan.npad, mode)
        h_last = C.c_void_p()
        ps = plan_struct(plan)
        img = None if image_embeds is None else image_embeds.data_ptr()
        if prefill:
            L.check(L.lib().iadr1_prefill(self.handle, src_index.data_ptr(), img, N, cos.data_ptr(), sin.data_ptr(), C.byref(ps),
                                          ws.data_ptr(), C.byref(sink), C.byref(h_last), L.stream_ptr()), "prefill")
        else:
            L.check(L.lib().iadr1_decoder_fwd(self.handle, src_index.data_ptr(), img, N, cos.data_ptr(), sin.data_ptr(),
                                              C.byref(ps), ws.data_ptr(), mode, None if sink is None else C.byref(sink),
                                              C.byref(h_last), L.stream_ptr()), "decoder_fwd")
        off = h_last.value - ws.data_ptr()
        h = ws[off:off + N * H * 2].view(torch.bfloat16).view(N, H)
        return h, ws

    def decoder_bwd(self, dh, src_index, cos, sin, plan, ws, mode: int, dimg32, on_layer_done=None):
        ps = plan_struct(plan)
        cb = LAYER_CB(lambda layer, _u: on_layer_done(layer)) if on_layer_done is not None else None
        L.check(L.lib().iadr1_decoder_bwd(self.handle, dh.data_ptr(), src_index.data_ptr(), cos.data_ptr(), sin.data_ptr(),
                                          C.byref(ps), ws.data_ptr(), mode, int(src_index.shape[0]),
                                          None if dimg32 is None else dimg32.data_ptr(),
                                          C.cast(cb, C.c_void_p) if cb is not None else None, None, L.stream_ptr()), "decoder_bwd")

    # ---- log-prob / cross-entropy head ------------------------------------------------------------------------------------
    def logprob_fwd(self, h, sel_index, labels, temperature: float, for_backward: bool):
        M = int(sel_index.shape[0])
        nbytes = C.c_longlong()
        L.check(L.lib().iadr1_logprob_workspace_bytes(self.handle, M, int(for_backward), C.byref(nbytes)), "logprob_workspace_bytes")
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=h.device)
        logp = torch.empty(M, dtype=torch.float32, device=h.device)
        L.check(L.lib().iadr1_logprob_fwd(self.handle, h.data_ptr(), sel_index.data_ptr(), labels.data_ptr(), M, temperature,
                                          ws.data_ptr(), logp.data_ptr(), L.stream_ptr()), "logprob_fwd")
        return logp, ws

    def logprob_bwd(self, dlogp, sel_index, labels, temperature: float, ws, n_tokens: int):
        H = self.cfg.text.hidden_size
        dh32 = torch.zeros(n_tokens, H, dtype=torch.float32, device=dlogp.device)
        d = dlogp.to(torch.float32).contiguous()
        L.check(L.lib().iadr1_logprob_bwd(self.handle, d.data_ptr(), sel_index.data_ptr(), labels.data_ptr(), int(sel_index.shape[0]),
                                          temperature, ws.data_ptr(), dh32.data_ptr(), L.stream_ptr()), "logprob_bwd")
        return dh32

    # ---- vision tower ------------------------------------------------------------------------------------------------------
    def vision_geom(self, geo, plan_full, plan_win, n_out: int) -> VisionGeom:
        """ctypes view of a geometry object (geometry.VisionGeometry / SiglipGeometry) + its attention plans; cached on it."""
        s = geo.__dict__.get("_cgeom")
        if s is None:
            def ptr(name):
                t_ = getattr(geo, name, None)
                return None if t_ is None else t_.data_ptr()
            pf, pw = plan_struct(plan_full), (plan_struct(plan_win) if plan_win is not None else None)
            s = VisionGeom(geo.n_patches, n_out, ptr("cos"), ptr("sin"), ptr("window_index"), ptr("reverse_index"), C.pointer(pf),
                           C.pointer(pw) if pw is not None else None, ptr("pos_index"), ptr("pack_index"))
            geo._cgeom = s
        return s

    def vision_fwd(self, px, cgeom: VisionGeom, npad: int, save: bool):
        """px bf16 [n_patches, patch_dim_padded] -> (image embeddings bf16 [n_out, out_hidden], workspace)."""
        nbytes = C.c_longlong()
        L.check(L.lib().iadr1_vision_workspace_bytes(self.handle, cgeom.n_patches, cgeom.n_out, npad, int(save), C.byref(nbytes)),
                "vision_workspace_bytes")
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=px.device)
        out = torch.empty(cgeom.n_out, self.cfg.vision.out_hidden_size, dtype=torch.bfloat16, device=px.device)
        L.check(L.lib().iadr1_vision_fwd(self.handle, px.data_ptr(), C.byref(cgeom), ws.data_ptr(), int(save), out.data_ptr(),
                                         L.stream_ptr()), "vision_fwd")
        return out, ws

    def vision_bwd(self, d_out, px, cgeom: VisionGeom, ws):
        L.check(L.lib().iadr1_vision_bwd(self.handle, d_out.data_ptr(), px.data_ptr(), C.byref(cgeom), ws.data_ptr(),
                                         L.stream_ptr()), "vision_bwd")
