"""In-rank rollout engine: prefill once per prompt, fork the KV prefix G ways, decode all rows of all in-flight groups
in lock-step through one CUDA-graph-captured step.

Replaces the reference's dedicated-GPU vLLM engine and its per-step weight copy
(ref: train/stage_rl/trainer/sc_grpo_trainer.py:314-358 engine, :569-579 `_move_model_to_vllm`, :643-677 generate):
  * weights are read IN PLACE from the live training buffer - no policy->engine sync (SURVEY.md K20 eliminated);
  * the prompt (ViT + prefill) runs once per group and its K/V are shared by the G rows (what vLLM's
    `enable_prefix_caching=True`, :351, achieves by hashing);
  * several groups decode together so each weight byte streamed from HBM serves G x n_groups rows;
  * sampler contract = SamplingParams(temperature, top_p=0.9, top_k=50, max_tokens=C), :353-358.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import lib as L
from . import ops
from .geometry import decode_rope_table

bf16, f32, i32 = torch.bfloat16, torch.float32, torch.int32
NUM_SMS = 148


def _split_for(m_feat: int, k: int) -> int:
    tiles = (m_feat + 127) // 128
    nkb = (k + 63) // 64
    return max(1, min(nkb, NUM_SMS // max(tiles, 1)))


class RolloutEngine:
    def __init__(self, vlm, max_groups: int, num_generations: int, p_max: int, c_max: int, temperature: float = 0.9,
                 top_k: int = 50, top_p: float = 0.9, forbid_eos: bool = False, use_cuda_graph: bool = True):
        if top_k is None or top_k <= 0 or top_k > 128:
            raise ValueError("rollout sampler supports 1 <= top_k <= 128 (reference uses top_k=50)")
        self.vlm, self.cfg = vlm, vlm.cfg
        t = self.cfg.text
        self.G, self.n_groups = num_generations, max_groups
        self.R = R = max_groups * num_generations
        if R > 256:
            raise ValueError("at most 256 rows decode together (one UMMA N tile)")
        self.p_max, self.c_max = p_max, c_max
        self.temperature, self.top_k, self.top_p, self.forbid_eos = temperature, top_k, top_p, forbid_eos
        self.use_graph = use_cuda_graph
        self.use_pdl = os.environ.get("IADR1_PDL", "1") != "0"
        dev = vlm.device
        Lyr, nkv, hd, H, I, V = t.num_layers, t.num_kv_heads, t.head_dim, t.hidden_size, t.intermediate_size, t.vocab_size
        self.kp = torch.zeros(Lyr, max_groups, p_max, nkv, hd, dtype=bf16, device=dev)
        self.vp = torch.zeros_like(self.kp)
        self.kc = torch.zeros(Lyr, R, c_max, nkv, hd, dtype=bf16, device=dev)
        self.vc = torch.zeros_like(self.kc)
        self.state = torch.zeros(8, dtype=i32, device=dev)
        self.tok = torch.zeros(R, dtype=i32, device=dev)
        self.finished = torch.zeros(R, dtype=i32, device=dev)
        self.out_tokens = torch.zeros(R, c_max, dtype=i32, device=dev)
        self.rope_delta = torch.zeros(R, dtype=i32, device=dev)
        self.row_plen = torch.zeros(R, dtype=i32, device=dev)
        self.row_group = (torch.arange(R, device=dev) // num_generations).to(i32)
        self.h = torch.zeros(R, H, dtype=f32, device=dev)
        self.xn = torch.zeros(R, H, dtype=bf16, device=dev)
        self.qkv = torch.zeros(R, t.qkv_dim, dtype=f32, device=dev)
        self.q = torch.zeros(R, t.num_heads * hd, dtype=bf16, device=dev)
        self.attn = torch.zeros(R, t.num_heads * hd, dtype=bf16, device=dev)
        if hd in (64, 128):
            # tensor-core kernel: ONE wave of (row, kv head, split) CTAs at 3 CTAs per SM; each CTA pipelines its balanced
            # share of the live context (decode.cu: decode_attn_mma_kernel)
            # 4-warp CTAs (64-key chunks, 3 per SM) by default; IADR1_DECODE_ATTN_NW=2 selects 2-warp CTAs (32-key chunks,
            # 5 per SM, finer context splits) for experiments
            nw = int(os.environ.get("IADR1_DECODE_ATTN_NW", "0"))
            s4 = max(1, min(16, 3 * NUM_SMS // (R * nkv)))
            s2 = max(1, min(16, 5 * NUM_SMS // (R * nkv)))
            if nw == 0:
                nw = 4     # measured (profiles/r01_decode_timeline.txt): the 2-warp variant is slower at R = 64 and R = 128
            self.nsplit = s4 if nw == 4 else s2
            if os.environ.get("IADR1_DECODE_ATTN_NSPLIT"):
                self.nsplit = max(1, min(16, int(os.environ["IADR1_DECODE_ATTN_NSPLIT"])))
            self.attn_nw = nw
            # shared-prefix form (decode.cu: decode_attn_grouped_kernel): the prompt keys once per group in P-kind CTAs, each
            # row's own keys in C-kind CTAs, all in one wave of <= 3 CTAs per SM; `part` then holds psplit + csplit slots
            # (measured at 3B widths, P = 297, 128 rows: 31 us per layer against 22 us for the per-row kernel - building the
            # 64-row query tile and merging 4-6 slots costs more than the shared staging saves while the prompt is about half
            # of the context; it is the default only when a LONG prompt dominates (p_max >= 4 c_max and >= 1024 keys), e.g. LLaVA-OneVision's 3.7 k image tokens)
            grouped = os.environ.get("IADR1_DECODE_ATTN_GROUPED", "auto")
            if (grouped == "1" or (grouped == "auto" and p_max >= 4 * c_max and p_max >= 1024)) and nw == 4:
                budget, rblocks = 3 * NUM_SMS, (num_generations + 7) // 8
                # prompt CTAs should take about as long as the row CTAs (mid-rollout: c_max / 2 keys): ~2 chunks of 64 keys each
                ps = max(1, min(8, (p_max + 127) // 128))
                if os.environ.get("IADR1_DECODE_ATTN_PSPLIT"):
                    ps = max(1, min(8, int(os.environ["IADR1_DECODE_ATTN_PSPLIT"])))
                while ps > 1 and max_groups * rblocks * nkv * ps > budget // 2:
                    ps -= 1
                cs_ = max(1, min(8, (budget - max_groups * rblocks * nkv * ps) // (R * nkv)))
                self.attn_psplit, self.nsplit = ps, ps + cs_
        else:
            self.nsplit = max(1, min(32, (p_max + c_max + 127) // 128))   # scalar kernel: one 128-key chunk per CTA
            self.attn_nw = 4
        self.attn_psplit = getattr(self, "attn_psplit", 0)
        self.part = torch.zeros(R, t.num_heads, self.nsplit, hd + 2, dtype=f32, device=dev)
        self.tickets = torch.zeros(R * nkv, dtype=i32, device=dev)
        self.chain_counters = torch.zeros((t.num_layers + 1) * 8, dtype=i32, device=dev)   # persistent decode-layer chain
        # fp32 gate|up accumulator of the stream-K product; decode_silu_mul_f32 leaves it zero for the next layer
        self.block_n = max(16, ops.ceil_to(R, 16))
        self.co_resident = os.environ.get("IADR1_DECODE_CORES", "0") != "0" and self.block_n <= 128
        self.no_bulk_red = os.environ.get("IADR1_DECODE_BULKRED", "1") == "0"
        self.gu_mode = os.environ.get("IADR1_DECODE_GU", "fused")      # fused | streamk | f32 | bf16 (probe A/B)
        self.gu = torch.zeros(R, 2 * I, dtype=f32 if self.gu_mode in ("streamk", "f32") else bf16, device=dev)
        self.act = torch.zeros(R, I, dtype=bf16, device=dev)
        self.logits = torch.zeros(R, V, dtype=f32, device=dev)
        self.max_pos = p_max + c_max + 8
        self.cos_tab, self.sin_tab = decode_rope_table(t, self.max_pos, dev)
        self._graph = None
        self._dstate = None
        # the C++ decode step covers the default configuration; probe modes (IADR1_DECODE_*) keep the per-kernel Python body
        self._native_ok = (self.gu_mode == "fused" and not self.co_resident and not self.no_bulk_red
                           and os.environ.get("IADR1_DECODE_NATIVE", "1") != "0")
        self._seed = 0
        self.replays = 0            # graph replays so far (bench.py counts kernels = replays * kernels_per_step)
        self.kernels_per_step = 0

    # ---------------------------------------------------------------------------------------------------------------
    def _skinny(self, W, x, out, split_k=1, atomic=False, bias=None, stream_k=False):
        """out[R, F] (+)= x[R, K] @ W[F, K]^T with W as the 128-row MMA operand (swap-AB), transposed store.
        stream_k: the weight's k-blocks are spread evenly over all SMs (fp32 atomics into `out`)."""
        if stream_k:
            split_k, atomic = 1, True
        L.gemm(W, x, out=out, trans_out=True, split_k=split_k, atomic=atomic or split_k > 1, bias=bias, bias_per_m=True,
               block_n=self.block_n, a_static=True, stream_k=stream_k, co_resident=self.co_resident,
               no_bulk_red=self.no_bulk_red)

    def _head_and_sample(self, first: int):
        if self._native_ok:
            import ctypes as C
            L.check(L.lib().iadr1_decode_head(self.vlm.native.handle, C.byref(self._native_state()), first, L.stream_ptr()), "decode_head")
            return
        t, p, lib = self.cfg.text, self.vlm.p, L.lib()
        s = L.stream_ptr()
        L.check(lib.iadr1_rmsnorm_f32in(self.h.data_ptr(), p["norm.weight"].data_ptr(), self.xn.data_ptr(), self.R,
                                        t.hidden_size, t.rms_norm_eps, None, 0, s), "rmsnorm_f32in")
        self._skinny(self.vlm.params.lm_head, self.xn, self.logits)
        # seed argument 0: the per-call seed lives in state[4:6] on the device (a captured graph freezes its arguments)
        L.check(lib.iadr1_sample(self.logits.data_ptr(), self.R, t.vocab_size, self.temperature, self.top_k, self.top_p,
                                 0, self.state.data_ptr(), self.tok.data_ptr(), self.finished.data_ptr(),
                                 self.out_tokens.data_ptr(), self.c_max, self.cfg.eos_token_id, self.cfg.pad_token_id,
                                 int(self.forbid_eos), first, s), "sample")

    def _decode_step(self):
        lib = L.lib()
        lib.iadr1_set_pdl(int(self.use_pdl))
        try:
            self._decode_step_body()
        finally:
            lib.iadr1_set_pdl(0)

    def _native_state(self):
        """ctypes view of the engine for the model-level C ABI (iadr1_decode_step / iadr1_decode_head)."""
        if self._dstate is None:
            from .native import DecodeState
            c = self.cfg
            self._dstate = DecodeState(
                self.R, self.n_groups, self.p_max, self.c_max, -self.nsplit if self.attn_nw == 2 else self.nsplit, self.max_pos,
                self.block_n, self.kp.data_ptr(), self.vp.data_ptr(), self.kc.data_ptr(), self.vc.data_ptr(),
                self.state.data_ptr(), self.tok.data_ptr(), self.finished.data_ptr(), self.out_tokens.data_ptr(),
                self.rope_delta.data_ptr(), self.row_plen.data_ptr(), self.row_group.data_ptr(), self.h.data_ptr(),
                self.xn.data_ptr(), self.qkv.data_ptr(), self.attn.data_ptr(), self.part.data_ptr(), self.tickets.data_ptr(),
                self.act.data_ptr(), self.logits.data_ptr(), self.cos_tab.data_ptr(), self.sin_tab.data_ptr(),
                self.temperature, self.top_k, self.top_p, c.eos_token_id, c.pad_token_id, int(self.forbid_eos),
                self.chain_counters.data_ptr(), self.attn_psplit)
        return self._dstate

    def _decode_step_body(self):
        if self._native_ok:
            import ctypes as C
            L.check(L.lib().iadr1_decode_step(self.vlm.native.handle, C.byref(self._native_state()), L.stream_ptr()), "decode_step")
            return
        t, p, lib = self.cfg.text, self.vlm.p, L.lib()
        R, H, I, nq, nkv, hd = self.R, t.hidden_size, t.intermediate_size, t.num_heads, t.num_kv_heads, t.head_dim
        s = L.stream_ptr()
        L.check(lib.iadr1_decode_embed(p["embed_tokens.weight"].data_ptr(), self.tok.data_ptr(), self.h.data_ptr(), R, H, s),
                "decode_embed")
        sk_qkv, sk_o, sk_d = _split_for(t.qkv_dim, H), _split_for(H, nq * hd), _split_for(H, I)
        for i in range(t.num_layers):
            b = f"layers.{i}."
            # RMSNorm also clears the fp32 qkv accumulator the split-K GEMM adds into
            L.check(lib.iadr1_rmsnorm_f32in(self.h.data_ptr(), p[b + "ln1.weight"].data_ptr(), self.xn.data_ptr(), R, H,
                                            t.rms_norm_eps, self.qkv.data_ptr(), t.qkv_dim, s), "rmsnorm_f32in")
            self._skinny(p[b + "qkv.weight"], self.xn, self.qkv, split_k=sk_qkv, atomic=True, bias=p.get(b + "qkv.bias"))
            L.check(lib.iadr1_decode_attention_fused(
                self.qkv.data_ptr(), self.cos_tab.data_ptr(), self.sin_tab.data_ptr(), self.rope_delta.data_ptr(),
                self.kp[i].data_ptr(), self.vp[i].data_ptr(), self.kc[i].data_ptr(), self.vc[i].data_ptr(),
                self.state.data_ptr(), self.row_group.data_ptr(), self.row_plen.data_ptr(), self.finished.data_ptr(), self.part.data_ptr(),
                self.tickets.data_ptr(), self.attn.data_ptr(), R, nq, nkv, hd, self.p_max, self.c_max,
                -self.nsplit if self.attn_nw == 2 else self.nsplit, self.max_pos, float(hd) ** -0.5, s), "decode_attention_fused")
            self._skinny(p[b + "o.weight"], self.attn, self.h, split_k=sk_o, atomic=True)        # h += attn @ Wo^T
            L.check(lib.iadr1_rmsnorm_f32in(self.h.data_ptr(), p[b + "ln2.weight"].data_ptr(), self.xn.data_ptr(), R, H,
                                            t.rms_norm_eps, None, 0, s), "rmsnorm_f32in")
            if self.gu_mode.startswith("fused"):
                L.gemm_swiglu(p[b + "gate_up.weight"], self.xn, self.act, block_n=self.block_n,
                              co_resident=self.gu_mode == "fused")
            elif self.gu_mode == "bf16":
                self._skinny(p[b + "gate_up.weight"], self.xn, self.gu)
                ops.act_mul_fwd(self.gu, I, ops.ACT_SILU, gated=True, out=self.act)
            else:
                self._skinny(p[b + "gate_up.weight"], self.xn, self.gu, atomic=True, stream_k=self.gu_mode == "streamk")
                L.check(lib.iadr1_decode_silu_mul_f32(self.gu.data_ptr(), self.act.data_ptr(), R, I, s), "decode_silu_mul_f32")
            self._skinny(p[b + "down.weight"], self.act, self.h, split_k=sk_d, atomic=True)      # h += mlp
        self._head_and_sample(first=0)
        L.check(lib.iadr1_decode_advance(self.state.data_ptr(), s), "decode_advance")

    # ---------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, prompts: list, seed: int = 0, max_new_tokens: int | None = None, sync_every: int = 32,
                 logits_hook=None, image_embeds: torch.Tensor | None = None):
        """prompts: list (<= max_groups) of dicts {input_ids [P] (list/array), pixel_values, grid_thw}.
        Returns (completion_ids int32 [n_groups_used * G, C] on device, stats dict)."""
        n = len(prompts)
        if n == 0 or n > self.n_groups:
            raise ValueError(f"need 1..{self.n_groups} prompts, got {n}")
        C = self.c_max if max_new_tokens is None else min(max_new_tokens, self.c_max)
        t, vlm, G = self.cfg.text, self.vlm, self.G
        self._seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.state.zero_()
        lo, hi = self._seed & 0xFFFFFFFF, self._seed >> 32
        self.state[4:6] = torch.tensor([lo - (1 << 32) if lo >= (1 << 31) else lo, hi - (1 << 32) if hi >= (1 << 31) else hi],
                                       dtype=i32, device=self.state.device)
        # every id of generation_config.json's eos list stops a row (the sampler takes one id as an argument, a second one
        # through the device state)
        eos_ids = [int(x) for x in self.cfg.extra.get("eos_token_ids", [])] if hasattr(self.cfg, "extra") else []
        others = [x for x in eos_ids if x != self.cfg.eos_token_id]
        if others:
            self.state[1] = others[0] + 1
        self.finished.zero_()
        self.gu.zero_()
        self.out_tokens.fill_(self.cfg.pad_token_id)
        plen = np.zeros(self.R, dtype=np.int32)
        delta = np.zeros(self.R, dtype=np.int32)
        # ---- prefill: ALL prompts in one right-padded [n, Pmax] pass through the training forward kernels (one vision
        # tower call over every image, decoder GEMMs at M = n * Pmax); each layer's post-rotary K/V land in the shared
        # prefix cache. Padding sits after the real tokens, so causal attention never lets it influence them.
        from .geometry import position_ids as mrope_position_ids   # family dispatch (M-RoPE or plain 1-D positions)
        id_list = [np.asarray(pr["input_ids"], dtype=np.int64).reshape(-1) for pr in prompts]
        lens = [len(x) for x in id_list]
        Pmax = max(lens)
        if Pmax > self.p_max:
            raise ValueError(f"prompt of {Pmax} tokens exceeds p_max={self.p_max}")
        ids = np.full((n, Pmax), self.cfg.pad_token_id, dtype=np.int64)
        am = np.zeros((n, Pmax), dtype=np.int64)
        for i, x in enumerate(id_list):
            ids[i, :lens[i]], am[i, :lens[i]] = x, 1
        pvs = [pr.get("pixel_values") for pr in prompts if pr.get("pixel_values") is not None]
        grids = [g_ for pr in prompts if pr.get("grid_thw") is not None
                 for g_ in (pr["grid_thw"].tolist() if torch.is_tensor(pr["grid_thw"]) else pr["grid_thw"])]
        pv = torch.cat([x.to(vlm.device) for x in pvs], 0) if pvs else None
        pos, _ = mrope_position_ids(ids, grids, self.cfg, am)
        batch = vlm.prepare_batch(ids, pv, grids if grids else None, position_ids=torch.from_numpy(pos), attention_mask=am)
        img = None
        if batch["n_img_tokens"] > 0:
            # `image_embeds`: the trainer already ran the vision tower over the window's images (same weights)
            img = image_embeds if image_embeds is not None else vlm.vision_forward(batch["pixel_values"], batch["grid"], save=False)[0]
        nq, nkv, hd = t.num_heads, t.num_kv_heads, t.head_dim

        sink = dict(kp=self.kp, vp=self.vp, n=n, p_len=Pmax)      # every layer's post-rotary K / V -> shared-prefix cache

        h, _ = vlm.decoder_forward(batch["src_index"], img, vlm.full_attention(n, Pmax), batch["cos"], batch["sin"],
                                   save=False, kv_sink=sink)
        h = h.view(n, Pmax, -1)
        for gi in range(n):
            P = lens[gi]
            self.h[gi * G:(gi + 1) * G].copy_(h[gi, P - 1].float()[None, :].expand(G, -1))
            plen[gi * G:(gi + 1) * G] = P
            # generated token k of this row sits at position P + k + delta on all three axes
            delta[gi * G:(gi + 1) * G] = int(pos[:, gi, :P].max()) + 1 - P
        for gi in range(n, self.n_groups):  # unused groups: mark rows finished so they only emit pad
            self.finished[gi * G:(gi + 1) * G] = 1
            plen[gi * G:(gi + 1) * G] = 1
        self.row_plen.copy_(torch.from_numpy(plen))
        self.rope_delta.copy_(torch.from_numpy(delta))
        self.state[2] = n * G
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        self._head_and_sample(first=1)
        if logits_hook is not None:
            logits_hook(0, self.logits)
        # ---- decode: C - 1 replays of the captured step
        use_graph = self.use_graph and logits_hook is None
        if use_graph and self._graph is None:
            self._capture()
        steps_done = 0
        for sidx in range(C - 1):
            if use_graph:
                self._graph.replay()
                self.replays += 1
            else:
                self._decode_step()
                if logits_hook is not None:
                    logits_hook(sidx + 1, self.logits)
            steps_done += 1
            if not self.forbid_eos and sync_every and (sidx + 1) % sync_every == 0:
                if int(self.state[2].item()) <= 0:
                    break
        ev1.record()
        torch.cuda.synchronize()
        out = self.out_tokens[: n * G, :C].clone()
        ms = ev0.elapsed_time(ev1)
        return out, dict(decode_ms=ms, steps=steps_done + 1, rows=n * G)

    def _capture(self):
        """Warm up once on a side stream (also sets kernel attributes / fills the tensor-map cache), then capture.
        The captured step is position-independent: step / tokens / lengths are read from device memory."""
        snap = (self.state.clone(), self.tok.clone(), self.finished.clone(), self.out_tokens.clone(), self.h.clone())
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        n0 = L.launch_count()
        with torch.cuda.stream(side):
            self._decode_step()
        self.kernels_per_step = L.launch_count() - n0
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.state.copy_(snap[0]); self.tok.copy_(snap[1]); self.finished.copy_(snap[2])
        self.out_tokens.copy_(snap[3]); self.h.copy_(snap[4])
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._decode_step()
        # the capture pass did not execute; state is as restored above
        self._graph = g
