"""GPU tests of the rollout engine and the trainer loop on the tiny twin (all through the C-ABI kernels)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tiny_trainer(cuda, family="qwen2_5_vl", **over):
    from iad_r1_b200.config import tiny_config
    from iad_r1_b200.grpo_config import GRPOConfig
    from iad_r1_b200.synthetic import SyntheticProcessor, format_reward, make_noise_reward
    from iad_r1_b200.trainer import SCGRPOTrainer
    cfg = tiny_config(family)
    kw = dict(output_dir="/tmp/iadr1_test", per_device_train_batch_size=1, gradient_accumulation_steps=2,
              num_generations=4, max_completion_length=16, max_prompt_length=512, learning_rate=1e-3, beta=0.04,
              logging_steps=1, save_strategy="no", max_steps=2, seed=7)
    kw.update(over)
    args = GRPOConfig(**kw)
    proc = SyntheticProcessor(cfg)
    return cfg, SCGRPOTrainer(model=cfg, reward_funcs=[format_reward, make_noise_reward(0)], args=args, processing_class=proc)


@pytest.mark.parametrize("family", ["qwen2_5_vl", "llava_onevision", "llava", "llava_next"])
def test_decode_matches_training_forward(cuda, family):
    """Teacher-forcing check of the whole rollout path: the log-prob the DECODE kernels assign to each sampled token
    (KV cache, prefix sharing, rope deltas, split-K atomics, fp32 residual) equals the TRAINING forward's log-prob of the
    same token. Tolerance 0.05 abs (decode keeps an fp32 residual stream, training rounds it to bf16 as HF does)."""
    from iad_r1_b200.rollout import RolloutEngine
    from iad_r1_b200.synthetic import synthetic_dataset
    cfg, tr = _tiny_trainer(cuda, family)
    data = synthetic_dataset(2, 112)
    encs = [tr._encode_prompt(ex) for ex in data]
    G, C = 4, 12
    eng = RolloutEngine(tr.model, 2, G, 192, C, temperature=0.9, top_k=50, top_p=0.9, use_cuda_graph=False)
    rec = []
    out, stats = eng.generate(encs, seed=3, logits_hook=lambda s, lg: rec.append(lg.clone()))
    assert out.shape == (2 * G, C) and len(rec) == C
    dec_logp = torch.stack([torch.log_softmax(rec[s], -1).gather(1, out[:, s].long()[:, None]).squeeze(1) for s in range(C)], 1)
    fin_after = torch.zeros(2 * G, dtype=torch.bool, device=cuda)
    for gi, enc in enumerate(encs):
        P = len(enc["input_ids"])
        comp = out[gi * G:(gi + 1) * G]
        ids = torch.cat([torch.from_numpy(enc["input_ids"]).to(cuda)[None].expand(G, -1), comp.long()], 1)
        batch = tr.model.prepare_batch(ids, enc["pixel_values"], enc["grid_thw"], prompt_len=P)
        T = P + C
        rows = (torch.arange(G, device=cuda)[:, None] * T + (P - 1) + torch.arange(C, device=cuda)[None]).reshape(-1).to(torch.int32)
        logp, _ = tr.model.logprobs_forward(batch, rows, comp.reshape(-1).to(torch.int32).contiguous(), save=False)
        logp = logp.view(G, C)
        from iad_r1_b200.grpo_loss import completion_mask
        m = completion_mask(comp.long(), cfg.eos_token_id).bool()
        err = (logp - dec_logp[gi * G:(gi + 1) * G]).abs()[m].max().item()
        print(f"\ngroup {gi}: P={P} decode-vs-train logp max err {err:.4f}")
        assert err < 0.05
        # rows that hit EOS emit pad afterwards
        assert (comp[~m] == cfg.pad_token_id).all()


def test_cuda_graph_rollout_equals_eager(cuda):
    from iad_r1_b200.rollout import RolloutEngine
    from iad_r1_b200.synthetic import synthetic_dataset
    cfg, tr = _tiny_trainer(cuda)
    enc = [tr._encode_prompt(synthetic_dataset(1, 112)[0])]
    outs = []
    for graph in (False, True):
        eng = RolloutEngine(tr.model, 1, 4, 128, 10, use_cuda_graph=graph, forbid_eos=True)
        o1, _ = eng.generate(enc, seed=5)
        o2, _ = eng.generate(enc, seed=5)      # second call reuses the captured graph
        assert torch.equal(o1, o2)
        outs.append(o1)
        assert (o1 != cfg.eos_token_id).all()
    # split-K fp32 atomics make the logits order-dependent in the last bits; sampled tokens agree except at exact ties
    assert (outs[0] == outs[1]).float().mean().item() > 0.9


def test_cuda_graph_rollout_follows_the_seed(cuda):
    """the captured decode graph froze the seed of the first generate() call, so tokens 1..C-1 of every later
    rollout reused one Philox stream. The seed now lives in device memory: a second call with another seed must change the
    tokens BEYOND position 0, and the graph engine must agree with the eager engine for that second seed."""
    from iad_r1_b200.rollout import RolloutEngine
    from iad_r1_b200.synthetic import synthetic_dataset
    cfg, tr = _tiny_trainer(cuda)
    enc = [tr._encode_prompt(synthetic_dataset(1, 112)[0])]
    res = {}
    for graph in (True, False):
        eng = RolloutEngine(tr.model, 1, 4, 128, 24, use_cuda_graph=graph, forbid_eos=True)
        a, _ = eng.generate(enc, seed=5)
        b, _ = eng.generate(enc, seed=6)          # graph engine: replays the graph captured during the seed=5 call
        a2, _ = eng.generate(enc, seed=5)
        assert torch.equal(a, a2), "same seed must reproduce the rollout"
        assert (a[:, 1:] != b[:, 1:]).float().mean().item() > 0.3, "tokens past position 0 ignore the seed"
        res[graph] = b
    assert (res[True] == res[False]).float().mean().item() > 0.9


@pytest.mark.parametrize("family", ["qwen2_5_vl", "llava_onevision", "llava", "llava_next"])
def test_trainer_two_steps(cuda, family):
    from iad_r1_b200.synthetic import synthetic_dataset
    cfg, tr = _tiny_trainer(cuda, family)
    tr.train_dataset = synthetic_dataset(8, 112)
    before = tr.params.flat.clone()
    ref_before = tr.ref_model.params.flat.clone()
    out = tr.train()
    assert out["global_step"] == 2 and tr.state.global_step == 2
    assert not torch.equal(before, tr.params.flat), "parameters did not change"
    assert torch.equal(ref_before, tr.ref_model.params.flat), "reference model must stay frozen"
    assert torch.isfinite(tr.params.flat.float()).all()
    logs = [l for l in tr.state.log_history if "loss" in l]
    assert len(logs) == 2
    for k in ("loss", "grad_norm", "learning_rate", "completion_length", "reward", "reward_std", "kl",
              "rewards/format_reward", "rewards/noise_reward"):
        assert k in logs[0], k
    assert (tr.params.grad_flat == 0).all(), "fused AdamW zeroes the gradient buffer"
    assert logs[0]["kl"] < 1e-3  # first step: reference == policy


@pytest.mark.parametrize("family", ["qwen2_5_vl", "llava_onevision"])
def test_window_vision_matches_per_pass_vision(cuda, family):
    """Running the vision tower once per accumulation window (features shared by the rollout prefill and every scoring
    pass, one backward over the accumulated feature gradient) gives the same completions and the same accumulated
    gradient as running it inside every pass (bf16 summation order aside)."""
    from iad_r1_b200.synthetic import synthetic_dataset
    data = synthetic_dataset(4, 112)
    grads, base = [], None
    for wv in (True, False):
        cfg, tr = _tiny_trainer(cuda, family, window_vision=wv, per_device_train_batch_size=2, gradient_accumulation_steps=2)
        tr.prepare_window(data)
        assert (tr._window is not None) == wv
        got = [tr._rollout_cache[id(ex)][1] for ex in data]
        if base is None:
            base = [c.clone() for c in got]
        else:
            # same features -> same sampled completions, except where the split-K reductions' nondeterministic summation
            # order moves a logit across an exact sampling tie; the gradient comparison scores the SAME completions
            assert (torch.cat(got) == torch.cat(base)).float().mean().item() > 0.9
            for ex, c in zip(data, base):
                tr._rollout_cache[id(ex)] = (tr._rollout_cache[id(ex)][0], c)
        for j in range(0, 4, 2):
            tr.training_step(data[j:j + 2])
        tr._flush_window_vision()
        torch.cuda.synchronize()
        grads.append(tr.params.grad_flat.clone())
    cosv = torch.nn.functional.cosine_similarity(grads[0], grads[1], dim=0).item()
    rel = ((grads[0] - grads[1]).norm() / grads[1].norm()).item()
    print(f"\n[{family}] window-vision vs per-pass gradient: cos {cosv:.6f}, rel {rel:.4f}")
    assert cosv > 0.9995 and rel < 0.03
    vis = tr.params.g["visual.blocks.0.qkv.weight"]
    assert vis.abs().sum() > 0


def test_clip_mode_multi_iteration(cuda):
    """Vendored-TRL semantics behind GRPOConfig.num_iterations (ref: trl/trl/trainer/grpo_trainer.py:872-901, 1182-1219):
    one rollout per window, num_iterations optimizer steps over it, the ratio of the 2nd pass is taken against the
    log-probs recorded in the 1st; truncated rows are masked out when mask_truncated_completions is set."""
    from iad_r1_b200.synthetic import synthetic_dataset
    cfg, tr = _tiny_trainer(cuda, loss_mode="clip", num_iterations=2, max_steps=4, loss_type="bnpo",
                            mask_truncated_completions=False, learning_rate=5e-3)
    tr.train_dataset = synthetic_dataset(8, 112)
    calls = []
    orig = tr._rollout
    tr._rollout = lambda enc, **kw: (calls.append(len(enc)), orig(enc, **kw))[1]
    out = tr.train()
    assert out["global_step"] == 4
    assert len(calls) == 2, "two windows -> two rollouts for four optimizer steps"
    assert not tr._rollout_cache and not tr._old_logps
    assert torch.isfinite(tr.params.flat.float()).all()
    # all rows truncated (EOS forbidden) + mask_truncated_completions -> zero loss, zero gradient, parameters unchanged
    cfg, tr = _tiny_trainer(cuda, loss_mode="clip", mask_truncated_completions=True, rollout_forbid_eos=True, beta=0.0,
                            max_steps=1, loss_type="grpo")
    tr.train_dataset = synthetic_dataset(4, 112)
    before = tr.params.flat.clone()
    tr.train()
    assert torch.equal(before, tr.params.flat)


def test_all_masked_gradient_is_zero(cuda):
    """TRL's 'no parameter change when every advantage is zero' twin (ref: trl/tests/test_grpo_trainer.py:1010-1047):
    a constant reward gives zero advantages, and with beta = 0 the gradient must vanish exactly."""
    from iad_r1_b200.synthetic import synthetic_dataset
    from iad_r1_b200.config import tiny_config
    from iad_r1_b200.grpo_config import GRPOConfig
    from iad_r1_b200.synthetic import SyntheticProcessor
    from iad_r1_b200.trainer import SCGRPOTrainer

    def const_reward(prompts, completions, **kw):
        return [1.0] * len(completions)

    cfg = tiny_config("qwen2_5_vl")
    args = GRPOConfig(output_dir="/tmp/iadr1_test", gradient_accumulation_steps=1, num_generations=4,
                      max_completion_length=8, beta=0.0, logging_steps=1, save_strategy="no", max_steps=1)
    tr = SCGRPOTrainer(model=cfg, reward_funcs=const_reward, args=args, processing_class=SyntheticProcessor(cfg))
    tr.train_dataset = synthetic_dataset(2, 112)
    before = tr.params.flat.clone()
    tr.train()
    assert torch.equal(before, tr.params.flat)
    assert tr.ref_model is None


def test_sft_trainer_reduces_loss(cuda, tmp_path):
    """PA-SFT step: teacher-forced CE on assistant tokens through the same kernels; a few steps on two
    fixed examples must lower the loss, frozen-tower quirk Q17 holds for the qwen2_vl family only."""
    from iad_r1_b200.config import tiny_config
    from iad_r1_b200.sft_trainer import PASFTTrainer, SFTArguments
    from iad_r1_b200.synthetic import SyntheticProcessor, synthetic_image
    for family, frozen in (("qwen2_5_vl", False), ("qwen2_vl", True), ("llava_onevision", False), ("llava", False), ("llava_next", False)):
        cfg = tiny_config(family)
        data = [{"messages": [{"role": "user", "content": "<image>Is there a defect in the image?"},
                              {"role": "assistant", "content": "<think> the surface is scratch </think> <answer> yes </answer>"}],
                 "images": [synthetic_image(i, 112)]} for i in range(2)]
        args = SFTArguments(output_dir=str(tmp_path / family), do_train=True, learning_rate=2e-3, lr_scheduler_type="cosine",
                            warmup_steps=1, weight_decay=0.1, gradient_accumulation_steps=2, logging_steps=1, max_steps=6,
                            save_strategy="no", cutoff_len=512, bf16=True)
        tr = PASFTTrainer(cfg, args, train_dataset=data, processing_class=SyntheticProcessor(cfg))
        vis_before = tr.params.p["visual.blocks.0.qkv.weight"].clone()
        txt_before = tr.params.p["layers.0.qkv.weight"].clone()
        tr.train()
        losses = [l["loss"] for l in tr.state.log_history if "loss" in l]
        assert len(losses) == 6 and losses[-1] < losses[0] - 0.2, losses
        assert tr.freeze_vision == frozen
        assert torch.equal(vis_before, tr.params.p["visual.blocks.0.qkv.weight"]) == frozen
        assert not torch.equal(txt_before, tr.params.p["layers.0.qkv.weight"])


def test_multi_group_pass_equals_single_group_passes(cuda):
    """Packing two groups into one forward/backward (per_device_train_batch_size=2) gives the same log-probs and the
    same accumulated gradient as two single-group passes (up to bf16 summation order in the weight-gradient GEMMs)."""
    from iad_r1_b200.synthetic import synthetic_dataset
    cfg, tr = _tiny_trainer(cuda)
    data = synthetic_dataset(2, 112)
    encs = [tr._encode_prompt(ex) for ex in data]
    torch.manual_seed(0)
    comps = [torch.randint(10, 900, (4, 10), device=cuda, dtype=torch.int32) for _ in range(2)]
    groups = [dict(prompt_ids=e["input_ids"], completion_ids=c, pixel_values=e["pixel_values"], grid_thw=e["grid_thw"])
              for e, c in zip(encs, comps)]
    merged = tr.model.prepare_groups(groups)
    lp_m, ctx_m = tr.model.logprobs_forward(merged, merged["sel_index"], merged["labels"])
    d = torch.randn_like(lp_m)
    tr.params.zero_grad()
    tr.model.logprobs_backward(d, ctx_m)
    g_merged = tr.params.grad_flat.clone()
    tr.params.zero_grad()
    lps = []
    for g_, (lo, hi) in zip(groups, merged["group_slices"]):
        b = tr.model.prepare_groups([g_])
        lp, ctx = tr.model.logprobs_forward(b, b["sel_index"], b["labels"])
        lps.append(lp)
        tr.model.logprobs_backward(d[lo:hi].contiguous(), ctx)
    assert (torch.cat(lps) - lp_m).abs().max().item() < 2e-3
    rel = ((tr.params.grad_flat - g_merged).norm() / g_merged.norm()).item()
    assert rel < 2e-2, rel


def test_two_image_prompt_layouts_agree(cuda):
    """The 1-shot prompt of the reference (`--single_img 0`: a reference image and a test image in one user turn,
    ref: train/stage_rl/grpo_ad.py:92-116) - two images of different sizes per prompt: the shared-prefix layout and the
    reference [G, P + C] layout give the same log-probs, and a trainer step runs on it."""
    from iad_r1_b200.synthetic import synthetic_dataset, synthetic_image
    cfg, tr = _tiny_trainer(cuda, max_steps=1, gradient_accumulation_steps=1)
    rows = synthetic_dataset(2, 112)
    for i, ex in enumerate(rows):
        ex["image"] = [synthetic_image(10 + i, 112), synthetic_image(20 + i, 84)]
        ex["prompt"][0]["content"] = [{"type": "image"}, {"type": "image"}, ex["prompt"][0]["content"][-1]]
    enc = tr._encode_prompt(rows[0])
    assert len(enc["grid_thw"]) == 2 and enc["grid_thw"][0] != enc["grid_thw"][1]
    G, C, P = 4, 6, len(enc["input_ids"])
    comp = torch.randint(10, 900, (G, C), device=cuda, dtype=torch.int32)
    b1 = tr.model.prepare_group(enc["input_ids"], comp, enc["pixel_values"], enc["grid_thw"])
    lp1, _ = tr.model.logprobs_forward(b1, b1["sel_index"], b1["labels"], save=False)
    ids = torch.cat([torch.from_numpy(enc["input_ids"]).to(cuda)[None].expand(G, -1), comp.long()], 1)
    b2 = tr.model.prepare_batch(ids, enc["pixel_values"], enc["grid_thw"], prompt_len=P)
    sel = (torch.arange(G, device=cuda)[:, None] * (P + C) + (P - 1) + torch.arange(C, device=cuda)[None]).reshape(-1).to(torch.int32)
    lp2, _ = tr.model.logprobs_forward(b2, sel, comp.reshape(-1).contiguous(), save=False)
    assert (lp1 - lp2).abs().max().item() < 0.02
    tr.train_dataset = rows
    out = tr.train()
    assert out["global_step"] == 1 and torch.isfinite(tr.params.flat.float()).all()


@pytest.mark.parametrize("family", ["llava_onevision", "qwen2_5_vl", "llava"])
def test_two_stage_pa_sft_then_sc_grpo(cuda, tmp_path, family):
    """BASELINE config 5, the chain of the reference's launch scripts: PA-SFT writes an HF-layout directory
    (`--output_dir`, PA_SFT_*.sh), SC-GRPO is pointed at it by path (`MODEL_NAME_OR_PATH=<PA-SFT dir>`,
    ref: scripts/train/SC_GRPO/SC_GRPO_LLaVA_OneVision_SI_0.5B.sh:25). The directory name matches none of the family
    substrings (as the reference's `..._PA_SFT` names do not): the family is read from config.json."""
    from iad_r1_b200.config import tiny_config
    from iad_r1_b200.grpo_config import GRPOConfig
    from iad_r1_b200.sft_trainer import PASFTTrainer, SFTArguments
    from iad_r1_b200.synthetic import SyntheticProcessor, format_reward, make_noise_reward, synthetic_dataset, synthetic_image
    from iad_r1_b200.trainer import SCGRPOTrainer
    cfg = tiny_config(family)
    sft_dir = str(tmp_path / "Expert_AD_PA_SFT")
    data = [{"messages": [{"role": "user", "content": "<image>Is there a defect in the image?"},
                          {"role": "assistant", "content": "<think> the surface is scratch </think> <answer> yes </answer>"}],
             "images": [synthetic_image(i, 112)]} for i in range(2)]
    sargs = SFTArguments(output_dir=sft_dir, do_train=True, learning_rate=1e-3, gradient_accumulation_steps=1, logging_steps=1,
                         max_steps=2, save_strategy="no", cutoff_len=512, bf16=True)
    sft = PASFTTrainer(cfg, sargs, train_dataset=data, processing_class=SyntheticProcessor(cfg))
    sft.train()
    sft.save_model(sft_dir)
    stage1 = sft.params.flat.clone()
    del sft
    gargs = GRPOConfig(output_dir=str(tmp_path / "grpo"), per_device_train_batch_size=1, gradient_accumulation_steps=2,
                       num_generations=4, max_completion_length=12, learning_rate=1e-3, beta=0.04, logging_steps=1,
                       save_strategy="no", max_steps=1, seed=3)
    tr = SCGRPOTrainer(model=sft_dir, reward_funcs=[format_reward, make_noise_reward(0)], args=gargs,
                       processing_class=SyntheticProcessor(cfg), train_dataset=synthetic_dataset(2, 112))
    assert tr.cfg.family == family
    assert torch.equal(tr.params.flat, stage1), "stage 2 must start from the stage-1 weights bit for bit"
    assert torch.equal(tr.ref_model.params.flat, stage1), "the KL reference is the PA-SFT policy"
    out = tr.train()
    assert out["global_step"] == 1 and not torch.equal(tr.params.flat, stage1)
    assert torch.isfinite(tr.params.flat.float()).all()


def test_activation_recompute_gives_the_same_gradient(cuda):
    """Per-layer recompute (`activation_recompute="on"`, what --gradient_checkpointing asks of the reference and the 7B memory
    plan switches on) re-runs the same deterministic kernels: log-probs and the accumulated gradient equal the
    resident-activation path's."""
    from iad_r1_b200.synthetic import synthetic_dataset
    cfg, tr = _tiny_trainer(cuda)
    data = synthetic_dataset(1, 112)
    enc = tr._encode_prompt(data[0])
    torch.manual_seed(0)
    comp = torch.randint(10, 900, (4, 10), device=cuda, dtype=torch.int32)
    batch = tr.model.prepare_groups([dict(prompt_ids=enc["input_ids"], completion_ids=comp, pixel_values=enc["pixel_values"],
                                          grid_thw=enc["grid_thw"])])
    res = []
    for rc in (False, True):
        tr.model.recompute = rc
        tr.params.zero_grad()
        lp, ctx = tr.model.logprobs_forward(batch, batch["sel_index"], batch["labels"])
        assert ctx["dctx"].mode == (2 if rc else 1)
        ws_bytes = ctx["dctx"].ws.numel()
        if rc:
            assert ws_bytes < 0.8 * res_bytes, "only the layer inputs (+ one layer of scratch) may stay resident"
        res_bytes = ws_bytes
        tr.model.logprobs_backward(torch.linspace(-1, 1, lp.numel(), device=cuda), ctx)
        torch.cuda.synchronize()
        res.append((lp.clone(), tr.params.grad_flat.clone()))
    assert torch.equal(res[0][0], res[1][0])
    rel = ((res[0][1] - res[1][1]).norm() / res[0][1].norm()).item()
    assert rel < 1e-5, rel


def test_resume_from_checkpoint_continues_the_run(cuda, tmp_path):
    """`--save_steps` checkpoints carry fp32 master weights, Adam moments and the step counters (trainer_base.save_checkpoint);
    `train(resume_from_checkpoint=...)` continues with the same data order, learning-rate schedule and optimizer state: the
    resumed run ends where the uninterrupted one does (PA-SFT: no sampling, so up to fp32 summation order)."""
    import os
    from iad_r1_b200.config import tiny_config
    from iad_r1_b200.sft_trainer import PASFTTrainer, SFTArguments
    from iad_r1_b200.synthetic import SyntheticProcessor, synthetic_image
    cfg = tiny_config("qwen2_5_vl")
    data = [{"messages": [{"role": "user", "content": "<image>Is there a defect in the image?"},
                          {"role": "assistant", "content": f"<think> region {i} looks scratched </think> <answer> yes </answer>"}],
             "images": [synthetic_image(i, 112)]} for i in range(6)]

    def make(out, **kw):
        args = SFTArguments(output_dir=str(out), do_train=True, learning_rate=2e-3, lr_scheduler_type="cosine", warmup_steps=1,
                            weight_decay=0.1, gradient_accumulation_steps=2, logging_steps=1, max_steps=3, cutoff_len=512,
                            bf16=True, **kw)
        return PASFTTrainer(cfg, args, train_dataset=data, processing_class=SyntheticProcessor(cfg))

    full = make(tmp_path / "full", save_strategy="steps", save_steps=2)
    full.train()
    ckpt = os.path.join(str(tmp_path / "full"), "checkpoint-2")
    assert os.path.isfile(os.path.join(ckpt, "optimizer.pt")) and os.path.isfile(os.path.join(ckpt, "trainer_state.json"))
    resumed = make(tmp_path / "resumed", save_strategy="no")
    resumed.train(resume_from_checkpoint=ckpt)
    assert resumed.state.global_step == 3 and resumed._opt_step == 3
    d = (resumed.params.master - full.params.master).abs().max().item()
    moved = (full.params.master - make(tmp_path / "init", save_strategy="no").params.master).abs().max().item()
    print(f"\nresume: max |master_resumed - master_uninterrupted| = {d:.2e} (parameters moved by up to {moved:.2e})")
    assert d <= 1e-3 * moved + 1e-7
    l_full = [l["loss"] for l in full.state.log_history if "loss" in l]
    l_res = [l["loss"] for l in resumed.state.log_history if "loss" in l]
    assert len(l_res) == 3 and abs(l_res[-1] - l_full[-1]) < 1e-3
    with pytest.raises(FileNotFoundError):
        make(tmp_path / "bad", save_strategy="no").train(resume_from_checkpoint=str(tmp_path / "init"))


@pytest.mark.parametrize("family", ["qwen2_5_vl", "llava_onevision", "llava"])
def test_decode_chain_matches_per_op_kernels(cuda, monkeypatch, family):
    """The persistent decode-layer chain (csrc/decode_chain.cu: o -> norm -> gate_up + SwiGLU -> down -> norm -> qkv in ONE
    launch per layer, phases ordered by global counters) against the one-kernel-per-op decode step (IADR1_DECODE_CHAIN=0) on
    the same engine state: same logits at every step up to fp32 summation order, same sampled tokens."""
    from iad_r1_b200.rollout import RolloutEngine
    from iad_r1_b200.synthetic import synthetic_dataset
    cfg, tr = _tiny_trainer(cuda, family)
    vlm = tr.model
    encs = [tr._encode_prompt(ex) for ex in synthetic_dataset(3, 112)]
    G, C = 4, 10
    pmax = (max(len(e["input_ids"]) for e in encs) + 63) // 64 * 64
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("IADR1_DECODE_CHAIN", mode)
        for graph in (False, True):
            eng = RolloutEngine(vlm, 3, G, pmax, C, use_cuda_graph=graph, forbid_eos=True)
            rec = []
            out, _ = eng.generate(encs, seed=7, logits_hook=(None if graph else (lambda s_, lg: rec.append(lg.float().cpu().clone()))))
            torch.cuda.synchronize()
            res[(mode, graph)] = (out.cpu().clone(), rec)
    monkeypatch.delenv("IADR1_DECODE_CHAIN")
    out1, rec1 = res[("1", False)]
    out0, rec0 = res[("0", False)]
    assert torch.equal(res[("1", True)][0], out1), "graph replay of the chained step differs from eager"
    worst = 0.0
    for k in range(C):
        same = (out1[:, :k] == out0[:, :k]).all(dim=1) if k else torch.ones(out1.shape[0], dtype=torch.bool)
        if same.any():
            worst = max(worst, (rec1[k][same] - rec0[k][same]).abs().max().item())
    frac = (out1 == out0).float().mean().item()
    print(f"\n[{family}] chained vs per-op decode: max logit diff {worst:.2e} on matching prefixes, {frac * 100:.1f}% identical tokens")
    assert worst < 2e-2 and frac > 0.9
