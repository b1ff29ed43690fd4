"""CPU tests that PIN the oracle: the restated reference arithmetic (oracle/grpo_ref.py) against the known-answer tests
the reference's vendored TRL holds for these helpers, and the product's Python loss against the oracle."""
import os

import pytest
import torch

from oracle import grpo_ref as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- trl/tests/test_utils.py:47-130 (TestPad) -----------------------------------------------------------------------
@pytest.mark.parametrize("xs,kw,expected", [
    ([[1, 2, 3], [4, 5]], dict(padding_side="left"), [[1, 2, 3], [0, 4, 5]]),
    ([[1, 2, 3], [4, 5]], dict(padding_side="right"), [[1, 2, 3], [4, 5, 0]]),
    ([[[1, 2], [3, 4]], [[5, 6]]], dict(padding_side="left"), [[[1, 2], [3, 4]], [[0, 0], [5, 6]]]),
    ([[[1, 2], [3, 4]], [[5, 6]]], dict(padding_side="right"), [[[1, 2], [3, 4]], [[5, 6], [0, 0]]]),
    ([[[1, 2], [3, 4]], [[5]]], dict(padding_side="right"), [[[1, 2], [3, 4]], [[5, 0], [0, 0]]]),
    ([[1, 2, 3], [4, 5]], dict(padding_side="right", pad_to_multiple_of=4), [[1, 2, 3, 0], [4, 5, 0, 0]]),
    ([[1, 2, 3, 4, 5], [6, 7, 8]], dict(padding_side="right", pad_to_multiple_of=4),
     [[1, 2, 3, 4, 5, 0, 0, 0], [6, 7, 8, 0, 0, 0, 0, 0]]),
    ([[1, 2, 3, 4, 5], [6, 7, 8]], dict(padding_side="left", pad_to_multiple_of=4),
     [[0, 0, 0, 1, 2, 3, 4, 5], [0, 0, 0, 0, 0, 6, 7, 8]]),
    ([[1, 2, 3, 4], [5, 6, 7, 8]], dict(padding_side="left", pad_to_multiple_of=4), [[1, 2, 3, 4], [5, 6, 7, 8]]),
])
def test_pad_kats(xs, kw, expected):
    out = R.pad_ref([torch.tensor(x) for x in xs], padding_value=0, **kw)
    assert torch.equal(out, torch.tensor(expected))


# ---- trl/tests/test_utils.py:494-512 (TestSelectiveLogSoftmax) ----------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32, torch.float16, torch.bfloat16])
def test_selective_log_softmax_kat(dtype):
    torch.manual_seed(0)
    x = torch.randn(4, 8, 32).to(dtype)
    idx = torch.randint(0, 32, (4, 8))
    want = torch.gather(x.log_softmax(-1), -1, idx.unsqueeze(-1)).squeeze(-1)
    got = R.selective_log_softmax_ref(x, idx)
    if dtype in (torch.float16, torch.bfloat16):
        assert torch.equal(got, want)
    else:
        torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)


# ---- trl/tests/test_grpo_trainer.py:36-142 (RepeatSampler) ---------------------------------------------------------------
def test_repeat_sampler_properties():
    s = R.repeat_sampler_ref(7, 2)
    assert len(s) == 14 and set(s) == set(range(7)) and all(s[i] == s[i + 1] for i in range(0, 14, 2))
    assert R.repeat_sampler_ref(7, 2, shuffle=False) == [0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6]
    s = R.repeat_sampler_ref(8, 1, batch_size=2, repeat_count=2)
    assert len(s) == 16 and all(s[i:i + 1] == s[i + 2:i + 3] for i in range(0, 16, 4))
    s = R.repeat_sampler_ref(7, 1, batch_size=2, repeat_count=2)
    assert len(s) == 12 and set(s).issubset(set(range(7)))
    s = R.repeat_sampler_ref(7, 2, batch_size=3, repeat_count=2)
    assert len(s) == 24 and s[0:6] == s[6:12] and s[12:18] == s[18:24]
    s = R.repeat_sampler_ref(7, 3, batch_size=2, repeat_count=2)
    assert len(s) == 36 and all(s[i] == s[i + 1] == s[i + 2] for i in range(0, 36, 3))
    s = R.repeat_sampler_ref(7, 2, batch_size=2, repeat_count=3)
    assert s[0:4] == s[4:8] == s[8:12] and s[24:28] == s[28:32] == s[32:36]


# ---- mask semantics with the injected completions of trl/tests/test_grpo_trainer.py:959-1008 -------------------------------
def test_completion_mask_injected_ids():
    eos, pad = 151645, 151643
    comp = torch.tensor([[1, 2, 3, 4, 5, 6, 7, 8], [9, 10, 11, eos, pad, pad, pad, pad], [12, 13, 14, 15, 16, 17, 18, eos]])
    m = R.completion_mask_ref(comp, eos)
    assert m.tolist() == [[1] * 8, [1, 1, 1, 1, 0, 0, 0, 0], [1] * 8]   # Q8: first EOS is inside the mask
    from iad_r1_b200.grpo_loss import completion_mask
    assert torch.equal(completion_mask(comp, eos), m)


def test_mask_truncated_completions_kat():
    """TRL's injected-completion case (ref: trl/tests/test_grpo_trainer.py:959-1008): ids [1..8] (no EOS, truncated),
    [9,10,11,EOS,pad x4], [12..18,EOS]; with mask_truncated_completions the first row leaves the loss."""
    from iad_r1_b200 import grpo_loss as P
    eos, pad = 151645, 151643
    comp = torch.tensor([[1, 2, 3, 4, 5, 6, 7, 8], [9, 10, 11, eos, pad, pad, pad, pad], [12, 13, 14, 15, 16, 17, 18, eos]])
    m = P.completion_mask(comp, eos)
    assert m.tolist() == [[1] * 8, [1, 1, 1, 1, 0, 0, 0, 0], [1] * 8]
    assert P.mask_truncated(m, comp, eos).tolist() == [[0] * 8, [1, 1, 1, 1, 0, 0, 0, 0], [1] * 8]


def test_advantages_and_losses_match_oracle():
    from iad_r1_b200 import grpo_loss as P
    torch.manual_seed(1)
    G, C = 8, 16
    rew = torch.rand(2 * G, 2)
    a_ref, r_ref, s_ref = R.advantages_ref(rew, G)
    a, r, s = P.group_advantages(rew, G)
    assert torch.equal(a, a_ref) and torch.equal(r, r_ref) and torch.equal(s, s_ref)
    lp = torch.randn(G, C, requires_grad=True)
    lp2 = lp.detach().clone().requires_grad_(True)
    ref = lp.detach() + 0.2 * torch.randn(G, C)
    mask = R.completion_mask_ref(torch.randint(0, 5, (G, C)), 0)
    l_ref, k_ref = R.sc_grpo_loss_ref(lp, ref, a_ref[:G], mask, 0.04)
    l, k = P.sc_grpo_loss(lp2, ref, a[:G], mask, 0.04)
    assert torch.equal(l, l_ref) and torch.equal(k, k_ref)
    l_ref.backward(); l.backward()
    assert torch.allclose(lp.grad, lp2.grad, rtol=1e-6, atol=1e-9)
    old = lp.detach() + 0.1 * torch.randn(G, C)
    for lt in ("grpo", "bnpo", "dr_grpo"):
        c_ref = R.clip_loss_ref(lp.detach(), old, ref, a_ref[:G], mask, 0.04, 0.2, 0.28, lt, C)
        c, _ = P.clip_grpo_loss(lp.detach(), old, ref, a[:G], mask, 0.04, 0.2, 0.28, lt, C)
        assert torch.equal(c, c_ref)
    # Q13: G = 1 -> unbiased std of one sample is NaN
    assert torch.isnan(P.group_advantages(torch.ones(1, 1), 1)[0]).all()


@pytest.mark.parametrize("family", ["qwen2_5_vl", "qwen2_vl", "llava_onevision"])
def test_golden_fixture_is_self_consistent(family):
    """The committed fixture reproduces from its own parts with the oracle (guards against a stale / edited fixture)."""
    fix = torch.load(os.path.join(GOLD, f"tiny_{family}.pt"), map_location="cpu", weights_only=False)
    mask = R.completion_mask_ref(fix["input_ids"][:, fix["P"]:], 1005)
    assert torch.equal(mask, fix["completion_mask"])
    adv, _, _ = R.advantages_ref(fix["rewards_per_func"], fix["G"])
    assert torch.allclose(adv, fix["advantages"])
    lp = fix["logp_fp32"].clone().requires_grad_(True)
    loss, kl = R.sc_grpo_loss_ref(lp, fix["ref_logp"], adv, mask, fix["beta"])
    assert torch.allclose(loss, fix["loss"], atol=1e-7) and torch.allclose(kl, fix["mean_kl"], atol=1e-7)
    loss.backward()
    assert torch.allclose(lp.grad, fix["dlogp"], atol=1e-7)
    # reference-form bf16 log-probs sit within bf16 noise of the fp32 oracle
    assert (fix["logp_bf16_ref"] - fix["logp_fp32"]).abs().max() < 0.1


@pytest.mark.slow
def test_golden_fixture_regenerates_from_hf():
    """Re-run HF on CPU and compare with the committed fixture (needs transformers; ~10 s)."""
    transformers = pytest.importorskip("transformers")
    from iad_r1_b200.config import tiny_config
    from oracle.hf_oracle import build_hf_model, hf_logits, per_token_logps
    fix = torch.load(os.path.join(GOLD, "tiny_qwen2_5_vl.pt"), map_location="cpu", weights_only=False)
    cfg = tiny_config("qwen2_5_vl")
    m = build_hf_model(cfg, seed=0)
    G, P = fix["G"], fix["P"]
    logits = hf_logits(m, fix["input_ids"], fix["pixel_values"].float().repeat(G, 1), torch.tensor([fix["grid"]] * G),
                       fix["position_ids"], fix["attention_mask"])
    lp = per_token_logps(logits, fix["input_ids"])[:, P - 1:]
    assert torch.allclose(lp, fix["logp_fp32"], atol=1e-5)


@pytest.mark.parametrize("family", ["qwen2_5_vl", "llava_onevision"])
def test_cpu_reference_group_step_runs(family):
    """The reference-form CPU step (HF generate + two log-prob forwards + restated loss + autograd + AdamW) that bench.py
    times as `cpu_baseline` / `--impl reference`, on the tiny twins of both processor layouts."""
    pytest.importorskip("transformers")
    from iad_r1_b200.config import tiny_config
    from iad_r1_b200.synthetic import SyntheticProcessor, synthetic_dataset
    from oracle.cpu_reference import CPUReference, reference_form_flops
    cfg = tiny_config(family)
    proc = SyntheticProcessor(cfg)
    ex = synthetic_dataset(1, 112)[0]
    enc = proc(text=[proc.apply_chat_template(ex["prompt"])], images=ex["image"])
    ids = enc["input_ids"][0]
    px, grid = (enc["pixel_values"], enc["image_sizes"].tolist()) if family == "llava_onevision" else \
        (enc["pixel_values"], enc["image_grid_thw"].tolist())
    ref = CPUReference(cfg, seed=0, threads=2)
    before = [p.detach().clone() for p in ref.policy.parameters()]
    dt, loss = ref.group_step(ids, px, grid, 4, 6, lambda comp: torch.randn(comp.shape[0]).tolist(), seed=0)
    assert dt > 0 and loss == loss
    assert any(not torch.equal(a, b) for a, b in zip(before, ref.policy.parameters())), "AdamW step must move the policy"
    assert reference_form_flops(cfg, 4, ids.shape[0], 6, 80) > 0
