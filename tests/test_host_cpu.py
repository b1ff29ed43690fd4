"""CPU tests of the host logic: config parsing, CLI surface, geometry (vs the HF functions it restates), parameter
naming round trip, C-ABI export check, and the world_size-2 gloo plumbing of the data-parallel step."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    from iad_r1_b200 import lib as L
    protos = L.header_prototypes()
    assert len(protos) >= 25
    if not os.path.exists(L._LIB_PATH):
        import __graft_entry__ as g
        g.build()
    handle = ctypes.CDLL(L._LIB_PATH)
    for name in list(protos) + ["iadr1_last_error", "iadr1_launch_count", "iadr1_reset_launch_count"]:
        assert hasattr(handle, name), f"{name} declared in include/iadr1_b200.h but not exported"
    assert handle.iadr1_version() == 100


def test_no_cpu_fallback_without_gpu():
    """The product path must fail loudly without CUDA (no oracle / torch fallback)."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from iad_r1_b200 import lib as L
    from iad_r1_b200.config import tiny_config
    from iad_r1_b200.grpo_config import GRPOConfig
    from iad_r1_b200.trainer import SCGRPOTrainer
    with pytest.raises(L.NativeLibraryError):
        SCGRPOTrainer(model=tiny_config(), reward_funcs=lambda **k: [0.0], args=GRPOConfig("/tmp/x"))
    import inspect
    import iad_r1_b200.model as m, iad_r1_b200.ops as o, iad_r1_b200.trainer as t, iad_r1_b200.rollout as r
    for mod in (m, o, t, r):
        assert "oracle" not in inspect.getsource(mod).replace("HF oracle", ""), f"{mod.__name__} must not touch oracle/"


def test_cli_surface_accepts_every_script_flag():
    from iad_r1_b200.grpo_config import GRPOConfig, ModelConfig, ScriptArguments, TrlParser
    argv = ("--deepspeed scripts/train/zero3.json --output_dir /tmp/o --model_name_or_path /m/Qwen2.5-VL-3B "
            "--dataset_name d.json --max_prompt_length 4096 --max_completion_length 512 --num_generations 4 "
            "--per_device_train_batch_size 1 --gradient_accumulation_steps 2 --logging_steps 1 --bf16 --report_to wandb "
            "--gradient_checkpointing true --attn_implementation flash_attention_2 --save_steps 100 --num_train_epochs 1 "
            "--run_name r").split()
    s, t, m = TrlParser((ScriptArguments, GRPOConfig, ModelConfig)).parse_args_and_config(argv)
    assert t.bf16 and t.gradient_checkpointing and t.num_generations == 4 and t.max_completion_length == 512
    assert t.gradient_accumulation_steps == 2 and t.report_to == ["wandb"] and m.attn_implementation == "flash_attention_2"
    d = GRPOConfig("/tmp/x")  # defaults of trl/trl/trainer/grpo_config.py
    assert (d.max_prompt_length, d.num_generations, d.max_completion_length, d.temperature, d.top_k, d.top_p) == (512, 8, 256, 0.9, 50, 1.0)
    assert (d.learning_rate, d.beta, d.num_iterations, d.epsilon, d.loss_type, d.scale_rewards) == (1e-6, 0.04, 1, 0.2, "bnpo", True)
    with pytest.raises(ValueError):
        GRPOConfig("/tmp/x", loss_type="nope")


def test_trlparser_yaml_and_env(tmp_path):
    from iad_r1_b200.grpo_config import GRPOConfig, TrlParser
    y = tmp_path / "c.yaml"
    y.write_text("env:\n  IADR1_TEST_ENV: hello\nnum_generations: 6\nbeta: 0.1\n")
    (t,) = TrlParser(GRPOConfig).parse_args_and_config(["--config", str(y), "--output_dir", "/tmp/o", "--beta", "0.2"])
    assert os.environ["IADR1_TEST_ENV"] == "hello" and t.num_generations == 6 and t.beta == 0.2


def test_config_from_both_hf_schemas():
    from iad_r1_b200.config import PRESETS, VLMConfig
    cfg = PRESETS["qwen2.5-vl-3b"]()
    flat = cfg.to_hf_dict()
    back = VLMConfig.from_hf_dict(flat)
    assert back.text == cfg.text and back.vision == cfg.vision
    nested = {"model_type": "qwen2_5_vl", "tie_word_embeddings": True, "image_token_id": 151655,
              "text_config": {"vocab_size": 151936, "hidden_size": 2048, "intermediate_size": 11008, "num_hidden_layers": 36,
                              "num_attention_heads": 16, "num_key_value_heads": 2, "rms_norm_eps": 1e-6,
                              "rope_parameters": {"rope_theta": 1e6, "mrope_section": [16, 24, 24]}},
              "vision_config": flat["vision_config"]}
    assert VLMConfig.from_hf_dict(nested).text == cfg.text
    # parameter totals reproduce the published model sizes (SURVEY.md §8 checksum)
    from iad_r1_b200.params import ParamStore
    for name, want in (("qwen2.5-vl-3b", 3.75e9), ("qwen2-vl-2b", 2.21e9), ("qwen2.5-vl-7b", 8.29e9)):
        ps = ParamStore.__new__(ParamStore)
        ps.cfg, ps.shapes, ps.decay = PRESETS[name](), {}, {}
        ps.shapes = __import__("collections").OrderedDict()
        ps._declare()
        total = sum(int(np.prod(s)) for s in ps.shapes.values())
        pad = ps.cfg.vision.depth * 3 * (ps.cfg.vision.intermediate_padded - ps.cfg.vision.intermediate_size) * ps.cfg.vision.hidden_size
        assert abs(total - pad - want) / want < 0.01, (name, total)


def test_mrope_positions_kat():
    """SURVEY.md Appendix B worked example: 448x448 image, P=320 -> axes (b, b+row, b+col), max position 79, delta -240."""
    from iad_r1_b200.config import PRESETS
    from iad_r1_b200.geometry import mrope_position_ids
    cfg = PRESETS["qwen2.5-vl-3b"]()
    ids = np.array([[11] * 23 + [cfg.vision_start_token_id] + [cfg.image_token_id] * 256 + [cfg.vision_end_token_id] + [12] * 39])
    assert ids.shape[1] == 320
    pos, delta = mrope_position_ids(ids, [(1, 32, 32)], cfg)
    b = 24
    assert (pos[0, 0, b:b + 256] == b).all()
    assert pos[1, 0, b:b + 256].tolist() == [b + r for r in range(16) for _ in range(16)]
    assert pos[2, 0, b:b + 256].tolist() == [b + c for _ in range(16) for c in range(16)]
    assert pos[:, 0, b + 256].tolist() == [b + 16] * 3
    assert pos.max() == 79 and delta[0] == -240
    with pytest.raises(ValueError):
        mrope_position_ids(ids[:, 30:], [(1, 32, 32)], cfg)   # Q14: truncating into the image span is an error


def test_vision_geometry_matches_hf():
    transformers = pytest.importorskip("transformers")
    from iad_r1_b200.config import tiny_config
    from iad_r1_b200.geometry import VisionGeometry
    from oracle.hf_oracle import build_hf_model
    cfg = tiny_config("qwen2_5_vl")
    vis = build_hf_model(cfg).model.visual
    for grid in ([(1, 8, 8)], [(1, 6, 10)], [(1, 16, 12), (1, 4, 6)]):
        geo = VisionGeometry(cfg.vision, grid, "cpu")
        wi, cu = vis.get_window_index(torch.tensor(grid))
        assert torch.equal(wi.int(), geo.window_index)
        cu = torch.unique_consecutive(torch.tensor(cu))
        assert sorted(set(geo.win_hi.tolist())) == cu[1:].tolist()
        rp = vis.rot_pos_emb(torch.tensor(grid))
        n = rp.shape[0]
        rp = rp.reshape(n // 4, 4, -1)[wi].reshape(n, -1)
        assert torch.allclose(torch.cat((rp, rp), -1).cos(), geo.cos, atol=1e-6)
        assert torch.equal(geo.reverse_index.long(), torch.argsort(wi))


def test_vision_geometry_segments_and_llava_packing_offsets():
    """Host bookkeeping behind the batched vision attention and the window-level tower pass: equal-length segment
    detection, per-image relative key ranges for ragged batches, and the anyres packing index over several images."""
    from iad_r1_b200.config import PRESETS, tiny_config
    from iad_r1_b200 import geometry as G
    v = PRESETS["qwen2.5-vl-3b"]().vision
    geo = G.VisionGeometry(v, [(1, 32, 32)] * 3, "cpu")
    assert geo.full_seg == 1024 and geo.win_seg == 64 and geo.n_patches == 3072       # 448 x 448: 16 windows of 64 patches
    lo, hi = geo.seg_ranges[64]
    assert lo.tolist() == [0] * 64 and hi.tolist() == [64] * 64
    tv = tiny_config("qwen2_5_vl").vision
    rag = G.VisionGeometry(tv, [(1, 8, 8), (1, 10, 8)], "cpu")
    assert rag.full_seg == 0 and rag.image_ranges == [(0, 64), (64, 144)]
    for kind in ("full", "win"):
        for j, (a, b) in enumerate(rag.image_ranges):
            lo_j, hi_j = rag.relative_ranges(kind, j)
            lo_abs, hi_abs = rag._abs[kind]
            assert torch.equal(lo_j + a, lo_abs[a:b]) and torch.equal(hi_j + a, hi_abs[a:b])
            assert int(lo_j.min()) >= 0 and int(hi_j.max()) <= b - a                   # keys never leave the image
    cfg = tiny_config("llava_onevision")
    tpc = cfg.vision.tokens_per_crop
    sg = G.SiglipGeometry(cfg, [(5, 80, 100), (2, 56, 56)], "cpu")
    i0, n0 = G.llava_pack_index(cfg, (80, 100))
    i1, n1 = G.llava_pack_index(cfg, (56, 56))
    assert (n0, n1) == (5, 2) and sg.n_crops == 7 and sg.n_patches == 7 * tpc and sg.n_tokens == len(i0) + len(i1)
    pk = sg.pack_index.numpy()
    assert (pk[:len(i0)] == i0).all()
    second = pk[len(i0):]
    assert (second[i1 >= 0] == i1[i1 >= 0] + n0 * tpc).all() and (second[i1 < 0] == -1).all()
    with pytest.raises(ValueError):
        G.SiglipGeometry(cfg, [(3, 80, 100)], "cpu")                                  # crop count must match the image size


def test_param_store_hf_names_roundtrip_cpu():
    from iad_r1_b200.config import tiny_config
    from iad_r1_b200.params import ParamStore
    fix = torch.load(os.path.join(ROOT, "tests", "golden", "tiny_qwen2_5_vl.pt"), map_location="cpu", weights_only=False)
    ps = ParamStore(tiny_config("qwen2_5_vl"), "cpu", with_grads=True, with_optimizer=True)
    ps.load_hf_state_dict(fix["state_dict"])
    sd = ps.hf_state_dict()
    for k, v in fix["state_dict"].items():
        k = ps.canonical_name(k)
        if k != "lm_head.weight":
            assert torch.equal(sd[k].reshape(v.shape), v), k
    assert torch.equal(ps.master, ps.flat.float())
    # vision MLP padding rows stay exactly zero
    gu = ps.p["visual.blocks.0.gate_up.weight"]
    I, Ip = ps.cfg.vision.intermediate_size, ps.cfg.vision.intermediate_padded
    assert Ip % 8 == 0 and (gu[I:Ip] == 0).all() and (gu[Ip + I:] == 0).all()
    assert ps.n_decay % 8 == 0 and all(o % 8 == 0 for o in ps.offsets.values())


def test_llava_onevision_host_geometry_matches_hf():
    """anyres bookkeeping of LLaVA-OneVision restated on the host (geometry.py) against the HF functions it follows:
    select_best_resolution, and pack_image_features' row map (crop re-tiling, unpadding, image_newline per row)."""
    transformers = pytest.importorskip("transformers")
    from transformers.image_processing_utils import select_best_resolution as hf_sbr
    from transformers.models.llava_onevision import modeling_llava_onevision as M
    from iad_r1_b200.config import PRESETS, tiny_config, VLMConfig
    from iad_r1_b200 import geometry as G
    full = PRESETS["llava-ov-0.5b"]()
    assert VLMConfig.from_hf_dict(full.to_hf_dict()).to_hf_dict() == full.to_hf_dict()
    for hw in [(448, 448), (600, 800), (100, 80), (1000, 300), (384, 384), (500, 1500)]:
        assert G.select_best_resolution(hw, full.extra["image_grid_pinpoints"]) == hf_sbr(hw, full.extra["image_grid_pinpoints"])
    assert G.image_token_count(full, (5, 448, 448)) == 3699          # 729 + 54 * (54 + 1)
    cfg = tiny_config("llava_onevision")
    tpc, H = cfg.vision.tokens_per_crop, 8

    class _Cfg:                                                      # the only fields pack_image_features reads
        class vision_config:
            image_size, patch_size = cfg.vision.image_size, cfg.vision.patch_size
        image_grid_pinpoints = cfg.extra["image_grid_pinpoints"]

    class _Self:
        config = _Cfg
    for hw in [(80, 100), (100, 80), (56, 56), (112, 112), (60, 160), (150, 50)]:
        idx, n_crops = G.llava_pack_index(cfg, hw)
        feat = torch.arange(n_crops * tpc * H, dtype=torch.float32).view(n_crops, tpc, H)
        newline = -torch.ones(H)
        want, lens = M.LlavaOnevisionModel.pack_image_features(_Self, [feat], torch.tensor([hw]), image_newline=newline)
        flat = feat.view(-1, H)
        got = torch.where(torch.from_numpy(idx)[:, None] >= 0, flat[torch.from_numpy(idx).clamp(min=0).long()], newline[None])
        assert torch.equal(got, want[0]) and int(lens[0]) == len(idx) == G.image_token_count(cfg, (n_crops, *hw)), hw
    # plain 1-D positions with padding (Qwen2 under LLaVA): cumulative count of unmasked tokens
    am = np.array([[0, 0, 1, 1, 1], [1, 1, 1, 1, 1]])
    pos, delta = G.position_ids(np.zeros((2, 5), dtype=np.int64), [], cfg, am)
    assert pos[0, 0].tolist() == [1, 1, 0, 1, 2] and pos[2, 1].tolist() == [0, 1, 2, 3, 4] and delta.tolist() == [-2, 0]
    # patchify = the Conv2d patch embedding as a GEMM
    x = torch.randn(3, 3, 56, 56)
    w = torch.randn(96, 3, 14, 14)
    ref = torch.nn.functional.conv2d(x, w, stride=14).flatten(2).transpose(1, 2).reshape(3 * 16, 96)
    assert torch.allclose(G.patchify_crops(x, 14) @ w.reshape(96, -1).t(), ref, atol=1e-4)


def test_param_store_llava_names_roundtrip_cpu():
    from iad_r1_b200.config import tiny_config
    from iad_r1_b200.params import ParamStore
    fix = torch.load(os.path.join(ROOT, "tests", "golden", "tiny_llava_onevision.pt"), map_location="cpu", weights_only=False)
    ps = ParamStore(tiny_config("llava_onevision"), "cpu", with_grads=True)
    ps.load_hf_state_dict(fix["state_dict"])
    sd = ps.hf_state_dict()
    seen = 0
    for k, v in fix["state_dict"].items():
        k = ps.canonical_name(k)
        if not k.endswith("lm_head.weight"):
            assert torch.equal(sd[k].reshape(v.shape), v), k
            seen += 1
    assert seen == len(sd)


@pytest.mark.parametrize("family", ["qwen2_vl", "qwen2_5_vl", "llava_onevision"])
def test_checkpoint_roundtrip(tmp_path, family):
    """save_model -> HF-layout directory -> load (what SC_GRPO_*.sh:25 does with the PA-SFT output), and the written
    directory loads into the HF class itself with the same tensors (the reference's from_pretrained path)."""
    from iad_r1_b200.checkpoint import load_pretrained, save_pretrained
    from iad_r1_b200.config import tiny_config
    from iad_r1_b200.params import ParamStore
    ps = ParamStore(tiny_config(family), "cpu")
    ps.init_random(seed=3)
    save_pretrained(ps, str(tmp_path / "ckpt"))
    cfg2, ps2 = load_pretrained(str(tmp_path / "ckpt"), "cpu", with_grads=False, with_optimizer=False)
    assert cfg2.text == ps.cfg.text and cfg2.vision == ps.cfg.vision
    assert torch.equal(ps.flat, ps2.flat)
    # every tensor of the written state dict has the name and shape the HF model expects
    from oracle.hf_oracle import build_hf_model
    hf = build_hf_model(ps.cfg, seed=0, dtype=torch.float32)
    want = {ps.canonical_name(k): tuple(v.shape) for k, v in hf.state_dict().items()}
    got = {k: tuple(v.shape) for k, v in ps.hf_state_dict().items()}
    for k, shp in got.items():
        assert want.get(k) == shp, (k, shp, want.get(k))
    missing = [k for k in want if k not in got and not k.endswith("lm_head.weight")]
    assert not missing, missing[:5]


_DDP = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# the data-parallel exchange of the hot path: ONE all-reduce of the flat fp32 gradient, then scale 1/world
g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
dist.all_reduce(g)
g *= 1.0 / world
assert torch.allclose(g, torch.arange(1000, dtype=torch.float32) * 1.5), g[:4]
# dataset sharding: rank r takes shuffled indices r, r + W, ...
gen = torch.Generator().manual_seed(42)
order = torch.randperm(10, generator=gen).tolist()
mine = order[rank::world]
allv = [None] * world
dist.all_gather_object(allv, mine)
assert sorted(sum(allv, [])) == list(range(10))
# metric reduction (trainer.log): packed vector mean over ranks
v = torch.tensor([float(rank), 2.0]); dist.all_reduce(v); v /= world
assert v.tolist() == [0.5, 2.0]
dist.destroy_process_group()
print("ok", rank)
'''


def test_data_parallel_plumbing_gloo_world2(tmp_path):
    script = tmp_path / "ddp.py"
    script.write_text(_DDP % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29517")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29517", str(script)], capture_output=True, text=True, env=env, timeout=180)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_sft_example_encoding_labels_assistant_only():
    """ref: llamafactory/data/processors/supervised.py:34-88 - labels are IGNORE_INDEX everywhere but assistant turns."""
    from iad_r1_b200.config import tiny_config
    from iad_r1_b200.sft_trainer import IGNORE_INDEX, SFTArguments, encode_supervised_example
    from iad_r1_b200.synthetic import SyntheticProcessor, synthetic_image
    cfg = tiny_config()
    proc = SyntheticProcessor(cfg)
    ex = {"messages": [{"role": "user", "content": "<image>Is there a defect?"}, {"role": "assistant", "content": "yes a scratch"}],
          "images": [synthetic_image(0, 112)]}
    enc = encode_supervised_example(ex, proc, 512, None, 512 * 512)
    ids, lab = enc["input_ids"], enc["labels"]
    assert (ids == cfg.image_token_id).sum() == 16 and enc["pixel_values"].shape[0] == 64
    sup = lab != IGNORE_INDEX
    assert sup.sum() >= 3 and (lab[sup] == ids[sup]).all()
    first = int(np.argmax(sup))
    assert not sup[:first].any() and (ids[:first] == cfg.image_token_id).sum() == 16   # prompt + image unlabelled
    assert ids[sup][-2] == cfg.eos_token_id or cfg.eos_token_id in ids[sup]              # end-of-turn is supervised
    with pytest.raises(ValueError):
        SFTArguments(stage="dpo")
    from transformers import HfArgumentParser
    with pytest.raises(ValueError):
        HfArgumentParser(SFTArguments).parse_args_into_dataclasses(args=["--not_a_flag", "1"])
    argv = ("--deepspeed z.json --stage sft --do_train --model_name_or_path /m --dataset D --image_dir /i --template qwen2_vl "
            "--finetuning_type full --output_dir /o --overwrite_cache --overwrite_output_dir --warmup_steps 100 --weight_decay 0.1 "
            "--per_device_train_batch_size 1 --gradient_accumulation_steps 2 --ddp_timeout 90000 --learning_rate 1e-5 "
            "--lr_scheduler_type cosine --logging_steps 5 --cutoff_len 4096 --save_steps 365 --plot_loss --num_train_epochs 1 --bf16").split()
    (a,) = HfArgumentParser(SFTArguments).parse_args_into_dataclasses(args=argv)
    assert a.cutoff_len == 4096 and a.lr_scheduler_type == "cosine" and a.bf16 and a.plot_loss


def test_sft_truncation_follows_reference_infer_seqlen():
    """Per-turn truncation to cutoff_len (ref: llamafactory/data/processors/processor_utils.py:51-65 `infer_seqlen`,
    supervised.py:50-74). The expected tuples were produced by the reference function itself in this container."""
    from iad_r1_b200.config import tiny_config
    from iad_r1_b200.sft_trainer import IGNORE_INDEX, encode_supervised_example, infer_seqlen
    from iad_r1_b200.synthetic import SyntheticProcessor
    kats = [((10, 5, 100), (10, 5)), ((300, 50, 100), (86, 14)), ((30, 500, 100), (30, 70)), ((300, 300, 100), (50, 50)),
            ((1, 1, 0), (0, 0)), ((4000, 600, 4096), (3496, 600)), ((4000, 600, 1000), (870, 130)), ((50, 49, 100), (50, 49)),
            ((100, 10, 19), (18, 1)), ((7, 93, 50), (7, 43))]
    for args, want in kats:
        assert infer_seqlen(*args) == want, args
    proc = SyntheticProcessor(tiny_config())
    ex = {"messages": [{"role": "user", "content": "is the first surface scratch " * 6},
                       {"role": "assistant", "content": "yes a scratch top left " * 5},
                       {"role": "user", "content": "and the second image " * 4},
                       {"role": "assistant", "content": "no defect " * 8}], "images": []}
    full = encode_supervised_example(ex, proc, 4096, None, 512 * 512)
    n = len(full["input_ids"])
    assert (full["labels"] != IGNORE_INDEX).sum() > 0 and len(full["labels"]) == n
    for cutoff in (n, n - 7, n // 2, 24, 9):
        enc = encode_supervised_example(ex, proc, cutoff, None, 512 * 512)
        ids, lab = enc["input_ids"], enc["labels"]
        assert len(ids) == len(lab) <= cutoff
        sup = lab != IGNORE_INDEX
        assert (lab[sup] == ids[sup]).all()
        if cutoff >= n:
            assert (ids == full["input_ids"]).all() and (lab == full["labels"]).all()
        else:
            assert len(ids) == cutoff, (cutoff, len(ids))       # the budget is used up exactly while tokens remain


@pytest.mark.slow
def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints ONE JSON line with the contract keys:
    same metric / unit / config as our arm, impl = reference, a cpu_baseline block and a zero-copy e2e block."""
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    for k in ("metric", "value", "unit", "impl", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "groups/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "qwen2.5-vl-3b" in line["config"]["workload"]


def _reference_rewards():
    """The reference's own reward callbacks (train/stage_rl/reward.py + reward_process/), imported from /root/reference
    behind a stub for `sentence_transformers` (only description_reward, unused by the default registry, needs it).
    CPU-container only: the reference tree does not exist on the GPU box."""
    import importlib
    import types
    ref = "/root/reference/train/stage_rl"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present")
    if "sentence_transformers" not in sys.modules:
        stub = types.ModuleType("sentence_transformers")
        stub.SentenceTransformer = lambda *a, **k: None
        stub.models = types.SimpleNamespace()
        stub.util = types.SimpleNamespace()
        sys.modules["sentence_transformers"] = stub
    sys.path.insert(0, ref)
    try:
        return importlib.import_module("reward")
    except Exception as e:      # anything else the reference imports at module scope
        pytest.skip(f"reference reward module not importable here: {type(e).__name__}: {e}")
    finally:
        sys.path.remove(ref)


def test_reference_reward_callbacks_drop_in(capsys):
    """Boundary (b): the reference's UNMODIFIED accuracy_reward / consistency_reward run through our calling convention
    (ref: sc_grpo_trainer.py:749-781, Q10) and through `SCGRPOTrainer._group_loss` - rewards, group advantages (unbiased
    std + 1e-4, :784-793) and the SC loss (:796-798) come out as the reference's arithmetic gives them."""
    reward = _reference_rewards()
    from collections import defaultdict
    from contextlib import nullcontext
    from types import SimpleNamespace
    from iad_r1_b200 import grpo_loss
    from iad_r1_b200.grpo_config import GRPOConfig
    from iad_r1_b200.trainer import SCGRPOTrainer, call_reward_funcs
    from oracle import grpo_ref
    yes = "<think>a</think><location>top left corner</location><type>scratch</type><answer>yes</answer>"
    texts = [yes, "<think>a</think><answer>no</answer>", "<think>a</think><location>center</location><type>hole</type><answer>yes</answer>",
             "no tags at all"]
    example = {"prompt": [{"role": "user", "content": [{"type": "image"}, {"type": "text", "text": "q"}]}], "image": ["/x.png"],
               "solution": "<answer>yes</answer><location>upper left</location><type>scratch</type>", "problem": "q", "id": 7}
    funcs = [reward.accuracy_reward, reward.consistency_reward]
    r = call_reward_funcs(funcs, example, texts, current_step=3)
    capsys.readouterr()                                    # the reference callbacks print debug lines
    assert r.shape == (4, 2)
    assert r[0, 0].item() == 2.0 and r[0, 1].item() == 1.0   # right answer (1) + (type 1.0 + location 1) / 2; format ok
    assert r[1].tolist() == [0.0, 0.0] and r[3].tolist() == [0.0, 0.0]
    assert 1.0 <= r[2, 0].item() < 2.0 and r[2, 1].item() == 1.0
    # the same callbacks through the trainer's per-group loss on the CPU (a stand-in `self`: no CUDA needed for this part)
    G, C = 4, 5
    eos = 99
    comp = torch.tensor([[5, 6, 7, eos, 0], [5, 6, eos, 0, 0], [1, 2, 3, 4, 5], [eos, 0, 0, 0, 0]])
    proc = SimpleNamespace(eos_token_id=eos, batch_decode=lambda ids, skip_special_tokens=True: texts)
    args = GRPOConfig(output_dir="/tmp/x", num_generations=G, beta=0.04)
    me = SimpleNamespace(num_generations=G, device=torch.device("cpu"), args=args, processing_class=proc, reward_funcs=funcs,
                         state=SimpleNamespace(global_step=3), beta=0.04, _metrics=defaultdict(list), _iteration=0, _old_logps={},
                         max_completion_length=C, _phase=lambda name: nullcontext())
    torch.manual_seed(0)
    logps = (torch.randn(G, C) - 3).requires_grad_(True)
    ref_logps = logps.detach() + 0.1 * torch.randn(G, C)
    loss = SCGRPOTrainer._group_loss(me, example, comp, logps, ref_logps)
    capsys.readouterr()
    mask = grpo_ref.completion_mask_ref(comp, eos)
    adv, _, _ = grpo_ref.advantages_ref(r, G)
    want, kl = grpo_ref.sc_grpo_loss_ref(logps.detach(), ref_logps, adv, mask, 0.04)
    assert torch.allclose(loss.detach(), want, atol=1e-6)
    assert abs(me._metrics["reward"][0].item() - r.sum(1).mean().item()) < 1e-6
    assert set(me._metrics) >= {"completion_length", "rewards/accuracy_reward", "rewards/consistency_reward", "reward", "reward_std", "kl"}


def test_fmha_plan_host_logic():
    """Work lists of the fused attention (iad-r1_b200/fmha.py): the query tiles reach every allowed (query, key) pair and
    the key tiles of the backward cover every allowed pair exactly once (split items add, overlaps would double count)."""
    import numpy as np
    from iad_r1_b200 import fmha
    r1, p1, s1 = fmha.shared_prefix_geometry(140, 2, 70, 0)
    r2, p2, s2 = fmha.shared_prefix_geometry(90, 2, 130, 280)
    lo = np.arange(552) // 64 * 64
    layouts = {
        "shared": fmha.shared_prefix_geometry(297, 3, 150),
        "two_groups": (np.concatenate([r1, r2]), p1 + p2, s1 + s2),
        "windows": (fmha.range_rows(lo, np.minimum(lo + 64, 552)), None, None),
        "causal": (fmha.causal_rows(2, 200), [(0, 200), (200, 400)], [(0, 200), (200, 400)]),
    }
    for name, (rng, probs, segs) in layouts.items():
        qi, ki = fmha.build_items(rng, probs, segs, nkv=2)
        N = rng.shape[0]
        mask = fmha.reference_mask(rng)
        cover = np.zeros((N, N), dtype=np.int32)
        for q0, nr, kv0, kv1, p0, p1_ in qi:
            assert 0 < nr <= 128
            rows = slice(q0, q0 + nr)
            if kv1 > kv0:
                cover[rows, kv0:min(N, kv0 + (kv1 - kv0 + 127) // 128 * 128)] += 1
            if p1_ > p0:
                cover[rows, p0:min(N, p0 + (p1_ - p0 + 127) // 128 * 128)] += 1
        assert (cover[mask] >= 1).all(), name
        # every query row belongs to exactly one query tile
        rows_cov = np.zeros(N, dtype=np.int32)
        for q0, nr, *_ in qi:
            rows_cov[q0:q0 + nr] += 1
        assert (rows_cov == 1).all(), name
        kcover = np.zeros((N, N), dtype=np.int32)
        for k0, nk, q0, q1, f0, f1 in ki:
            assert 0 < nk <= 128 and q1 > q0 and q0 % 4 == 0
            kcover[q0:q1, k0:k0 + nk] += 1
            for t in range(f0, f1):     # tiles flagged "allowed in full" really are
                assert mask[q0 + 64 * t:q0 + 64 * t + 64, k0:k0 + nk].all() and q0 + 64 * t + 64 <= q1 and nk == 128, name
        assert (kcover[mask] == 1).all() and kcover.max() == 1, name
        assert sum(f1 - f0 for *_, f0, f1 in ki) > 0 or name in ("windows",), name
    with __import__("pytest").raises(ValueError):
        fmha.build_items(np.array([[0, 0, 0, 0]], dtype=np.int32))


def test_gradient_bucket_bookkeeping():
    """Overlapped all-reduce (trainer_base): decoder-layer matrix ranges are disjoint, retire last -> first, and together with
    the remainder computed at the optimizer step cover the flat gradient exactly once."""
    from iad_r1_b200.config import tiny_config
    from iad_r1_b200.params import ParamStore
    from iad_r1_b200.trainer_base import complement_ranges
    for fam in ("qwen2_5_vl", "llava_onevision"):
        cfg = tiny_config(fam)
        ps = ParamStore(cfg, "meta")
        done = [ps.layer_matrix_range(i) for i in reversed(range(cfg.text.num_layers))]
        for (lo, hi), i in zip(done, reversed(range(cfg.text.num_layers))):
            assert lo == ps.offsets[f"layers.{i}.qkv.weight"] and hi > lo and hi <= ps.n_decay
            names = [n for n, o in ps.offsets.items() if lo <= o < hi]
            assert sorted(names) == sorted(f"layers.{i}.{k}.weight" for k in ("qkv", "o", "gate_up", "down")), names
        rest = complement_ranges(done, ps.numel)
        cover = sorted(done + rest)
        assert cover[0][0] == 0 and cover[-1][1] == ps.numel
        assert all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
    assert complement_ranges([], 10) == [(0, 10)] and complement_ranges([(0, 10)], 10) == []
    assert complement_ranges([(2, 4), (6, 8)], 10) == [(0, 2), (4, 6), (8, 10)]


def test_smart_resize_matches_hf():
    """Host half of the GPU image preprocessing: the target-size rule equals HF's `smart_resize`."""
    from transformers.models.qwen2_vl.image_processing_qwen2_vl import smart_resize as hf_smart_resize
    from iad_r1_b200.preprocess import smart_resize
    for h, w in [(448, 448), (224, 224), (300, 500), (1000, 700), (90, 120), (37, 2000), (4000, 3000), (28, 28), (15, 15)]:
        for mx in (480000, 12845056):
            assert smart_resize(h, w, 28, 3136, mx) == tuple(hf_smart_resize(h, w, 28, 3136, mx)), (h, w, mx)


def test_sft_native_data_path():
    """PA-SFT data path without LLaMA-Factory (sft_data.py): chatml templates of `--template qwen2_vl` / `llava_next_qwen`
    (ref: llamafactory/data/template.py:901-913, 1121-1133) with the default system prompt, `<image>` expansion per plugin
    (mm_plugin.py:327-367, 850-897), multi-turn pairs with prompt masking, multi-image rows, pad-to-8 collation
    (collator.py:79-161) and the reference's error messages for mismatched images."""
    import numpy as np
    from iad_r1_b200 import sft_data as D
    from iad_r1_b200.config import tiny_config
    from iad_r1_b200.sft_trainer import encode_supervised_example
    from iad_r1_b200.synthetic import SyntheticProcessor, synthetic_image
    tpl = D.get_template("qwen2_vl")
    msgs = [{"role": "user", "content": "<image>first?"}, {"role": "assistant", "content": "yes"},
            {"role": "user", "content": "and <image> this?"}, {"role": "assistant", "content": "no"}]
    exp = D.expand_image_placeholders(msgs, [3, 2], tpl)
    assert exp[0]["content"] == "<|vision_start|><|image_pad|><|image_pad|><|image_pad|><|vision_end|>first?"
    assert exp[2]["content"] == "and <|vision_start|><|image_pad|><|image_pad|><|vision_end|> this?"
    pairs = D.render_pairs(exp, tpl)
    assert pairs[0][0] == ("<|im_start|>system\nYou are a helpful assistant.<|im_end|>\n<|im_start|>user\n" + exp[0]["content"] +
                           "<|im_end|>\n<|im_start|>assistant\n")
    assert pairs[0][1] == "yes<|im_end|>\n"
    assert pairs[1] == ("<|im_start|>user\n" + exp[2]["content"] + "<|im_end|>\n<|im_start|>assistant\n", "no<|im_end|>\n")
    sys_pairs = D.render_pairs([{"role": "system", "content": "Inspect parts."}] + exp[:2], tpl)
    assert sys_pairs[0][0].startswith("<|im_start|>system\nInspect parts.<|im_end|>\n")
    lt = D.get_template(None, "llava_onevision")
    assert lt.name == "llava_next_qwen"
    assert D.expand_image_placeholders(msgs[:2], [4], lt)[0]["content"] == "<image><image><image><image>first?"
    # `--template llava` (LLaVA-1.5, template.py:832-841 "copied from vicuna"): system text straight into "USER:", eos after answers
    vt = D.get_template(None, "llava")
    assert vt.name == "llava" and vt.plugin == "llava"
    vexp = D.expand_image_placeholders(msgs[:2], [2], vt)
    vp = D.render_pairs(vexp, vt, eos="</s>")
    assert vp[0] == ("A chat between a curious user and an artificial intelligence assistant. The assistant gives helpful, "
                     "detailed, and polite answers to the user's questions.USER: <image><image>first? ASSISTANT:", "yes</s>")
    for bad in ([3], [3, 2, 1]):
        with pytest.raises(ValueError, match="number of <image> tokens|less than the number"):
            D.expand_image_placeholders(msgs, bad, tpl)
    mt = D.get_template(None, "llava_next")          # PA_SFT_LLaVA_1_6.sh: --template llava_next_mistral (template.py:885-896)
    assert mt.name == "llava_next_mistral" and D.render_pairs(msgs[:2], mt) == [("<s>[INST] <image>first?[/INST]", " yes</s>")]
    with pytest.raises(ValueError):
        D.get_template("paligemma")
    ids, labels = D.encode_pairs([([1, 2, 3], [4, 5]), ([6], [7, 8, 9])], 100)
    assert ids == [1, 2, 3, 4, 5, 6, 7, 8, 9] and labels == [-100, -100, -100, 4, 5, -100, 7, 8, 9]
    col = D.collate([dict(input_ids=ids, labels=labels), dict(input_ids=[1, 2], labels=[-100, 2])], pad_token_id=0)
    assert col["input_ids"].shape == (2, 16) and (col["labels"][1, 2:] == -100).all() and col["attention_mask"].sum() == 11
    # end to end on the twins: two images in one conversation, two assistant turns
    for fam in ("qwen2_5_vl", "llava_onevision", "llava", "llava_next"):
        cfg = tiny_config(fam)
        proc = SyntheticProcessor(cfg)
        ex = {"messages": msgs, "images": [synthetic_image(0, 112), synthetic_image(1, 84)]}
        enc = encode_supervised_example(ex, proc, 4096, None, 512 * 512, cfg=cfg)
        ids, lab = enc["input_ids"], enc["labels"]
        from iad_r1_b200.geometry import image_token_count
        n_img = sum(image_token_count(cfg, g) for g in enc["grid_thw"])
        assert (ids == cfg.image_token_id).sum() == n_img and len(enc["grid_thw"]) == 2
        sup = lab != D.IGNORE_INDEX
        assert (lab[sup] == ids[sup]).all() and (ids[sup] == cfg.eos_token_id).sum() == 2      # both answers + their <|im_end|>
        assert not sup[ids == cfg.image_token_id].any()
        # Q16: a large image is pre-shrunk with NEAREST below image_resolution pixels
        big = synthetic_image(2, 448)
        small = D.preshrink_image(big, 128 * 128, D.get_template(None, fam).plugin)
        assert small.width * small.height <= 128 * 128 and D.preshrink_image(big, 10 ** 7, "llava_next").size == big.size


def test_checkpoint_keeps_the_source_config_and_generation_config(tmp_path):
    """A directory written by `save_pretrained` carries the SOURCE checkpoint's config.json (keys this path does not model: bos,
    sliding window, max positions, eos lists) and generation_config.json unchanged, and the loaded config exposes every stop id of
    the generation config (the reference's vLLM engine stops on all of them)."""
    import json
    from iad_r1_b200.checkpoint import load_config, save_pretrained
    from iad_r1_b200.config import tiny_config
    from iad_r1_b200.params import ParamStore
    src = tmp_path / "src"
    src.mkdir()
    raw = tiny_config("qwen2_5_vl").to_hf_dict()
    raw.update({"bos_token_id": 7, "sliding_window": 4096, "max_window_layers": 3, "max_position_embeddings": 32768,
                "initializer_range": 0.02})
    gen = {"eos_token_id": [raw["eos_token_id"], 1001], "pad_token_id": raw["pad_token_id"], "do_sample": True, "top_k": 1}
    (src / "config.json").write_text(json.dumps(raw))
    (src / "generation_config.json").write_text(json.dumps(gen))
    cfg = load_config(str(src))
    assert cfg.extra["eos_token_ids"] == [raw["eos_token_id"], 1001]
    ps = ParamStore(cfg, "cpu")
    ps.init_random(seed=0)
    out = tmp_path / "out"
    save_pretrained(ps, str(out))
    assert json.loads((out / "config.json").read_text()) == raw
    assert json.loads((out / "generation_config.json").read_text()) == gen
    cfg2 = load_config(str(out))
    assert cfg2.text.hidden_size == cfg.text.hidden_size and cfg2.extra["eos_token_ids"] == [raw["eos_token_id"], 1001]


def test_reward_model_branch_scores_prompt_plus_completion():
    """Reward MODELS (ref: sc_grpo_trainer.py:232-258, 759-772): a sequence-classification network scores the chat-templated
    prompt + completion, right-padded, `logits[:, 0]`; callbacks and models can be mixed in one reward list."""
    import torch
    from iad_r1_b200.trainer import call_reward_funcs

    class Tok:
        pad_token_id = 0

        def apply_chat_template(self, messages, tokenize=False):
            return " | ".join(f"{m['role']}:{m['content'] if isinstance(m['content'], str) else m['content'][-1]['text']}" for m in messages)

        def __call__(self, texts, return_tensors, padding, padding_side, add_special_tokens):
            assert padding_side == "right" and add_special_tokens is False
            n = max(len(t) for t in texts)
            ids = torch.zeros(len(texts), n, dtype=torch.long)
            for i, t in enumerate(texts):
                ids[i, :len(t)] = torch.tensor([ord(c) % 97 + 1 for c in t])
            return {"input_ids": ids, "attention_mask": (ids > 0).long()}

    class RM(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.tensor(0.5))

        def forward(self, input_ids, attention_mask):
            class O:
                pass
            o = O()
            o.logits = (attention_mask.sum(1, keepdim=True).float() * self.w)      # score = 0.5 * length of the rendered text
            return o

    example = {"prompt": [{"role": "user", "content": [{"type": "image"}, {"type": "text", "text": "q"}]}], "solution": "<answer>no</answer>"}
    texts = ["abc", "abcdef"]
    r = call_reward_funcs([RM(), lambda prompts, completions, current_step, **kw: [1.0] * len(prompts)], example, texts, 0, [Tok(), None])
    rendered = [len(Tok().apply_chat_template(example["prompt"] + [{"role": "assistant", "content": t}])) for t in texts]
    assert r.shape == (2, 2) and torch.allclose(r[:, 0], torch.tensor(rendered, dtype=torch.float32) * 0.5) and (r[:, 1] == 1).all()
