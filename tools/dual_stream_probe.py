"""Experiment: two half-size rollout engines (R = 64 rows each) replaying their decode graphs concurrently on two streams
vs one engine with R = 128: does overlapping the per-kernel fixed latency of one chain with the other chain's work help?
python tools/dual_stream_probe.py [layers]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iad_r1_b200.config import PRESETS
from iad_r1_b200.params import ParamStore
from iad_r1_b200.model import VLM
from iad_r1_b200.rollout import RolloutEngine
from iad_r1_b200.synthetic import SyntheticProcessor, synthetic_dataset

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 12
dev = torch.device("cuda:0")
cfg = PRESETS["qwen2.5-vl-3b"]()
cfg.text.num_layers = layers
cfg.vision.depth = 2
cfg.vision.fullatt_block_indexes = (1,)
ps = ParamStore(cfg, dev)
ps.init_random(0)
vlm = VLM(cfg, ps)
proc = SyntheticProcessor(cfg, max_pixels=480000)
G, C, NEW = 8, 512, 96
encs = []
for ex in synthetic_dataset(16, 448):
    e = proc(text=[proc.apply_chat_template(ex["prompt"])], images=ex["image"])
    encs.append(dict(input_ids=e["input_ids"][0].numpy(), pixel_values=e["pixel_values"].to(dev), grid_thw=e["image_grid_thw"].tolist()))


def steady(engines, streams, steps=200):
    """replay every engine's captured decode graph `steps` times, engines interleaved on their streams"""
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for eng in engines:
        eng.state[0] = 200
    torch.cuda.synchronize()
    e0.record()
    for st in streams:
        st.wait_stream(torch.cuda.current_stream())
    for _ in range(steps):
        for eng, st in zip(engines, streams):
            with torch.cuda.stream(st):
                eng._graph.replay()
    for st in streams:
        torch.cuda.current_stream().wait_stream(st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3


one = RolloutEngine(vlm, 16, G, 320, C, forbid_eos=True)
one.generate(encs, seed=1, max_new_tokens=NEW)
t1 = steady([one], [torch.cuda.Stream()])
print(f"one engine  R=128: {t1:8.1f} us per decode step ({t1 / layers:.1f} us per layer incl. head)")
a = RolloutEngine(vlm, 8, G, 320, C, forbid_eos=True)
b = RolloutEngine(vlm, 8, G, 320, C, forbid_eos=True)
a.generate(encs[:8], seed=1, max_new_tokens=NEW)
b.generate(encs[8:], seed=2, max_new_tokens=NEW)
ta = steady([a], [torch.cuda.Stream()])
t2 = steady([a, b], [torch.cuda.Stream(), torch.cuda.Stream()])
print(f"one engine  R=64 : {ta:8.1f} us per decode step")
print(f"two engines R=64 + R=64 on two streams: {t2:8.1f} us per step pair (128 rows) -> {t1 / t2:.2f}x of the single R=128 engine")
for pdl in ("0",):
    os.environ["IADR1_PDL"] = pdl
    a2 = RolloutEngine(vlm, 8, G, 320, C, forbid_eos=True)
    b2 = RolloutEngine(vlm, 8, G, 320, C, forbid_eos=True)
    a2.generate(encs[:8], seed=1, max_new_tokens=NEW)
    b2.generate(encs[8:], seed=2, max_new_tokens=NEW)
    t3 = steady([a2, b2], [torch.cuda.Stream(), torch.cuda.Stream()])
    print(f"two engines, PDL off: {t3:8.1f} us per step pair")
