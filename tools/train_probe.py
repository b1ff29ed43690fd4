"""Per-kernel GPU time of ONE scoring pass (policy forward + reference forward + backward, two packed groups, Qwen2.5-VL-3B,
decoder only - vision features supplied) via torch.profiler (CUPTI sees the kernels launched through the C ABI).
Usage (GPU box): python tools/train_probe.py [layers]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from iad_r1_b200.config import PRESETS
from iad_r1_b200.params import ParamStore
from iad_r1_b200.model import VLM

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 6
dev = torch.device("cuda:0")
cfg = PRESETS["qwen2.5-vl-3b"]()
cfg.text.num_layers = layers
cfg.vision.depth = 1
cfg.vision.fullatt_block_indexes = (0,)
ps = ParamStore(cfg, dev, with_grads=True)
ps.init_random(0)
vlm = VLM(cfg, ps)
G, C, P, n_img = 8, 512, 297, 256
torch.manual_seed(0)
groups = []
for g in range(2):
    ids = torch.randint(1000, 100000, (P,)).numpy()
    ids[24] = cfg.vision_start_token_id
    ids[25:25 + n_img] = cfg.image_token_id
    ids[25 + n_img] = cfg.vision_end_token_id
    groups.append(dict(prompt_ids=ids, completion_ids=torch.randint(1000, 100000, (G, C), dtype=torch.int32),
                       pixel_values=torch.randn(1024, cfg.vision.patch_dim), grid_thw=[(1, 32, 32)]))
batch = vlm.prepare_groups(groups)
img = (torch.randn(2 * n_img, cfg.text.hidden_size, device=dev) * 0.02).to(torch.bfloat16)
dlogp = torch.randn(2 * G * C, device=dev) * 1e-3


def one_pass():
    sink = lambda d: None
    logp, ctx = vlm.logprobs_forward(batch, batch["sel_index"], batch["labels"], save=True, image_embeds=img, dimg_sink=sink)
    vlm.logprobs_forward(batch, batch["sel_index"], batch["labels"], save=False, image_embeds=img)
    vlm.logprobs_backward(dlogp, ctx)


one_pass()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    one_pass()
    torch.cuda.synchronize()
ev = prof.key_averages()
tot = sum(e.device_time_total for e in ev)
print(f"layers={layers}: total kernel time {tot / 1e3:.1f} ms")
for e in sorted(ev, key=lambda e: -e.device_time_total)[:24]:
    print(f"{100 * e.device_time_total / tot:5.1f}%  {e.device_time_total / 1e3:8.2f} ms  {e.count:5d} x {e.device_time_total / max(1, e.count):8.1f} us  {e.key[:70]}")
