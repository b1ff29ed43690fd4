"""Launches the dominant GEMM shapes of the step once each (after a warm-up pass) so that ONE `ncu --set full` capture
yields DRAM traffic / tensor-pipe numbers per shape. Usage (GPU box, under ncu, -k regex:gemm_bf16 -s <warm> -c <n>):
    python tools/ncu_shapes.py            # prints the launch order it used
Shapes are Qwen2.5-VL-3B, two groups per pass (M = 2 x 4393 tokens), and the decode products at ROWS (default 128) rows in flight."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iad_r1_b200 import lib as L

dev = torch.device("cuda:0")
bf16, f32 = torch.bfloat16, torch.float32
torch.manual_seed(0)


def rnd(*s):
    return (torch.randn(*s, device=dev) * 0.05).to(bf16)


M, H, I, R, V = 8786 // 2 * 2, 2048, 11008, int(os.environ.get("ROWS", "128")), 151936
x, w_gu, w_dn = rnd(M, H), rnd(2 * I, H), rnd(H, I)
act, dy = rnd(M, I), rnd(M, 2 * I)
gw = torch.zeros(2 * I, H, device=dev, dtype=f32)
xr, ar = rnd(R, H), rnd(R, I)
h32 = torch.zeros(R, H, device=dev, dtype=f32)
actr = torch.empty(R, I, device=dev, dtype=bf16)
w_head = rnd(V, H)
logits = torch.empty(R, V, device=dev, dtype=f32)
cases = [
    ("train gate_up fwd  [M,2048]x[22016,2048]^T", lambda: L.gemm(x, w_gu)),
    ("train down fwd     [M,11008]x[2048,11008]^T", lambda: L.gemm(act, w_dn)),
    ("train gate_up dgrad [M,22016]x[22016,2048]", lambda: L.gemm(dy, w_gu.t())),
    ("train gate_up wgrad [22016,M]x[M,2048] f32 accumulate", lambda: L.gemm(dy.t(), x.t(), out=gw, accumulate=True)),
    (f"decode gate_up+SwiGLU R={R}", lambda: L.gemm_swiglu(w_gu, xr, actr, block_n=R)),
    (f"decode down split-K bulk-reduce R={R}", lambda: L.gemm(w_dn, ar, out=h32, trans_out=True, split_k=9, atomic=True, block_n=R, a_static=True)),
    (f"decode lm_head R={R}", lambda: L.gemm(w_head, xr, out=logits, trans_out=True, block_n=R, a_static=True)),
]
if os.environ.get("TIME"):
    # A/B of the tile order on the training shapes (CUDA events, 10 back-to-back launches; operands >> L2)
    tcases = [
        ("gate_up fwd", lambda r: L.gemm(x, w_gu, raster=r), 2.0 * M * 2 * I * H),
        ("down fwd", lambda r: L.gemm(act, w_dn, raster=r), 2.0 * M * H * I),
        ("gate_up dgrad", lambda r: L.gemm(dy, w_gu.t(), raster=r), 2.0 * M * 2 * I * H),
        ("down dgrad", lambda r: L.gemm(x, w_dn.t(), raster=r), 2.0 * M * H * I),
        ("gate_up wgrad", lambda r: L.gemm(dy.t(), x.t(), out=gw, accumulate=True, raster=r), 2.0 * M * 2 * I * H),
        ("down wgrad", lambda r: L.gemm(x.t(), act.t(), out=gw.view(-1)[: H * I].view(H, I), accumulate=True, raster=r), 2.0 * M * H * I),
    ]
    for name, fn, fl in tcases:
        line = f"{name:14s}"
        for r in (1, 2, 0):
            for _ in range(3):
                fn(r)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn(r)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            line += f"  raster={r}: {ms * 1e3:7.1f} us {fl / ms / 1e9:7.1f} TFLOP/s"
        print(line, flush=True)
    sys.exit(0)
for name, fn in cases:       # warm-up pass (tensor-map cache, attributes)
    fn()
torch.cuda.synchronize()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for i, (name, fn) in enumerate(cases):
    flush.zero_()            # evict L2 so DRAM traffic is the cold-cache figure
    torch.cuda.synchronize()
    fn()
    torch.cuda.synchronize()
    print(f"launch {len(cases) + i}: {name}")
