"""Two launches for an `ncu --set full -k regex:gemm_bf16 -s 2 -c 2` capture: the gate_up forward product (bf16 TMA tile store)
and the gate_up weight gradient (fp32 TMA reduce-add) at the 3B headline shapes, each after one warm-up launch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iad_r1_b200 import lib as L  # noqa: E402

dev = torch.device("cuda:0")
T, H, I2 = 8786, 2048, 22016
x = torch.randn(T, H, device=dev).bfloat16()
w = (torch.randn(I2, H, device=dev) * 0.02).bfloat16()
out = torch.empty(T, I2, device=dev, dtype=torch.bfloat16)
dy = torch.randn(T, I2, device=dev).bfloat16()
dw = torch.zeros(I2, H, device=dev)
for _ in range(2):
    L.gemm(x, w, out=out)
    L.gemm(dy.t(), x.t(), out=dw, accumulate=True, out_dtype=torch.float32)
torch.cuda.synchronize()
