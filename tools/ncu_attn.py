"""One forward + backward of the shared-prefix attention composition (Qwen2.5-VL-3B: P = 297, G = 8, C = 512, 16 q heads,
2 kv heads, hd = 128) so that `ncu --set full -k regex:gemm_bf16` shows each batched product of the scoring pass.
TIME=1: CUDA-event time of every launch instead (no ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iad_r1_b200 import lib as L, ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
P, G, C, nq, nkv, hd = 297, 8, 512, 16, 2, 128
att = ops.SharedPrefixAttention(P, G, C, nq, nkv, hd, dev)
N = P + G * C
qkv = (torch.randn(N, (nq + 2 * nkv) * hd, device=dev) * 0.5).to(torch.bfloat16)
dattn = (torch.randn(N, nq * hd, device=dev) * 0.1).to(torch.bfloat16)
for _ in range(2):
    out, saved = att.forward(qkv)
    att.backward(dattn, qkv, saved)
torch.cuda.synchronize()
if os.environ.get("TIME"):
    import ctypes as C_
    L.check(L.lib().iadr1_gemm_profile_enable(1))
    for _ in range(5):
        out, saved = att.forward(qkv)
        att.backward(dattn, qkv, saved)
    torch.cuda.synchronize()
    tms, tfl, tmax, nl = C_.c_double(), C_.c_double(), C_.c_double(), C_.c_longlong()
    L.lib().iadr1_gemm_profile_collect.argtypes = [C_.POINTER(C_.c_double)] * 3 + [C_.POINTER(C_.c_longlong), C_.c_char_p]
    L.check(L.lib().iadr1_gemm_profile_collect(C_.byref(tms), C_.byref(tfl), C_.byref(tmax), C_.byref(nl), b"gpurun_out/attn_shapes.csv"))
    print(open("gpurun_out/attn_shapes.csv").read())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out, saved = att.forward(qkv)
    e1.record(); torch.cuda.synchronize()
    print("forward  ms", e0.elapsed_time(e1) / 10)
    e0.record()
    for _ in range(10):
        att.backward(dattn, qkv, saved)
    e1.record(); torch.cuda.synchronize()
    print("backward ms", e0.elapsed_time(e1) / 10)
else:
    out, saved = att.forward(qkv)
    att.backward(dattn, qkv, saved)
    torch.cuda.synchronize()
