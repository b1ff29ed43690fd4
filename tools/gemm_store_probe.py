"""Forward and weight-gradient products of the 3B decoder at the headline shapes, timed through CUDA graphs:
IADR1_GEMM_TMA_STORE=0 (per-thread row stores / read-modify-write) against the default TMA-store epilogue.

    python tools/gemm_store_probe.py ; IADR1_GEMM_TMA_STORE=0 python tools/gemm_store_probe.py
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iad_r1_b200 import lib as L
from tools.gemm_tile_probe import graph_time
dev = torch.device("cuda:0")
T = 8786
for name, M, N, K in [("gate_up fwd", T, 22016, 2048), ("down fwd", T, 2048, 11008), ("o fwd", T, 2048, 2048)]:
    x = [torch.randn(M, K, device=dev).bfloat16() for _ in range(3)]
    w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    t = graph_time(lambda i: L.gemm(x[i], w, out=out))
    print(f"{name} [{M}x{N}x{K}] {t:7.1f} us {2.0*M*N*K/t/1e6:6.0f} TFLOP/s", flush=True)
for name, N, K in [("gate_up wgrad", 22016, 2048), ("down wgrad", 2048, 11008)]:
    dy = [torch.randn(T, N, device=dev).bfloat16() for _ in range(3)]
    x = [torch.randn(T, K, device=dev).bfloat16() for _ in range(3)]
    out = torch.zeros(N, K, device=dev)
    t = graph_time(lambda i: L.gemm(dy[i].t(), x[i].t(), out=out, accumulate=True, out_dtype=torch.float32))
    print(f"{name} [{N}x{K}x{T}] {t:7.1f} us {2.0*T*N*K/t/1e6:6.0f} TFLOP/s", flush=True)
