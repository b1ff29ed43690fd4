"""Does operand majorness (K-major vs MN-major smem tiles) change the tcgen05 GEMM's throughput? Same FLOPs, 4 layouts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iad_r1_b200 import lib as L
dev = torch.device("cuda:0")
bf16 = torch.bfloat16
def rnd(*s): return (torch.randn(*s, device=dev) * 0.05).to(bf16)
def t(fn, fl, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    return f"{ms*1e3:7.1f} us {fl/ms/1e9:7.1f} TF/s"
for (M, N, K) in [(8768, 2048, 22016), (22016, 2048, 8768), (8192, 8192, 8192)]:
    Kp = (K + 7) // 8 * 8
    fl = 2.0 * M * N * K
    a_k, a_mn = rnd(M, K), rnd(K, (M + 7) // 8 * 8)[:, :M].t()
    b_k, b_mn = rnd(N, K), rnd(K, N).t()
    out = torch.empty(M, N, dtype=bf16, device=dev)
    print(f"M={M} N={N} K={K}")
    for an, a in (("A K-major ", a_k), ("A MN-major", a_mn)):
        for bn, b in (("B K-major ", b_k), ("B MN-major", b_mn)):
            for nc in (False, True):
                if nc and a is a_k and b is b_k:
                    continue
                print(f"   {an} {bn} {'2-D boxes   ' if nc else 'chunked maps'}: {t(lambda: L.gemm(a, b, out=out, no_chunked_maps=nc), fl)}", flush=True)
