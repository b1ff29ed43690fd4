"""Weight-gradient products dW[N, K] += dy[T, N]^T x[T, K] (fp32 accumulate, MN-major operands) at the headline shapes: the
library's wave-quantisation-aware tile width (block_n = 0) against fixed 256-column tiles, same process, same clocks.

    python tools/gemm_tile_probe.py      -> one line per shape
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from iad_r1_b200 import lib as L
    dev = torch.device("cuda:0")
    T = 8786
    shapes = [("qkv (3B)", 2560, 2048, T), ("o (3B)", 2048, 2048, T), ("down (3B)", 2048, 11008, T), ("gate_up (3B)", 22016, 2048, T),
              ("vision qkv", 3840, 1280, 16384), ("vision proj", 1280, 1280, 16384), ("vision up", 6848, 1280, 16384),
              ("qkv (7B)", 4608, 3584, T), ("o (7B)", 3584, 3584, T)]
    for name, N, K, rows in shapes:
        dy = [torch.randn(rows, N, device=dev).bfloat16() for _ in range(3)]
        x = [torch.randn(rows, K, device=dev).bfloat16() for _ in range(3)]
        out = torch.zeros(N, K, device=dev)
        res = {}
        for bn in (256, 0):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(3):
                    L.gemm(dy[i].t(), x[i].t(), out=out, accumulate=True, out_dtype=torch.float32, block_n=bn)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    for i in range(3):
                        L.gemm(dy[i].t(), x[i].t(), out=out, accumulate=True, out_dtype=torch.float32, block_n=bn)
                g.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    g.replay()
                e1.record()
                torch.cuda.synchronize()
            res[bn] = 1000 * e0.elapsed_time(e1) / 30
        fl = 2.0 * N * K * rows
        print(f"{name:14s} [{N} x {K}] K={rows}: block_n 256 {res[256]:7.1f} us ({fl / res[256] / 1e6:6.0f} TFLOP/s)   auto {res[0]:7.1f} us "
              f"({fl / res[0] / 1e6:6.0f} TFLOP/s)   {res[256] / res[0]:.2f}x", flush=True)


def graph_time(fn, n=3, iters=10):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(n):
            fn(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for i in range(n):
                fn(i)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return 1000 * e0.elapsed_time(e1) / (iters * n)


def swiglu_bwd():
    """Down-projection input gradient + SwiGLU backward: GEMM -> act_mul_bwd against the fused epilogue (epi 5)."""
    from iad_r1_b200 import lib as L, ops
    dev = torch.device("cuda:0")
    for name, M, I, H in [("3B", 8786, 11008, 2048), ("7B", 8786, 18944, 3584)]:
        dy = [torch.randn(M, H, device=dev).bfloat16() for _ in range(3)]
        w = (torch.randn(H, I, device=dev) * 0.02).bfloat16()
        gu = [torch.randn(M, 2 * I, device=dev).bfloat16() for _ in range(3)]
        dact = torch.empty(M, I, device=dev, dtype=torch.bfloat16)

        def unfused(i):
            L.gemm(dy[i], w.t(), out=dact)
            ops.act_mul_bwd(dact, gu[i], I, 0, True, dgu=gu[i])
        t_gemm = graph_time(lambda i: L.gemm(dy[i], w.t(), out=dact))
        t_un = graph_time(unfused)
        t_f = graph_time(lambda i: L.gemm_swiglu_bwd(dy[i], w, gu[i]))
        print(f"swiglu backward {name} [M={M}, I={I}, H={H}]: GEMM alone {t_gemm:7.1f} us, GEMM + act_mul_bwd {t_un:7.1f} us, fused epilogue "
              f"{t_f:7.1f} us ({t_un / t_f:.2f}x)", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "swiglu_bwd":
        swiglu_bwd()
    else:
        main()
