"""HBM roofline of the non-GEMM training kernels at the headline shapes (Qwen2.5-VL-3B, two groups per pass:
8786 token rows, H = 2048, I = 11008, 16 + 2 heads of 128): CUDA-event time per launch over inputs larger than L2
(a ring of buffers), algorithmic bytes / time against MEASURED_PEAKS.json's HBM copy bandwidth.

    python tools/rowops_probe.py [--rows 8786] [--iters 20]      -> one line per kernel + a JSON summary line
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=8786)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--eager", action="store_true", help="plain launches, no timing (for an ncu capture)")
    a = ap.parse_args()
    from iad_r1_b200 import ops
    dev = torch.device("cuda:0")
    bf16, f32 = torch.bfloat16, torch.float32
    T, H, I, nq, nkv, hd = a.rows, 2048, 11008, 16, 2, 128
    QKV = (nq + 2 * nkv) * hd
    peak = 6449.1
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f).get("hbm_gbs", peak))
    except Exception:
        pass
    RING = 6                                               # 6 x (>= 36 MB per operand) > 126 MB of L2

    def ring(*shape, dtype=bf16, scale=1.0, n=RING):
        return [(torch.randn(*shape, device=dev) * scale).to(dtype) for _ in range(n)]

    results = []

    def timed(name, nbytes, fn):
        """One CUDA graph of RING launches (one per ring buffer) replayed `iters` times: the Python / ctypes issue time of a
        launch (~20 us) would otherwise hide every kernel shorter than that."""
        if a.only and a.only not in name:
            return
        if a.eager:                                          # ncu capture: plain launches
            for i in range(a.iters):
                fn(i % RING)
            torch.cuda.synchronize()
            return
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(RING):
                fn(i)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for i in range(RING):
                    fn(i)
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
        us = 1000 * e0.elapsed_time(e1) / (a.iters * RING)
        gbs = nbytes / us / 1e3
        results.append({"kernel": name, "us": round(us, 1), "MB": round(nbytes / 1e6, 1), "GB/s": round(gbs), "frac": round(gbs / peak, 3)})
        print(f"{name:38s} {us:8.1f} us  {nbytes / 1e6:8.1f} MB  {gbs:7.0f} GB/s  {gbs / peak:5.2f} of HBM", flush=True)

    x, dy, dx = ring(T, H), ring(T, H), ring(T, H)
    w = torch.ones(H, device=dev, dtype=bf16)
    y = torch.empty(T, H, device=dev, dtype=bf16)
    dw = torch.zeros(H, device=dev, dtype=f32)
    rstd = [ops.rmsnorm_fwd(x[i], w, 1e-6, out=y)[1] for i in range(RING)]
    timed("rmsnorm_fwd [T,2048]", T * H * 2 * 2, lambda i: ops.rmsnorm_fwd(x[i], w, 1e-6, out=y))
    timed("rmsnorm_bwd [T,2048] (dx +=)", T * H * 2 * 4, lambda i: ops.rmsnorm_bwd(dy[i], x[i], w, rstd[i], dx[i], dw, True))
    timed("rmsnorm_bwd [T,2048] (dx +=, no dw)", T * H * 2 * 4, lambda i: ops.rmsnorm_bwd(dy[i], x[i], w, rstd[i], dx[i], None, True))
    timed("rmsnorm_bwd [T,2048] (dx =)", T * H * 2 * 3, lambda i: ops.rmsnorm_bwd(dy[i], x[i], w, rstd[i], dx[i], dw, False))

    gu, dact = ring(T, 2 * I, n=3), ring(T, I, n=3)
    act = torch.empty(T, I, device=dev, dtype=bf16)
    dgu = torch.empty(T, 2 * I, device=dev, dtype=bf16)
    timed("act_mul_fwd (SwiGLU) [T,11008]", T * I * 2 * 3, lambda i: ops.act_mul_fwd(gu[i % 3], I, 0, True, out=act))
    timed("act_mul_bwd (SwiGLU) [T,11008]", T * I * 2 * 5, lambda i: ops.act_mul_bwd(dact[i % 3], gu[i % 3], I, 0, True, dgu=dgu))

    qkv = ring(T, QKV)
    cos = torch.randn(T, hd, device=dev, dtype=f32)
    sin = torch.randn(T, hd, device=dev, dtype=f32)
    timed("rope (q, k heads in place)", T * (nq + nkv) * hd * 2 * 2, lambda i: ops.rope_(qkv[i], cos, sin, nq + nkv, hd, True))

    a1, a2 = ring(T, H), ring(T, H)
    timed("add_bf16 [T,2048]", T * H * 2 * 3, lambda i: ops.add_bf16(a1[i], a2[i], out=y))
    cs = torch.zeros(QKV, device=dev, dtype=f32)
    timed("colsum [T,2560] (bias gradient)", T * QKV * 2, lambda i: ops.colsum(qkv[i], cs))
    print(json.dumps({"peak_gbs": peak, "rows": T, "results": results}), flush=True)


if __name__ == "__main__":
    main()
