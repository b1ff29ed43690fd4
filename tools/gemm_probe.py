"""First-light probe for the tcgen05 GEMM: runs every layout / epilogue variant against torch and PRINTS errors
(never asserts), then times training-sized shapes. Usage (GPU box): python tools/gemm_probe.py"""
import os, sys, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iad_r1_b200 import lib as L

torch.manual_seed(0)
dev = torch.device("cuda:0")


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(torch.bfloat16)


def report(name, got, want, tol=2e-2):
    got = got.float(); want = want.float()
    err = (got - want).abs().max().item()
    ref = want.abs().max().item() + 1e-6
    bad = (~torch.isfinite(got)).sum().item()
    print(f"{'PASS' if err / ref < tol and bad == 0 else 'FAIL'} {name}: max_abs_err={err:.4g} ref_max={ref:.4g} rel={err/ref:.3g} nonfinite={bad}", flush=True)


def case(name, fn):
    try:
        fn()
        torch.cuda.synchronize()
    except Exception as e:
        print(f"ERROR {name}: {e}")
        traceback.print_exc()


def layouts(M, N, K, amn, bmn):
    a = rnd(K, M).t() if amn else rnd(M, K)
    b = rnd(K, N).t() if bmn else rnd(N, K)
    return a, b


def basic():
    for (M, N, K) in [(128, 256, 64), (128, 256, 256), (256, 512, 512), (200, 72, 1176), (1024, 1280, 1176), (333, 3420, 1280), (4096, 2048, 2048)]:
        for amn in (0, 1):
            for bmn in (0, 1):
                if (amn and M % 8) or (bmn and N % 8):
                    continue
                def f():
                    a, b = layouts(M, N, K, amn, bmn)
                    out = L.gemm(a, b)
                    report(f"gemm M{M} N{N} K{K} a_mn{amn} b_mn{bmn}", out, a.float() @ b.float().t())
                case(f"gemm {M},{N},{K},{amn},{bmn}", f)


def blockn():
    M, N, K = 300, 208, 320
    for bn in (16, 32, 48, 80, 128, 208, 256):
        def f():
            a, b = layouts(M, N, K, 0, 0)
            out = L.gemm(a, b, block_n=bn)
            report(f"block_n={bn}", out, a.float() @ b.float().t())
        case(f"bn{bn}", f)


def epilogues():
    M, N, K = 384, 512, 256
    a, b = layouts(M, N, K, 0, 0)
    want = a.float() @ b.float().t()
    bias = rnd(N); res = rnd(M, N)
    case("bias", lambda: report("bias", L.gemm(a, b, bias=bias), want + bias.float()))
    case("residual", lambda: report("residual", L.gemm(a, b, residual=res), want + res.float()))
    case("alpha f32", lambda: report("alpha f32", L.gemm(a, b, alpha=0.5, out_dtype=torch.float32), 0.5 * want, 1e-3))
    def acc():
        out = torch.ones(M, N, device=dev)
        L.gemm(a, b, out=out, accumulate=True)
        report("accumulate f32", out, want + 1, 1e-3)
    case("accumulate", acc)
    def sk():
        out = L.gemm(a, b, split_k=4, out_dtype=torch.float32)
        report("split_k=4 atomic", out, want, 1e-3)
    case("splitk", sk)
    def tr():
        out = L.gemm(a, b, trans_out=True, bias=rnd(M) * 0, bias_per_m=True)
        report("trans_out", out, want.t())
    case("trans", tr)
    def skinny():
        w = rnd(2048, 1024); x = rnd(8, 1024)
        out = L.gemm(w, x, trans_out=True, split_k=4, out_dtype=torch.float32, block_n=16)
        report("skinny swapAB splitk", out, x.float() @ w.float().t(), 1e-3)
    case("skinny", skinny)


def timing():
    for (M, N, K, amn, bmn) in [(6656, 2048, 2048, 0, 0), (6656, 22016, 2048, 0, 0), (6656, 2048, 11008, 0, 0),
                                (6656, 2048, 22016 // 2, 0, 1), (2048, 11008, 6656, 1, 1), (8192, 8192, 8192, 0, 0), (4096, 151936, 2048, 0, 0)]:
        def f():
            a, b = layouts(M, N, K, amn, bmn)
            out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
            for _ in range(3):
                L.gemm(a, b, out=out)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            n = 10
            e0.record()
            for _ in range(n):
                L.gemm(a, b, out=out)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            bt = b.t() if not bmn else b.t()
            t0.record()
            for _ in range(n):
                torch.matmul(a, b.t())
            t1.record(); torch.cuda.synchronize()
            ms_t = t0.elapsed_time(t1) / n
            print(f"TIME M{M} N{N} K{K} a_mn{amn} b_mn{bmn}: {ms:.3f} ms {2*M*N*K/ms/1e9:.1f} TFLOP/s | torch {ms_t:.3f} ms {2*M*N*K/ms_t/1e9:.1f} TFLOP/s", flush=True)
        case(f"time {M},{N},{K}", f)


if __name__ == "__main__":
    which = sys.argv[1:] or ["basic", "blockn", "epilogues", "timing"]
    for w in which:
        print(f"== {w} ==", flush=True)
        globals()[w]()
