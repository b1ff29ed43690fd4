"""ctypes binding of libiadr1_b200.so (the C ABI in include/iadr1_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, this raises. The product path
never routes through ``oracle/`` or a torch re-implementation ("no CPU fallback").
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libiadr1_b200.so")
_lib = None


class NativeLibraryError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("batch", C.c_int), ("batch_lo", C.c_int), ("b_lo_div", C.c_int),
        ("A", C.c_void_p), ("lda", C.c_longlong), ("a_bs_lo", C.c_longlong), ("a_bs_hi", C.c_longlong), ("a_mn", C.c_int),
        ("B", C.c_void_p), ("ldb", C.c_longlong), ("b_bs_lo", C.c_longlong), ("b_bs_hi", C.c_longlong), ("b_mn", C.c_int),
        ("C", C.c_void_p), ("ldc", C.c_longlong), ("c_bs_lo", C.c_longlong), ("c_bs_hi", C.c_longlong),
        ("c_f32", C.c_int), ("trans_c", C.c_int), ("accumulate", C.c_int), ("atomic", C.c_int), ("split_k", C.c_int),
        ("alpha", C.c_float),
        ("bias", C.c_void_p), ("bias_per_m", C.c_int),
        ("residual", C.c_void_p),
        ("kmode", C.c_int), ("skip_mode", C.c_int), ("causal_off", C.c_int),
        ("epi", C.c_int),
        ("labels", C.c_void_p), ("part_max", C.c_void_p), ("part_sum", C.c_void_p), ("tgt_logit", C.c_void_p),
        ("lse_tiles_n", C.c_int),
        ("lse", C.c_void_p), ("gscale", C.c_void_p),
        ("block_n", C.c_int), ("stages", C.c_int), ("max_ctas", C.c_int), ("a_static", C.c_int),
        ("stream_k", C.c_int), ("up_row_off", C.c_int), ("raster", C.c_int), ("no_chunked_maps", C.c_int), ("no_bulk_red", C.c_int), ("co_resident", C.c_int),
        ("gu_out", C.c_void_p), ("gu_ld", C.c_longlong),
    ]


def lib():
    """Load (once) and return the ctypes handle. Raises NativeLibraryError when the .so is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise NativeLibraryError(
                f"{_LIB_PATH} not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc for sm_100a). There is no CPU or PyTorch fallback for the hot path.")
        L = C.CDLL(_LIB_PATH)
        L.iadr1_last_error.restype = C.c_char_p
        L.iadr1_launch_count.restype = C.c_longlong
        L.iadr1_gemm_bf16.argtypes = [C.POINTER(GemmDesc), C.c_void_p]
        _declare(L)
        _lib = L
    return _lib


HEADER = os.path.join(os.path.dirname(_HERE), "include", "iadr1_b200.h")
_CTYPES = {"int": C.c_int, "long long": C.c_longlong, "float": C.c_float, "unsigned long long": C.c_ulonglong,
           "iadr1_layer_cb": C.c_void_p}


def header_prototypes(path: str = HEADER) -> dict:
    """Parse `int iadr1_xxx(args);` prototypes out of the public header -> {name: [ctypes argtypes]}."""
    import re
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\bint\s+(iadr1_\w+)\s*\(([^)]*)\)\s*;", text):
        name, args = m.group(1), m.group(2).strip()
        types = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "char*" in a.replace(" *", "*"):
                    types.append(C.c_char_p)
                elif "*" in a:
                    types.append(C.c_void_p)
                else:
                    base = " ".join(a.replace("const ", "").split()[:-1])
                    types.append(_CTYPES[base])
        protos[name] = types
    return protos


def _declare(L):
    """argtypes for every `int iadr1_*(...)` entry point the header declares (all return int status)."""
    for name, args in header_prototypes().items():
        fn = getattr(L, name, None)
        if fn is None:
            raise NativeLibraryError(f"{_LIB_PATH} does not export {name}; rebuild the library")
        if name == "iadr1_gemm_bf16":
            continue
        fn.argtypes = args
        fn.restype = C.c_int


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise NativeLibraryError(f"{what}: {lib().iadr1_last_error().decode()}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(lib().iadr1_launch_count())


def reset_launch_count() -> None:
    lib().iadr1_reset_launch_count()


def _ptr(t):
    return None if t is None else t.data_ptr()


def _mat(t: torch.Tensor, what: str):
    """(ptr, ld, mn_major, rows, cols) for a 2-D bf16 view: rows x cols logical, either dim contiguous."""
    if t.dtype != torch.bfloat16:
        raise TypeError(f"{what} must be bfloat16, got {t.dtype}")
    if t.dim() != 2:
        raise ValueError(f"{what} must be 2-D")
    if t.stride(1) == 1:
        return t.data_ptr(), t.stride(0), 0
    if t.stride(0) == 1:
        return t.data_ptr(), t.stride(1), 1
    raise ValueError(f"{what}: one of the two dims must be contiguous (strides {t.stride()})")


def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor | None = None, *, bias=None, residual=None,
         alpha: float = 1.0, accumulate: bool = False, out_dtype=torch.bfloat16, split_k: int = 1,
         atomic: bool | None = None, trans_out: bool = False, bias_per_m: bool = False, block_n: int = 0, stages: int = 0,
         max_ctas: int = 0, a_static: bool = False, stream_k: bool = False,
         co_resident: bool = False, no_bulk_red: bool = False, raster: int = 0,
         no_chunked_maps: bool = False) -> torch.Tensor:
    """out[M,N] (+)= alpha * a[M,K] @ b[N,K]^T (+ bias) (+ residual).

    ``a`` and ``b`` are 2-D bf16 *views*; either of their dims may be the contiguous one, so ``x @ W.T`` is
    ``gemm(x, W)``, ``dy @ W`` is ``gemm(dy, W.t())`` and ``dy.T @ x`` is ``gemm(dy.t(), x.t())`` with no copies.
    """
    M, K = a.shape
    N, K2 = b.shape
    if K != K2:
        raise ValueError(f"inner dims differ: {a.shape} vs {b.shape}")
    if out is None:
        shape = (N, M) if trans_out else (M, N)
        out = torch.zeros(shape, dtype=out_dtype, device=a.device) if (accumulate or split_k > 1) else \
            torch.empty(shape, dtype=out_dtype, device=a.device)
    d = GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.batch = d.batch_lo = d.b_lo_div = 1
    d.A, d.lda, d.a_mn = _mat(a, "a")
    d.B, d.ldb, d.b_mn = _mat(b, "b")
    if out.stride(1) != 1:
        raise ValueError("out must be row-major")
    if tuple(out.shape) != ((N, M) if trans_out else (M, N)):
        raise ValueError(f"out has shape {tuple(out.shape)}, expected {(N, M) if trans_out else (M, N)}")
    d.C, d.ldc = out.data_ptr(), out.stride(0)
    d.c_f32 = 1 if out.dtype == torch.float32 else 0
    d.trans_c = int(trans_out)
    atomic = (split_k > 1) if atomic is None else atomic
    d.accumulate = int(accumulate and not atomic)
    d.atomic = int(atomic)
    d.split_k = split_k
    d.alpha = alpha
    d.bias, d.bias_per_m = _ptr(bias), int(bias_per_m)
    d.residual = _ptr(residual)
    d.block_n, d.stages, d.max_ctas, d.a_static = block_n, stages, max_ctas, int(a_static)
    d.stream_k, d.co_resident, d.no_bulk_red, d.raster = int(stream_k), int(co_resident), int(no_bulk_red), raster
    d.no_chunked_maps = int(no_chunked_maps)
    check(lib().iadr1_gemm_bf16(C.byref(d), stream_ptr()), "iadr1_gemm_bf16")
    return out


def gemm_swiglu(w_gate_up: torch.Tensor, x: torch.Tensor, out: torch.Tensor, *, block_n: int = 0, a_static: bool = True,
                co_resident: bool = True) -> torch.Tensor:
    """Decode MLP front half in one launch: out[R, I] = bf16(silu(x @ Wg^T)) * (x @ Wu^T) with w_gate_up = [Wg; Wu]
    ([2I, K] row-major, the training layout) as the 128-row MMA operand (64 gate + 64 up rows per tile)."""
    M2, K = w_gate_up.shape
    R, K2 = x.shape
    if K != K2 or tuple(out.shape) != (R, M2 // 2) or out.dtype != torch.bfloat16:
        raise ValueError("gemm_swiglu: shape/dtype mismatch")
    d = GemmDesc()
    d.M, d.N, d.K = M2, R, K
    d.batch = d.batch_lo = d.b_lo_div = 1
    d.A, d.lda, d.a_mn = _mat(w_gate_up, "w_gate_up")
    d.B, d.ldb, d.b_mn = _mat(x, "x")
    d.C, d.ldc = out.data_ptr(), out.stride(0)
    d.split_k, d.alpha, d.epi, d.up_row_off = 1, 1.0, 3, M2 // 2
    d.block_n, d.a_static, d.co_resident = block_n, int(a_static), int(co_resident)
    check(lib().iadr1_gemm_bf16(C.byref(d), stream_ptr()), "iadr1_gemm_bf16(swiglu)")
    return out


def gemm_swiglu_train(x: torch.Tensor, w_gate_up: torch.Tensor, act: torch.Tensor, gu: torch.Tensor | None = None) -> torch.Tensor:
    """Training MLP front half in one launch (epi 4): act[M, I] = bf16(silu(bf16(x @ Wg^T))) * bf16(x @ Wu^T) with the fused weight
    w_gate_up = [Wg; Wu] ([2I, K]); `gu` [M, 2I] optionally receives the gate | up pre-activations the backward needs."""
    M, K = x.shape
    I2, K2 = w_gate_up.shape
    if K != K2 or tuple(act.shape) != (M, I2 // 2) or act.dtype != torch.bfloat16:
        raise ValueError("gemm_swiglu_train: shape/dtype mismatch")
    d = GemmDesc()
    d.M, d.N, d.K = M, I2 // 2, K
    d.batch = d.batch_lo = d.b_lo_div = 1
    d.A, d.lda, d.a_mn = _mat(x, "x")
    d.B, d.ldb, d.b_mn = _mat(w_gate_up, "w_gate_up")
    d.C, d.ldc = act.data_ptr(), act.stride(0)
    d.split_k, d.alpha, d.epi = 1, 1.0, 4
    if gu is not None:
        d.gu_out, d.gu_ld = gu.data_ptr(), gu.stride(0)
    check(lib().iadr1_gemm_bf16(C.byref(d), stream_ptr()), "iadr1_gemm_bf16(swiglu, training)")
    return act


def gemm_swiglu_bwd(dy: torch.Tensor, w_down: torch.Tensor, gu: torch.Tensor) -> torch.Tensor:
    """Down-projection input gradient with the SwiGLU backward in its epilogue (epi 5): for dact = dy[M, H] @ w_down[H, I] (never
    written), gu[M, 2I] = gate | up is overwritten in place with dgate | dup - bit-identical to gemm + act_mul_bwd."""
    M, H = dy.shape
    H2, I = w_down.shape
    if H != H2 or gu.shape[0] != M or gu.shape[1] != 2 * I or gu.dtype != torch.bfloat16:
        raise ValueError("gemm_swiglu_bwd: shape/dtype mismatch")
    d = GemmDesc()
    d.M, d.N, d.K = M, I, H
    d.batch = d.batch_lo = d.b_lo_div = 1
    d.A, d.lda, d.a_mn = _mat(dy, "dy")
    d.B, d.ldb, d.b_mn = _mat(w_down.t(), "w_down")
    d.split_k, d.alpha, d.epi = 1, 1.0, 5
    d.gu_out, d.gu_ld = gu.data_ptr(), gu.stride(0)
    check(lib().iadr1_gemm_bf16(C.byref(d), stream_ptr()), "iadr1_gemm_bf16(swiglu backward)")
    return gu


def gemm_batched(a, b, out, *, M, N, K, batch, batch_lo=0, b_lo_div=1, lda, a_bs_lo=0, a_bs_hi=0, a_mn=0, ldb,
                 b_bs_lo=0, b_bs_hi=0, b_mn=0, ldc, c_bs_lo=0, c_bs_hi=0, alpha=1.0, kmode=0, skip_mode=0,
                 causal_off=0, block_n=0, a_off=0, b_off=0, c_off=0, residual=None):
    """Raw batched entry (element strides) used by the attention compositions. Offsets are in elements."""
    d = GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.batch, d.batch_lo, d.b_lo_div = batch, batch_lo or batch, b_lo_div
    d.A, d.lda, d.a_bs_lo, d.a_bs_hi, d.a_mn = a.data_ptr() + 2 * a_off, lda, a_bs_lo, a_bs_hi, a_mn
    d.B, d.ldb, d.b_bs_lo, d.b_bs_hi, d.b_mn = b.data_ptr() + 2 * b_off, ldb, b_bs_lo, b_bs_hi, b_mn
    esz = out.element_size()
    d.C, d.ldc, d.c_bs_lo, d.c_bs_hi = out.data_ptr() + esz * c_off, ldc, c_bs_lo, c_bs_hi
    d.c_f32 = 1 if out.dtype == torch.float32 else 0
    d.split_k = 1
    d.alpha = alpha
    d.kmode, d.skip_mode, d.causal_off = kmode, skip_mode, causal_off
    d.block_n = block_n
    if residual is not None:  # bf16, indexed exactly like C (same offset / strides)
        d.residual = residual.data_ptr() + 2 * c_off
    check(lib().iadr1_gemm_bf16(C.byref(d), stream_ptr()), "iadr1_gemm_bf16(batched)")
    return out
