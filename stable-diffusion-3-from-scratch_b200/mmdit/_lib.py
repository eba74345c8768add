"""ctypes binding of libmmdit_b200.so (the C-ABI declared in include/mmdit_b200.h).

PyTorch is imported first so that the library's libcudart.so.12 dependency
resolves to the runtime torch already loaded.  There is no fallback: if the
shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

import torch  # noqa: F401  (loads libcudart before our library)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libmmdit_b200.so")

c_i32, c_i64, c_f32, c_vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class GemmArgs(C.Structure):
    """Mirror of `mmdit_gemm_args` (include/mmdit_b200.h)."""

    _fields_ = [
        ("A", c_vp), ("B", c_vp), ("D", c_vp),
        ("M", c_i64), ("N", c_i64), ("K", c_i64),
        ("lda", c_i64), ("ldb", c_i64), ("ldd", c_i64),
        ("a_major", c_i32), ("b_major", c_i32),
        ("d_fp32", c_i32), ("accumulate", c_i32),
        ("split_k", c_i32), ("epilogue", c_i32),
        ("bias", c_vp), ("bias_fp32", c_i32),
        ("gate", c_vp), ("rows_per_gate", c_i64), ("ld_gate", c_i64),
        ("resid", c_vp), ("ldr", c_i64),
        ("aux", c_vp), ("ld_aux", c_i64),
        ("remap_rows", c_i64), ("remap_batch_rows", c_i64), ("remap_offset", c_i64),
        ("force_block_n", c_i32), ("reserved", c_i32),
        ("qk_wq", c_vp), ("qk_wk", c_vp), ("rope_cos", c_vp), ("rope_sin", c_vp),
        ("qk_tokens", c_i32), ("qk_eps", c_f32),
        ("colsum_partial", c_vp),
    ]


class AttnArgs(C.Structure):
    """Mirror of `mmdit_attn_args` (include/mmdit_b200.h)."""

    _fields_ = [
        ("q", c_vp * 2), ("k", c_vp * 2), ("v", c_vp * 2),
        ("ld_q", c_i64 * 2), ("ld_k", c_i64 * 2), ("ld_v", c_i64 * 2),
        ("o", c_vp * 2), ("ld_o", c_i64 * 2),
        ("lse", c_vp),
        ("B", c_i32), ("H", c_i32), ("N", c_i32), ("M", c_i32), ("head_dim", c_i32),
        ("scale", c_f32),
        ("d_o", c_vp * 2), ("ld_do", c_i64 * 2),
        ("dq", c_vp * 2), ("dk", c_vp * 2), ("dv", c_vp * 2),
        ("ld_dq", c_i64 * 2), ("ld_dk", c_i64 * 2), ("ld_dv", c_i64 * 2),
        ("delta", c_vp), ("dq_acc", c_vp),
        ("logit_bound", c_vp),
    ]


EPI_NONE, EPI_GATE_RESID, EPI_SILU, EPI_RESID, EPI_SWIGLU, EPI_QKNORM, EPI_SWIGLU_BWD = 0, 1, 2, 3, 4, 5, 6

_lib = None


def lib():
    """Load the shared library once; raise loudly when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or eager fallback for the MMDiT hot path)")
    L = C.CDLL(LIB_PATH)
    L.mmdit_last_error.restype = C.c_char_p
    L.mmdit_last_error.argtypes = []
    L.mmdit_abi_version.restype = c_i32
    L.mmdit_device_check.restype = c_i32
    L.mmdit_launch_count.restype = C.c_ulonglong
    _declare(L)
    _lib = L
    return L


def _declare(L):
    from . import _abi
    for name, argtypes in _abi.SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = _abi.RESTYPES.get(name, c_i32)
        fn.argtypes = argtypes


def check(rc, what):
    if rc != 0:
        msg = lib().mmdit_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return 0 if t is None else t.data_ptr()


def launch_count():
    """CUDA kernels launched by the library so far in this process."""
    return int(lib().mmdit_launch_count())
