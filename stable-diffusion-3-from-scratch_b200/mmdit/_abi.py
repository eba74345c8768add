"""Argument types of every compute entry point in include/mmdit_b200.h.

Kept as plain data so tests can check (without a GPU) that the built library
exports each symbol and that this table and the header agree.
"""
import ctypes as C

vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float

SIGNATURES = {
    # name: argtypes (all return int)
    "mmdit_gemm_bf16": [vp, vp],        # (const mmdit_gemm_args*, stream)
    "mmdit_gemm_bf16_simt": [vp, vp],
}
