"""Argument types of every compute entry point in include/mmdit_b200.h.

Kept as plain data so tests can check (without a GPU) that the built library
exports each symbol and that this table and the header agree.
"""
import ctypes as C

vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float

SIGNATURES = {
    # name: argtypes (all return int); vp = device pointer / struct pointer / stream
    "mmdit_gemm_bf16": [vp, vp],
    "mmdit_gemm_bf16_simt": [vp, vp],
    "mmdit_attn_fwd": [vp, vp],
    "mmdit_attn_bwd": [vp, vp],
    "mmdit_ln_modulate_fwd": [vp, vp, vp, vp, vp, vp, i64, i32, i64, i64, f32, vp],
    "mmdit_ln_modulate_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, i64, i32, i64, i64, i64, vp],
    "mmdit_gate_residual_fwd": [vp, vp, vp, vp, i64, i32, i64, i64, vp],
    "mmdit_gate_residual_ln_fwd": [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i64, i64, i64, f32, vp],
    "mmdit_gate_bwd": [vp, vp, vp, vp, vp, i32, vp, vp, i64, i32, i64, i64, i64, i64, vp],
    "mmdit_text_norm_fwd": [vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, f32, vp],
    "mmdit_text_norm_bwd": [vp, i32, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, vp],
    "mmdit_qknorm_rope_fwd": [vp, vp, vp, vp, vp, vp, i64, i32, i64, i64, i32, f32, vp],
    "mmdit_qk_logit_bound": [vp, vp, vp, vp, f32, vp, vp],
    "mmdit_qknorm_rope_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i64, i64, i64, i32, f32, vp],
    "mmdit_qknorm_rope_bwd_acc": [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i64, i64, i64, i32,
                                  f32, vp],
    "mmdit_swiglu_fwd": [vp, vp, i64, i32, vp],
    "mmdit_swiglu_bwd": [vp, vp, vp, vp, vp, i64, i32, vp],
    "mmdit_timestep_embed_fwd": [vp, vp, vp, vp, i32, i32, vp],
    "mmdit_timestep_embed_bwd": [vp, vp, vp, vp, vp, i32, i32, vp],
    "mmdit_patchify": [vp, i32, vp, i32, i32, i32, i32, i32, vp],
    "mmdit_unpatchify": [vp, vp, i32, i32, i32, i32, i32, i32, vp],
    "mmdit_rf_noise": [vp, vp, i32, vp, vp, i64, i64, vp],
    "mmdit_rf_loss_fwd": [vp, i32, vp, vp, i32, vp, vp, i64, vp],
    "mmdit_rf_loss_bwd": [vp, vp, vp, i32, i64, vp],
    "mmdit_cfg_euler_step": [vp, vp, i32, i64, f32, f32, vp],
    "mmdit_colsum_bf16": [vp, vp, i64, i32, i64, vp],
    "mmdit_fold_rows_f32": [vp, vp, i32, i32, i64, vp],
    "mmdit_cast_f32_bf16": [vp, vp, i64, vp],
    "mmdit_fold_slices_f32": [vp, vp, i64, i32, i64, i32, vp],
    "mmdit_adamw_step": [vp, vp, i32, vp, f32, f32, f32, f32, f32, f32, vp],
    "mmdit_adamw_chunk_elems": [],
    "mmdit_comm_alloc": [vp, i64],
    "mmdit_comm_free": [vp],
    "mmdit_comm_handle_bytes": [],
    "mmdit_comm_export": [vp, vp],
    "mmdit_comm_import": [vp, vp],
    "mmdit_comm_close": [vp],
    "mmdit_allreduce_mean_f32": [vp, i64, i64, i32, vp],
    "mmdit_rowreduce_workspace_floats": [i64, i32, i64],
    "mmdit_set_row_kernel_generation": [i32],
    "mmdit_swiglu_bwd_workspace_floats": [i64, i32],
}

# entry points that do not return the int status code
RESTYPES = {
    "mmdit_rowreduce_workspace_floats": i64,
    "mmdit_swiglu_bwd_workspace_floats": i64,
}
