/*
 * mmdit_b200.h -- C-ABI of the B200-native MMDiT hot path (libmmdit_b200.so).
 *
 * The reference (gmongaras/Stable-Diffusion-3-From-Scratch) has no FFI of its
 * own: its hot path is Python modules calling torch / flash-attn / xformers.
 * Each entry point below names the reference call site it replaces
 * (file:line under the reference root).  All pointers are DEVICE pointers
 * owned by the caller (PyTorch's allocator); the library never allocates,
 * frees or retains device memory (one exception: the IPC-exportable gradient
 * arena of mmdit_comm_alloc), never synchronises the device, launches
 * only on the stream passed in, and is CUDA-graph capturable.
 *
 * Conventions
 *   - activations are bf16 row-major unless a parameter says otherwise
 *   - return value: 0 = ok, < 0 = argument error, > 0 = cudaError_t;
 *     mmdit_last_error() returns a thread-local message for the last failure
 *   - `stream` is a cudaStream_t passed as void*
 */
#ifndef MMDIT_B200_H_
#define MMDIT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMDIT_ABI_VERSION 1

const char* mmdit_last_error(void);
int mmdit_abi_version(void);
/* 0 when the current CUDA device is sm_100 (B200); error otherwise. */
int mmdit_device_check(void);
/* Number of CUDA kernels this library has launched in this process (bench.py's gpu_launches). */
unsigned long long mmdit_launch_count(void);

/* ------------------------------------------------------------------ GEMM --
 * D[M,N] = epilogue( A[M,K] * B[N,K]^T )  on tcgen05 tensor cores, fp32
 * accumulation in TMEM, operands staged by TMA.
 *
 * Replaces every nn.Linear / Conv2d-as-GEMM on the path and their autograd
 * dgrad / wgrad: Attention.py:36-45,130-135,425 (QKV / out projections),
 * MLP.py:19,32 (SwiGLU w12 / w3), Norm.py:13-14, Transformer_Block_Dual.py:
 * 25-28,49-53 (adaLN modulation), diff_model.py:157,160,168-169,176,210
 * (front/back-end linears), ImagePositionalEncoding.py:114-116 (patch conv).
 *
 * a_major / b_major: 0 = reduction dim contiguous (A stored [M,K], B stored
 * [N,K]); 1 = M/N contiguous (A stored [K,M], B stored [K,N]).  That covers
 * fprop (0,0), dgrad dX = dY * W (0,1) and wgrad dW = dY^T * X (1,1) with no
 * transposed copies.
 */
enum {
  MMDIT_EPI_NONE = 0,      /* D = acc (+bias)                                   */
  MMDIT_EPI_GATE_RESID = 1,/* D = (acc+bias) * gate[m/rows_per_gate, n] + resid */
  MMDIT_EPI_SILU = 2,      /* D = silu(acc+bias)                                */
  MMDIT_EPI_RESID = 3,     /* D = acc + bias + resid                            */
  MMDIT_EPI_SWIGLU = 4,    /* B = [gate rows; up rows] (xformers w12, MLP.py:19): D[M,N/2] = silu(g)*u,
                              aux[M,N] = acc+bias (bf16; NULL = not kept, inference); needs N % 256 == 0, M > 128 */
  MMDIT_EPI_QKNORM = 5,    /* B = [q; k; v] rows (N = 3d): D[M,N] = acc (+bias), aux[M,2d] = per-head
                              RMSNorm * weight (+ 2-D RoPE) of the q and k columns (Attention.py:61-64,
                              130-134,174-194); validated, measured slower than the separate kernel */
  MMDIT_EPI_SWIGLU_BWD = 6 /* data gradient of xformers' w3 fused with the SwiGLU backward (MLP.py:19,32):
                              g = bf16(acc) [M,N = hidden]; aux[M,2N] = [x1 | x2] (INPUT, the saved
                              pre-activations); D[M,2N] = [g x2 silu'(x1) | g silu(x1)];
                              colsum_partial (optional) = column sums of every 32-row strip of D.
                              Needs N % 256 == 0, M % 128 == 0, M > 128, no bias */
};

typedef struct mmdit_gemm_args {
  const void* A;      /* bf16 */
  const void* B;      /* bf16 */
  void* D;            /* bf16 or fp32 */
  int64_t M, N, K;
  int64_t lda, ldb, ldd; /* leading dimensions in elements */
  int32_t a_major, b_major;
  int32_t d_fp32;     /* 0: bf16 output, 1: fp32 output */
  int32_t accumulate; /* fp32 output only: D += result (atomic when split_k > 1) */
  int32_t split_k;    /* 0 = auto (only >1 when accumulate=1); >1 with accumulate=0: slices mode,
                         split s writes D + s*M*ldd and the caller folds (mmdit_fold_slices_f32) */
  int32_t epilogue;   /* MMDIT_EPI_* */
  const void* bias;   /* [N] or NULL */
  int32_t bias_fp32;  /* dtype of bias */
  const void* gate;   /* bf16 [M/rows_per_gate, N] (row stride ld_gate) */
  int64_t rows_per_gate;
  int64_t ld_gate;
  const void* resid;  /* bf16 [M,N] row stride ldr (may alias D) */
  int64_t ldr;
  void* aux;          /* optional bf16 [M,N]: pre-epilogue value (acc+bias), for backward */
  int64_t ld_aux;
  /* optional output-row remap (scatter rows of a per-batch sub-sequence):
   * d_row = (m / remap_rows) * remap_batch_rows + (m % remap_rows) + remap_offset; 0 = identity */
  int64_t remap_rows, remap_batch_rows, remap_offset;
  int32_t force_block_n; /* 0 = auto; else 64/128/256 (testing) */
  int32_t reserved;
  /* MMDIT_EPI_QKNORM only */
  const void* qk_wq;     /* fp32 [64] */
  const void* qk_wk;     /* fp32 [64] */
  const void* rope_cos;  /* fp32 [tokens, 32] or NULL (no rotation: text stream) */
  const void* rope_sin;
  int32_t qk_tokens;     /* tokens per sample: position of row m is m % qk_tokens */
  float qk_eps;
  /* MMDIT_EPI_SWIGLU_BWD only */
  float* colsum_partial; /* fp32 [M/32, 2N] or NULL */
} mmdit_gemm_args;

int mmdit_gemm_bf16(const mmdit_gemm_args* args, void* stream);
/* Debug twin: same contract on CUDA cores (no tensor cores, no TMA). Used by
 * tests to cross-check the tcgen05 path on the device; never on the product path. */
int mmdit_gemm_bf16_simt(const mmdit_gemm_args* args, void* stream);


/* ------------------------------------------------------------- attention --
 * Joint text+image softmax attention, head_dim 64, non-causal, fp32 softmax.
 * Replaces flash_attn_func and the concat/transpose/split copies around it
 * (Attention.py:259-263, 293, 411-417) and their autograd.
 * Streams: [0] = image tokens (N per sample), [1] = text tokens (M per sample);
 * the joint sequence order is image first, then text (Attention.py:259-261).
 * Every operand is a [B*rows, ld] bf16 matrix in which head h occupies columns
 * [h*64, h*64+64) from the given base pointer, so q/k/v may live inside packed
 * QKV buffers.  lse / delta: fp32 [B, H, N+M]; dq_acc: fp32 [B, N+M, H*64].
 */
typedef struct mmdit_attn_args {
  const void* q[2]; const void* k[2]; const void* v[2];
  int64_t ld_q[2], ld_k[2], ld_v[2];
  void* o[2];            /* fwd: output; bwd: forward output (input) */
  int64_t ld_o[2];
  float* lse;            /* fwd: output; bwd: input */
  int32_t B, H, N, M, head_dim;
  float scale;
  /* backward only */
  const void* d_o[2]; int64_t ld_do[2];
  void* dq[2]; void* dk[2]; void* dv[2];   /* dq[s] may be NULL: dq then stays in dq_acc (fp32) for the caller */
  int64_t ld_dq[2], ld_dk[2], ld_dv[2];
  float* delta;          /* workspace */
  float* dq_acc;         /* workspace; holds dq (fp32, joint layout) on return */
  /* forward only, optional: device scalar holding an upper bound of |scale * q.k| (see
   * mmdit_qk_logit_bound).  When present and <= 24 the softmax runs in one pass against this
   * fixed reference (no running maximum, no rescaling); otherwise the online-softmax path runs. */
  const float* logit_bound;
} mmdit_attn_args;

int mmdit_attn_fwd(const mmdit_attn_args* args, void* stream);
int mmdit_attn_bwd(const mmdit_attn_args* args, void* stream);

/* ------------------------------------------------- adaLN LayerNorm-modulate --
 * y = LN(x) * (1 + scale[b]) + shift[b], LN eps, no affine (Norm.py:16-23).
 * shift/scale: bf16 [B, d] with row stride ld_mod; b = row / rows_per_batch.
 * bwd: dx = LN-backward(dy * (1+scale)) (+ dres), dshift/dscale (fp32, row
 * stride ld_dmod; bf16 when dmod_bf16) are WRITTEN (two-stage reduction through `workspace`). */
int mmdit_ln_modulate_fwd(const void* x, const void* shift, const void* scale, void* y, float* mean,
                          float* rstd, int64_t rows, int32_t d, int64_t rows_per_batch,
                          int64_t ld_mod, float eps, void* stream);
int mmdit_ln_modulate_bwd(const void* dy, const void* x, const float* mean, const float* rstd,
                          const void* scale, const void* dres, void* dx, void* dshift,
                          void* dscale, int32_t dmod_bf16, float* workspace, int64_t rows, int32_t d,
                          int64_t rows_per_batch, int64_t ld_mod, int64_t ld_dmod, void* stream);
/* fp32 elements of `workspace` needed by ln_modulate_bwd / gate_bwd (per-block partial sums;
 * the column reductions are two-stage, no atomics). */
int64_t mmdit_rowreduce_workspace_floats(int64_t rows, int32_t d, int64_t rows_per_batch);

/* Backward of the gated residual o = a * gate[b] + x (Transformer_Block_Dual.py:64-76):
 * da = dout * gate[b]; dgate[b] += sum_rows dout * a; dab[b] += sum_rows da (optional,
 * per-sample partial of the bias gradient of the producing linear). fp32 outputs are written;
 * workspace: mmdit_rowreduce_workspace_floats(rows, d, rows_per_batch) floats. */
/* Forward of the same gated residual as a standalone kernel: out = a * gate[b] + resid. */
int mmdit_gate_residual_fwd(const void* a, const void* gate, const void* resid, void* out,
                            int64_t rows, int32_t d, int64_t rows_per_batch, int64_t ld_gate,
                            void* stream);
/* The gated residual followed by the next adaLN LayerNorm-modulate in ONE pass
 * (Transformer_Block_Dual.py:64-72: `X = attn * scale1(y) + X` then `norm2(X, y)`; same for the text
 * stream): x_out = a * gate[b] + resid (bf16, bit-identical to mmdit_gate_residual_fwd),
 * y = LN(x_out) * (1 + scale[b]) + shift[b] (bit-identical to mmdit_ln_modulate_fwd on x_out),
 * mean / rstd of x_out for the backward. */
int mmdit_gate_residual_ln_fwd(const void* a, const void* gate, const void* resid, const void* shift,
                               const void* scale, void* x_out, void* y, float* mean, float* rstd,
                               int64_t rows, int32_t d, int64_t rows_per_batch, int64_t ld_gate,
                               int64_t ld_mod, float eps, void* stream);
int mmdit_gate_bwd(const void* dout, const void* a, const void* gate, void* da, void* dgate,
                   int32_t dgate_bf16, float* dab, float* workspace, int64_t rows, int32_t d,
                   int64_t rows_per_batch, int64_t ld_gate, int64_t ld_dgate, int64_t ld_dab,
                   void* stream);
/* Which generation of the row kernels (LayerNorm-modulate fwd / bwd, gated residual + LN, gate bwd,
 * QK-RMSNorm + RoPE fwd / bwd) the entry points above and below launch: 2 (default; csrc/rowwise2.cu,
 * csrc/qknorm2.cu) or 1 (csrc/rowwise.cu, csrc/elementwise.cu; also what shapes the second generation
 * does not cover fall back to).  Same results (forward kernels bit-identical); the environment variable
 * MMDIT_ROW_KERNELS=1 sets the initial value.  Not a reference interface: an A/B and regression knob. */
int mmdit_set_row_kernel_generation(int32_t generation);

/* Text front-end (diff_model.py:168-172,323-326): out = bf16(sigma * RMSNorm_fp32(c) * w).
 * Tokens [0,split) of each sample use (w1,sigma1) -> out1 [B*split, d]; tokens [split,M)
 * use (w2,sigma2) -> out2 [B*(M-split), d]. rstd: fp32 [B*M] saved for backward. */
int mmdit_text_norm_fwd(const void* c, const float* w1, const float* w2, const float* sigma1,
                        const float* sigma2, void* out1, void* out2, float* rstd, int64_t batch,
                        int32_t tokens, int32_t split, int32_t d, float eps, void* stream);
/* One half: dn = grad wrt out half [B*ntok, d]; accumulates dw [d] and dsigma [1]. */
int mmdit_text_norm_bwd(const void* dn, int32_t dn_fp32, const void* c, const float* rstd,
                        const float* w, const float* sigma, float* dw, float* dsigma, int64_t batch,
                        int32_t tokens, int32_t tok0, int32_t ntok, int32_t d, void* stream);

/* Per-head QK RMSNorm (+ 2-D axial RoPE on image tokens) (Attention.py:61-64,130-134,
 * 174-194; rotary_embedding.py:36-76,269-288).  qkv: raw projections, q at column 0 and
 * k at column d of each row (stride ld_in).  out: q at 0, k at d (stride ld_out).
 * rope_cos/rope_sin: fp32 [tokens_per_sample, 32] or NULL (text stream). */
int mmdit_qknorm_rope_fwd(const void* qkv, const float* wq, const float* wk, const float* rope_cos,
                          const float* rope_sin, void* out, int64_t rows, int32_t d, int64_t ld_in,
                          int64_t ld_out, int32_t tokens_per_sample, float eps, void* stream);
/* out[0] = 1.02 * scale * 64 * max|w_q| * max|w_k| over both streams: bound of the scaled logits
 * that per-head RMSNorm guarantees (head_dim 64; RoPE is a rotation). wq_c / wk_c may be NULL. */
int mmdit_qk_logit_bound(const float* wq_x, const float* wk_x, const float* wq_c, const float* wk_c,
                         float scale, float* out, void* stream);
int mmdit_qknorm_rope_bwd(const void* dqk, const void* qkv, const float* wq, const float* wk,
                          const float* rope_cos, const float* rope_sin, void* dqkv, float* dwq,
                          float* dwk, int64_t rows, int32_t d, int64_t ld_g, int64_t ld_in,
                          int64_t ld_dout, int32_t tokens_per_sample, float eps, void* stream);
/* Same, with the q half of the incoming gradient read from the attention backward's fp32 accumulator
 * (mmdit_attn_args.dq_acc, [B, acc_tokens, d]; this stream's rows start at acc_tok_off inside every sample)
 * and rounded to bf16 on the fly: call mmdit_attn_bwd with dq[s] = NULL and no bf16 copy of dq is ever
 * written.  Only the k half (columns d .. 2d) of dqk is read. */
int mmdit_qknorm_rope_bwd_acc(const float* dq_acc, int32_t acc_tokens, int32_t acc_tok_off, const void* dqk,
                              const void* qkv, const float* wq, const float* wk, const float* rope_cos,
                              const float* rope_sin, void* dqkv, float* dwq, float* dwk, int64_t rows,
                              int32_t d, int64_t ld_g, int64_t ld_in, int64_t ld_dout,
                              int32_t tokens_per_sample, float eps, void* stream);

/* SwiGLU activation, xformers semantics (MLP.py:19,32): h12 = [x1 | x2], a = silu(x1) * x2.
 * bwd writes dh12 and adds the column sums to db12 [2*hidden] (may be NULL; else needs workspace). */
int mmdit_swiglu_fwd(const void* h12, void* a, int64_t rows, int32_t hidden, void* stream);
int mmdit_swiglu_bwd(const void* da, const void* h12, void* dh12, float* db12, float* workspace,
                     int64_t rows, int32_t hidden, void* stream);
int64_t mmdit_swiglu_bwd_workspace_floats(int64_t rows, int32_t hidden);

/* Timestep embedding (PositionalEncoding.py:23-30 with t * time_scale, diff_model.py:306):
 * out bf16 [B, d]; denom fp32 [d] as built by the reference ctor. bwd accumulates dscale[1]. */
int mmdit_timestep_embed_fwd(const float* t, const float* time_scale, const float* denom, void* out,
                             int32_t batch, int32_t d, void* stream);
int mmdit_timestep_embed_bwd(const void* de, const float* t, const float* time_scale,
                             const float* denom, float* dscale, int32_t batch, int32_t d,
                             void* stream);

/* [B,C,H,W] <-> tokens [B*(H/p)*(W/p), C*p*p] bf16, column c*p*p + i*p + j
 * (ImagePositionalEncoding.py:114-116,181-183 conv-as-GEMM input; patchify.py:41-72). */
int mmdit_patchify(const void* img, int32_t img_fp32, void* tokens, int32_t B, int32_t C, int32_t H,
                   int32_t W, int32_t p, void* stream);
int mmdit_unpatchify(const void* tokens, void* img, int32_t img_fp32, int32_t B, int32_t C,
                     int32_t H, int32_t W, int32_t p, void* stream);

/* Rectified flow: x_t = (1-t) x0 + t eps (diff_model.py:229-241); loss = mean((v-(eps-x0))^2)
 * with diff kept for backward (model_trainer.py:429-446); dv = upstream * 2/numel * diff;
 * x -= ((1+w) v[:B] - w v[B:]) * dt (diff_model.py:419-429). */
int mmdit_rf_noise(const void* x0, const void* eps, int32_t in_fp32, const float* t, float* xt,
                   int64_t batch, int64_t per_sample, void* stream);
int mmdit_rf_loss_fwd(const void* v, int32_t v_fp32, const void* eps, const void* x0,
                      int32_t in_fp32, float* diff, float* loss, int64_t numel, void* stream);
int mmdit_rf_loss_bwd(const float* diff, const float* upstream, void* dv, int32_t dv_fp32,
                      int64_t numel, void* stream);
int mmdit_cfg_euler_step(float* x, const void* v, int32_t v_fp32, int64_t half_numel,
                         float cfg_scale, float dt, void* stream);

/* out[n] += sum_rows in[row, n] (bias gradients); fold of per-sample fp32 partials; cast. */
int mmdit_colsum_bf16(const void* in, float* out, int64_t rows, int32_t n, int64_t ld, void* stream);
int mmdit_fold_rows_f32(const float* in, float* out, int32_t rows, int32_t n, int64_t ld,
                        void* stream);
int mmdit_cast_f32_bf16(const float* in, void* out, int64_t n, void* stream);
/* out[i] (+)= sum_s ws[s*stride + i]: folds the slices of a split-K GEMM launched with
 * split_k > 1, accumulate = 0 (each split then writes its partial product to D + s*M*ldd). */
int mmdit_fold_slices_f32(const float* ws, float* out, int64_t n, int32_t slices, int64_t stride,
                          int32_t accumulate, void* stream);

/* ---------------------------------------------------------- optimizer step --
 * Fused gradient-norm clip + AdamW + bf16 shadow refresh over many tensors
 * (model_trainer.py:483-503; "next" row (f)-1 of the scope table).  table: device array of
 * mmdit_param_desc; chunks: device array of int32 pairs (tensor index, chunk index) with
 * mmdit_adamw_chunk_elems() elements per chunk; state: device float[4] = {grad sum of squares,
 * step count, learning rate, reserved}, owned by the caller and persistent across steps (the step
 * count is incremented on the device, so the call is CUDA-graph replayable).  lr < 0 means "read
 * the learning rate from state[2]": a warm-up / decay scheduler (model_trainer.py:25-41,
 * 263, 495-496) then only rewrites that float between replays of a captured step.
 * max_norm <= 0 disables clipping. */
typedef struct mmdit_param_desc {
  float* p; const float* g; float* m; float* v;
  void* shadow;          /* bf16 copy of p refreshed in the same pass, or NULL */
  int64_t n;
} mmdit_param_desc;
int mmdit_adamw_step(const void* table, const void* chunks, int32_t n_chunks, float* state,
                     float lr, float beta1, float beta2, float eps, float weight_decay,
                     float max_norm, void* stream);
int mmdit_adamw_chunk_elems(void);

/* ------------------------------------------- data-parallel gradient exchange --
 * Replaces the DDP reducer's bucketed NCCL all-reduce (model_trainer.py:224) with one kernel
 * per bucket over NVLink peer memory: reduce-scatter + all-gather + mean in a single launch,
 * CUDA-graph capturable, meant to run on a side stream while the backward continues.
 * The gradient arena and the signal pad live in ONE allocation per rank made by
 * mmdit_comm_alloc (cudaMalloc; the only device memory this library ever allocates -- it must be
 * exportable through CUDA IPC, which PyTorch's sub-allocated blocks are not); peers map it with
 * mmdit_comm_export / mmdit_comm_import (cudaIpc* handles, exchanged by the host over
 * torch.distributed).  world in {1, 2, 4, 8}.  Summation order is rank 0..W-1 on every rank:
 * results are bit-identical across ranks and across runs. */
typedef struct mmdit_comm {
  void* buf[8];    /* fp32 gradient arena of every rank (buf[rank] is the local one)        */
  void* flag[8];   /* signal pad of every rank: uint32[16], zero-initialised                */
  void* state;     /* local uint32[2] (epoch, finished-CTA counter), zero-initialised        */
  int32_t world, rank;
} mmdit_comm;
int mmdit_comm_alloc(void** ptr, int64_t bytes);          /* cudaMalloc + zero */
int mmdit_comm_free(void* ptr);
int mmdit_comm_handle_bytes(void);                          /* sizeof(cudaIpcMemHandle_t) */
int mmdit_comm_export(void* ptr, void* handle_out);
int mmdit_comm_import(const void* handle, void** peer_ptr_out);
int mmdit_comm_close(void* peer_ptr);
/* arena[offset, offset+n) <- mean over ranks; offset, n in floats, multiples of 4.
 * ctas <= 0: default grid.  Every rank must issue the same sequence of calls. */
int mmdit_allreduce_mean_f32(const mmdit_comm* comm, int64_t offset, int64_t n, int32_t ctas,
                             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMDIT_B200_H_ */
