/*
 * mmdit_b200.h -- C-ABI of the B200-native MMDiT hot path (libmmdit_b200.so).
 *
 * The reference (gmongaras/Stable-Diffusion-3-From-Scratch) has no FFI of its
 * own: its hot path is Python modules calling torch / flash-attn / xformers.
 * Each entry point below names the reference call site it replaces
 * (file:line under the reference root).  All pointers are DEVICE pointers
 * owned by the caller (PyTorch's allocator); the library never allocates,
 * frees or retains device memory, never synchronises the device, launches
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
  MMDIT_EPI_SWIGLU = 4     /* B rows interleaved per 64: [gate|up]; D[M,N/2] = silu(g)*u; aux = raw acc+bias */
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
  int32_t split_k;    /* 0 = auto (only >1 when accumulate=1) */
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
} mmdit_gemm_args;

int mmdit_gemm_bf16(const mmdit_gemm_args* args, void* stream);
/* Debug twin: same contract on CUDA cores (no tensor cores, no TMA). Used by
 * tests to cross-check the tcgen05 path on the device; never on the product path. */
int mmdit_gemm_bf16_simt(const mmdit_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMDIT_B200_H_ */
