// Debug twin of the tcgen05 GEMM on CUDA cores: same argument contract, one
// thread per output element, fp32 accumulation.  Used by the GPU tests to
// cross-check the tensor-core path (and its epilogues) on the device itself.
// Never called from the product path.
#include "common.cuh"
#include "mmdit_b200.h"

namespace mmdit {

struct SimtParams {
  const bf16 *A, *B;
  void* D;
  long long M, N, K, lda, ldb, ldd;
  int a_mn, b_mn, d_fp32, accumulate, epi;
  const void* bias;
  int bias_fp32;
  const bf16* gate;
  long long rows_per_gate, ld_gate;
  const bf16* resid;
  long long ldr;
  bf16* aux;
  long long ld_aux;
  long long remap_rows, remap_batch_rows, remap_offset;
};

__global__ void gemm_simt_kernel(SimtParams p) {
  const long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long m = blockIdx.y;
  if (n >= p.N || m >= p.M) return;
  float acc = 0.f;
  for (long long k = 0; k < p.K; ++k) {
    const float a = __bfloat162float(p.a_mn ? p.A[k * p.lda + m] : p.A[m * p.lda + k]);
    const float b = __bfloat162float(p.b_mn ? p.B[k * p.ldb + n] : p.B[n * p.ldb + k]);
    acc = fmaf(a, b, acc);
  }
  if (p.bias)
    acc += p.bias_fp32 ? reinterpret_cast<const float*>(p.bias)[n]
                       : __bfloat162float(reinterpret_cast<const bf16*>(p.bias)[n]);
  long long drow = m;
  if (p.remap_rows > 0)
    drow = (m / p.remap_rows) * p.remap_batch_rows + (m % p.remap_rows) + p.remap_offset;
  if (p.aux) p.aux[drow * p.ld_aux + n] = __float2bfloat16(acc);
  if (p.epi == MMDIT_EPI_GATE_RESID) {
    const float g = __bfloat162float(p.gate[(m / p.rows_per_gate) * p.ld_gate + n]);
    acc = fmaf(acc, g, __bfloat162float(p.resid[drow * p.ldr + n]));
  } else if (p.epi == MMDIT_EPI_RESID) {
    acc += __bfloat162float(p.resid[drow * p.ldr + n]);
  } else if (p.epi == MMDIT_EPI_SILU) {
    acc = silu_f(acc);
  }
  if (p.d_fp32) {
    float* d = reinterpret_cast<float*>(p.D) + drow * p.ldd + n;
    *d = p.accumulate ? *d + acc : acc;
  } else {
    reinterpret_cast<bf16*>(p.D)[drow * p.ldd + n] = __float2bfloat16(acc);
  }
}

}  // namespace mmdit

using namespace mmdit;

extern "C" int mmdit_gemm_bf16_simt(const mmdit_gemm_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MMDIT_REQUIRE(a && a->A && a->B && a->D, MMDIT_ERR_ARG, "gemm_simt: null pointer argument");
  MMDIT_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0 && a->M < 65536, MMDIT_ERR_ARG,
                "gemm_simt: bad shape");
  SimtParams p;
  p.A = static_cast<const bf16*>(a->A); p.B = static_cast<const bf16*>(a->B); p.D = a->D;
  p.M = a->M; p.N = a->N; p.K = a->K; p.lda = a->lda; p.ldb = a->ldb; p.ldd = a->ldd;
  p.a_mn = a->a_major; p.b_mn = a->b_major; p.d_fp32 = a->d_fp32; p.accumulate = a->accumulate;
  p.epi = a->epilogue; p.bias = a->bias; p.bias_fp32 = a->bias_fp32;
  p.gate = static_cast<const bf16*>(a->gate); p.rows_per_gate = a->rows_per_gate;
  p.ld_gate = a->ld_gate; p.resid = static_cast<const bf16*>(a->resid); p.ldr = a->ldr;
  p.aux = static_cast<bf16*>(a->aux); p.ld_aux = a->ld_aux;
  p.remap_rows = a->remap_rows; p.remap_batch_rows = a->remap_batch_rows;
  p.remap_offset = a->remap_offset;
  dim3 grid((unsigned)((a->N + 127) / 128), (unsigned)a->M);
  gemm_simt_kernel<<<grid, 128, 0, stream>>>(p);
  return check_launch("gemm_simt_kernel");
}
