// Memory-bound row kernels of the MMDiT block (bf16 activations, fp32 math):
//   adaLN LayerNorm-modulate fwd/bwd      (reference Norm.py:16-23)
//   gate backward                          (Transformer_Block_Dual.py:64-76, autograd)
//   text RMSNorm * scalar fwd/bwd          (diff_model.py:168-172,323-326)
// One warp owns one row; every lane moves 128-bit vectors (8 bf16) of
// consecutive 256-element chunks, so a warp reads/writes 512 contiguous bytes
// per instruction.  Row statistics use warp shuffles.  Column reductions
// (dshift/dscale/dgate, per batch element) are accumulated in registers across
// the rows a warp owns, folded once per block in shared memory and added to
// the fp32 output with one atomic per column per block.
#include <stdlib.h>

#include "common.cuh"
#include "mmdit_b200.h"

namespace mmdit {

constexpr int ROW_THREADS = 256;
constexpr int ROW_WARPS = ROW_THREADS / 32;

// Column reductions of the backward kernels: no atomics.  Each warp keeps per-lane column
// partials in registers; the block folds its warps through shared memory and writes one
// [2][d] partial; a second tiny kernel folds the partials of each sample.
constexpr int BWD_THREADS = 128;
constexpr int BWD_WARPS = BWD_THREADS / 32;

template <int NC>
__device__ __forceinline__ void block_fold_partials(float* red, const float (&a0)[NC][8],
                                                    const float (&a1)[NC][8], float* out, int d,
                                                    int warp, int lane) {
  float* mine = red + (long long)warp * 2 * d;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = c * 256 + lane * 8;
    if (col < d) {
      *reinterpret_cast<float4*>(mine + col) = make_float4(a0[c][0], a0[c][1], a0[c][2], a0[c][3]);
      *reinterpret_cast<float4*>(mine + col + 4) = make_float4(a0[c][4], a0[c][5], a0[c][6], a0[c][7]);
      *reinterpret_cast<float4*>(mine + d + col) = make_float4(a1[c][0], a1[c][1], a1[c][2], a1[c][3]);
      *reinterpret_cast<float4*>(mine + d + col + 4) = make_float4(a1[c][4], a1[c][5], a1[c][6], a1[c][7]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * d; i += BWD_THREADS) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < BWD_WARPS; ++w) acc += red[(long long)w * 2 * d + i];
    out[i] = acc;
  }
}

// out0[b, col] = sum_k partial[b, k, 0, col]; out1[b, col] = sum_k partial[b, k, 1, col]
// Outputs are fp32, or bf16 when out_bf16 is set (gradients of the bf16 modulation vectors).
__global__ void fold_batch_partials_kernel(const float* __restrict__ partial, void* __restrict__ out0,
                                           void* __restrict__ out1, int d, int blocks_per_batch,
                                           long long ld0, long long ld1, int out0_bf16, int out1_bf16) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // 0 .. 2d
  if (i >= 2 * d) return;
  const long long b = blockIdx.y;
  const float* src = partial + b * blocks_per_batch * 2 * d + i;
  float acc = 0.f;
  for (int k = 0; k < blocks_per_batch; ++k) acc += src[(long long)k * 2 * d];
  if (i < d) {
    if (out0_bf16) reinterpret_cast<bf16*>(out0)[b * ld0 + i] = __float2bfloat16(acc);
    else reinterpret_cast<float*>(out0)[b * ld0 + i] = acc;
  } else if (out1) {
    if (out1_bf16) reinterpret_cast<bf16*>(out1)[b * ld1 + (i - d)] = __float2bfloat16(acc);
    else reinterpret_cast<float*>(out1)[b * ld1 + (i - d)] = acc;
  }
}

template <int NC>
__device__ __forceinline__ void load_row(const bf16* p, int d, int lane, float (&v)[NC][8]) {
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = c * 256 + lane * 8;
    if (col < d) {
      load8(p + col, v[c]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[c][j] = 0.f;
    }
  }
}

// ------------------------------------------------------------ LN-modulate fwd
// y = LN(x) * bf16(1 + scale[b]) + shift[b]      (the reference adds 1 in bf16)
template <int NC>
__global__ void __launch_bounds__(ROW_THREADS)
ln_mod_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ shift,
                  const bf16* __restrict__ scale, bf16* __restrict__ y, float* __restrict__ mean_out,
                  float* __restrict__ rstd_out, long long R, int d, long long rows_per_batch,
                  long long ld_mod, float eps) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * ROW_WARPS + warp;
  if (row >= R) return;
  float v[NC][8];
  load_row<NC>(x + row * d, d, lane, v);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[c][j];
  const float mean = warp_sum(s) / d;
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = c * 256 + lane * 8;
    if (col < d) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = v[c][j] - mean;
        q += t * t;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / d + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  const long long b = row / rows_per_batch;
  const bf16* sh = shift + b * ld_mod;
  const bf16* sc = scale + b * ld_mod;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = c * 256 + lane * 8;
    if (col < d) {
      float fs[8], fc[8], o[8];
      load8(sh + col, fs);
      load8(sc + col, fc);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float one_plus = __bfloat162float(__float2bfloat16(1.f + fc[j]));
        o[j] = (v[c][j] - mean) * rstd * one_plus + fs[j];
      }
      store8(y + row * d + col, o);
    }
  }
}

// ------------------------------------------- gated residual + LN-modulate, one pass
// The two always follow each other inside a block (Transformer_Block_Dual.py:64-72):
//   x' = a * gate[b] + resid            (rounded to bf16, exactly like gate_residual_fwd_kernel)
//   y  = LN(x') * bf16(1 + scale[b]) + shift[b]
// One warp per row: a, resid read once, x' and y written once -- the separate kernels re-read x'.
template <int NC>
__global__ void __launch_bounds__(ROW_THREADS)
gate_res_ln_fwd_kernel(const bf16* __restrict__ a, const bf16* __restrict__ gate,
                       const bf16* __restrict__ resid, const bf16* __restrict__ shift,
                       const bf16* __restrict__ scale, bf16* __restrict__ xo, bf16* __restrict__ y,
                       float* __restrict__ mean_out, float* __restrict__ rstd_out, long long R, int d,
                       long long rows_per_batch, long long ld_gate, long long ld_mod, float eps) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * ROW_WARPS + warp;
  if (row >= R) return;
  const long long b = row / rows_per_batch;
  float v[NC][8], r[NC][8], g[NC][8];
  load_row<NC>(a + row * d, d, lane, v);
  load_row<NC>(resid + row * d, d, lane, r);
  load_row<NC>(gate + b * ld_gate, d, lane, g);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = c * 256 + lane * 8;
    if (col < d) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[c][j] = fmaf(v[c][j], g[c][j], r[c][j]);
      store8(xo + row * d + col, v[c]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[c][j] = __bfloat162float(__float2bfloat16(v[c][j]));   // what the LN of the stored x' sees
        s += v[c][j];
      }
    }
  }
  const float mean = warp_sum(s) / d;
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = c * 256 + lane * 8;
    if (col < d) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = v[c][j] - mean;
        q += t * t;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / d + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  const bf16* sh = shift + b * ld_mod;
  const bf16* sc = scale + b * ld_mod;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = c * 256 + lane * 8;
    if (col < d) {
      float fs[8], fc[8], o[8];
      load8(sh + col, fs);
      load8(sc + col, fc);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float one_plus = __bfloat162float(__float2bfloat16(1.f + fc[j]));
        o[j] = (v[c][j] - mean) * rstd * one_plus + fs[j];
      }
      store8(y + row * d + col, o);
    }
  }
}

// ------------------------------------------------------------ LN-modulate bwd
// g = dy*(1+s); dx = rstd*(g - mean(g) - xhat*mean(g*xhat)) (+ dres)
// dshift[b] += sum_rows dy ; dscale[b] += sum_rows dy*xhat      (fp32 atomics)
template <int NC>
__global__ void __launch_bounds__(BWD_THREADS, NC <= 3 ? 4 : 2)
ln_mod_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                  const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                  const bf16* __restrict__ scale, const bf16* __restrict__ dres,
                  bf16* __restrict__ dx, float* __restrict__ partial,
                  int d, long long rows_per_batch, long long ld_mod,
                  int rows_per_block, int blocks_per_batch) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  extern __shared__ float red[];  // [BWD_WARPS][2][d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = blockIdx.x / blocks_per_batch;
  const int chunk = blockIdx.x % blocks_per_batch;
  const long long r0 = b * rows_per_batch + (long long)chunk * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > (b + 1) * rows_per_batch) r1 = (b + 1) * rows_per_batch;


  const bf16* sc = scale + b * ld_mod;
  float a_sh[NC][8], a_sc[NC][8];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) a_sh[c][j] = a_sc[c][j] = 0.f;

  for (long long row = r0 + warp; row < r1; row += BWD_WARPS) {
    float g[NC][8], xh[NC][8];
    uint4 rres[NC];   // residual-path gradient, fetched together with dy and x (one latency, not two)
    load_row<NC>(dy + row * d, d, lane, g);
    load_row<NC>(x + row * d, d, lane, xh);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int col = c * 256 + lane * 8;
      rres[c] = (dres && col < d) ? *reinterpret_cast<const uint4*>(dres + row * d + col)
                                  : make_uint4(0u, 0u, 0u, 0u);
    }
    const float mean = mean_in[row], rstd = rstd_in[row];
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int col = c * 256 + lane * 8;
      if (col < d) {
        float fc[8];
        load8(sc + col, fc);   // L1-resident; recomputing 1+scale keeps 8*NC registers free
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float one_plus = __bfloat162float(__float2bfloat16(1.f + fc[j]));
          const float xhat = (xh[c][j] - mean) * rstd;
          const float dyv = g[c][j];
          a_sh[c][j] += dyv;
          a_sc[c][j] += dyv * xhat;
          const float gv = dyv * one_plus;
          xh[c][j] = xhat;
          g[c][j] = gv;
          sg += gv;
          sgx += gv * xhat;
        }
      }
    }
    const float mg = warp_sum(sg) / d, mgx = warp_sum(sgx) / d;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int col = c * 256 + lane * 8;
      if (col < d) {
        const float2 r0v = unpack_bf16x2(rres[c].x), r1v = unpack_bf16x2(rres[c].y),
                     r2v = unpack_bf16x2(rres[c].z), r3v = unpack_bf16x2(rres[c].w);
        float o[8] = {r0v.x, r0v.y, r1v.x, r1v.y, r2v.x, r2v.y, r3v.x, r3v.y};
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += rstd * (g[c][j] - mg - xh[c][j] * mgx);
        store8(dx + row * d + col, o);
      }
    }
  }
  block_fold_partials<NC>(red, a_sh, a_sc, partial + (long long)blockIdx.x * 2 * d, d, warp, lane);
}

// ------------------------------------------------------------------ gate bwd
// forward was o = a*g[b] + x.  da = do*g[b]; dg[b] += sum_rows do*a;
// dab[b] += sum_rows da (per-batch partial of the bias gradient, optional).
template <int NC>
__global__ void __launch_bounds__(BWD_THREADS, NC <= 3 ? 4 : 2)
gate_bwd_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ a,
                const bf16* __restrict__ gate, bf16* __restrict__ da, float* __restrict__ partial,
                int d, long long rows_per_batch, long long ld_gate, int rows_per_block,
                int blocks_per_batch) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  extern __shared__ float red[];  // [BWD_WARPS][2][d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = blockIdx.x / blocks_per_batch;
  const int chunk = blockIdx.x % blocks_per_batch;
  const long long r0 = b * rows_per_batch + (long long)chunk * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > (b + 1) * rows_per_batch) r1 = (b + 1) * rows_per_batch;
  float g[NC][8];
  load_row<NC>(gate + b * ld_gate, d, lane, g);
  float a_g[NC][8], a_b[NC][8];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) a_g[c][j] = a_b[c][j] = 0.f;
  for (long long row = r0 + warp; row < r1; row += BWD_WARPS) {
    float dv[NC][8], av[NC][8];
    load_row<NC>(dout + row * d, d, lane, dv);
    load_row<NC>(a + row * d, d, lane, av);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int col = c * 256 + lane * 8;
      if (col < d) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a_g[c][j] += dv[c][j] * av[c][j];
          o[j] = dv[c][j] * g[c][j];
          a_b[c][j] += o[j];
        }
        store8(da + row * d + col, o);
      }
    }
  }
  block_fold_partials<NC>(red, a_g, a_b, partial + (long long)blockIdx.x * 2 * d, d, warp, lane);
}

// ------------------------------------------------------- text RMSNorm * scalar
// out = bf16( sigma * rmsnorm_fp32(c) * w ), first `split` tokens of every
// sample use (w1, sigma1) and go to out1 [B*split, d]; the rest use (w2, sigma2)
// and go to out2 [B*(M-split), d].
template <int NC>
__global__ void __launch_bounds__(ROW_THREADS)
text_norm_fwd_kernel(const bf16* __restrict__ c, const float* __restrict__ w1,
                     const float* __restrict__ w2, const float* __restrict__ sigma1,
                     const float* __restrict__ sigma2, bf16* __restrict__ out1,
                     bf16* __restrict__ out2, float* __restrict__ rstd_out, long long R, int d,
                     int M, int split, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * ROW_WARPS + warp;
  if (row >= R) return;
  const long long b = row / M;
  const int tok = (int)(row % M);
  float v[NC][8];
  load_row<NC>(c + row * d, d, lane, v);
  float q = 0.f;
#pragma unroll
  for (int cc = 0; cc < NC; ++cc)
#pragma unroll
    for (int j = 0; j < 8; ++j) q += v[cc][j] * v[cc][j];
  const float rstd = rsqrtf(warp_sum(q) / d + eps);
  if (lane == 0 && rstd_out) rstd_out[row] = rstd;
  const bool first = tok < split;
  const float* w = first ? w1 : w2;
  const float sg = first ? *sigma1 : *sigma2;
  bf16* o = first ? out1 + (b * split + tok) * (long long)d
                  : out2 + (b * (M - split) + (tok - split)) * (long long)d;
#pragma unroll
  for (int cc = 0; cc < NC; ++cc) {
    const int col = cc * 256 + lane * 8;
    if (col < d) {
      float r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = sg * (v[cc][j] * rstd * w[col + j]);
      store8(o + col, r);
    }
  }
}

// backward of one half: dn = dL/d(out) [rows, d] (rows of this half, half-major),
// dsigma += sum dn * (chat*w) ; dw[col] += sum_rows dn*chat*sigma.  No grad to c.
template <int NC>
__device__ __forceinline__ void load_row_f32(const float* p, int d, int lane, float (&v)[NC][8]) {
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = c * 256 + lane * 8;
    if (col < d) {
      const float4 a = *reinterpret_cast<const float4*>(p + col);
      const float4 b = *reinterpret_cast<const float4*>(p + col + 4);
      v[c][0] = a.x; v[c][1] = a.y; v[c][2] = a.z; v[c][3] = a.w;
      v[c][4] = b.x; v[c][5] = b.y; v[c][6] = b.z; v[c][7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[c][j] = 0.f;
    }
  }
}

template <int NC, bool DN_F32>
__global__ void __launch_bounds__(ROW_THREADS)
text_norm_bwd_kernel(const void* __restrict__ dn_, const bf16* __restrict__ c,
                     const float* __restrict__ rstd_in, const float* __restrict__ w,
                     const float* __restrict__ sigma, float* __restrict__ dw,
                     float* __restrict__ dsigma, long long rows, int d, int M, int tok0, int ntok,
                     int rows_per_block) {
  extern __shared__ float red[];  // [d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  for (int i = threadIdx.x; i < d; i += ROW_THREADS) red[i] = 0.f;
  __syncthreads();
  const float sg = *sigma;
  float acc[NC][8];
#pragma unroll
  for (int cc = 0; cc < NC; ++cc)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[cc][j] = 0.f;
  float ds = 0.f;
  for (long long r = r0 + warp; r < r1; r += ROW_WARPS) {
    const long long b = r / ntok;
    const int tok = tok0 + (int)(r % ntok);
    const long long crow = b * M + tok;
    float g[NC][8], cv[NC][8];
    if constexpr (DN_F32) load_row_f32<NC>(reinterpret_cast<const float*>(dn_) + r * d, d, lane, g);
    else load_row<NC>(reinterpret_cast<const bf16*>(dn_) + r * d, d, lane, g);
    load_row<NC>(c + crow * d, d, lane, cv);
    const float rstd = rstd_in[crow];
#pragma unroll
    for (int cc = 0; cc < NC; ++cc) {
      const int col = cc * 256 + lane * 8;
      if (col < d) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float chat = cv[cc][j] * rstd;
          acc[cc][j] += g[cc][j] * chat;
          ds += g[cc][j] * chat * w[col + j];
        }
      }
    }
  }
#pragma unroll
  for (int cc = 0; cc < NC; ++cc) {
    const int col = cc * 256 + lane * 8;
    if (col < d) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&red[col + j], acc[cc][j] * sg);
    }
  }
  ds = warp_sum(ds);
  if (lane == 0) atomicAdd(dsigma, ds);
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += ROW_THREADS) atomicAdd(dw + i, red[i]);
}

#define DISPATCH_NC(d, CALL)                                                        \
  do {                                                                              \
    const int nc_ = ((d) + 255) / 256;                                              \
    switch (nc_) {                                                                  \
      case 1: { constexpr int NC = 1; CALL; } break;                                \
      case 2: { constexpr int NC = 2; CALL; } break;                                \
      case 3: { constexpr int NC = 3; CALL; } break;                                \
      case 4: { constexpr int NC = 4; CALL; } break;                                \
      case 5: { constexpr int NC = 5; CALL; } break;                                \
      case 6: { constexpr int NC = 6; CALL; } break;                                \
      case 7: { constexpr int NC = 7; CALL; } break;                                \
      case 8: { constexpr int NC = 8; CALL; } break;                                \
      case 9: { constexpr int NC = 9; CALL; } break;                                \
      default:                                                                      \
        set_last_error("row width %d unsupported (max 2304)", (int)(d));            \
        return MMDIT_ERR_UNSUPPORTED;                                               \
    }                                                                               \
  } while (0)

}  // namespace mmdit

using namespace mmdit;

extern "C" {

int mmdit_ln_modulate_fwd(const void* x, const void* shift, const void* scale, void* y, float* mean,
                          float* rstd, int64_t rows, int32_t d, int64_t rows_per_batch,
                          int64_t ld_mod, float eps, void* stream) {
  MMDIT_REQUIRE(x && shift && scale && y && rows > 0 && d > 0 && d % 8 == 0 && rows_per_batch > 0,
                MMDIT_ERR_ARG, "ln_modulate_fwd: bad arguments (d must be a multiple of 8)");
  if (row_kernel_generation() >= 2) {
    const int rc = ln_modulate_fwd_v2(x, shift, scale, y, mean, rstd, rows, d, rows_per_batch, ld_mod, eps,
                                      (cudaStream_t)stream);
    if (rc != ROW_V2_UNSUPPORTED) return rc;
  }
  const unsigned grid = (unsigned)((rows + ROW_WARPS - 1) / ROW_WARPS);
  DISPATCH_NC(d, MMDIT_CARVEOUT(ln_mod_fwd_kernel<NC>); (launch_k(ln_mod_fwd_kernel<NC>, grid, dim3(ROW_THREADS), 0, (cudaStream_t)stream, 
                     (const bf16*)x, (const bf16*)shift, (const bf16*)scale, (bf16*)y, mean, rstd,
                     rows, d, rows_per_batch, ld_mod, eps)));
  return check_launch("ln_mod_fwd_kernel");
}

int mmdit_gate_residual_ln_fwd(const void* a, const void* gate, const void* resid, const void* shift,
                               const void* scale, void* x_out, void* y, float* mean, float* rstd,
                               int64_t rows, int32_t d, int64_t rows_per_batch, int64_t ld_gate,
                               int64_t ld_mod, float eps, void* stream) {
  MMDIT_REQUIRE(a && gate && resid && shift && scale && x_out && y && rows > 0 && d > 0 && d % 8 == 0 &&
                    rows_per_batch > 0 && ld_gate % 8 == 0 && ld_mod % 8 == 0,
                MMDIT_ERR_ARG, "gate_residual_ln_fwd: bad arguments (d, ld_gate, ld_mod must be multiples of 8)");
  if (row_kernel_generation() >= 2) {
    const int rc = gate_residual_ln_fwd_v2(a, gate, resid, shift, scale, x_out, y, mean, rstd, rows, d,
                                           rows_per_batch, ld_gate, ld_mod, eps, (cudaStream_t)stream);
    if (rc != ROW_V2_UNSUPPORTED) return rc;
  }
  const unsigned grid = (unsigned)((rows + ROW_WARPS - 1) / ROW_WARPS);
  DISPATCH_NC(d, MMDIT_CARVEOUT(gate_res_ln_fwd_kernel<NC>); (launch_k(gate_res_ln_fwd_kernel<NC>, grid, dim3(ROW_THREADS), 0, (cudaStream_t)stream, 
                     (const bf16*)a, (const bf16*)gate, (const bf16*)resid, (const bf16*)shift,
                     (const bf16*)scale, (bf16*)x_out, (bf16*)y, mean, rstd, rows, d, rows_per_batch,
                     ld_gate, ld_mod, eps)));
  return check_launch("gate_res_ln_fwd_kernel");
}

// Rows a block of the backward row kernels reduces before it writes one partial: chosen so that the
// grid is ONE wave of resident blocks (blocks per SM from the occupancy calculator x SMs, split evenly
// over the samples).  With the fixed 16 rows of round 1 the cfg2 grids were 1.73 (image) and 1.08 (text)
// waves.  MMDIT_ROW_RPB overrides.  Never below 8 rows (bounds the partial workspace).
constexpr int ROW_RPB_MIN = 8;
static int rows_per_block(const void* kernel, size_t smem, long long rows_per_batch, int nb) {
  static const int forced = [] {
    const char* e = getenv("MMDIT_ROW_RPB");
    const int x = e ? atoi(e) : 0;
    return x >= ROW_RPB_MIN && x <= 1024 ? x : 0;
  }();
  if (forced) return forced;
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, BWD_THREADS, smem) != cudaSuccess || occ < 1) {
    cudaGetLastError();
    occ = 2;
  }
  long long per_sample = (long long)occ * num_sms() / nb;
  if (per_sample < 1) per_sample = 1;
  long long rpb = (rows_per_batch + per_sample - 1) / per_sample;
  if (rpb < ROW_RPB_MIN) rpb = ROW_RPB_MIN;
  return (int)rpb;
}

int64_t mmdit_rowreduce_workspace_floats(int64_t rows, int32_t d, int64_t rows_per_batch) {
  if (rows <= 0 || d <= 0 || rows_per_batch <= 0) return 0;
  const int64_t bpb = (rows_per_batch + ROW_RPB_MIN - 1) / ROW_RPB_MIN;   // upper bound over every block size
  return (rows / rows_per_batch) * bpb * 2 * d;
}

int mmdit_ln_modulate_bwd(const void* dy, const void* x, const float* mean, const float* rstd,
                          const void* scale, const void* dres, void* dx, void* dshift,
                          void* dscale, int32_t dmod_bf16, float* workspace, int64_t rows, int32_t d,
                          int64_t rows_per_batch, int64_t ld_mod, int64_t ld_dmod, void* stream) {
  MMDIT_REQUIRE(dy && x && mean && rstd && scale && dx && dshift && dscale && workspace && rows > 0 &&
                    d % 8 == 0 && rows_per_batch > 0 && rows % rows_per_batch == 0,
                MMDIT_ERR_ARG, "ln_modulate_bwd: bad arguments");
  const int nb = (int)(rows / rows_per_batch);
  const size_t smem = (size_t)BWD_WARPS * 2 * d * sizeof(float);
  MMDIT_REQUIRE(smem <= 48 * 1024, MMDIT_ERR_UNSUPPORTED, "ln_modulate_bwd: d=%d too wide", d);
  int bpb = 0, v2 = ROW_V2_UNSUPPORTED;
  if (row_kernel_generation() >= 2)
    v2 = ln_modulate_bwd_v2(dy, x, mean, rstd, scale, dres, dx, dshift, dscale, dmod_bf16, ld_dmod, workspace, rows,
                            d, rows_per_batch, ld_mod, &bpb, (cudaStream_t)stream);
  if (v2 == ROW_V2_UNSUPPORTED) {
    int rpb = 16;
    DISPATCH_NC(d, rpb = rows_per_block((const void*)ln_mod_bwd_kernel<NC>, smem, rows_per_batch, nb));
    bpb = (int)((rows_per_batch + rpb - 1) / rpb);
    const unsigned grid = (unsigned)(nb * bpb);
    DISPATCH_NC(d, MMDIT_CARVEOUT(ln_mod_bwd_kernel<NC>); (launch_k(ln_mod_bwd_kernel<NC>, grid, dim3(BWD_THREADS), smem, (cudaStream_t)stream,
                       (const bf16*)dy, (const bf16*)x, mean, rstd, (const bf16*)scale,
                       (const bf16*)dres, (bf16*)dx, workspace, d, rows_per_batch, ld_mod, rpb, bpb)));
  } else if (v2 != MMDIT_OK) {
    return v2;
  }
  if (bpb == 0) return MMDIT_OK;   // the sample's strips ran as one cluster and folded their column sums
  dim3 g2((2 * d + 255) / 256, nb);
  MMDIT_CARVEOUT(fold_batch_partials_kernel);
  launch_k(fold_batch_partials_kernel, g2, dim3(256), 0, (cudaStream_t)stream, workspace, dshift, dscale, d, bpb,
                                                                  ld_dmod, ld_dmod, dmod_bf16, dmod_bf16);
  return check_launch(v2 == MMDIT_OK ? "fold_batch_partials_kernel" : "ln_mod_bwd_kernel", v2 == MMDIT_OK ? 1 : 2);
}

int mmdit_gate_bwd(const void* dout, const void* a, const void* gate, void* da, void* dgate,
                   int32_t dgate_bf16, float* dab, float* workspace, int64_t rows, int32_t d,
                   int64_t rows_per_batch, int64_t ld_gate, int64_t ld_dgate, int64_t ld_dab,
                   void* stream) {
  MMDIT_REQUIRE(dout && a && gate && da && dgate && workspace && rows > 0 && d % 8 == 0 &&
                    rows_per_batch > 0 && rows % rows_per_batch == 0,
                MMDIT_ERR_ARG, "gate_bwd: bad arguments");
  const int nb = (int)(rows / rows_per_batch);
  const size_t smem = (size_t)BWD_WARPS * 2 * d * sizeof(float);
  MMDIT_REQUIRE(smem <= 48 * 1024, MMDIT_ERR_UNSUPPORTED, "gate_bwd: d=%d too wide", d);
  int bpb = 0, v2 = ROW_V2_UNSUPPORTED;
  if (row_kernel_generation() >= 2)
    v2 = gate_bwd_v2(dout, a, gate, da, dgate, dgate_bf16, dab, ld_dgate, ld_dab, workspace, rows, d, rows_per_batch,
                     ld_gate, &bpb, (cudaStream_t)stream);
  if (v2 == ROW_V2_UNSUPPORTED) {
    int rpb = 16;
    DISPATCH_NC(d, rpb = rows_per_block((const void*)gate_bwd_kernel<NC>, smem, rows_per_batch, nb));
    bpb = (int)((rows_per_batch + rpb - 1) / rpb);
    const unsigned grid = (unsigned)(nb * bpb);
    DISPATCH_NC(d, MMDIT_CARVEOUT(gate_bwd_kernel<NC>); (launch_k(gate_bwd_kernel<NC>, grid, dim3(BWD_THREADS), smem, (cudaStream_t)stream,
                       (const bf16*)dout, (const bf16*)a, (const bf16*)gate, (bf16*)da, workspace, d,
                       rows_per_batch, ld_gate, rpb, bpb)));
  } else if (v2 != MMDIT_OK) {
    return v2;
  }
  if (bpb == 0) return MMDIT_OK;   // folded inside the cluster
  dim3 g2((2 * d + 255) / 256, nb);
  launch_k(fold_batch_partials_kernel, g2, dim3(256), 0, (cudaStream_t)stream, workspace, dgate, dab, d, bpb,
                                                                  ld_dgate, ld_dab, dgate_bf16, 0);
  return check_launch(v2 == MMDIT_OK ? "fold_batch_partials_kernel" : "gate_bwd_kernel", v2 == MMDIT_OK ? 1 : 2);
}

int mmdit_text_norm_fwd(const void* c, const float* w1, const float* w2, const float* sigma1,
                        const float* sigma2, void* out1, void* out2, float* rstd, int64_t batch,
                        int32_t tokens, int32_t split, int32_t d, float eps, void* stream) {
  MMDIT_REQUIRE(c && w1 && w2 && sigma1 && sigma2 && out1 && batch > 0 && tokens >= split &&
                    split >= 0 && d % 8 == 0 && (out2 || tokens == split),
                MMDIT_ERR_ARG, "text_norm_fwd: bad arguments");
  const long long R = batch * (long long)tokens;
  const unsigned grid = (unsigned)((R + ROW_WARPS - 1) / ROW_WARPS);
  DISPATCH_NC(d, (text_norm_fwd_kernel<NC><<<grid, ROW_THREADS, 0, (cudaStream_t)stream>>>(
                     (const bf16*)c, w1, w2, sigma1, sigma2, (bf16*)out1, (bf16*)out2, rstd, R, d,
                     tokens, split, eps)));
  return check_launch("text_norm_fwd_kernel");
}

int mmdit_text_norm_bwd(const void* dn, int32_t dn_fp32, const void* c, const float* rstd,
                        const float* w, const float* sigma, float* dw, float* dsigma, int64_t batch,
                        int32_t tokens, int32_t tok0, int32_t ntok, int32_t d, void* stream) {
  MMDIT_REQUIRE(dn && c && rstd && w && sigma && dw && dsigma && batch > 0 && ntok > 0 &&
                    tok0 + ntok <= tokens && d % 8 == 0,
                MMDIT_ERR_ARG, "text_norm_bwd: bad arguments");
  const long long rows = batch * (long long)ntok;
  const int rpb = 64;
  const unsigned grid = (unsigned)((rows + rpb - 1) / rpb);
  const size_t smem = (size_t)d * sizeof(float);
  if (dn_fp32) {
    DISPATCH_NC(d, (text_norm_bwd_kernel<NC, true><<<grid, ROW_THREADS, smem, (cudaStream_t)stream>>>(
                       dn, (const bf16*)c, rstd, w, sigma, dw, dsigma, rows, d, tokens, tok0, ntok, rpb)));
  } else {
    DISPATCH_NC(d, (text_norm_bwd_kernel<NC, false><<<grid, ROW_THREADS, smem, (cudaStream_t)stream>>>(
                       dn, (const bf16*)c, rstd, w, sigma, dw, dsigma, rows, d, tokens, tok0, ntok, rpb)));
  }
  return check_launch("text_norm_bwd_kernel");
}

}  // extern "C"
