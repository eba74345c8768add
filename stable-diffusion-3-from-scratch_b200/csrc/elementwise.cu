// Elementwise / small-reduction kernels of the MMDiT step (bf16 activations,
// fp32 math, 128-bit vector accesses):
//   per-head QK RMSNorm + 2-D axial RoPE fwd/bwd   (Attention.py:61-64,130-134,174-194;
//                                                   rotary_embedding.py:36-76,269-288)
//   SwiGLU activation fwd/bwd                      (MLP.py:32, xformers SwiGLU semantics)
//   timestep embedding fwd/bwd                     (PositionalEncoding.py:23-30, diff_model.py:306)
//   patchify / unpatchify                          (ImagePositionalEncoding.py:181-183, patchify.py:41-72)
//   rectified-flow noising, velocity loss, CFG+Euler update
//                                                  (diff_model.py:229-241,419-429; model_trainer.py:429-446)
//   column sums (bias gradients), row folds, dtype casts
#include <stdlib.h>

#include "common.cuh"
#include "mmdit_b200.h"

namespace mmdit {

// --------------------------------------------------------- QK RMSNorm + RoPE
// qkv: [R, ld_in] with q at column 0 and k at column d (raw projections).
// out: [R, ld_out] with normalised(+rotated) q at column 0 and k at column d.
// rope_cos/sin: [tokens_per_sample, 32] fp32 (angle per interleaved pair), or
// NULL for the text stream.  One thread = 8 consecutive columns of q and of k;
// a head (64 columns) is 8 adjacent lanes -> xor-shuffle reduction.
__global__ void __launch_bounds__(256)
qknorm_rope_fwd_kernel(const bf16* __restrict__ qkv, const float* __restrict__ wq,
                       const float* __restrict__ wk, const float* __restrict__ rope_cos,
                       const float* __restrict__ rope_sin, bf16* __restrict__ out, long long R,
                       int d, long long ld_in, long long ld_out, int tokens_per_sample, float eps) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  const int groups = d / 8;
  const long long total = R * groups;
  // total and the grid stride are multiples of 8, so the 8 lanes of a head are
  // active together and every shuffle below runs with a full warp.
  const long long start = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long iters = (total + stride - 1) / stride;
  for (long long it = 0; it < iters; ++it) {
    const long long idx = start + it * stride;
    const bool active = idx < total;
    const long long row = active ? idx / groups : 0;
    const int col = active ? (int)(idx % groups) * 8 : 0;
    const int pos = col & 63;  // position inside the head
    float q[8], k[8];
    load8(qkv + row * ld_in + col, q);
    load8(qkv + row * ld_in + d + col, k);
    float sq = 0.f, sk = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { sq += q[j] * q[j]; sk += k[j] * k[j]; }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
      sk += __shfl_xor_sync(0xffffffffu, sk, o);
    }
    const float rq = rsqrtf(sq * (1.f / 64.f) + eps), rk = rsqrtf(sk * (1.f / 64.f) + eps);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // the reference rounds the RMSNorm output to bf16 before the fp32 rotation
      q[j] = __bfloat162float(__float2bfloat16(q[j] * rq * wq[pos + j]));
      k[j] = __bfloat162float(__float2bfloat16(k[j] * rk * wk[pos + j]));
    }
    if (rope_cos) {
      const int tok = (int)(row % tokens_per_sample);
      const float* cs = rope_cos + (long long)tok * 32 + (pos >> 1);
      const float* sn = rope_sin + (long long)tok * 32 + (pos >> 1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float c = cs[j], s = sn[j];
        const float q0 = q[2 * j], q1 = q[2 * j + 1], k0 = k[2 * j], k1 = k[2 * j + 1];
        q[2 * j] = q0 * c - q1 * s;
        q[2 * j + 1] = q1 * c + q0 * s;
        k[2 * j] = k0 * c - k1 * s;
        k[2 * j + 1] = k1 * c + k0 * s;
      }
    }
    if (active) {
      store8(out + row * ld_out + col, q);
      store8(out + row * ld_out + d + col, k);
    }
  }
}

// backward: dqk = grad wrt the normalised/rotated q,k ([R, ld_g], q at 0, k at d);
// raw qkv as in forward; writes grad wrt raw q,k into dqkv ([R, ld_dout], q at 0, k at d)
// and accumulates dwq/dwk (64 floats each, fp32 atomics).
// DQ_F32: the q half of the incoming gradient is read straight from the attention backward's fp32
// accumulator dq_acc [B, acc_tokens, d] (rows acc_tok_off .. of every sample belong to this stream)
// instead of a bf16 copy: the separate convert pass (write + re-read of a [R, d] bf16 tensor and one
// launch per stream) disappears; the value is rounded to bf16 first, exactly as the copy was.
template <bool DQ_F32>
__global__ void __launch_bounds__(256, 3)
qknorm_rope_bwd_kernel(const float* __restrict__ dq_acc, int acc_tokens, int acc_tok_off,
                       const bf16* __restrict__ dqk, const bf16* __restrict__ qkv,
                       const float* __restrict__ wq, const float* __restrict__ wk,
                       const float* __restrict__ rope_cos, const float* __restrict__ rope_sin,
                       bf16* __restrict__ dqkv, float* __restrict__ dwq, float* __restrict__ dwk,
                       long long R, int d, long long ld_g, long long ld_in, long long ld_dout,
                       int tokens_per_sample, float eps) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  __shared__ float red[2][64];
  if (threadIdx.x < 128) (&red[0][0])[threadIdx.x] = 0.f;
  __syncthreads();
  const int groups = d / 8;
  const long long total = R * groups;
  float awq[8], awk[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) awq[j] = awk[j] = 0.f;
  // grid stride is a multiple of 8 and groups is a multiple of 8, so (idx % 8)
  // -- the 8-column slot inside the head -- is fixed per thread.
  const long long start = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long iters = (total + stride - 1) / stride;
  for (long long it = 0; it < iters; ++it) {
    const long long idx = start + it * stride;
    const bool active = idx < total;
    const long long row = active ? idx / groups : 0;
    const int col = active ? (int)(idx % groups) * 8 : 0;
    const int pos = col & 63;
    float q[8], k[8], gq[8], gk[8];
    if (active) {
      load8(qkv + row * ld_in + col, q);
      load8(qkv + row * ld_in + d + col, k);
      if constexpr (DQ_F32) {
        const long long arow = (row / tokens_per_sample) * acc_tokens + acc_tok_off + row % tokens_per_sample;
        const float4 a0 = *reinterpret_cast<const float4*>(dq_acc + arow * d + col);
        const float4 a1 = *reinterpret_cast<const float4*>(dq_acc + arow * d + col + 4);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) gq[j] = __bfloat162float(__float2bfloat16(av[j]));
      } else {
        load8(dqk + row * ld_g + col, gq);
      }
      load8(dqk + row * ld_g + d + col, gk);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) q[j] = k[j] = gq[j] = gk[j] = 0.f;
    }
    if (rope_cos && active) {  // rotate the incoming gradient by -theta
      const int tok = (int)(row % tokens_per_sample);
      const float* cs = rope_cos + (long long)tok * 32 + (pos >> 1);
      const float* sn = rope_sin + (long long)tok * 32 + (pos >> 1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float c = cs[j], s = sn[j];
        const float a0 = gq[2 * j], a1 = gq[2 * j + 1], b0 = gk[2 * j], b1 = gk[2 * j + 1];
        gq[2 * j] = a0 * c + a1 * s;
        gq[2 * j + 1] = a1 * c - a0 * s;
        gk[2 * j] = b0 * c + b1 * s;
        gk[2 * j + 1] = b1 * c - b0 * s;
      }
    }
    float sq = 0.f, sk = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { sq += q[j] * q[j]; sk += k[j] * k[j]; }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
      sk += __shfl_xor_sync(0xffffffffu, sk, o);
    }
    const float rq = rsqrtf(sq * (1.f / 64.f) + eps), rk = rsqrtf(sk * (1.f / 64.f) + eps);
    float mq = 0.f, mk = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float qh = q[j] * rq, kh = k[j] * rk;
      awq[j] += gq[j] * qh;
      awk[j] += gk[j] * kh;
      gq[j] *= wq[pos + j];
      gk[j] *= wk[pos + j];
      mq += gq[j] * qh;
      mk += gk[j] * kh;
      q[j] = qh;
      k[j] = kh;
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      mq += __shfl_xor_sync(0xffffffffu, mq, o);
      mk += __shfl_xor_sync(0xffffffffu, mk, o);
    }
    mq *= (1.f / 64.f);
    mk *= (1.f / 64.f);
    if (active) {
      float oq[8], ok[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        oq[j] = rq * (gq[j] - q[j] * mq);
        ok[j] = rk * (gk[j] - k[j] * mk);
      }
      store8(dqkv + row * ld_dout + col, oq);
      store8(dqkv + row * ld_dout + d + col, ok);
    }
  }
  const int slot = (threadIdx.x & 7) * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float a = awq[j], b = awk[j];
    a += __shfl_xor_sync(0xffffffffu, a, 8);
    a += __shfl_xor_sync(0xffffffffu, a, 16);
    b += __shfl_xor_sync(0xffffffffu, b, 8);
    b += __shfl_xor_sync(0xffffffffu, b, 16);
    if ((threadIdx.x & 31) < 8) {
      atomicAdd(&red[0][slot + j], a);
      atomicAdd(&red[1][slot + j], b);
    }
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    atomicAdd(dwq + threadIdx.x, red[0][threadIdx.x]);
    atomicAdd(dwk + threadIdx.x, red[1][threadIdx.x]);
  }
}

// ------------------------------------------------------------------- SwiGLU
// h12: [R, 2*hid] (x1 = gate half at column 0, x2 at column hid); a = silu(x1)*x2.
__global__ void __launch_bounds__(256)
swiglu_fwd_kernel(const bf16* __restrict__ h12, bf16* __restrict__ a, long long R, int hid) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  const int groups = hid / 8;
  const long long total = R * groups;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / groups;
    const int col = (int)(idx % groups) * 8;
    float x1[8], x2[8], o[8];
    load8(h12 + row * 2 * hid + col, x1);
    load8(h12 + row * 2 * hid + hid + col, x2);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = silu_f(x1[j]) * x2[j];
    store8(a + row * hid + col, o);
  }
}

// da: [R, hid]; writes dh12 [R, 2*hid] and accumulates db12 [2*hid] (fp32 atomics).
// Block = 128 threads x 8 columns = 1024 hidden columns, strip of rows_per_block rows.
// 80 registers (6 blocks of 128 per SM allowed): with the default heuristics ptxas held this kernel to 60
// registers and serialised part of the 12 loads of a trip behind the math; with the room all 12 are issued
// first.  Measured: cfg2 text 74.0 -> 57.2 us (0.62 -> 0.81 of the HBM rate), cfg3 image 193.6 -> 165.9 us
// (0.80 -> 0.93).  A register ping-pong over the next four rows (168 registers) was slower: 74.4 / 180.7 us.
__global__ void __launch_bounds__(128, 6)
swiglu_bwd_kernel(const bf16* __restrict__ da, const bf16* __restrict__ h12,
                  bf16* __restrict__ dh12, float* __restrict__ partial, long long R, int hid,
                  int rows_per_block) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  const int col = (blockIdx.x * 128 + threadIdx.x) * 8;
  if (col >= hid) return;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > R) r1 = R;
  float b1[8], b2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) b1[j] = b2[j] = 0.f;
  // 4 rows per trip: all 12 vector loads are issued before any math, so a warp keeps 6 KiB in
  // flight (the grid alone does not fill the machine: ~20 warps per SM at cfg2 sizes)
  auto body = [&](const uint4& ug, const uint4& u1, const uint4& u2, long long row) {
    const float2 ga = unpack_bf16x2(ug.x), gb = unpack_bf16x2(ug.y), gc = unpack_bf16x2(ug.z), gd = unpack_bf16x2(ug.w);
    const float2 xa = unpack_bf16x2(u1.x), xb = unpack_bf16x2(u1.y), xc = unpack_bf16x2(u1.z), xd = unpack_bf16x2(u1.w);
    const float2 ya = unpack_bf16x2(u2.x), yb = unpack_bf16x2(u2.y), yc = unpack_bf16x2(u2.z), yd = unpack_bf16x2(u2.w);
    const float g[8] = {ga.x, ga.y, gb.x, gb.y, gc.x, gc.y, gd.x, gd.y};
    const float x1[8] = {xa.x, xa.y, xb.x, xb.y, xc.x, xc.y, xd.x, xd.y};
    const float x2[8] = {ya.x, ya.y, yb.x, yb.y, yc.x, yc.y, yd.x, yd.y};
    float d1[8], d2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sg = __fdividef(1.f, 1.f + __expf(-x1[j]));   // MUFU.RCP: the IEEE division made this kernel issue-bound
      const float sl = x1[j] * sg;
      d1[j] = g[j] * x2[j] * sg * (1.f + x1[j] * (1.f - sg));
      d2[j] = g[j] * sl;
      b1[j] += d1[j];
      b2[j] += d2[j];
    }
    store8(dh12 + row * 2 * hid + col, d1);
    store8(dh12 + row * 2 * hid + hid + col, d2);
  };
  long long row = r0;
  struct Rows4 { uint4 ug[4], u1[4], u2[4]; };
  auto load4 = [&](Rows4& t, long long rw) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      t.ug[k] = *reinterpret_cast<const uint4*>(da + (rw + k) * hid + col);
      t.u1[k] = *reinterpret_cast<const uint4*>(h12 + (rw + k) * 2 * hid + col);
      t.u2[k] = *reinterpret_cast<const uint4*>(h12 + (rw + k) * 2 * hid + hid + col);
    }
  };
  auto body4 = [&](const Rows4& t, long long rw) {
#pragma unroll
    for (int k = 0; k < 4; ++k) body(t.ug[k], t.u1[k], t.u2[k], rw + k);
  };
  for (; row + 3 < r1; row += 4) {
    Rows4 t;
    load4(t, row);
    body4(t, row);
  }
  for (; row < r1; ++row) {
    const uint4 ug = *reinterpret_cast<const uint4*>(da + row * hid + col);
    const uint4 u1 = *reinterpret_cast<const uint4*>(h12 + row * 2 * hid + col);
    const uint4 u2 = *reinterpret_cast<const uint4*>(h12 + row * 2 * hid + hid + col);
    body(ug, u1, u2, row);
  }
  if (partial) {  // one fp32 row of column sums per row strip; folded by fold_rows_f32_kernel
    float* dst = partial + (long long)blockIdx.y * 2 * hid;
    *reinterpret_cast<float4*>(dst + col) = make_float4(b1[0], b1[1], b1[2], b1[3]);
    *reinterpret_cast<float4*>(dst + col + 4) = make_float4(b1[4], b1[5], b1[6], b1[7]);
    *reinterpret_cast<float4*>(dst + hid + col) = make_float4(b2[0], b2[1], b2[2], b2[3]);
    *reinterpret_cast<float4*>(dst + hid + col + 4) = make_float4(b2[4], b2[5], b2[6], b2[7]);
  }
}

// ------------------------------------------------------------ gated residual
// out[r,:] = a[r,:] * gate[r / rows_per_batch, :] + resid[r,:]   (Transformer_Block_Dual.py:64-76)
__global__ void __launch_bounds__(256)
gate_residual_fwd_kernel(const bf16* __restrict__ a, const bf16* __restrict__ gate,
                         const bf16* __restrict__ resid, bf16* __restrict__ out, long long R, int d,
                         long long rows_per_batch, long long ld_gate) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  const int groups = d / 8;
  const long long total = R * groups;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / groups;
    const int col = (int)(idx % groups) * 8;
    float av[8], gv[8], rv[8], o[8];
    load8(a + row * d + col, av);
    load8(resid + row * d + col, rv);
    load8(gate + (row / rows_per_batch) * ld_gate + col, gv);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(av[j], gv[j], rv[j]);
    store8(out + row * d + col, o);
  }
}

// ------------------------------------------------------------ timestep embed
// e[b, j]      = sin(t_b * s / den[2j])        j <  d/2
// e[b, d/2+j]  = cos(t_b * s / den[2j+1])
__global__ void timestep_embed_fwd_kernel(const float* __restrict__ t, const float* __restrict__ scale,
                                          const float* __restrict__ denom, bf16* __restrict__ out,
                                          int B, int d) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * d) return;
  const int b = idx / d, j = idx % d, half = d / 2;
  const float ts = t[b] * scale[0];
  const float v = j < half ? sinf(ts / denom[2 * j]) : cosf(ts / denom[2 * (j - half) + 1]);
  out[idx] = __float2bfloat16(v);
}
// dscale += sum_{b,j} de[b,j] * (t_b/den) * (cos | -sin)(t_b*s/den)
__global__ void timestep_embed_bwd_kernel(const bf16* __restrict__ de, const float* __restrict__ t,
                                          const float* __restrict__ scale,
                                          const float* __restrict__ denom, float* __restrict__ dscale,
                                          int B, int d) {
  __shared__ float red[32];
  const int half = d / 2;
  const float s = scale[0];
  float acc = 0.f;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < B * d; idx += gridDim.x * blockDim.x) {
    const int b = idx / d, j = idx % d;
    const float den = j < half ? denom[2 * j] : denom[2 * (j - half) + 1];
    const float u = t[b] / den, arg = t[b] * s / den;
    const float g = j < half ? cosf(arg) : -sinf(arg);
    acc += __bfloat162float(de[idx]) * u * g;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(dscale, v);
  }
}

// ----------------------------------------------------- patchify / unpatchify
// img [B,C,H,W] (fp32 or bf16) <-> tokens [B*(H/p)*(W/p), C*p*p] bf16, column = c*p*p + i*p + j
template <typename ImgT>
__global__ void patchify_kernel(const ImgT* __restrict__ img, bf16* __restrict__ tok, int B, int C,
                                int H, int W, int p) {
  const long long total = (long long)B * C * H * W;
  const int nw = W / p, nh = H / p;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(idx % W);
    const int h = (int)((idx / W) % H);
    const int c = (int)((idx / ((long long)W * H)) % C);
    const long long b = idx / ((long long)W * H * C);
    const long long row = (b * nh + h / p) * nw + w / p;
    const int col = c * p * p + (h % p) * p + (w % p);
    tok[row * (C * p * p) + col] = __float2bfloat16((float)img[idx]);
  }
}
template <typename ImgT>
__global__ void unpatchify_kernel(const bf16* __restrict__ tok, ImgT* __restrict__ img, int B, int C,
                                  int H, int W, int p) {
  const long long total = (long long)B * C * H * W;
  const int nw = W / p, nh = H / p;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(idx % W);
    const int h = (int)((idx / W) % H);
    const int c = (int)((idx / ((long long)W * H)) % C);
    const long long b = idx / ((long long)W * H * C);
    const long long row = (b * nh + h / p) * nw + w / p;
    const int col = c * p * p + (h % p) * p + (w % p);
    img[idx] = (ImgT)__bfloat162float(tok[row * (C * p * p) + col]);
  }
}

// ------------------------------------------------ rectified flow elementwise
// x_t = (1 - t_b) * x0 + t_b * eps   (fp32 out)
template <typename InT>
__global__ void rf_noise_kernel(const InT* __restrict__ x0, const InT* __restrict__ eps,
                                const float* __restrict__ t, float* __restrict__ xt,
                                long long per_sample, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const float tb = t[idx / per_sample];
    xt[idx] = (1.f - tb) * (float)x0[idx] + tb * (float)eps[idx];
  }
}
// loss += sum (v - (eps - x0))^2 * inv_numel ; diff = v - (eps - x0)  (fp32)
template <typename VT, typename InT>
__global__ void rf_loss_kernel(const VT* __restrict__ v, const InT* __restrict__ eps,
                               const InT* __restrict__ x0, float* __restrict__ diff,
                               float* __restrict__ loss, long long total, float inv_numel) {
  __shared__ float red[32];
  float acc = 0.f;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const float df = (float)v[idx] - ((float)eps[idx] - (float)x0[idx]);
    if (diff) diff[idx] = df;
    acc += df * df;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(loss, s * inv_numel);
  }
}
// dv = diff * (2/numel) * upstream   (bf16 or fp32 out)
template <typename VT>
__global__ void rf_loss_bwd_kernel(const float* __restrict__ diff, const float* __restrict__ upstream,
                                   VT* __restrict__ dv, long long total, float two_inv_numel) {
  const float g = upstream[0] * two_inv_numel;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x)
    dv[idx] = (VT)(diff[idx] * g);
}
// x -= ((1+w) v[:B] - w v[B:]) * dt      (x fp32 in place)
template <typename VT>
__global__ void cfg_euler_kernel(float* __restrict__ x, const VT* __restrict__ v, long long half,
                                 float w, float dt) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < half;
       idx += (long long)gridDim.x * blockDim.x) {
    const float vc = (float)v[idx], vu = (float)v[half + idx];
    x[idx] -= ((1.f + w) * vc - w * vu) * dt;
  }
}

// -------------------------------------------------------------- reductions
// out[n] += sum_rows in[row, n]   (bf16 in, fp32 out). Block: 128 threads x 8 cols, row strip.
__global__ void __launch_bounds__(128)
colsum_bf16_kernel(const bf16* __restrict__ in, float* __restrict__ out, long long R, int n,
                   long long ld, int rows_per_block) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  const int col = (blockIdx.x * 128 + threadIdx.x) * 8;
  if (col >= n) return;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > R) r1 = R;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (long long row = r0; row < r1; ++row) {
    float v[8];
    load8(in + row * ld + col, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += v[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(out + col + j, acc[j]);
}
// out[n] += sum_rows in[row, n] (fp32 in), small row counts (per-batch partials).
__global__ void fold_rows_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int R,
                                     int n, long long ld) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  // grid.y row groups, each adds its partial column sum with one atomic (few groups -> cheap)
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= n) return;
  const int per = (R + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * per, r1 = min(R, r0 + per);
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  int r = r0;
  for (; r + 3 < r1; r += 4) {
    acc0 += in[(long long)r * ld + col];
    acc1 += in[(long long)(r + 1) * ld + col];
    acc2 += in[(long long)(r + 2) * ld + col];
    acc3 += in[(long long)(r + 3) * ld + col];
  }
  for (; r < r1; ++r) acc0 += in[(long long)r * ld + col];
  if (r1 > r0) atomicAdd(out + col, (acc0 + acc1) + (acc2 + acc3));
}
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out,
                                     long long n) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  const long long n8 = n / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(in)[2 * i];
    const float4 b = reinterpret_cast<const float4*>(in)[2 * i + 1];
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    store8(out + 8 * i, v);
  }
  for (long long i = n8 * 8 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16(in[i]);
}

// out[i] (+)= sum_s ws[s*stride + i]   -- folds split-K slices (fp32, 16-byte vectors)
__global__ void __launch_bounds__(256)
fold_slices_kernel(const float* __restrict__ ws, float* __restrict__ out, long long n, int slices,
                   long long stride, int accumulate) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  const long long n4 = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 acc = accumulate ? reinterpret_cast<const float4*>(out)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < slices; ++s) {
      const float4 v = reinterpret_cast<const float4*>(ws + s * stride)[i];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = acc;
  }
}

// out[0] = 1.02 * scale * 64 * max(|wq_x|,|wq_c|) * max(|wk_x|,|wk_c|): an upper bound of the scaled
// attention logits after per-head RMSNorm (unit-RMS vectors of length 64 times the norm weight;
// RoPE is a rotation), with 2 % slack for the bf16 rounding of q and k.
__global__ void qk_logit_bound_kernel(const float* __restrict__ wq_x, const float* __restrict__ wk_x,
                                      const float* __restrict__ wq_c, const float* __restrict__ wk_c,
                                      float scale, float* __restrict__ out) {
  const int i = threadIdx.x;  // 64 threads
  float q = fmaxf(fabsf(wq_x[i]), wq_c ? fabsf(wq_c[i]) : 0.f);
  float k = fmaxf(fabsf(wk_x[i]), wk_c ? fabsf(wk_c[i]) : 0.f);
  __shared__ float sq[2], sk[2];
  q = warp_max(q);
  k = warp_max(k);
  if ((i & 31) == 0) { sq[i >> 5] = q; sk[i >> 5] = k; }
  __syncthreads();
  if (i == 0) out[0] = 1.02f * scale * 64.f * fmaxf(sq[0], sq[1]) * fmaxf(sk[0], sk[1]);
}

static inline unsigned grid_for(long long work_items, int threads) {
  long long blocks = (work_items + threads - 1) / threads;
  static const int per_sm = [] {   // blocks per SM of the grid-stride kernels (tuning knob)
    const char* e = getenv("MMDIT_GRID_CAP");
    const int x = e ? atoi(e) : 16;
    return x >= 1 && x <= 64 ? x : 16;
  }();
  const long long cap = (long long)num_sms() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace mmdit

using namespace mmdit;

extern "C" {

int mmdit_qknorm_rope_fwd(const void* qkv, const float* wq, const float* wk, const float* rope_cos,
                          const float* rope_sin, void* out, int64_t rows, int32_t d, int64_t ld_in,
                          int64_t ld_out, int32_t tokens_per_sample, float eps, void* stream) {
  MMDIT_REQUIRE(qkv && wq && wk && out && rows > 0 && d > 0 && d % 64 == 0 && ld_in % 8 == 0 &&
                    ld_out % 8 == 0 && tokens_per_sample > 0 && (!rope_cos == !rope_sin),
                MMDIT_ERR_ARG, "qknorm_rope_fwd: bad arguments (head_dim is fixed at 64)");
  if (row_kernel_generation() >= 2) {
    const int rc = qknorm_rope_fwd_v2(qkv, wq, wk, rope_cos, rope_sin, out, rows, d, ld_in, ld_out,
                                      tokens_per_sample, eps, (cudaStream_t)stream);
    if (rc != ROW_V2_UNSUPPORTED) return rc;
  }
  const long long work = rows * (long long)(d / 8);
  MMDIT_CARVEOUT(qknorm_rope_fwd_kernel);
  launch_k(qknorm_rope_fwd_kernel, dim3(grid_for(work, 256)), dim3(256), 0, (cudaStream_t)stream, 
      (const bf16*)qkv, wq, wk, rope_cos, rope_sin, (bf16*)out, rows, d, ld_in, ld_out,
      tokens_per_sample, eps);
  return check_launch("qknorm_rope_fwd_kernel");
}

static unsigned qkn_cap() {   // resident blocks per SM x waves of the QK-norm backward grid (tuning knob)
  static const unsigned v = [] {
    const char* e = getenv("MMDIT_QKN_CAP");
    const int x = e ? atoi(e) : 3;
    return (unsigned)(x >= 1 && x <= 64 ? x : 3);
  }();
  return v;
}

int mmdit_qknorm_rope_bwd(const void* dqk, const void* qkv, const float* wq, const float* wk,
                          const float* rope_cos, const float* rope_sin, void* dqkv, float* dwq,
                          float* dwk, int64_t rows, int32_t d, int64_t ld_g, int64_t ld_in,
                          int64_t ld_dout, int32_t tokens_per_sample, float eps, void* stream) {
  MMDIT_REQUIRE(dqk && qkv && wq && wk && dqkv && dwq && dwk && rows > 0 && d % 64 == 0 &&
                    ld_g % 8 == 0 && ld_in % 8 == 0 && ld_dout % 8 == 0 && tokens_per_sample > 0,
                MMDIT_ERR_ARG, "qknorm_rope_bwd: bad arguments");
  if (row_kernel_generation() >= 2) {
    const int rc = qknorm_rope_bwd_v2(nullptr, 0, 0, dqk, qkv, wq, wk, rope_cos, rope_sin, dqkv, dwq, dwk, rows, d,
                                      ld_g, ld_in, ld_dout, tokens_per_sample, eps, (cudaStream_t)stream);
    if (rc != ROW_V2_UNSUPPORTED) return rc;
  }
  const long long work = rows * (long long)(d / 8);
  unsigned grid = grid_for(work, 256);
  const unsigned cap = (unsigned)num_sms() * qkn_cap();  // fewer, longer-lived blocks: fewer atomics
  if (grid > cap) grid = cap;
  MMDIT_CARVEOUT(qknorm_rope_bwd_kernel<false>);
  launch_k(qknorm_rope_bwd_kernel<false>, grid, dim3(256), 0, (cudaStream_t)stream, (const float*)nullptr, 0, 0,
      (const bf16*)dqk, (const bf16*)qkv, wq, wk, rope_cos, rope_sin, (bf16*)dqkv, dwq, dwk, rows,
      d, ld_g, ld_in, ld_dout, tokens_per_sample, eps);
  return check_launch("qknorm_rope_bwd_kernel");
}

int mmdit_qknorm_rope_bwd_acc(const float* dq_acc, int32_t acc_tokens, int32_t acc_tok_off, const void* dqk,
                              const void* qkv, const float* wq, const float* wk, const float* rope_cos,
                              const float* rope_sin, void* dqkv, float* dwq, float* dwk, int64_t rows,
                              int32_t d, int64_t ld_g, int64_t ld_in, int64_t ld_dout,
                              int32_t tokens_per_sample, float eps, void* stream) {
  MMDIT_REQUIRE(dq_acc && dqk && qkv && wq && wk && dqkv && dwq && dwk && rows > 0 && d % 64 == 0 &&
                    ld_g % 8 == 0 && ld_in % 8 == 0 && ld_dout % 8 == 0 && tokens_per_sample > 0 &&
                    acc_tok_off >= 0 && acc_tok_off + tokens_per_sample <= acc_tokens &&
                    rows % tokens_per_sample == 0,
                MMDIT_ERR_ARG, "qknorm_rope_bwd_acc: bad arguments");
  if (row_kernel_generation() >= 2) {
    const int rc = qknorm_rope_bwd_v2(dq_acc, acc_tokens, acc_tok_off, dqk, qkv, wq, wk, rope_cos, rope_sin, dqkv,
                                      dwq, dwk, rows, d, ld_g, ld_in, ld_dout, tokens_per_sample, eps,
                                      (cudaStream_t)stream);
    if (rc != ROW_V2_UNSUPPORTED) return rc;
  }
  const long long work = rows * (long long)(d / 8);
  unsigned grid = grid_for(work, 256);
  const unsigned cap = (unsigned)num_sms() * qkn_cap();  // fewer, longer-lived blocks: fewer atomics
  if (grid > cap) grid = cap;
  MMDIT_CARVEOUT(qknorm_rope_bwd_kernel<true>);
  launch_k(qknorm_rope_bwd_kernel<true>, grid, dim3(256), 0, (cudaStream_t)stream, dq_acc, (int)acc_tokens,
      (int)acc_tok_off, (const bf16*)dqk, (const bf16*)qkv, wq, wk, rope_cos, rope_sin, (bf16*)dqkv, dwq, dwk, rows,
      d, ld_g, ld_in, ld_dout, tokens_per_sample, eps);
  return check_launch("qknorm_rope_bwd_kernel<acc>");
}

int mmdit_gate_residual_fwd(const void* a, const void* gate, const void* resid, void* out,
                            int64_t rows, int32_t d, int64_t rows_per_batch, int64_t ld_gate,
                            void* stream) {
  MMDIT_REQUIRE(a && gate && resid && out && rows > 0 && d > 0 && d % 8 == 0 && rows_per_batch > 0 &&
                    ld_gate % 8 == 0,
                MMDIT_ERR_ARG, "gate_residual_fwd: bad arguments");
  const long long work = rows * (long long)(d / 8);
  MMDIT_CARVEOUT(gate_residual_fwd_kernel);
  launch_k(gate_residual_fwd_kernel, dim3(grid_for(work, 256)), dim3(256), 0, (cudaStream_t)stream, 
      (const bf16*)a, (const bf16*)gate, (const bf16*)resid, (bf16*)out, rows, d, rows_per_batch,
      ld_gate);
  return check_launch("gate_residual_fwd_kernel");
}

int mmdit_qk_logit_bound(const float* wq_x, const float* wk_x, const float* wq_c, const float* wk_c,
                         float scale, float* out, void* stream) {
  MMDIT_REQUIRE(wq_x && wk_x && out, MMDIT_ERR_ARG, "qk_logit_bound: bad arguments");
  qk_logit_bound_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(wq_x, wk_x, wq_c, wk_c, scale, out);
  return check_launch("qk_logit_bound_kernel");
}

int mmdit_swiglu_fwd(const void* h12, void* a, int64_t rows, int32_t hidden, void* stream) {
  MMDIT_REQUIRE(h12 && a && rows > 0 && hidden > 0 && hidden % 8 == 0, MMDIT_ERR_ARG,
                "swiglu_fwd: bad arguments");
  const long long work = rows * (long long)(hidden / 8);
  MMDIT_CARVEOUT(swiglu_fwd_kernel);
  launch_k(swiglu_fwd_kernel, dim3(grid_for(work, 256)), dim3(256), 0, (cudaStream_t)stream, (const bf16*)h12,
                                                                          (bf16*)a, rows, hidden);
  return check_launch("swiglu_fwd_kernel");
}

int64_t mmdit_swiglu_bwd_workspace_floats(int64_t rows, int32_t hidden) {
  return rows > 0 && hidden > 0 ? ((rows + 63) / 64) * 2 * (int64_t)hidden : 0;
}

int mmdit_swiglu_bwd(const void* da, const void* h12, void* dh12, float* db12, float* workspace,
                     int64_t rows, int32_t hidden, void* stream) {
  MMDIT_REQUIRE(da && h12 && dh12 && rows > 0 && hidden > 0 && hidden % 8 == 0 &&
                    (!db12 || workspace),
                MMDIT_ERR_ARG, "swiglu_bwd: bad arguments (db12 needs a workspace)");
  const int rpb = 64;
  const int nrb = (int)((rows + rpb - 1) / rpb);
  dim3 grid((unsigned)((hidden / 8 + 127) / 128), (unsigned)nrb);
  MMDIT_CARVEOUT(swiglu_bwd_kernel);
  launch_k(swiglu_bwd_kernel, grid, dim3(128), 0, (cudaStream_t)stream, (const bf16*)da, (const bf16*)h12,
           (bf16*)dh12, db12 ? workspace : nullptr, rows, hidden, rpb);
  if (db12)
    launch_k(fold_rows_f32_kernel, dim3((2 * hidden + 255) / 256, nrb >= 32 ? 16 : 1), dim3(256), 0, (cudaStream_t)stream, 
        workspace, db12, nrb, 2 * hidden, 2 * (long long)hidden);
  return check_launch("swiglu_bwd_kernel", db12 ? 2 : 1);
}

int mmdit_timestep_embed_fwd(const float* t, const float* time_scale, const float* denom, void* out,
                             int32_t batch, int32_t d, void* stream) {
  MMDIT_REQUIRE(t && time_scale && denom && out && batch > 0 && d > 0 && d % 2 == 0, MMDIT_ERR_ARG,
                "timestep_embed_fwd: bad arguments");
  const int n = batch * d;
  timestep_embed_fwd_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      t, time_scale, denom, (bf16*)out, batch, d);
  return check_launch("timestep_embed_fwd_kernel");
}

int mmdit_timestep_embed_bwd(const void* de, const float* t, const float* time_scale,
                             const float* denom, float* dscale, int32_t batch, int32_t d,
                             void* stream) {
  MMDIT_REQUIRE(de && t && time_scale && denom && dscale && batch > 0 && d > 0, MMDIT_ERR_ARG,
                "timestep_embed_bwd: bad arguments");
  const int n = batch * d;
  int grid = (n + 255) / 256;
  if (grid > 64) grid = 64;
  timestep_embed_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)de, t, time_scale,
                                                                   denom, dscale, batch, d);
  return check_launch("timestep_embed_bwd_kernel");
}

int mmdit_patchify(const void* img, int32_t img_fp32, void* tokens, int32_t B, int32_t C, int32_t H,
                   int32_t W, int32_t p, void* stream) {
  MMDIT_REQUIRE(img && tokens && B > 0 && C > 0 && p > 0 && H % p == 0 && W % p == 0, MMDIT_ERR_ARG,
                "patchify: bad arguments (H, W must be multiples of the patch size)");
  const long long total = (long long)B * C * H * W;
  if (img_fp32)
    patchify_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float*)img, (bf16*)tokens, B, C, H, W, p);
  else
    patchify_kernel<bf16><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const bf16*)img, (bf16*)tokens, B, C, H, W, p);
  return check_launch("patchify_kernel");
}

int mmdit_unpatchify(const void* tokens, void* img, int32_t img_fp32, int32_t B, int32_t C,
                     int32_t H, int32_t W, int32_t p, void* stream) {
  MMDIT_REQUIRE(img && tokens && B > 0 && C > 0 && p > 0 && H % p == 0 && W % p == 0, MMDIT_ERR_ARG,
                "unpatchify: bad arguments");
  const long long total = (long long)B * C * H * W;
  if (img_fp32)
    unpatchify_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const bf16*)tokens, (float*)img, B, C, H, W, p);
  else
    unpatchify_kernel<bf16><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const bf16*)tokens, (bf16*)img, B, C, H, W, p);
  return check_launch("unpatchify_kernel");
}

int mmdit_rf_noise(const void* x0, const void* eps, int32_t in_fp32, const float* t, float* xt,
                   int64_t batch, int64_t per_sample, void* stream) {
  MMDIT_REQUIRE(x0 && eps && t && xt && batch > 0 && per_sample > 0, MMDIT_ERR_ARG,
                "rf_noise: bad arguments");
  const long long total = batch * per_sample;
  if (in_fp32)
    rf_noise_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float*)x0, (const float*)eps, t, xt, per_sample, total);
  else
    rf_noise_kernel<bf16><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (const bf16*)x0, (const bf16*)eps, t, xt, per_sample, total);
  return check_launch("rf_noise_kernel");
}

int mmdit_rf_loss_fwd(const void* v, int32_t v_fp32, const void* eps, const void* x0,
                      int32_t in_fp32, float* diff, float* loss, int64_t numel, void* stream) {
  MMDIT_REQUIRE(v && eps && x0 && loss && numel > 0, MMDIT_ERR_ARG, "rf_loss_fwd: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), s);
  if (e != cudaSuccess) {
    set_last_error("rf_loss_fwd: memset: %s", cudaGetErrorString(e));
    return (int)e;
  }
  const float inv = 1.f / (float)numel;
  const unsigned grid = grid_for(numel, 256);
#define RF_LOSS_LAUNCH(VT, IT)                                                                   \
  rf_loss_kernel<VT, IT><<<grid, 256, 0, s>>>((const VT*)v, (const IT*)eps, (const IT*)x0, diff, \
                                              loss, numel, inv)
  if (v_fp32 && in_fp32) RF_LOSS_LAUNCH(float, float);
  else if (v_fp32) RF_LOSS_LAUNCH(float, bf16);
  else if (in_fp32) RF_LOSS_LAUNCH(bf16, float);
  else RF_LOSS_LAUNCH(bf16, bf16);
#undef RF_LOSS_LAUNCH
  return check_launch("rf_loss_kernel");
}

int mmdit_rf_loss_bwd(const float* diff, const float* upstream, void* dv, int32_t dv_fp32,
                      int64_t numel, void* stream) {
  MMDIT_REQUIRE(diff && upstream && dv && numel > 0, MMDIT_ERR_ARG, "rf_loss_bwd: bad arguments");
  const float g = 2.f / (float)numel;
  if (dv_fp32)
    rf_loss_bwd_kernel<float><<<grid_for(numel, 256), 256, 0, (cudaStream_t)stream>>>(
        diff, upstream, (float*)dv, numel, g);
  else
    rf_loss_bwd_kernel<bf16><<<grid_for(numel, 256), 256, 0, (cudaStream_t)stream>>>(
        diff, upstream, (bf16*)dv, numel, g);
  return check_launch("rf_loss_bwd_kernel");
}

int mmdit_cfg_euler_step(float* x, const void* v, int32_t v_fp32, int64_t half_numel,
                         float cfg_scale, float dt, void* stream) {
  MMDIT_REQUIRE(x && v && half_numel > 0, MMDIT_ERR_ARG, "cfg_euler_step: bad arguments");
  if (v_fp32)
    cfg_euler_kernel<float><<<grid_for(half_numel, 256), 256, 0, (cudaStream_t)stream>>>(
        x, (const float*)v, half_numel, cfg_scale, dt);
  else
    cfg_euler_kernel<bf16><<<grid_for(half_numel, 256), 256, 0, (cudaStream_t)stream>>>(
        x, (const bf16*)v, half_numel, cfg_scale, dt);
  return check_launch("cfg_euler_kernel");
}

int mmdit_colsum_bf16(const void* in, float* out, int64_t rows, int32_t n, int64_t ld,
                      void* stream) {
  MMDIT_REQUIRE(in && out && rows > 0 && n > 0 && n % 8 == 0 && ld % 8 == 0, MMDIT_ERR_ARG,
                "colsum_bf16: bad arguments");
  int rpb = 64;
  if (rows > 65535LL * rpb) rpb = (int)((rows + 65534) / 65535);
  dim3 grid((unsigned)((n / 8 + 127) / 128), (unsigned)((rows + rpb - 1) / rpb));
  launch_k(colsum_bf16_kernel, grid, dim3(128), 0, (cudaStream_t)stream, (const bf16*)in, out, rows, n, ld, rpb);
  return check_launch("colsum_bf16_kernel");
}

int mmdit_fold_rows_f32(const float* in, float* out, int32_t rows, int32_t n, int64_t ld,
                        void* stream) {
  MMDIT_REQUIRE(in && out && rows > 0 && n > 0, MMDIT_ERR_ARG, "fold_rows_f32: bad arguments");
  launch_k(fold_rows_f32_kernel, dim3((n + 255) / 256, rows >= 32 ? 16 : 1), dim3(256), 0, (cudaStream_t)stream, in, out, rows, n, ld);
  return check_launch("fold_rows_f32_kernel");
}

int mmdit_fold_slices_f32(const float* ws, float* out, int64_t n, int32_t slices, int64_t stride,
                          int32_t accumulate, void* stream) {
  MMDIT_REQUIRE(ws && out && n > 0 && n % 4 == 0 && slices > 0 && stride % 4 == 0 &&
                    ((uintptr_t)ws & 15) == 0 && ((uintptr_t)out & 15) == 0,
                MMDIT_ERR_ARG, "fold_slices_f32: bad arguments");
  MMDIT_CARVEOUT(fold_slices_kernel);
  launch_k(fold_slices_kernel, dim3(grid_for(n / 4, 256)), dim3(256), 0, (cudaStream_t)stream, ws, out, n, slices, stride,
                                                                             accumulate);
  return check_launch("fold_slices_kernel");
}

int mmdit_cast_f32_bf16(const float* in, void* out, int64_t n, void* stream) {
  MMDIT_REQUIRE(in && out && n > 0 && ((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0,
                MMDIT_ERR_ARG, "cast_f32_bf16: bad arguments (16-byte aligned buffers)");
  launch_k(cast_f32_bf16_kernel, dim3(grid_for(n / 8 + 1, 256)), dim3(256), 0, (cudaStream_t)stream, in, (bf16*)out, n);
  return check_launch("cast_f32_bf16_kernel");
}

}  // extern "C"
