// Second generation of the per-head QK RMSNorm + 2-D axial RoPE kernels (reference Attention.py:61-64,
// 130-134,174-194; rotary_embedding.py:36-76,269-288).  Same math as the kernels in elementwise.cu;
// what changes is everything around the math, which is what bounded them (~20 issued instructions per
// element against a budget of 19 at the HBM rate):
//   * a thread keeps ONE 8-column slot of q and of k for its whole life (thread = column slot x row
//     phase), so there is no 64-bit index division per vector and the norm weights of the slot sit in
//     registers instead of being re-read for every row;
//   * a block owns a strip of tokens of one sample: the RoPE table row is tok, no modulo;
//   * the cos / sin pairs of a slot are two 128-bit loads, issued with the row's q / k vectors one row
//     ahead of the math (register ping-pong), so every thread has two rows in flight;
//   * packed fp32x2 math for the scaling; the row reductions and the rotation keep the first
//     generation's operation order.
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"
#include "mmdit_b200.h"

namespace mmdit {

constexpr int QK2_FWD_MAX_THREADS = 384;
constexpr int QK2_BWD_MAX_THREADS = 256;
// Rows in flight per thread + 1.  Measured at cfg2 (16384 x 768, RoPE): forward 2 stages 24.4 us, 4 stages
// (112 registers) 34.5 us; backward from the fp32 dq accumulator 39.1 us with 2, 38.6 us with 3 (168
// registers): deeper rings cost registers and code size and buy nothing.
constexpr int QK2_FWD_STAGES = 2;
constexpr int QK2_BWD_STAGES = 2;

__device__ __forceinline__ float head_sum(float v) {   // over the 8 adjacent lanes that hold one head
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
__device__ __forceinline__ float sumsq8(const float2 (&v)[4]) {
  float ss = 0.f;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    ss = __fmaf_rn(v[p].x, v[p].x, ss);
    ss = __fmaf_rn(v[p].y, v[p].y, ss);
  }
  return ss;
}
__device__ __forceinline__ void unpack8(const uint4& u, float2 (&v)[4]) {
  v[0] = bf2_unpack(u.x); v[1] = bf2_unpack(u.y); v[2] = bf2_unpack(u.z); v[3] = bf2_unpack(u.w);
}
__device__ __forceinline__ void load_w8(const float* __restrict__ w, int pos, float2 (&o)[4]) {
#pragma unroll
  for (int p = 0; p < 4; ++p) o[p] = make_float2(w[pos + 2 * p], w[pos + 2 * p + 1]);
}

// body(stage, it) for it = 0 .. iters-1 with the loads of iteration it + S - 1 issued before the math of
// iteration it: every thread keeps S - 1 rows in flight.  `iters` is uniform over the block.
template <int S, typename Stage, typename Load, typename Body>
__device__ __forceinline__ void stage_ring(int iters, Load&& load, Body&& body) {
  Stage st[S];
#pragma unroll
  for (int k = 0; k < S - 1; ++k)
    if (k < iters) load(st[k], k);
  for (int it = 0; it < iters; it += S) {
#pragma unroll
    for (int k = 0; k < S; ++k) {
      if (it + k < iters) {
        if (it + k + S - 1 < iters) load(st[(k + S - 1) % S], it + k + S - 1);
        body(st[k], it + k);
      }
    }
  }
}

struct QkStrip {
  int gx, ry, col, pos;
  long long row0;     // first row (b * T + tok0) of the strip
  int tok0, ntok;     // tokens tok0 .. tok0 + ntok - 1 of the sample
  int iters;          // uniform over the block
};
__device__ __forceinline__ QkStrip qk_strip(int G, int RY, int T, int rows_per_block, int blocks_per_sample) {
  QkStrip s;
  s.gx = threadIdx.x % G;
  s.ry = threadIdx.x / G;
  s.col = s.gx * 8;
  s.pos = s.col & 63;
  const int b = blockIdx.x / blocks_per_sample;
  const int chunk = blockIdx.x % blocks_per_sample;
  s.tok0 = chunk * rows_per_block;
  s.ntok = min(rows_per_block, T - s.tok0);
  s.row0 = (long long)b * T + s.tok0;
  s.iters = (rows_per_block + RY - 1) / RY;
  return s;
}

// -------------------------------------------------------------------- forward
struct QkFwdStage {
  uint4 q, k;
  float4 cs, sn;
};

template <bool ROPE>
__global__ void __launch_bounds__(QK2_FWD_MAX_THREADS)
qknorm_rope_fwd2_kernel(const bf16* __restrict__ qkv, const float* __restrict__ wq,
                        const float* __restrict__ wk, const float* __restrict__ rope_cos,
                        const float* __restrict__ rope_sin, bf16* __restrict__ out, int d,
                        long long ld_in, long long ld_out, int T, float eps, int G, int RY,
                        int rows_per_block, int blocks_per_sample) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  const QkStrip s = qk_strip(G, RY, T, rows_per_block, blocks_per_sample);
  float2 wq2[4], wk2[4];
  load_w8(wq, s.pos, wq2);
  load_w8(wk, s.pos, wk2);

  auto load = [&](QkFwdStage& t, int it) {
    const int i = it * RY + s.ry;
    if (i < s.ntok) {
      const bf16* p = qkv + (s.row0 + i) * ld_in + s.col;
      t.q = *reinterpret_cast<const uint4*>(p);
      t.k = *reinterpret_cast<const uint4*>(p + d);
      if (ROPE) {
        const long long off = (long long)(s.tok0 + i) * 32 + (s.pos >> 1);
        t.cs = *reinterpret_cast<const float4*>(rope_cos + off);
        t.sn = *reinterpret_cast<const float4*>(rope_sin + off);
      }
    } else {
      t.q = t.k = make_uint4(0u, 0u, 0u, 0u);
      t.cs = t.sn = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  // RMS-normalise the slot's 8 columns, scale by the norm weight, round to bf16 (the reference rounds
  // the RMSNorm output before the fp32 rotation), rotate, round again
  auto norm_rot = [&](const uint4& raw, const float2 (&w)[4], const QkFwdStage& t, uint32_t (&o)[4]) {
    float2 v[4];
    unpack8(raw, v);
    const float r = rsqrtf(__fmaf_rn(head_sum(sumsq8(v)), 1.f / 64.f, eps));
    const float2 r2 = f2_dup(r);
#pragma unroll
    for (int p = 0; p < 4; ++p) o[p] = bf2_pack(f2_mul(f2_mul(v[p], r2), w[p]));
    if (ROPE) {
      const float c[4] = {t.cs.x, t.cs.y, t.cs.z, t.cs.w}, sn[4] = {t.sn.x, t.sn.y, t.sn.z, t.sn.w};
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float2 n = bf2_unpack(o[p]);
        const float x0 = __fmaf_rn(n.x, c[p], -__fmul_rn(n.y, sn[p]));
        const float x1 = __fmaf_rn(n.x, sn[p], __fmul_rn(n.y, c[p]));
        o[p] = pack_bf16x2(x0, x1);
      }
    }
  };
  auto body = [&](const QkFwdStage& t, int it) {
    uint32_t oq[4], ok[4];
    norm_rot(t.q, wq2, t, oq);
    norm_rot(t.k, wk2, t, ok);
    const int i = it * RY + s.ry;
    if (i < s.ntok) {
      bf16* p = out + (s.row0 + i) * ld_out + s.col;
      *reinterpret_cast<uint4*>(p) = make_uint4(oq[0], oq[1], oq[2], oq[3]);
      *reinterpret_cast<uint4*>(p + d) = make_uint4(ok[0], ok[1], ok[2], ok[3]);
    }
  };

  stage_ring<QK2_FWD_STAGES, QkFwdStage>(s.iters, load, body);
}

// ------------------------------------------------------------------- backward
// dqk = grad wrt the normalised / rotated q, k; writes grad wrt raw q, k; accumulates dwq / dwk.
// DQ_F32: the q half of the incoming gradient comes from the attention backward's fp32 accumulator
// dq_acc [B, acc_tokens, d] (rounded to bf16 first, exactly as the bf16 copy was).
template <bool DQ_F32>
struct QkBwdStage {
  uint4 q, k, gk;
  uint4 gq;           // bf16 dq            (!DQ_F32)
  float4 gq0, gq1;    // fp32 dq_acc        (DQ_F32)
  float4 cs, sn;
};

template <bool ROPE, bool DQ_F32>
__global__ void __launch_bounds__(QK2_BWD_MAX_THREADS, 2)
qknorm_rope_bwd2_kernel(const float* __restrict__ dq_acc, int acc_tokens, int acc_tok_off,
                        const bf16* __restrict__ dqk, const bf16* __restrict__ qkv,
                        const float* __restrict__ wq, const float* __restrict__ wk,
                        const float* __restrict__ rope_cos, const float* __restrict__ rope_sin,
                        bf16* __restrict__ dqkv, float* __restrict__ dwq, float* __restrict__ dwk, int d,
                        long long ld_g, long long ld_in, long long ld_dout, int T, float eps, int G,
                        int RY, int rows_per_block, int blocks_per_sample) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  __shared__ float red[2][64];
  if (threadIdx.x < 128) (&red[0][0])[threadIdx.x] = 0.f;
  __syncthreads();
  const QkStrip s = qk_strip(G, RY, T, rows_per_block, blocks_per_sample);
  const int b = blockIdx.x / blocks_per_sample;
  float2 wq2[4], wk2[4], awq[4], awk[4];
  load_w8(wq, s.pos, wq2);
  load_w8(wk, s.pos, wk2);
#pragma unroll
  for (int p = 0; p < 4; ++p) awq[p] = awk[p] = make_float2(0.f, 0.f);

  auto load = [&](QkBwdStage<DQ_F32>& t, int it) {
    const int i = it * RY + s.ry;
    if (i < s.ntok) {
      const bf16* p = qkv + (s.row0 + i) * ld_in + s.col;
      t.q = *reinterpret_cast<const uint4*>(p);
      t.k = *reinterpret_cast<const uint4*>(p + d);
      const bf16* g = dqk + (s.row0 + i) * ld_g + s.col;
      t.gk = *reinterpret_cast<const uint4*>(g + d);
      if constexpr (DQ_F32) {
        const float* a = dq_acc + ((long long)b * acc_tokens + acc_tok_off + s.tok0 + i) * d + s.col;
        t.gq0 = *reinterpret_cast<const float4*>(a);
        t.gq1 = *reinterpret_cast<const float4*>(a + 4);
      } else {
        t.gq = *reinterpret_cast<const uint4*>(g);
      }
      if (ROPE) {
        const long long off = (long long)(s.tok0 + i) * 32 + (s.pos >> 1);
        t.cs = *reinterpret_cast<const float4*>(rope_cos + off);
        t.sn = *reinterpret_cast<const float4*>(rope_sin + off);
      }
    } else {
      t.q = t.k = t.gk = t.gq = make_uint4(0u, 0u, 0u, 0u);
      t.gq0 = t.gq1 = t.cs = t.sn = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  // one of q / k: g = incoming gradient pairs (rotated back in place), returns the packed output words
  auto half = [&](const uint4& raw, float2 (&g)[4], const float2 (&w)[4], float2 (&aw)[4],
                  const QkBwdStage<DQ_F32>& t, uint32_t (&o)[4]) {
    if (ROPE) {   // rotate the incoming gradient by -theta
      const float c[4] = {t.cs.x, t.cs.y, t.cs.z, t.cs.w}, sn[4] = {t.sn.x, t.sn.y, t.sn.z, t.sn.w};
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float a0 = g[p].x, a1 = g[p].y;
        g[p].x = __fmaf_rn(a0, c[p], __fmul_rn(a1, sn[p]));
        g[p].y = __fmaf_rn(a1, c[p], -__fmul_rn(a0, sn[p]));
      }
    }
    float2 v[4];
    unpack8(raw, v);
    const float r = rsqrtf(__fmaf_rn(head_sum(sumsq8(v)), 1.f / 64.f, eps));
    const float2 r2 = f2_dup(r);
    float m = 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      v[p] = f2_mul(v[p], r2);                 // normalised value
      aw[p] = f2_fma(v[p], g[p], aw[p]);       // dw += g * xhat
      g[p] = f2_mul(g[p], w[p]);
      m = __fmaf_rn(v[p].x, g[p].x, m);
      m = __fmaf_rn(v[p].y, g[p].y, m);
    }
    m = head_sum(m) * (1.f / 64.f);
    const float2 nm = f2_dup(-m);
#pragma unroll
    for (int p = 0; p < 4; ++p) o[p] = bf2_pack(f2_mul(r2, f2_fma(v[p], nm, g[p])));
  };
  auto body = [&](const QkBwdStage<DQ_F32>& t, int it) {
    float2 gq[4], gk[4];
    if constexpr (DQ_F32) {
      gq[0] = bf2_unpack(pack_bf16x2(t.gq0.x, t.gq0.y));
      gq[1] = bf2_unpack(pack_bf16x2(t.gq0.z, t.gq0.w));
      gq[2] = bf2_unpack(pack_bf16x2(t.gq1.x, t.gq1.y));
      gq[3] = bf2_unpack(pack_bf16x2(t.gq1.z, t.gq1.w));
    } else {
      unpack8(t.gq, gq);
    }
    unpack8(t.gk, gk);
    uint32_t oq[4], ok[4];
    half(t.q, gq, wq2, awq, t, oq);
    half(t.k, gk, wk2, awk, t, ok);
    const int i = it * RY + s.ry;
    if (i < s.ntok) {
      bf16* p = dqkv + (s.row0 + i) * ld_dout + s.col;
      *reinterpret_cast<uint4*>(p) = make_uint4(oq[0], oq[1], oq[2], oq[3]);
      *reinterpret_cast<uint4*>(p + d) = make_uint4(ok[0], ok[1], ok[2], ok[3]);
    }
  };

  stage_ring<QK2_BWD_STAGES, QkBwdStage<DQ_F32>>(s.iters, load, body);

  // dwq / dwk: lanes l, l^8, l^16, l^24 hold the same slot of different heads (G is a multiple of 8)
  const int slot = (threadIdx.x & 7) * 8;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    float v4[4] = {awq[p].x, awq[p].y, awk[p].x, awk[p].y};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v4[e] += __shfl_xor_sync(0xffffffffu, v4[e], 8);
      v4[e] += __shfl_xor_sync(0xffffffffu, v4[e], 16);
    }
    if ((threadIdx.x & 31) < 8) {
      atomicAdd(&red[0][slot + 2 * p], v4[0]);
      atomicAdd(&red[0][slot + 2 * p + 1], v4[1]);
      atomicAdd(&red[1][slot + 2 * p], v4[2]);
      atomicAdd(&red[1][slot + 2 * p + 1], v4[3]);
    }
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    atomicAdd(dwq + threadIdx.x, red[0][threadIdx.x]);
    atomicAdd(dwk + threadIdx.x, red[1][threadIdx.x]);
  }
}

// ------------------------------------------------------------------ host side
struct QkPlan {
  int G, RY, NT, rows_per_block, blocks_per_sample, B;
};
static int gcd_i(int a, int b) { return b ? gcd_i(b, a % b) : a; }

// threads = column slots (d / 8) x row phases; a multiple of 32 so that every warp is whole
static bool qk_plan(QkPlan& p, const void* kernel, long long rows, int d, int T, int target_threads,
                    int max_threads, int max_blocks_per_sm) {
  if (rows % T != 0 || rows / T > 0x7fffffffLL / 2) return false;
  p.B = (int)(rows / T);
  p.G = d / 8;
  const int m = 32 / gcd_i(p.G, 32);
  p.RY = (target_threads + p.G - 1) / p.G;
  p.RY = (p.RY + m - 1) / m * m;
  p.NT = p.G * p.RY;
  if (p.NT > max_threads) return false;
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, p.NT, 0) != cudaSuccess || occ < 1) {
    cudaGetLastError();
    occ = 1;
  }
  if (max_blocks_per_sm > 0 && occ > max_blocks_per_sm) occ = max_blocks_per_sm;
  long long per_sample = (long long)occ * num_sms() / p.B;   // one wave, split evenly over the samples
  if (per_sample < 1) per_sample = 1;
  long long rpb = (T + per_sample - 1) / per_sample;
  rpb = (rpb + p.RY - 1) / p.RY * p.RY;
  p.rows_per_block = (int)rpb;
  p.blocks_per_sample = (int)((T + rpb - 1) / rpb);
  return true;
}
static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

int qknorm_rope_fwd_v2(const void* qkv, const float* wq, const float* wk, const float* rope_cos,
                       const float* rope_sin, void* out, long long rows, int d, long long ld_in,
                       long long ld_out, int T, float eps, cudaStream_t stream) {
  if (!aligned16(rope_cos) || !aligned16(rope_sin)) return ROW_V2_UNSUPPORTED;
  QkPlan p;
  const void* k = rope_cos ? (const void*)qknorm_rope_fwd2_kernel<true> : (const void*)qknorm_rope_fwd2_kernel<false>;
  if (!qk_plan(p, k, rows, d, T, 256, QK2_FWD_MAX_THREADS, 0)) return ROW_V2_UNSUPPORTED;
  const dim3 grid((unsigned)(p.B * p.blocks_per_sample)), block((unsigned)p.NT);
  if (rope_cos)
    launch_k(qknorm_rope_fwd2_kernel<true>, grid, block, 0, stream, (const bf16*)qkv, wq, wk, rope_cos, rope_sin,
             (bf16*)out, d, ld_in, ld_out, T, eps, p.G, p.RY, p.rows_per_block, p.blocks_per_sample);
  else
    launch_k(qknorm_rope_fwd2_kernel<false>, grid, block, 0, stream, (const bf16*)qkv, wq, wk, rope_cos, rope_sin,
             (bf16*)out, d, ld_in, ld_out, T, eps, p.G, p.RY, p.rows_per_block, p.blocks_per_sample);
  return check_launch("qknorm_rope_fwd2_kernel");
}

static int qkn2_cap() {   // resident blocks per SM of the backward grid: every block ends with 128 global atomics
  static const int v = [] {
    const char* e = getenv("MMDIT_QKN2_CAP");
    const int x = e ? atoi(e) : 2;
    return x >= 1 && x <= 16 ? x : 2;
  }();
  return v;
}

int qknorm_rope_bwd_v2(const float* dq_acc, int acc_tokens, int acc_tok_off, const void* dqk, const void* qkv,
                       const float* wq, const float* wk, const float* rope_cos, const float* rope_sin,
                       void* dqkv, float* dwq, float* dwk, long long rows, int d, long long ld_g,
                       long long ld_in, long long ld_dout, int T, float eps, cudaStream_t stream) {
  if (!aligned16(rope_cos) || !aligned16(rope_sin) || !aligned16(dq_acc) || (dq_acc && d % 4 != 0))
    return ROW_V2_UNSUPPORTED;
  QkPlan p;
#define QKB_CALL(ROPEV, ACCV)                                                                               \
  do {                                                                                                      \
    if (!qk_plan(p, (const void*)qknorm_rope_bwd2_kernel<ROPEV, ACCV>, rows, d, T, 192, QK2_BWD_MAX_THREADS, \
                 qkn2_cap()))                                                                               \
      return ROW_V2_UNSUPPORTED;                                                                            \
    launch_k(qknorm_rope_bwd2_kernel<ROPEV, ACCV>, dim3((unsigned)(p.B * p.blocks_per_sample)),             \
             dim3((unsigned)p.NT), 0, stream, dq_acc, acc_tokens, acc_tok_off, (const bf16*)dqk,            \
             (const bf16*)qkv, wq, wk, rope_cos, rope_sin, (bf16*)dqkv, dwq, dwk, d, ld_g, ld_in, ld_dout,  \
             T, eps, p.G, p.RY, p.rows_per_block, p.blocks_per_sample);                                     \
  } while (0)
  if (rope_cos) {
    if (dq_acc) QKB_CALL(true, true); else QKB_CALL(true, false);
  } else {
    if (dq_acc) QKB_CALL(false, true); else QKB_CALL(false, false);
  }
#undef QKB_CALL
  return check_launch("qknorm_rope_bwd2_kernel");
}

}  // namespace mmdit
