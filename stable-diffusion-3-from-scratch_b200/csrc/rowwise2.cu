// Second generation of the LayerNorm-modulate / gate row kernels (reference Norm.py:16-23,
// Transformer_Block_Dual.py:64-76).  Same math as rowwise.cu, restructured around what the
// first generation was bound by: ~28 issued instructions per element and one row of loads per
// warp in flight (in-kernel timelines, profiles/r02_experiments_not_shipped.md).
//   * a block owns a strip of rows of ONE sample, so the per-sample vectors (1 + scale, shift, gate)
//     are expanded to fp32 once per block into shared memory (lane-major: every LDS.128 of a warp is
//     conflict-free) instead of being re-read, unpacked and re-rounded for every row;
//   * rows stay packed (bf16x2 words) in registers until they are used, which leaves room for the
//     NEXT row of every warp to be in flight while the current one is computed (register ping-pong);
//   * the elementwise math is packed fp32x2 (FADD2 / FMUL2 / FFMA2: one issue slot, two columns).
//     Row statistics keep the first generation's sequential summation order: y, x', mean, rstd, dx and
//     da are bit-identical to the first-generation kernels (tools/row_probe.py check, 0 differing
//     elements at every shape); the per-sample column sums differ in the last fp32 bits because the
//     strips are cut differently.
// Measured (B200, graph replay over buffers larger than the L2, cfg2 image stream 16384 x 768, of the
// 6549.8 GB/s copy rate): LN-modulate fwd 13.9 -> 9.9 us (0.55 -> 0.77), gate+residual+LN fwd 23.5 ->
// 17.1 us (0.65 -> 0.90), gate bwd 18.9 -> 17.1 us; at d = 1536 LN bwd 57.8 -> 45.4 us, gate bwd 39.4 ->
// 32.5 us, gate+LN fwd 107 -> 38 us (profiles/r02_row_kernels_gen2.log).
// One warp owns one row; lane l holds columns c*256 + l*8 .. +7 of every 256-column chunk c.
#include <stdlib.h>

#include "common.cuh"
#include "mmdit_b200.h"

namespace mmdit {

constexpr int R2_THREADS = 128;
constexpr int R2_WARPS = R2_THREADS / 32;

// ---- per-sample vector tables -----------------------------------------------------------------
// Entry (c, h, lane) holds columns c*256 + lane*8 + h*4 .. +3: the 32 lanes of a warp read 32
// consecutive float4's.  A table is NC*256 floats; columns >= d hold zeros.
__device__ __forceinline__ int tab_off(int c, int h, int lane) { return ((c * 2 + h) * 32 + lane) * 4; }

template <int NC, bool FULL, bool ONE_PLUS>
__device__ __forceinline__ void fill_table(float* tab, const bf16* __restrict__ vec, int d, int warp,
                                           int lane) {
  for (int c = warp; c < NC; c += R2_WARPS) {
    const int col = c * 256 + lane * 8;
    float f[8];
    if (FULL || col < d) {
      load8(vec + col, f);
      if (ONE_PLUS) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = __bfloat162float(__float2bfloat16(1.f + f[j]));   // the reference adds 1 in bf16
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = 0.f;
    }
    *reinterpret_cast<float4*>(tab + tab_off(c, 0, lane)) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(tab + tab_off(c, 1, lane)) = make_float4(f[4], f[5], f[6], f[7]);
  }
}
// the four column pairs of chunk c held by this lane
// (volatile: the table is loop-invariant, and a compiler that hoists these loads out of the row loop
// turns the table back into 8*NC live registers per vector -- the spills this layout exists to avoid)
__device__ __forceinline__ void tab_read(const float* tab, int c, int lane, float2 (&t)[4]) {
  const uint32_t addr = smem_u32(tab + tab_off(c, 0, lane));
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(t[0].x), "=f"(t[0].y), "=f"(t[1].x), "=f"(t[1].y) : "r"(addr));
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4 + 512];"
               : "=f"(t[2].x), "=f"(t[2].y), "=f"(t[3].x), "=f"(t[3].y) : "r"(addr));
}

template <int NC, bool FULL>
__device__ __forceinline__ void load_raw(uint4 (&r)[NC], const bf16* __restrict__ p, int d, int lane) {
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = c * 256 + lane * 8;
    r[c] = (FULL || col < d) ? *reinterpret_cast<const uint4*>(p + col) : make_uint4(0u, 0u, 0u, 0u);
  }
}
template <int NC>
__device__ __forceinline__ void zero_raw(uint4 (&r)[NC]) {
#pragma unroll
  for (int c = 0; c < NC; ++c) r[c] = make_uint4(0u, 0u, 0u, 0u);
}

// warps -> shared [warp][2][d] -> one [2][d] partial per block (same layout and order as rowwise.cu)
template <int NC>
__device__ __forceinline__ void block_fold2(float* red, const float2 (&a0)[NC][4], const float2 (&a1)[NC][4],
                                            float* out, int d, int warp, int lane) {
  float* mine = red + (long long)warp * 2 * d;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = c * 256 + lane * 8;
    if (col < d) {
      *reinterpret_cast<float4*>(mine + col) = make_float4(a0[c][0].x, a0[c][0].y, a0[c][1].x, a0[c][1].y);
      *reinterpret_cast<float4*>(mine + col + 4) = make_float4(a0[c][2].x, a0[c][2].y, a0[c][3].x, a0[c][3].y);
      *reinterpret_cast<float4*>(mine + d + col) = make_float4(a1[c][0].x, a1[c][0].y, a1[c][1].x, a1[c][1].y);
      *reinterpret_cast<float4*>(mine + d + col + 4) = make_float4(a1[c][2].x, a1[c][2].y, a1[c][3].x, a1[c][3].y);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * d; i += R2_THREADS) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < R2_WARPS; ++w) acc += red[(long long)w * 2 * d + i];
    out[i] = acc;
  }
}

// Column sums of a whole sample without a second launch: the blocks of one sample are ONE thread-block
// cluster (2..8 CTAs).  Every CTA reduces its warps in its own shared memory, the cluster synchronises,
// and CTA r adds slice r of the 2*d columns over all CTAs through distributed shared memory
// (ld.shared::cluster), in rank order -- the order fold_batch_partials_kernel uses over the workspace --
// and writes the final bf16 / fp32 vectors.  No workspace traffic, no fold kernel (94 launches of ~4 us
// per cfg2 step).  out0 / out1 as in fold_batch_partials_kernel.
__device__ __forceinline__ float ld_dsmem_f32(uint32_t cluster_saddr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(cluster_saddr));
  return v;
}
template <int NC>
__device__ __forceinline__ void cluster_fold2(float* red, const float2 (&a0)[NC][4], const float2 (&a1)[NC][4],
                                              int d, int warp, int lane, int csize, long long b, void* out0,
                                              void* out1, long long ld0, long long ld1, int out0_bf16,
                                              int out1_bf16) {
  float* mine = red + (long long)warp * 2 * d;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = c * 256 + lane * 8;
    if (col < d) {
      *reinterpret_cast<float4*>(mine + col) = make_float4(a0[c][0].x, a0[c][0].y, a0[c][1].x, a0[c][1].y);
      *reinterpret_cast<float4*>(mine + col + 4) = make_float4(a0[c][2].x, a0[c][2].y, a0[c][3].x, a0[c][3].y);
      *reinterpret_cast<float4*>(mine + d + col) = make_float4(a1[c][0].x, a1[c][0].y, a1[c][1].x, a1[c][1].y);
      *reinterpret_cast<float4*>(mine + d + col + 4) = make_float4(a1[c][2].x, a1[c][2].y, a1[c][3].x, a1[c][3].y);
    }
  }
  __syncthreads();
  const int n2 = 2 * d;
  for (int i = threadIdx.x; i < n2; i += R2_THREADS) {   // in place: column i is touched by one thread only
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < R2_WARPS; ++w) acc += red[(long long)w * n2 + i];
    red[i] = acc;
  }
  cluster_sync_all();   // every CTA's block sums are in its shared memory and visible to the cluster
  const int rank = (int)cluster_ctarank();
  const int per = (n2 + csize - 1) / csize;
  const int i1 = min(n2, (rank + 1) * per);
  for (int i = rank * per + threadIdx.x; i < i1; i += R2_THREADS) {
    const uint32_t local = smem_u32(red + i);
    float acc = 0.f;
    for (int k = 0; k < csize; ++k) acc += ld_dsmem_f32(mapa_u32(local, (uint32_t)k));
    if (i < d) {
      if (out0_bf16) reinterpret_cast<bf16*>(out0)[b * ld0 + i] = __float2bfloat16(acc);
      else reinterpret_cast<float*>(out0)[b * ld0 + i] = acc;
    } else if (out1) {
      if (out1_bf16) reinterpret_cast<bf16*>(out1)[b * ld1 + (i - d)] = __float2bfloat16(acc);
      else reinterpret_cast<float*>(out1)[b * ld1 + (i - d)] = acc;
    }
  }
  cluster_sync_all();   // no CTA leaves while a peer may still read its shared memory
}

// Runs body(stage, row) over rows r0 + warp, r0 + warp + R2_WARPS, ... < r1 with the loads of the
// next row issued before the current row is processed.  `load(stage, row)` fills stage 0 or 1.
template <typename Load, typename Body>
__device__ __forceinline__ void pingpong_rows(long long r0, long long r1, int warp, Load&& load, Body&& body) {
  long long row = r0 + warp;
  if (row >= r1) return;
  load(0, row);
  while (true) {
    long long nrow = row + R2_WARPS;
    if (nrow < r1) load(1, nrow);
    body(0, row);
    row = nrow;
    if (row >= r1) break;
    nrow = row + R2_WARPS;
    if (nrow < r1) load(0, nrow);
    body(1, row);
    row = nrow;
    if (row >= r1) break;
  }
}

// ------------------------------------------------------------ LN-modulate bwd
// g = dy*(1+s); dx = rstd*(g - mean(g) - xhat*mean(g*xhat)) (+ dres)
// partial[block] = { sum_rows dy , sum_rows dy*xhat }   (folded per sample by fold_batch_partials_kernel)
template <int NC>
struct LnBwdRow {
  uint4 dy[NC], x[NC];
  float mean, rstd;
};

template <int NC, bool FULL, bool PREFETCH>
__global__ void __launch_bounds__(R2_THREADS, PREFETCH ? 3 : 2)
ln_mod_bwd2_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                   const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                   const bf16* __restrict__ scale, const bf16* __restrict__ dres,
                   bf16* __restrict__ dx, float* __restrict__ partial, int d, long long rows_per_batch,
                   long long ld_mod, void* __restrict__ out0, void* __restrict__ out1, long long ld_out,
                   int out_bf16, int cluster_fold, int rows_per_block, int blocks_per_batch) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  extern __shared__ float red[];  // (1 + scale) table during the row loop, then [R2_WARPS][2][d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = blockIdx.x / blocks_per_batch;
  const int chunk = blockIdx.x % blocks_per_batch;
  const long long r0 = b * rows_per_batch + (long long)chunk * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > (b + 1) * rows_per_batch) r1 = (b + 1) * rows_per_batch;

  const float* tab = red;
  fill_table<NC, FULL, true>(red, scale + b * ld_mod, d, warp, lane);
  __syncthreads();

  const bool has_res = dres != nullptr;
  float2 a_sh[NC][4], a_sc[NC][4];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int p = 0; p < 4; ++p) a_sh[c][p] = a_sc[c][p] = make_float2(0.f, 0.f);

  auto load = [&](LnBwdRow<NC>& t, long long row) {
    load_raw<NC, FULL>(t.dy, dy + row * d, d, lane);
    load_raw<NC, FULL>(t.x, x + row * d, d, lane);
    t.mean = mean_in[row];
    t.rstd = rstd_in[row];
  };
  auto body = [&](const LnBwdRow<NC>& t, long long row) {
    // the residual-path gradient is only needed after the row reductions: fetched here, not a row
    // ahead (12 more live registers per stage would spill at three blocks per SM)
    uint4 dr[NC];
    if (has_res) load_raw<NC, FULL>(dr, dres + row * d, d, lane);
    else zero_raw<NC>(dr);
    const float2 nmean = f2_dup(-t.mean), rs = f2_dup(t.rstd);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float2 op[4];
      tab_read(tab, c, lane, op);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float2 dyv = bf2_unpack(word(t.dy[c], p));
        const float2 xh = f2_mul(f2_add(bf2_unpack(word(t.x[c], p)), nmean), rs);
        a_sh[c][p] = f2_add(a_sh[c][p], dyv);
        a_sc[c][p] = f2_fma(dyv, xh, a_sc[c][p]);
        const float2 gv = f2_mul(dyv, op[p]);
        sg += gv.x;                          // sequential, the first generation's order
        sgx = __fmaf_rn(gv.x, xh.x, sgx);
        sg += gv.y;
        sgx = __fmaf_rn(gv.y, xh.y, sgx);
      }
    }
    const float mg = warp_sum(sg) / d, mgx = warp_sum(sgx) / d;
    const float2 nmg = f2_dup(-mg), nmgx = f2_dup(-mgx);
    bf16* orow = dx + row * d;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int col = c * 256 + lane * 8;
      if (FULL || col < d) {
        float2 op[4];
        tab_read(tab, c, lane, op);
        uint32_t o[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float2 dyv = bf2_unpack(word(t.dy[c], p));
          const float2 xh = f2_mul(f2_add(bf2_unpack(word(t.x[c], p)), nmean), rs);
          float2 u = f2_add(f2_mul(dyv, op[p]), nmg);   // g - mean(g)
          u = f2_fma(xh, nmgx, u);                      //   - xhat * mean(g * xhat)
          o[p] = bf2_pack(f2_fma(rs, u, bf2_unpack(word(dr[c], p))));
        }
        *reinterpret_cast<uint4*>(orow + col) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  };

  if constexpr (PREFETCH) {
    LnBwdRow<NC> st[2];
    pingpong_rows(r0, r1, warp, [&](int s, long long row) { load(st[s], row); },
                  [&](int s, long long row) { body(st[s], row); });
  } else {
    for (long long row = r0 + warp; row < r1; row += R2_WARPS) {
      LnBwdRow<NC> t;
      load(t, row);
      body(t, row);
    }
  }
  __syncthreads();   // every warp is done with the table before the fold overwrites it
  if (cluster_fold)
    cluster_fold2<NC>(red, a_sh, a_sc, d, warp, lane, blocks_per_batch, b, out0, out1, ld_out, ld_out, out_bf16,
                      out_bf16);
  else
    block_fold2<NC>(red, a_sh, a_sc, partial + (long long)blockIdx.x * 2 * d, d, warp, lane);
}

// ------------------------------------------------------------------ gate bwd
// forward was o = a*g[b] + x.  da = do*g[b]; partial[block] = { sum_rows do*a , sum_rows da }
template <int NC>
struct GateBwdRow {
  uint4 dv[NC], av[NC];
};

template <int NC, bool FULL, bool PREFETCH>
__global__ void __launch_bounds__(R2_THREADS, PREFETCH ? 3 : 2)
gate_bwd2_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ a,
                 const bf16* __restrict__ gate, bf16* __restrict__ da, float* __restrict__ partial,
                 int d, long long rows_per_batch, long long ld_gate, void* __restrict__ out0,
                 void* __restrict__ out1, long long ld0, long long ld1, int out0_bf16, int cluster_fold,
                 int rows_per_block, int blocks_per_batch) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  extern __shared__ float red[];  // gate table during the row loop, then [R2_WARPS][2][d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = blockIdx.x / blocks_per_batch;
  const int chunk = blockIdx.x % blocks_per_batch;
  const long long r0 = b * rows_per_batch + (long long)chunk * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > (b + 1) * rows_per_batch) r1 = (b + 1) * rows_per_batch;

  const float* tab = red;
  fill_table<NC, FULL, false>(red, gate + b * ld_gate, d, warp, lane);
  __syncthreads();

  float2 a_g[NC][4], a_b[NC][4];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int p = 0; p < 4; ++p) a_g[c][p] = a_b[c][p] = make_float2(0.f, 0.f);

  auto load = [&](GateBwdRow<NC>& t, long long row) {
    load_raw<NC, FULL>(t.dv, dout + row * d, d, lane);
    load_raw<NC, FULL>(t.av, a + row * d, d, lane);
  };
  auto body = [&](const GateBwdRow<NC>& t, long long row) {
    bf16* orow = da + row * d;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int col = c * 256 + lane * 8;
      float2 g[4];
      tab_read(tab, c, lane, g);
      uint32_t o[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float2 dv = bf2_unpack(word(t.dv[c], p));
        a_g[c][p] = f2_fma(dv, bf2_unpack(word(t.av[c], p)), a_g[c][p]);
        const float2 ov = f2_mul(dv, g[p]);
        a_b[c][p] = f2_add(a_b[c][p], ov);
        o[p] = bf2_pack(ov);
      }
      if (FULL || col < d) *reinterpret_cast<uint4*>(orow + col) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  };

  if constexpr (PREFETCH) {
    GateBwdRow<NC> st[2];
    pingpong_rows(r0, r1, warp, [&](int s, long long row) { load(st[s], row); },
                  [&](int s, long long row) { body(st[s], row); });
  } else {
    for (long long row = r0 + warp; row < r1; row += R2_WARPS) {
      GateBwdRow<NC> t;
      load(t, row);
      body(t, row);
    }
  }
  __syncthreads();
  if (cluster_fold)
    cluster_fold2<NC>(red, a_g, a_b, d, warp, lane, blocks_per_batch, b, out0, out1, ld0, ld1, out0_bf16, 0);
  else
    block_fold2<NC>(red, a_g, a_b, partial + (long long)blockIdx.x * 2 * d, d, warp, lane);
}

// ------------------------------------------------------------ LN-modulate fwd
// y = LN(x) * bf16(1 + scale[b]) + shift[b].  `v` holds the row as fp32 pairs on entry.
template <int NC, bool FULL>
__device__ __forceinline__ void ln_tail(float2 (&v)[NC][4], const float* tab_scale, const float* tab_shift,
                                        bf16* __restrict__ yrow, float* __restrict__ mean_out,
                                        float* __restrict__ rstd_out, long long row, int d, int lane, float eps) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int p = 0; p < 4; ++p) { s += v[c][p].x; s += v[c][p].y; }
  const float mean = warp_sum(s) / d;
  const float2 nmean = f2_dup(-mean);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = c * 256 + lane * 8;
#pragma unroll
    for (int p = 0; p < 4; ++p) v[c][p] = f2_add(v[c][p], nmean);
    if (FULL || col < d) {
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        q = __fmaf_rn(v[c][p].x, v[c][p].x, q);
        q = __fmaf_rn(v[c][p].y, v[c][p].y, q);
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / d + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  const float2 rs = f2_dup(rstd);
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int col = c * 256 + lane * 8;
    if (FULL || col < d) {
      float2 op[4], sh[4];
      tab_read(tab_scale, c, lane, op);
      tab_read(tab_shift, c, lane, sh);
      uint32_t o[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) o[p] = bf2_pack(f2_fma(f2_mul(v[c][p], rs), op[p], sh[p]));
      *reinterpret_cast<uint4*>(yrow + col) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

template <int NC, bool FULL>
__global__ void __launch_bounds__(R2_THREADS, NC <= 3 ? 5 : 3)
ln_mod_fwd2_kernel(const bf16* __restrict__ x, const bf16* __restrict__ shift,
                   const bf16* __restrict__ scale, bf16* __restrict__ y, float* __restrict__ mean_out,
                   float* __restrict__ rstd_out, int d, long long rows_per_batch, long long ld_mod,
                   float eps, int rows_per_block, int blocks_per_batch) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  extern __shared__ float red[];  // [2][NC*256]: (1 + scale), shift
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = blockIdx.x / blocks_per_batch;
  const int chunk = blockIdx.x % blocks_per_batch;
  const long long r0 = b * rows_per_batch + (long long)chunk * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > (b + 1) * rows_per_batch) r1 = (b + 1) * rows_per_batch;
  float* tab_scale = red;
  float* tab_shift = red + NC * 256;
  fill_table<NC, FULL, true>(tab_scale, scale + b * ld_mod, d, warp, lane);
  fill_table<NC, FULL, false>(tab_shift, shift + b * ld_mod, d, warp, lane);
  __syncthreads();

  uint4 st[2][NC];
  pingpong_rows(
      r0, r1, warp, [&](int s, long long row) { load_raw<NC, FULL>(st[s], x + row * d, d, lane); },
      [&](int s, long long row) {
        float2 v[NC][4];
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int p = 0; p < 4; ++p) v[c][p] = bf2_unpack(word(st[s][c], p));
        ln_tail<NC, FULL>(v, tab_scale, tab_shift, y + row * d, mean_out, rstd_out, row, d, lane, eps);
      });
}

// ------------------------------------------- gated residual + LN-modulate, one pass
//   x' = a * gate[b] + resid   (rounded to bf16, stored);   y = LN(x') * bf16(1 + scale[b]) + shift[b]
template <int NC, bool FULL>
__global__ void __launch_bounds__(R2_THREADS, NC <= 3 ? 4 : 2)
gate_res_ln_fwd2_kernel(const bf16* __restrict__ a, const bf16* __restrict__ gate,
                        const bf16* __restrict__ resid, const bf16* __restrict__ shift,
                        const bf16* __restrict__ scale, bf16* __restrict__ xo, bf16* __restrict__ y,
                        float* __restrict__ mean_out, float* __restrict__ rstd_out, int d,
                        long long rows_per_batch, long long ld_gate, long long ld_mod, float eps,
                        int rows_per_block, int blocks_per_batch) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  extern __shared__ float red[];  // [3][NC*256]: (1 + scale), shift, gate
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = blockIdx.x / blocks_per_batch;
  const int chunk = blockIdx.x % blocks_per_batch;
  const long long r0 = b * rows_per_batch + (long long)chunk * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > (b + 1) * rows_per_batch) r1 = (b + 1) * rows_per_batch;
  float* tab_scale = red;
  float* tab_shift = red + NC * 256;
  float* tab_gate = red + 2 * NC * 256;
  fill_table<NC, FULL, true>(tab_scale, scale + b * ld_mod, d, warp, lane);
  fill_table<NC, FULL, false>(tab_shift, shift + b * ld_mod, d, warp, lane);
  fill_table<NC, FULL, false>(tab_gate, gate + b * ld_gate, d, warp, lane);
  __syncthreads();

  uint4 sa[2][NC], sr[2][NC];
  pingpong_rows(
      r0, r1, warp,
      [&](int s, long long row) {
        load_raw<NC, FULL>(sa[s], a + row * d, d, lane);
        load_raw<NC, FULL>(sr[s], resid + row * d, d, lane);
      },
      [&](int s, long long row) {
        float2 v[NC][4];
        bf16* xrow = xo + row * d;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int col = c * 256 + lane * 8;
          float2 g[4];
          tab_read(tab_gate, c, lane, g);
          uint32_t o[4];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            o[p] = bf2_pack(f2_fma(bf2_unpack(word(sa[s][c], p)), g[p], bf2_unpack(word(sr[s][c], p))));
            v[c][p] = bf2_unpack(o[p]);   // what the LN of the stored x' sees
          }
          if (FULL || col < d) *reinterpret_cast<uint4*>(xrow + col) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        ln_tail<NC, FULL>(v, tab_scale, tab_shift, y + row * d, mean_out, rstd_out, row, d, lane, eps);
      });
}

// ------------------------------------------------------------------ host side
static int g_row_generation = [] {
  const char* e = getenv("MMDIT_ROW_KERNELS");
  const int x = e ? atoi(e) : 2;
  return x == 1 ? 1 : 2;
}();
int row_kernel_generation() { return g_row_generation; }

// one wave of resident blocks, split evenly over the samples (same policy as rowwise.cu)
static int strip_rows(const void* kernel, size_t smem, long long rows_per_batch, int nb, int min_rows) {
  static const int forced = [] {
    const char* e = getenv("MMDIT_ROW_RPB");
    const int x = e ? atoi(e) : 0;
    return x >= 8 && x <= 1024 ? x : 0;
  }();
  if (forced) return forced;
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, R2_THREADS, smem) != cudaSuccess || occ < 1) {
    cudaGetLastError();
    occ = 2;
  }
  long long per_sample = (long long)occ * num_sms() / nb;
  if (per_sample < 1) per_sample = 1;
  long long rpb = (rows_per_batch + per_sample - 1) / per_sample;
  if (rpb < min_rows) rpb = min_rows;
  return (int)rpb;
}

#define R2_DISPATCH(d, KERNEL_FULL, KERNEL_TAIL)                                     \
  do {                                                                               \
    const int nc_ = ((d) + 255) / 256;                                               \
    const bool full_ = (d) % 256 == 0;                                               \
    switch (nc_) {                                                                   \
      case 1: { constexpr int NC = 1; if (full_) { KERNEL_FULL; } else { KERNEL_TAIL; } } break; \
      case 2: { constexpr int NC = 2; if (full_) { KERNEL_FULL; } else { KERNEL_TAIL; } } break; \
      case 3: { constexpr int NC = 3; if (full_) { KERNEL_FULL; } else { KERNEL_TAIL; } } break; \
      case 4: { constexpr int NC = 4; if (full_) { KERNEL_FULL; } else { KERNEL_TAIL; } } break; \
      case 5: { constexpr int NC = 5; if (full_) { KERNEL_FULL; } else { KERNEL_TAIL; } } break; \
      case 6: { constexpr int NC = 6; if (full_) { KERNEL_FULL; } else { KERNEL_TAIL; } } break; \
      default: return ROW_V2_UNSUPPORTED;   /* wider rows stay on the first generation */ \
    }                                                                                \
  } while (0)

template <typename K, typename... Args>
static void launch_strips(K kernel, size_t smem, long long rows_per_batch, int nb, int min_rows,
                          cudaStream_t stream, int* bpb_out, Args... args) {
  if (smem > 48 * 1024) cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int rpb = strip_rows((const void*)kernel, smem, rows_per_batch, nb, min_rows);
  const int bpb = (int)((rows_per_batch + rpb - 1) / rpb);
  if (bpb_out) *bpb_out = bpb;
  launch_k(kernel, dim3((unsigned)(nb * bpb)), dim3(R2_THREADS), smem, stream, args..., rpb, bpb);
}

// MMDIT_ROW_CLUSTER=0: always the workspace + fold-kernel path
static bool cluster_fold_enabled() {
  static const bool on = [] {
    const char* e = getenv("MMDIT_ROW_CLUSTER");
    return e ? atoi(e) != 0 : true;
  }();
  return on;
}
constexpr int MAX_FOLD_CLUSTER = 8;   // portable cluster size

template <typename... KArgs>
static bool launch_cluster(void (*kernel)(KArgs...), unsigned grid, unsigned cluster, size_t smem,
                           cudaStream_t stream, KArgs... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(R2_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  int nclusters = 0;
  if (cudaOccupancyMaxActiveClusters(&nclusters, kernel, &cfg) != cudaSuccess || nclusters < 1) {
    cudaGetLastError();
    return false;   // this cluster shape cannot be scheduled: the caller takes the workspace path
  }
  cudaLaunchKernelEx(&cfg, kernel, args...);   // errors surface in check_launch()
  return true;
}

// Each returns ROW_V2_UNSUPPORTED when the shape is left to the first generation, else the launch status.
int ln_modulate_fwd_v2(const void* x, const void* shift, const void* scale, void* y, float* mean, float* rstd,
                       long long rows, int d, long long rows_per_batch, long long ld_mod, float eps,
                       cudaStream_t stream) {
  if (rows % rows_per_batch != 0 || ld_mod % 8 != 0) return ROW_V2_UNSUPPORTED;
  const int nb = (int)(rows / rows_per_batch);
#define LNF(FULLV)                                                                                     \
  launch_strips(ln_mod_fwd2_kernel<NC, FULLV>, (size_t)2 * NC * 256 * sizeof(float), rows_per_batch, nb, 4, \
                stream, nullptr, (const bf16*)x, (const bf16*)shift, (const bf16*)scale, (bf16*)y, mean,   \
                rstd, d, rows_per_batch, ld_mod, eps)
  R2_DISPATCH(d, LNF(true), LNF(false));
#undef LNF
  return check_launch("ln_mod_fwd2_kernel");
}

int gate_residual_ln_fwd_v2(const void* a, const void* gate, const void* resid, const void* shift,
                            const void* scale, void* x_out, void* y, float* mean, float* rstd, long long rows,
                            int d, long long rows_per_batch, long long ld_gate, long long ld_mod, float eps,
                            cudaStream_t stream) {
  if (rows % rows_per_batch != 0) return ROW_V2_UNSUPPORTED;
  const int nb = (int)(rows / rows_per_batch);
#define GLF(FULLV)                                                                                          \
  launch_strips(gate_res_ln_fwd2_kernel<NC, FULLV>, (size_t)3 * NC * 256 * sizeof(float), rows_per_batch, nb, 4, \
                stream, nullptr, (const bf16*)a, (const bf16*)gate, (const bf16*)resid, (const bf16*)shift,     \
                (const bf16*)scale, (bf16*)x_out, (bf16*)y, mean, rstd, d, rows_per_batch, ld_gate, ld_mod, eps)
  R2_DISPATCH(d, GLF(true), GLF(false));
#undef GLF
  return check_launch("gate_res_ln_fwd2_kernel");
}

// workspace: [nb * bpb][2][d] floats, bpb <= ceil(rows_per_batch / 8) (mmdit_rowreduce_workspace_floats).
// *bpb_out = partials per sample left in the workspace for fold_batch_partials_kernel, or 0 when the
// sample's strips ran as one cluster and wrote out0 / out1 themselves.
template <typename K, typename Tail>
static void launch_bwd(K kernel, size_t smem, long long rows_per_batch, int nb, cudaStream_t stream, int* bpb_out,
                       Tail&& launch_with) {
  const int rpb = strip_rows((const void*)kernel, smem, rows_per_batch, nb, 8);
  const int bpb = (int)((rows_per_batch + rpb - 1) / rpb);
  if (cluster_fold_enabled() && bpb <= MAX_FOLD_CLUSTER && launch_with(true, rpb, bpb)) {
    *bpb_out = 0;
    return;
  }
  launch_with(false, rpb, bpb);
  *bpb_out = bpb;
}

int ln_modulate_bwd_v2(const void* dy, const void* x, const float* mean, const float* rstd, const void* scale,
                       const void* dres, void* dx, void* dshift, void* dscale, int dmod_bf16, long long ld_dmod,
                       float* workspace, long long rows, int d, long long rows_per_batch, long long ld_mod,
                       int* bpb_out, cudaStream_t stream) {
  if (ld_mod % 8 != 0) return ROW_V2_UNSUPPORTED;
  const int nb = (int)(rows / rows_per_batch);
  const int nc = (d + 255) / 256;
  size_t smem = (size_t)R2_WARPS * 2 * d * sizeof(float);
  if (smem < (size_t)nc * 256 * sizeof(float)) smem = (size_t)nc * 256 * sizeof(float);
#define LNB(FULLV)                                                                                          \
  launch_bwd(ln_mod_bwd2_kernel<NC, FULLV, (NC <= 3)>, smem, rows_per_batch, nb, stream, bpb_out,              \
             [&](bool cluster, int rpb, int bpb) {                                                          \
               auto k = ln_mod_bwd2_kernel<NC, FULLV, (NC <= 3)>;                                           \
               if (cluster)                                                                                 \
                 return launch_cluster(k, (unsigned)(nb * bpb), (unsigned)bpb, smem, stream, (const bf16*)dy, \
                                       (const bf16*)x, mean, rstd, (const bf16*)scale, (const bf16*)dres,   \
                                       (bf16*)dx, workspace, d, rows_per_batch, ld_mod, dshift, dscale,     \
                                       ld_dmod, dmod_bf16, 1, rpb, bpb);                                    \
               launch_k(k, dim3((unsigned)(nb * bpb)), dim3(R2_THREADS), smem, stream, (const bf16*)dy,     \
                        (const bf16*)x, mean, rstd, (const bf16*)scale, (const bf16*)dres, (bf16*)dx,       \
                        workspace, d, rows_per_batch, ld_mod, dshift, dscale, ld_dmod, dmod_bf16, 0, rpb,   \
                        bpb);                                                                               \
               return true;                                                                                 \
             })
  R2_DISPATCH(d, LNB(true), LNB(false));
#undef LNB
  return check_launch("ln_mod_bwd2_kernel");
}

int gate_bwd_v2(const void* dout, const void* a, const void* gate, void* da, void* dgate, int dgate_bf16,
                float* dab, long long ld_dgate, long long ld_dab, float* workspace, long long rows, int d,
                long long rows_per_batch, long long ld_gate, int* bpb_out, cudaStream_t stream) {
  if (ld_gate % 8 != 0) return ROW_V2_UNSUPPORTED;
  const int nb = (int)(rows / rows_per_batch);
  const int nc = (d + 255) / 256;
  size_t smem = (size_t)R2_WARPS * 2 * d * sizeof(float);
  if (smem < (size_t)nc * 256 * sizeof(float)) smem = (size_t)nc * 256 * sizeof(float);
#define GB(FULLV)                                                                                           \
  launch_bwd(gate_bwd2_kernel<NC, FULLV, (NC <= 3)>, smem, rows_per_batch, nb, stream, bpb_out,                \
             [&](bool cluster, int rpb, int bpb) {                                                          \
               auto k = gate_bwd2_kernel<NC, FULLV, (NC <= 3)>;                                             \
               if (cluster)                                                                                 \
                 return launch_cluster(k, (unsigned)(nb * bpb), (unsigned)bpb, smem, stream,                \
                                       (const bf16*)dout, (const bf16*)a, (const bf16*)gate, (bf16*)da,     \
                                       workspace, d, rows_per_batch, ld_gate, dgate, (void*)dab, ld_dgate,  \
                                       ld_dab, dgate_bf16, 1, rpb, bpb);                                    \
               launch_k(k, dim3((unsigned)(nb * bpb)), dim3(R2_THREADS), smem, stream, (const bf16*)dout,   \
                        (const bf16*)a, (const bf16*)gate, (bf16*)da, workspace, d, rows_per_batch, ld_gate, \
                        dgate, (void*)dab, ld_dgate, ld_dab, dgate_bf16, 0, rpb, bpb);                      \
               return true;                                                                                 \
             })
  R2_DISPATCH(d, GB(true), GB(false));
#undef GB
  return check_launch("gate_bwd2_kernel");
}

}  // namespace mmdit

using namespace mmdit;

extern "C" int mmdit_set_row_kernel_generation(int32_t generation) {
  MMDIT_REQUIRE(generation == 1 || generation == 2, MMDIT_ERR_ARG,
                "set_row_kernel_generation: 1 (first generation) or 2");
  g_row_generation = generation;
  return MMDIT_OK;
}
