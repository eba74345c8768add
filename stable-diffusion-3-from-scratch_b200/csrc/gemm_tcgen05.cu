// bf16 GEMM on tcgen05 tensor cores for sm_100a.
//
//   D[M,N] = epilogue(A[M,K] * B[N,K]^T), fp32 accumulation in TMEM.
//
// One persistent CTA per SM, 256 threads, warp-specialised:
//   warp 0      TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring)
//   warp 1      MMA issuer     (one lane issues tcgen05.mma, commits to mbarriers)
//   warp 2      TMEM allocator
//   warps 4..7  epilogue       (tcgen05.ld -> registers -> fused epilogue -> global)
// Pipelines: smem full/empty ring (TMA <-> MMA) and a 2-deep TMEM accumulator
// ring (MMA <-> epilogue) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Tile: 128 x block_n x 64 (block_n in {64,128,256}, chosen at launch).  Both
// operands may be K-major or MN-major (see mmdit_b200.h), which gives fprop,
// dgrad and wgrad from the same kernel without transposed copies.  Split-K
// (fp32 atomics) keeps all SMs busy on the small-output / long-reduction wgrads.
#include <stdlib.h>

#include "common.cuh"
#include "mmdit_b200.h"

namespace mmdit {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB
constexpr int MAX_STAGES = 8;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_THREADS_SWIGLU_BWD = 384;  // 8 epilogue warps: two per TMEM lane quarter, 32 columns of a chunk each
constexpr int SMEM_BUDGET = 227 * 1024;
constexpr int BAR_BYTES = 256;
constexpr int EPI_STAGE_BYTES = 4 * 32 * 64 * 4;  // 4 warps x (32 rows x 64 fp32)
constexpr int EPI_STAGE_BYTES_SWIGLU_BWD = 4 * 6 * 4096;  // 4 warps x (2 x (x1, x2) in; d1, d2 out) 32 x 64 bf16 tiles

struct GemmParams {
  CUtensorMap tmA, tmB;
  CUtensorMap tmD;  // bf16 output as {N, M}, box {64, 32}, 128B swizzle (TMA-store epilogues)
  CUtensorMap tmAux;  // EV_SWIGLU_TMA: the pre-activation [M, N] (D is then [M, N/2])
  int swiglu_half;    // N/2 in SwiGLU mode (B rows [0,N/2) = gate, [N/2,N) = up), else 0
  int M, N, K;
  int block_n, stages;
  int a_mn, b_mn;
  int tiles_m, tiles_n, split_k, kb_total, kb_per_split;
  // epilogue
  void* D;
  long long ldd;
  int d_fp32, accumulate, epi;
  const void* bias;
  int bias_fp32;
  const bf16* gate;
  long long rows_per_gate, ld_gate;
  const bf16* resid;
  long long ldr;
  bf16* aux;
  long long ld_aux;
  long long remap_rows, remap_batch_rows, remap_offset;
  long long slice_stride;  // split-K slices mode: split ks writes D + ks*slice_stride (no atomics)
  int debug;  // bit0: skip global stores, bit1: skip TMEM loads (perf experiments only)
  // EV_QKNORM_TMA (appended: the offsets of everything above stay what the other variants use)
  const float* qk_wq;      // [64] RMSNorm weight of q
  const float* qk_wk;      // [64] RMSNorm weight of k
  const float* rope_cos;   // [tokens, 32] or null (text stream)
  const float* rope_sin;
  int qk_d;                // model width: columns [0,d) = q, [d,2d) = k, [2d,3d) = v
  int qk_tokens;           // tokens per sample (row -> position for RoPE)
  float qk_eps;
  // EV_SWIGLU_BWD
  float* colsum_partial;   // [M/32, 2N] fp32 column sums of every 32-row strip of D, or null
};

__device__ __forceinline__ float bias_at(const GemmParams& p, int n) {
  if (!p.bias) return 0.f;
  return p.bias_fp32 ? reinterpret_cast<const float*>(p.bias)[n]
                     : __bfloat162float(reinterpret_cast<const bf16*>(p.bias)[n]);
}

// Fused epilogue for 8 consecutive columns [n, n+8) of output row m (physical row drow).
__device__ __forceinline__ void epilogue8(const GemmParams& p, long long m, long long drow, int n,
                                          float (&v)[8], int nvalid) {
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < nvalid) v[j] += bias_at(p, n + j);
  }
  const bool full = (nvalid == 8);
  if (p.aux) {
    bf16* ap = p.aux + drow * p.ld_aux + n;
    if (full && ((reinterpret_cast<uintptr_t>(ap) & 15) == 0)) {
      store8(ap, v);
    } else {
      for (int j = 0; j < nvalid; ++j) ap[j] = __float2bfloat16(v[j]);
    }
  }
  if (p.epi == MMDIT_EPI_GATE_RESID || p.epi == MMDIT_EPI_RESID) {
    float g[8], r[8];
    if (p.epi == MMDIT_EPI_GATE_RESID) {
      const bf16* gp = p.gate +
          static_cast<long long>(static_cast<unsigned>(m) / static_cast<unsigned>(p.rows_per_gate)) * p.ld_gate + n;
      if (full && ((reinterpret_cast<uintptr_t>(gp) & 15) == 0)) {
        load8(gp, g);
      } else {
        for (int j = 0; j < 8; ++j) g[j] = j < nvalid ? __bfloat162float(gp[j]) : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = 1.f;
    }
    const bf16* rp = p.resid + drow * p.ldr + n;
    if (full && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
      load8(rp, r);
    } else {
      for (int j = 0; j < 8; ++j) r[j] = j < nvalid ? __bfloat162float(rp[j]) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], g[j], r[j]);
  } else if (p.epi == MMDIT_EPI_SILU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = silu_f(v[j]);
  }
  if (!p.d_fp32) {
    bf16* dp = reinterpret_cast<bf16*>(p.D) + drow * p.ldd + n;
    if (full && ((reinterpret_cast<uintptr_t>(dp) & 15) == 0)) {
      store8(dp, v);
    } else {
      for (int j = 0; j < nvalid; ++j) dp[j] = __float2bfloat16(v[j]);
    }
  } else {
    float* dp = reinterpret_cast<float*>(p.D) + drow * p.ldd + n;
    if (p.accumulate) {
      if (p.split_k > 1) {
        for (int j = 0; j < nvalid; ++j) atomicAdd(dp + j, v[j]);
      } else if (full && ((reinterpret_cast<uintptr_t>(dp) & 15) == 0)) {
        float4 a = *reinterpret_cast<float4*>(dp), b = *reinterpret_cast<float4*>(dp + 4);
        a.x += v[0]; a.y += v[1]; a.z += v[2]; a.w += v[3];
        b.x += v[4]; b.y += v[5]; b.z += v[6]; b.w += v[7];
        *reinterpret_cast<float4*>(dp) = a;
        *reinterpret_cast<float4*>(dp + 4) = b;
      } else {
        for (int j = 0; j < nvalid; ++j) dp[j] += v[j];
      }
    } else if (full && ((reinterpret_cast<uintptr_t>(dp) & 15) == 0)) {
      *reinterpret_cast<float4*>(dp) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(dp + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      for (int j = 0; j < nvalid; ++j) dp[j] = v[j];
    }
  }
}

// Epilogue variants.  The fast ones are straight-line code for the hot GEMMs of the block
// (all run-time flags resolved at compile time, full 8-column vectors, 16-byte aligned rows);
// EV_GENERIC keeps every option behind run-time flags (tails, remap, SiLU, odd alignments).
enum { EV_GENERIC = 0, EV_BF16 = 1, EV_BF16_BIAS = 2, EV_GATE = 3, EV_F32 = 4, EV_F32_ATOMIC = 5,
       EV_BF16_TMA = 6, EV_SWIGLU_TMA = 7, EV_QKNORM_TMA = 8, EV_SWIGLU_BWD = 9 };
// EV_SWIGLU_BWD (CTA pairs only): the GEMM is the data gradient of xformers' w3 (MLP.py:19,32),
// acc = dY W3 = d(silu(x1) * x2) [M, N = hidden]; the epilogue applies the SwiGLU backward to it against
// the saved pre-activations aux = [x1 | x2] ([M, 2N]) and writes D = [d x1 | d x2] ([M, 2N]) -- the
// activation-gradient tensor is never written or re-read and the separate swiglu_bwd pass (10 B per
// hidden element at 0.7-0.8 of the HBM rate, 2.0 ms of the cfg2 step) disappears.  Per epilogue warp and
// 64-column chunk: the x1 / x2 tiles of its 32 rows arrive by TMA (issued one chunk ahead by the warp
// itself), each thread reads its row from the swizzled tiles, computes d1 / d2 from the bf16-rounded
// accumulator exactly as swiglu_bwd_kernel does, and the two output tiles leave by TMA stores; the
// column sums of the strip (w12 bias gradient) are read back from the output tiles, two columns per lane.
// Validated bit-identical to gemm + swiglu_bwd (tools/gemm_probe.py swiglu_bwd).  Eight epilogue warps (two
// per TMEM lane quarter, 32 columns of every chunk each, sharing the quarter's tiles) instead of four changed
// little (142 -> 134 us at the cfg2 image shape against 57 + 95 us for the two kernels): the fused kernel
// moves 430 MB (h12 in, dh12 out) with one 8 KB prefetch in flight per lane quarter, i.e. it is bound by
// bytes in flight (Little: 32 KB x 148 SMs / ~1.5 us = 3.2 TB/s, the measured rate), and a deeper x-tile
// ring does not fit next to a 4-stage operand ring.  Inside the cfg2 step it buys nothing (the separate
// swiglu_bwd overlaps the other stream's GEMMs), so the trainer keeps the two kernels unless
// MMDIT_FUSED_SWIGLU_BWD=1.  Two hazards found on hardware are written down where they bit: the x tiles
// must be double-buffered, and the output tiles need a bar.sync (not __syncwarp) between the st.shared
// and the TMA store.
// EV_QKNORM_TMA (experimental, MMDIT_FUSED_QKNORM=1; NOT yet validated on hardware): the packed
// q|k|v projection writes the raw projection (D, needed by the backward) and, for the q and k
// columns, the per-head RMSNorm * weight followed by the 2-D RoPE rotation (aux = [M, 2d]) --
// Attention.py:61-64,130-134,174-194 in the producer's epilogue.  A 64-column chunk is exactly one
// head and the epilogue thread owns the whole row of it, so the norm needs no shuffles; the
// arithmetic (bf16-rounded projection in, fp32 norm, bf16 rounding before the rotation, summation
// order of the squares) restates qknorm_rope_fwd_kernel so that both paths agree bit for bit.
// EV_SWIGLU_TMA (CTA pairs only): B = xformers' w12 ([gate rows; up rows], MLP.py:19).  The leader
// CTA stages 128 gate rows, its peer the matching 128 up rows, so every accumulator row holds
// gate[128] | up[128] of the same hidden columns: the epilogue writes the pre-activation (aux,
// bf16, needed by the backward) AND silu(gate) * up (D) -- the activation kernel and its re-read
// of the 8d-wide pre-activation disappear.  The activation is computed from the bf16-rounded
// pre-activation, i.e. bit-identical to mmdit_swiglu_fwd on the stored aux.
// EV_BF16_TMA (bf16 output, optional fp32 bias): each epilogue thread owns one accumulator row
// (its TMEM lane); a 64-column chunk is converted in registers, written as one 128-byte row of
// a 128B-swizzled [32 x 64] bf16 staging tile (conflict-free 16-byte stores), and the tile leaves
// through one TMA store.  No shared-memory read-back, no per-thread global stores, ragged M / N
// edges clipped by the TMA unit; two staging tiles per warp keep a store in flight.

// Operands of the gated-residual epilogue for one 64-column chunk: 8 passes x (gate, resid),
// loaded as a batch (and before the accumulator is needed) so their latency is paid once.
// 16-byte streaming load that does not allocate in L1 (which is almost entirely carved out as
// shared memory here, so allocating loads would serialise on a few KB of cache).
__device__ __forceinline__ uint4 ld_stream16(const void* ptr) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(ptr));
  return r;
}
struct GatePrefetch {
  uint4 g[8], r[8];
};
__device__ __forceinline__ void gate_prefetch(const GemmParams& p, int lane, long long m0, int n,
                                              GatePrefetch& pf) {
  const int sub_row = lane >> 3;
#pragma unroll
  for (int pass = 0; pass < 8; ++pass) {
    const long long m = m0 + pass * 4 + sub_row;
    if (m < p.M && n < p.N && !(p.debug & 8)) {
      const unsigned gb = static_cast<unsigned>(m) / static_cast<unsigned>(p.rows_per_gate);
      pf.g[pass] = __ldg(reinterpret_cast<const uint4*>(p.gate + static_cast<long long>(gb) * p.ld_gate + n));
      pf.r[pass] = ld_stream16(p.resid + m * p.ldr + n);
    }
  }
}

template <int EV>
__device__ __forceinline__ void epilogue_fast(const GemmParams& p, const float* wbuf, int lane,
                                              long long m0, int n, const GatePrefetch* pf,
                                              long long slice_off) {
  const int sub_row = lane >> 3, seg = lane & 7;
  float bias[8];
  if constexpr (EV == EV_BF16_BIAS || EV == EV_GATE) {
    if (p.bias) {
      const float4 b0 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.bias) + n);
      const float4 b1 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.bias) + n + 4);
      bias[0] = b0.x; bias[1] = b0.y; bias[2] = b0.z; bias[3] = b0.w;
      bias[4] = b1.x; bias[5] = b1.y; bias[6] = b1.z; bias[7] = b1.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) bias[j] = 0.f;
    }
  }
#pragma unroll
  for (int pass = 0; pass < 8; ++pass) {
    const int row = pass * 4 + sub_row;
    const long long m = m0 + row;
    const float4* rrow = reinterpret_cast<const float4*>(wbuf + row * 64);
    const float4 a = rrow[(2 * seg) ^ (row & 7)];
    const float4 b = rrow[(2 * seg + 1) ^ (row & 7)];
    if (m >= p.M) continue;
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if constexpr (EV == EV_BF16_BIAS || EV == EV_GATE) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += bias[j];
    }
    if constexpr (EV == EV_GATE) {
      if (!(p.debug & 4)) store8(p.aux + m * p.ld_aux + n, v);  // pre-gate value, needed for dgate
      const uint4 gu = pf->g[pass], ru = pf->r[pass];
      const float2 g0 = unpack_bf16x2(gu.x), g1 = unpack_bf16x2(gu.y), g2 = unpack_bf16x2(gu.z),
                   g3 = unpack_bf16x2(gu.w);
      const float2 r0 = unpack_bf16x2(ru.x), r1 = unpack_bf16x2(ru.y), r2 = unpack_bf16x2(ru.z),
                   r3 = unpack_bf16x2(ru.w);
      v[0] = fmaf(v[0], g0.x, r0.x); v[1] = fmaf(v[1], g0.y, r0.y);
      v[2] = fmaf(v[2], g1.x, r1.x); v[3] = fmaf(v[3], g1.y, r1.y);
      v[4] = fmaf(v[4], g2.x, r2.x); v[5] = fmaf(v[5], g2.y, r2.y);
      v[6] = fmaf(v[6], g3.x, r3.x); v[7] = fmaf(v[7], g3.y, r3.y);
    }
    if constexpr (EV == EV_BF16 || EV == EV_BF16_BIAS || EV == EV_GATE) {
      store8(reinterpret_cast<bf16*>(p.D) + m * p.ldd + n, v);
    } else if constexpr (EV == EV_F32) {
      float* dp = reinterpret_cast<float*>(p.D) + slice_off + m * p.ldd + n;
      *reinterpret_cast<float4*>(dp) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(dp + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else if constexpr (EV == EV_F32_ATOMIC) {
      float* dp = reinterpret_cast<float*>(p.D) + m * p.ldd + n;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dp), "f"(v[0]), "f"(v[1]),
                   "f"(v[2]), "f"(v[3]) : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dp + 4), "f"(v[4]), "f"(v[5]),
                   "f"(v[6]), "f"(v[7]) : "memory");
    }
  }
}

// PAIR = true: launched as clusters of two CTAs (one TPC).  The pair owns a 256 x block_n output
// tile: CTA `rank` holds A rows [rank*128, +128), the B rows [rank*block_n/2, +block_n/2) and the
// accumulator rows [rank*128, +128) in its own TMEM; the leader issues tcgen05.mma.cta_group::2,
// which reads both halves of B from both SMs.  Per SM and k-block that is 32 KB of operand
// traffic (L2 -> smem) instead of 48 KB for the same FLOPs -- the single-CTA 128x256 tile is
// L2-bandwidth bound at ~2/3 of the MMA rate on B200 (profiles/r01_gemm_shapes_*.log).
template <int EV, bool PAIR>
__global__ void __launch_bounds__(EV == EV_SWIGLU_BWD ? GEMM_THREADS_SWIGLU_BWD : GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024 B alignment in the shared window.
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int block_n = p.block_n;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;         // 0 = leader
  const int unit0 = PAIR ? (blockIdx.x >> 1) : blockIdx.x;     // persistent work-unit stride
  const int nunits = PAIR ? (gridDim.x >> 1) : gridDim.x;
  const int b_rows = PAIR ? block_n / 2 : block_n;             // B rows staged by this CTA
  const int b_stage_bytes = b_rows * BLOCK_K * 2;
  const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;
  const int stages = p.stages;

  // epilogue staging: per epilogue warp 32 rows x 64 fp32 (16-byte chunks XOR-swizzled by row)
  uint8_t* stage_buf = smem + stages * stage_bytes;
  constexpr int kEpiBytes = EV == EV_SWIGLU_BWD ? EPI_STAGE_BYTES_SWIGLU_BWD : EPI_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_buf + kEpiBytes);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full = empty_bar + MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  [[maybe_unused]] uint64_t* epi_in = tmem_empty + 3;   // [4 warps][2 buffers] EV_SWIGLU_BWD: x1 / x2 tiles landed

  const uint32_t tmem_cols = 2u * block_n;  // 128 / 256 / 512: powers of two

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], (PAIR ? 8 : 4) * (EV == EV_SWIGLU_BWD ? 2 : 1));  // one arrive per epilogue warp (of both CTAs)
    }
    if constexpr (EV == EV_SWIGLU_BWD)
      for (int w = 0; w < 8; ++w) mbar_init(&epi_in[w], 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) {
      tmem_alloc_pair(tmem_slot, tmem_cols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, tmem_cols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();  // peer barriers initialised before any remote arrive / TMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // everything above overlapped the previous kernel's tail; its outputs are visible from here

  const int total_work = p.tiles_m * p.tiles_n * p.split_k;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // The whole warp walks the loop and waits on the barriers; ONE elected lane issues (guarding with
    // `lane == 0` instead makes ptxas wrap every TMA / MMA instruction in an ELECT loop, see elect_one()).
    {
      int stage = 0;
      uint32_t phase = 0;
      // In a pair both CTAs load their halves and credit the LEADER's full barrier, which
      // expects the bytes of both (only the leader's MMA thread waits on it).
      auto load = [&](void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
        if constexpr (PAIR) tma_load_2d_pair(dst, tm, mapa_u32(smem_u32(bar), 0), c0, c1);
        else tma_load_2d(dst, tm, bar, c0, c1);
      };
      for (int work = unit0; work < total_work; work += nunits) {
        const int tile = work / p.split_k, ks = work % p.split_k;
        const int m_blk = PAIR ? (tile % p.tiles_m) * 2 + (int)rank : tile % p.tiles_m;  // 128-row units
        const int n_blk = tile / p.tiles_m;
        const int n0 = p.swiglu_half ? n_blk * b_rows + (int)rank * p.swiglu_half
                                     : n_blk * block_n + (int)rank * (PAIR ? b_rows : 0);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * stage_bytes;
          uint8_t* sB = sA + A_STAGE_BYTES;
          if (elect_one()) {
            if (!PAIR || rank == 0) mbar_expect_tx(&full_bar[stage], PAIR ? 2 * stage_bytes : stage_bytes);
            if (!p.a_mn) {
              load(sA, &p.tmA, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
            } else {
              for (int i = 0; i < BLOCK_M / 64; ++i)
                load(sA + i * (BLOCK_K * 128), &p.tmA, &full_bar[stage], m_blk * BLOCK_M + i * 64, kb * BLOCK_K);
            }
            if (!p.b_mn) {
              load(sB, &p.tmB, &full_bar[stage], kb * BLOCK_K, n0);
            } else {
              for (int i = 0; i < b_rows / 64; ++i)
                load(sB + i * (BLOCK_K * 128), &p.tmB, &full_bar[stage], n0 + i * 64, kb * BLOCK_K);
            }
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(PAIR ? 2 * BLOCK_M : BLOCK_M, block_n, p.a_mn, p.b_mn);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int work = unit0; work < total_work; work += nunits) {
        const int ks = work % p.split_k;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * block_n;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
          const uint32_t b_addr = a_addr + A_STAGE_BYTES;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {
              const uint64_t adesc =
                  p.a_mn ? desc_mnmajor(a_addr, k, BLOCK_K * 128) : desc_kmajor(a_addr, k);
              const uint64_t bdesc =
                  p.b_mn ? desc_mnmajor(b_addr, k, BLOCK_K * 128) : desc_kmajor(b_addr, k);
              if constexpr (PAIR) umma_bf16_pair(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              else umma_bf16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            // frees the smem slot (in both CTAs of a pair) when these MMAs retire
            if constexpr (PAIR) umma_commit_pair(&empty_bar[stage]);
            else umma_commit(&empty_bar[stage]);
            // accumulator ready for the epilogue warps (of both CTAs)
            if (kb == kb1 - 1) {
              if constexpr (PAIR) umma_commit_pair(&tmem_full[as]);
              else umma_commit(&tmem_full[as]);
            }
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue
    const int ew = (warp - 4) & 3;  // == warp % 4 -> TMEM lane quarter
    [[maybe_unused]] const int hw = (warp - 4) >> 2;  // EV_SWIGLU_BWD: which 32-column half of every chunk (8 epilogue warps)
    int as = 0;
    uint32_t aphase = 0;
    const uint32_t leader_tmem_empty = PAIR ? mapa_u32(smem_u32(&tmem_empty[0]), 0) : 0u;
    [[maybe_unused]] int tma_buf = 0;
    [[maybe_unused]] uint32_t epi_q = 0;   // EV_SWIGLU_BWD: chunks processed so far (x-tile buffer = q & 1)
    if constexpr (EV == EV_SWIGLU_BWD) {
      if (hw == 0 && unit0 < total_work && elect_one()) {   // x1 / x2 tiles of this lane quarter's first chunk
        const int tile0 = unit0 / p.split_k;
        const long long mf = static_cast<long long>(PAIR ? (tile0 % p.tiles_m) * 2 + (int)rank : tile0 % p.tiles_m) * BLOCK_M + ew * 32;
        const int nf = (tile0 / p.tiles_m) * 256;
        uint8_t* wb = stage_buf + ew * 24576;
        mbar_expect_tx(&epi_in[2 * ew], 2 * 4096);
        tma_load_2d(wb, &p.tmAux, &epi_in[2 * ew], nf, static_cast<int>(mf));
        tma_load_2d(wb + 4096, &p.tmAux, &epi_in[2 * ew], p.N + nf, static_cast<int>(mf));
      }
      __syncwarp();
    }
    for (int work = unit0; work < total_work; work += nunits) {
      const int tile = work / p.split_k;
      const int m_blk = PAIR ? (tile % p.tiles_m) * 2 + (int)rank : tile % p.tiles_m;
      const int n_blk = tile / p.tiles_m;
      const long long m0 = static_cast<long long>(m_blk) * BLOCK_M + ew * 32;
      [[maybe_unused]] GatePrefetch pf;
      if constexpr (EV == EV_GATE) gate_prefetch(p, lane, m0, n_blk * block_n + (lane & 7) * 8, pf);
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      // Accumulator rows live one per thread (TMEM lane).  Each 64-column chunk is transposed
      // through shared memory so that 8 lanes cover one row's 64 columns: every global load /
      // store of the fused epilogue is then a full, coalesced 128-byte (bf16) line per row.
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * block_n;
      float* wbuf = reinterpret_cast<float*>(stage_buf + ew * (32 * 64 * 4));
      const int nchunks = block_n / 64;
      const int sub_row = lane >> 3, seg = lane & 7;
      if constexpr (EV == EV_QKNORM_TMA) {
        uint8_t* sbase = stage_buf + ew * (32 * 64 * 4);
        const float* bp = reinterpret_cast<const float*>(p.bias);
        auto stage_tile = [&](const uint32_t (&packed)[32], const CUtensorMap* tm, int col) {
          uint8_t* tile = sbase + (tma_buf & 1) * 4096;
          if (elect_one()) tma_wait_group_read1();
          __syncwarp();
          uint8_t* prow = tile + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(prow + ((j ^ (lane & 7)) << 4)) =
                make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (!(p.debug & 1) && elect_one()) {
            tma_store_2d(tm, tile, col, static_cast<int>(m0));
            tma_commit_group();
          }
          ++tma_buf;
        };
        const int nchunks_q = block_n / 64;
        for (int c = 0; c < nchunks_q; ++c) {
          const int n = n_blk * block_n + c * 64;
          uint32_t pk[32];
          {
            uint32_t r0[32], r1[32];
            tmem_ld32(taddr + c * 64, r0);
            tmem_ld32(taddr + c * 64 + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const uint32_t* src = j < 8 ? &r0[4 * j] : &r1[4 * (j - 8)];
              float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
              if (bp && n + 4 * j < p.N) b = __ldg(reinterpret_cast<const float4*>(bp + n + 4 * j));
              pk[2 * j] = pack_bf16x2(__uint_as_float(src[0]) + b.x, __uint_as_float(src[1]) + b.y);
              pk[2 * j + 1] = pack_bf16x2(__uint_as_float(src[2]) + b.z, __uint_as_float(src[3]) + b.w);
            }
          }
          if (c == nchunks_q - 1) {  // accumulator drained: hand the TMEM stage back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (PAIR) mbar_arrive_cluster(leader_tmem_empty + as * 8);
              else mbar_arrive(&tmem_empty[as]);
            }
          }
          if (n >= p.N || m0 >= p.M) continue;
          stage_tile(pk, &p.tmD, n);                 // raw projection (bf16), all of q | k | v
          const int part = n / p.qk_d;               // 0: q, 1: k, 2: v
          if (part >= 2) continue;
          const float* w = part == 0 ? p.qk_wq : p.qk_wk;
          // sum of squares: 8 sequential partial sums of 8, combined as the 8-lane xor tree of
          // qknorm_rope_fwd_kernel does ((s0+s1)+(s2+s3)) + ((s4+s5)+(s6+s7))
          float sg[8];
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 v = unpack_bf16x2(pk[4 * g + j]);
              a += v.x * v.x;
              a += v.y * v.y;
            }
            sg[g] = a;
          }
          const float ss = ((sg[0] + sg[1]) + (sg[2] + sg[3])) + ((sg[4] + sg[5]) + (sg[6] + sg[7]));
          const float rn = rsqrtf(ss * (1.f / 64.f) + p.qk_eps);
          const long long mrow = m0 + lane;
          const int tok = static_cast<int>(static_cast<unsigned long long>(mrow) %
                                           static_cast<unsigned>(p.qk_tokens > 0 ? p.qk_tokens : 1));
          const float* cs = p.rope_cos ? p.rope_cos + static_cast<long long>(tok) * 32 : nullptr;
          const float* sn = p.rope_sin ? p.rope_sin + static_cast<long long>(tok) * 32 : nullptr;
#pragma unroll
          for (int j = 0; j < 16; ++j) {   // 4 columns = 2 rotation pairs per trip
            const float4 wv = __ldg(reinterpret_cast<const float4*>(w + 4 * j));
            const float2 a = unpack_bf16x2(pk[2 * j]), b = unpack_bf16x2(pk[2 * j + 1]);
            // the reference rounds the RMSNorm output to bf16 before the fp32 rotation
            float y0 = __bfloat162float(__float2bfloat16(a.x * rn * wv.x));
            float y1 = __bfloat162float(__float2bfloat16(a.y * rn * wv.y));
            float y2 = __bfloat162float(__float2bfloat16(b.x * rn * wv.z));
            float y3 = __bfloat162float(__float2bfloat16(b.y * rn * wv.w));
            if (cs) {
              const float2 cc = *reinterpret_cast<const float2*>(cs + 2 * j);
              const float2 sv = *reinterpret_cast<const float2*>(sn + 2 * j);
              const float t0 = y0 * cc.x - y1 * sv.x, t1 = y1 * cc.x + y0 * sv.x;
              const float t2 = y2 * cc.y - y3 * sv.y, t3 = y3 * cc.y + y2 * sv.y;
              y0 = t0; y1 = t1; y2 = t2; y3 = t3;
            }
            pk[2 * j] = pack_bf16x2(y0, y1);
            pk[2 * j + 1] = pack_bf16x2(y2, y3);
          }
          stage_tile(pk, &p.tmAux, n);               // aux is [M, 2d]: q at column 0, k at column d
        }
      } else if constexpr (EV == EV_SWIGLU_TMA) {
        uint8_t* sbase = stage_buf + ew * (32 * 64 * 4);
        const float* bp = reinterpret_cast<const float*>(p.bias);
        auto stage_tile = [&](const uint32_t (&packed)[32], const CUtensorMap* tm, int col) {
          uint8_t* tile = sbase + (tma_buf & 1) * 4096;
          if (elect_one()) tma_wait_group_read1();
          __syncwarp();
          uint8_t* prow = tile + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(prow + ((j ^ (lane & 7)) << 4)) =
                make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (!(p.debug & 1) && elect_one()) {
            tma_store_2d(tm, tile, col, static_cast<int>(m0));
            tma_commit_group();
          }
          ++tma_buf;
        };
        // bf16-round 64 accumulator columns (+bias) into 32 packed words
        auto round64 = [&](const uint32_t (&lo)[32], const uint32_t (&hi)[32], int bias_col, uint32_t (&out)[32]) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint32_t* src = j < 8 ? &lo[4 * j] : &hi[4 * (j - 8)];
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bp) b = __ldg(reinterpret_cast<const float4*>(bp + bias_col + 4 * j));
            out[2 * j] = pack_bf16x2(__uint_as_float(src[0]) + b.x, __uint_as_float(src[1]) + b.y);
            out[2 * j + 1] = pack_bf16x2(__uint_as_float(src[2]) + b.z, __uint_as_float(src[3]) + b.w);
          }
        };
        for (int c = 0; c < 2; ++c) {
          const int n = n_blk * 128 + c * 64;   // column inside the gate (and the up) half
          uint32_t gq[32], uq[32];
          {
            uint32_t r0[32], r1[32];
            tmem_ld32(taddr + c * 64, r0);
            tmem_ld32(taddr + c * 64 + 32, r1);
            tmem_ld_wait();
            round64(r0, r1, n, gq);
            tmem_ld32(taddr + 128 + c * 64, r0);
            tmem_ld32(taddr + 128 + c * 64 + 32, r1);
            tmem_ld_wait();
            round64(r0, r1, p.swiglu_half + n, uq);
          }
          if (c == 1) {  // accumulator drained: hand the TMEM stage back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (PAIR) mbar_arrive_cluster(leader_tmem_empty + as * 8);
              else mbar_arrive(&tmem_empty[as]);
            }
          }
          if (m0 >= p.M) continue;
          if (p.aux) {   // the pre-activation is only kept for the backward (inference passes aux = NULL)
            stage_tile(gq, &p.tmAux, n);
            stage_tile(uq, &p.tmAux, p.swiglu_half + n);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float2 g = unpack_bf16x2(gq[j]), u = unpack_bf16x2(uq[j]);
            gq[j] = pack_bf16x2(silu_f(g.x) * u.x, silu_f(g.y) * u.y);
          }
          stage_tile(gq, &p.tmD, n);
        }
      } else if constexpr (EV == EV_SWIGLU_BWD) {
        uint8_t* wbase = stage_buf + ew * 24576;   // [buf 0: x1, x2][buf 1: x1, x2][d1][d2], 4 KiB each
        uint8_t* tD1 = wbase + 16384, *tD2 = wbase + 20480;
        const int hid = p.N;
        for (int c = 0; c < 4; ++c, ++epi_q) {
          const int n = n_blk * 256 + c * 64;
          const int buf = epi_q & 1;
          uint8_t* tX1 = wbase + buf * 8192, *tX2 = tX1 + 4096;
          {
            // x1 / x2 tiles of the NEXT chunk (of this tile or of the next work unit) into the other buffer:
            // its previous contents were consumed by the arithmetic of the chunk before this one.  (A
            // single-buffered refill issued right after this chunk's ld.shared corrupted ~1e-6 of the
            // elements even with the loads consumed first; issued after the chunk's stores it was clean.)
            int n_next = n + 64;
            long long m_next = m0;
            bool has_next = true;
            if (c == 3) {
              const int work2 = work + nunits;
              has_next = work2 < total_work;
              const int tile2 = work2 / p.split_k;
              m_next = static_cast<long long>(PAIR ? (tile2 % p.tiles_m) * 2 + (int)rank : tile2 % p.tiles_m) * BLOCK_M + ew * 32;
              n_next = (tile2 / p.tiles_m) * 256;
            }
            if (hw == 0 && has_next && elect_one()) {
              uint64_t* nb = &epi_in[2 * ew + (buf ^ 1)];
              uint8_t* nx = wbase + (buf ^ 1) * 8192;
              mbar_expect_tx(nb, 2 * 4096);
              tma_load_2d(nx, &p.tmAux, nb, n_next, static_cast<int>(m_next));
              tma_load_2d(nx + 4096, &p.tmAux, nb, hid + n_next, static_cast<int>(m_next));
            }
            __syncwarp();
          }
          uint32_t r0[32];
          tmem_ld32(taddr + c * 64 + hw * 32, r0);
          tmem_ld_wait();
          if (c == 3) {  // accumulator drained: hand the TMEM stage back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (PAIR) mbar_arrive_cluster(leader_tmem_empty + as * 8);
              else mbar_arrive(&tmem_empty[as]);
            }
          }
          // this thread's half row of the x1 / x2 tiles (128B-swizzled, as TMA wrote them)
          mbar_wait(&epi_in[2 * ew + buf], (epi_q >> 1) & 1u);
          uint32_t x1p[16], x2p[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int jj = hw * 4 + j;
            const uint4 u = *reinterpret_cast<const uint4*>(tX1 + lane * 128 + ((jj ^ (lane & 7)) << 4));
            const uint4 v = *reinterpret_cast<const uint4*>(tX2 + lane * 128 + ((jj ^ (lane & 7)) << 4));
            x1p[4 * j] = u.x; x1p[4 * j + 1] = u.y; x1p[4 * j + 2] = u.z; x1p[4 * j + 3] = u.w;
            x2p[4 * j] = v.x; x2p[4 * j + 1] = v.y; x2p[4 * j + 2] = v.z; x2p[4 * j + 3] = v.w;
          }
          // d1 = g x2 sg (1 + x1 (1 - sg)),  d2 = g x1 sg   with g rounded to bf16 (as the unfused path stores it)
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float g0 = __bfloat162float(__float2bfloat16(__uint_as_float(r0[2 * j])));
            const float g1 = __bfloat162float(__float2bfloat16(__uint_as_float(r0[2 * j + 1])));
            const float2 a = unpack_bf16x2(x1p[j]), b = unpack_bf16x2(x2p[j]);
            const float s0 = __fdividef(1.f, 1.f + __expf(-a.x)), s1 = __fdividef(1.f, 1.f + __expf(-a.y));
            x1p[j] = pack_bf16x2(g0 * b.x * s0 * (1.f + a.x * (1.f - s0)), g1 * b.y * s1 * (1.f + a.y * (1.f - s1)));
            x2p[j] = pack_bf16x2(g0 * (a.x * s0), g1 * (a.y * s1));
          }
          // the previous chunk's stores (issued by the hw == 0 warp) have read the output tiles
          if (hw == 0 && elect_one()) tma_wait_group_read0();
          named_bar_sync(8 + ew, 64);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int jj = hw * 4 + j;
            *reinterpret_cast<uint4*>(tD1 + lane * 128 + ((jj ^ (lane & 7)) << 4)) =
                make_uint4(x1p[4 * j], x1p[4 * j + 1], x1p[4 * j + 2], x1p[4 * j + 3]);
            *reinterpret_cast<uint4*>(tD2 + lane * 128 + ((jj ^ (lane & 7)) << 4)) =
                make_uint4(x2p[4 * j], x2p[4 * j + 1], x2p[4 * j + 2], x2p[4 * j + 3]);
          }
          fence_proxy_async_smem();
          // bar.sync over the two warps of the lane quarter (it also drains their pending st.shared: with a
          // plain __syncwarp the store below occasionally read a tile whose last 16-byte chunk of a few rows
          // was still the previous chunk's -- single-buffered output tiles leave no slack)
          named_bar_sync(8 + ew, 64);
          if (hw == 0 && !(p.debug & 1) && elect_one()) {
            tma_store_2d(&p.tmD, tD1, n, static_cast<int>(m0));
            tma_store_2d(&p.tmD, tD2, hid + n, static_cast<int>(m0));
            tma_commit_group();
          }
          __syncwarp();   // reconverge before the next chunk's warp-collective tcgen05.ld
          if (p.colsum_partial) {   // column sums of this 32-row strip: lane l owns column 32 hw + l of the chunk
            float a0 = 0.f, b0 = 0.f;
            const int jch = hw * 4 + (lane >> 3), bofs = (lane & 7) * 2;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
              const int off = r * 128 + ((jch ^ (r & 7)) << 4) + bofs;
              a0 += __bfloat162float(*reinterpret_cast<const bf16*>(tD1 + off));
              b0 += __bfloat162float(*reinterpret_cast<const bf16*>(tD2 + off));
            }
            float* dst = p.colsum_partial + (m0 >> 5) * (2LL * hid) + n + hw * 32 + lane;
            dst[0] = a0;
            dst[hid] = b0;
          }
        }
      } else if constexpr (EV == EV_BF16_TMA) {
        uint8_t* sbase = stage_buf + ew * (32 * 64 * 4);   // two 4 KiB bf16 tiles (1024 B aligned)
        for (int c = 0; c < nchunks; ++c) {
          const int n = n_blk * block_n + c * 64;
          uint32_t r0[32], r1[32];
          tmem_ld32(taddr + c * 64, r0);
          tmem_ld32(taddr + c * 64 + 32, r1);
          tmem_ld_wait();
          if (c == nchunks - 1) {  // accumulator drained: hand the TMEM stage back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (PAIR) mbar_arrive_cluster(leader_tmem_empty + as * 8);
              else mbar_arrive(&tmem_empty[as]);
            }
          }
          if (n >= p.N || m0 >= p.M) continue;  // whole chunk beyond a ragged edge
          uint8_t* tile = sbase + (tma_buf & 1) * 4096;
          // the store that last used this tile (two chunks ago) must have finished reading it
          if (elect_one()) tma_wait_group_read1();
          __syncwarp();
          uint8_t* prow = tile + lane * 128;
          const float* bp = reinterpret_cast<const float*>(p.bias);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t* src = j < 4 ? &r0[8 * j] : &r1[8 * (j - 4)];
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(src[i]);
            if (bp) {   // N % 8 == 0 and n + 8j < N is checked per 8-column group
              if (n + 8 * j < p.N) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(bp + n + 8 * j));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(bp + n + 8 * j + 4));
                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
              }
            }
            uint4 u;
            u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
            u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
            *reinterpret_cast<uint4*>(prow + ((j ^ (lane & 7)) << 4)) = u;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (!(p.debug & 1) && elect_one()) {
            tma_store_2d(&p.tmD, tile, n, static_cast<int>(m0));
            tma_commit_group();
          }
          ++tma_buf;
        }
      } else
      for (int c = 0; c < nchunks; ++c) {
        if constexpr (EV == EV_GATE) {
          if (c > 0) gate_prefetch(p, lane, m0, n_blk * block_n + c * 64 + seg * 8, pf);
        }
        uint32_t r0[32], r1[32];
        if (!(p.debug & 2)) {
          tmem_ld32(taddr + c * 64, r0);
          tmem_ld32(taddr + c * 64 + 32, r1);
          tmem_ld_wait();
        }
        if (c == nchunks - 1) {  // accumulator fully drained: hand the TMEM stage back early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (PAIR) mbar_arrive_cluster(leader_tmem_empty + as * 8);
            else mbar_arrive(&tmem_empty[as]);
          }
        }
        float4* wrow = reinterpret_cast<float4*>(wbuf + lane * 64);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          wrow[j ^ (lane & 7)] = make_float4(__uint_as_float(r0[4 * j]), __uint_as_float(r0[4 * j + 1]),
                                             __uint_as_float(r0[4 * j + 2]), __uint_as_float(r0[4 * j + 3]));
          wrow[(8 + j) ^ (lane & 7)] =
              make_float4(__uint_as_float(r1[4 * j]), __uint_as_float(r1[4 * j + 1]),
                          __uint_as_float(r1[4 * j + 2]), __uint_as_float(r1[4 * j + 3]));
        }
        __syncwarp();
        const int n = n_blk * block_n + c * 64 + seg * 8;
        if constexpr (EV != EV_GENERIC) {
          if (n < p.N && !(p.debug & 1))
            epilogue_fast<EV>(p, wbuf, lane, m0, n, &pf, (work % p.split_k) * p.slice_stride);
        } else {
          const int nvalid = min(8, p.N - n);
#pragma unroll 1
          for (int pass = 0; pass < 8; ++pass) {
            const int row = pass * 4 + sub_row;
            const long long m = static_cast<long long>(m_blk) * BLOCK_M + ew * 32 + row;
            const float4* rrow = reinterpret_cast<const float4*>(wbuf + row * 64);
            const float4 a = rrow[(2 * seg) ^ (row & 7)];
            const float4 b = rrow[(2 * seg + 1) ^ (row & 7)];
            if (m < p.M && nvalid > 0 && !(p.debug & 1)) {
              long long drow = m;
              if (p.remap_rows > 0) {
                const unsigned mm = static_cast<unsigned>(m), rr = static_cast<unsigned>(p.remap_rows);
                drow = static_cast<long long>(mm / rr) * p.remap_batch_rows + (mm % rr) + p.remap_offset;
              }
              float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
              epilogue8(p, m, drow, n, v, nvalid);
            }
          }
        }
        __syncwarp();  // staging buffer is reused by the next chunk
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    if constexpr (EV == EV_BF16_TMA || EV == EV_SWIGLU_TMA || EV == EV_QKNORM_TMA || EV == EV_SWIGLU_BWD) {
      if (elect_one()) tma_wait_group0();  // staging tiles must outlive their stores
    }
  }

  // This CTA's last tile is stored: the next kernel of the stream may be scheduled on the SMs that free
  // up (it runs its prologue and blocks in its own pdl_wait() until this grid has completed).  Triggering
  // at kernel entry instead parks the dependents' CTAs on the SMs for the whole kernel, which takes the
  // slots the other stream's kernels use to fill this kernel's tail (measured: step 29.6 -> 30.4 ms).
  if (warp >= 4) pdl_trigger();
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) {
    cluster_sync_all();  // no CTA exits (or frees TMEM) while its peer can still signal it
    if (warp == 2) tmem_dealloc_pair(tmem_base, tmem_cols);
  } else {
    if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
  }
}

static int pick_block_n(long long M, long long N, int sms) {
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  // prefer 256 unless it leaves the machine visibly under-filled
  const long long tm = (M + BLOCK_M - 1) / BLOCK_M;
  const long long t256 = tm * ((N + 255) / 256);
  const long long t128 = tm * ((N + 127) / 128);
  auto eff = [&](long long t, double cost) {
    long long waves = (t + sms - 1) / sms;
    return (double)t / (double)(waves * sms) / cost;
  };
  // 128-wide tiles run as single CTAs (no cta_group::2) with twice the operand traffic per FLOP,
  // which is L2-bandwidth bound on B200: they only pay off when 256-wide tiles leave most SMs idle
  return eff(t256, 1.0) >= eff(t128, 1.35) ? 256 : 128;
}

template <int EV, bool PAIR>
static int launch_variant(int grid, int smem_bytes, cudaStream_t stream, const GemmParams& p) {
  static const cudaError_t attr_rc = cudaFuncSetAttribute(   // thread-safe one-time initialisation
      gemm_tcgen05_kernel<EV, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET);
  if (attr_rc != cudaSuccess) {
    set_last_error("gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(attr_rc));
    return (int)attr_rc;
  }
  if constexpr (PAIR) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(EV == EV_SWIGLU_BWD ? GEMM_THREADS_SWIGLU_BWD : GEMM_THREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<EV, true>, p);
    if (e != cudaSuccess) {
      set_last_error("gemm: cudaLaunchKernelEx (CTA pair): %s", cudaGetErrorString(e));
      return (int)e;
    }
  } else {
    launch_k(gemm_tcgen05_kernel<EV, false>, dim3(grid), dim3(GEMM_THREADS), smem_bytes, stream, p);
  }
  return 0;
}

}  // namespace mmdit

using namespace mmdit;

extern "C" int mmdit_gemm_bf16(const mmdit_gemm_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MMDIT_REQUIRE(a && a->A && a->B && a->D, MMDIT_ERR_ARG, "gemm: null pointer argument");
  MMDIT_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, MMDIT_ERR_ARG, "gemm: bad shape %lld %lld %lld",
                (long long)a->M, (long long)a->N, (long long)a->K);
  MMDIT_REQUIRE(a->M < (1ll << 31) && a->N < (1ll << 31) && a->K < (1ll << 31), MMDIT_ERR_ARG,
                "gemm: dimension exceeds int32");
  MMDIT_REQUIRE(a->lda % 8 == 0 && a->ldb % 8 == 0, MMDIT_ERR_ALIGN,
                "gemm: lda/ldb must be multiples of 8 elements (16 B) for TMA, got %lld %lld",
                (long long)a->lda, (long long)a->ldb);
  MMDIT_REQUIRE(!(a->accumulate && !a->d_fp32), MMDIT_ERR_ARG, "gemm: accumulate needs fp32 D");
  MMDIT_REQUIRE(a->epilogue == MMDIT_EPI_NONE || a->epilogue == MMDIT_EPI_GATE_RESID ||
                    a->epilogue == MMDIT_EPI_SILU || a->epilogue == MMDIT_EPI_RESID ||
                    a->epilogue == MMDIT_EPI_SWIGLU || a->epilogue == MMDIT_EPI_QKNORM ||
                    a->epilogue == MMDIT_EPI_SWIGLU_BWD,
                MMDIT_ERR_UNSUPPORTED, "gemm: epilogue %d not supported", a->epilogue);
  if (a->epilogue == MMDIT_EPI_GATE_RESID)
    MMDIT_REQUIRE(a->gate && a->rows_per_gate > 0 && a->resid, MMDIT_ERR_ARG,
                  "gemm: gate/resid epilogue needs gate, rows_per_gate, resid");
  if (a->epilogue == MMDIT_EPI_RESID)
    MMDIT_REQUIRE(a->resid != nullptr, MMDIT_ERR_ARG, "gemm: resid epilogue needs resid");

  const bool swiglu = a->epilogue == MMDIT_EPI_SWIGLU;
  if (swiglu) {
    auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    MMDIT_REQUIRE(!a->d_fp32 && !a->accumulate && a->split_k <= 1 && a->remap_rows == 0 &&
                      a->b_major == 0 && a->N % 256 == 0 && a->M > BLOCK_M && al(a->D) && al(a->aux) &&
                      a->ldd % 8 == 0 && a->ld_aux % 8 == 0 && (!a->bias || (a->bias_fp32 && al(a->bias))),
                  MMDIT_ERR_UNSUPPORTED,
                  "gemm: the SwiGLU epilogue needs bf16 D [M,N/2] (and aux [M,N] unless NULL = inference), "
                  "K-major B, N %% 256 == 0, M > 128, 16-byte aligned rows and an fp32 bias");
  }
  const bool swiglu_bwd = a->epilogue == MMDIT_EPI_SWIGLU_BWD;
  if (swiglu_bwd) {
    auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    MMDIT_REQUIRE(a->aux && !a->d_fp32 && !a->accumulate && a->split_k <= 1 && a->remap_rows == 0 && !a->bias &&
                      a->N % 256 == 0 && a->M % BLOCK_M == 0 && a->M > BLOCK_M && al(a->D) && al(a->aux) &&
                      a->ldd % 8 == 0 && a->ld_aux % 8 == 0 && (!a->colsum_partial || al(a->colsum_partial)),
                  MMDIT_ERR_UNSUPPORTED,
                  "gemm: the SwiGLU-backward epilogue needs aux [M,2N], bf16 D [M,2N], no bias, N %% 256 == 0, "
                  "M %% 128 == 0, M > 128 and 16-byte aligned rows");
  }
  const bool qknorm = a->epilogue == MMDIT_EPI_QKNORM;
  if (qknorm) {
    auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    MMDIT_REQUIRE(a->aux && a->qk_wq && a->qk_wk && !a->d_fp32 && !a->accumulate && a->split_k <= 1 &&
                      a->remap_rows == 0 && a->N % 192 == 0 && a->M > BLOCK_M && al(a->D) && al(a->aux) &&
                      al(a->qk_wq) && al(a->qk_wk) && a->ldd % 8 == 0 && a->ld_aux % 8 == 0 &&
                      (!a->bias || (a->bias_fp32 && al(a->bias))) && (!a->rope_cos == !a->rope_sin) &&
                      (!a->rope_cos || (a->qk_tokens > 0 && al(a->rope_cos) && al(a->rope_sin))),
                  MMDIT_ERR_UNSUPPORTED,
                  "gemm: the QK-norm epilogue needs aux [M, 2N/3], bf16 D, N = 3*d with d %% 64 == 0, M > 128, "
                  "fp32 16-byte aligned norm weights / RoPE tables");
  }
  const int sms = num_sms();
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = (int)a->M; p.N = (int)a->N; p.K = (int)a->K;
  p.a_mn = a->a_major ? 1 : 0;
  p.b_mn = a->b_major ? 1 : 0;
  // split-K GEMMs (explicit slices / fp32 accumulate) fill the machine through the split, not through
  // narrower tiles: the split planners (here and ops._plan_split) assume 128x256 tiles
  const bool will_split = a->split_k > 1 || (a->split_k <= 0 && a->accumulate && a->d_fp32 && a->K >= 8 * BLOCK_K);
  p.block_n = (swiglu || qknorm || swiglu_bwd) ? 256 : a->force_block_n ? a->force_block_n
              : (will_split && a->N > 128) ? 256 : pick_block_n(a->M, a->N, sms);
  MMDIT_REQUIRE(p.block_n == 64 || p.block_n == 128 || p.block_n == 256, MMDIT_ERR_ARG,
                "gemm: block_n %d", p.block_n);
  // CTA pairs (cta_group::2) for every 256-wide tile with more than one 128-row block
  static int env_pair = -1;
  if (env_pair < 0) {
    const char* e = getenv("MMDIT_GEMM_PAIR");
    env_pair = e ? atoi(e) : 1;
  }
  const bool pair = swiglu || qknorm || swiglu_bwd || (env_pair && p.block_n == 256 && a->M > BLOCK_M && !(a->reserved & 16));
  const int workers = pair ? sms / 2 : sms;  // persistent work units running concurrently
  const int stage_bytes = A_STAGE_BYTES + (pair ? p.block_n / 2 : p.block_n) * BLOCK_K * 2;
  const int epi_bytes = swiglu_bwd ? EPI_STAGE_BYTES_SWIGLU_BWD : EPI_STAGE_BYTES;
  p.stages = (SMEM_BUDGET - 1024 - BAR_BYTES - epi_bytes) / stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  {
    static int env_stages = -1;  // perf experiments: MMDIT_GEMM_STAGES caps the smem ring depth
    if (env_stages < 0) {
      const char* e = getenv("MMDIT_GEMM_STAGES");
      env_stages = e ? atoi(e) : 0;
    }
    if (env_stages > 0 && p.stages > env_stages) p.stages = env_stages;
  }
  p.tiles_m = pair ? (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M) : (p.M + BLOCK_M - 1) / BLOCK_M;
  p.tiles_n = (p.N + p.block_n - 1) / p.block_n;   // SwiGLU: N/256 tiles of 128 gate + 128 up columns
  p.swiglu_half = swiglu ? p.N / 2 : 0;
  p.kb_total = (p.K + BLOCK_K - 1) / BLOCK_K;
  int split = a->split_k;
  const int tiles = p.tiles_m * p.tiles_n;
  if (split <= 0) {
    split = 1;
    if (a->accumulate && a->d_fp32 && p.kb_total >= 8) {
      // model: time ~ waves * (k-blocks per unit + epilogue), epilogue ~ 10 k-block times
      double best = 1e30;
      const int max_split = p.kb_total / 4 > 0 ? p.kb_total / 4 : 1;
      for (int s = 1; s <= max_split && s <= 64; ++s) {
        const int per = (p.kb_total + s - 1) / s;
        const long long units = (long long)tiles * ((p.kb_total + per - 1) / per);
        const long long waves = (units + workers - 1) / workers;
        const double t = (double)waves * (per + 10.0);
        if (t < best - 1e-9) { best = t; split = s; }
      }
    }
  }
  // split_k > 1 with accumulate=0: "slices" mode -- split s writes its partial product to
  // D + s*M*ldd (D must hold split_k such slices); the caller folds them (no atomics).
  MMDIT_REQUIRE(split == 1 || a->d_fp32, MMDIT_ERR_ARG, "gemm: split_k > 1 needs fp32 output");
  const bool slices = split > 1 && !a->accumulate;
  MMDIT_REQUIRE(!slices || (a->epilogue == MMDIT_EPI_NONE && !a->bias && !a->aux && a->remap_rows == 0),
                MMDIT_ERR_ARG, "gemm: split-K slices mode supports the plain fp32 epilogue only");
  if (split > p.kb_total) split = p.kb_total;
  p.kb_per_split = (p.kb_total + split - 1) / split;
  p.split_k = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;

  p.D = a->D; p.ldd = a->ldd; p.d_fp32 = a->d_fp32; p.accumulate = a->accumulate;
  p.epi = a->epilogue;
  p.bias = a->bias; p.bias_fp32 = a->bias_fp32;
  p.gate = static_cast<const bf16*>(a->gate);
  p.rows_per_gate = a->rows_per_gate; p.ld_gate = a->ld_gate;
  p.resid = static_cast<const bf16*>(a->resid); p.ldr = a->ldr;
  p.aux = static_cast<bf16*>(a->aux); p.ld_aux = a->ld_aux;
  p.remap_rows = a->remap_rows; p.remap_batch_rows = a->remap_batch_rows;
  p.remap_offset = a->remap_offset;
  p.debug = a->reserved;
  p.slice_stride = (split > 1 && !a->accumulate) ? (long long)a->M * a->ldd : 0;

  // Tensor maps. dims are innermost-first.
  {
    uint64_t dims[2], strides[1];
    uint32_t box[2];
    if (!p.a_mn) { dims[0] = a->K; dims[1] = a->M; box[0] = BLOCK_K; box[1] = BLOCK_M; }
    else         { dims[0] = a->M; dims[1] = a->K; box[0] = 64;      box[1] = BLOCK_K; }
    strides[0] = (uint64_t)a->lda * 2;
    int rc = encode_tmap(&p.tmA, a->A, 2, dims, strides, box, 2, true);
    if (rc) return rc;
    if (!p.b_mn) { dims[0] = a->K; dims[1] = a->N; box[0] = BLOCK_K; box[1] = (uint32_t)(pair ? p.block_n / 2 : p.block_n); }
    else         { dims[0] = a->N; dims[1] = a->K; box[0] = 64;      box[1] = BLOCK_K; }
    strides[0] = (uint64_t)a->ldb * 2;
    rc = encode_tmap(&p.tmB, a->B, 2, dims, strides, box, 2, true);
    if (rc) return rc;
  }

  const int smem_bytes = p.stages * stage_bytes + epi_bytes + BAR_BYTES + 1024;
  const int total_work = tiles * p.split_k;
  const int grid = pair ? 2 * (total_work < workers ? total_work : workers)
                        : (total_work < sms ? total_work : sms);

  // pick the epilogue variant: fast paths need whole 16-byte vectors everywhere
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool vec_ok = a->N % 8 == 0 && al16(a->D) && a->remap_rows == 0 &&
                      (a->d_fp32 ? a->ldd % 4 == 0 : a->ldd % 8 == 0);
  const bool bias_ok = !a->bias || (a->bias_fp32 && al16(a->bias));
  int ev = EV_GENERIC;
  if (swiglu) {
    ev = EV_SWIGLU_TMA;
  } else if (swiglu_bwd) {
    ev = EV_SWIGLU_BWD;
    p.colsum_partial = static_cast<float*>(a->colsum_partial);
  } else if (qknorm) {
    ev = EV_QKNORM_TMA;
    p.qk_wq = static_cast<const float*>(a->qk_wq);
    p.qk_wk = static_cast<const float*>(a->qk_wk);
    p.rope_cos = static_cast<const float*>(a->rope_cos);
    p.rope_sin = static_cast<const float*>(a->rope_sin);
    p.qk_d = (int)(a->N / 3);
    p.qk_tokens = a->qk_tokens;
    p.qk_eps = a->qk_eps;
  } else if (vec_ok) {
    if (a->d_fp32) {
      if (a->epilogue == MMDIT_EPI_NONE && !a->bias && !a->aux) {
        if (!a->accumulate) ev = EV_F32;
        else if (p.split_k > 1 || a->accumulate) ev = EV_F32_ATOMIC;  // red.add is also a valid "+="
      }
    } else if (a->epilogue == MMDIT_EPI_NONE && !a->aux) {
      static int env_tma = -1;
      if (env_tma < 0) {
        const char* e = getenv("MMDIT_GEMM_TMA_STORE");
        env_tma = e ? atoi(e) : 1;
      }
      if (env_tma && bias_ok && a->N >= 64) ev = EV_BF16_TMA;
      else if (!a->bias) ev = EV_BF16;
      else if (bias_ok) ev = EV_BF16_BIAS;
    } else if (a->epilogue == MMDIT_EPI_GATE_RESID && a->aux && bias_ok && al16(a->aux) &&
               al16(a->gate) && al16(a->resid) && a->ld_aux % 8 == 0 && a->ld_gate % 8 == 0 &&
               a->ldr % 8 == 0) {
      ev = EV_GATE;
    }
  }
  MMDIT_REQUIRE(p.slice_stride == 0 || ev == EV_F32, MMDIT_ERR_ALIGN,
                "gemm: split-K slices mode needs N %% 8 == 0 and 16-byte aligned D");
  if (ev == EV_BF16_TMA || ev == EV_SWIGLU_TMA || ev == EV_QKNORM_TMA || ev == EV_SWIGLU_BWD) {
    uint64_t dims[2] = {(uint64_t)(swiglu ? a->N / 2 : swiglu_bwd ? 2 * a->N : a->N), (uint64_t)a->M}, strides[1] = {(uint64_t)a->ldd * 2};
    uint32_t box[2] = {64, 32};
    int rc_d = encode_tmap(&p.tmD, a->D, 2, dims, strides, box, 2, true);
    if (rc_d) return rc_d;
    if ((swiglu && a->aux) || qknorm || swiglu_bwd) {
      dims[0] = (uint64_t)(qknorm ? a->N / 3 * 2 : swiglu_bwd ? 2 * a->N : a->N);
      strides[0] = (uint64_t)a->ld_aux * 2;
      rc_d = encode_tmap(&p.tmAux, a->aux, 2, dims, strides, box, 2, true);
      if (rc_d) return rc_d;
    }
  }
  int rc = 0;
#define LAUNCH_EV(EVV)                                                                 \
  case EVV:                                                                            \
    rc = pair ? launch_variant<EVV, true>(grid, smem_bytes, stream, p)                 \
              : launch_variant<EVV, false>(grid, smem_bytes, stream, p);               \
    break;
  switch (ev) {
    LAUNCH_EV(EV_GENERIC)
    LAUNCH_EV(EV_BF16)
    LAUNCH_EV(EV_BF16_BIAS)
    LAUNCH_EV(EV_GATE)
    LAUNCH_EV(EV_F32)
    LAUNCH_EV(EV_F32_ATOMIC)
    LAUNCH_EV(EV_BF16_TMA)
    case EV_SWIGLU_TMA:
      rc = launch_variant<EV_SWIGLU_TMA, true>(grid, smem_bytes, stream, p);
      break;
    case EV_SWIGLU_BWD:
      rc = launch_variant<EV_SWIGLU_BWD, true>(grid, smem_bytes, stream, p);
      break;
    case EV_QKNORM_TMA:
      rc = launch_variant<EV_QKNORM_TMA, true>(grid, smem_bytes, stream, p);
      break;
  }
#undef LAUNCH_EV
  if (rc) return rc;
  return check_launch("gemm_tcgen05_kernel");
}
