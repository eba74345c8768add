// Joint text+image softmax attention, backward, on tcgen05 tensor cores
// (the autograd of flash_attn_func at Attention.py:293; math in SURVEY App. E).
//
// One CTA = one (sample, head, 128-key tile); it loops over all query tiles.
// Everything is computed transposed so that each compute thread owns a key row:
//   S^T  = K Q^T            dP^T = V dO^T                 (TMEM, 128 cols each)
//   P^T  = exp2(S^T*c - LSE),   dS^T = P^T (dP^T - delta) * scale   -> bf16 smem
//   dV  += P^T dO           dK  += dS^T Q                 (TMEM, 64 cols each)
//   dQ_i = dS K   (A operand = dS^T read MN-major)        (TMEM, 64 cols)
// dQ tiles are added into an fp32 accumulator with 128-bit vector reductions
// (red.global.add.v4.f32); a small kernel converts it to bf16 afterwards.
//   warp 0      TMA producer (K,V once; Q_i,dO_i through a 2-deep ring)
//   warp 1      MMA issuer
//   warps 2..9  compute: thread = (key row, 64-query half)
//   warps 10..13 drain each finished dQ tile from TMEM into the fp32 accumulator (vector RED),
//               off the compute warps' critical path
#include <stdlib.h>

#include "common.cuh"
#include "mmdit_b200.h"

namespace mmdit {

constexpr int ATT_TILE = 128;
constexpr int ATT_HD = 64;
constexpr int ATT_TILE_BYTES = ATT_TILE * ATT_HD * 2;  // 16 KiB
constexpr int BWD_THREADS = 448;  // TMA warp, MMA warp, 8 compute warps, 4 dQ-drain warps
constexpr int BWD_SMEM = 14 * ATT_TILE_BYTES + 2 * 2 * 128 * 4 + 256;  // 226.25 KiB

int make_attn_tmap(CUtensorMap* map, const void* base, long long ld, int H, int rows, int B);

// Optional in-kernel timeline (debugging / tuning): one chosen CTA records (event id, clock64).
__device__ long long* g_bwd_timeline = nullptr;
__device__ int g_bwd_timeline_block = -1;
#define TL(slot, id)                                                                   \
  do {                                                                                 \
    if (tl) { const int s_ = (slot); tl[2 * s_] = (id); tl[2 * s_ + 1] = clock64(); }  \
  } while (0)

struct AttnBwdParams {
  CUtensorMap tmQ[2], tmK[2], tmV[2], tmdO[2];
  CUtensorMap tmdQ[2];  // fp32 dQ accumulator [B, T, H*64] viewed per stream, box 32 x 128, 128B swizzle
  bf16* dk[2];
  bf16* dv[2];
  long long ld_dk[2], ld_dv[2];
  const float* lse;    // [B,H,T]
  const float* delta;  // [B,H,T]
  float* dq_acc;       // [B, T, H*64] fp32
  int B, H, N, M;
  float scale, scale_log2;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

// First generation (one CTA per item); kept as the MMDIT_ATTN_BWD_V2=0 fallback.
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ AttnBwdParams p) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = smem + ATT_TILE_BYTES;
  uint8_t* sQ = smem + 2 * ATT_TILE_BYTES;    // [2]
  uint8_t* sdO = smem + 4 * ATT_TILE_BYTES;   // [2]
  uint8_t* sPt = smem + 6 * ATT_TILE_BYTES;    // [2 buffers][2 halves]
  uint8_t* sdSt = smem + 10 * ATT_TILE_BYTES;  // [2 buffers][2 halves]
  float* sLse = reinterpret_cast<float*>(smem + 14 * ATT_TILE_BYTES);  // [2][128]
  float* sDelta = sLse + 256;                                          // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDelta + 256);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;   // [2]
  uint64_t* qdo_empty = bars + 3;  // [2]
  uint64_t* st_full = bars + 5;
  uint64_t* pt_full = bars + 6;
  uint64_t* dq_full = bars + 7;
  uint64_t* dq_empty = bars + 8;
  uint64_t* buf_free = bars + 9;  // [2]: P^T/dS^T smem buffer no longer read by any MMA
  uint64_t* all_done = bars + 11; // every MMA of this CTA has retired
  uint64_t* stage_free = bars + 12;  // [2]: dQ staging (aliases the P^T buffer) read out by the TMA reduce
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  uint64_t* st_free = bars + 16;  // 256: every compute thread holds its S^T / dP^T values in registers

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.N + p.M;
  const int ntx = (p.N + ATT_TILE - 1) / ATT_TILE;
  const int ntc = (p.M + ATT_TILE - 1) / ATT_TILE;
  const int nt = ntx + ntc;
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int ks = kt < ntx ? 0 : 1;
  const int k_row0 = (ks == 0 ? kt : kt - ntx) * ATT_TILE;
  const int k_rows = ks == 0 ? p.N : p.M;
  const int k_valid = min(ATT_TILE, k_rows - k_row0);

  const int linear_block = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  long long* tl = nullptr;   // [0,128): MMA warp events, [128,256): compute warp 2 events
  if (g_bwd_timeline && linear_block == g_bwd_timeline_block && lane == 0) {
    if (warp == 1) tl = g_bwd_timeline;
    if (warp == 2) tl = g_bwd_timeline + 128;
  }
  int tls = 0;
  TL(tls++, 1);
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(kv_full, 1);
      for (int s = 0; s < 2; ++s) { mbar_init(&qdo_full[s], 1); mbar_init(&qdo_empty[s], 1); }
      mbar_init(st_full, 1);
      mbar_init(pt_full, 256);
      mbar_init(st_free, 256);
      mbar_init(dq_full, 1);
      mbar_init(dq_empty, 128);   // the 4 drain warps
      mbar_init(&buf_free[0], 1);
      mbar_init(&buf_free[1], 1);
      mbar_init(all_done, 1);
      mbar_init(&stage_free[0], 1);
      mbar_init(&stage_free[1], 1);
      mbar_fence_init();
      // K, V and the first two (Q, dO) tiles fly while TMEM is being allocated
      mbar_expect_tx(kv_full, 2 * ATT_TILE_BYTES);
      tma_load_4d(sK, &p.tmK[ks], kv_full, 0, h, k_row0, b);
      tma_load_4d(sV, &p.tmV[ks], kv_full, 0, h, k_row0, b);
      for (int i = 0; i < 2 && i < nt; ++i) {
        const int qs = i < ntx ? 0 : 1;
        const int row0 = (qs == 0 ? i : i - ntx) * ATT_TILE;
        mbar_expect_tx(&qdo_full[i], 2 * ATT_TILE_BYTES);
        tma_load_4d(sQ + i * ATT_TILE_BYTES, &p.tmQ[qs], &qdo_full[i], 0, h, row0, b);
        tma_load_4d(sdO + i * ATT_TILE_BYTES, &p.tmdO[qs], &qdo_full[i], 0, h, row0, b);
      }
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_St = tmem_base, tm_dPt = tmem_base + 128, tm_dV = tmem_base + 256,
                 tm_dK = tmem_base + 320, tm_dQ = tmem_base + 384;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 2; i < nt; ++i) {   // tiles 0 and 1 were issued in the prologue
        const int st = i & 1;
        const int qs = i < ntx ? 0 : 1;
        const int row0 = (qs == 0 ? i : i - ntx) * ATT_TILE;
        mbar_wait(&qdo_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&qdo_full[st], 2 * ATT_TILE_BYTES);
        tma_load_4d(sQ + st * ATT_TILE_BYTES, &p.tmQ[qs], &qdo_full[st], 0, h, row0, b);
        tma_load_4d(sdO + st * ATT_TILE_BYTES, &p.tmdO[qs], &qdo_full[st], 0, h, row0, b);
      }
    }
  } else if (warp == 1) {
    // the whole warp walks the loop (all lanes wait on the barriers); one elected lane issues the MMAs
    {
      // Software pipeline: S^T/dP^T of tile i+1 are issued as soon as the compute warps have read
      // tile i's, BEFORE the dV/dK/dQ MMAs of tile i, so exp/dS math of tile i+1 overlaps them.
      const uint32_t id_kn = make_idesc_bf16(128, 64, 0, 1);   // dV, dK
      const uint32_t id_nn = make_idesc_bf16(128, 64, 1, 1);   // dQ
      constexpr uint64_t kStepK = 32 >> 4, kStepMN = 2048 >> 4, kTile = ATT_TILE_BYTES >> 4;
      const uint64_t k_desc = desc_kmajor(smem_u32(sK), 0), v_desc = desc_kmajor(smem_u32(sV), 0);
      const uint64_t k_desc_mn = desc_mnmajor(smem_u32(sK), 0, ATT_TILE_BYTES);
      const uint64_t q_desc0 = desc_kmajor(smem_u32(sQ), 0), do_desc0 = desc_kmajor(smem_u32(sdO), 0);
      const uint64_t q_desc0_mn = desc_mnmajor(smem_u32(sQ), 0, ATT_TILE_BYTES);
      const uint64_t do_desc0_mn = desc_mnmajor(smem_u32(sdO), 0, ATT_TILE_BYTES);
      const uint64_t pt_desc0 = desc_kmajor(smem_u32(sPt), 0), dst_desc0 = desc_kmajor(smem_u32(sdSt), 0);
      const uint64_t dst_desc0_mn = desc_mnmajor(smem_u32(sdSt), 0, ATT_TILE_BYTES);
      auto issue_s = [&](int i) {
        const int st = i & 1;
        const int qs = i < ntx ? 0 : 1;
        const int row0 = (qs == 0 ? i : i - ntx) * ATT_TILE;
        const int nq = (min(ATT_TILE, (qs == 0 ? p.N : p.M) - row0) + 15) & ~15;  // queries that exist
        const uint32_t id_kq = make_idesc_bf16(128, nq, 0, 0);
        mbar_wait(&qdo_full[st], (i >> 1) & 1);
        tc_fence_after();
        TL(tls++, 10 + i);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tm_St, k_desc + k * kStepK, q_desc0 + st * kTile + k * kStepK, id_kq, k > 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tm_dPt, v_desc + k * kStepK, do_desc0 + st * kTile + k * kStepK, id_kq, k > 0);
          umma_commit(st_full);
        }
        __syncwarp();
        TL(tls++, 20 + i);
      };
      mbar_wait(kv_full, 0);
      issue_s(0);
      for (int i = 0; i < nt; ++i) {
        const int st = i & 1;
        const int qs = i < ntx ? 0 : 1;
        const int row0 = (qs == 0 ? i : i - ntx) * ATT_TILE;
        const int nq = (min(ATT_TILE, (qs == 0 ? p.N : p.M) - row0) + 15) & ~15;
        const uint64_t pt_desc = pt_desc0 + st * 2 * kTile, dst_desc = dst_desc0 + st * 2 * kTile;
        if (i + 1 < nt) {
          // S^T / dP^T of tile i sit in the compute threads' registers: the next tile's may be issued
          // now, while tile i's exponentials are still being computed
          mbar_wait(st_free, i & 1);
          tc_fence_after();
          issue_s(i + 1);
        }
        mbar_wait(pt_full, i & 1);
        tc_fence_after();
        TL(tls++, 30 + i);
        if (elect_one()) {
          for (int k = 0; k < nq / 16; ++k)
            umma_bf16(tm_dV, pt_desc + (k >> 2) * kTile + (k & 3) * kStepK,
                      do_desc0_mn + st * kTile + k * kStepMN, id_kn, (i > 0 || k > 0) ? 1u : 0u);
          for (int k = 0; k < nq / 16; ++k)
            umma_bf16(tm_dK, dst_desc + (k >> 2) * kTile + (k & 3) * kStepK,
                      q_desc0_mn + st * kTile + k * kStepMN, id_kn, (i > 0 || k > 0) ? 1u : 0u);
        }
        __syncwarp();
        TL(tls++, 40 + i);
        if (i > 0) {
          mbar_wait(dq_empty, (i - 1) & 1);
          tc_fence_after();
        }
        TL(tls++, 50 + i);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tm_dQ, dst_desc0_mn + st * 2 * kTile + k * kStepMN, k_desc_mn + k * kStepMN, id_nn, k > 0);
          umma_commit(&qdo_empty[st]);
          umma_commit(&buf_free[st]);
          umma_commit(dq_full);
          if (i == nt - 1) umma_commit(all_done);
        }
        __syncwarp();
        TL(tls++, 60 + i);
      }
    }
  } else if (warp >= 10) {
    // ------------------------------------------------------------ dQ drain warps
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // query row of the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    for (int i = 0; i < nt; ++i) {
      const int qs = i < ntx ? 0 : 1;
      const int row0 = (qs == 0 ? i : i - ntx) * ATT_TILE;
      const int q_valid = min(ATT_TILE, (qs == 0 ? p.N : p.M) - row0);
      const int t0 = (qs == 0 ? 0 : p.N) + row0;
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
      uint32_t q0[32], q1[32];
      tmem_ld32(tm_dQ + lane_off, q0);
      tmem_ld32(tm_dQ + lane_off + 32, q1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(dq_empty);       // TMEM tile is in registers: the MMA warp may overwrite it
      // dq_full(i) also means every MMA that read P^T buffer (i & 1) has retired: reuse it as the
      // staging tile (two 128-row x 128-byte halves, 128B swizzle) of an asynchronous TMA reduce-add.
      uint8_t* stage = sPt + (i & 1) * 2 * ATT_TILE_BYTES;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        *reinterpret_cast<uint4*>(stage + r * 128 + ((g ^ (r & 7)) << 4)) =
            make_uint4(q0[4 * g], q0[4 * g + 1], q0[4 * g + 2], q0[4 * g + 3]);
        *reinterpret_cast<uint4*>(stage + ATT_TILE_BYTES + r * 128 + ((g ^ (r & 7)) << 4)) =
            make_uint4(q1[4 * g], q1[4 * g + 1], q1[4 * g + 2], q1[4 * g + 3]);
      }
      fence_proxy_async_smem();
      named_bar_sync(2, 128);
      if (warp == 10 && lane == 0) {
        tma_reduce_add_3d(&p.tmdQ[qs], stage, h * ATT_HD, row0, b);
        tma_reduce_add_3d(&p.tmdQ[qs], stage + ATT_TILE_BYTES, h * ATT_HD + 32, row0, b);
        tma_commit_group();
        tma_wait_group_read0();          // smem has been read: compute warps may refill the buffer
        mbar_arrive(&stage_free[i & 1]);
      }
      (void)q_valid; (void)t0;
    }
    if (warp == 10 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
    // ---------------------------------------------------------------- compute
    const int cw = warp - 2;
    const int quarter = warp & 3;
    const int hf = cw >> 2;             // which 64-wide half of the query tile
    const int r = quarter * 32 + lane;  // key row in the tile == TMEM lane
    const int ct = cw * 32 + lane;      // 0..255
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const bool k_ok = r < k_valid;
    const float sl2 = p.scale_log2;
    const long long lse_base = ((long long)b * p.H + h) * T;
    float pre_lse = INFINITY, pre_del = 0.f;
    if (ct < 128) {   // first query tile
      const int qv0 = min(ATT_TILE, (ntx > 0 ? p.N : p.M));
      if (ct < qv0) {
        pre_lse = p.lse[lse_base + ct] * 1.4426950408889634f;
        pre_del = p.delta[lse_base + ct];
      }
    }
    for (int i = 0; i < nt; ++i) {
      const int qs = i < ntx ? 0 : 1;
      const int row0 = (qs == 0 ? i : i - ntx) * ATT_TILE;
      const int q_valid = min(ATT_TILE, (qs == 0 ? p.N : p.M) - row0);
      const int t0 = (qs == 0 ? 0 : p.N) + row0;
      float* lse_s = sLse + (i & 1) * 128;
      float* del_s = sDelta + (i & 1) * 128;
      if (ct < 128) {   // values were fetched from global one iteration ago
        lse_s[ct] = pre_lse;
        del_s[ct] = pre_del;
      }
      if (ct < 128 && i + 1 < nt) {   // prefetch for the next query tile; consumed after this tile's math
        const int qs1 = i + 1 < ntx ? 0 : 1;
        const int row1 = (qs1 == 0 ? i + 1 : i + 1 - ntx) * ATT_TILE;
        const int qv1 = min(ATT_TILE, (qs1 == 0 ? p.N : p.M) - row1);
        const int t1 = (qs1 == 0 ? 0 : p.N) + row1;
        const bool ok = ct < qv1;
        pre_lse = ok ? p.lse[lse_base + t1 + ct] * 1.4426950408889634f : INFINITY;
        pre_del = ok ? p.delta[lse_base + t1 + ct] : 0.f;
      }
      named_bar_sync(1, 256);
      TL(tls++, 70 + i);
      mbar_wait(st_full, i & 1);
      tc_fence_after();
      TL(tls++, 80 + i);
      // P^T/dS^T buffer (i & 1): tile i-2's MMAs have retired and its dQ staging has been read out
      mbar_wait(&stage_free[i & 1], ((i >> 1) & 1) ^ 1);
      uint8_t* bufP = sPt + (i & 1) * 2 * ATT_TILE_BYTES;
      uint8_t* bufD = sdSt + (i & 1) * 2 * ATT_TILE_BYTES;
      const int nq = (q_valid + 15) & ~15;
      // chunks of 32 query columns this thread will load from TMEM (warp-uniform)
      const int n_ld = (quarter * 32 >= k_valid) ? 0 : (hf * 64 >= nq ? 0 : (hf * 64 + 32 >= nq ? 1 : 2));
      if (n_ld == 0) {
        tc_fence_before();
        mbar_arrive(st_free);
      }
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        if (hf * 64 + c * 32 >= nq) break;  // warp-uniform: these query columns do not exist
        if (quarter * 32 >= k_valid) {
          // partial key tile: no key row of this warp exists -> P^T = dS^T = 0, no exp / TMEM traffic
          // (dS^T must be exact zeros: dQ sums over the key rows)
          uint8_t* prow = bufP + hf * ATT_TILE_BYTES + r * 128;
          uint8_t* drow = bufD + hf * ATT_TILE_BYTES + r * 128;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int off = ((c * 4 + g) ^ (r & 7)) << 4;
            *reinterpret_cast<uint4*>(prow + off) = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(drow + off) = make_uint4(0u, 0u, 0u, 0u);
          }
          continue;
        }
        uint32_t s[32], dp[32];
        tmem_ld32(tm_St + lane_off + hf * 64 + c * 32, s);
        tmem_ld32(tm_dPt + lane_off + hf * 64 + c * 32, dp);
        tmem_ld_wait();
        if (c == n_ld - 1) {   // last TMEM read of this tile by this thread
          tc_fence_before();
          mbar_arrive(st_free);
        }
        uint8_t* prow = bufP + hf * ATT_TILE_BYTES + r * 128;
        uint8_t* drow = bufD + hf * ATT_TILE_BYTES + r * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float pe[8], de[8];
          const int col0 = hf * 64 + c * 32 + g * 8;
          const float4 l0 = *reinterpret_cast<const float4*>(lse_s + col0);
          const float4 l1 = *reinterpret_cast<const float4*>(lse_s + col0 + 4);
          const float4 e0 = *reinterpret_cast<const float4*>(del_s + col0);
          const float4 e1 = *reinterpret_cast<const float4*>(del_s + col0 + 4);
          const float lv[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
          const float dv8[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float pv = k_ok ? ex2_approx(fmaf(__uint_as_float(s[g * 8 + j]), sl2, -lv[j])) : 0.f;
            pe[j] = pv;
            de[j] = pv * (__uint_as_float(dp[g * 8 + j]) - dv8[j]) * p.scale;
          }
          uint4 u, w;
          u.x = pack_bf16x2(pe[0], pe[1]); u.y = pack_bf16x2(pe[2], pe[3]);
          u.z = pack_bf16x2(pe[4], pe[5]); u.w = pack_bf16x2(pe[6], pe[7]);
          w.x = pack_bf16x2(de[0], de[1]); w.y = pack_bf16x2(de[2], de[3]);
          w.z = pack_bf16x2(de[4], de[5]); w.w = pack_bf16x2(de[6], de[7]);
          const int off = ((c * 4 + g) ^ (r & 7)) << 4;
          *reinterpret_cast<uint4*>(prow + off) = u;
          *reinterpret_cast<uint4*>(drow + off) = w;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(pt_full);
      TL(tls++, 90 + i);
    }
    mbar_wait(all_done, 0);   // every MMA (incl. the last dV/dK updates) has retired
    tc_fence_after();
    {
      uint32_t a[32], c2[32];
      tmem_ld32(tm_dV + lane_off + hf * 32, a);
      tmem_ld32(tm_dK + lane_off + hf * 32, c2);
      tmem_ld_wait();
      if (k_ok) {
        const long long grow = (long long)b * k_rows + k_row0 + r;
        bf16* dvp = p.dv[ks] + grow * p.ld_dv[ks] + h * ATT_HD + hf * 32;
        bf16* dkp = p.dk[ks] + grow * p.ld_dk[ks] + h * ATT_HD + hf * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float v[8], w[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[j] = __uint_as_float(a[g * 8 + j]);
            w[j] = __uint_as_float(c2[g * 8 + j]);
          }
          store8(dvp + g * 8, v);
          store8(dkp + g * 8, w);
        }
      }
    }
  }
  TL(tls++, 200);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// Second generation: PERSISTENT CTAs (one per SM) that walk work items (sample, head, 128-key tile).
// The first-generation kernel above starts a CTA per item: barrier init + TMEM allocation + the first
// TMA round trip (~2800 cycles before the first exponential) and the drain of dV / dK after the last
// tile (~2400 cycles) are a third of a 16k-cycle item at 4 query tiles (T = 410).  Here
//   * TMEM is allocated and the barriers are initialised once per CTA,
//   * K / V are double-buffered over ITEMS and Q / dO ride a 2-deep ring over all tiles of all items, so
//     the next item's operands are in shared memory before the current item ends,
//   * the MMA warp issues S^T / dP^T of the next item's first tile as soon as the compute threads hold
//     the current (last) tile in registers: the dV / dK / dQ MMAs of the last tile, the dV / dK drain
//     and the dQ reduction of one item overlap the first exponentials of the next,
//   * P^T goes to the dV MMA through tensor memory (A operand in TMEM): its shared-memory buffers are
//     what pays for the second K / V stage.
//   warp 0       TMA producer          warp 1       MMA issuer
//   warps 2..9   compute: thread = (key row, 64-query half); also drain dV / dK at the end of an item
//   warps 10..13 drain each finished dQ tile from TMEM through a staging tile into the fp32
//                accumulator (TMA reduce-add)
constexpr int BWD2_SMEM = 14 * ATT_TILE_BYTES + 2 * 2 * 128 * 4 + 256;

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd2_kernel(const __grid_constant__ AttnBwdParams p, int items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;                            // [2] over items
  uint8_t* sV = smem + 2 * ATT_TILE_BYTES;       // [2] over items
  uint8_t* sQ = smem + 4 * ATT_TILE_BYTES;       // [2] ring over tiles
  uint8_t* sdO = smem + 6 * ATT_TILE_BYTES;      // [2] ring over tiles
  uint8_t* sdSt = smem + 8 * ATT_TILE_BYTES;     // [2 buffers][2 halves]
  uint8_t* sStage = smem + 12 * ATT_TILE_BYTES;  // dQ staging: two 128-row x 128-byte halves
  float* sLse = reinterpret_cast<float*>(smem + 14 * ATT_TILE_BYTES);  // [2][128]
  float* sDelta = sLse + 256;                                          // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDelta + 256);
  uint64_t* kv_full = bars + 0;     // [2]
  uint64_t* kv_empty = bars + 2;    // [2] every MMA of the item in this slot has retired
  uint64_t* qdo_full = bars + 4;    // [2]
  uint64_t* qdo_empty = bars + 6;   // [2]
  uint64_t* st_full = bars + 8;
  uint64_t* st_free = bars + 9;     // 256: every compute thread holds its S^T / dP^T values in registers
  uint64_t* pt_full = bars + 10;    // 256: P^T in tensor memory, dS^T in shared memory
  uint64_t* dv_done = bars + 11;    // the dV MMAs reading P^T from TMEM have retired
  uint64_t* buf_free = bars + 12;   // [2] dS^T buffer no longer read by any MMA
  uint64_t* dq_full = bars + 14;
  uint64_t* dq_empty = bars + 15;   // 128: the dQ tile has left TMEM
  uint64_t* item_done = bars + 16;  // every MMA of the item has retired: dV / dK are final
  uint64_t* dvdk_free = bars + 17;  // 256: dV / dK of the finished item have left TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.N + p.M;
  const int ntx = (p.N + ATT_TILE - 1) / ATT_TILE;
  const int ntc = (p.M + ATT_TILE - 1) / ATT_TILE;
  const int nt = ntx + ntc;
  const int BH = p.B * p.H;
  const int n_items = items > (int)blockIdx.x ? (items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  // item n of this CTA -> (sample, head, key tile).  The key tiles of one (sample, head) run on
  // neighbouring CTAs at the same time, so its Q / dO tiles are read from HBM once and its dQ rows are
  // reduced while they sit in the L2 (key-tile-major numbering re-reads them nt times: 853 vs 738 us at
  // T = 1178).  When the grid is a multiple of nt the tile index is rotated by n, otherwise a CTA would
  // see the same (possibly partial, cheaper) key tile in every item.
  const bool rotate = ((int)gridDim.x % nt) == 0;
  auto item_of = [&](int n, int& b, int& h, int& kt) {
    const int id = (int)blockIdx.x + n * (int)gridDim.x;
    const int bh = id / nt;
    kt = id - bh * nt;
    if (rotate) kt = (kt + n) % nt;
    h = bh % p.H;
    b = bh / p.H;
  };
  // tile i of the joint sequence -> (stream, first row, rows that exist)
  auto tile_of = [&](int i, int& s_, int& row0, int& valid) {
    s_ = i < ntx ? 0 : 1;
    row0 = (s_ == 0 ? i : i - ntx) * ATT_TILE;
    valid = min(ATT_TILE, (s_ == 0 ? p.N : p.M) - row0);
  };

  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1);
        mbar_init(&qdo_full[s], 1); mbar_init(&qdo_empty[s], 1);
        mbar_init(&buf_free[s], 1);
      }
      mbar_init(st_full, 1);
      mbar_init(st_free, 256);
      mbar_init(pt_full, 256);
      mbar_init(dv_done, 1);
      mbar_init(dq_full, 1);
      mbar_init(dq_empty, 128);
      mbar_init(item_done, 1);
      mbar_init(dvdk_free, 256);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_St = tmem_base, tm_dPt = tmem_base + 128, tm_dV = tmem_base + 256,
                 tm_dK = tmem_base + 320, tm_dQ = tmem_base + 384, tm_Pt = tmem_base + 448;
  pdl_wait();   // barrier init and the TMEM allocation overlapped the previous kernel's tail

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // (the whole warp walks the loop; one elected lane issues)
    auto load_kv = [&](int n) {
      int b, h, kt;
      item_of(n, b, h, kt);
      int ks, k_row0, k_valid;
      tile_of(kt, ks, k_row0, k_valid);
      const int slot = n & 1;
      mbar_wait(&kv_empty[slot], ((n >> 1) & 1) ^ 1);   // the item two back (same slot) has retired
      if (elect_one()) {
        mbar_expect_tx(&kv_full[slot], 2 * ATT_TILE_BYTES);
        tma_load_4d(sK + slot * ATT_TILE_BYTES, &p.tmK[ks], &kv_full[slot], 0, h, k_row0, b);
        tma_load_4d(sV + slot * ATT_TILE_BYTES, &p.tmV[ks], &kv_full[slot], 0, h, k_row0, b);
      }
      __syncwarp();
    };
    int G = 0;
    if (n_items > 0) load_kv(0);
    for (int n = 0; n < n_items; ++n) {
      int b, h, kt;
      item_of(n, b, h, kt);
      for (int i = 0; i < nt; ++i, ++G) {
        int qs, row0, qv;
        tile_of(i, qs, row0, qv);
        const int st = G & 1;
        mbar_wait(&qdo_empty[st], ((G >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&qdo_full[st], 2 * ATT_TILE_BYTES);
          tma_load_4d(sQ + st * ATT_TILE_BYTES, &p.tmQ[qs], &qdo_full[st], 0, h, row0, b);
          tma_load_4d(sdO + st * ATT_TILE_BYTES, &p.tmdO[qs], &qdo_full[st], 0, h, row0, b);
        }
        __syncwarp();
        // K / V of the NEXT item, one tile into this item: its slot's previous occupant (the item before
        // this one) retires about now, long before the next item's first S^T needs the tiles
        if (i == (nt > 1 ? 1 : 0) && n + 1 < n_items) load_kv(n + 1);
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer
    // the whole warp walks the loop (all lanes wait on the barriers); one elected lane issues the MMAs
    const uint32_t id_kn = make_idesc_bf16(128, 64, 0, 1);   // dV, dK
    const uint32_t id_nn = make_idesc_bf16(128, 64, 1, 1);   // dQ
    constexpr uint64_t kStepK = 32 >> 4, kStepMN = 2048 >> 4, kTile = ATT_TILE_BYTES >> 4;
    const uint64_t k_desc0 = desc_kmajor(smem_u32(sK), 0), v_desc0 = desc_kmajor(smem_u32(sV), 0);
    const uint64_t k_desc0_mn = desc_mnmajor(smem_u32(sK), 0, ATT_TILE_BYTES);
    const uint64_t q_desc0 = desc_kmajor(smem_u32(sQ), 0), do_desc0 = desc_kmajor(smem_u32(sdO), 0);
    const uint64_t q_desc0_mn = desc_mnmajor(smem_u32(sQ), 0, ATT_TILE_BYTES);
    const uint64_t do_desc0_mn = desc_mnmajor(smem_u32(sdO), 0, ATT_TILE_BYTES);
    const uint64_t dst_desc0 = desc_kmajor(smem_u32(sdSt), 0);
    const uint64_t dst_desc0_mn = desc_mnmajor(smem_u32(sdSt), 0, ATT_TILE_BYTES);
    // S^T = K Q^T and dP^T = V dO^T of global tile number g (query tile i of the item in K / V slot kvs)
    auto issue_s = [&](int g, int i, int kvs) {
      const int st = g & 1;
      int qs, row0, qv;
      tile_of(i, qs, row0, qv);
      const int nq = (qv + 15) & ~15;   // queries that exist
      const uint32_t id_kq = make_idesc_bf16(128, nq, 0, 0);
      mbar_wait(&qdo_full[st], (g >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tm_St, k_desc0 + kvs * kTile + k * kStepK, q_desc0 + st * kTile + k * kStepK, id_kq, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tm_dPt, v_desc0 + kvs * kTile + k * kStepK, do_desc0 + st * kTile + k * kStepK, id_kq, k > 0);
        umma_commit(st_full);
      }
      __syncwarp();
    };
    int G = 0;
    if (n_items > 0) {
      mbar_wait(&kv_full[0], 0);
      issue_s(0, 0, 0);
    }
    for (int n = 0; n < n_items; ++n) {
      const int kvs = n & 1;
      for (int i = 0; i < nt; ++i, ++G) {
        const int st = G & 1;
        int qs, row0, qv;
        tile_of(i, qs, row0, qv);
        const int nq = (qv + 15) & ~15;
        const uint64_t dst_desc = dst_desc0 + st * 2 * kTile;
        // S^T / dP^T of tile G sit in the compute threads' registers: the next tile's (of this item or
        // of the next one) may be issued now, while tile G's exponentials are still being computed
        if (i + 1 < nt) {
          mbar_wait(st_free, G & 1);
          tc_fence_after();
          issue_s(G + 1, i + 1, kvs);
        } else if (n + 1 < n_items) {
          mbar_wait(st_free, G & 1);
          tc_fence_after();
          mbar_wait(&kv_full[(n + 1) & 1], ((n + 1) >> 1) & 1);
          issue_s(G + 1, 0, (n + 1) & 1);
        }
        mbar_wait(pt_full, G & 1);
        tc_fence_after();
        if (i == 0 && n > 0) {
          mbar_wait(dvdk_free, (n - 1) & 1);   // the previous item's dV / dK have been read out
          tc_fence_after();
        }
        if (elect_one()) {
          for (int k = 0; k < nq / 16; ++k)
            umma_bf16_ts(tm_dV, tm_Pt + k * 8, do_desc0_mn + st * kTile + k * kStepMN, id_kn,
                         (i > 0 || k > 0) ? 1u : 0u);
          umma_commit(dv_done);   // P^T in TMEM may be overwritten once these retire
          for (int k = 0; k < nq / 16; ++k)
            umma_bf16(tm_dK, dst_desc + (k >> 2) * kTile + (k & 3) * kStepK,
                      q_desc0_mn + st * kTile + k * kStepMN, id_kn, (i > 0 || k > 0) ? 1u : 0u);
        }
        __syncwarp();
        if (G > 0) {
          mbar_wait(dq_empty, (G - 1) & 1);
          tc_fence_after();
        }
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tm_dQ, dst_desc0_mn + st * 2 * kTile + k * kStepMN, k_desc0_mn + kvs * kTile + k * kStepMN,
                      id_nn, k > 0);
          umma_commit(&qdo_empty[st]);
          umma_commit(&buf_free[st]);
          umma_commit(dq_full);
          if (i == nt - 1) {
            umma_commit(item_done);
            umma_commit(&kv_empty[kvs]);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp >= 10) {
    // ------------------------------------------------------------ dQ drain warps
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // query row of the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    int G = 0;
    for (int n = 0; n < n_items; ++n) {
      int b, h, kt;
      item_of(n, b, h, kt);
      for (int i = 0; i < nt; ++i, ++G) {
        int qs, row0, qv;
        tile_of(i, qs, row0, qv);
        mbar_wait(dq_full, G & 1);
        tc_fence_after();
        uint32_t q0[32], q1[32];
        tmem_ld32(tm_dQ + lane_off, q0);
        tmem_ld32(tm_dQ + lane_off + 32, q1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(dq_empty);       // TMEM tile is in registers: the MMA warp may overwrite it
        named_bar_sync(2, 128);      // the previous tile's reduce has read the staging tile (see below)
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          *reinterpret_cast<uint4*>(sStage + r * 128 + ((g ^ (r & 7)) << 4)) =
              make_uint4(q0[4 * g], q0[4 * g + 1], q0[4 * g + 2], q0[4 * g + 3]);
          *reinterpret_cast<uint4*>(sStage + ATT_TILE_BYTES + r * 128 + ((g ^ (r & 7)) << 4)) =
              make_uint4(q1[4 * g], q1[4 * g + 1], q1[4 * g + 2], q1[4 * g + 3]);
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (warp == 10 && lane == 0) {
          tma_reduce_add_3d(&p.tmdQ[qs], sStage, h * ATT_HD, row0, b);
          tma_reduce_add_3d(&p.tmdQ[qs], sStage + ATT_TILE_BYTES, h * ATT_HD + 32, row0, b);
          tma_commit_group();
          tma_wait_group_read0();    // before this thread reaches the next tile's first barrier
        }
      }
    }
    if (warp == 10 && lane == 0) tma_wait_group0();
  } else {
    // ---------------------------------------------------------------- compute
    const int cw = warp - 2;
    const int quarter = warp & 3;
    const int hf = cw >> 2;             // which 64-wide half of the query tile
    const int r = quarter * 32 + lane;  // key row in the tile == TMEM lane
    const int ct = cw * 32 + lane;      // 0..255
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const float sl2 = p.scale_log2;
    // lse / delta of the query rows of global tile (n, i), fetched one tile ahead by threads 0..127
    auto fetch = [&](int n, int i, float& l, float& dl) {
      int b, h, kt;
      item_of(n, b, h, kt);
      int qs, row0, qv;
      tile_of(i, qs, row0, qv);
      const long long base = ((long long)b * p.H + h) * T + (qs == 0 ? 0 : p.N) + row0;
      const bool ok = ct < qv;
      l = ok ? p.lse[base + ct] * 1.4426950408889634f : INFINITY;
      dl = ok ? p.delta[base + ct] : 0.f;
    };
    // dV / dK of item m are final once every MMA of the item has retired; they leave TMEM for global
    // memory AFTER the first tile of item m+1 has been computed (the wait for the last MMAs and the
    // stores hide behind that tile's exponentials; only the MMA warp's first dV / dK of item m+1 wait)
    auto drain_dvdk = [&](int m) {
      int b, h, kt;
      item_of(m, b, h, kt);
      int ks, k_row0, k_valid;
      tile_of(kt, ks, k_row0, k_valid);
      const int k_rows = ks == 0 ? p.N : p.M;
      mbar_wait(item_done, m & 1);
      tc_fence_after();
      uint32_t a[32], c2[32];
      tmem_ld32(tm_dV + lane_off + hf * 32, a);
      tmem_ld32(tm_dK + lane_off + hf * 32, c2);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(dvdk_free);      // the next item's first dV / dK MMAs may overwrite the accumulators
      if (r < k_valid) {
        const long long grow = (long long)b * k_rows + k_row0 + r;
        bf16* dvp = p.dv[ks] + grow * p.ld_dv[ks] + h * ATT_HD + hf * 32;
        bf16* dkp = p.dk[ks] + grow * p.ld_dk[ks] + h * ATT_HD + hf * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float v[8], w[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[j] = __uint_as_float(a[g * 8 + j]);
            w[j] = __uint_as_float(c2[g * 8 + j]);
          }
          store8(dvp + g * 8, v);
          store8(dkp + g * 8, w);
        }
      }
    };
    float pre_lse = INFINITY, pre_del = 0.f;
    if (ct < 128 && n_items > 0) fetch(0, 0, pre_lse, pre_del);
    int G = 0;
    for (int n = 0; n < n_items; ++n) {
      int b, h, kt;
      item_of(n, b, h, kt);
      int ks, k_row0, k_valid;
      tile_of(kt, ks, k_row0, k_valid);
      const bool k_ok = r < k_valid;
      for (int i = 0; i < nt; ++i, ++G) {
        int qs, row0, q_valid;
        tile_of(i, qs, row0, q_valid);
        float* lse_s = sLse + (G & 1) * 128;
        float* del_s = sDelta + (G & 1) * 128;
        if (ct < 128) {   // values were fetched from global one tile ago
          lse_s[ct] = pre_lse;
          del_s[ct] = pre_del;
          if (i + 1 < nt) fetch(n, i + 1, pre_lse, pre_del);
          else if (n + 1 < n_items) fetch(n + 1, 0, pre_lse, pre_del);
        }
        named_bar_sync(1, 256);
        mbar_wait(st_full, G & 1);
        tc_fence_after();
        // dS^T buffer (G & 1): the dK / dQ MMAs of tile G-2 have retired
        mbar_wait(&buf_free[G & 1], ((G >> 1) & 1) ^ 1);
        uint8_t* bufD = sdSt + (G & 1) * 2 * ATT_TILE_BYTES;
        const int nq = (q_valid + 15) & ~15;
        uint32_t ppk[32];   // this thread's 64 P^T values as packed bf16 pairs
#pragma unroll
        for (int q = 0; q < 32; ++q) ppk[q] = 0u;
        // chunks of 32 query columns this thread will load from TMEM (warp-uniform)
        const int n_ld = (quarter * 32 >= k_valid) ? 0 : (hf * 64 >= nq ? 0 : (hf * 64 + 32 >= nq ? 1 : 2));
        if (n_ld == 0) {
          tc_fence_before();
          mbar_arrive(st_free);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (hf * 64 + c * 32 >= nq) break;  // warp-uniform: these query columns do not exist
          uint8_t* drow = bufD + hf * ATT_TILE_BYTES + r * 128;
          if (quarter * 32 >= k_valid) {
            // partial key tile: no key row of this warp exists -> P^T = dS^T = 0, no exp / TMEM traffic
            // (dS^T must be exact zeros: dQ sums over the key rows)
#pragma unroll
            for (int g = 0; g < 4; ++g)
              *reinterpret_cast<uint4*>(drow + (((c * 4 + g) ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
            continue;
          }
          uint32_t s[32], dp[32];
          tmem_ld32(tm_St + lane_off + hf * 64 + c * 32, s);
          tmem_ld32(tm_dPt + lane_off + hf * 64 + c * 32, dp);
          tmem_ld_wait();
          if (c == n_ld - 1) {   // last TMEM read of this tile by this thread
            tc_fence_before();
            mbar_arrive(st_free);
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float pe[8], de[8];
            const int col0 = hf * 64 + c * 32 + g * 8;
            const float4 l0 = *reinterpret_cast<const float4*>(lse_s + col0);
            const float4 l1 = *reinterpret_cast<const float4*>(lse_s + col0 + 4);
            const float4 e0 = *reinterpret_cast<const float4*>(del_s + col0);
            const float4 e1 = *reinterpret_cast<const float4*>(del_s + col0 + 4);
            const float lv[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
            const float dv8[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float pv = k_ok ? ex2_approx(fmaf(__uint_as_float(s[g * 8 + j]), sl2, -lv[j])) : 0.f;
              pe[j] = pv;
              de[j] = pv * (__uint_as_float(dp[g * 8 + j]) - dv8[j]) * p.scale;
            }
            ppk[c * 16 + g * 4] = pack_bf16x2(pe[0], pe[1]);
            ppk[c * 16 + g * 4 + 1] = pack_bf16x2(pe[2], pe[3]);
            ppk[c * 16 + g * 4 + 2] = pack_bf16x2(pe[4], pe[5]);
            ppk[c * 16 + g * 4 + 3] = pack_bf16x2(pe[6], pe[7]);
            uint4 w;
            w.x = pack_bf16x2(de[0], de[1]); w.y = pack_bf16x2(de[2], de[3]);
            w.z = pack_bf16x2(de[4], de[5]); w.w = pack_bf16x2(de[6], de[7]);
            *reinterpret_cast<uint4*>(drow + (((c * 4 + g) ^ (r & 7)) << 4)) = w;
          }
        }
        if (hf * 64 < nq) {   // this warp's 64 query columns (or their first half) exist
          if (G > 0) {
            mbar_wait(dv_done, (G - 1) & 1);   // the previous tile's dV MMAs no longer read P^T
            tc_fence_after();
          }
          tmem_st32(tm_Pt + lane_off + hf * 32, ppk);
          tmem_st_wait();
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(pt_full);
        if (i == 0 && n > 0) drain_dvdk(n - 1);
      }
    }
    if (n_items > 0) drain_dvdk(n_items - 1);
  }
  if (warp >= 2) pdl_trigger();   // late trigger, see gemm_tcgen05.cu
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// delta[b,h,t] = sum_e dO[b,t,h,e] * O[b,t,h,e]; 8 lanes per (row, head).  Both streams in one launch;
// every thread also clears the 8 floats of the fp32 dQ accumulator that belong to its (row, head, part)
// -- the accumulator is exactly 8 floats per thread of this grid, so the separate 80 MB memset node
// (13 us per attention backward at the cfg2 shape) disappears.
struct AttnDeltaParams {
  const bf16* o[2];
  const bf16* d_o[2];
  long long ld_o[2], ld_do[2];
  long long rows[2];          // B * rows_per_sample
  int rows_per_sample[2];
  float* delta;               // [B, H, T]
  float* dq_acc;              // [B, T, H*64]
  int H, T, N;
};
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __grid_constant__ AttnDeltaParams p) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  const long long total0 = p.rows[0] * p.H * 8;
  const long long total = total0 + p.rows[1] * p.H * 8;
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool active = idx < total;
  const int s = (active && idx >= total0) ? 1 : 0;      // total0 is a multiple of 8: a head never straddles
  if (s) idx -= total0;
  const long long rh = active ? idx / 8 : 0;
  const int part = (int)(idx & 7);
  const long long row = rh / p.H;
  const int h = (int)(rh % p.H);
  float a[8], g[8];
  float acc = 0.f;
  if (active) {
    load8(p.o[s] + row * p.ld_o[s] + h * 64 + part * 8, a);
    load8(p.d_o[s] + row * p.ld_do[s] + h * 64 + part * 8, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += a[j] * g[j];
  }
#pragma unroll
  for (int off = 1; off < 8; off <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (active) {
    const long long bb = row / p.rows_per_sample[s];
    const int t = (s ? p.N : 0) + (int)(row % p.rows_per_sample[s]);
    if (part == 0) p.delta[(bb * p.H + h) * p.T + t] = acc;
    float* z = p.dq_acc + ((bb * p.T + t) * (long long)p.H + h) * 64 + part * 8;
    *reinterpret_cast<float4*>(z) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(z + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// dq (bf16, stream layout) = dq_acc (fp32 [B,T,H*64]) rows of one stream
__global__ void __launch_bounds__(256)
attn_dq_convert_kernel(const float* __restrict__ acc, bf16* __restrict__ dq, int B, int T, int t_off,
                       int rows_per_sample, int dmodel, long long ld_dq) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  const int groups = dmodel / 8;
  const long long total = (long long)B * rows_per_sample * groups;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(idx % groups) * 8;
    const long long row = idx / groups;
    const long long bb = row / rows_per_sample;
    const int tr = (int)(row % rows_per_sample);
    const float* src = acc + ((bb * T + t_off + tr) * (long long)dmodel) + col;
    const float4 x = *reinterpret_cast<const float4*>(src);
    const float4 y = *reinterpret_cast<const float4*>(src + 4);
    float v[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
    store8(dq + row * ld_dq + col, v);
  }
}

}  // namespace mmdit

using namespace mmdit;

extern "C" int mmdit_debug_attn_bwd_timeline(long long* buf, int block) {
  cudaError_t e = cudaMemcpyToSymbol(g_bwd_timeline, &buf, sizeof(buf));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_bwd_timeline_block, &block, sizeof(block));
  return (int)e;
}

extern "C" int mmdit_attn_bwd(const mmdit_attn_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MMDIT_REQUIRE(a, MMDIT_ERR_ARG, "attn_bwd: null args");
  MMDIT_REQUIRE(a->head_dim == 64, MMDIT_ERR_UNSUPPORTED, "attn_bwd: head_dim must be 64");
  MMDIT_REQUIRE(a->B > 0 && a->H > 0 && a->N > 0 && a->M >= 0, MMDIT_ERR_ARG, "attn_bwd: bad shape");
  MMDIT_REQUIRE(a->lse && a->delta && a->dq_acc, MMDIT_ERR_ARG,
                "attn_bwd: lse / delta / dq_acc workspaces required");
  AttnBwdParams p;
  memset(&p, 0, sizeof(p));
  const int rows[2] = {a->N, a->M};
  const int T = a->N + a->M;
  const int dmodel = a->H * 64;
  AttnDeltaParams dp;
  memset(&dp, 0, sizeof(dp));
  for (int s = 0; s < 2; ++s) {
    if (rows[s] == 0) continue;
    MMDIT_REQUIRE(a->q[s] && a->k[s] && a->v[s] && a->o[s] && a->d_o[s] && a->dk[s] && a->dv[s],
                  MMDIT_ERR_ARG, "attn_bwd: null pointer in stream %d", s);
    MMDIT_REQUIRE(a->ld_q[s] % 8 == 0 && a->ld_k[s] % 8 == 0 && a->ld_v[s] % 8 == 0 &&
                      a->ld_o[s] % 8 == 0 && a->ld_do[s] % 8 == 0 && a->ld_dq[s] % 8 == 0 &&
                      a->ld_dk[s] % 8 == 0 && a->ld_dv[s] % 8 == 0,
                  MMDIT_ERR_ALIGN, "attn_bwd: row strides must be multiples of 8 elements");
    int rc = make_attn_tmap(&p.tmQ[s], a->q[s], a->ld_q[s], a->H, rows[s], a->B);
    if (rc) return rc;
    rc = make_attn_tmap(&p.tmK[s], a->k[s], a->ld_k[s], a->H, rows[s], a->B);
    if (rc) return rc;
    rc = make_attn_tmap(&p.tmV[s], a->v[s], a->ld_v[s], a->H, rows[s], a->B);
    if (rc) return rc;
    rc = make_attn_tmap(&p.tmdO[s], a->d_o[s], a->ld_do[s], a->H, rows[s], a->B);
    if (rc) return rc;
    {
      uint64_t dims[3] = {(uint64_t)dmodel, (uint64_t)rows[s], (uint64_t)a->B};
      uint64_t strides[2] = {(uint64_t)dmodel * 4, (uint64_t)T * dmodel * 4};
      uint32_t box[3] = {32, ATT_TILE, 1};
      rc = encode_tmap(&p.tmdQ[s], a->dq_acc + (s == 0 ? 0 : (size_t)a->N * dmodel), 3, dims, strides, box, 4,
                       true);
      if (rc) return rc;
    }
    p.dk[s] = static_cast<bf16*>(a->dk[s]);
    p.dv[s] = static_cast<bf16*>(a->dv[s]);
    p.ld_dk[s] = a->ld_dk[s];
    p.ld_dv[s] = a->ld_dv[s];
    dp.o[s] = static_cast<const bf16*>(a->o[s]);
    dp.d_o[s] = static_cast<const bf16*>(a->d_o[s]);
    dp.ld_o[s] = a->ld_o[s];
    dp.ld_do[s] = a->ld_do[s];
    dp.rows[s] = (long long)a->B * rows[s];
    dp.rows_per_sample[s] = rows[s];
  }
  // delta = rowsum(dO * O) of both streams, and the fp32 dQ accumulator cleared, in one launch
  dp.delta = a->delta; dp.dq_acc = a->dq_acc;
  dp.H = a->H; dp.T = T; dp.N = a->N;
  if (dp.rows_per_sample[1] == 0) dp.rows_per_sample[1] = 1;
  {
    const long long work = (dp.rows[0] + dp.rows[1]) * a->H * 8;
    MMDIT_CARVEOUT(attn_delta_kernel);
    launch_k(attn_delta_kernel, dim3((unsigned)((work + 255) / 256)), dim3(256), 0, stream, dp);
  }
  int rc = check_launch("attn_delta_kernel");
  if (rc) return rc;
  p.lse = a->lse; p.delta = a->delta; p.dq_acc = a->dq_acc;
  p.B = a->B; p.H = a->H; p.N = a->N; p.M = a->M;
  p.scale = a->scale;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  static const cudaError_t attr_rc =   // thread-safe one-time initialisation (C++11 magic static)
      cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM);
  if (attr_rc != cudaSuccess) {
    set_last_error("attn_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(attr_rc));
    return (int)attr_rc;
  }
  const int nt = (a->N + ATT_TILE - 1) / ATT_TILE + (a->M + ATT_TILE - 1) / ATT_TILE;
  static const int use_v2 = [] {   // MMDIT_ATTN_BWD_V2=0: the first-generation (CTA per item) kernel
    const char* ev = getenv("MMDIT_ATTN_BWD_V2");
    return ev ? atoi(ev) : 1;
  }();
  if (use_v2) {
    static const cudaError_t attr2_rc =
        cudaFuncSetAttribute(attn_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD2_SMEM);
    if (attr2_rc != cudaSuccess) {
      set_last_error("attn_bwd2: cudaFuncSetAttribute: %s", cudaGetErrorString(attr2_rc));
      return (int)attr2_rc;
    }
    const int items = nt * a->H * a->B;
    const int grid2 = items < num_sms() ? items : num_sms();
    launch_k(attn_bwd2_kernel, dim3(grid2), dim3(BWD_THREADS), BWD2_SMEM, stream, p, items);
  } else {
    dim3 grid(nt, a->H, a->B);
    launch_k(attn_bwd_kernel, grid, dim3(BWD_THREADS), BWD_SMEM, stream, p);
  }
  rc = check_launch("attn_bwd_kernel");
  if (rc) return rc;
  int converts = 0;
  for (int s = 0; s < 2; ++s) {
    // dq[s] == NULL: the caller consumes the fp32 accumulator dq_acc itself (mmdit_qknorm_rope_bwd_acc)
    if (rows[s] == 0 || !a->dq[s]) continue;
    ++converts;
    const long long work = (long long)a->B * rows[s] * (dmodel / 8);
    long long blocks = (work + 255) / 256;
    if (blocks > num_sms() * 16LL) blocks = num_sms() * 16LL;
    MMDIT_CARVEOUT(attn_dq_convert_kernel);
    launch_k(attn_dq_convert_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, 
        a->dq_acc, static_cast<bf16*>(a->dq[s]), a->B, T, s == 0 ? 0 : a->N, rows[s], dmodel,
        a->ld_dq[s]);
  }
  return converts ? check_launch("attn_dq_convert_kernel", converts) : 0;
}
