// Joint text+image softmax attention, forward, second generation (the kernel the model runs):
// persistent CTAs, P kept in tensor memory, 8 softmax warps, TMA-store epilogue.
// Replaces flash_attn_func (Attention.py:293) and the concat / transpose / split copies around it
// (Attention.py:259-263, 411-417) exactly like attn_fwd.cu; that first-generation kernel stays as
// the online-softmax path for callers without a logit bound.
//
// This kernel runs the single-pass softmax: with QK-RMSNorm (Attention.py:61-64,130-134) the scaled
// logits are bounded by `bound` = 8 max|w_q| max|w_k| (mmdit_qk_logit_bound), so
//     P = exp(s - bound),   O = (sum_j P_j V_j) / l,   l = sum P,   LSE = bound + log l
// needs neither a running maximum nor a rescale of O.
//
// Work item = (sample, head, 128-query tile); 2 CTAs per SM, each loops over items
// (blockIdx.x, + gridDim.x, ...): TMEM is allocated and the mbarriers are initialised ONCE, and the
// Q / K / V tiles of the next item are already in flight while the current one finishes.
//   warp 0      TMA producer: Q (2-deep ring over items), K and V (separate 2-deep rings over key tiles)
//   warp 1      MMA issuer:   S = Q K^T (TMEM cols 0..127), O += P V with P read from TMEM
//                             (tcgen05.mma A operand in tensor memory, cols 128..191), O in cols 192..255
//   warps 2..9  softmax: thread = (query row, 64-key half).  tcgen05.ld S -> exp2 -> bf16 pairs ->
//               tcgen05.st P.  No shared-memory round trip for P (the first generation spent 32 KB of
//               st.shared + a proxy fence per tile, and the P.V MMA read its A operand from smem at
//               192 B/clk -- above the 128 B/clk an SM delivers).
// S of the next key tile (or of the next ITEM's first tile) is issued as soon as every thread holds its
// S values in registers, so the tensor pipe works under the exponentials; the epilogue (O / l -> bf16
// -> swizzled smem tile -> one TMA store, rows past the end of a stream clipped by the TMA unit)
// overlaps the next item's first QK^T.
// Tiles never straddle the image/text boundary: each stream is tiled separately, partial tiles are
// zero-filled by TMA and masked, so any N, M work.
#include <type_traits>

#include "common.cuh"
#include "mmdit_b200.h"

namespace mmdit {

constexpr int F2_TILE = 128;
constexpr int F2_HD = 64;
constexpr int F2_THREADS = 320;
constexpr int F2_TILE_BYTES = F2_TILE * F2_HD * 2;          // 16 KiB
constexpr int F2_SMEM = 6 * F2_TILE_BYTES + 1024 + 256;     // Q[2], K[2], V[2], row-sum exchange, barriers

int make_attn_tmap(CUtensorMap* map, const void* base, long long ld, int H, int rows, int B);

// Optional in-kernel timeline (tuning): one chosen CTA records (event id, clock64) pairs.
// [0,256): MMA thread, [256,768): softmax warp 2 lane 0, [768,1280): softmax warp 7 lane 0
__device__ long long* g_attn2_timeline = nullptr;
__device__ int g_attn2_timeline_block = -1;
#ifdef MMDIT_ATTN_TIMELINE   // make EXTRA=-DMMDIT_ATTN_TIMELINE: the stamps cost registers in the hot loops
#define TL2(id)                                                                          \
  do {                                                                                   \
    if (tl && tls < tl_cap) { tl[2 * tls] = (id); tl[2 * tls + 1] = clock64(); ++tls; }  \
  } while (0)
#else
#define TL2(id) do { } while (0)
#endif

struct AttnFwd2Params {
  CUtensorMap tmQ[2], tmK[2], tmV[2], tmO[2];  // [0] image stream, [1] text stream
  float* lse;                                  // [B, H, N+M]
  int B, H, N, M;
  int items;
  float scale, scale_log2;
  const float* logit_bound;                    // device scalar (never null here)
};

__device__ __forceinline__ float f2_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2,
                                             int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(m),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// 32 lanes x 16 consecutive fp32 columns / 8 packed columns (smaller register blocks than the x32 forms
// of common.cuh: two loads in flight + one packed block fit the 96-register budget of this kernel)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// key tile j of the joint sequence -> (stream, first row, valid rows)
struct KvTile {
  int s, row0, nv;
};
__device__ __forceinline__ KvTile kv_tile(int j, int ntx, int N, int M) {
  KvTile t;
  t.s = j < ntx ? 0 : 1;
  t.row0 = (t.s == 0 ? j : j - ntx) * F2_TILE;
  t.nv = min(F2_TILE, (t.s == 0 ? N : M) - t.row0);
  return t;
}

__global__ void __launch_bounds__(F2_THREADS, 2)
attn_fwd2_kernel(const __grid_constant__ AttnFwd2Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                        // [2]; the finished item's slot doubles as the O staging tile
  uint8_t* sK = smem + 2 * F2_TILE_BYTES;    // [2]
  uint8_t* sV = smem + 4 * F2_TILE_BYTES;    // [2]
  float* sL = reinterpret_cast<float*>(smem + 6 * F2_TILE_BYTES);   // [2][128] partial row sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * F2_TILE_BYTES + 1024);
  uint64_t* q_full = bars + 0;    // [2]
  uint64_t* q_empty = bars + 2;   // [2]
  uint64_t* k_full = bars + 4;    // [2]
  uint64_t* k_empty = bars + 6;   // [2]
  uint64_t* v_full = bars + 8;    // [2]
  uint64_t* v_empty = bars + 10;  // [2]
  uint64_t* s_full = bars + 12;
  uint64_t* s_free = bars + 13;   // 256: every softmax thread holds its S values in registers
  uint64_t* p_full = bars + 14;   // 256: P of this tile is in tensor memory
  uint64_t* pv_done = bars + 15;  // P V of this tile has retired (P region free; at the last tile: O final)
  uint64_t* o_free = bars + 16;   // 256: O of the finished item is in registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntx = (p.N + F2_TILE - 1) / F2_TILE;
  const int ntc = (p.M + F2_TILE - 1) / F2_TILE;
  const int nt = ntx + ntc;                  // query tiles == key tiles per (sample, head)

#ifdef MMDIT_ATTN_TIMELINE
  long long* tl = nullptr;
  int tls = 0, tl_cap = 0;
  if (g_attn2_timeline && (int)blockIdx.x == g_attn2_timeline_block && lane == 0) {
    if (warp == 1) { tl = g_attn2_timeline; tl_cap = 128; }
    if (warp == 2) { tl = g_attn2_timeline + 256; tl_cap = 256; }
    if (warp == 7) { tl = g_attn2_timeline + 768; tl_cap = 256; }
  }
#endif
  TL2(1);
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1);
        mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
        mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
      }
      mbar_init(s_full, 1);
      mbar_init(s_free, 256);
      mbar_init(p_full, 256);
      mbar_init(pv_done, 1);
      mbar_init(o_free, 256);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_S = tmem_base, tm_P = tmem_base + 128, tm_O = tmem_base + 192;
  pdl_wait();   // barrier init and the TMEM allocation overlapped the previous kernel's tail
  const float bound = *p.logit_bound;
  const bool usable = bound >= 0.f && bound <= 24.f;   // else: the online-softmax kernel handles this launch

  auto decode = [&](int item, int& b, int& h, int& qt) {
    qt = item % nt;
    const int bh = item / nt;
    h = bh % p.H;
    b = bh / p.H;
  };

  if (!usable) {
    // nothing to do (the TMEM allocation is still released below)
  } else if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int n = 0;   // key tiles issued so far (ring position), over all items
      int k = 0;   // items issued so far
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++k) {
        int b, h, qt;
        decode(item, b, h, qt);
        const int qs = qt < ntx ? 0 : 1;
        const int q_row0 = (qs == 0 ? qt : qt - ntx) * F2_TILE;
        const int slot = k & 1;
        mbar_wait(&q_empty[slot], ((k >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[slot], F2_TILE_BYTES);
        tma_load_4d(sQ + slot * F2_TILE_BYTES, &p.tmQ[qs], &q_full[slot], 0, h, q_row0, b);
        for (int j = 0; j < nt; ++j, ++n) {
          const KvTile t = kv_tile(j, ntx, p.N, p.M);
          const int st = n & 1;
          const uint32_t par = ((n >> 1) & 1) ^ 1;
          mbar_wait(&k_empty[st], par);
          mbar_expect_tx(&k_full[st], F2_TILE_BYTES);
          tma_load_4d(sK + st * F2_TILE_BYTES, &p.tmK[t.s], &k_full[st], 0, h, t.row0, b);
          mbar_wait(&v_empty[st], par);
          mbar_expect_tx(&v_full[st], F2_TILE_BYTES);
          tma_load_4d(sV + st * F2_TILE_BYTES, &p.tmV[t.s], &v_full[st], 0, h, t.row0, b);
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer
    // The whole warp walks the loop (every lane waits on the barriers); one elected lane issues.
    {
      const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);
      const uint64_t q_desc0 = desc_kmajor(smem_u32(sQ), 0);
      const uint64_t k_desc0 = desc_kmajor(smem_u32(sK), 0);
      const uint64_t v_desc0 = desc_mnmajor(smem_u32(sV), 0, F2_TILE_BYTES);
      constexpr uint64_t kStepK = 32 >> 4, kStepMN = 2048 >> 4, kTile = F2_TILE_BYTES >> 4;
      const int my_items = p.items > (int)blockIdx.x ? (p.items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
      // S of global tile number n (item k, key tile j): Q slot k&1, K stage n&1
      auto issue_s = [&](int n, int k, int j) {
        const KvTile t = kv_tile(j, ntx, p.N, p.M);
        const int st = n & 1;
        const uint32_t idesc_s = make_idesc_bf16(128, (t.nv + 15) & ~15, 0, 0);
        if (j == 0) {
          mbar_wait(&q_full[k & 1], (k >> 1) & 1);
        }
        mbar_wait(&k_full[st], (n >> 1) & 1);
        tc_fence_after();
        const uint64_t qd = q_desc0 + (k & 1) * kTile, kd = k_desc0 + st * kTile;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < F2_HD / 16; ++kk)
            umma_bf16(tm_S, qd + kk * kStepK, kd + kk * kStepK, idesc_s, kk > 0);
          umma_commit(&k_empty[st]);
          umma_commit(s_full);
        }
        __syncwarp();
      };
      int n = 0;
      if (my_items > 0) issue_s(0, 0, 0);
      for (int k = 0; k < my_items; ++k) {
        for (int j = 0; j < nt; ++j, ++n) {
          const KvTile t = kv_tile(j, ntx, p.N, p.M);
          const int st = n & 1;
          const int n_mma = (t.nv + 15) & ~15;
          const bool last = (j + 1 == nt);
          if (!last || k + 1 < my_items) {
            mbar_wait(s_free, n & 1);      // S_n sits in the softmax threads' registers: TMEM S is free
            tc_fence_after();
            TL2(1000 + n);
            if (!last) issue_s(n + 1, k, j + 1);
            else issue_s(n + 1, k + 1, 0);  // the next item's first tile: runs under this item's epilogue
            TL2(2000 + n);
          }
          mbar_wait(p_full, n & 1);
          TL2(3000 + n);
          if (j == 0 && k > 0) mbar_wait(o_free, (k - 1) & 1);   // previous item's O has been read out
          mbar_wait(&v_full[st], (n >> 1) & 1);
          tc_fence_after();
          const uint64_t vd = v_desc0 + st * kTile;
          if (elect_one()) {
            if (n_mma == F2_TILE) {
#pragma unroll
              for (int kk = 0; kk < F2_TILE / 16; ++kk)
                umma_bf16_ts(tm_O, tm_P + kk * 8, vd + kk * kStepMN, idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
            } else {
              for (int kk = 0; kk < n_mma / 16; ++kk)
                umma_bf16_ts(tm_O, tm_P + kk * 8, vd + kk * kStepMN, idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
            }
            umma_commit(&v_empty[st]);
            umma_commit(pv_done);
          }
          __syncwarp();
          TL2(4000 + n);
        }
      }
    }
  } else {
    // ----------------------------------------------------------------------- softmax
    const int sw = warp - 2;                 // 0..7
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int hf = sw >> 2;                  // which 64-key half of a tile
    const int r = quarter * 32 + lane;       // query row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const float sl2 = p.scale_log2;
    const float mb = bound * 1.4426950408889634f;
    int n = 0, k = 0;
    int pending = -1;   // warp 2 / lane 0: Q slot whose O store has been issued but not yet drained
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++k) {
      int b, h, qt;
      decode(item, b, h, qt);
      const int qs = qt < ntx ? 0 : 1;
      const int q_row0 = (qs == 0 ? qt : qt - ntx) * F2_TILE;
      const int q_rows = qs == 0 ? p.N : p.M;
      const int q_valid = min(F2_TILE, q_rows - q_row0);
      // a partial query tile leaves whole warps without a valid row: they keep the barrier protocol
      // going but skip the math (their P / O rows are never stored)
      const bool warp_active = quarter * 32 < q_valid;
      float l = 0.f;
      for (int j = 0; j < nt; ++j, ++n) {
        const KvTile t = kv_tile(j, ntx, p.N, p.M);
        const int n_mma = (t.nv + 15) & ~15;
        const int c0 = hf * 64;                        // first key column of this thread
        const bool have = warp_active && c0 < n_mma;   // warp-uniform
        mbar_wait(s_full, n & 1);
        tc_fence_after();
        TL2(1000 + n);
        if (have) {
          // 16 key columns at a time, two TMEM loads in flight: exp2 -> 8 packed bf16 pairs -> tcgen05.st
          const int nch = min(4, (n_mma - c0 + 15) >> 4);     // 16-column chunks of this thread (warp-uniform)
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
          bool p_free = (n == 0);
          auto chunk = [&](const uint32_t (&sv)[16], int ci) {
            const int col = c0 + ci * 16;
            uint32_t pk[8];
            if (col + 16 <= t.nv) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const float e0 = f2_ex2(fmaf(__uint_as_float(sv[i]), sl2, -mb));
                const float e1 = f2_ex2(fmaf(__uint_as_float(sv[i + 1]), sl2, -mb));
                const float e2 = f2_ex2(fmaf(__uint_as_float(sv[i + 2]), sl2, -mb));
                const float e3 = f2_ex2(fmaf(__uint_as_float(sv[i + 3]), sl2, -mb));
                a0 += e0; a1 += e1; a2 += e2; a3 += e3;
                pk[i >> 1] = pack_bf16x2(e0, e1);
                pk[(i >> 1) + 1] = pack_bf16x2(e2, e3);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; i += 2) {
                float e0 = f2_ex2(fmaf(__uint_as_float(sv[i]), sl2, -mb));
                float e1 = f2_ex2(fmaf(__uint_as_float(sv[i + 1]), sl2, -mb));
                if (col + i >= t.nv) e0 = 0.f;
                if (col + i + 1 >= t.nv) e1 = 0.f;
                a0 += e0; a1 += e1;
                pk[i >> 1] = pack_bf16x2(e0, e1);
              }
            }
            if (!p_free) {   // the P region is free once the previous tile's P V has retired
              mbar_wait(pv_done, (n - 1) & 1);
              tc_fence_after();
              p_free = true;
            }
            tmem_st8(tm_P + lane_off + hf * 32 + ci * 8, pk);
          };
          uint32_t sa[16], sb[16];
          tmem_ld16(tm_S + lane_off + c0, sa);
          if (nch > 1) tmem_ld16(tm_S + lane_off + c0 + 16, sb);
          tmem_ld_wait();
          TL2(2000 + n);
          if (nch <= 2) {
            tc_fence_before();
            mbar_arrive(s_free);       // S has left TMEM: the next QK^T may overwrite it
          }
          chunk(sa, 0);
          TL2(3000 + n);
          if (nch > 2) tmem_ld16(tm_S + lane_off + c0 + 32, sa);
          if (nch > 1) chunk(sb, 1);
          if (nch > 2) {
            if (nch > 3) tmem_ld16(tm_S + lane_off + c0 + 48, sb);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(s_free);
            TL2(4000 + n);
            chunk(sa, 2);
            if (nch > 3) chunk(sb, 3);
          }
          l += (a0 + a1) + (a2 + a3);
          TL2(5000 + n);
          tmem_st_wait();
          TL2(6000 + n);
        } else {
          tc_fence_before();
          mbar_arrive(s_free);
          // (also keeps a thread without work from arriving on p_full of a phase that is still open)
          if (n > 0) mbar_wait(pv_done, (n - 1) & 1);
        }
        tc_fence_before();
        mbar_arrive(p_full);
        if (pending >= 0) {
          // the previous item's O store has had a whole key tile to read its staging tile: hand the
          // Q slot back to the producer now (waiting right after the store would stall this warp)
          tma_wait_group_read0();
          mbar_arrive(&q_empty[pending]);
          pending = -1;
        }
      }
      // ---------------------------------------------------------------- epilogue
      sL[hf * 128 + r] = l;
      TL2(7000 + k);
      mbar_wait(pv_done, (n - 1) & 1);       // the last P V has retired: O is final
      tc_fence_after();
      TL2(7100 + k);
      named_bar_sync(1, 256);                // partial row sums visible
      TL2(7200 + k);
      const float lt = sL[r] + sL[128 + r];
      uint32_t o[32];
      if (warp_active) {
        tmem_ld32(tm_O + lane_off + hf * 32, o);
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(o_free);                   // the next item's first P V may overwrite O
      uint8_t* stage = sQ + (k & 1) * F2_TILE_BYTES;   // this item's Q tile: all its MMAs have retired
      if (warp_active) {
        const float inv_l = 1.f / lt;
        uint8_t* prow = stage + r * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[8 * g]) * inv_l, __uint_as_float(o[8 * g + 1]) * inv_l);
          u.y = pack_bf16x2(__uint_as_float(o[8 * g + 2]) * inv_l, __uint_as_float(o[8 * g + 3]) * inv_l);
          u.z = pack_bf16x2(__uint_as_float(o[8 * g + 4]) * inv_l, __uint_as_float(o[8 * g + 5]) * inv_l);
          u.w = pack_bf16x2(__uint_as_float(o[8 * g + 6]) * inv_l, __uint_as_float(o[8 * g + 7]) * inv_l);
          *reinterpret_cast<uint4*>(prow + (((hf * 4 + g) ^ (r & 7)) << 4)) = u;
        }
        if (hf == 0 && r < q_valid) {
          const int tq = (qs == 0 ? 0 : p.N) + q_row0 + r;
          p.lse[((long long)b * p.H + h) * (p.N + p.M) + tq] = bound + __logf(lt);
        }
      }
      TL2(7300 + k);
      fence_proxy_async_smem();
      named_bar_sync(1, 256);                // the staging tile is complete (and sL may be rewritten)
      TL2(7400 + k);
      if (warp == 2 && lane == 0) {
        tma_store_4d(&p.tmO[qs], stage, 0, h, q_row0, b);   // rows >= q_rows are clipped by the TMA unit
        tma_commit_group();
        pending = k & 1;
      }
    }
    if (warp == 2 && lane == 0) tma_wait_group0();   // staging tiles must outlive their stores
  }
  if (warp >= 2) pdl_trigger();   // late trigger, see gemm_tcgen05.cu
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace mmdit

using namespace mmdit;

extern "C" int mmdit_debug_attn2_timeline(long long* buf, int block) {
  cudaError_t e = cudaMemcpyToSymbol(g_attn2_timeline, &buf, sizeof(buf));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_attn2_timeline_block, &block, sizeof(block));
  return (int)e;
}

// Launches the second-generation forward for `a` (requires a->logit_bound); the kernel returns at
// once when the bound turns out to be unusable (> 24), in which case the caller's online kernel runs.
int launch_attn_fwd2(const mmdit_attn_args* a, cudaStream_t stream) {
  AttnFwd2Params p;
  memset(&p, 0, sizeof(p));
  const int rows[2] = {a->N, a->M};
  for (int s = 0; s < 2; ++s) {
    if (rows[s] == 0) continue;
    int rc = make_attn_tmap(&p.tmQ[s], a->q[s], a->ld_q[s], a->H, rows[s], a->B);
    if (rc) return rc;
    rc = make_attn_tmap(&p.tmK[s], a->k[s], a->ld_k[s], a->H, rows[s], a->B);
    if (rc) return rc;
    rc = make_attn_tmap(&p.tmV[s], a->v[s], a->ld_v[s], a->H, rows[s], a->B);
    if (rc) return rc;
    rc = make_attn_tmap(&p.tmO[s], a->o[s], a->ld_o[s], a->H, rows[s], a->B);
    if (rc) return rc;
  }
  p.lse = a->lse;
  p.B = a->B; p.H = a->H; p.N = a->N; p.M = a->M;
  p.scale = a->scale;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.logit_bound = a->logit_bound;
  const int nt = (a->N + F2_TILE - 1) / F2_TILE + (a->M + F2_TILE - 1) / F2_TILE;
  p.items = nt * a->H * a->B;
  static const cudaError_t attr_rc =
      cudaFuncSetAttribute(attn_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM);
  if (attr_rc != cudaSuccess) {
    set_last_error("attn_fwd2: cudaFuncSetAttribute: %s", cudaGetErrorString(attr_rc));
    return (int)attr_rc;
  }
  const int grid = p.items < 2 * num_sms() ? p.items : 2 * num_sms();
  launch_k(attn_fwd2_kernel, dim3(grid), dim3(F2_THREADS), F2_SMEM, stream, p);
  return check_launch("attn_fwd2_kernel");
}
