// Joint text+image softmax attention, forward, on tcgen05 tensor cores.
// Replaces flash_attn_func (Attention.py:293) plus the concat / transpose /
// split copies around it (Attention.py:259-263, 411-417): Q, K, V are read in
// place from the two streams (image rows first, then text rows -- the
// reference's concat order) through 4-D TMA descriptors, and the output is
// written straight into the per-stream [rows, dim] buffers the out-projection
// GEMMs consume.
//
// One CTA = one (sample, head, 128-query tile); two CTAs per SM so that the
// softmax of one overlaps the MMAs of the other.  head_dim = 64.
//   warp 0      TMA producer (Q once, then K/V tiles through a 2-deep ring)
//   warp 1      MMA issuer: S = Q K^T (TMEM, 128 cols), O += P V (TMEM, 64 cols)
//   warps 2..5  softmax: one thread per query row (TMEM lane), online softmax in
//               fp32 with exp2, P written as bf16 into 128B-swizzled smem (the A
//               operand of the second MMA), O rescaled in TMEM.
// Pipelining inside a CTA: a softmax thread turns its S row into packed bf16 REGISTERS; the
// moment its last S column has left TMEM it signals `s_free`, and the MMA warp issues S of the
// NEXT key tile while this tile's exponentials are still being computed; P is stored to smem only
// after the previous P.V has retired, and that P.V overlaps the next tile's softmax math.
// Tiles never straddle the image/text boundary: each stream is tiled
// separately and partial tiles are masked, so any N, M work.
#include <stdlib.h>
#include <type_traits>

#include "common.cuh"
#include "mmdit_b200.h"

namespace mmdit {

constexpr int ATT_TILE = 128;
constexpr int ATT_HD = 64;
constexpr int ATT_THREADS = 192;
constexpr int ATT_TILE_BYTES = ATT_TILE * ATT_HD * 2;  // 16 KiB
constexpr int ATT_SMEM = 7 * ATT_TILE_BYTES + 256;     // Q, K[2], V[2], P(2 tiles) + barriers

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Optional in-kernel timeline (debugging / tuning): one chosen CTA records (event id, clock64).
__device__ long long* g_attn_timeline = nullptr;
__device__ int g_attn_timeline_block = -1;
#define TL(slot, id)                                                        \
  do {                                                                      \
    if (tl) { const int s_ = (slot); tl[2 * s_] = (id); tl[2 * s_ + 1] = clock64(); } \
  } while (0)

struct AttnFwdParams {
  CUtensorMap tmQ[2], tmK[2], tmV[2];  // [0] image stream, [1] text stream
  bf16* o[2];
  long long ldo[2];
  float* lse;  // [B, H, N+M]
  int B, H, N, M;
  float scale, scale_log2;
  const float* logit_bound;  // device scalar: upper bound of |scale * q.k| (QK-RMSNorm makes it small), or null
  int skip_if_bounded;       // the second-generation kernel (attn_fwd2.cu) serves launches whose bound is usable
  int items;                 // > 0: 1-D grid whose CTAs loop over (sample, head, query tile) items -- the launch
                             // behind attn_fwd2, which normally has nothing to do: a few hundred CTAs that exit at
                             // once instead of B*H*tiles of them (ncu: 16 us of CTA launches for an empty grid)
};

__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ AttnFwdParams p) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + ATT_TILE_BYTES;
  uint8_t* sV = smem + 3 * ATT_TILE_BYTES;
  uint8_t* sP = smem + 5 * ATT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * ATT_TILE_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* pv_done = bars + 7;
  uint64_t* s_free = bars + 8;    // every softmax thread holds its S row in registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  if (p.skip_if_bounded) {
    const float bd = *p.logit_bound;
    if (bd >= 0.f && bd <= 24.f) return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntx = (p.N + ATT_TILE - 1) / ATT_TILE;
  const int ntc = (p.M + ATT_TILE - 1) / ATT_TILE;
  const int nkv = ntx + ntc;
  // one item per CTA (3-D grid), or a loop over items (1-D grid, p.items > 0); every item sets up and
  // tears down its barriers and its TMEM allocation exactly like a CTA of its own
  for (int item = p.items > 0 ? (int)blockIdx.x : 0; item < (p.items > 0 ? p.items : 1); item += (int)gridDim.x) {
  const bool first_item = p.items > 0 ? item == (int)blockIdx.x : true;
  const bool last_item = p.items > 0 ? item + (int)gridDim.x >= p.items : true;
  const int qt = p.items > 0 ? item % nkv : (int)blockIdx.x;
  const int h = p.items > 0 ? (item / nkv) % p.H : (int)blockIdx.y;
  const int b = p.items > 0 ? item / (nkv * p.H) : (int)blockIdx.z;
  const int qs = qt < ntx ? 0 : 1;                       // stream of the query tile
  const int q_row0 = (qs == 0 ? qt : qt - ntx) * ATT_TILE;
  const int q_rows = qs == 0 ? p.N : p.M;
  const int q_valid = min(ATT_TILE, q_rows - q_row0);

  const int linear_block = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  long long* tl = nullptr;   // [0,64): MMA warp events, [64,192): softmax warp 2 events
  if (g_attn_timeline && linear_block == g_attn_timeline_block && lane == 0) {
    if (warp == 1) tl = g_attn_timeline;
    if (warp == 2) tl = g_attn_timeline + 128;
  }
  int tls = 0;
  TL(tls++, 1);
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int s = 0; s < 2; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
      mbar_init(s_full, 1);
      mbar_init(p_full, 128);
      mbar_init(pv_done, 1);
      mbar_init(s_free, 128);
      mbar_fence_init();
      // fire Q and the first two K/V tiles right away: they fly while TMEM is allocated
      mbar_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_4d(sQ, &p.tmQ[qs], q_full, 0, h, q_row0, b);
      for (int j = 0; j < 2 && j < nkv; ++j) {
        const int ks = j < ntx ? 0 : 1;
        const int row0 = (ks == 0 ? j : j - ntx) * ATT_TILE;
        mbar_expect_tx(&kv_full[j], 2 * ATT_TILE_BYTES);
        tma_load_4d(sK + j * ATT_TILE_BYTES, &p.tmK[ks], &kv_full[j], 0, h, row0, b);
        tma_load_4d(sV + j * ATT_TILE_BYTES, &p.tmV[ks], &kv_full[j], 0, h, row0, b);
      }
    }
    __syncwarp();
    if (first_item) {   // one TMEM allocation per CTA (the permit is relinquished: no second alloc possible)
      tmem_alloc(tmem_slot, 256);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      for (int j = 2; j < nkv; ++j) {   // tiles 0 and 1 were issued in the prologue
        const int st = j & 1;
        const int ks = j < ntx ? 0 : 1;
        const int row0 = (ks == 0 ? j : j - ntx) * ATT_TILE;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&kv_full[st], 2 * ATT_TILE_BYTES);
        tma_load_4d(sK + st * ATT_TILE_BYTES, &p.tmK[ks], &kv_full[st], 0, h, row0, b);
        tma_load_4d(sV + st * ATT_TILE_BYTES, &p.tmV[ks], &kv_full[st], 0, h, row0, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);
      // descriptors are built once; stepping K only adds a constant to the 14-bit address field
      const uint64_t q_desc = desc_kmajor(smem_u32(sQ), 0);
      const uint64_t p_desc = desc_kmajor(smem_u32(sP), 0);
      const uint64_t k_desc0 = desc_kmajor(smem_u32(sK), 0);
      const uint64_t v_desc0 = desc_mnmajor(smem_u32(sV), 0, ATT_TILE_BYTES);
      constexpr uint64_t kStepK = 32 >> 4, kStepMN = 2048 >> 4, kTile = ATT_TILE_BYTES >> 4;
      TL(tls++, 2);
      mbar_wait(q_full, 0);
      TL(tls++, 3);
      auto tile_cols = [&](int j) {   // key columns of tile j that exist, rounded up to the MMA N step
        const int ks = j < ntx ? 0 : 1;
        const int row0 = (ks == 0 ? j : j - ntx) * ATT_TILE;
        const int nv = min(ATT_TILE, (ks == 0 ? p.N : p.M) - row0);
        return (nv + 15) & ~15;
      };
      auto issue_s = [&](int j) {
        const int st = j & 1;
        const uint32_t idesc_s = make_idesc_bf16(128, tile_cols(j), 0, 0);
        const uint64_t k_desc = k_desc0 + st * kTile;
        mbar_wait(&kv_full[st], (j >> 1) & 1);
        tc_fence_after();
        TL(tls++, 10 + j);
#pragma unroll
        for (int k = 0; k < ATT_HD / 16; ++k)
          umma_bf16(tmem_S, q_desc + k * kStepK, k_desc + k * kStepK, idesc_s, k > 0);
        umma_commit(s_full);
        TL(tls++, 20 + j);
      };
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        const int st = j & 1;
        const int n_mma = tile_cols(j);
        const uint64_t v_desc = v_desc0 + st * kTile;
        if (j + 1 < nkv) {
          mbar_wait(s_free, j & 1);   // S_j sits in the softmax threads' registers: TMEM S is free
          tc_fence_after();
          issue_s(j + 1);
        }
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        TL(tls++, 30 + j);
        if (n_mma == ATT_TILE) {
#pragma unroll
          for (int k = 0; k < ATT_TILE / 16; ++k)
            umma_bf16(tmem_O, p_desc + (k >> 2) * kTile + (k & 3) * kStepK, v_desc + k * kStepMN, idesc_o,
                      (j > 0 || k > 0) ? 1u : 0u);
        } else {
          for (int k = 0; k < n_mma / 16; ++k)
            umma_bf16(tmem_O, p_desc + (k >> 2) * kTile + (k & 3) * kStepK, v_desc + k * kStepMN, idesc_o,
                      (j > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&kv_empty[st]);
        umma_commit(pv_done);
        TL(tls++, 40 + j);
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    float m = -INFINITY, l = 0.f;
    const float sl2 = p.scale_log2;
    // With QK-RMSNorm the scaled logits are bounded by 8*max|w_q|*max|w_k|; when the caller
    // supplies that bound (and it is small enough for exp2 not to underflow) the softmax needs
    // neither a running maximum nor O rescaling: P = exp(s - bound) in one pass over S.
    float bound = p.logit_bound ? *p.logit_bound : INFINITY;
    const bool use_bound = bound >= 0.f && bound <= 24.f;
    const float bound_l2 = bound * 1.4426950408889634f;
    if (use_bound) m = bound / p.scale;   // so that lse = m*scale + log(l) below stays valid
    // A partial query tile (e.g. the 26 text rows left after a full 128-row tile) leaves whole warps
    // without a valid row: they keep the barrier protocol going but skip the softmax math (their P
    // rows stay whatever the buffer held -- every output row depends on its own P row only, and
    // those rows are never stored).
    const bool warp_active = quarter * 32 < q_valid;
    for (int j = 0; j < nkv; ++j) {
      const int ks = j < ntx ? 0 : 1;
      const int row0 = (ks == 0 ? j : j - ntx) * ATT_TILE;
      const int nv = min(ATT_TILE, (ks == 0 ? p.N : p.M) - row0);  // valid keys in this tile
      const int n_mma = (nv + 15) & ~15;
      const int nchunk = (n_mma + 31) / 32;
      TL(tls++, 50 + j);
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      TL(tls++, 60 + j);
      if (!warp_active) {
        tc_fence_before();
        mbar_arrive(s_free);
        mbar_arrive(p_full);
        // stay in step with the working warps: S of the next tile is issued as soon as s_free
        // completes, so without this wait an idle warp could see s_full(j+1) and arrive on p_full
        // a second time while phase j is still open (the phase would complete without P_j)
        mbar_wait(p_full, j & 1);
        continue;
      }
      float alpha = 1.f, mb, m_new = m;
      if (!use_bound) {
        // pass 1: row maximum (4 independent chains; TMEM loads issued in pairs)
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        for (int c = 0; c < nchunk; c += 2) {
          uint32_t s0[32], s1[32];
          tmem_ld32(tmem_S + lane_off + c * 32, s0);
          if (c + 1 < nchunk) tmem_ld32(tmem_S + lane_off + (c + 1) * 32, s1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < nv) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(s0[i]));
          if (c + 1 < nchunk) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if ((c + 1) * 32 + i < nv) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(s1[i]));
          }
        }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        m_new = fmaxf(m, mx);
        alpha = ex2_approx((m - m_new) * sl2);
        mb = m_new * sl2;
      } else {
        mb = bound_l2;   // fixed reference point: no running maximum, no rescale of O
      }
      TL(tls++, 70 + j);
      if (j > 0 && !use_bound) {
        mbar_wait(pv_done, (j - 1) & 1);   // previous P V retired: O is final so far
        tc_fence_after();
        TL(tls++, 80 + j);
        // rescale O only if some row of this warp actually raised its maximum
        if (!__all_sync(0xffffffffu, m_new == m)) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld32(tmem_O + lane_off + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tmem_O + lane_off + c * 32, o);
          }
          tmem_st_wait();
        }
      }
      TL(tls++, 90 + j);
      // P row of this thread as 64 packed bf16 pairs, kept in registers until the smem tile is free
      uint32_t pk[64];
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
      auto exp_row = [&](auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (FULL || c < nchunk) {
            uint32_t sv[32];
            tmem_ld32(tmem_S + lane_off + c * 32, sv);
            tmem_ld_wait();
            if (c == (FULL ? 3 : nchunk - 1)) {   // S has left TMEM: the next QK^T may overwrite it
              tc_fence_before();
              mbar_arrive(s_free);
            }
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float e0 = ex2_approx(fmaf(__uint_as_float(sv[i]), sl2, -mb));
              float e1 = ex2_approx(fmaf(__uint_as_float(sv[i + 1]), sl2, -mb));
              if (!FULL) {
                if (c * 32 + i >= nv) e0 = 0.f;
                if (c * 32 + i + 1 >= nv) e1 = 0.f;
              }
              rs4[(i >> 1) & 3] += e0 + e1;
              pk[c * 16 + (i >> 1)] = pack_bf16x2(e0, e1);
            }
          }
        }
      };
      if (nv == ATT_TILE) exp_row(std::true_type{});
      else exp_row(std::false_type{});
      if (j > 0 && use_bound) {
        mbar_wait(pv_done, (j - 1) & 1);   // previous P V retired: the P tile in smem is free
        tc_fence_after();
        TL(tls++, 80 + j);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < nchunk) {
          uint8_t* prow = sP + (c >> 1) * ATT_TILE_BYTES + r * 128;
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(prow + ((((c & 1) * 4 + g) ^ (r & 7)) << 4)) =
                make_uint4(pk[c * 16 + g * 4], pk[c * 16 + g * 4 + 1], pk[c * 16 + g * 4 + 2],
                           pk[c * 16 + g * 4 + 3]);
        }
      }
      const float rs = (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
      l = l * alpha + rs;
      m = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      TL(tls++, 100 + j);
    }
    mbar_wait(pv_done, (nkv - 1) & 1);
    tc_fence_after();
    TL(tls++, 110);
    const float inv_l = 1.f / l;
    const long long grow = (long long)b * q_rows + q_row0 + r;
    bf16* orow = p.o[qs] + grow * p.ldo[qs] + h * ATT_HD;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld32(tmem_O + lane_off + c * 32, o);
      tmem_ld_wait();
      if (r < q_valid) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(o[g * 8 + i]) * inv_l;
          store8(orow + c * 32 + g * 8, v);
        }
      }
    }
    if (r < q_valid) {
      const int t = (qs == 0 ? 0 : p.N) + q_row0 + r;
      p.lse[((long long)b * p.H + h) * (p.N + p.M) + t] = m * p.scale + logf(l);
    }
  }
  TL(tls++, 120);
  tc_fence_before();
  __syncthreads();
  if (warp == 0 && last_item) tmem_dealloc(tmem_base, 256);
  if (p.items > 0) {   // the next item re-initialises the barriers: invalidate them first
    if (threadIdx.x == 0)
      for (int i = 0; i < 9; ++i) mbar_inval(bars + i);
    __syncthreads();
  }
  }   // item loop
}

// 4-D view of one operand of one stream: (64 | H | rows | B), head h at column h*64.
int make_attn_tmap(CUtensorMap* map, const void* base, long long ld, int H, int rows, int B) {
  uint64_t dims[4] = {64, (uint64_t)H, (uint64_t)rows, (uint64_t)B};
  uint64_t strides[3] = {64 * 2, (uint64_t)ld * 2, (uint64_t)ld * 2 * (uint64_t)rows};
  uint32_t box[4] = {64, 1, ATT_TILE, 1};
  return encode_tmap(map, base, 4, dims, strides, box, 2, true);
}

}  // namespace mmdit

using namespace mmdit;

int launch_attn_fwd2(const mmdit_attn_args* a, cudaStream_t stream);   // attn_fwd2.cu

extern "C" int mmdit_debug_attn_timeline(long long* buf, int block) {
  cudaError_t e = cudaMemcpyToSymbol(g_attn_timeline, &buf, sizeof(buf));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_attn_timeline_block, &block, sizeof(block));
  return (int)e;
}

extern "C" int mmdit_attn_fwd(const mmdit_attn_args* a, void* stream) {
  MMDIT_REQUIRE(a, MMDIT_ERR_ARG, "attn_fwd: null args");
  MMDIT_REQUIRE(a->head_dim == 64, MMDIT_ERR_UNSUPPORTED, "attn_fwd: head_dim must be 64, got %d",
                a->head_dim);
  MMDIT_REQUIRE(a->B > 0 && a->H > 0 && a->N > 0 && a->M >= 0, MMDIT_ERR_ARG, "attn_fwd: bad shape");
  MMDIT_REQUIRE(a->q[0] && a->k[0] && a->v[0] && a->o[0] && a->lse, MMDIT_ERR_ARG,
                "attn_fwd: null image-stream pointer");
  MMDIT_REQUIRE(a->M == 0 || (a->q[1] && a->k[1] && a->v[1] && a->o[1]), MMDIT_ERR_ARG,
                "attn_fwd: null text-stream pointer");
  AttnFwdParams p;
  memset(&p, 0, sizeof(p));
  const int rows[2] = {a->N, a->M};
  for (int s = 0; s < 2; ++s) {
    if (rows[s] == 0) continue;
    MMDIT_REQUIRE(a->ld_q[s] % 8 == 0 && a->ld_k[s] % 8 == 0 && a->ld_v[s] % 8 == 0 &&
                      a->ld_o[s] % 8 == 0,
                  MMDIT_ERR_ALIGN, "attn_fwd: row strides must be multiples of 8 elements");
    int rc = make_attn_tmap(&p.tmQ[s], a->q[s], a->ld_q[s], a->H, rows[s], a->B);
    if (rc) return rc;
    rc = make_attn_tmap(&p.tmK[s], a->k[s], a->ld_k[s], a->H, rows[s], a->B);
    if (rc) return rc;
    rc = make_attn_tmap(&p.tmV[s], a->v[s], a->ld_v[s], a->H, rows[s], a->B);
    if (rc) return rc;
    p.o[s] = static_cast<bf16*>(a->o[s]);
    p.ldo[s] = a->ld_o[s];
  }
  p.lse = a->lse;
  p.B = a->B; p.H = a->H; p.N = a->N; p.M = a->M;
  p.scale = a->scale;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.logit_bound = a->logit_bound;
  static const cudaError_t attr_rc =   // thread-safe one-time initialisation (C++11 magic static)
      cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
  if (attr_rc != cudaSuccess) {
    set_last_error("attn_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(attr_rc));
    return (int)attr_rc;
  }
  // With a logit bound the second-generation kernel (attn_fwd2.cu) runs; should the bound turn out
  // to be unusable (> 24, only known on the device) it returns at once and the kernel below does the
  // work -- and vice versa.  MMDIT_ATTN_FWD_V2=0 keeps everything on the first-generation kernel.
  static const int use_v2 = [] {
    const char* e = getenv("MMDIT_ATTN_FWD_V2");
    return e ? atoi(e) : 1;
  }();
  if (a->logit_bound && use_v2) {
    const int rc2 = launch_attn_fwd2(a, static_cast<cudaStream_t>(stream));
    if (rc2) return rc2;
    p.skip_if_bounded = 1;
  }
  const int nt = (a->N + ATT_TILE - 1) / ATT_TILE + (a->M + ATT_TILE - 1) / ATT_TILE;
  dim3 grid(nt, a->H, a->B);
  if (p.skip_if_bounded) {
    p.items = nt * a->H * a->B;
    grid = dim3(p.items < 2 * num_sms() ? p.items : 2 * num_sms());
  }
  launch_k(attn_fwd_kernel, grid, dim3(ATT_THREADS), ATT_SMEM, static_cast<cudaStream_t>(stream), p);
  return check_launch("attn_fwd_kernel");
}
