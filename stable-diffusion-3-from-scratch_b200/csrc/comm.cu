// Data-parallel gradient all-reduce (mean) over NVLink peer memory -- the exchange step of the
// reference's DDP wrapper (model_trainer.py:224), as ONE kernel per gradient bucket that can be
// captured into the training step's CUDA graph and runs on a side stream next to the backward.
//
// Every rank keeps its fp32 gradient arena in a cudaMalloc allocation that its peers map through
// CUDA IPC (NVSwitch gives every GPU full-bandwidth loads/stores to every peer).  For a bucket
// [off, off+n) the kernel on rank r
//   1. tells every peer "my gradients of this bucket are final" (flag store, release.sys) and
//      waits for the same from all of them;
//   2. reduces ITS 1/W shard of the bucket: loads the shard from all W arenas (peer loads over
//      NVLink), sums in rank order 0..W-1 (the same order everywhere -> bit-identical replicas,
//      run-to-run deterministic), scales by 1/W and stores the result into all W arenas;
//   3. the last CTA to finish tells every peer "my shard is written" and waits for all peers.
// Flags carry a device-resident epoch (advanced by the kernel itself), so the same captured
// launch can be replayed forever.  All spins are bounded (trap instead of hanging the GPU).
#include <stdlib.h>

#include "common.cuh"
#include "mmdit_b200.h"

namespace mmdit {

constexpr int COMM_MAX_WORLD = 8;
constexpr int COMM_THREADS = 256;

struct CommParams {
  float* buf[COMM_MAX_WORLD];        // arena base of every rank (own = local pointer)
  uint32_t* flag[COMM_MAX_WORLD];    // signal pad of every rank: [2][COMM_MAX_WORLD] uint32
  uint32_t* state;                   // local: [0] epoch, [1] finished-CTA counter
  int world, rank;
  long long off, n;                  // bucket range in floats (off and n multiples of 4)
  float scale;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer16(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_peer16(float* p, const float4& v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
// wait until *p has reached `target` (epochs only grow; compare modulo 2^32)
__device__ __forceinline__ void spin_until(const uint32_t* p, uint32_t target) {
  const long long t0 = clock64();
  while (static_cast<int32_t>(ld_acquire_sys(p) - target) < 0) {
    __nanosleep(64);
    if (clock64() - t0 > 20000000000LL) __trap();  // ~10 s: a peer never arrived
  }
}

// U: batches of W peer loads issued before the first use (U x W x 16 bytes in flight per thread).
// Default U = 8 / W (8 loads in flight); MMDIT_COMM_UNROLL=2|4 multiplies it (W = 8: the
// 16-byte peer loads reached only 19 % of the NVLink rate on a 24 MB bucket -- more bytes in
// flight per thread is the first thing to sweep).
template <int W, int U>
__global__ void __launch_bounds__(COMM_THREADS) allreduce_mean_kernel(const CommParams p) {
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile uint32_t*>(p.state);
  __syncthreads();
  const uint32_t v1 = s_epoch + 1, v2 = s_epoch + 2;

  // 1. gradients final everywhere (this kernel is stream-ordered after the producers of the bucket)
  if (blockIdx.x == 0 && threadIdx.x < W) st_release_sys(p.flag[threadIdx.x] + p.rank, v1);
  if (threadIdx.x < W) spin_until(p.flag[p.rank] + threadIdx.x, v1);
  __syncthreads();

  // 2. reduce-scatter + all-gather of this rank's shard, 16 bytes per thread and peer;
  //    U x W = 8 peer loads in flight per thread whatever the world size
  const long long n4 = p.n >> 2;
  const long long per = (n4 + W - 1) / W;
  const long long lo = p.rank * per, hi = min(n4, lo + per);
  const long long base4 = p.off >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * COMM_THREADS;
  for (long long i = lo + static_cast<long long>(blockIdx.x) * COMM_THREADS + threadIdx.x; i < hi;
       i += U * stride) {
    float4 a[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long iu = i + u * stride;
      if (iu < hi) {
#pragma unroll
        for (int r = 0; r < W; ++r) a[u][r] = ld_peer16(p.buf[r] + ((base4 + iu) << 2));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long iu = i + u * stride;
      if (iu < hi) {
        float4 s = a[u][0];
#pragma unroll
        for (int r = 1; r < W; ++r) { s.x += a[u][r].x; s.y += a[u][r].y; s.z += a[u][r].z; s.w += a[u][r].w; }
        s.x *= p.scale; s.y *= p.scale; s.z *= p.scale; s.w *= p.scale;
#pragma unroll
        for (int r = 0; r < W; ++r) st_peer16(p.buf[r] + ((base4 + iu) << 2), s);
      }
    }
  }

  // 3. shard written everywhere: the last CTA of this rank signals and waits for all peers
  __threadfence_system();
  __syncthreads();
  __shared__ uint32_t s_last;
  if (threadIdx.x == 0) s_last = (atomicAdd(p.state + 1, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last) {
    __threadfence_system();
    if (threadIdx.x < W) {
      st_release_sys(p.flag[threadIdx.x] + COMM_MAX_WORLD + p.rank, v2);
      spin_until(p.flag[p.rank] + COMM_MAX_WORLD + threadIdx.x, v2);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      p.state[1] = 0;
      p.state[0] = v2;
      __threadfence();
    }
  }
}

}  // namespace mmdit

using namespace mmdit;

extern "C" int mmdit_comm_alloc(void** ptr, int64_t bytes) {
  MMDIT_REQUIRE(ptr && bytes > 0, MMDIT_ERR_ARG, "comm_alloc: bad arguments");
  cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
  if (e == cudaSuccess) e = cudaMemset(*ptr, 0, (size_t)bytes);
  if (e != cudaSuccess) {
    set_last_error("comm_alloc(%lld bytes): %s", (long long)bytes, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

extern "C" int mmdit_comm_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) {
    set_last_error("comm_free: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

extern "C" int mmdit_comm_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

extern "C" int mmdit_comm_export(void* ptr, void* handle_out) {
  MMDIT_REQUIRE(ptr && handle_out, MMDIT_ERR_ARG, "comm_export: null argument");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) {
    set_last_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    return (int)e;
  }
  memcpy(handle_out, &h, sizeof(h));
  return 0;
}

extern "C" int mmdit_comm_import(const void* handle, void** ptr_out) {
  MMDIT_REQUIRE(handle && ptr_out, MMDIT_ERR_ARG, "comm_import: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    set_last_error("cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

extern "C" int mmdit_comm_close(void* peer_ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(peer_ptr);
  if (e != cudaSuccess) {
    set_last_error("cudaIpcCloseMemHandle: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

extern "C" int mmdit_allreduce_mean_f32(const mmdit_comm* c, int64_t offset, int64_t n, int32_t ctas,
                                        void* stream) {
  MMDIT_REQUIRE(c && c->state, MMDIT_ERR_ARG, "allreduce: null communicator");
  MMDIT_REQUIRE(c->world >= 1 && c->world <= COMM_MAX_WORLD && c->rank >= 0 && c->rank < c->world,
                MMDIT_ERR_ARG, "allreduce: world %d rank %d", c->world, c->rank);
  MMDIT_REQUIRE(offset >= 0 && n > 0 && offset % 4 == 0 && n % 4 == 0, MMDIT_ERR_ALIGN,
                "allreduce: bucket offset/length must be multiples of 4 floats (got %lld, %lld)",
                (long long)offset, (long long)n);
  CommParams p;
  memset(&p, 0, sizeof(p));
  for (int r = 0; r < c->world; ++r) {
    MMDIT_REQUIRE(c->buf[r] && c->flag[r], MMDIT_ERR_ARG, "allreduce: rank %d not mapped", r);
    p.buf[r] = static_cast<float*>(c->buf[r]);
    p.flag[r] = static_cast<uint32_t*>(c->flag[r]);
  }
  p.state = static_cast<uint32_t*>(c->state);
  p.world = c->world; p.rank = c->rank;
  p.off = offset; p.n = n;
  p.scale = 1.0f / (float)c->world;
  if (ctas <= 0) ctas = 48;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // peer loads in flight per thread, as a multiple of the per-world base.  Measured at 8 GPUs (cfg2 step,
  // profiles/r02_bench_cfg2_n8_*): x1 31.5 ms, x2 30.6 ms, x4 31.6 ms -> x2 at world 8; MMDIT_COMM_UNROLL overrides.
  static const int forced = [] {
    const char* e = getenv("MMDIT_COMM_UNROLL");
    const int m = e ? atoi(e) : 0;
    return (m == 1 || m == 2 || m == 4) ? m : 0;
  }();
  const int mult = forced ? forced : (c->world == 8 ? 2 : 1);
#define COMM_LAUNCH(WW, UU)                                            \
  do {                                                                 \
    MMDIT_CARVEOUT((allreduce_mean_kernel<WW, UU>));                   \
    allreduce_mean_kernel<WW, UU><<<ctas, COMM_THREADS, 0, s>>>(p);    \
  } while (0)
#define COMM_CASE(WW, U1)                                              \
  case WW:                                                             \
    if (mult == 4) COMM_LAUNCH(WW, 4 * U1);                            \
    else if (mult == 2) COMM_LAUNCH(WW, 2 * U1);                       \
    else COMM_LAUNCH(WW, U1);                                          \
    break;
  switch (c->world) {
    COMM_CASE(1, 8)
    COMM_CASE(2, 4)
    COMM_CASE(4, 2)
    COMM_CASE(8, 1)
    default:
      set_last_error("allreduce: world size %d not supported (1, 2, 4, 8)", c->world);
      return MMDIT_ERR_UNSUPPORTED;
  }
#undef COMM_CASE
#undef COMM_LAUNCH
  return check_launch("allreduce_mean_kernel");
}
