// Fused multi-tensor optimizer step for the MMDiT trainer (reference: model_trainer.py:483-503 --
// GradScaler.unscale_, clip_grad_norm_(1.0), AdamW(lr, betas, eps 1e-8, weight_decay 0.01),
// zero_grad -- and the per-forward fp32->bf16 weight casts that torch.autocast performs).
//   kernel 1: sum of squared gradients over every tensor (one fp32 atomic per block)
//   kernel 2: clip coefficient + AdamW update of p/m/v + refresh of the bf16 shadow weight
// All scalars that change per step (step count, gradient norm) live in device memory, so the
// pair is CUDA-graph replayable.  Memory-bound: 16 B read + 14 B written per parameter.
#include "common.cuh"
#include "mmdit_b200.h"

namespace mmdit {

struct ParamDesc {        // mirrors mmdit_param_desc
  float* p;
  const float* g;
  float* m;
  float* v;
  bf16* shadow;           // may be null
  long long n;
};
constexpr int OPT_THREADS = 256;
constexpr int OPT_CHUNK = 16384;  // elements per block

__global__ void __launch_bounds__(OPT_THREADS)
grad_sumsq_kernel(const ParamDesc* __restrict__ table, const int2* __restrict__ chunks,
                  float* __restrict__ state /* [0]=sumsq */) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  __shared__ float red[OPT_THREADS / 32];
  const int2 w = chunks[blockIdx.x];
  const ParamDesc d = table[w.x];
  const long long beg = (long long)w.y * OPT_CHUNK;
  const long long end = min(beg + OPT_CHUNK, d.n);
  float acc = 0.f;
  if ((reinterpret_cast<uintptr_t>(d.g) & 15) == 0) {
    const long long n4 = (end - beg) / 4;
    const float4* g4 = reinterpret_cast<const float4*>(d.g + beg);
    for (long long i = threadIdx.x; i < n4; i += OPT_THREADS) {
      const float4 x = g4[i];
      acc += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
    for (long long i = beg + n4 * 4 + threadIdx.x; i < end; i += OPT_THREADS) acc += d.g[i] * d.g[i];
  } else {
    for (long long i = beg + threadIdx.x; i < end; i += OPT_THREADS) acc += d.g[i] * d.g[i];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s = threadIdx.x < OPT_THREADS / 32 ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(state, s);
  }
}

__global__ void __launch_bounds__(OPT_THREADS)
adamw_kernel(const ParamDesc* __restrict__ table, const int2* __restrict__ chunks,
             const float* __restrict__ state /* [0]=sumsq [1]=step (already incremented) [2]=lr */,
             float lr_arg, float beta1, float beta2, float eps, float wd, float max_norm) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  // lr < 0: the learning rate is the device-resident state[2] (a scheduler rewrites it between
  // replays of a captured step; by-value arguments are frozen into a CUDA graph)
  const float lr = lr_arg < 0.f ? state[2] : lr_arg;
  const int2 w = chunks[blockIdx.x];
  const ParamDesc d = table[w.x];
  const long long beg = (long long)w.y * OPT_CHUNK;
  const long long end = min(beg + OPT_CHUNK, d.n);
  const float norm = sqrtf(state[0]);
  const float clip = max_norm > 0.f ? fminf(1.f, max_norm / (norm + 1e-6f)) : 1.f;  // clip_grad_norm_
  const float step = state[1];
  const float bc1 = 1.f - powf(beta1, step), bc2 = 1.f - powf(beta2, step);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2), decay = 1.f - lr * wd;
  const bool vec = ((reinterpret_cast<uintptr_t>(d.p) | reinterpret_cast<uintptr_t>(d.g) |
                     reinterpret_cast<uintptr_t>(d.m) | reinterpret_cast<uintptr_t>(d.v)) & 15) == 0 &&
                   (!d.shadow || (reinterpret_cast<uintptr_t>(d.shadow + beg) & 7) == 0);
  auto upd = [&](float& p, float g, float& m, float& v) {
    g *= clip;
    p *= decay;                                    // decoupled weight decay
    m = beta1 * m + (1.f - beta1) * g;
    v = beta2 * v + (1.f - beta2) * g * g;
    p -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + eps);
  };
  if (vec) {
    const long long n4 = (end - beg) / 4;
    for (long long i = threadIdx.x; i < n4; i += OPT_THREADS) {
      const long long o = beg + 4 * i;
      float4 p = *reinterpret_cast<float4*>(d.p + o);
      const float4 g = *reinterpret_cast<const float4*>(d.g + o);
      float4 m = *reinterpret_cast<float4*>(d.m + o);
      float4 v = *reinterpret_cast<float4*>(d.v + o);
      upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
      *reinterpret_cast<float4*>(d.p + o) = p;
      *reinterpret_cast<float4*>(d.m + o) = m;
      *reinterpret_cast<float4*>(d.v + o) = v;
      if (d.shadow) {
        uint2 s;
        s.x = pack_bf16x2(p.x, p.y);
        s.y = pack_bf16x2(p.z, p.w);
        *reinterpret_cast<uint2*>(d.shadow + o) = s;
      }
    }
    for (long long i = beg + n4 * 4 + threadIdx.x; i < end; i += OPT_THREADS) {
      float p = d.p[i], m = d.m[i], v = d.v[i];
      upd(p, d.g[i], m, v);
      d.p[i] = p; d.m[i] = m; d.v[i] = v;
      if (d.shadow) d.shadow[i] = __float2bfloat16(p);
    }
  } else {
    for (long long i = beg + threadIdx.x; i < end; i += OPT_THREADS) {
      float p = d.p[i], m = d.m[i], v = d.v[i];
      upd(p, d.g[i], m, v);
      d.p[i] = p; d.m[i] = m; d.v[i] = v;
      if (d.shadow) d.shadow[i] = __float2bfloat16(p);
    }
  }
}

__global__ void opt_begin_kernel(float* state) {
  // (no early pdl_trigger: dependents are released when this grid exits)
  pdl_wait();   // programmatic dependent launch: the previous kernel's writes are visible from here
  state[0] = 0.f;     // sum of squares
  state[1] += 1.f;    // step count
}

}  // namespace mmdit

using namespace mmdit;

extern "C" int mmdit_adamw_step(const void* table, const void* chunks, int32_t n_chunks, float* state,
                                float lr, float beta1, float beta2, float eps, float weight_decay,
                                float max_norm, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MMDIT_REQUIRE(table && chunks && state && n_chunks > 0, MMDIT_ERR_ARG, "adamw_step: bad arguments");
  static_assert(sizeof(ParamDesc) == sizeof(mmdit_param_desc), "param desc layout");
  launch_k(opt_begin_kernel, dim3(1), dim3(1), 0, stream, state);
  launch_k(grad_sumsq_kernel, dim3(n_chunks), dim3(OPT_THREADS), 0, stream, static_cast<const ParamDesc*>(table),
                                                          static_cast<const int2*>(chunks), state);
  launch_k(adamw_kernel, dim3(n_chunks), dim3(OPT_THREADS), 0, stream, static_cast<const ParamDesc*>(table),
                                                     static_cast<const int2*>(chunks), state, lr, beta1,
                                                     beta2, eps, weight_decay, max_norm);
  return check_launch("adamw_kernel", 3);
}

extern "C" int mmdit_adamw_chunk_elems(void) { return OPT_CHUNK; }
