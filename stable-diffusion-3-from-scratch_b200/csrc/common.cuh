// Shared device/host helpers for the sm_100a MMDiT kernels: PTX wrappers for
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM), and small bf16
// vector utilities.  Everything here is inline PTX; no CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace mmdit {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- errors ---
// Negative codes are argument errors, positive codes are cudaError_t values.
enum {
  MMDIT_OK = 0,
  MMDIT_ERR_ARG = -1,
  MMDIT_ERR_ALIGN = -2,
  MMDIT_ERR_UNSUPPORTED = -3,
  MMDIT_ERR_DRIVER = -4,
};
void set_last_error(const char* fmt, ...);
int check_launch(const char* what, int kernels = 1);  // cudaGetLastError -> code + message; counts launches

#define MMDIT_REQUIRE(cond, code, ...)  \
  do {                                  \
    if (!(cond)) {                      \
      set_last_error(__VA_ARGS__);      \
      return (code);                    \
    }                                   \
  } while (0)

// host: TMA descriptor encode (driver entry point resolved at run time)
int encode_tmap(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes /* rank-1 entries */, const uint32_t* box,
                int elem_bytes, bool swizzle128);
int num_sms();
// Ask for the maximum shared-memory carve-out for a kernel that itself needs (almost) none.
// An SM runs CTAs of different kernels side by side only under ONE L1/shared split; the GEMM and
// attention kernels need the 228 KB split, so a streaming / exchange kernel that keeps the default
// split cannot share an SM with them (and, once resident, keeps their CTAs out).  With the same
// preference the row kernels of one stream and the peer-memory exchange overlap the tensor-core
// kernels of another.  An experiment knob: no effect was measurable at N = 1, so it is OFF unless
// MMDIT_CARVEOUT=1 (multi-GPU overlap of the exchange kernel not yet measured with it).
void prefer_max_smem_carveout(const void* kernel);
#define MMDIT_CARVEOUT(kernel)                                                            \
  do {                                                                                    \
    static bool once_ = (mmdit::prefer_max_smem_carveout((const void*)(kernel)), true);   \
    (void)once_;                                                                          \
  } while (0)

// ---- second-generation row kernels (rowwise2.cu / qknorm2.cu) ------------------------------
// row_kernel_generation(): 2 unless MMDIT_ROW_KERNELS=1 or mmdit_set_row_kernel_generation(1).
// Each *_v2 launcher returns ROW_V2_UNSUPPORTED when the shape stays on the first generation.
constexpr int ROW_V2_UNSUPPORTED = -1000;
int row_kernel_generation();
int ln_modulate_fwd_v2(const void* x, const void* shift, const void* scale, void* y, float* mean, float* rstd,
                       long long rows, int d, long long rows_per_batch, long long ld_mod, float eps,
                       cudaStream_t stream);
int gate_residual_ln_fwd_v2(const void* a, const void* gate, const void* resid, const void* shift,
                            const void* scale, void* x_out, void* y, float* mean, float* rstd, long long rows,
                            int d, long long rows_per_batch, long long ld_gate, long long ld_mod, float eps,
                            cudaStream_t stream);
int ln_modulate_bwd_v2(const void* dy, const void* x, const float* mean, const float* rstd, const void* scale,
                       const void* dres, void* dx, void* dshift, void* dscale, int dmod_bf16, long long ld_dmod,
                       float* workspace, long long rows, int d, long long rows_per_batch, long long ld_mod,
                       int* bpb_out, cudaStream_t stream);
int gate_bwd_v2(const void* dout, const void* a, const void* gate, void* da, void* dgate, int dgate_bf16,
                float* dab, long long ld_dgate, long long ld_dab, float* workspace, long long rows, int d,
                long long rows_per_batch, long long ld_gate, int* bpb_out, cudaStream_t stream);
int qknorm_rope_fwd_v2(const void* qkv, const float* wq, const float* wk, const float* rope_cos,
                       const float* rope_sin, void* out, long long rows, int d, long long ld_in,
                       long long ld_out, int T, float eps, cudaStream_t stream);
int qknorm_rope_bwd_v2(const float* dq_acc, int acc_tokens, int acc_tok_off, const void* dqk, const void* qkv,
                       const float* wq, const float* wk, const float* rope_cos, const float* rope_sin,
                       void* dqkv, float* dwq, float* dwk, long long rows, int d, long long ld_g,
                       long long ld_in, long long ld_dout, int T, float eps, cudaStream_t stream);

// ---------------------------------------------------------------- launch ---
// Programmatic dependent launch: a kernel launched through launch_k() may be scheduled while the
// previous kernel of its stream is still running (as soon as every CTA of that kernel has executed
// pdl_trigger() or exited), runs its prologue (barrier init, TMEM allocation, descriptor prefetch)
// and blocks in pdl_wait() until the previous kernel has completed and its writes are visible.  Every
// kernel launched this way calls pdl_wait() before its first global-memory access.  Inside a captured
// CUDA graph the launch becomes a programmatic edge.  MMDIT_PDL=0 turns the attribute off (the device
// instructions are then no-ops).
bool pdl_enabled();
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                     Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface in check_launch()
}
#endif

// ---------------------------------------------------------------- device ---
#ifdef __CUDACC__

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier -------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// ---- proxies / fences -------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// One lane of a CONVERGED warp (elect.sync).  tcgen05.mma / tcgen05.commit / TMA issue code guarded by
// this predicate compiles to straight-line uniform-datapath instructions; guarded by `lane == 0`
// instead, ptxas wraps every such instruction in an ELECT / BRA.U.ANY loop over the "possibly several"
// active lanes, which costs tens of cycles per MMA issue.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- TMA --------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// TMA reduce-add (fp32) of a swizzled smem tile into a global tensor; bulk-group completion.
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, const void* src, int c0, int c1,
                                                  int c2) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(m),
      "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// TMA store of a (swizzled) smem tile into a global tensor; bulk-group completion.  Elements of
// the box that fall outside the tensor are not written (ragged M / N edges need no masking).
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_wait_group_read1() {
  asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}
__device__ __forceinline__ void tma_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_group_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---- tcgen05: TMEM management ------------------------------------------------
// Warp-collective. ncols: power of two in [32, 512].
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}

// ---- tcgen05: descriptors + MMA ------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_128B, sm_100 version bits.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  d |= 2ull << 61;  // SWIZZLE_128B
  return d;
}
// K-major operand tile: rows of 64 bf16 (128 B), 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_saddr, int k16) {
  return make_smem_desc(tile_saddr + k16 * 32, 16, 1024);
}
// MN-major operand tile: 64-wide MN chunks, each [block_k rows][128 B]; chunks
// chunk_bytes apart (LBO); 8-k-row groups 1024 B apart (SBO).
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_saddr, int k16, uint32_t chunk_bytes) {
  return make_smem_desc(tile_saddr + k16 * 2048, chunk_bytes, 1024);
}
// Instruction descriptor, kind::f16, bf16 x bf16 -> fp32.
__device__ __forceinline__ uint32_t make_idesc_bf16(int m, int n, int a_mn, int b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;   // D format: f32
  d |= 1u << 7;   // A format: bf16
  d |= 1u << 10;  // B format: bf16
  d |= static_cast<uint32_t>(a_mn & 1) << 15;
  d |= static_cast<uint32_t>(b_mn & 1) << 16;
  d |= static_cast<uint32_t>(n >> 3) << 17;
  d |= static_cast<uint32_t>(m >> 4) << 24;
  return d;
}
// D[tmem] (+)= A[smem] * B[smem]; one thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the A operand in tensor memory (rows = TMEM lanes, 16 bf16 of K packed into 8
// consecutive 32-bit columns starting at `a_tmem`); B still comes from shared memory.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued MMAs of this thread retire.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// ---- CTA pairs (cluster of 2, tcgen05 cta_group::2) --------------------------------
// Two CTAs on the two SMs of one TPC run ONE 256-row MMA: each holds its own 128 rows of A,
// half of the B tile and its 128 accumulator rows in its own TMEM; the leader (rank 0) issues.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_saddr` (a shared::cta address) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_saddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_saddr) : "memory");
}
// TMA load into this CTA's smem whose completion bytes are credited to an mbarrier that may live
// in the peer CTA (`bar_cluster_saddr` is a shared::cluster address).
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster_saddr,
                                                 int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(m), "r"(bar_cluster_saddr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive (when the issued MMAs retire) on the barrier at this smem offset in BOTH CTAs of the pair.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .b16 m;\n\t"
      "mov.b16 m, 3;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(
          smem_u32(bar))
      : "memory");
}

// ---- tcgen05: TMEM <-> registers ------------------------------------------------
// 32 lanes (this warp's sub-partition) x 32 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
      "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
      "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
      "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ---- bf16 vector helpers ---------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
// 8 bf16 <-> 8 floats through one 128-bit access
__device__ __forceinline__ void load8(const bf16* p, float (&f)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z),
         d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void store8(bf16* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// ---- packed fp32x2 (FADD2 / FMUL2 / FFMA2: one issue slot, two columns) ---------------------
// ptxas may still contract a mul.f32x2 feeding an add into one FFMA2 / FFMA.
__device__ __forceinline__ unsigned long long f2_bits(float2 v) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y));
  return r;
}
__device__ __forceinline__ float2 f2_from(unsigned long long r) {
  float2 v;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r));
  return v;
}
__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
  return f2_from(r);
}
__device__ __forceinline__ float2 f2_mul(float2 a, float2 b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return f2_from(r);
}
__device__ __forceinline__ float2 f2_add(float2 a, float2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return f2_from(r);
}
__device__ __forceinline__ float2 f2_dup(float v) { return make_float2(v, v); }
// bf16x2 word -> two floats (exact): low half = element 2p, high half = element 2p + 1
__device__ __forceinline__ float2 bf2_unpack(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ uint32_t bf2_pack(float2 v) { return pack_bf16x2(v.x, v.y); }
__device__ __forceinline__ uint32_t word(const uint4& u, int p) {
  return p == 0 ? u.x : p == 1 ? u.y : p == 2 ? u.z : u.w;
}

// x * sigmoid(x); fast reciprocal (MUFU.RCP + FMUL, ~2 ulp): every consumer rounds to bf16 anyway,
// and the IEEE division sequence is ~3x the instructions inside the 4-warp GEMM epilogues.
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.f + __expf(-x)); }

#endif  // __CUDACC__
}  // namespace mmdit
