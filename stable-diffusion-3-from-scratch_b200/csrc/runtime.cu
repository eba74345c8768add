// Host-side runtime glue for the C-ABI: last-error string, launch checking,
// TMA descriptor encoding through the driver entry point (no -lcuda link).
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "mmdit_b200.h"

namespace mmdit {

static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;

int check_launch(const char* what, int kernels) {
  __atomic_fetch_add(&g_launches, (unsigned long long)kernels, __ATOMIC_RELAXED);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: %s", what, cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return MMDIT_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int encode_tmap(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int elem_bytes,
                bool swizzle128) {
  EncodeTiledFn fn = get_encode_fn();
  MMDIT_REQUIRE(fn != nullptr, MMDIT_ERR_DRIVER,
                "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  MMDIT_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, MMDIT_ERR_ALIGN,
                "TMA base address %p not 16-byte aligned", base);
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    MMDIT_REQUIRE((gstr[i] & 15) == 0, MMDIT_ERR_ALIGN,
                  "TMA stride %llu bytes (dim %d) not a multiple of 16",
                  (unsigned long long)gstr[i], i + 1);
  }
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                           : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(map, dt, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MMDIT_REQUIRE(r == CUDA_SUCCESS, MMDIT_ERR_DRIVER,
                "cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)",
                (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                box[0], rank > 1 ? box[1] : 0);
  return MMDIT_OK;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("MMDIT_PDL");
    return e ? atoi(e) != 0 : true;
  }();
  return on;
}

void prefer_max_smem_carveout(const void* kernel) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MMDIT_CARVEOUT");
    on = e ? atoi(e) : 0;   // measured at N = 1 (cfg2, same box): 31.9 / 32.1 ms on vs 32.0 / 31.8 ms off
  }
  if (on)
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaGetLastError();   // a hint: never an error for the caller
}

int num_sms() {
  static int n = 0;
  if (n) return n;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  return n;
}

}  // namespace mmdit

extern "C" {

const char* mmdit_last_error(void) { return mmdit::g_err; }

int mmdit_abi_version(void) { return MMDIT_ABI_VERSION; }

unsigned long long mmdit_launch_count(void) {
  return __atomic_load_n(&mmdit::g_launches, __ATOMIC_RELAXED);
}

int mmdit_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    mmdit::set_last_error("cudaGetDevice: %s", cudaGetErrorString(e));
    return (int)e;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    mmdit::set_last_error("device compute capability %d.%d is not sm_100 (B200)", major, minor);
    return mmdit::MMDIT_ERR_UNSUPPORTED;
  }
  return 0;
}

}  // extern "C"
