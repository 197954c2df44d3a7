// Host-side runtime glue: error reporting, driver entry points (TMA descriptor encode),
// version query. No allocation, no synchronisation.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {

static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};

bool pdl_enabled() {
  static const bool on = getenv("VDN_NO_PDL") == nullptr;
  return on;
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return VDN_E_CUDA;
  }
  return VDN_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    cudaGetLastError();
  });
  return fn;
}

// Encodes a bf16 tiled tensor map. dims/strides innermost first; strides in BYTES for
// dims 1..rank-1 (dim 0 is contiguous). swizzle_bytes in {0,32,64,128}.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  VDN_REQUIRE(fn != nullptr, VDN_E_ARCH, "cuTensorMapEncodeTiled driver entry point unavailable");
  VDN_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, VDN_E_ALIGN, "TMA base pointer must be 16B aligned");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i - 1];
      VDN_REQUIRE((gstr[i - 1] & 15) == 0, VDN_E_ALIGN, "TMA stride %d (%llu B) must be a multiple of 16", i,
                  (unsigned long long)gstr[i - 1]);
    }
  }
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx,
                  es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VDN_REQUIRE(r == CUDA_SUCCESS, VDN_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return VDN_OK;
}

}  // namespace vdn

extern "C" int vdn_version(void) { return 100; }
// Number of kernels this library has launched (or recorded into a CUDA graph) in this process.
extern "C" unsigned long long vdn_launch_count(void) { return vdn::g_launches.load(); }
extern "C" const char* vdn_last_error(void) { return vdn::g_err; }
