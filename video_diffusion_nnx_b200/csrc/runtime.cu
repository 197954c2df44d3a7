// Host-side runtime glue: error reporting, driver entry points (TMA descriptor encode),
// version query. No allocation, no synchronisation.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

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

// ---- experiment / test switches -------------------------------------------------------------------------------
namespace {
struct TuneEntry {
  char name[40];
  int value;
};
constexpr int kMaxTune = 64;
TuneEntry g_tune[kMaxTune];
std::atomic<int> g_tune_n{0};
std::mutex g_tune_mu;

int tune_find(const char* name) {
  const int n = g_tune_n.load(std::memory_order_acquire);
  for (int i = 0; i < n; ++i)
    if (strcmp(g_tune[i].name, name) == 0) return i;
  return -1;
}
void tune_store(const char* name, int value) {
  std::lock_guard<std::mutex> lk(g_tune_mu);
  int i = tune_find(name);
  if (i < 0) {
    i = g_tune_n.load();
    if (i >= kMaxTune) return;
    strncpy(g_tune[i].name, name, sizeof(g_tune[i].name) - 1);
    g_tune[i].value = value;
    g_tune_n.store(i + 1, std::memory_order_release);
  } else {
    g_tune[i].value = value;
  }
}
constexpr int kTuneUnset = INT32_MIN;
#ifdef VDN_DEBUG
// debug builds: the first lookup of a name falls back to the environment variable of the same name
int tune_env(const char* name) {
  const char* e = getenv(name);
  const int v = e ? (e[0] ? atoi(e) : 1) : kTuneUnset;
  tune_store(name, v);
  return v;
}
#endif
int tune_value(const char* name) {
  const int i = tune_find(name);
  if (i >= 0) return g_tune[i].value;
#ifdef VDN_DEBUG
  return tune_env(name);
#else
  return kTuneUnset;
#endif
}
}  // namespace

bool tune_is_set(const char* name) { return tune_value(name) != kTuneUnset; }
int tune_int(const char* name, int dflt) {
  const int v = tune_value(name);
  return v == kTuneUnset ? dflt : v;
}
bool tune_on(const char* name) {
  const int v = tune_value(name);
  return v != kTuneUnset && v != 0;
}

bool pdl_enabled() {
  return !tune_on("VDN_NO_PDL");
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

// Tensor-map cache: a CUtensorMap is a pure function of (base pointer, element type, rank, dims, strides, box,
// swizzle), and cuTensorMapEncodeTiled costs ~2-5 us of host time per call - several per launch. Engines call the
// same (pointer, shape) combinations every step, so eager / XLA-FFI callers (no CUDA graph to hide the host work)
// would be host-bound without it. Entries are never stale (the key is the whole input of the encode).
namespace {
struct TmapKey {
  uint64_t base;
  uint64_t dims[5];
  uint64_t strides[4];
  uint32_t box[5];
  uint32_t rank, swizzle, dtype;
  bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return (size_t)h;
  }
};
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps;
std::mutex g_tmap_mu;
std::atomic<unsigned long long> g_tmap_hits{0}, g_tmap_misses{0};
constexpr size_t kTmapCacheMax = 1 << 16;
}  // namespace

void tmap_cache_clear() {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  g_tmaps.clear();
}

static int encode_tmap(CUtensorMap* out, CUtensorMapDataType dt, int esize, const void* base, int rank,
                       const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  VDN_REQUIRE(rank >= 1 && rank <= 5, VDN_E_SHAPE, "TMA rank %d out of range", rank);
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = reinterpret_cast<uint64_t>(base);
  key.rank = (uint32_t)rank;
  key.swizzle = (uint32_t)swizzle_bytes;
  key.dtype = (uint32_t)dt;
  for (int i = 0; i < rank; ++i) {
    key.dims[i] = dims[i];
    key.box[i] = box[i];
    if (i > 0) key.strides[i - 1] = strides_bytes[i - 1];
  }
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmaps.find(key);
    if (it != g_tmaps.end()) {
      *out = it->second;
      g_tmap_hits.fetch_add(1, std::memory_order_relaxed);
      return VDN_OK;
    }
  }
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
  (void)esize;
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VDN_REQUIRE(r == CUDA_SUCCESS, VDN_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  g_tmap_misses.fetch_add(1, std::memory_order_relaxed);
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmaps.size() >= kTmapCacheMax) g_tmaps.clear();
    g_tmaps.emplace(key, *out);
  }
  return VDN_OK;
}

// Encodes (or fetches from the cache) a bf16 tiled tensor map. dims/strides innermost first; strides in BYTES for
// dims 1..rank-1 (dim 0 is contiguous). swizzle_bytes in {0,32,64,128}.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  return encode_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, rank, dims, strides_bytes, box, swizzle_bytes);
}
// Same for fp32 elements (the fp32-grade path).
int encode_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  return encode_tmap(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, rank, dims, strides_bytes, box, swizzle_bytes);
}

}  // namespace vdn

extern "C" int vdn_version(void) { return 200; }
// Test / tool hook for the experiment switches: set (enable != 0) or clear one by its VDN_* name.
extern "C" int vdn_debug_set(const char* name, int value, int enable) {
  VDN_REQUIRE(name && strlen(name) < 40, VDN_E_SHAPE, "vdn_debug_set: bad name");
  vdn::tune_store(name, enable ? value : vdn::kTuneUnset);
  return VDN_OK;
}
// Debug: device buffer (>= 1024 int64) that kernels with a timeline hook stamp with clock64(); NULL switches it off.
static long long* g_debug_trace = nullptr;
namespace vdn {
long long* debug_trace_ptr() { return g_debug_trace; }
}
extern "C" void vdn_debug_trace_buffer(void* dev_buf) { g_debug_trace = reinterpret_cast<long long*>(dev_buf); }
// Number of kernels this library has launched (or recorded into a CUDA graph) in this process.
extern "C" unsigned long long vdn_launch_count(void) { return vdn::g_launches.load(); }
extern "C" const char* vdn_last_error(void) { return vdn::g_err; }
// Tensor-map cache statistics (hits, misses) since load: evidence that repeated launches do not re-encode.
extern "C" int vdn_tmap_cache_stats(unsigned long long* hits, unsigned long long* misses) {
  if (hits) *hits = vdn::g_tmap_hits.load();
  if (misses) *misses = vdn::g_tmap_misses.load();
  return VDN_OK;
}
// Releases everything the library owns (cached tensor maps, communicators). The library never owns device memory.
extern "C" int vdn_shutdown(void) {
  vdn::tmap_cache_clear();
  vdn::comm_shutdown();
  return VDN_OK;
}

// ---- scratch sizes (bytes) of the entry points that take a workspace argument (include/vdn.h) ----
extern "C" size_t vdn_sla_workspace_floats(int n_img, int N);
extern "C" size_t vdn_sla_core_fwd_workspace(int n_img, int N) { return vdn_sla_workspace_floats(n_img, N) * sizeof(float); }
extern "C" size_t vdn_sla_core_bwd_workspace(int n_img) { return (size_t)n_img * 8 * 32 * 32 * sizeof(float); }
extern "C" size_t vdn_gn_silu_bwd_workspace(int B, int C) { return (size_t)B * C * 2 * sizeof(float); }
extern "C" size_t vdn_mha_core_bwd_workspace(long P) { return (size_t)P * 8 * sizeof(float); }
extern "C" size_t vdn_time_heads_bwd_workspace(int B, int ss_ld) { return (size_t)B * ss_ld * sizeof(float); }
extern "C" size_t vdn_time_mlp_bwd_workspace(int B, int dim) { return (size_t)B * 4 * dim * sizeof(float); }
