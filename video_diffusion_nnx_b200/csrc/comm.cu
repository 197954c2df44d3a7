// Gradient exchange of the data-parallel training step behind the C ABI.
//
// The reference shards the batch over the `data` mesh axis and lets GSPMD insert the gradient all-reduce of the
// pjit'd step (trainer.py:307-326, 363-364). Here the exchange is explicit: one NCCL communicator per process
// (one process per GPU), vdn_allreduce_bucket() enqueues ncclAllReduce(sum) of one contiguous slice of the flat
// gradient on the caller's stream. The call is CUDA-graph capturable, so the whole training step - forward,
// backward, bucket reductions on a forked stream, optimizer - replays as ONE graph launch.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2, which is already mapped in a torch process), so the
// library itself has no link-time dependency and loads on a box without NCCL; the comm entry points then
// return VDN_E_ARCH.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <vector>

#include "vdn_common.cuh"
#include "vdn_host.h"

namespace vdn {
namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  bool ok = false;
};

NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    // RTLD_NOLOAD first: reuse the copy torch (or the host framework) already mapped, so both see one NCCL
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    a.handle = h;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    a.CommInitRankConfig = reinterpret_cast<decltype(a.CommInitRankConfig)>(dlsym(h, "ncclCommInitRankConfig"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(h, "ncclAllReduce"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(dlsym(h, "ncclGetVersion"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.AllReduce && a.CommDestroy && a.GetErrorString;
  });
  return a;
}

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};
std::mutex g_comm_mu;
std::vector<Comm*> g_comms;

}  // namespace

void comm_shutdown() {
  std::lock_guard<std::mutex> lk(g_comm_mu);
  for (Comm* c : g_comms) {
    if (c->comm && api().ok) api().CommDestroy(c->comm);
    delete c;
  }
  g_comms.clear();
}

}  // namespace vdn

using namespace vdn;

#define VDN_NCCL(call, what)                                                                      \
  do {                                                                                            \
    ncclResult_t r_ = (call);                                                                     \
    VDN_REQUIRE(r_ == ncclSuccess, VDN_E_CUDA, "%s: NCCL error %d (%s)", what, (int)r_, api().GetErrorString(r_)); \
  } while (0)

extern "C" int vdn_comm_unique_id_bytes(void) { return (int)sizeof(ncclUniqueId); }

extern "C" int vdn_comm_unique_id(void* id_host) {
  VDN_REQUIRE(id_host != nullptr, VDN_E_SHAPE, "comm_unique_id: null buffer");
  VDN_REQUIRE(api().ok, VDN_E_ARCH, "NCCL (libnccl.so.2) could not be loaded");
  ncclUniqueId id;
  VDN_NCCL(api().GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(id_host, &id, sizeof(id));
  return VDN_OK;
}

extern "C" int vdn_comm_init(void** comm_out, const void* id_host, int rank, int world, int max_ctas) {
  VDN_REQUIRE(comm_out && id_host && world >= 1 && rank >= 0 && rank < world, VDN_E_SHAPE, "comm_init: bad arguments");
  VDN_REQUIRE(api().ok, VDN_E_ARCH, "NCCL (libnccl.so.2) could not be loaded");
  ncclUniqueId id;
  memcpy(&id, id_host, sizeof(id));
  Comm* c = new Comm;
  c->rank = rank;
  c->world = world;
  ncclResult_t r;
  if (max_ctas > 0 && api().CommInitRankConfig) {
    // cap the CTAs NCCL may occupy: the reduction overlaps a backward pass that is itself latency-bound, and every
    // SM NCCL holds is one the dependency chain does not get
    ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
    cfg.maxCTAs = max_ctas;
    cfg.minCTAs = 1;
    r = api().CommInitRankConfig(&c->comm, world, id, rank, &cfg);
  } else {
    r = api().CommInitRank(&c->comm, world, id, rank);
  }
  if (r != ncclSuccess) {
    set_last_error("ncclCommInitRank: NCCL error %d (%s)", (int)r, api().GetErrorString(r));
    delete c;
    return VDN_E_CUDA;
  }
  {
    std::lock_guard<std::mutex> lk(g_comm_mu);
    g_comms.push_back(c);
  }
  *comm_out = c;
  return VDN_OK;
}

extern "C" int vdn_comm_world(const void* comm, int* rank, int* world) {
  VDN_REQUIRE(comm != nullptr, VDN_E_SHAPE, "comm_world: null communicator");
  const Comm* c = reinterpret_cast<const Comm*>(comm);
  if (rank) *rank = c->rank;
  if (world) *world = c->world;
  return VDN_OK;
}

extern "C" int vdn_allreduce_bucket(void* comm, void* buf, long count, int dtype, void* stream) {
  VDN_REQUIRE(comm && buf && count > 0, VDN_E_SHAPE, "allreduce_bucket: bad arguments");
  VDN_REQUIRE(dtype == VDN_F32 || dtype == VDN_BF16, VDN_E_SHAPE, "allreduce_bucket: dtype must be VDN_F32 or VDN_BF16");
  Comm* c = reinterpret_cast<Comm*>(comm);
  VDN_NCCL(api().AllReduce(buf, buf, (size_t)count, dtype == VDN_F32 ? ncclFloat32 : ncclBfloat16, ncclSum, c->comm,
                           reinterpret_cast<cudaStream_t>(stream)),
           "ncclAllReduce");
  return VDN_OK;
}

extern "C" int vdn_comm_destroy(void* comm) {
  if (!comm) return VDN_OK;
  Comm* c = reinterpret_cast<Comm*>(comm);
  {
    std::lock_guard<std::mutex> lk(g_comm_mu);
    for (size_t i = 0; i < g_comms.size(); ++i)
      if (g_comms[i] == c) {
        g_comms.erase(g_comms.begin() + i);
        break;
      }
  }
  if (c->comm && api().ok) api().CommDestroy(c->comm);
  delete c;
  return VDN_OK;
}
