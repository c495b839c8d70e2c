// Multi-GPU helpers of the C ABI: one process per GPU, the per-image path shards with no data-path
// dependency, so the only exchanges are a start-up broadcast of the constants (FLAME bases, folded encoder weights) and a
// per-batch all-gather of the outputs (params, vertices, landmarks).  Both go through NCCL over NVLink / NVSwitch.
// libnccl is bound at RUN time with dlopen (the copy PyTorch has already loaded when there is one, else the system's), so
// libdad3d.so keeps loading on boxes without NCCL and has no link-time dependency on it.
#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <string>

#include "../../include/dad3d.h"
#include "common.h"

namespace {

// the handful of NCCL entry points used, declared locally (ABI-stable since NCCL 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { kNcclInt8 = 0 };

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);        // the copy already in the process (torch's)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.lib = h;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(dlsym(h, "ncclBroadcast"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  });
  return (api.lib && api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Broadcast && api.AllGather) ? &api : nullptr;
}

int nccl_fail(const char* what, ncclResult_t r) {
  NcclApi* a = nccl();
  dad3d::set_error(std::string(what) + " -> " + (a && a->GetErrorString ? a->GetErrorString(r) : "NCCL error"));
  return DAD3D_ERR_CUDA;
}

}  // namespace

struct dad3d_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
};

extern "C" {

int dad3d_comm_unique_id(uint8_t* id128_h) {
  DAD3D_REQUIRE(id128_h, "null pointer");
  NcclApi* a = nccl();
  if (!a) { dad3d::set_error("libnccl.so.2 could not be loaded"); return DAD3D_ERR_UNSUPPORTED; }
  ncclUniqueId id;
  ncclResult_t r = a->GetUniqueId(&id);
  if (r != 0) return nccl_fail("ncclGetUniqueId", r);
  std::memcpy(id128_h, id.internal, 128);
  return DAD3D_OK;
}

int dad3d_comm_init(dad3d_comm** out, const uint8_t* id128_h, int32_t rank, int32_t world, int32_t device) {
  DAD3D_REQUIRE(out && id128_h && world >= 1 && rank >= 0 && rank < world, "arguments");
  NcclApi* a = nccl();
  if (!a) { dad3d::set_error("libnccl.so.2 could not be loaded"); return DAD3D_ERR_UNSUPPORTED; }
  DAD3D_CUDA_OK(cudaSetDevice(device));
  ncclUniqueId id;
  std::memcpy(id.internal, id128_h, 128);
  dad3d_comm* c = new dad3d_comm();
  c->rank = rank; c->world = world; c->device = device;
  ncclResult_t r = a->CommInitRank(&c->comm, world, id, rank);
  if (r != 0) { delete c; return nccl_fail("ncclCommInitRank", r); }
  *out = c;
  return DAD3D_OK;
}

void dad3d_comm_destroy(dad3d_comm* c) {
  if (!c) return;
  NcclApi* a = nccl();
  if (a && c->comm) a->CommDestroy(c->comm);
  delete c;
}

int dad3d_bcast_constants(dad3d_comm* c, void* buf_d, size_t bytes, int32_t root, dad3d_stream stream) {
  DAD3D_REQUIRE(c && buf_d && root >= 0 && root < c->world, "arguments");
  if (bytes == 0) return DAD3D_OK;
  ncclResult_t r = nccl()->Broadcast(buf_d, buf_d, bytes, kNcclInt8, root, c->comm, reinterpret_cast<cudaStream_t>(stream));
  if (r != 0) return nccl_fail("ncclBroadcast", r);
  return DAD3D_OK;
}

int dad3d_allgather_outputs(dad3d_comm* c, const void* send_d, void* recv_d, size_t bytes_per_rank, dad3d_stream stream) {
  DAD3D_REQUIRE(c && send_d && recv_d, "arguments");
  if (bytes_per_rank == 0) return DAD3D_OK;
  ncclResult_t r = nccl()->AllGather(send_d, recv_d, bytes_per_rank, kNcclInt8, c->comm, reinterpret_cast<cudaStream_t>(stream));
  if (r != 0) return nccl_fail("ncclAllGather", r);
  return DAD3D_OK;
}

}  // extern "C"
