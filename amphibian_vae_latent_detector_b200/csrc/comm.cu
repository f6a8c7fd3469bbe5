// comm.cu -- the two exchanges of the sharded radial fit for hosts without torch.distributed: a per-context
// NCCL communicator, the all-reduce of the per-species sums / counts (every rank then forms the same centroids) and the
// all-gather of the local radii + labels (every rank then selects the same order statistics).  NCCL is resolved at run
// time with dlopen (the instance the process already holds, e.g. the one PyTorch ships, else libnccl.so.2 from the loader
// path), so libavld.so has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "common.cuh"

namespace {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.ok ? &api : nullptr;
  tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // already in the process (PyTorch's)?
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return nullptr;
  auto sym = [&](const char* name) { return dlsym(h, name); };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather && api.GroupStart &&
           api.GroupEnd && api.GetErrorString;
  return api.ok ? &api : nullptr;
}

#define AVLD_NCCL(api, expr)                                                                   \
  do {                                                                                         \
    const ncclResult_t r_ = (expr);                                                            \
    if (r_ != ncclSuccess) {                                                                   \
      avld::set_error("%s failed: %s", #expr, (api)->GetErrorString(r_));                      \
      return AVLD_ERR_CUDA;                                                                    \
    }                                                                                          \
  } while (0)

}  // namespace

using namespace avld;

extern "C" int avld_comm_unique_id(void* id_out) {
  AVLD_CHECK(id_out != nullptr, AVLD_ERR_INVALID, "NULL argument");
  NcclApi* api = nccl_api();
  AVLD_CHECK(api != nullptr, AVLD_ERR_UNSUPPORTED, "NCCL (libnccl.so.2) could not be loaded");
  ncclUniqueId id;
  AVLD_NCCL(api, api->GetUniqueId(&id));
  static_assert(sizeof(id) == AVLD_COMM_ID_BYTES, "ncclUniqueId size");
  std::memcpy(id_out, &id, sizeof(id));
  return AVLD_OK;
}

extern "C" int avld_comm_init(avld_ctx* c, const void* nccl_unique_id, int32_t rank, int32_t world) {
  AVLD_ENTER(c);
  AVLD_CHECK(nccl_unique_id != nullptr && world >= 1 && rank >= 0 && rank < world, AVLD_ERR_INVALID, "bad id / rank / world");
  AVLD_CHECK(c->comm == nullptr, AVLD_ERR_STATE, "this context already has a communicator");
  NcclApi* api = nccl_api();
  AVLD_CHECK(api != nullptr, AVLD_ERR_UNSUPPORTED, "NCCL (libnccl.so.2) could not be loaded");
  ncclUniqueId id;
  std::memcpy(&id, nccl_unique_id, sizeof(id));
  ncclComm_t comm = nullptr;
  AVLD_NCCL(api, api->CommInitRank(&comm, world, id, rank));
  c->comm = comm;
  c->comm_rank = rank;
  c->comm_world = world;
  return AVLD_OK;
}

extern "C" int avld_comm_destroy(avld_ctx* c) {
  AVLD_ENTER(c);
  if (c->comm == nullptr) return AVLD_OK;
  NcclApi* api = nccl_api();
  cudaDeviceSynchronize();
  if (api) api->CommDestroy(static_cast<ncclComm_t>(c->comm));
  c->comm = nullptr;
  c->comm_world = 0;
  if (c->d_gather_r) { cudaFree(c->d_gather_r); c->d_gather_r = nullptr; }
  if (c->d_gather_l) { cudaFree(c->d_gather_l); c->d_gather_l = nullptr; }
  c->gather_rows = 0;
  return AVLD_OK;
}

// 08:316 np.mean over ALL ranks' rows: sum [K, D] float64 and cnt [K] int64 are summed in place, one NCCL group
extern "C" int avld_allreduce_centroids(avld_ctx* c, double* sum, int64_t* cnt, int32_t K, int32_t D, void* stream) {
  AVLD_ENTER(c);
  AVLD_CHECK(c->comm != nullptr, AVLD_ERR_STATE, "avld_comm_init has not been called");
  AVLD_CHECK(sum && cnt && K >= 1 && D >= 1, AVLD_ERR_INVALID, "bad argument");
  NcclApi* api = nccl_api();
  ncclComm_t comm = static_cast<ncclComm_t>(c->comm);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AVLD_NCCL(api, api->GroupStart());
  AVLD_NCCL(api, api->AllReduce(sum, sum, static_cast<size_t>(K) * D, ncclFloat64, ncclSum, comm, st));
  AVLD_NCCL(api, api->AllReduce(cnt, cnt, static_cast<size_t>(K), ncclInt64, ncclSum, comm, st));
  AVLD_NCCL(api, api->GroupEnd());
  return AVLD_OK;
}

// radii [n_local, K] + label [n_local] of this rank -> radii_all [world * shard_rows, K], label_all [world * shard_rows]
// (rank r's rows at r * shard_rows; rows past a rank's n_local carry label -1, which the selection kernels skip)
extern "C" int avld_allgather_radii(avld_ctx* c, const float* radii, const int32_t* label, int64_t n_local, int64_t shard_rows,
                                    int32_t K, float* radii_all, int32_t* label_all, void* stream) {
  AVLD_ENTER(c);
  AVLD_CHECK(c->comm != nullptr, AVLD_ERR_STATE, "avld_comm_init has not been called");
  AVLD_CHECK(radii_all && label_all && K >= 1 && shard_rows >= 1 && n_local >= 0 && n_local <= shard_rows, AVLD_ERR_INVALID,
             "need 0 <= n_local <= shard_rows");
  AVLD_CHECK(n_local == 0 || (radii && label), AVLD_ERR_INVALID, "NULL argument");
  NcclApi* api = nccl_api();
  ncclComm_t comm = static_cast<ncclComm_t>(c->comm);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float* send_r = radii;
  const int32_t* send_l = label;
  if (n_local < shard_rows) {                   // pad this rank's block
    if (c->gather_rows < shard_rows * K) {
      if (c->d_gather_r) cudaFree(c->d_gather_r);
      if (c->d_gather_l) cudaFree(c->d_gather_l);
      c->d_gather_r = nullptr; c->d_gather_l = nullptr; c->gather_rows = 0;
      AVLD_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_gather_r), static_cast<size_t>(shard_rows) * K * sizeof(float)));
      AVLD_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_gather_l), static_cast<size_t>(shard_rows) * sizeof(int32_t)));
      c->gather_rows = shard_rows * K;
    }
    AVLD_CUDA(cudaMemsetAsync(c->d_gather_r, 0, static_cast<size_t>(shard_rows) * K * sizeof(float), st));
    AVLD_CUDA(cudaMemsetAsync(c->d_gather_l, 0xFF, static_cast<size_t>(shard_rows) * sizeof(int32_t), st));   // label -1
    if (n_local > 0) {
      AVLD_CUDA(cudaMemcpyAsync(c->d_gather_r, radii, static_cast<size_t>(n_local) * K * sizeof(float), cudaMemcpyDeviceToDevice, st));
      AVLD_CUDA(cudaMemcpyAsync(c->d_gather_l, label, static_cast<size_t>(n_local) * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    }
    send_r = c->d_gather_r;
    send_l = c->d_gather_l;
  }
  AVLD_NCCL(api, api->GroupStart());
  AVLD_NCCL(api, api->AllGather(send_r, radii_all, static_cast<size_t>(shard_rows) * K, ncclFloat32, comm, st));
  AVLD_NCCL(api, api->AllGather(send_l, label_all, static_cast<size_t>(shard_rows), ncclInt32, comm, st));
  AVLD_NCCL(api, api->GroupEnd());
  return AVLD_OK;
}
