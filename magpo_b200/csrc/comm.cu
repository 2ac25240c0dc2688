// Gradient exchange of the data-parallel update behind the C ABI: the two `jax.lax.pmean(..., "device")` of rec_magpo.py:399-409
// as NCCL all-reduces (sum; the 1/Nd factor is applied by magpo_clip_adam's grad_scale) over NVLink / NVSwitch, issued by the
// library itself on the streams its kernels run on — no framework collective, no stream hand-off. libnccl is resolved at run time
// (dlopen of "libnccl.so.2": the copy already mapped into the process if there is one), so the library has no link-time dependency
// on it and single-GPU users never load it. Only the few NCCL declarations used here are restated (nccl.h 2.27: opaque
// communicator, 128-byte unique id, ncclFloat32 = 7, ncclSum = 0, ncclMax = 2).
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

namespace magpo {
namespace {

typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId { char internal[128]; };
typedef int ncclResult_t;  // 0 = ncclSuccess

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) return api;
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.handle, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.handle, "ncclCommInitRank"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(api.handle, "ncclAllReduce"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.handle, "ncclCommDestroy"));
  api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(dlsym(api.handle, "ncclGetVersion"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.handle, "ncclGetErrorString"));
  api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy;
  return api;
}

void set_nccl_error(ncclResult_t r, const char* what) {
  char msg[256];
  snprintf(msg, sizeof(msg), "NCCL %s: %s", what, nccl().GetErrorString ? nccl().GetErrorString(r) : "error");
  set_error_text(msg);
}

}  // namespace
}  // namespace magpo

struct MagpoComm {
  magpo::ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0, device = -1;
};

namespace magpo {
int comm_allreduce(MagpoComm* c, cudaStream_t s, float* buf, int64_t n, int op) {
  if (!c || !buf || n < 0) return MAGPO_ERR_ARG;
  if (n == 0 || c->nranks == 1) return MAGPO_OK;
  const ncclResult_t r = nccl().AllReduce(buf, buf, (size_t)n, /*ncclFloat32*/ 7, op, c->comm, s);
  if (r != 0) {
    set_nccl_error(r, "ncclAllReduce");
    return MAGPO_ERR_CUDA;
  }
  note_launch();
  return MAGPO_OK;
}
}  // namespace magpo

using namespace magpo;

extern "C" {

int magpo_comm_available(void) { return nccl().ok ? 1 : 0; }

int magpo_comm_version(void) {
  int v = 0;
  if (nccl().ok && nccl().GetVersion) nccl().GetVersion(&v);
  return v;
}

int magpo_comm_unique_id(void* id128) {
  if (!id128) return MAGPO_ERR_ARG;
  if (!nccl().ok) return MAGPO_ERR_UNSUPPORTED;
  ncclUniqueId id;
  const ncclResult_t r = nccl().GetUniqueId(&id);
  if (r != 0) {
    set_nccl_error(r, "ncclGetUniqueId");
    return MAGPO_ERR_CUDA;
  }
  memcpy(id128, id.internal, sizeof(id.internal));
  return MAGPO_OK;
}

int magpo_comm_init(int32_t nranks, int32_t rank, const void* id128, MagpoComm** out) {
  if (!out || nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !id128)) return MAGPO_ERR_ARG;
  MagpoComm* c = new MagpoComm();
  c->nranks = nranks;
  c->rank = rank;
  cudaGetDevice(&c->device);
  if (nranks > 1) {
    if (!nccl().ok) {
      delete c;
      return MAGPO_ERR_UNSUPPORTED;
    }
    ncclUniqueId id;
    memcpy(id.internal, id128, sizeof(id.internal));
    const ncclResult_t r = nccl().CommInitRank(&c->comm, nranks, id, rank);
    if (r != 0) {
      set_nccl_error(r, "ncclCommInitRank");
      delete c;
      return MAGPO_ERR_CUDA;
    }
  }
  *out = c;
  return MAGPO_OK;
}

int magpo_comm_destroy(MagpoComm* c) {
  if (!c) return MAGPO_OK;
  if (c->comm) nccl().CommDestroy(c->comm);
  delete c;
  return MAGPO_OK;
}

int magpo_comm_allreduce_sum(MagpoComm* c, magpo_stream_t s, float* buf, int64_t n) { return comm_allreduce(c, as_stream(s), buf, n, /*ncclSum*/ 0); }
int magpo_comm_allreduce_max(MagpoComm* c, magpo_stream_t s, float* buf, int64_t n) { return comm_allreduce(c, as_stream(s), buf, n, /*ncclMax*/ 2); }

// The communicator magpo_minibatch_grads reduces through when asked to (reduce_grads != 0); NULL detaches it.
int magpo_context_set_comm(MagpoContext* ctx, MagpoComm* c) {
  if (!ctx) return MAGPO_ERR_ARG;
  ctx->comm = c;
  return MAGPO_OK;
}

}  // extern "C"
