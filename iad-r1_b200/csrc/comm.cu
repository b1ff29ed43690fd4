// The one data-path collective of the SC-GRPO step: the sum of the flat fp32 gradient over the
// data-parallel ranks, issued bucket by bucket on a caller-supplied stream so that it overlaps the tail of the backward.
//
// Replaces DeepSpeed ZeRO-3's reduce-scatter / all-gather traffic around every parameter use
// (ref: scripts/train/zero3.json:14-33; launcher ref: scripts/train/SC_GRPO/SC_GRPO_Qwen_Instruct_2_5_VL_3B.sh:40-63).
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 the process already carries through torch), so the library
// itself has no link-time dependency on it and loads on a box without NCCL; only these entry points need it.
#include "runtime.h"

#include <dlfcn.h>
#include <mutex>

namespace iadr1 {

// the slice of nccl.h this file uses (NCCL 2.x ABI: stable enum values and a 128-byte unique id)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { kNcclFloat32 = 7, kNcclBfloat16 = 9, kNcclSum = 0 };

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

static NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(h, "ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString;
  });
  return api;
}

static int nccl_fail(const char* what, ncclResult_t r) {
  return set_error("%s: %s", what, nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error");
}

}  // namespace iadr1

using namespace iadr1;

extern "C" {

int iadr1_comm_unique_id(void* out128) {
  if (!nccl().ok) return set_error("NCCL (libnccl.so.2) could not be loaded");
  ncclUniqueId id;
  ncclResult_t r = nccl().GetUniqueId(&id);
  if (r != 0) return nccl_fail("ncclGetUniqueId", r);
  memcpy(out128, id.internal, 128);
  return 0;
}

int iadr1_comm_create(const void* unique_id128, int rank, int world, void** comm_out) {
  if (!nccl().ok) return set_error("NCCL (libnccl.so.2) could not be loaded");
  if (!comm_out || world < 1 || rank < 0 || rank >= world) return set_error("comm_create: bad arguments");
  ncclUniqueId id;
  memcpy(id.internal, unique_id128, 128);
  ncclComm_t c = nullptr;
  ncclResult_t r = nccl().CommInitRank(&c, world, id, rank);
  if (r != 0) return nccl_fail("ncclCommInitRank", r);
  *comm_out = c;
  return 0;
}

int iadr1_comm_destroy(void* comm) {
  if (!comm) return 0;
  ncclResult_t r = nccl().CommDestroy(static_cast<ncclComm_t>(comm));
  return r == 0 ? 0 : nccl_fail("ncclCommDestroy", r);
}

// In-place SUM of grad[0 .. n) over the communicator, cut into buckets of `bucket_elems` (0 = one call) so that a caller
// that retires gradient ranges layer by layer can enqueue each range as soon as it is final. `stream` is the stream the
// collective is ordered on (the trainer uses a side stream that waits on an event of the compute stream).
int iadr1_grad_allreduce(void* comm, float* grad, long long n, long long bucket_elems, void* stream) {
  if (n <= 0) return 0;
  if (!comm) return set_error("grad_allreduce: null communicator");
  if (!nccl().ok) return set_error("NCCL (libnccl.so.2) could not be loaded");
  if (bucket_elems <= 0) bucket_elems = n;
  for (long long off = 0; off < n; off += bucket_elems) {
    const long long cnt = n - off < bucket_elems ? n - off : bucket_elems;
    ncclResult_t r = nccl().AllReduce(grad + off, grad + off, (size_t)cnt, kNcclFloat32, kNcclSum,
                                      static_cast<ncclComm_t>(comm), static_cast<cudaStream_t>(stream));
    if (r != 0) return nccl_fail("ncclAllReduce", r);
  }
  return 0;
}

}  // extern "C"
