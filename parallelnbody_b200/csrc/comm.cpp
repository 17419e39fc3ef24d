#include "comm.h"

#include <dlfcn.h>
#include <nccl.h>

#include <mutex>
#include <string>

#include "common.cuh"

namespace nbody {
namespace {

struct Api {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  std::string err;
};

Api* api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (a.lib) break;
    }
    if (!a.lib) { a.err = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
#define L(field, sym)                                                       \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, sym));         \
  if (!a.field) { a.err = std::string("libnccl lacks ") + sym; return; }
    L(GetUniqueId, "ncclGetUniqueId")
    L(CommInitRank, "ncclCommInitRank")
    L(CommDestroy, "ncclCommDestroy")
    L(GetErrorString, "ncclGetErrorString")
    L(AllGather, "ncclAllGather")
    L(AllReduce, "ncclAllReduce")
    L(Send, "ncclSend")
    L(Recv, "ncclRecv")
    L(GroupStart, "ncclGroupStart")
    L(GroupEnd, "ncclGroupEnd")
#undef L
  });
  return &a;
}

int fail(ncclResult_t r, const char* what) {
  Api* a = api();
  set_error(std::string("NCCL: ") + what + ": " + (a->GetErrorString ? a->GetErrorString(r) : "?"));
  return -3;
}
#define NB_NCCL(call, what)                   \
  do {                                        \
    ncclResult_t _r = (call);                 \
    if (_r != ncclSuccess) return fail(_r, what); \
  } while (0)

int ready() {
  Api* a = api();
  if (!a->err.empty() || !a->lib) { set_error("NCCL: " + a->err); return -3; }
  return 0;
}

}  // namespace

int Comm::unique_id(uint8_t out128[128]) {
  NB_TRY(ready());
  ncclUniqueId id;
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  NB_NCCL(api()->GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(out128, &id, 128);
  return 0;
}

int Comm::create(Comm** out, const uint8_t id128[128], int rank, int world) {
  NB_TRY(ready());
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t c;
  NB_NCCL(api()->CommInitRank(&c, world, id, rank), "ncclCommInitRank");
  Comm* k = new Comm();
  k->comm_ = c; k->rank_ = rank; k->world_ = world;
  *out = k;
  return 0;
}

Comm::~Comm() {
  if (comm_) api()->CommDestroy((ncclComm_t)comm_);
}

int Comm::all_gather_f32_inplace(float* buf, size_t count, cudaStream_t s) {
  NB_NCCL(api()->AllGather(buf + (size_t)rank_ * count, buf, count, ncclFloat32, (ncclComm_t)comm_, s), "ncclAllGather");
  return 0;
}
int Comm::all_gather_bytes(const void* send, void* recv, size_t bytes, cudaStream_t s) {
  NB_NCCL(api()->AllGather(send, recv, bytes, ncclInt8, (ncclComm_t)comm_, s), "ncclAllGather");
  return 0;
}
int Comm::all_reduce_f64_sum(double* buf, size_t count, cudaStream_t s) {
  NB_NCCL(api()->AllReduce(buf, buf, count, ncclFloat64, ncclSum, (ncclComm_t)comm_, s), "ncclAllReduce");
  return 0;
}
int Comm::all_reduce_u32_max(uint32_t* buf, size_t count, cudaStream_t s) {
  NB_NCCL(api()->AllReduce(buf, buf, count, ncclUint32, ncclMax, (ncclComm_t)comm_, s), "ncclAllReduce");
  return 0;
}
int Comm::all_reduce_u32_min(uint32_t* buf, size_t count, cudaStream_t s) {
  NB_NCCL(api()->AllReduce(buf, buf, count, ncclUint32, ncclMin, (ncclComm_t)comm_, s), "ncclAllReduce");
  return 0;
}
int Comm::all_reduce_i64_sum(int64_t* buf, size_t count, cudaStream_t s) {
  NB_NCCL(api()->AllReduce(buf, buf, count, ncclInt64, ncclSum, (ncclComm_t)comm_, s), "ncclAllReduce");
  return 0;
}
int Comm::all_to_all_v(const void* send, const size_t* send_bytes, const size_t* send_off, void* recv,
                       const size_t* recv_bytes, const size_t* recv_off, cudaStream_t s) {
  NB_NCCL(api()->GroupStart(), "ncclGroupStart");
  for (int p = 0; p < world_; p++) {
    if (send_bytes[p]) NB_NCCL(api()->Send((const char*)send + send_off[p], send_bytes[p], ncclInt8, p, (ncclComm_t)comm_, s), "ncclSend");
    if (recv_bytes[p]) NB_NCCL(api()->Recv((char*)recv + recv_off[p], recv_bytes[p], ncclInt8, p, (ncclComm_t)comm_, s), "ncclRecv");
  }
  NB_NCCL(api()->GroupEnd(), "ncclGroupEnd");
  return 0;
}

}  // namespace nbody
