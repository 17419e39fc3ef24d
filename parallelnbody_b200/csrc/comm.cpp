#include "comm.h"

#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace nbody {
namespace {

// ===================================================================================================== NCCL back end
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

class NcclComm : public Comm {
 public:
  NcclComm(ncclComm_t c, int rank, int world) : comm_(c) { rank_ = rank; world_ = world; }
  ~NcclComm() override { if (comm_) api()->CommDestroy(comm_); }
  const char* backend() const override { return "nccl"; }
  int all_gather_f32_inplace(float* buf, size_t count, cudaStream_t s) override {
    NB_NCCL(api()->AllGather(buf + (size_t)rank_ * count, buf, count, ncclFloat32, comm_, s), "ncclAllGather");
    return 0;
  }
  int all_gather_bytes(const void* send, void* recv, size_t bytes, cudaStream_t s) override {
    NB_NCCL(api()->AllGather(send, recv, bytes, ncclInt8, comm_, s), "ncclAllGather");
    return 0;
  }
  int all_reduce_f64_sum(double* buf, size_t count, cudaStream_t s) override {
    NB_NCCL(api()->AllReduce(buf, buf, count, ncclFloat64, ncclSum, comm_, s), "ncclAllReduce");
    return 0;
  }
  int all_reduce_u32_max(uint32_t* buf, size_t count, cudaStream_t s) override {
    NB_NCCL(api()->AllReduce(buf, buf, count, ncclUint32, ncclMax, comm_, s), "ncclAllReduce");
    return 0;
  }
  int all_reduce_u32_min(uint32_t* buf, size_t count, cudaStream_t s) override {
    NB_NCCL(api()->AllReduce(buf, buf, count, ncclUint32, ncclMin, comm_, s), "ncclAllReduce");
    return 0;
  }
  int all_reduce_i64_sum(int64_t* buf, size_t count, cudaStream_t s) override {
    NB_NCCL(api()->AllReduce(buf, buf, count, ncclInt64, ncclSum, comm_, s), "ncclAllReduce");
    return 0;
  }
  int all_to_all_v(const void* send, const size_t* send_bytes, const size_t* send_off, void* recv, const size_t* recv_bytes,
                   const size_t* recv_off, cudaStream_t s) override {
    // what a rank keeps for itself is a plain device copy (HBM speed), not a send / receive pair through NCCL's channels
    static const bool self_nccl = getenv("NBODY_A2A_SELF_NCCL") != nullptr;   // development knob: the own share through NCCL as well
    const bool self_done = !self_nccl && send_bytes[rank_] == recv_bytes[rank_];
    if (self_done && send_bytes[rank_] && (const char*)send + send_off[rank_] != (char*)recv + recv_off[rank_])
      NB_CUDA(cudaMemcpyAsync((char*)recv + recv_off[rank_], (const char*)send + send_off[rank_], send_bytes[rank_], cudaMemcpyDeviceToDevice, s));
    NB_NCCL(api()->GroupStart(), "ncclGroupStart");
    // an error inside the group must still close it, or every later NCCL call of this thread joins the open group
    ncclResult_t bad = ncclSuccess;
    const char* where = "";
    for (int p = 0; p < world_ && bad == ncclSuccess; p++) {
      if (p == rank_ && self_done) continue;
      if (send_bytes[p]) { bad = api()->Send((const char*)send + send_off[p], send_bytes[p], ncclInt8, p, comm_, s); where = "ncclSend"; }
      if (bad == ncclSuccess && recv_bytes[p]) { bad = api()->Recv((char*)recv + recv_off[p], recv_bytes[p], ncclInt8, p, comm_, s); where = "ncclRecv"; }
    }
    const ncclResult_t end = api()->GroupEnd();
    if (bad != ncclSuccess) return fail(bad, where);
    NB_NCCL(end, "ncclGroupEnd");
    return 0;
  }

  int all_to_all_v_multi(int nbuf, const void* const* send, void* const* recv, const size_t* elem, const size_t* send_cnt,
                         const size_t* send_off, const size_t* recv_cnt, const size_t* recv_off, cudaStream_t s) override {
    static const bool self_nccl = getenv("NBODY_A2A_SELF_NCCL") != nullptr;
    const bool self_done = !self_nccl && send_cnt[rank_] == recv_cnt[rank_];
    if (self_done && send_cnt[rank_])
      for (int k = 0; k < nbuf; k++)
        NB_CUDA(cudaMemcpyAsync((char*)recv[k] + recv_off[rank_] * elem[k], (const char*)send[k] + send_off[rank_] * elem[k], send_cnt[rank_] * elem[k],
                                cudaMemcpyDeviceToDevice, s));
    NB_NCCL(api()->GroupStart(), "ncclGroupStart");
    ncclResult_t bad = ncclSuccess;
    const char* where = "";
    for (int k = 0; k < nbuf && bad == ncclSuccess; k++)
      for (int p = 0; p < world_ && bad == ncclSuccess; p++) {
        if (p == rank_ && self_done) continue;
        if (send_cnt[p]) { bad = api()->Send((const char*)send[k] + send_off[p] * elem[k], send_cnt[p] * elem[k], ncclInt8, p, comm_, s); where = "ncclSend"; }
        if (bad == ncclSuccess && recv_cnt[p]) { bad = api()->Recv((char*)recv[k] + recv_off[p] * elem[k], recv_cnt[p] * elem[k], ncclInt8, p, comm_, s); where = "ncclRecv"; }
      }
    const ncclResult_t end = api()->GroupEnd();
    if (bad != ncclSuccess) return fail(bad, where);
    NB_NCCL(end, "ncclGroupEnd");
    return 0;
  }

 private:
  ncclComm_t comm_ = nullptr;
};

// ================================================================================================ loop-back back end
constexpr char kLoopMagic[8] = {'N', 'B', 'L', 'O', 'O', 'P', 'B', 'K'};
constexpr int kLoopMaxWorld = 64;

struct LoopSlot {
  const char* send = nullptr;
  size_t send_bytes[kLoopMaxWorld], send_off[kLoopMaxWorld];
  cudaEvent_t ready = nullptr, done = nullptr;
};

struct LoopGroup {
  uint64_t id = 0;
  int world = 0;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0, members = 0;
  uint64_t generation = 0;
  bool broken = false;
  LoopSlot slot[kLoopMaxWorld];
};

std::mutex g_loop_mu;
std::map<uint64_t, std::shared_ptr<LoopGroup>> g_loop_groups;

double loop_timeout_s() {
  const char* e = getenv("NBODY_LOOPBACK_TIMEOUT_S");
  const double v = e ? atof(e) : 0.0;
  return v > 0 ? v : 120.0;
}

// Host rendezvous of the group's threads. A rank that never arrives (it failed earlier) breaks the group after the
// timeout instead of hanging the others for ever.
int loop_barrier(LoopGroup* g) {
  std::unique_lock<std::mutex> lk(g->mu);
  if (g->broken) { set_error("loop-back comm: group is broken (a rank failed or timed out)"); return -3; }
  const uint64_t gen = g->generation;
  if (++g->arrived == g->world) {
    g->arrived = 0;
    g->generation++;
    g->cv.notify_all();
    return 0;
  }
  const auto deadline = std::chrono::steady_clock::now() + std::chrono::duration<double>(loop_timeout_s());
  while (g->generation == gen && !g->broken) {
    if (g->cv.wait_until(lk, deadline) == std::cv_status::timeout && g->generation == gen) {
      g->broken = true;
      g->cv.notify_all();
    }
  }
  if (g->generation == gen) { set_error("loop-back comm: timed out waiting for the other ranks (each rank needs its own host thread)"); return -3; }
  return 0;
}

template <class T, int OP>   // OP 0 sum, 1 max, 2 min; in = [world][count], summed in rank order on every rank
__global__ void loop_reduce_kernel(const T* __restrict__ in, const int world, const int count, T* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  T v = in[i];
  for (int p = 1; p < world; p++) {
    const T x = in[(size_t)p * count + i];
    v = OP == 0 ? v + x : OP == 1 ? (x > v ? x : v) : (x < v ? x : v);
  }
  out[i] = v;
}

class LoopComm : public Comm {
 public:
  LoopComm(std::shared_ptr<LoopGroup> g, int rank, int world) : g_(std::move(g)) { rank_ = rank; world_ = world; }
  ~LoopComm() override {
    LoopSlot& me = g_->slot[rank_];
    if (me.ready) cudaEventDestroy(me.ready);
    if (me.done) cudaEventDestroy(me.done);
    me.ready = me.done = nullptr;
    cudaFree(scratch_);
    std::lock_guard<std::mutex> lk(g_loop_mu);
    bool last = false;
    { std::lock_guard<std::mutex> lk2(g_->mu); last = --g_->members == 0; }
    if (last) g_loop_groups.erase(g_->id);
  }
  int init() {
    LoopSlot& me = g_->slot[rank_];
    NB_CUDA(cudaEventCreateWithFlags(&me.ready, cudaEventDisableTiming));
    NB_CUDA(cudaEventCreateWithFlags(&me.done, cudaEventDisableTiming));
    return loop_barrier(g_.get());   // every rank's events exist before anyone communicates
  }
  const char* backend() const override { return "loopback"; }

  int all_to_all_v(const void* send, const size_t* send_bytes, const size_t* send_off, void* recv, const size_t* recv_bytes,
                   const size_t* recv_off, cudaStream_t s) override {
    LoopSlot& me = g_->slot[rank_];
    me.send = (const char*)send;
    for (int p = 0; p < world_; p++) { me.send_bytes[p] = send_bytes[p]; me.send_off[p] = send_off[p]; }
    NB_CUDA(cudaEventRecord(me.ready, s));                 // my send buffer is final once this fires
    NB_TRY(loop_barrier(g_.get()));
    for (int p = 0; p < world_; p++) {
      const LoopSlot& peer = g_->slot[p];
      if (peer.send_bytes[rank_] != recv_bytes[p]) {
        set_error("loop-back comm: rank " + std::to_string(p) + " sends " + std::to_string(peer.send_bytes[rank_]) + " bytes, rank " +
                  std::to_string(rank_) + " expects " + std::to_string(recv_bytes[p]));
        std::lock_guard<std::mutex> lk(g_->mu);
        g_->broken = true;
        g_->cv.notify_all();
        return -3;
      }
      if (!recv_bytes[p]) continue;
      const char* src = peer.send + peer.send_off[rank_];
      char* dst = (char*)recv + recv_off[p];
      if (src == dst) continue;                            // in-place all-gather: the own slot is already there
      if (p != rank_) NB_CUDA(cudaStreamWaitEvent(s, peer.ready, 0));
      NB_CUDA(cudaMemcpyAsync(dst, src, recv_bytes[p], cudaMemcpyDefault, s));
    }
    NB_CUDA(cudaEventRecord(me.done, s));                  // I have read everything I need from the peers
    NB_TRY(loop_barrier(g_.get()));
    for (int p = 0; p < world_; p++)                        // nothing after this call may touch my send buffer before
      if (p != rank_) NB_CUDA(cudaStreamWaitEvent(s, g_->slot[p].done, 0));   // every peer has copied out of it
    // a rank may destroy its communicator (and these events) right after its last collective: nobody leaves before
    // everybody has enqueued its waits
    return loop_barrier(g_.get());
  }
  int all_gather_bytes(const void* send, void* recv, size_t bytes, cudaStream_t s) override {
    size_t sb[kLoopMaxWorld], so[kLoopMaxWorld], rb[kLoopMaxWorld], ro[kLoopMaxWorld];
    for (int p = 0; p < world_; p++) { sb[p] = bytes; so[p] = 0; rb[p] = bytes; ro[p] = (size_t)p * bytes; }
    return all_to_all_v(send, sb, so, recv, rb, ro, s);
  }
  int all_gather_f32_inplace(float* buf, size_t count, cudaStream_t s) override {
    return all_gather_bytes(buf + (size_t)rank_ * count, buf, count * sizeof(float), s);
  }
  int all_reduce_f64_sum(double* buf, size_t count, cudaStream_t s) override { return reduce<double, 0>(buf, count, s); }
  int all_reduce_u32_max(uint32_t* buf, size_t count, cudaStream_t s) override { return reduce<uint32_t, 1>(buf, count, s); }
  int all_reduce_u32_min(uint32_t* buf, size_t count, cudaStream_t s) override { return reduce<uint32_t, 2>(buf, count, s); }
  int all_reduce_i64_sum(int64_t* buf, size_t count, cudaStream_t s) override { return reduce<int64_t, 0>(buf, count, s); }

 private:
  template <class T, int OP>
  int reduce(T* buf, size_t count, cudaStream_t s) {
    const size_t need = (size_t)world_ * count * sizeof(T);
    if (need > scratch_bytes_) {
      NB_CUDA(cudaStreamSynchronize(s));
      if (scratch_) NB_CUDA(cudaFree(scratch_));
      scratch_ = nullptr;
      NB_CUDA(cudaMalloc(&scratch_, std::max<size_t>(need, 4096)));
      scratch_bytes_ = std::max<size_t>(need, 4096);
    }
    NB_TRY(all_gather_bytes(buf, scratch_, count * sizeof(T), s));   // on return the peers are done reading buf
    loop_reduce_kernel<T, OP><<<(unsigned)((count + 127) / 128), 128, 0, s>>>((const T*)scratch_, world_, (int)count, buf);
    NB_CUDA(cudaGetLastError());
    return 0;
  }
  std::shared_ptr<LoopGroup> g_;
  void* scratch_ = nullptr;
  size_t scratch_bytes_ = 0;
};

}  // namespace

// Default: one exchange per buffer (the loop-back back end); NcclComm overrides it with a single group.
int Comm::all_to_all_v_multi(int nbuf, const void* const* send, void* const* recv, const size_t* elem, const size_t* send_cnt,
                             const size_t* send_off, const size_t* recv_cnt, const size_t* recv_off, cudaStream_t s) {
  size_t a[64], b[64], c[64], d[64];
  for (int k = 0; k < nbuf; k++) {
    for (int p = 0; p < world_; p++) { a[p] = send_cnt[p] * elem[k]; b[p] = send_off[p] * elem[k]; c[p] = recv_cnt[p] * elem[k]; d[p] = recv_off[p] * elem[k]; }
    NB_TRY(all_to_all_v(send[k], a, b, recv[k], c, d, s));
  }
  return 0;
}

int Comm::unique_id(uint8_t out128[128]) {
  NB_TRY(ready());
  ncclUniqueId id;
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  NB_NCCL(api()->GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(out128, &id, 128);
  return 0;
}

int Comm::loopback_id(uint8_t out128[128]) {
  static std::mutex mu;
  static uint64_t counter = 0;
  std::lock_guard<std::mutex> lk(mu);
  memset(out128, 0, 128);
  memcpy(out128, kLoopMagic, 8);
  const uint64_t id = ((uint64_t)getpid() << 32) | ++counter;
  memcpy(out128 + 8, &id, 8);
  return 0;
}

int Comm::create(Comm** out, const uint8_t id128[128], int rank, int world) {
  *out = nullptr;
  if (memcmp(id128, kLoopMagic, 8) == 0) {
    if (world > kLoopMaxWorld) { set_error("loop-back comm: at most 64 ranks"); return -1; }
    uint64_t id = 0;
    memcpy(&id, id128 + 8, 8);
    std::shared_ptr<LoopGroup> g;
    {
      std::lock_guard<std::mutex> lk(g_loop_mu);
      auto& slot = g_loop_groups[id];
      if (!slot) { slot = std::make_shared<LoopGroup>(); slot->id = id; slot->world = world; }
      g = slot;
      std::lock_guard<std::mutex> lk2(g->mu);
      if (g->world != world || g->members >= world) { set_error("loop-back comm: world size mismatch or too many ranks joined this id"); return -1; }
      g->members++;
    }
    LoopComm* c = new LoopComm(g, rank, world);
    const int rc = c->init();
    if (rc) { delete c; return rc; }
    *out = c;
    return 0;
  }
  NB_TRY(ready());
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t c;
  NB_NCCL(api()->CommInitRank(&c, world, id, rank), "ncclCommInitRank");
  *out = new NcclComm(c, rank, world);
  return 0;
}

}  // namespace nbody
