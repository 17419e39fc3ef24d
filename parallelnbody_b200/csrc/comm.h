// Collective plumbing for the multi-GPU paths. Two back ends behind one interface:
//   * NCCL   - one process per GPU, communicator bootstrapped from the 128-byte ncclUniqueId that the launcher
//              broadcasts (torch.distributed in bench.py / tests; any out-of-band channel works). libnccl.so.2 is
//              resolved at run time with dlopen, so a single-GPU user never needs NCCL and a process that already
//              loaded torch's bundled NCCL shares that copy.
//   * loop-back - the ranks are handles of ONE process (one host thread each), usually on one GPU: collectives are
//              device-to-device copies ordered by CUDA events, rendezvous on a host barrier. Selected by an id minted
//              with Comm::loopback_id. It exists so that the domain-split Barnes-Hut path (migration, locally-essential
//              tree exchange) runs its real code on a single-GPU box (tests, compute-sanitizer).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace nbody {

class Comm {
 public:
  static int unique_id(uint8_t out128[128]);     // NCCL id (rank 0)
  static int loopback_id(uint8_t out128[128]);   // id of a fresh in-process group
  // Collective: every rank calls with the same id. device must be current.
  static int create(Comm** out, const uint8_t id128[128], int rank, int world);
  virtual ~Comm() {}
  int rank() const { return rank_; }
  int world() const { return world_; }
  virtual const char* backend() const = 0;
  // In-place all-gather of `count` floats per rank inside buf (rank r's slot at buf + r*count).
  virtual int all_gather_f32_inplace(float* buf, size_t count, cudaStream_t s) = 0;
  // send may alias recv + rank * bytes_per_rank.
  virtual int all_gather_bytes(const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s) = 0;
  virtual int all_reduce_f64_sum(double* buf, size_t count, cudaStream_t s) = 0;
  virtual int all_reduce_u32_max(uint32_t* buf, size_t count, cudaStream_t s) = 0;
  virtual int all_reduce_u32_min(uint32_t* buf, size_t count, cudaStream_t s) = 0;
  virtual int all_reduce_i64_sum(int64_t* buf, size_t count, cudaStream_t s) = 0;
  // Variable all-to-all of bytes: send_off/recv_off and counts are per peer, in bytes. The send and receive regions
  // must not overlap.
  virtual int all_to_all_v(const void* send, const size_t* send_bytes, const size_t* send_off, void* recv,
                           const size_t* recv_bytes, const size_t* recv_off, cudaStream_t s) = 0;
  // Several all_to_all_v with the same per-peer ELEMENT counts / offsets in one go (bodies travel as positions,
  // velocities and ids): elem[k] bytes per element of buffer pair k. One NCCL group = one launch.
  virtual int all_to_all_v_multi(int nbuf, const void* const* send, void* const* recv, const size_t* elem, const size_t* send_cnt,
                                 const size_t* send_off, const size_t* recv_cnt, const size_t* recv_off, cudaStream_t s);

 protected:
  int rank_ = 0, world_ = 1;
};

}  // namespace nbody
