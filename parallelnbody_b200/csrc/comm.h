// NCCL plumbing for the multi-GPU paths: one process per GPU, communicator bootstrapped from a 128-byte
// unique id that the launcher broadcasts (torch.distributed in bench.py/tests; any out-of-band channel works).
// libnccl.so.2 is resolved at run time with dlopen, so a single-GPU user never needs NCCL and a process that
// already loaded torch's bundled NCCL shares that copy.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace nbody {

class Comm {
 public:
  static int unique_id(uint8_t out128[128]);
  // Collective: every rank calls with the same id. device must be current.
  static int create(Comm** out, const uint8_t id128[128], int rank, int world);
  ~Comm();
  int rank() const { return rank_; }
  int world() const { return world_; }
  // In-place all-gather of `count` floats per rank inside buf (rank r's slot at buf + r*count).
  int all_gather_f32_inplace(float* buf, size_t count, cudaStream_t s);
  int all_gather_bytes(const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s);
  int all_reduce_f64_sum(double* buf, size_t count, cudaStream_t s);
  int all_reduce_u32_max(uint32_t* buf, size_t count, cudaStream_t s);
  int all_reduce_u32_min(uint32_t* buf, size_t count, cudaStream_t s);
  int all_reduce_i64_sum(int64_t* buf, size_t count, cudaStream_t s);
  // Variable all-to-all of bytes: send_off/recv_off and counts are per peer, in bytes (grouped send/recv).
  int all_to_all_v(const void* send, const size_t* send_bytes, const size_t* send_off, void* recv,
                   const size_t* recv_bytes, const size_t* recv_off, cudaStream_t s);

 private:
  Comm() {}
  void* comm_ = nullptr;
  int rank_ = 0, world_ = 1;
};

}  // namespace nbody
