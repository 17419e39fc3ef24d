// Shared helpers for the sm_100a kernels and the host runtime (no torch, no Unreal types).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <string>

namespace nbody {

// Thread-local error string behind nbody_last_error().
void set_error(const std::string& msg);

struct Status {
  int code;
  explicit operator bool() const { return code == 0; }
};

#define NB_CUDA(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) {                                                                            \
      ::nbody::set_error(std::string("CUDA: ") + cudaGetErrorString(_e) + " at " + __FILE__ + ":" +     \
                         std::to_string(__LINE__) + " (" #expr ")");                                    \
      return (_e == cudaErrorMemoryAllocation) ? -4 : -2;                                               \
    }                                                                                                   \
  } while (0)

#define NB_TRY(expr)            \
  do {                          \
    int _s = (expr);            \
    if (_s != 0) return _s;     \
  } while (0)

constexpr int kNumSMsB200 = 148;

// SMs of the current device (148 on a full B200). Cooperative launches must not exceed what is co-resident, so their
// grids are sized from this, not from the constant.
inline int sm_count() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
    return kNumSMsB200;
  return sms;
}

__host__ __device__ inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Streaming 128-bit loads/stores for data touched once per kernel (keeps L1 for reused lines).
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(float4* p, const float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Order-preserving float <-> uint32 map (so min / max of signed floats can use integer atomics).
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float ordered_to_float(uint32_t k) {
  const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}


}  // namespace nbody
