// Development probe (not part of the product library): how fast can the walk's EVALUATION half run on its own?
// Each warp holds 32 x B bodies in registers and evaluates a list of source entries from shared memory (SoA ring, one
// broadcast LDS.128 per four entries per component), exactly the inner loop of bh_walk_group_kernel - no traversal.
// Variants: B bodies per lane, warps per SM (register cap), one shared list or one list per body slot ("own").
// Build + run on the GPU box: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/eval_probe tools/eval_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../parallelnbody_b200/csrc/direct_kernels.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
using namespace nbody;

constexpr int kList = 128;   // entries per list

template <int B, bool OWN, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) eval_kernel(const float4* __restrict__ src, float4* __restrict__ out, const int iters, const float eps2) {
  extern __shared__ __align__(16) float smem[];
  constexpr int LISTS = OWN ? B : 1;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* ring = smem + (size_t)w * LISTS * 4 * kList;
  for (int l = 0; l < LISTS; l++)
    for (int j = lane; j < kList; j += 32) {
      const float4 s = src[(blockIdx.x * 131 + w * 17 + l * 7 + j) & 4095];
      ring[l * 4 * kList + j] = s.x; ring[l * 4 * kList + kList + j] = s.y; ring[l * 4 * kList + 2 * kList + j] = s.z; ring[l * 4 * kList + 3 * kList + j] = s.w;
    }
  __syncwarp();
  float2 nx[B], ny[B], nz[B], ax[B], ay[B], az[B];
#pragma unroll
  for (int k = 0; k < B; k++) {
    const float4 p = src[(blockIdx.x * 977 + threadIdx.x + 32 * k) & 4095];
    nx[k] = f2(-p.x, -p.x); ny[k] = f2(-p.y, -p.y); nz[k] = f2(-p.z, -p.z);
    ax[k] = f2(0.f, 0.f); ay[k] = f2(0.f, 0.f); az[k] = f2(0.f, 0.f);
  }
  const float2 e2 = f2(eps2, eps2);
  for (int it = 0; it < iters; it++) {
    const int head = (it & 1) * 64;   // 64-entry flushes, as the walk
#pragma unroll 2
    for (int j = 0; j < 64; j += 4) {
      if (!OWN) {
        const float4 X = *reinterpret_cast<const float4*>(ring + head + j);
        const float4 Y = *reinterpret_cast<const float4*>(ring + kList + head + j);
        const float4 Z = *reinterpret_cast<const float4*>(ring + 2 * kList + head + j);
        const float4 M = *reinterpret_cast<const float4*>(ring + 3 * kList + head + j);
#pragma unroll
        for (int k = 0; k < B; k++) {
          interact2<false>(f2(X.x, X.y), f2(Y.x, Y.y), f2(Z.x, Z.y), f2(M.x, M.y), nx[k], ny[k], nz[k], e2, ax[k], ay[k], az[k]);
          interact2<false>(f2(X.z, X.w), f2(Y.z, Y.w), f2(Z.z, Z.w), f2(M.z, M.w), nx[k], ny[k], nz[k], e2, ax[k], ay[k], az[k]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < B; k++) {
          const float* r = ring + k * 4 * kList;
          const float4 X = *reinterpret_cast<const float4*>(r + head + j);
          const float4 Y = *reinterpret_cast<const float4*>(r + kList + head + j);
          const float4 Z = *reinterpret_cast<const float4*>(r + 2 * kList + head + j);
          const float4 M = *reinterpret_cast<const float4*>(r + 3 * kList + head + j);
          interact2<false>(f2(X.x, X.y), f2(Y.x, Y.y), f2(Z.x, Z.y), f2(M.x, M.y), nx[k], ny[k], nz[k], e2, ax[k], ay[k], az[k]);
          interact2<false>(f2(X.z, X.w), f2(Y.z, Y.w), f2(Z.z, Z.w), f2(M.z, M.w), nx[k], ny[k], nz[k], e2, ax[k], ay[k], az[k]);
        }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < B; k++) s += ax[k].x + ax[k].y + ay[k].x + ay[k].y + az[k].x + az[k].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = make_float4(s, 0, 0, 0);
}

template <int B, bool OWN, int THREADS, int MINB>
void run(const float4* src, float4* out, int sms, const char* what) {
  auto k = eval_kernel<B, OWN, THREADS, MINB>;
  const size_t smem = (size_t)(THREADS / 32) * (OWN ? B : 1) * 4 * kList * sizeof(float);
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, THREADS, smem));
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k));
  const int grid = sms * per_sm, iters = 4000;
  k<<<grid, THREADS, smem>>>(src, out, 200, 1e-4f);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  k<<<grid, THREADS, smem>>>(src, out, iters, 1e-4f);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  const double inter = (double)grid * THREADS * B * 64.0 * iters;
  printf("%-34s B=%d regs=%3d warps/SM=%2d  %.3e interactions/s  (%.1f TFLOP/s at 20)\n", what, B, fa.numRegs, per_sm * THREADS / 32, inter / (ms * 1e-3), inter * 20 / (ms * 1e-3) / 1e12);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  float4* src; float4* out;
  CK(cudaMalloc(&src, 4096 * sizeof(float4))); CK(cudaMalloc(&out, (size_t)sms * 64 * 1024 * sizeof(float4)));
  float4 h[4096];
  srand(1);
  for (int i = 0; i < 4096; i++) h[i] = make_float4(rand() / (float)RAND_MAX, rand() / (float)RAND_MAX, rand() / (float)RAND_MAX, 1e-3f);
  CK(cudaMemcpy(src, h, sizeof(h), cudaMemcpyHostToDevice));
  run<1, false, 256, 3>(src, out, sms, "shared list, 80-reg cap");
  run<1, false, 256, 4>(src, out, sms, "shared list, 64-reg cap");
  run<1, false, 256, 6>(src, out, sms, "shared list, 40-reg cap");
  run<2, false, 256, 2>(src, out, sms, "shared list, 128-reg cap");
  run<2, false, 256, 3>(src, out, sms, "shared list, 80-reg cap");
  run<2, false, 256, 4>(src, out, sms, "shared list, 64-reg cap");
  run<4, false, 256, 1>(src, out, sms, "shared list, 255-reg cap");
  run<4, false, 256, 2>(src, out, sms, "shared list, 128-reg cap");
  run<4, false, 128, 5>(src, out, sms, "shared list, 96-reg cap");
  run<8, false, 256, 1>(src, out, sms, "shared list, 255-reg cap (K1-like)");
  run<8, false, 128, 3>(src, out, sms, "shared list, 168-reg cap");
  run<4, true, 256, 2>(src, out, sms, "own list per slot, 128-reg cap");
  run<4, true, 128, 5>(src, out, sms, "own list per slot, 96-reg cap");
  run<8, true, 256, 1>(src, out, sms, "own list per slot, 255-reg cap");
  run<8, true, 128, 3>(src, out, sms, "own list per slot, 168-reg cap");
  return 0;
}
