// Development microbenchmarks (not part of the product library): FP32 issue-rate probes and the K1 inner-loop
// variants, run once on a B200 to choose the shipping configuration. Build: see tools/Makefile.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../parallelnbody_b200/csrc/direct_kernels.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

using namespace nbody;

// ---- FP32 pipe probes -------------------------------------------------------------------------------
template <int ILP>
__global__ void __launch_bounds__(256) ffma_shared_ops(float* out, float a, float b, int iters, long long* cyc) {
  float acc[ILP];
#pragma unroll
  for (int k = 0; k < ILP; k++) acc[k] = threadIdx.x * 1e-3f + k;
  long long c0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int k = 0; k < ILP; k++) acc[k] = fmaf(acc[k], a, b);
  }
  long long c1 = clock64();
  float s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

// three distinct register operands per FFMA (acc = x*y + acc, x and y rotate)
template <int ILP>
__global__ void __launch_bounds__(256) ffma_3reg(float* out, const float* in, int iters, long long* cyc) {
  float acc[ILP], x[ILP], y[ILP];
#pragma unroll
  for (int k = 0; k < ILP; k++) { acc[k] = in[k]; x[k] = in[ILP + k] + threadIdx.x; y[k] = in[2 * ILP + k]; }
  long long c0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int k = 0; k < ILP; k++) acc[k] = fmaf(x[k], y[(k + u) % ILP], acc[k]);
  }
  long long c1 = clock64();
  float s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

template <int ILP>
__global__ void __launch_bounds__(256) ffma2_probe(float* out, const float* in, int iters, long long* cyc) {
  float2 acc[ILP], x[ILP], y[ILP];
#pragma unroll
  for (int k = 0; k < ILP; k++) {
    acc[k] = make_float2(in[k], in[k] + 1.f);
    x[k] = make_float2(in[ILP + k] + threadIdx.x, in[ILP + k]);
    y[k] = make_float2(in[2 * ILP + k], in[2 * ILP + k] * 0.5f);
  }
  long long c0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int k = 0; k < ILP; k++) acc[k] = __ffma2_rn(x[k], y[(k + u) % ILP], acc[k]);
  }
  long long c1 = clock64();
  float s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += acc[k].x + acc[k].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

// mixed: per group 12 FFMA + 1 MUFU.RSQ (the interaction's instruction mix), independent chains
template <int ILP>
__global__ void __launch_bounds__(256) mix_probe(float* out, const float* in, int iters, long long* cyc) {
  float acc[ILP], x[ILP], r[ILP];
#pragma unroll
  for (int k = 0; k < ILP; k++) { acc[k] = in[k]; x[k] = in[ILP + k] + threadIdx.x; r[k] = in[2 * ILP + k] + 1.f; }
  long long c0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) {
      r[k] = rsqrtf(r[k]) + 1.0f;  // 1 MUFU + 1 FADD
#pragma unroll
      for (int u = 0; u < 11; u++) acc[k] = fmaf(x[k], r[(k + u) % ILP], acc[k]);
    }
  }
  long long c1 = clock64();
  float s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += acc[k] + r[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

static float* d_out; static float* d_in; static long long* d_cyc;

template <class F>
static void run_probe(const char* name, F launch, int blocks, int threads, int iters, double ops_per_thread_iter, int sms) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(iters / 4);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    CK(cudaEventRecord(e0)); launch(iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = fminf(best, ms);
  }
  std::vector<long long> cyc(blocks);
  CK(cudaMemcpy(cyc.data(), d_cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost));
  double cavg = 0; for (auto c : cyc) cavg += (double)c; cavg /= blocks;
  double total_ops = ops_per_thread_iter * (double)iters * blocks * threads;
  double blocks_per_sm = (double)blocks / sms;
  // per-SM lane-ops per clock, from the in-kernel cycle counter (all resident CTAs run concurrently)
  double ops_per_clk_sm = ops_per_thread_iter * iters * threads * blocks_per_sm / cavg;
  printf("PROBE %-28s ms=%8.3f  lane-ops/clk/SM=%7.2f  Gops/s=%9.1f  eff_MHz=%7.1f\n", name, best, ops_per_clk_sm,
         total_ops / best * 1e-6, cavg / best * 1e-3);
}

// ---- direct-sum variants ----------------------------------------------------------------------------
__global__ void reduce_partials(const float4* partial, int jsplit, int n_pad, int n, float4* acc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 a = make_float4(0, 0, 0, 0);
  for (int s = 0; s < jsplit; s++) { float4 p = partial[(size_t)s * n_pad + i]; a.x += p.x; a.y += p.y; a.z += p.z; }
  acc[i] = a;
}

struct Variant { const char* name; int I; void (*launch)(dim3, const float4*, int, const float4*, int, float, float4*, int); };

template <int I, bool E0, int MINB> void launch_scalar(dim3 g, const float4* s, int chunk, const float4* t, int nt, float e2, float4* p, int npad) {
  direct_scalar_kernel<I, E0, MINB><<<g, kDirectTPB>>>(s, chunk, t, nt, e2, p, npad);
}
template <int I, bool E0, int MINB> void launch_packed(dim3 g, const float4* s, int chunk, const float4* t, int nt, float e2, float4* p, int npad) {
  direct_packed_kernel<I, E0, MINB><<<g, kDirectTPB>>>(s, chunk, t, nt, e2, p, npad);
}

// ---- (I, TPB) sweep of the packed kernel: more warps per SM vs more targets per thread
template <int I, int TPB, int MINB> static float time_packed(const float4* src, int n, int jsplit, float4* partial) {
  const int itile = TPB * I, nit = (n + itile - 1) / itile, chunk = n / jsplit;
  dim3 g(nit, jsplit);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  direct_packed_kernel<I, false, MINB, TPB><<<g, TPB>>>(src, chunk, src, n, 1e-4f, partial, n);
  CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    CK(cudaEventRecord(e0));
    direct_packed_kernel<I, false, MINB, TPB><<<g, TPB>>>(src, chunk, src, n, 1e-4f, partial, n);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = fminf(best, ms);
  }
  double ips = (double)n * n / (best * 1e-3);
  printf("SWEEP packed I=%d TPB=%d minb=%d jsplit=%d grid=%dx%d  ms=%8.3f  inter/s=%.4e  TFLOP/s(20)=%6.2f\n", I, TPB, MINB, jsplit, nit, jsplit,
         best, ips, ips * 20e-12);
  return best;
}

static int sweep_main() {
  const int n = 7680 * 48;   // divisible by 8 * {256, 320, 384, 512}
  std::vector<float4> h(n);
  srand(1234);
  for (int i = 0; i < n; i++) { h[i].x = 2.f * rand() / RAND_MAX - 1.f; h[i].y = 2.f * rand() / RAND_MAX - 1.f; h[i].z = 2.f * rand() / RAND_MAX - 1.f; h[i].w = 1e-4f / n; }
  float4 *d_src, *d_partial;
  CK(cudaMalloc(&d_src, (size_t)n * sizeof(float4)));
  CK(cudaMalloc(&d_partial, (size_t)8 * n * sizeof(float4)));
  CK(cudaMemcpy(d_src, h.data(), (size_t)n * sizeof(float4), cudaMemcpyHostToDevice));
  time_packed<8, 256, 1>(d_src, n, 8, d_partial);
  time_packed<6, 320, 1>(d_src, n, 8, d_partial);
  time_packed<5, 384, 1>(d_src, n, 8, d_partial);
  time_packed<6, 384, 1>(d_src, n, 8, d_partial);
  time_packed<4, 512, 1>(d_src, n, 8, d_partial);
  time_packed<4, 384, 1>(d_src, n, 8, d_partial);
  time_packed<4, 256, 2>(d_src, n, 8, d_partial);
  time_packed<3, 320, 2>(d_src, n, 8, d_partial);
  time_packed<8, 256, 1>(d_src, n, 8, d_partial);
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "sweep") return sweep_main();
  int n = argc > 1 ? atoi(argv[1]) : 262144;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("device %s  SMs=%d  clock=%d kHz\n", prop.name, sms, prop.clockRate);
  CK(cudaMalloc(&d_out, 148 * 16 * 256 * sizeof(float)));
  CK(cudaMalloc(&d_cyc, 148 * 16 * sizeof(long long)));
  std::vector<float> hin(256); for (int i = 0; i < 256; i++) hin[i] = 0.001f * (i + 1);
  CK(cudaMalloc(&d_in, 256 * sizeof(float))); CK(cudaMemcpy(d_in, hin.data(), 256 * sizeof(float), cudaMemcpyHostToDevice));

  const int iters = 20000;
  for (int bps : {2, 4}) {
    int blocks = sms * bps;
    printf("-- %d CTAs/SM x 256 threads\n", bps);
    run_probe("ffma shared-ops ILP8", [&](int it) { ffma_shared_ops<8><<<blocks, 256>>>(d_out, 1.0001f, 0.5f, it, d_cyc); }, blocks, 256, iters, 64, sms);
    run_probe("ffma 3reg ILP8", [&](int it) { ffma_3reg<8><<<blocks, 256>>>(d_out, d_in, it, d_cyc); }, blocks, 256, iters, 64, sms);
    run_probe("ffma 3reg ILP16", [&](int it) { ffma_3reg<16><<<blocks, 256>>>(d_out, d_in, it, d_cyc); }, blocks, 256, iters, 128, sms);
    run_probe("ffma2 ILP8 (x2 lanes)", [&](int it) { ffma2_probe<8><<<blocks, 256>>>(d_out, d_in, it, d_cyc); }, blocks, 256, iters, 128, sms);
    run_probe("ffma2 ILP4 (x2 lanes)", [&](int it) { ffma2_probe<4><<<blocks, 256>>>(d_out, d_in, it, d_cyc); }, blocks, 256, iters, 64, sms);
    run_probe("mix 11ffma+fadd+mufu ILP8", [&](int it) { mix_probe<8><<<blocks, 256>>>(d_out, d_in, it, d_cyc); }, blocks, 256, iters, 8 * 13, sms);
  }

  // ---- direct-sum variants
  std::vector<float4> h(n);
  srand(1234);
  for (int i = 0; i < n; i++) {
    h[i].x = 2.f * rand() / RAND_MAX - 1.f; h[i].y = 2.f * rand() / RAND_MAX - 1.f; h[i].z = 2.f * rand() / RAND_MAX - 1.f;
    h[i].w = 1e-4f / n;
  }
  float4 *d_src, *d_partial, *d_acc, *d_ref;
  const int max_split = 64;
  CK(cudaMalloc(&d_src, (size_t)n * sizeof(float4)));
  CK(cudaMalloc(&d_partial, (size_t)max_split * n * sizeof(float4)));
  CK(cudaMalloc(&d_acc, (size_t)n * sizeof(float4)));
  CK(cudaMalloc(&d_ref, (size_t)n * sizeof(float4)));
  CK(cudaMemcpy(d_src, h.data(), (size_t)n * sizeof(float4), cudaMemcpyHostToDevice));
  const float eps2 = 1e-4f;

  Variant vars[] = {
      {"scalar I=2 minb=3", 2, launch_scalar<2, false, 3>},
      {"scalar I=4 minb=2", 4, launch_scalar<4, false, 2>},
      {"scalar I=4 minb=3", 4, launch_scalar<4, false, 3>},
      {"scalar I=8 minb=2", 8, launch_scalar<8, false, 2>},
      {"scalar I=8 minb=1", 8, launch_scalar<8, false, 1>},
      {"scalar I=4 eps0   ", 4, launch_scalar<4, true, 2>},
      {"packed I=2 minb=3", 2, launch_packed<2, false, 3>},
      {"packed I=4 minb=2", 4, launch_packed<4, false, 2>},
      {"packed I=4 minb=3", 4, launch_packed<4, false, 3>},
      {"packed I=8 minb=1", 8, launch_packed<8, false, 1>},
      {"packed I=8 minb=2", 8, launch_packed<8, false, 2>},
      {"packed I=4 eps0   ", 4, launch_packed<4, true, 2>},
  };
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  bool have_ref = false;
  std::vector<float4> href(n), hacc(n);
  for (auto& v : vars) {
    int itile = kDirectTPB * v.I;
    int nit = (n + itile - 1) / itile;
    for (int jsplit : {1, 8, 32}) {
      if (n % (jsplit * kDirectTJ)) continue;
      if ((long long)nit * jsplit < sms) continue;
      int chunk = n / jsplit;
      dim3 g(nit, jsplit);
      v.launch(g, d_src, chunk, d_src, n, eps2, d_partial, n);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
      float best = 1e30f;
      for (int r = 0; r < 3; r++) {
        CK(cudaEventRecord(e0));
        v.launch(g, d_src, chunk, d_src, n, eps2, d_partial, n);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = fminf(best, ms);
      }
      reduce_partials<<<(n + 255) / 256, 256>>>(d_partial, jsplit, n, n, d_acc);
      CK(cudaMemcpy(hacc.data(), d_acc, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost));
      double err = 0, nrm = 0;
      if (!have_ref) { href = hacc; have_ref = true; }
      for (int i = 0; i < n; i++) {
        double dx = hacc[i].x - href[i].x, dy = hacc[i].y - href[i].y, dz = hacc[i].z - href[i].z;
        err += dx * dx + dy * dy + dz * dz;
        nrm += (double)href[i].x * href[i].x + (double)href[i].y * href[i].y + (double)href[i].z * href[i].z;
      }
      double ips = (double)n * n / (best * 1e-3);
      printf("DIRECT %-20s jsplit=%2d grid=%5dx%-2d ms=%9.3f  inter/s=%.4e  TFLOP/s(20)=%6.2f  relL2_vs_first=%.2e\n", v.name, jsplit,
             nit, jsplit, best, ips, ips * 20e-12, sqrt(err / nrm));
    }
  }
  return 0;
}
