// K2 - fused (partial-sum reduce) + kick-drift integrator, K3 - cube-size reduction, layout conversion,
// the device-side CreateSpacePoints generator and the energy diagnostics. All HBM-bound, one pass each.
#pragma once
#include "common.cuh"

namespace nbody {

// ---- K2 -------------------------------------------------------------------------------------------
// Consumes K1's partial accelerations (partial[jsplit][n_pad], summed in split order so the result is
// deterministic), applies G, stores acc, and - when INTEGRATE - performs the reference's integrator
// (AOctreeSearch::Tick, /root/reference/Source/NBody/OctreeSearch.cpp:28-31):
//     Velocity += PhDeltaTime * Acceleration;   Position += PhDeltaTime * Velocity;   (kick, then drift)
// product then add with no FMA contraction, exactly the reference's fp32 statement order. The new position
// is written straight into this rank's slot of the gathered source array (the all-gather send buffer).
// Algorithmic traffic per body: 60 B of reference fields; this float4 layout moves
// 16*jsplit (partials) + 16 (acc w) + 32 (vel r/w) + 32 (posm r/w) bytes.
template <bool INTEGRATE>
__global__ void __launch_bounds__(256)
reduce_kick_drift_kernel(const float4* __restrict__ partial, const int jsplit, const int64_t n_pad, const int n,
                         const float G, const float dt, float4* __restrict__ posm, float4* __restrict__ vel,
                         float4* __restrict__ acc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float ax = 0.f, ay = 0.f, az = 0.f;
  for (int s = 0; s < jsplit; s++) {
    const float4 p = ld_stream(partial + (size_t)s * n_pad + i);
    ax += p.x; ay += p.y; az += p.z;
  }
  ax *= G; ay *= G; az *= G;
  st_stream(acc + i, make_float4(ax, ay, az, 0.f));
  if (INTEGRATE) {
    float4 v = vel[i];
    float4 p = posm[i];
    v.x = __fadd_rn(v.x, __fmul_rn(dt, ax));
    v.y = __fadd_rn(v.y, __fmul_rn(dt, ay));
    v.z = __fadd_rn(v.z, __fmul_rn(dt, az));
    p.x = __fadd_rn(p.x, __fmul_rn(dt, v.x));
    p.y = __fadd_rn(p.y, __fmul_rn(dt, v.y));
    p.z = __fadd_rn(p.z, __fmul_rn(dt, v.z));
    vel[i] = v;
    posm[i] = p;
  }
}

// Kick-drift from an already reduced acceleration array (Barnes-Hut path; G already applied).
__global__ void __launch_bounds__(256)
kick_drift_kernel(const int n, const float dt, float4* __restrict__ posm, float4* __restrict__ vel,
                  const float4* __restrict__ acc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 a = acc[i];
  float4 v = vel[i];
  float4 p = posm[i];
  v.x = __fadd_rn(v.x, __fmul_rn(dt, a.x));
  v.y = __fadd_rn(v.y, __fmul_rn(dt, a.y));
  v.z = __fadd_rn(v.z, __fmul_rn(dt, a.z));
  p.x = __fadd_rn(p.x, __fmul_rn(dt, v.x));
  p.y = __fadd_rn(p.y, __fmul_rn(dt, v.y));
  p.z = __fadd_rn(p.z, __fmul_rn(dt, v.z));
  vel[i] = v;
  posm[i] = p;
}

// ---- K3 -------------------------------------------------------------------------------------------
// AOctreeSearch::ComputeCubeSize (OctreeSearch.cpp:47-56): Size = max_i max(|x|,|y|,|z|) about the world origin.
// Non-negative floats order like their bit patterns, so the cross-CTA combine is an integer atomicMax.
// Also produces the tight bounding box (min/max per axis) used by the non-reference root mode; box6 holds
// order-preserving uint keys (see float_to_ordered) so the same atomics work for signed values.
// out[0] = absmax bits; out[1..3] = ordered min x,y,z ; out[4..6] = ordered max x,y,z.
// Initialise out = {0, ~0,~0,~0, 0,0,0} before the launch.
__global__ void __launch_bounds__(256)
cube_size_kernel(const float4* __restrict__ posm, const int n, uint32_t* __restrict__ out) {
  float amax = 0.f;
  float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = ld_stream(posm + i);
    amax = fmaxf(amax, fmaxf(fmaxf(fabsf(p.x), fabsf(p.y)), fabsf(p.z)));
    mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
    mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
#pragma unroll
    for (int k = 0; k < 3; k++) {
      mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
      mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
    }
  }
  __shared__ float s[8][7];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s[w][0] = amax; for (int k = 0; k < 3; k++) { s[w][1 + k] = mn[k]; s[w][4 + k] = mx[k]; } }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < 8; q++) {
      amax = fmaxf(amax, s[q][0]);
      for (int k = 0; k < 3; k++) { mn[k] = fminf(mn[k], s[q][1 + k]); mx[k] = fmaxf(mx[k], s[q][4 + k]); }
    }
    atomicMax(out, __float_as_uint(amax));
    for (int k = 0; k < 3; k++) {
      atomicMin(out + 1 + k, float_to_ordered(mn[k]));
      atomicMax(out + 4 + k, float_to_ordered(mx[k]));
    }
  }
}

// ---- layout conversion: FParticle AoS (40 B, OctreeSearch.h:9-18) <-> float4 SoA -------------------------
// ids == nullptr: body i of the SoA arrays is record (first + i); else record ids[i].
__global__ void __launch_bounds__(256)
aos_to_soa_kernel(const uint8_t* __restrict__ aos, const size_t stride, const int64_t first, const int n,
                  float4* __restrict__ posm, float4* __restrict__ vel, float4* __restrict__ acc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* r = reinterpret_cast<const float*>(aos + (size_t)(first + i) * stride);
  posm[i] = make_float4(r[1], r[2], r[3], r[0]);
  vel[i] = make_float4(r[4], r[5], r[6], 0.f);
  acc[i] = make_float4(r[7], r[8], r[9], 0.f);
}
// Writes local body i into record i (ids == nullptr) or record ids[i] of a compact staging array (n records of 40 B).
__global__ void __launch_bounds__(256)
soa_to_aos_kernel(const float4* __restrict__ posm, const float4* __restrict__ vel, const float4* __restrict__ acc,
                  const int n, const int32_t* __restrict__ ids, float* __restrict__ aos10) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = posm[i], v = vel[i], a = acc[i];
  float* r = aos10 + (size_t)(ids ? ids[i] : i) * 10;
  r[0] = p.w; r[1] = p.x; r[2] = p.y; r[3] = p.z; r[4] = v.x; r[5] = v.y; r[6] = v.z; r[7] = a.x; r[8] = a.y; r[9] = a.z;
}

// ---- CreateSpacePoints on device (OctreeSearch.cpp:58-72) ------------------------------------------------
__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__host__ __device__ inline float u01(uint64_t h) { return (float)(h >> 40) * (1.0f / 16777216.0f); }  // [0,1)

// Counter-based: body g draws from splitmix64(seed, g, k); identical on every rank and reproducible on the host.
__host__ __device__ inline void space_point(uint64_t seed, int64_t g, float size, float* posm4, float* vel4) {
  const uint64_t base = splitmix64(seed ^ splitmix64((uint64_t)g));
  uint64_t h[7];
  for (int k = 0; k < 7; k++) h[k] = splitmix64(base + (uint64_t)k);
  posm4[0] = (2.f * u01(h[0]) - 1.f) * size;            // RandPointInBox(+-(S, S, S/10)), cpp:61,64
  posm4[1] = (2.f * u01(h[1]) - 1.f) * size;
  posm4[2] = (2.f * u01(h[2]) - 1.f) * (size / 10.f);
  const float speed = 10.f * (float)(25 + (int)(h[3] % 26));   // 10*RandRange(25,50), cpp:65
  const float z = 2.f * u01(h[4]) - 1.f, phi = 6.28318530718f * u01(h[5]);   // VRand(): isotropic direction
  const float s = sqrtf(fmaxf(0.f, 1.f - z * z));
  vel4[0] = speed * s * cosf(phi); vel4[1] = speed * s * sinf(phi); vel4[2] = speed * z; vel4[3] = 0.f;
  posm4[3] = (float)(1 + (int)(h[6] % 5000));                  // RandRange(1,5000), cpp:66
  if (g == 0) {                                                // central body, cpp:68-70
    posm4[0] = posm4[1] = posm4[2] = 0.f; vel4[0] = vel4[1] = vel4[2] = 0.f; posm4[3] = 5000.f;
  }
}
__global__ void __launch_bounds__(256)
space_points_kernel(const uint64_t seed, const int64_t first, const int n, const float size, float4* __restrict__ posm,
                    float4* __restrict__ vel, float4* __restrict__ acc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float p[4], v[4];
  space_point(seed, first + i, size, p, v);
  posm[i] = make_float4(p[0], p[1], p[2], p[3]);
  vel[i] = make_float4(v[0], v[1], v[2], 0.f);
  acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---- energy diagnostics (fp64 accumulation) ----------------------------------------------------------
// out[0] += sum 1/2 m v^2 ; out[1] += -1/2 G sum_i m_i sum_{j != i} m_j / sqrt(r^2 + eps^2).
// Targets = this rank's bodies (global index first_global + i); sources = all n_src bodies.
__global__ void __launch_bounds__(256)
energy_kernel(const float4* __restrict__ src, const int n_src, const float4* __restrict__ tgt, const float4* __restrict__ vel,
              const int n_tgt, const int64_t first_global, const float G, const float eps2, double* __restrict__ out) {
  __shared__ float4 tile[256];
  __shared__ double red[2][8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n_tgt;
  const float4 p = tgt[live ? i : n_tgt - 1];
  const int64_t gi = first_global + i;
  double phi = 0.0;
  for (int base = 0; base < n_src; base += 256) {
    const int j = base + threadIdx.x;
    tile[threadIdx.x] = j < n_src ? src[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int lim = min(256, n_src - base);
    float part = 0.f;
    for (int k = 0; k < lim; k++) {
      const float4 s = tile[k];
      const float dx = s.x - p.x, dy = s.y - p.y, dz = s.z - p.z;
      const float r2 = dx * dx + dy * dy + dz * dz;
      const bool skip = (base + k == gi) || (r2 == 0.f && eps2 == 0.f);
      part += skip ? 0.f : s.w * rsqrtf(r2 + eps2);
    }
    phi += (double)part;
    __syncthreads();
  }
  double ke = 0.0, pe = 0.0;
  if (live) {
    const float4 v = vel[i];
    ke = 0.5 * (double)p.w * ((double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z);
    pe = -0.5 * (double)G * (double)p.w * phi;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ke += __shfl_xor_sync(0xffffffffu, ke, o);
    pe += __shfl_xor_sync(0xffffffffu, pe, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ke; red[1][threadIdx.x >> 5] = pe; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < 8; q++) { ke += red[0][q]; pe += red[1][q]; }
    atomicAdd(out, ke);
    atomicAdd(out + 1, pe);
  }
}

// out[ids[i]] = in[i]: undoes the Morton reordering for a read-back.
__global__ void __launch_bounds__(256)
scatter_float4_kernel(const float4* __restrict__ in, const int32_t* __restrict__ ids, const int n, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[ids[i]] = in[i];
}

// Domain-split Barnes-Hut: after a migration the rank's bodies sit in two stretches (the ones that stayed, the ones that
// arrived); a read-back first makes them contiguous again. b0/n0/b1: logical body i = b0 + i for i < n0, else b1 + (i - n0).
__global__ void __launch_bounds__(256)
compact_segments_kernel(const int b0, const int n0, const int b1, const int n, const float4* __restrict__ posm_in,
                        const float4* __restrict__ vel_in, const float4* __restrict__ acc_in, const int32_t* __restrict__ ids_in,
                        float4* __restrict__ posm_out, float4* __restrict__ vel_out, float4* __restrict__ acc_out, int32_t* __restrict__ ids_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = i < n0 ? b0 + i : b1 + (i - n0);
  posm_out[i] = posm_in[j]; vel_out[i] = vel_in[j]; acc_out[i] = acc_in[j]; ids_out[i] = ids_in[j];
}

// out[0] = min, out[1] = max of the order-preserving keys of the masses (posm.w); initialise out = {~0, 0}.
__global__ void __launch_bounds__(256)
mass_range_kernel(const float4* __restrict__ posm, const int64_t n, uint32_t* __restrict__ out) {
  uint32_t lo = ~0u, hi = 0u;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t k = float_to_ordered(ld_stream(posm + i).w);
    lo = min(lo, k); hi = max(hi, k);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(out, lo); atomicMax(out + 1, hi); }
}

__global__ void fill_float4_kernel(float4* p, const int64_t n, const float4 v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace nbody
