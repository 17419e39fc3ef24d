// K1 - tiled all-pairs ("direct sum") force kernels for sm_100a.
//
// Reproduces the results of the reference's force law, Octree::ComputeForces with every source a one-body
// leaf (the Theta = 0 limit; /root/reference/Source/NBody/OctreeSearch.h:99-108):
//     a_i += G * m_j * (x_j - x_i) / d^3,   d = |x_j - x_i|,   pairs with d == 0 skipped (h:102)
// with optional Plummer softening d^2 -> d^2 + eps^2 (BASELINE north_star). G is applied by the consumer
// (the fused reduce + kick-drift kernel, integrate.cuh), not per interaction.
//
// Data layout: sources are float4 (x, y, z, m) in HBM, read coalesced, staged per CTA in shared memory.
// Work decomposition: grid = (i-tiles, j-splits). A CTA owns TPB*I targets (I per thread, in registers) and
// one contiguous chunk of sources; it writes one float4 partial acceleration per target into
// partial[jsplit][n_tgt_pad]. The consumer sums the splits in a fixed order, so results are deterministic
// (no float atomics). j-splitting is what keeps all 148 SMs busy and the tail small for any N.
//
// Two inner loops are provided:
//   * scalar  : FADD/FFMA/MUFU.RSQ/FMUL per interaction, sources broadcast from smem as LDS.128.
//   * packed  : Blackwell packed-fp32 (FADD2/FFMA2/FMUL2, PTX add/fma/mul.f32x2) over PAIRS OF SOURCES, sources
//               staged SoA in smem so one LDS.128 yields four x (or y, z, m) values. Each thread keeps two
//               partial accumulators per target and folds them at the end.
// Which one ships is decided by measurement (profiles/).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nbody {

constexpr int kDirectTPB = 256;  // threads per CTA
constexpr int kDirectTJ = 256;   // sources per smem tile (one per thread per load)

// Bare MUFU.RSQ. rsqrtf() without -ftz wraps the MUFU in a denormal-range test and two predicated FMULs
// (3 extra issue slots per interaction); r^2 + eps^2 is never denormal for eps > 0, and the eps == 0
// instantiation treats pairs closer than sqrt(FLT_MIN) ~ 1e-19 as coincident (skipped like d == 0).
__device__ __forceinline__ float rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kTinyR2 = 1.17549435e-38f;

// ---- scalar interaction ------------------------------------------------------------------------------
template <bool EPS0>
__device__ __forceinline__ void interact(const float4 s, const float xi, const float yi, const float zi,
                                         const float eps2, float& ax, float& ay, float& az) {
  const float dx = s.x - xi, dy = s.y - yi, dz = s.z - zi;
  float r2 = fmaf(dx, dx, eps2);
  r2 = fmaf(dy, dy, r2);
  r2 = fmaf(dz, dz, r2);
  float inv = rsqrt_approx(r2);
  if (EPS0) inv = r2 >= kTinyR2 ? inv : 0.f;  // d == 0 skip (OctreeSearch.h:102); a select, not a branch
  const float inv2 = inv * inv;
  const float minv = s.w * inv;
  const float w = minv * inv2;
  ax = fmaf(w, dx, ax);
  ay = fmaf(w, dy, ay);
  az = fmaf(w, dz, az);
}

// Scalar kernel. src must hold gridDim.y * chunk entries (padding = zero-mass bodies at the origin).
template <int I, bool EPS0, int MINB>
__global__ void __launch_bounds__(kDirectTPB, MINB)
direct_scalar_kernel(const float4* __restrict__ src, const int chunk, const float4* __restrict__ tgt,
                     const int n_tgt, const float eps2, float4* __restrict__ partial, const int n_tgt_pad) {
  __shared__ float4 tile[2][kDirectTJ];
  const int t = threadIdx.x;
  const int i_base = blockIdx.x * (kDirectTPB * I);
  float xi[I], yi[I], zi[I], ax[I], ay[I], az[I];
#pragma unroll
  for (int k = 0; k < I; k++) {
    int i = i_base + k * kDirectTPB + t;
    i = i < n_tgt ? i : n_tgt - 1;
    const float4 p = tgt[i];
    xi[k] = p.x; yi[k] = p.y; zi[k] = p.z;
    ax[k] = 0.f; ay[k] = 0.f; az[k] = 0.f;
  }
  const float4* s = src + (size_t)blockIdx.y * chunk;
  const int ntiles = chunk / kDirectTJ;
  float4 nxt = s[t];
  for (int tl = 0; tl < ntiles; tl++) {
    const int b = tl & 1;
    tile[b][t] = nxt;
    __syncthreads();  // one barrier per tile: buffer b^1 was last read before the previous barrier
    if (tl + 1 < ntiles) nxt = s[(size_t)(tl + 1) * kDirectTJ + t];
#pragma unroll 8
    for (int j = 0; j < kDirectTJ; j++) {
      const float4 sj = tile[b][j];
#pragma unroll
      for (int k = 0; k < I; k++) interact<EPS0>(sj, xi[k], yi[k], zi[k], eps2, ax[k], ay[k], az[k]);
    }
  }
  float4* out = partial + (size_t)blockIdx.y * n_tgt_pad;
#pragma unroll
  for (int k = 0; k < I; k++) {
    const int i = i_base + k * kDirectTPB + t;
    if (i < n_tgt) out[i] = make_float4(ax[k], ay[k], az[k], 0.f);
  }
}

// ---- packed-fp32 (f32x2) interaction over two sources ------------------------------------------------
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

// EQM = all sources carry the same mass: the mass factor leaves the inner loop (w = inv^3, 11 FMA-pipe lane-ops per
// interaction instead of 12) and is applied once per target by the consumer together with G.
template <bool EPS0, bool EQM = false>
__device__ __forceinline__ void interact2(const float2 X, const float2 Y, const float2 Z, const float2 M,
                                          const float2 nxi, const float2 nyi, const float2 nzi,
                                          const float2 eps2, float2& ax, float2& ay, float2& az) {
  const float2 dx = __fadd2_rn(X, nxi), dy = __fadd2_rn(Y, nyi), dz = __fadd2_rn(Z, nzi);
  float2 r2 = __ffma2_rn(dx, dx, eps2);
  r2 = __ffma2_rn(dy, dy, r2);
  r2 = __ffma2_rn(dz, dz, r2);
  float2 inv;
  inv.x = rsqrt_approx(r2.x);
  inv.y = rsqrt_approx(r2.y);
  if (EPS0) {
    inv.x = r2.x >= kTinyR2 ? inv.x : 0.f;
    inv.y = r2.y >= kTinyR2 ? inv.y : 0.f;
  }
  const float2 inv2 = __fmul2_rn(inv, inv);
  float2 w;
  if (EQM) {
    w = __fmul2_rn(inv, inv2);
  } else {
    const float2 minv = __fmul2_rn(M, inv);
    w = __fmul2_rn(minv, inv2);
  }
  ax = __ffma2_rn(w, dx, ax);
  ay = __ffma2_rn(w, dy, ay);
  az = __ffma2_rn(w, dz, az);
}

// Padding entries of the source array: so far away that r^2 overflows to +inf, rsqrt gives +0 and the pair contributes
// exactly nothing - with or without a mass factor (zero-mass padding would not do for the equal-mass kernel).
constexpr float kPadCoord = 3.0e19f;

// TPB threads per CTA = sources per shared-memory tile (default 256); chunk must be a multiple of TPB.
template <int I, bool EPS0, int MINB, int TPB = kDirectTPB, bool EQM = false>
__global__ void __launch_bounds__(TPB, MINB)
direct_packed_kernel(const float4* __restrict__ src, const int chunk, const float4* __restrict__ tgt,
                     const int n_tgt, const float eps2s, float4* __restrict__ partial, const int n_tgt_pad) {
  // SoA tile, double buffered: [buf][component][j]
  constexpr int NC = EQM ? 3 : 4;
  constexpr int UNR = EQM ? 2 : 4;   // deeper unrolling of the equal-mass loop spills at I = 8 (ptxas: 255 registers + 32 B stack)
  __shared__ __align__(16) float tile[2][NC][TPB];
  const int t = threadIdx.x;
  const int i_base = blockIdx.x * (TPB * I);
  float2 nxi[I], nyi[I], nzi[I], ax[I], ay[I], az[I];
#pragma unroll
  for (int k = 0; k < I; k++) {
    int i = i_base + k * TPB + t;
    i = i < n_tgt ? i : n_tgt - 1;
    const float4 p = tgt[i];
    nxi[k] = f2(-p.x, -p.x); nyi[k] = f2(-p.y, -p.y); nzi[k] = f2(-p.z, -p.z);
    ax[k] = f2(0.f, 0.f); ay[k] = f2(0.f, 0.f); az[k] = f2(0.f, 0.f);
  }
  const float2 eps2 = f2(eps2s, eps2s);
  const float4* s = src + (size_t)blockIdx.y * chunk;
  const int ntiles = chunk / TPB;
  float4 nxt = s[t];
  for (int tl = 0; tl < ntiles; tl++) {
    const int b = tl & 1;
    tile[b][0][t] = nxt.x; tile[b][1][t] = nxt.y; tile[b][2][t] = nxt.z;
    if (!EQM) tile[b][NC - 1][t] = nxt.w;
    __syncthreads();
    if (tl + 1 < ntiles) nxt = s[(size_t)(tl + 1) * TPB + t];
#pragma unroll UNR
    for (int j = 0; j < TPB; j += 4) {
      const float4 X = *reinterpret_cast<const float4*>(&tile[b][0][j]);
      const float4 Y = *reinterpret_cast<const float4*>(&tile[b][1][j]);
      const float4 Z = *reinterpret_cast<const float4*>(&tile[b][2][j]);
      float4 M = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!EQM) M = *reinterpret_cast<const float4*>(&tile[b][NC - 1][j]);
#pragma unroll
      for (int k = 0; k < I; k++) {
        interact2<EPS0, EQM>(f2(X.x, X.y), f2(Y.x, Y.y), f2(Z.x, Z.y), f2(M.x, M.y), nxi[k], nyi[k], nzi[k], eps2,
                             ax[k], ay[k], az[k]);
        interact2<EPS0, EQM>(f2(X.z, X.w), f2(Y.z, Y.w), f2(Z.z, Z.w), f2(M.z, M.w), nxi[k], nyi[k], nzi[k], eps2,
                             ax[k], ay[k], az[k]);
      }
    }
  }
  float4* out = partial + (size_t)blockIdx.y * n_tgt_pad;
#pragma unroll
  for (int k = 0; k < I; k++) {
    const int i = i_base + k * TPB + t;
    if (i < n_tgt) out[i] = make_float4(ax[k].x + ax[k].y, ay[k].x + ay[k].y, az[k].x + az[k].y, 0.f);
  }
}

}  // namespace nbody
