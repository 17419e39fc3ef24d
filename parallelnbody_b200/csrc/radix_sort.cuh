// K5 - device LSD radix sort of (64-bit Morton key, 32-bit body index) pairs, hand-written for sm_100a.
//
// Replaces the ordering that the reference obtains implicitly by inserting bodies one at a time into its pointer
// octree (Octree::Add, /root/reference/Source/NBody/OctreeSearch.h:60-81): sorting the bodies along the Morton curve
// makes every octree cell a contiguous index range.
//
// 8 passes of 8 bits (fewer when the tree needs fewer levels). Per pass: (1) per-CTA digit histogram, (2) exclusive scan
// of the digit-major table hist[digit][cta] (gives every CTA its global base per digit), (3) stable scatter of 4096-key
// tiles staged in shared memory (see radix_scatter_kernel). All HBM-bound: per pass 8 B key read (hist) + 12 B read +
// 12 B write (scatter) per body.
#pragma once
#include <algorithm>

#include "common.cuh"

namespace nbody {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;                              // keys per thread
constexpr int kSortTile = kSortThreads * kSortItems;        // 4096 keys per CTA
constexpr int kSortWarps = kSortThreads / 32;

// ---- exclusive scan of uint32 (three small kernels; n up to 2^31) --------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o) v += t;
  }
  return v;
}

// Block-wide exclusive scan of one value per thread (256 threads); returns the exclusive prefix, *total = block sum.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total) {
  __shared__ uint32_t wsum[kScanThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t inc = warp_incl_scan(v);
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t s = lane < kScanThreads / 32 ? wsum[lane] : 0u;
    s = warp_incl_scan(s);
    if (lane < kScanThreads / 32) wsum[lane] = s;
  }
  __syncthreads();
  const uint32_t base = w ? wsum[w - 1] : 0u;
  *total = wsum[kScanThreads / 32 - 1];
  __syncthreads();
  return base + inc - v;
}

__global__ void __launch_bounds__(kScanThreads)
scan_tile_sums_kernel(const uint32_t* __restrict__ in, const int64_t n, uint32_t* __restrict__ tile_sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) if (base + k < n) s += in[base + k];
  uint32_t total;
  block_excl_scan(s, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
// One CTA: in-place exclusive scan of m tile sums (m is small: n / 2048).
__global__ void __launch_bounds__(kScanThreads)
scan_spine_kernel(uint32_t* __restrict__ tile_sums, const int64_t m) {
  uint32_t carry = 0;
  for (int64_t base = 0; base < m; base += kScanThreads) {
    const int64_t i = base + threadIdx.x;
    const uint32_t v = i < m ? tile_sums[i] : 0u;
    uint32_t total;
    const uint32_t ex = block_excl_scan(v, &total);
    if (i < m) tile_sums[i] = carry + ex;
    carry += total;
  }
}
__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(uint32_t* __restrict__ data, const int64_t n, const uint32_t* __restrict__ tile_sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems], s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) { v[k] = base + k < n ? data[base + k] : 0u; s += v[k]; }
  uint32_t total;
  uint32_t ex = block_excl_scan(s, &total) + tile_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; k++) { if (base + k < n) data[base + k] = ex; ex += v[k]; }
}

// In-place exclusive scan; tile_sums must hold ceil(n / kScanTile) words. 3 launches.
inline void exclusive_scan_u32(uint32_t* data, int64_t n, uint32_t* tile_sums, cudaStream_t s, double* launches) {
  const int64_t tiles = ceil_div(n, kScanTile);
  scan_tile_sums_kernel<<<(unsigned)tiles, kScanThreads, 0, s>>>(data, n, tile_sums);
  scan_spine_kernel<<<1, kScanThreads, 0, s>>>(tile_sums, tiles);
  scan_apply_kernel<<<(unsigned)tiles, kScanThreads, 0, s>>>(data, n, tile_sums);
  if (launches) *launches += 3;
}

// ---- radix pass ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const uint64_t* __restrict__ keys, const int n, const int shift, uint32_t* __restrict__ hist,
                  const int nblocks) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * kSortTile;
  uint64_t key[kSortItems];
#pragma unroll
  for (int k = 0; k < kSortItems; k++) {
    const int i = base + k * kSortThreads + threadIdx.x;
    key[k] = i < n ? keys[i] : 0ull;
  }
#pragma unroll
  for (int k = 0; k < kSortItems; k++) {
    const int i = base + k * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(uint32_t)(key[k] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// hist must already hold the exclusive scan of the digit-major table. idx_in == nullptr: payload = position (first pass).
// Stable scatter in three steps: (1) every warp ranks its 512 keys among equal digits with warp ballots (round r,
// lane l = memory order); (2) warp counts are prefixed over the CTA's warps and digit totals over the digits, which
// gives each key its position in the CTA's sorted tile - the tile is staged in shared memory in that order; (3) the
// staged tile is written out front to back, so consecutive threads write consecutive addresses within each digit run.
constexpr int kSortSmemBytes = kSortTile * 12 + (kSortWarps * 256 + 3 * 256) * 4;

__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ idx_in, const int n,
                     const int shift, const uint32_t* __restrict__ hist, const int nblocks,
                     uint64_t* __restrict__ keys_out, uint32_t* __restrict__ idx_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* skey = reinterpret_cast<uint64_t*>(smem_raw);                       // [kSortTile]
  uint32_t* sidx = reinterpret_cast<uint32_t*>(smem_raw + (size_t)kSortTile * 8);   // [kSortTile]
  uint32_t* wcount = sidx + kSortTile;                                          // [kSortWarps][256]
  uint32_t* dbase = wcount + kSortWarps * 256;                                  // [256] global base of digit d for this CTA
  uint32_t* lbase = dbase + 256;                                                // [256] start of digit d inside the sorted tile
  uint32_t* dtot = lbase + 256;                                                 // [256]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kSortWarps; k++) wcount[k * 256 + threadIdx.x] = 0;
  __syncthreads();
  const int seg = blockIdx.x * kSortTile + w * (32 * kSortItems);
  uint64_t key[kSortItems];
  uint32_t rank[kSortItems];
  // all 16 loads first (independent, one memory round trip), then the ranking rounds, which synchronise the warp
#pragma unroll
  for (int r = 0; r < kSortItems; r++) {
    const int i = seg + 32 * r + lane;
    key[r] = i < n ? keys_in[i] : ~0ull;
  }
#pragma unroll
  for (int r = 0; r < kSortItems; r++) {
    const int i = seg + 32 * r + lane;
    const bool live = i < n;
    const uint32_t d = live ? ((uint32_t)(key[r] >> shift) & 255u) : 0u;
    // lanes holding the same digit: 8 ballots (one per digit bit) + 1 for liveness; much cheaper than MATCH.ANY on ~30
    // distinct values per warp
    uint32_t peers = __ballot_sync(0xffffffffu, live);
    if (!live) peers = ~peers;
#pragma unroll
    for (int b = 0; b < 8; b++) {
      const uint32_t bal = __ballot_sync(0xffffffffu, (d >> b) & 1u);
      peers &= ((d >> b) & 1u) ? bal : ~bal;
    }
    const uint32_t before = __popc(peers & ((1u << lane) - 1u));
    uint32_t prev = 0;
    if (live) prev = wcount[w * 256 + d];
    __syncwarp();
    if (live && before == 0) wcount[w * 256 + d] = prev + __popc(peers);
    __syncwarp();
    rank[r] = prev + before;
  }
  __syncthreads();
  uint32_t total_d;
  {  // thread d: prefix the warp counts of digit d over the CTA's warps, fetch the global base
    uint32_t run = 0;
#pragma unroll
    for (int k = 0; k < kSortWarps; k++) { const uint32_t t = wcount[k * 256 + threadIdx.x]; wcount[k * 256 + threadIdx.x] = run; run += t; }
    total_d = run;
    dtot[threadIdx.x] = run;
    dbase[threadIdx.x] = hist[(size_t)threadIdx.x * nblocks + blockIdx.x];
  }
  uint32_t tile_total;
  lbase[threadIdx.x] = block_excl_scan(total_d, &tile_total);
  __syncthreads();
  uint32_t pay[kSortItems];
#pragma unroll
  for (int r = 0; r < kSortItems; r++) {
    const int i = seg + 32 * r + lane;
    pay[r] = (idx_in && i < n) ? idx_in[i] : (uint32_t)i;
  }
#pragma unroll
  for (int r = 0; r < kSortItems; r++) {
    const int i = seg + 32 * r + lane;
    if (i < n) {
      const uint32_t d = (uint32_t)(key[r] >> shift) & 255u;
      const uint32_t lpos = lbase[d] + wcount[w * 256 + d] + rank[r];
      skey[lpos] = key[r];
      sidx[lpos] = pay[r];
    }
  }
  __syncthreads();
  for (uint32_t j = threadIdx.x; j < tile_total; j += kSortThreads) {
    const uint64_t k = skey[j];
    const uint32_t d = (uint32_t)(k >> shift) & 255u;
    const uint32_t pos = dbase[d] + (j - lbase[d]);
    keys_out[pos] = k;
    idx_out[pos] = sidx[j];
  }
}

struct RadixSortBuffers {
  uint64_t* keys[2] = {nullptr, nullptr};
  uint32_t* idx[2] = {nullptr, nullptr};
  uint32_t* hist = nullptr;       // 256 * nblocks
  uint32_t* tile_sums = nullptr;  // ceil(256 * nblocks / kScanTile)
};

// Sorts keys[0] (payload = iota) by bits [first_bit rounded down to a multiple of 8, key_bits); the result is in
// keys[out], idx[out] (returned index). Lower bits keep their input order (stable).
inline int radix_sort_pairs(RadixSortBuffers& b, int n, int key_bits, cudaStream_t s, double* launches, int first_bit = 0) {
  const int nblocks = (int)ceil_div(n, kSortTile);
  static const cudaError_t attr_rc = cudaFuncSetAttribute(radix_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortSmemBytes);
  (void)attr_rc;
  int cur = 0;
  const int passes = (key_bits + 7) / 8, p0 = std::max(0, std::min(first_bit / 8, passes - 1));
  for (int p = p0; p < passes; p++) {
    const int shift = 8 * p;
    radix_hist_kernel<<<nblocks, kSortThreads, 0, s>>>(b.keys[cur], n, shift, b.hist, nblocks);
    if (launches) *launches += 1;
    exclusive_scan_u32(b.hist, (int64_t)256 * nblocks, b.tile_sums, s, launches);
    radix_scatter_kernel<<<nblocks, kSortThreads, kSortSmemBytes, s>>>(b.keys[cur], p == p0 ? nullptr : b.idx[cur], n, shift, b.hist, nblocks,
                                                           b.keys[cur ^ 1], b.idx[cur ^ 1]);
    if (launches) *launches += 1;
    cur ^= 1;
  }
  return cur;
}

}  // namespace nbody
