// K5 - device LSD radix sort of (64-bit Morton key, 32-bit body index) pairs, hand-written for sm_100a.
//
// Replaces the ordering that the reference obtains implicitly by inserting bodies one at a time into its pointer
// octree (Octree::Add, /root/reference/Source/NBody/OctreeSearch.h:60-81): sorting the bodies along the Morton curve
// makes every octree cell a contiguous index range.
//
// 8-bit digits, only the passes the tree needs. ONE kernel per pass ("onesweep"): a CTA takes the next 4096-key tile
// (ticket from an atomic counter), ranks its keys among equal digits with warp ballots, publishes its per-digit counts
// and finds its global offsets by a decoupled look-back over the tiles before it (one status word per tile and digit:
// 2 flag bits + 30 count bits, so flag and value travel in one store), then writes the tile out through shared memory,
// digit run by digit run, coalesced. The digit histogram a pass needs up front is produced by the pass before it (the
// keys are in registers anyway); the first one comes from the kernel that generates the keys (or radix_hist_kernel).
// HBM traffic per pass: 12 B read + 12 B written per body (+ 1 KB of status per tile).
#pragma once
#include <algorithm>

#include "common.cuh"

namespace nbody {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;                              // keys per thread
constexpr int kSortTile = kSortThreads * kSortItems;        // 4096 keys per CTA
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortMaxPasses = 8;
constexpr uint32_t kStatusAggregate = 1u << 30, kStatusPrefix = 2u << 30, kStatusValue = (1u << 30) - 1u;

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
  __shared__ uint32_t wsum[kSortWarps];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t inc = warp_incl_scan(v);
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t s = lane < kSortWarps ? wsum[lane] : 0u;
    s = warp_incl_scan(s);
    if (lane < kSortWarps) wsum[lane] = s;
  }
  __syncthreads();
  const uint32_t base = w ? wsum[w - 1] : 0u;
  *total = wsum[kSortWarps - 1];
  __syncthreads();
  return base + inc - v;
}

// Histogram of one digit over all keys -> ghist[256] (zeroed by the caller). Only for sorts whose keys do not come out
// of a kernel that can count on the way (the Morton-key kernel does).
__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const uint64_t* __restrict__ keys, const int n, const int shift, uint32_t* __restrict__ ghist) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  for (int i = blockIdx.x * kSortThreads + threadIdx.x; i < n; i += gridDim.x * kSortThreads)
    atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(ghist + threadIdx.x, h[threadIdx.x]);
}

// One pass. ghist = histogram of this pass' digit over all keys; ghist_next (may be NULL) receives the histogram of the
// next pass' digit (shift_next); status = ntiles x 256 words, ticket = 1 word, all zeroed by the caller.
// idx_in == nullptr: payload = position (first pass). Stable.
constexpr int kSortSmemBytes = kSortTile * 12 + (kSortWarps * 256 + 4 * 256) * 4 + 16;

__global__ void __launch_bounds__(kSortThreads, 3)
radix_onesweep_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ idx_in, const int n, const int shift,
                      const uint32_t* __restrict__ ghist, uint32_t* __restrict__ ghist_next, const int shift_next,
                      uint32_t* __restrict__ status, uint32_t* __restrict__ ticket, uint64_t* __restrict__ keys_out,
                      uint32_t* __restrict__ idx_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* skey = reinterpret_cast<uint64_t*>(smem_raw);                           // [kSortTile]
  uint32_t* sidx = reinterpret_cast<uint32_t*>(smem_raw + (size_t)kSortTile * 8);   // [kSortTile]
  uint32_t* wcount = sidx + kSortTile;                                              // [kSortWarps][256]
  uint32_t* dbase = wcount + kSortWarps * 256;                                      // [256] global position of digit d's first key of this tile
  uint32_t* lbase = dbase + 256;                                                    // [256] start of digit d inside the sorted tile
  uint32_t* hnext = lbase + 256;                                                    // [256] next pass' digit histogram of this tile
  uint32_t* stile = hnext + 256;                                                    // [1]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) *stile = atomicAdd(ticket, 1u);
#pragma unroll
  for (int k = 0; k < kSortWarps; k++) wcount[k * 256 + threadIdx.x] = 0;
  hnext[threadIdx.x] = 0;
  __syncthreads();
  const int tile = (int)*stile;
  const int seg = tile * kSortTile + w * (32 * kSortItems);
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
    if (ghist_next && live) atomicAdd(&hnext[(uint32_t)(key[r] >> shift_next) & 255u], 1u);
  }
  __syncthreads();
  uint32_t total_d;
  {  // thread d: prefix the warp counts of digit d over the CTA's warps
    uint32_t run = 0;
#pragma unroll
    for (int k = 0; k < kSortWarps; k++) { const uint32_t t = wcount[k * 256 + threadIdx.x]; wcount[k * 256 + threadIdx.x] = run; run += t; }
    total_d = run;
  }
  // publish this tile's count of digit d, then look back over the earlier tiles for the running total
  uint32_t* mine = status + (size_t)tile * 256 + threadIdx.x;
  uint32_t excl = 0;
  if (tile == 0) {
    __stcg(mine, kStatusPrefix | total_d);
  } else {
    __stcg(mine, kStatusAggregate | total_d);
    const volatile uint32_t* look = status + (size_t)(tile - 1) * 256 + threadIdx.x;
    while (true) {
      const uint32_t v = *look;
      if ((v >> 30) == 0u) continue;          // that tile has not published yet (it holds an earlier ticket: it is running)
      excl += v & kStatusValue;
      if (v & kStatusPrefix) break;
      look -= 256;
    }
    __stcg(mine, kStatusPrefix | (excl + total_d));
  }
  // global base of digit d = number of keys with a smaller digit + keys of digit d in earlier tiles
  uint32_t all;
  const uint32_t gbase = block_excl_scan(ghist[threadIdx.x], &all);
  dbase[threadIdx.x] = gbase + excl;
  uint32_t tile_total;
  lbase[threadIdx.x] = block_excl_scan(total_d, &tile_total);
  if (ghist_next && hnext[threadIdx.x]) atomicAdd(ghist_next + threadIdx.x, hnext[threadIdx.x]);
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
  uint32_t* work = nullptr;       // [kSortMaxPasses + 1][256] digit histograms | [16] tickets | [passes][ntiles][256] status
  size_t work_words = 0;
};

inline size_t radix_work_words(int64_t n) {
  const size_t ntiles = (size_t)ceil_div(std::max<int64_t>(n, 1), kSortTile);
  return (size_t)(kSortMaxPasses + 1) * 256 + 16 + (size_t)kSortMaxPasses * ntiles * 256;
}

struct RadixPlan {
  int p0 = 0, passes = 0;            // 8-bit passes p0 .. passes-1 (shift = 8 p)
  uint32_t* ghist0 = nullptr;        // where the first pass expects its digit histogram
  int shift0 = 0;
};

// Zeroes the work area for a sort of bits [first_bit rounded down to a multiple of 8, key_bits) and tells the producer of
// the keys where to count the first digit.
inline RadixPlan radix_sort_begin(RadixSortBuffers& b, int n, int key_bits, cudaStream_t s, int first_bit = 0) {
  RadixPlan pl;
  pl.passes = (key_bits + 7) / 8;
  pl.p0 = std::max(0, std::min(first_bit / 8, pl.passes - 1));
  const size_t ntiles = (size_t)ceil_div(n, kSortTile);
  const size_t used = (size_t)(kSortMaxPasses + 1) * 256 + 16 + (size_t)(pl.passes - pl.p0) * ntiles * 256;
  cudaMemsetAsync(b.work, 0, used * sizeof(uint32_t), s);
  pl.ghist0 = b.work;
  pl.shift0 = 8 * pl.p0;
  return pl;
}

// Runs the passes; the histogram of the first digit must already be in pl.ghist0. The result is in keys[out], idx[out]
// (returned index); payload = position in the input. Lower bits keep their input order (stable).
inline int radix_sort_run(RadixSortBuffers& b, const RadixPlan& pl, int n, cudaStream_t s, double* launches) {
  const int ntiles = (int)ceil_div(n, kSortTile);
  cudaFuncSetAttribute(radix_onesweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortSmemBytes);   // per device: cheap, idempotent
  uint32_t* tickets = b.work + (size_t)(kSortMaxPasses + 1) * 256;
  uint32_t* status = tickets + 16;
  int cur = 0;
  for (int p = pl.p0; p < pl.passes; p++) {
    const int k = p - pl.p0;
    const bool last = p + 1 == pl.passes;
    radix_onesweep_kernel<<<ntiles, kSortThreads, kSortSmemBytes, s>>>(
        b.keys[cur], p == pl.p0 ? nullptr : b.idx[cur], n, 8 * p, b.work + (size_t)k * 256, last ? nullptr : b.work + (size_t)(k + 1) * 256,
        8 * (p + 1), status + (size_t)k * ntiles * 256, tickets + k, b.keys[cur ^ 1], b.idx[cur ^ 1]);
    if (launches) *launches += 1;
    cur ^= 1;
  }
  return cur;
}

// Sorts keys[0] (payload = iota) by bits [first_bit rounded down to a multiple of 8, key_bits).
inline int radix_sort_pairs(RadixSortBuffers& b, int n, int key_bits, cudaStream_t s, double* launches, int first_bit = 0) {
  const RadixPlan pl = radix_sort_begin(b, n, key_bits, s, first_bit);
  const int blocks = (int)std::min<int64_t>(ceil_div(n, kSortThreads * 8), sm_count() * 8);
  radix_hist_kernel<<<std::max(blocks, 1), kSortThreads, 0, s>>>(b.keys[0], n, pl.shift0, pl.ghist0);
  if (launches) *launches += 1;
  return radix_sort_run(b, pl, n, s, launches);
}

}  // namespace nbody
