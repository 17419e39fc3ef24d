// Barnes-Hut path on sm_100a: K4 Morton keys, K5 radix sort (radix_sort.cuh), K6 octree topology, K7 monopoles,
// K8 force walk. See bh.cuh for what each replaces in the reference.
//
// Tree representation (all in HBM, rebuilt every step like the reference does, OctreeSearch.cpp:78-81):
//   bodies sorted by 63-bit Morton key (21 bits per axis, X most significant as Octree::GetOctant, OctreeSearch.h:50-56),
//   so every octree cell is a contiguous body range. Nodes form a COMPRESSED octree: a node is the smallest cell that
//   contains its bodies (chains of one-child cells that the reference materialises, h:65-78, are skipped - they all carry
//   the same monopole and the reference accepts the chain iff it would accept its smallest cell, so the result is the same);
//   children of a node are contiguous in index and stored in octant order.
//     node_com[k]  = (centre of mass xyz, total mass)                       float4
//     node_meta[k] = (first child | first body, #children | #bodies, level | leaf flag, parent)   int4
//     node_range[k] = (first body, one past last body)                      int2
//   A node with <= leaf_size bodies (or at the deepest level) is a leaf. Cell half-width = root half-width / 2^level,
//   the reference's `Size` (h:70-74).
#include "bh.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "comm.h"
#include "direct_kernels.cuh"
#include "radix_sort.cuh"

namespace cg = cooperative_groups;

namespace nbody {
namespace {

constexpr int kMaxLevel = 21;            // 3 * 21 = 63 key bits
constexpr int kLeafFlag = 1 << 8;
// walk groups hold <= group_size bodies (32 * B, B bodies per lane, B in {1, 2, 4})
constexpr int kWalkThreads = 256;
constexpr int kWalkWarps = kWalkThreads / 32;
constexpr int kWalkMinCtas1 = 3;         // one body per lane: measured 2.19 ms at 3 CTAs per SM, 2.28 at 4 (64 registers), 2.45 at 2 (100 registers)
constexpr int kWalkMinCtas = 3;          // register budget: 80 per thread -> 24 warps per SM (64 registers spill the streamed-leaf walk)
constexpr int kStackCap = 8192;          // per-warp spill slab of the walk stack (HBM/L2 resident; cells only, rarely touched)
constexpr int kStackSmem = 512;          // per-warp stack window in shared memory
constexpr int kStackMask = kStackSmem - 1;
constexpr int kStackWin = 768;           // the single-warp walk's linear stack window (entries); a round pushes at most 256
constexpr int kListCap = 128;            // per-warp interaction ring in shared memory
constexpr int kFlush = 64;               // pending entries evaluated per flush (the ring also holds up to 32 more + 31 left over)
constexpr int kLeafChunk = 1 << 16;      // leaves at least this large are streamed on their own (bounds the packed scan of the walk)
constexpr int kLetSamples = 256;         // key samples per rank for the domain splitters
constexpr int kMaxWorld = 16;
constexpr int kCostBins = 1024;          // equal-count bins along the sorted bodies in which the walks record their work

struct Counters {
  int nnodes, ngroups, ticket, depth, next_group, overflow, pad0, pad1;
  int gen_off[kMaxLevel + 4];
  unsigned long long interactions;
  // the split's own grid barrier: arrival counter and release word, each in a cache line of its own (the release word is
  // polled by every block while the working ones hit nnodes / ngroups with atomics)
  int pad2[32 - ((8 + kMaxLevel + 4 + 2) & 31)];
  int bar_count, pad3[31];
  int bar_gen, pad4[31];
};

struct Impl {
  int64_t cap_n = 0;
  RadixSortBuffers sort;
  int sorted = 0;                   // which sort buffer holds the sorted keys
  int64_t cap_nodes = 0;
  float4* node_com = nullptr;
  int4* node_meta = nullptr;
  int2* node_range = nullptr;
  uint32_t* node_ready = nullptr;
  float4* node_bmin = nullptr;       // bounding box of every node's bodies (domain-split mode only)
  float4* node_bmax = nullptr;
  int64_t cap_boxes_nodes = 0;
  int2* groups = nullptr;            // walk groups: body ranges of <= group_size neighbours
  int* group_cost = nullptr;         // interaction-list length of each group in the last walk (work measure for the domain split)
  Counters* counters = nullptr;
  float4* root = nullptr;          // [0] = cube (centre, half-width); [1] = previous root COM (xyz) + valid flag (w)
  int* stacks = nullptr;
  int64_t cap_stacks = 0;
  float* boxes = nullptr; int64_t cap_boxes = 0;
  int n = 0;
  int built_group_size = 0;
  // --- multi-GPU domain decomposition + locally-essential-tree exchange (K9), allocated on first use
  int let_world = 0;
  uint64_t* samples = nullptr;     // [world * kLetSamples] gathered key samples
  uint64_t* splitters = nullptr;   // [world + 1] first key of every rank's domain (0 ... ~0)
  int* send_off = nullptr;         // [2 world + 1] body ranges per destination rank | export counts per peer
  int* all_off = nullptr;          // [world * (2 world + 1)] every rank's message
  char* peer_pub = nullptr;        // [world * kPubBytes] every rank's published boundary tree (cells + bounding boxes)
  int* pub_node = nullptr;         // [kLetPub] local tree node behind each published cell
  uint32_t* visit = nullptr;       // export descent: 2 x (cells, peer masks) frontiers of cap_visit entries + 2 counters
  int64_t cap_visit = 0;
  float4* let_out = nullptr;       // [world * cap_let] per-peer export lists
  int* let_cnt = nullptr;          // [world] export counts (the tail of the send_off message, not an allocation of its own)
  uint32_t* bin_cost = nullptr;    // [kCostBins] interactions evaluated for the bodies of each equal-count bin (last step)
  bool splitters_valid = false;    // the splitters describe balanced domains of the system this handle last ran
  int* h_counts = nullptr;         // pinned host copy of the gathered per-step count message
  cudaEvent_t ev_counts = nullptr; // fires when h_counts is filled
  float* ret = nullptr;            // read-back staging (return to owner)
  int64_t cap_ret = 0;
  int64_t cap_let = 0;
  float4* let_in = nullptr;        // received points
  float4* let_sorted = nullptr;    // the same in Morton order (sources of the LET tree)
  int64_t cap_let_in = 0;
  float4* all_pos = nullptr;       // diagnostics: every rank's bodies (energy)
  int64_t cap_all_pos = 0;
};

Impl* impl_of(BHState& st) {
  if (!st.impl) st.impl = new Impl();
  return static_cast<Impl*>(st.impl);
}

template <class T>
int realloc_dev(T** p, size_t count) {
  if (*p) { NB_CUDA(cudaFree(*p)); *p = nullptr; }
  NB_CUDA(cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)));
  return 0;
}

int ensure(Impl* m, int n, cudaStream_t s) {
  if (!m->counters) {
    NB_TRY(realloc_dev(&m->counters, 1));
    NB_TRY(realloc_dev(&m->root, 2));
    NB_CUDA(cudaMemsetAsync(m->counters, 0, sizeof(Counters), s));
    NB_CUDA(cudaMemsetAsync(m->root, 0, 2 * sizeof(float4), s));
  }
  if (n <= m->cap_n) return 0;
  NB_CUDA(cudaStreamSynchronize(s));
  const size_t c = (size_t)n + (size_t)n / 8 + 1024;
  for (int k = 0; k < 2; k++) { NB_TRY(realloc_dev(&m->sort.keys[k], c)); NB_TRY(realloc_dev(&m->sort.idx[k], c)); }
  m->sort.work_words = radix_work_words((int64_t)c);
  NB_TRY(realloc_dev(&m->sort.work, m->sort.work_words));
  const size_t nodes = 2 * c + 8;
  NB_TRY(realloc_dev(&m->node_com, nodes));
  NB_TRY(realloc_dev(&m->node_meta, nodes));
  NB_TRY(realloc_dev(&m->node_range, nodes));
  NB_TRY(realloc_dev(&m->node_ready, nodes));
  NB_TRY(realloc_dev(&m->groups, c));
  NB_TRY(realloc_dev(&m->group_cost, c));
  m->cap_nodes = (int64_t)nodes;
  m->cap_n = (int64_t)c;
  return 0;
}

// ---- K4: root cube + Morton keys -----------------------------------------------------------------------------
// box = cube_size_kernel output. reference_root: centre = previous root COM (0 on the first build), half-width =
// max |coordinate| (OctreeSearch.cpp:47-56,77-79) - such a cube need not contain every body, exactly as in the
// reference, where outliers are still routed by the octant comparisons; here their quantised coordinates clamp.
// Otherwise: the tight bounding cube.
__global__ void root_cube_kernel(const uint32_t* __restrict__ box, const int mode, float4* __restrict__ root) {
  float cx, cy, cz, half;
  if (mode == 1) {   // the reference's root
    const float4 prev = root[1];
    cx = prev.x; cy = prev.y; cz = prev.z;
    half = __uint_as_float(box[0]);
  } else {
    const float lx = ordered_to_float(box[1]), ly = ordered_to_float(box[2]), lz = ordered_to_float(box[3]);
    const float hx = ordered_to_float(box[4]), hy = ordered_to_float(box[5]), hz = ordered_to_float(box[6]);
    cx = 0.5f * (lx + hx); cy = 0.5f * (ly + hy); cz = 0.5f * (lz + hz);
    half = 0.5f * fmaxf(fmaxf(hx - lx, hy - ly), hz - lz);
    half = half * 1.0001f + 1e-30f;
    if (mode == 2) {
      // sticky cube (multi-GPU domain split): keep the previous cube while it still holds every body and is not far too
      // large, so that keys - and the domain splitters, which are keys - mean the same from step to step
      const float4 prev = root[0];
      const bool inside = prev.w > 0.f && lx >= prev.x - prev.w && hx <= prev.x + prev.w && ly >= prev.y - prev.w && hy <= prev.y + prev.w &&
                          lz >= prev.z - prev.w && hz <= prev.z + prev.w;
      if (inside && prev.w <= 3.f * half) return;
      half *= 1.25f;
    }
  }
  if (!(half > 0.f) || !isfinite(half)) half = 1.f;
  root[0] = make_float4(cx, cy, cz, half);
}

__device__ __forceinline__ uint64_t expand21(uint32_t v) {
  uint64_t x = v & 0x1fffffu;
  x = (x | x << 32) & 0x001f00000000ffffull;
  x = (x | x << 16) & 0x001f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
__device__ __forceinline__ uint32_t compact21(uint64_t x) {
  x &= 0x1249249249249249ull;
  x = (x | x >> 2) & 0x10c30c30c30c30c3ull;
  x = (x | x >> 4) & 0x100f00f00f00f00full;
  x = (x | x >> 8) & 0x001f0000ff0000ffull;
  x = (x | x >> 16) & 0x001f00000000ffffull;
  x = (x | x >> 32) & 0x1fffffull;
  return (uint32_t)x;
}
// Cell index along one axis: u = (x - centre) / half in [-1, 1) -> floor((u + 1) * 2^20), clamped. The division keeps the
// reference's octant comparisons exact where they are exact: x >= centre <=> the top bit is set (OctreeSearch.h:50-56),
// e.g. for the central body that CreateSpacePoints puts exactly at the root centre (OctreeSearch.cpp:68-70).
__device__ __forceinline__ uint32_t quantize(float x, float centre, float half) {
  const float u = __fdiv_rn(__fsub_rn(x, centre), half);
  const float f = floorf(__fmul_rn(__fadd_rn(u, 1.f), 1048576.f));
  return (uint32_t)fminf(fmaxf(f, 0.f), 2097151.f);
}

// ghist (may be NULL): histogram of the key digit at `shift` = what the first radix pass needs up front; counted here so the
// sort never re-reads the keys for it.
__global__ void __launch_bounds__(256)
morton_kernel(const float4* __restrict__ posm, const int n, const float4* __restrict__ root, uint64_t* __restrict__ keys,
              uint32_t* __restrict__ ghist, const int shift, const BodySegs segs = BodySegs{0, 0x7fffffff, 0}) {
  __shared__ uint32_t h[256];
  if (ghist) { h[threadIdx.x] = 0; __syncthreads(); }
  const float4 c = root[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = ld_stream(posm + segs.at(i));
    const uint32_t qx = quantize(p.x, c.x, c.w), qy = quantize(p.y, c.y, c.w), qz = quantize(p.z, c.z, c.w);
    const uint64_t k = expand21(qx) << 2 | expand21(qy) << 1 | expand21(qz);   // octant digit = 4*X + 2*Y + Z (OctreeSearch.h:50-56)
    keys[i] = k;
    if (ghist) atomicAdd(&h[(uint32_t)(k >> shift) & 255u], 1u);
  }
  if (ghist) {
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(ghist + threadIdx.x, h[threadIdx.x]);
  }
}

// order[i] = logical input index of the i-th sorted body; segs maps logical indices to storage (the bodies that stayed on
// this rank, then the ones that arrived: see bh_let_finish).
__global__ void __launch_bounds__(256)
gather_bodies_kernel(const uint32_t* __restrict__ order, const int n, const BodySegs segs, const float4* __restrict__ posm_in,
                     const float4* __restrict__ vel_in, const int32_t* __restrict__ ids_in, float4* __restrict__ posm_out,
                     float4* __restrict__ vel_out, int32_t* __restrict__ ids_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = segs.at((int)order[i]);
  st_stream(posm_out + i, posm_in[j]);
  if (vel_in) st_stream(vel_out + i, vel_in[j]);
  if (ids_in) ids_out[i] = ids_in[j];
}

// ---- K6: octree topology ------------------------------------------------------------------------------------
// Number of leading 3-bit digits two keys share = level of the smallest cell containing both.
__device__ __forceinline__ int common_levels(uint64_t a, uint64_t b) {
  const uint64_t x = a ^ b;
  if (x == 0) return kMaxLevel;
  return (__clzll((long long)x) - 1) / 3;
}

// Cuts the body range [b, e) into ceil(len / group_size) near-equal walk groups.
__device__ __forceinline__ void emit_groups(const int b, const int e, const int group_size, int2* __restrict__ groups,
                                            Counters* __restrict__ c) {
  const int len = e - b;
  if (len <= 0) return;
  const int chunks = (len + group_size - 1) / group_size;
  const int g0 = atomicAdd(&c->ngroups, chunks);
  for (int k = 0; k < chunks; k++)
    groups[g0 + k] = make_int2(b + (int)((long long)len * k / chunks), b + (int)((long long)len * (k + 1) / chunks));
}

__global__ void tree_init_kernel(const uint64_t* __restrict__ keys, const int n, const int group_size, const int super, int2* __restrict__ range,
                                 int4* __restrict__ meta, uint32_t* __restrict__ ready, int2* __restrict__ groups,
                                 Counters* __restrict__ c) {
  const int lvl = n > 1 ? common_levels(keys[0], keys[n - 1]) : kMaxLevel;
  range[0] = make_int2(0, n);
  meta[0] = make_int4(0, 0, lvl, -1);
  ready[0] = 0;
  c->nnodes = 1; c->ngroups = 0; c->ticket = 0; c->depth = 0; c->next_group = 0; c->overflow = 0; c->bar_count = 0; c->bar_gen = 0;
  for (int k = 0; k < kMaxLevel + 4; k++) c->gen_off[k] = 1;
  c->gen_off[0] = 0;
  c->interactions = 0;
  if (n <= super) emit_groups(0, n, group_size, groups, c);
}

// The top-down split, all generations in ONE cooperative launch (grid-wide barrier between generations): every node
// created by the previous generation either becomes a leaf or is cut at its level's octant digit into its non-empty
// children (8 lanes per node, one octant boundary each, found by binary search in the sorted keys). Equivalent of the
// recursive re-insertion in Octree::Add (OctreeSearch.h:65-78). `levels` = how many levels the keys are sorted to: a cell
// at that level is a leaf whatever it holds (21 = all 63 key bits).
__device__ __forceinline__ int split_generation(const int gb, const int ge, const uint64_t* __restrict__ keys, const int leaf_size,
                                                const int group_size, const int super, const int levels, int2* __restrict__ range,
                                                int4* __restrict__ meta, uint32_t* __restrict__ ready, int2* __restrict__ groups,
                                                Counters* __restrict__ c) {
  const int lane = threadIdx.x & 31, sub = lane & 7, gshift = lane & ~7;
  const unsigned gmask = 0xffu << gshift;
  const int stride = gridDim.x * blockDim.x / 8;
  int maxlvl = 0;
  // a warp takes 4 consecutive nodes per trip (8 lanes each); the trip count is uniform over the warp
  for (int node0 = gb + (blockIdx.x * blockDim.x + (threadIdx.x & ~31)) / 8; node0 < ge; node0 += stride) {
    const int node = node0 + (lane >> 3);
    const bool live = node < ge;
    int2 r = make_int2(0, 0);
    int4 m = make_int4(0, 0, 0, -1);
    if (live) { r = range[node]; m = meta[node]; }
    const int cnt = r.y - r.x, level = m.z;
    const bool split = live && cnt > leaf_size && level < levels;
    if (live && !split) {
      if (sub == 0) {
        meta[node] = make_int4(r.x, cnt, level | kLeafFlag, m.w);
        // > super bodies in one deepest-level cell (coincident to the key resolution): walk them in chunks
        if (cnt > super) emit_groups(r.x, r.y, group_size, groups, c);
      }
      int d = m.w >= 0 ? (meta[m.w].z & 255) + 1 : 0;   // depth of the leaf's cell in the reference's tree
      if (cnt > leaf_size && level < kMaxLevel) d = levels + 1;   // the sort was too shallow for this cell: ask for more next time
      maxlvl = max(maxlvl, d);
    }
    const int shift = 3 * (kMaxLevel - 1 - min(level, kMaxLevel - 1));
    // first body whose digit at this level is >= sub: 8-ary search (7 independent probes per round, so the chain of
    // dependent L2 round trips is log8 of the range), then a binary search on the last few bodies
    int lo = r.x, hi = split ? r.y : r.x;
    while (hi - lo > 7) {
      const int w = (hi - lo) >> 3;
      int d[7];
#pragma unroll
      for (int k = 0; k < 7; k++) d[k] = (int)((keys[lo + (k + 1) * w] >> shift) & 7ull);
      int nlo = lo, nhi = hi;
      bool found = false;
#pragma unroll
      for (int k = 0; k < 7; k++) {
        if (!found) {
          if (d[k] >= sub) { nhi = lo + (k + 1) * w; found = true; }
          else nlo = lo + (k + 1) * w + 1;
        }
      }
      lo = nlo; hi = nhi;
    }
    if (lo < hi) {   // <= 7 candidates left: probe them all at once (independent loads) instead of a dependent binary search
      const int len = hi - lo;
      int d[7];
#pragma unroll
      for (int k = 0; k < 7; k++) d[k] = k < len ? (int)((keys[lo + k] >> shift) & 7ull) : 8;
      int first = len;
#pragma unroll
      for (int k = 6; k >= 0; k--) if (d[k] >= sub) first = k;
      lo += min(first, len);
    }
    __syncwarp();
    const int start = lo;
    int next = __shfl_down_sync(0xffffffffu, start, 1, 8);
    if (sub == 7) next = r.y;
    const int cc = split ? next - start : 0;
    const unsigned nonempty = (__ballot_sync(0xffffffffu, cc > 0) >> gshift) & 0xffu;
    const int nchild = __popc(nonempty), slot = __popc(nonempty & ((1u << sub) - 1u));
    // node allocation: one atomic per warp (up to 4 nodes being split), not one per node - millions of same-address
    // atomics per generation serialise in L2
    const int n0 = __shfl_sync(0xffffffffu, nchild, 0), n1 = __shfl_sync(0xffffffffu, nchild, 8), n2 = __shfl_sync(0xffffffffu, nchild, 16),
              n3 = __shfl_sync(0xffffffffu, nchild, 24);
    int base = 0;
    if (lane == 0 && n0 + n1 + n2 + n3 > 0) base = atomicAdd(&c->nnodes, n0 + n1 + n2 + n3);
    base = __shfl_sync(0xffffffffu, base, 0) + (gshift >= 8 ? n0 : 0) + (gshift >= 16 ? n1 : 0) + (gshift >= 24 ? n2 : 0);
    if (cc > 0) {
      const int child = base + slot;
      range[child] = make_int2(start, next);
      // smallest cell holding the child's bodies; beyond `levels` the keys are unsorted, so the prefix is only known to there
      meta[child] = make_int4(0, 0, cc > 1 ? min(common_levels(keys[start], keys[next - 1]), levels) : kMaxLevel, node);
      ready[child] = 0;
    }
    // Walk groups. A cell with more than `super` bodies hands its small children (<= super bodies each) to the walk:
    // runs of consecutive small children (adjacent octants, contiguous bodies) are cut into equal chunks of <= group_size
    // bodies, so the walk's lanes are well filled while every chunk stays inside this cell.
    const bool grouping = split && cnt > super;
    if (__any_sync(0xffffffffu, grouping)) {
      int run_begin = 0, run_end = 0;
      for (int k = 0; k < 8; k++) {
        const int ck = __shfl_sync(0xffffffffu, cc, gshift + k), sk = __shfl_sync(0xffffffffu, start, gshift + k);
        if (!grouping || sub != 0 || ck == 0) continue;
        if (ck <= super) { if (run_end == run_begin) run_begin = sk; run_end = sk + ck; }
        else { emit_groups(run_begin, run_end, group_size, groups, c); run_begin = run_end = 0; }
      }
      if (grouping && sub == 0) emit_groups(run_begin, run_end, group_size, groups, c);
    }
    if (split && sub == 0) meta[node] = make_int4(base, nchild, level, m.w);
  }
  (void)gmask;
  return maxlvl;
}

__global__ void __launch_bounds__(256, 6)
tree_split_kernel(const uint64_t* __restrict__ keys, const int leaf_size, const int group_size, const int super, const int levels,
                  int2* __restrict__ range, int4* __restrict__ meta, uint32_t* __restrict__ ready, int2* __restrict__ groups,
                  Counters* __restrict__ c) {
  // One grid-wide barrier per generation, written out here instead of cg::grid::sync() (which would need a second one
  // so that nobody allocates nodes again before everyone has read the count): the LAST block to arrive records how many
  // nodes exist - everything the generation created - in gen_off and only then releases the others, who read it there.
  // The launch is cooperative, so all blocks are co-resident and the spin cannot starve anyone.
  int gb = 0, ge = 1, maxlvl = 0, gen = 0;
  for (; gen <= kMaxLevel; gen++) {
    maxlvl = max(maxlvl, split_generation(gb, ge, keys, leaf_size, group_size, super, levels, range, meta, ready, groups, c));
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const int arrived = atomicAdd(&c->bar_count, 1) + 1;
      if (arrived == (int)gridDim.x * (gen + 1)) {
        c->gen_off[gen + 2] = atomicAdd(&c->nnodes, 0);
        __threadfence();
        atomicExch(&c->bar_gen, gen + 1);
      } else {
        while (*((volatile int*)&c->bar_gen) < gen + 1) __nanosleep(20);
      }
      __threadfence();
    }
    __syncthreads();
    const int next = *((volatile int*)&c->gen_off[gen + 2]);
    gb = ge; ge = next;
    if (gb == ge) break;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) for (int g = gen + 3; g < kMaxLevel + 4; g++) c->gen_off[g] = ge;
  if (maxlvl) atomicMax(&c->depth, maxlvl);
}

// ---- K7: monopoles (Octree::ComputeMass, OctreeSearch.h:83-97) ---------------------------------------------------
// Leaves sum their bodies; the last child to arrive at a parent sums the parent's children in octant order:
//   TotalMass += Mc;  CenterOfMass += COMc * Mc;  CenterOfMass *= 1 / TotalMass   (fp32, as the reference; M == 0 keeps
// a position inside the cell, h:95), with the first-order sums taken about the first child.
template <bool BOXES>
__global__ void __launch_bounds__(256)
monopole_kernel(const float4* __restrict__ posm, const int2* __restrict__ range, const int4* __restrict__ meta,
                uint32_t* __restrict__ ready, float4* __restrict__ com, const Counters* __restrict__ c,
                float4* __restrict__ root, float4* __restrict__ bmin, float4* __restrict__ bmax) {
  // BOXES: also the bounding box of every node's bodies (bmin / bmax), carried up the same way - the domain-split mode
  // describes a rank's domain to its peers with them
  const int nn = c->nnodes;
  for (int node = blockIdx.x * blockDim.x + threadIdx.x; node < nn; node += gridDim.x * blockDim.x) {
    int4 m = meta[node];
    if (!(m.z & kLeafFlag)) continue;
    // sums are taken about the first member (mathematically the same COM; exact when the members coincide, so a
    // body never sees its own cell's monopole at a rounding-error distance)
    const float4 p0 = posm[m.x];
    float M = 0.f, x = 0.f, y = 0.f, z = 0.f;
    float lx = p0.x, ly = p0.y, lz = p0.z, hx = p0.x, hy = p0.y, hz = p0.z;
    for (int b = m.x; b < m.x + m.y; b++) {
      const float4 p = posm[b];
      M += p.w; x += (p.x - p0.x) * p.w; y += (p.y - p0.y) * p.w; z += (p.z - p0.z) * p.w;
      if (BOXES) { lx = fminf(lx, p.x); ly = fminf(ly, p.y); lz = fminf(lz, p.z); hx = fmaxf(hx, p.x); hy = fmaxf(hy, p.y); hz = fmaxf(hz, p.z); }
    }
    if (M != 0.f) { const float rv = 1.f / M; x = p0.x + x * rv; y = p0.y + y * rv; z = p0.z + z * rv; }
    else { x = p0.x; y = p0.y; z = p0.z; }
    __stcg(com + node, make_float4(x, y, z, M));
    if (BOXES) { __stcg(bmin + node, make_float4(lx, ly, lz, 0.f)); __stcg(bmax + node, make_float4(hx, hy, hz, 0.f)); }
    while (true) {
      const int parent = m.w;
      if (parent < 0) { root[1] = make_float4(x, y, z, 1.f); break; }   // next reference-mode root centre (OctreeSearch.cpp:77)
      __threadfence();
      m = meta[parent];
      if (atomicAdd(ready + parent, 1u) != (uint32_t)(m.y - 1)) break;
      __threadfence();
      const float4 q0 = __ldcg(com + m.x);
      M = 0.f; x = 0.f; y = 0.f; z = 0.f;
      for (int k = 0; k < m.y; k++) {
        const float4 q = __ldcg(com + m.x + k);
        M += q.w; x += (q.x - q0.x) * q.w; y += (q.y - q0.y) * q.w; z += (q.z - q0.z) * q.w;
        if (BOXES) {
          const float4 a = __ldcg(bmin + m.x + k), b = __ldcg(bmax + m.x + k);
          if (k == 0) { lx = a.x; ly = a.y; lz = a.z; hx = b.x; hy = b.y; hz = b.z; }
          else { lx = fminf(lx, a.x); ly = fminf(ly, a.y); lz = fminf(lz, a.z); hx = fmaxf(hx, b.x); hy = fmaxf(hy, b.y); hz = fmaxf(hz, b.z); }
        }
      }
      if (M != 0.f) { const float rv = 1.f / M; x = q0.x + x * rv; y = q0.y + y * rv; z = q0.z + z * rv; }
      else { x = q0.x; y = q0.y; z = q0.z; }
      __stcg(com + parent, make_float4(x, y, z, M));
      if (BOXES) { __stcg(bmin + parent, make_float4(lx, ly, lz, 0.f)); __stcg(bmax + parent, make_float4(hx, hy, hz, 0.f)); }
    }
  }
}

// ---- K8a: warp-coherent group walk ----------------------------------------------------------------------------
// One warp per group of <= group_size Morton-neighbouring bodies (B = 1, 2 or 4 per lane, in registers). The warp pops up to 32 stack
// entries per round, one per lane; each lane tests its cell against the GROUP's bounding box:
//     accept  <=>  half-width / dmin < theta,  dmin = distance(box, cell COM)          (cf. Size / d < Theta, h:103)
// dmin <= every member's own d, so an accepted cell is one the reference would accept for each member. Accepted cells
// and the bodies of opened leaves are appended to a 128-entry ring in shared memory; whenever 64 are pending, all lanes
// evaluate them against their bodies with the direct-sum interaction (no divergence; the one evaluation site). Opened cells
// push their children; the bodies of opened leaves never touch the stack - they are streamed through the ring 32 at a
// time. The stack holds cells only: its top lives in a linear shared-memory window (768 entries per warp, plain stores
// for the pushes), older entries spill to a per-warp slab in global memory.
// The pending ring is SoA (x[128] | y[128] | z[128] | m[128]) so that one LDS.128 yields four consecutive x (y, z, m) and the
// evaluation runs on PAIRS of entries with Blackwell packed fp32 (FADD2 / FFMA2 / FMUL2), as K1 does.
// What bounds it (ncu, profiles/r2_ncu_summary_direct_and_walk.md): instruction issue shared by traversal (38 % of the
// instructions) and evaluation; three CTAs per SM is the measured optimum (kWalkMinCtas1).
template <int B, bool EPS0>
__device__ __forceinline__ void eval_list(const float* __restrict__ ring, const int head, const int count, const float2 (&nx)[B],
                                          const float2 (&ny)[B], const float2 (&nz)[B], const float2 eps2,
                                          float2 (&ax)[B], float2 (&ay)[B], float2 (&az)[B]) {
#pragma unroll 2
  for (int j = 0; j < count; j += 4) {
    const float4 X = *reinterpret_cast<const float4*>(ring + head + j);
    const float4 Y = *reinterpret_cast<const float4*>(ring + kListCap + head + j);
    const float4 Z = *reinterpret_cast<const float4*>(ring + 2 * kListCap + head + j);
    const float4 M = *reinterpret_cast<const float4*>(ring + 3 * kListCap + head + j);
#pragma unroll
    for (int k = 0; k < B; k++) {
      interact2<EPS0>(f2(X.x, X.y), f2(Y.x, Y.y), f2(Z.x, Z.y), f2(M.x, M.y), nx[k], ny[k], nz[k], eps2, ax[k], ay[k], az[k]);
      interact2<EPS0>(f2(X.z, X.w), f2(Y.z, Y.w), f2(Z.z, Z.w), f2(M.z, M.w), nx[k], ny[k], nz[k], eps2, ax[k], ay[k], az[k]);
    }
  }
}

template <int B, bool EPS0>
__global__ void __launch_bounds__(kWalkThreads, B == 1 ? kWalkMinCtas1 : B == 2 ? kWalkMinCtas : 2)
bh_walk_group_kernel(const float4* __restrict__ posm, const float4* __restrict__ node_com, const int4* __restrict__ node_meta,
                     const float4* __restrict__ tgt, const int2* __restrict__ groups, Counters* __restrict__ c,
                     const float4* __restrict__ root, const float theta2, const float eps2, const float G, const int t0,
                     const int t1, const int accumulate, int* __restrict__ stacks, float4* __restrict__ acc,
                     int* __restrict__ group_cost, uint32_t* __restrict__ bin_cost, const int n_targets) {
  // posm / node_* = the SOURCE tree; tgt / groups / c = the targets and their walk groups (the same tree, or - for
  // the locally-essential points received from other ranks - the local tree whose bodies are being accelerated)
  __shared__ __align__(16) float ring_all[kWalkWarps][4 * kListCap];
  __shared__ int stk_all[kWalkWarps][kStackWin];    // top of the warp's cell stack (a LINEAR window: entry i of the window at stk[i]); older entries spill to the global slab
  __shared__ int offs_all[kWalkWarps][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* ring = ring_all[w];
  int* stk = stk_all[w];
  int* offs = offs_all[w];
  int* gstack = stacks + (size_t)(blockIdx.x * kWalkWarps + w) * kStackCap;
  const int ngroups = c->ngroups;
  const float root_half = root[0].w;
  const unsigned lt = (1u << lane) - 1u;
  const float2 eps2v = f2(eps2, eps2);
  unsigned long long inter = 0;
  while (true) {
    int g = 0;
    if (lane == 0) g = atomicAdd(&c->next_group, 1);
    g = __shfl_sync(0xffffffffu, g, 0);
    if (g >= ngroups) break;
    const int2 r = groups[g];
    const int ntarget = min(r.y, t1) - max(r.x, t0);
    if (ntarget <= 0) continue;
    float2 nx[B], ny[B], nz[B], ax[B], ay[B], az[B];   // negated target coordinates / accumulators, duplicated per half
    float lox = 3.4e38f, loy = 3.4e38f, loz = 3.4e38f, hix = -3.4e38f, hiy = -3.4e38f, hiz = -3.4e38f;
#pragma unroll
    for (int k = 0; k < B; k++) {
      ax[k] = f2(0.f, 0.f); ay[k] = f2(0.f, 0.f); az[k] = f2(0.f, 0.f);
      const int i = r.x + lane + 32 * k;
      const float4 p = tgt[i < r.y ? i : r.x];
      nx[k] = f2(-p.x, -p.x); ny[k] = f2(-p.y, -p.y); nz[k] = f2(-p.z, -p.z);
      lox = fminf(lox, p.x); loy = fminf(loy, p.y); loz = fminf(loz, p.z);
      hix = fmaxf(hix, p.x); hiy = fmaxf(hiy, p.y); hiz = fmaxf(hiz, p.z);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lox = fminf(lox, __shfl_xor_sync(0xffffffffu, lox, o)); hix = fmaxf(hix, __shfl_xor_sync(0xffffffffu, hix, o));
      loy = fminf(loy, __shfl_xor_sync(0xffffffffu, loy, o)); hiy = fmaxf(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
      loz = fminf(loz, __shfl_xor_sync(0xffffffffu, loz, o)); hiz = fmaxf(hiz, __shfl_xor_sync(0xffffffffu, hiz, o));
    }
    const float gcx = 0.5f * (lox + hix), gcy = 0.5f * (loy + hiy), gcz = 0.5f * (loz + hiz);
    const float ghx = 0.5f * (hix - lox), ghy = 0.5f * (hiy - loy), ghz = 0.5f * (hiz - loz);

    // logical stack = `base` spilled entries in the slab + the window stk[0 .. top); pushes are plain stores at stk[top + k]
    // (no wrap-around arithmetic on the hot path; when the window fills up its oldest part moves to the slab and the rest slides down)
    int top = 1, base = 0, head = 0, pending = 0;
    int leaf_first = 0, leaf_excl = 0, ltotal = 0, lpos = 0;   // bodies of the leaves opened by the last round, being streamed
    if (lane == 0) stk[0] = 0;
    __syncwarp();
    unsigned long long entries = 0;
    bool overflow = false;
    // One loop, one evaluation site. Each trip either streams the next 32 bodies of the opened leaves into the ring, or - when
    // none are left - pops up to 32 cells, tests them and queues what they yield; kFlush pending entries are evaluated at once.
    while (true) {
      if (lpos < ltotal) {
        // flat index j over all opened leaves of the round -> owner lane by binary search in the exclusive prefix of the sizes
        const int j = lpos + lane;
        int L = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) if (offs[L + step] <= j) L += step;
        const int first = __shfl_sync(0xffffffffu, leaf_first, L), off = __shfl_sync(0xffffffffu, leaf_excl, L);
        const int cnt = min(32, ltotal - lpos);
        if (lane < cnt) {
          const float4 bd = posm[first + (j - off)];
          const int slot = (head + pending + lane) & (kListCap - 1);
          ring[slot] = bd.x; ring[kListCap + slot] = bd.y; ring[2 * kListCap + slot] = bd.z; ring[3 * kListCap + slot] = bd.w;
        }
        pending += cnt;
        lpos += 32;
      } else if (top + base > 0) {
        if (top == 0) {   // window empty: bring the youngest spilled entries back
          const int cnt = min(base, kStackWin / 2);
          for (int k = lane; k < cnt; k += 32) stk[k] = gstack[base - cnt + k];
          base -= cnt;
          top = cnt;
          __syncwarp();
        }
        const int nb = min(32, top);
        top -= nb;
        const int e = lane < nb ? stk[top + lane] : -1;
        float4 item = make_float4(0.f, 0.f, 0.f, 0.f);
        bool has_item = false;
        int push_first = 0, push_n = 0, leaf_n = 0;
        leaf_first = 0;
        if (e >= 0) {
          const float4 cm = node_com[e];
          const int4 m = node_meta[e];
          const bool leaf = (m.z & kLeafFlag) != 0;
          const float size = root_half * __int_as_float((127 - (m.z & 255)) << 23);   // half-width / 2^level
          const float dx = fmaxf(fabsf(cm.x - gcx) - ghx, 0.f), dy = fmaxf(fabsf(cm.y - gcy) - ghy, 0.f),
                      dz = fmaxf(fabsf(cm.z - gcz) - ghz, 0.f);
          const float dmin2 = dx * dx + dy * dy + dz * dz;
          if (size * size < theta2 * dmin2 || (leaf && m.y == 1)) { item = cm; has_item = true; }
          else if (leaf) { leaf_first = m.x; leaf_n = m.y; }
          else { push_first = m.x; push_n = m.y; }
        }
        // a leaf too large for the packed scan below (a deepest-level cell full of coincident bodies): stream it on its own
        unsigned big = __ballot_sync(0xffffffffu, leaf_n >= kLeafChunk);
        while (big) {
          const int src = __ffs(big) - 1;
          big &= big - 1;
          const int f = __shfl_sync(0xffffffffu, leaf_first, src), n = __shfl_sync(0xffffffffu, leaf_n, src);
          for (int b0 = 0; b0 < n; b0 += 32) {
            const int cnt = min(32, n - b0);
            if (lane < cnt) {
              const float4 bd = posm[f + b0 + lane];
              const int slot = (head + pending + lane) & (kListCap - 1);
              ring[slot] = bd.x; ring[kListCap + slot] = bd.y; ring[2 * kListCap + slot] = bd.z; ring[3 * kListCap + slot] = bd.w;
            }
            pending += cnt;
            __syncwarp();
            if (pending >= 32) {
              eval_list<B, EPS0>(ring, head, 32, nx, ny, nz, eps2v, ax, ay, az);
              head = (head + 32) & (kListCap - 1);
              pending -= 32;
              entries += 32;
              __syncwarp();
            }
          }
          if (lane == src) leaf_n = 0;
        }
        // one warp scan for both counts: children to push (<= 8 per lane, low 10 bits) and leaf bodies to stream (high bits)
        int incl = push_n | (leaf_n << 10);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const int both = __shfl_sync(0xffffffffu, incl, 31);
        const int total = both & 1023;
        ltotal = both >> 10;
        lpos = 0;
        leaf_excl = (incl >> 10) - leaf_n;
        if (ltotal) offs[lane] = leaf_excl;
        if (total) {   // children of the opened cells go on the stack
          if (top + total > kStackWin) {   // make room (rare): the oldest entries of the window move to the slab, the rest slides down
            const int sp = min(top, max(kStackWin / 2, top + total - kStackWin));
            if (base + sp > kStackCap) { overflow = true; break; }
            for (int k = lane; k < sp; k += 32) gstack[base + k] = stk[k];
            base += sp;
            __syncwarp();
            for (int k0 = 0; k0 < top - sp; k0 += 32) {
              const int v = k0 + lane < top - sp ? stk[sp + k0 + lane] : 0;
              __syncwarp();
              if (k0 + lane < top - sp) stk[k0 + lane] = v;
            }
            top -= sp;
            __syncwarp();
          }
          int* dst = stk + top + (incl & 1023) - push_n;
#pragma unroll
          for (int k = 0; k < 8; k++) if (k < push_n) dst[k] = push_first + k;
          top += total;
        }
        // accepted cells join the pending interaction ring
        const unsigned mask = __ballot_sync(0xffffffffu, has_item);
        if (has_item) {
          const int slot = (head + pending + __popc(mask & lt)) & (kListCap - 1);
          ring[slot] = item.x; ring[kListCap + slot] = item.y; ring[2 * kListCap + slot] = item.z; ring[3 * kListCap + slot] = item.w;
        }
        pending += __popc(mask);
      } else {
        break;
      }
      __syncwarp();
      if (pending >= kFlush) {
        eval_list<B, EPS0>(ring, head, kFlush, nx, ny, nz, eps2v, ax, ay, az);
        head = (head + kFlush) & (kListCap - 1);
        pending -= kFlush;
        entries += kFlush;
        __syncwarp();
      }
    }
    if (overflow) { if (lane == 0) atomicExch(&c->overflow, 1); pending = 0; }
    if (pending > 0) {  // tail: pad the ring with massless entries up to a multiple of 4 so the evaluation stays branch-free
      const int padded = (pending + 3) & ~3;
      if (lane < padded - pending) {
        const int slot = (head + pending + lane) & (kListCap - 1);
        ring[slot] = 0.f; ring[kListCap + slot] = 0.f; ring[2 * kListCap + slot] = 0.f; ring[3 * kListCap + slot] = 0.f;
      }
      __syncwarp();
      eval_list<B, EPS0>(ring, head, padded, nx, ny, nz, eps2v, ax, ay, az);
      entries += pending;
      __syncwarp();
    }
    inter += entries * (unsigned long long)ntarget;
    if (group_cost && lane == 0) group_cost[g] = (accumulate ? group_cost[g] : 0) + (int)min(entries, 0x3fffffffull);
    // work record for the domain split: interactions of this group, binned by its position in the sorted order
    if (bin_cost && lane == 0)
      atomicAdd(bin_cost + min((int)((long long)r.x * kCostBins / max(n_targets, 1)), kCostBins - 1), (uint32_t)min(entries * (unsigned long long)ntarget, 0xffffffffull));
#pragma unroll
    for (int k = 0; k < B; k++) {
      const int i = r.x + lane + 32 * k;
      if (i < r.y && i >= t0 && i < t1) {
        float4 a = make_float4(G * (ax[k].x + ax[k].y), G * (ay[k].x + ay[k].y), G * (az[k].x + az[k].y), 0.f);
        if (accumulate) { const float4 o = acc[i]; a.x += o.x; a.y += o.y; a.z += o.z; }
        acc[i] = a;
      }
    }
  }
  if (lane == 0 && inter) atomicAdd(&c->interactions, inter);
}

// ---- K8a': the same walk, warp-specialised ---------------------------------------------------------------------
// The walk above alternates, inside every warp, a latency-bound traversal (dependent smem / L2 accesses, scans) with an
// FMA-pipe-bound evaluation, so the FMA pipe idles whenever a scheduler's warps are traversing at the same time (ncu: 68 %
// busy). Here the two halves run in different warps: warps 0..3 of a CTA TRAVERSE (one group each, exactly the rule and
// order of the kernel above) and stream what they accept through a shared-memory ring of kWsChunks x 64 entries to
// warps 4..7, which hold the groups' bodies in registers and only EVALUATE. Hand-off per 64-entry chunk through a pair
// of mbarriers (full / empty, one elected lane arrives, every lane of the other warp waits); a chunk carries its group and
// a LAST flag, so the evaluating warp writes the accelerations when the group ends. One producer feeds one consumer, chunks
// are consumed in order: the summation order - hence the result - is that of the single-warp kernel, bit for bit.
constexpr int kWsPairs = 4;
constexpr int kWsThreads = 64 * kWsPairs;
constexpr int kWsChunk = 64;
constexpr int kWsChunks = 4;
constexpr int kWsLast = 1 << 16;
// register split between the two warpgroups of a CTA (launch: 64 per thread; setmaxnreg moves the budget to where the
// instruction-level parallelism pays - the evaluation keeps eight interaction chains in flight per warp)
template <int SPLIT> struct WsRegs { static constexpr int traverse = SPLIT == 1 ? 40 : 48, evaluate = SPLIT == 1 ? 88 : 80; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, const int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, const uint32_t parity) {
  asm volatile(
      "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

template <int B, bool EPS0>
__device__ __forceinline__ void eval_chunk(const float* __restrict__ ch, const int count, const float2 (&nx)[B], const float2 (&ny)[B],
                                           const float2 (&nz)[B], const float2 eps2, float2 (&ax)[B], float2 (&ay)[B], float2 (&az)[B]) {
#pragma unroll 2
  for (int j = 0; j < count; j += 4) {
    const float4 X = *reinterpret_cast<const float4*>(ch + j);
    const float4 Y = *reinterpret_cast<const float4*>(ch + kWsChunk + j);
    const float4 Z = *reinterpret_cast<const float4*>(ch + 2 * kWsChunk + j);
    const float4 M = *reinterpret_cast<const float4*>(ch + 3 * kWsChunk + j);
#pragma unroll
    for (int k = 0; k < B; k++) {
      interact2<EPS0>(f2(X.x, X.y), f2(Y.x, Y.y), f2(Z.x, Z.y), f2(M.x, M.y), nx[k], ny[k], nz[k], eps2, ax[k], ay[k], az[k]);
      interact2<EPS0>(f2(X.z, X.w), f2(Y.z, Y.w), f2(Z.z, Z.w), f2(M.z, M.w), nx[k], ny[k], nz[k], eps2, ax[k], ay[k], az[k]);
    }
  }
}

template <int B, bool EPS0, int SPLIT>
__global__ void __launch_bounds__(kWsThreads, B <= 2 ? 4 : 2)
bh_walk_ws_kernel(const float4* __restrict__ posm, const float4* __restrict__ node_com, const int4* __restrict__ node_meta,
                  const float4* __restrict__ tgt, const int2* __restrict__ groups, Counters* __restrict__ c,
                  const float4* __restrict__ root, const float theta2, const float eps2, const float G, const int t0,
                  const int t1, const int accumulate, int* __restrict__ stacks, float4* __restrict__ acc,
                  int* __restrict__ group_cost, uint32_t* __restrict__ bin_cost, const int n_targets) {
  __shared__ __align__(16) float ring_all[kWsPairs][kWsChunks * 4 * kWsChunk];   // chunk-major: x[64] | y[64] | z[64] | m[64]
  __shared__ int stk_all[kWsPairs][kStackSmem];
  __shared__ int offs_all[kWsPairs][32];
  __shared__ int2 info_all[kWsPairs][kWsChunks];               // (group, entries | LAST) of a chunk; group < 0 = no more work
  __shared__ __align__(8) uint64_t full_all[kWsPairs][kWsChunks], empty_all[kWsPairs][kWsChunks];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, w = warp % kWsPairs;
  if (threadIdx.x < kWsPairs * kWsChunks) {
    mbar_init(&full_all[0][0] + threadIdx.x, 1);
    mbar_init(&empty_all[0][0] + threadIdx.x, 1);
  }
  __syncthreads();
  float* ring = ring_all[w];
  int2* info = info_all[w];
  uint64_t* full = full_all[w];
  uint64_t* empty = empty_all[w];

  if (warp >= kWsPairs) {
    // ------------------------------------------------------------------ evaluating warp
    if (SPLIT > 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WsRegs<SPLIT>::evaluate));
    const float2 eps2v = f2(eps2, eps2);
    float2 nx[B], ny[B], nz[B], ax[B], ay[B], az[B];
    int2 r = make_int2(0, 0);
    int cur = -1;
    for (unsigned q = 0;; q++) {
      const int ch = q & (kWsChunks - 1);
      mbar_wait(full + ch, (q / kWsChunks) & 1);
      const int2 inf = info[ch];
      if (inf.x < 0) break;
      if (inf.x != cur) {
        cur = inf.x;
        r = groups[cur];
#pragma unroll
        for (int k = 0; k < B; k++) {
          ax[k] = f2(0.f, 0.f); ay[k] = f2(0.f, 0.f); az[k] = f2(0.f, 0.f);
          const int i = r.x + lane + 32 * k;
          const float4 p = tgt[i < r.y ? i : r.x];
          nx[k] = f2(-p.x, -p.x); ny[k] = f2(-p.y, -p.y); nz[k] = f2(-p.z, -p.z);
        }
      }
      eval_chunk<B, EPS0>(ring + ch * (4 * kWsChunk), inf.y & (kWsLast - 1), nx, ny, nz, eps2v, ax, ay, az);
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + ch);
      if (inf.y & kWsLast) {
#pragma unroll
        for (int k = 0; k < B; k++) {
          const int i = r.x + lane + 32 * k;
          if (i < r.y && i >= t0 && i < t1) {
            float4 a = make_float4(G * (ax[k].x + ax[k].y), G * (ay[k].x + ay[k].y), G * (az[k].x + az[k].y), 0.f);
            if (accumulate) { const float4 o = acc[i]; a.x += o.x; a.y += o.y; a.z += o.z; }
            acc[i] = a;
          }
        }
        cur = -1;
      }
    }
    return;
  }

  // -------------------------------------------------------------------- traversing warp
  if (SPLIT > 0) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WsRegs<SPLIT>::traverse));
  int* stk = stk_all[w];
  int* offs = offs_all[w];
  int* gstack = stacks + (size_t)(blockIdx.x * kWsPairs + w) * kStackCap;
  const int ngroups = c->ngroups;
  const float root_half = root[0].w;
  const unsigned lt = (1u << lane) - 1u;
  unsigned long long inter = 0;
  unsigned wpos = 0;             // entries written so far (chunk = wpos / 64)
  unsigned acq = 0, done = 0;    // chunks acquired for writing / handed over
  int g = 0;
  // every chunk up to the one holding entry `upto` must have been released by the evaluating warp
  auto ensure = [&](const unsigned upto) {
    const unsigned need = upto / kWsChunk + 1;
    while (acq < need) { mbar_wait(empty + (acq & (kWsChunks - 1)), ((acq / kWsChunks) & 1) ^ 1); acq++; }
  };
  auto put = [&](const unsigned pos, const float4 v) {
    float* d = ring + ((pos / kWsChunk) & (kWsChunks - 1)) * (4 * kWsChunk) + (pos & (kWsChunk - 1));
    d[0] = v.x; d[kWsChunk] = v.y; d[2 * kWsChunk] = v.z; d[3 * kWsChunk] = v.w;
  };
  auto commit = [&]() {   // hand over the chunks that filled up
    while (done < wpos / kWsChunk) {
      __syncwarp();
      if (lane == 0) { info[done & (kWsChunks - 1)] = make_int2(g, kWsChunk); mbar_arrive(full + (done & (kWsChunks - 1))); }
      done++;
    }
  };
  while (true) {
    if (lane == 0) g = atomicAdd(&c->next_group, 1);
    g = __shfl_sync(0xffffffffu, g, 0);
    if (g >= ngroups) break;
    const int2 r = groups[g];
    const int ntarget = min(r.y, t1) - max(r.x, t0);
    if (ntarget <= 0) continue;
    float lox = 3.4e38f, loy = 3.4e38f, loz = 3.4e38f, hix = -3.4e38f, hiy = -3.4e38f, hiz = -3.4e38f;
#pragma unroll
    for (int k = 0; k < B; k++) {
      const int i = r.x + lane + 32 * k;
      const float4 p = tgt[i < r.y ? i : r.x];
      lox = fminf(lox, p.x); loy = fminf(loy, p.y); loz = fminf(loz, p.z);
      hix = fmaxf(hix, p.x); hiy = fmaxf(hiy, p.y); hiz = fmaxf(hiz, p.z);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lox = fminf(lox, __shfl_xor_sync(0xffffffffu, lox, o)); hix = fmaxf(hix, __shfl_xor_sync(0xffffffffu, hix, o));
      loy = fminf(loy, __shfl_xor_sync(0xffffffffu, loy, o)); hiy = fmaxf(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
      loz = fminf(loz, __shfl_xor_sync(0xffffffffu, loz, o)); hiz = fmaxf(hiz, __shfl_xor_sync(0xffffffffu, hiz, o));
    }
    const float gcx = 0.5f * (lox + hix), gcy = 0.5f * (loy + hiy), gcz = 0.5f * (loz + hiz);
    const float ghx = 0.5f * (hix - lox), ghy = 0.5f * (hiy - loy), ghz = 0.5f * (hiz - loz);

    int top = 1, base = 0;
    int leaf_first = 0, leaf_excl = 0, ltotal = 0, lpos = 0;
    if (lane == 0) stk[0] = 0;
    __syncwarp();
    const unsigned wstart = wpos;
    bool overflow = false;
    while (true) {
      if (lpos < ltotal) {   // stream the next 32 bodies of the leaves opened by the last round
        const int j = lpos + lane;
        int L = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) if (offs[L + step] <= j) L += step;
        const int first = __shfl_sync(0xffffffffu, leaf_first, L), off = __shfl_sync(0xffffffffu, leaf_excl, L);
        const int cnt = min(32, ltotal - lpos);
        float4 bd = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < cnt) bd = posm[first + (j - off)];
        ensure(wpos + cnt - 1);
        if (lane < cnt) put(wpos + lane, bd);
        wpos += cnt;
        lpos += 32;
      } else if (top > 0) {
        if (top == base) {   // window empty: bring the youngest spilled entries back
          const int cnt = min(base, kStackSmem / 2);
          for (int k = lane; k < cnt; k += 32) stk[(base - cnt + k) & kStackMask] = gstack[base - cnt + k];
          base -= cnt;
          __syncwarp();
        }
        const int nb = min(32, top - base);
        top -= nb;
        const int e = lane < nb ? stk[(top + lane) & kStackMask] : -1;
        float4 item = make_float4(0.f, 0.f, 0.f, 0.f);
        bool has_item = false;
        int push_first = 0, push_n = 0, leaf_n = 0;
        leaf_first = 0;
        if (e >= 0) {
          const float4 cm = node_com[e];
          const int4 m = node_meta[e];
          const bool leaf = (m.z & kLeafFlag) != 0;
          const float size = root_half * __int_as_float((127 - (m.z & 255)) << 23);
          const float dx = fmaxf(fabsf(cm.x - gcx) - ghx, 0.f), dy = fmaxf(fabsf(cm.y - gcy) - ghy, 0.f),
                      dz = fmaxf(fabsf(cm.z - gcz) - ghz, 0.f);
          const float dmin2 = dx * dx + dy * dy + dz * dz;
          if (size * size < theta2 * dmin2 || (leaf && m.y == 1)) { item = cm; has_item = true; }
          else if (leaf) { leaf_first = m.x; leaf_n = m.y; }
          else { push_first = m.x; push_n = m.y; }
        }
        // a leaf too large for the packed scan below (a deepest-level cell full of coincident bodies): stream it on its own
        unsigned big = __ballot_sync(0xffffffffu, leaf_n >= kLeafChunk);
        while (big) {
          const int src = __ffs(big) - 1;
          big &= big - 1;
          const int f = __shfl_sync(0xffffffffu, leaf_first, src), n = __shfl_sync(0xffffffffu, leaf_n, src);
          for (int b0 = 0; b0 < n; b0 += 32) {
            const int cnt = min(32, n - b0);
            ensure(wpos + cnt - 1);
            if (lane < cnt) put(wpos + lane, posm[f + b0 + lane]);
            wpos += cnt;
            commit();
          }
          if (lane == src) leaf_n = 0;
        }
        int incl = push_n | (leaf_n << 10);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const int both = __shfl_sync(0xffffffffu, incl, 31);
        const int total = both & 1023;
        ltotal = both >> 10;
        lpos = 0;
        leaf_excl = (incl >> 10) - leaf_n;
        if (ltotal) offs[lane] = leaf_excl;
        if (total) {
          const int need = (top - base) + total - kStackSmem;
          if (need > 0) {
            const int sp = min((need + 31) & ~31, top - base);
            if (base + sp > kStackCap) { overflow = true; break; }
            for (int k = lane; k < sp; k += 32) gstack[base + k] = stk[(base + k) & kStackMask];
            base += sp;
            __syncwarp();
          }
          const int dst = top + (incl & 1023) - push_n;
#pragma unroll
          for (int k = 0; k < 8; k++) if (k < push_n) stk[(dst + k) & kStackMask] = push_first + k;
          top += total;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, has_item);
        if (mask) {
          const int na = __popc(mask);
          ensure(wpos + na - 1);
          if (has_item) put(wpos + __popc(mask & lt), item);
          wpos += na;
        }
      } else {
        break;
      }
      __syncwarp();
      commit();
    }
    if (overflow && lane == 0) atomicExch(&c->overflow, 1);
    // end of the group: pad the open chunk with massless entries up to a multiple of 4 and hand it over with the LAST flag
    const unsigned long long entries = wpos - wstart;
    {
      const unsigned rem = wpos & (kWsChunk - 1), padded = (rem + 3) & ~3u;
      ensure(wpos);
      if (lane < padded - rem) put(wpos + lane, make_float4(0.f, 0.f, 0.f, 0.f));
      __syncwarp();
      if (lane == 0) { info[done & (kWsChunks - 1)] = make_int2(g, (int)padded | kWsLast); mbar_arrive(full + (done & (kWsChunks - 1))); }
      done++;
      wpos = done * kWsChunk;
    }
    inter += entries * (unsigned long long)ntarget;
    if (group_cost && lane == 0) group_cost[g] = (accumulate ? group_cost[g] : 0) + (int)min(entries, 0x3fffffffull);
    if (bin_cost && lane == 0)
      atomicAdd(bin_cost + min((int)((long long)r.x * kCostBins / max(n_targets, 1)), kCostBins - 1), (uint32_t)min(entries * (unsigned long long)ntarget, 0xffffffffull));
  }
  ensure(wpos);
  __syncwarp();
  if (lane == 0) { info[done & (kWsChunks - 1)] = make_int2(-1, 0); mbar_arrive(full + (done & (kWsChunks - 1))); }
  if (lane == 0 && inter) atomicAdd(&c->interactions, inter);
}

// ---- K8b: per-body walk with the reference's exact rule and order (parity mode) ---------------------------------
// Octree::ComputeForces (OctreeSearch.h:99-108): d = |COM - x| (fp32 sqrtf); d == 0 -> skip; accept when Size / d < Theta or
// one-body leaf: a += (float)(G * M / d^3) * (COM - x), scalar in double; else the children in octant order.
__device__ __forceinline__ void ref_pair(const float4 s, const float4 p, const float G, const float eps2, float& ax, float& ay,
                                         float& az, int& count) {
  const float dx = __fsub_rn(p.x, s.x), dy = __fsub_rn(p.y, s.y), dz = __fsub_rn(p.z, s.z);
  const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  float d = sqrtf(d2);
  if (d == 0.f) return;
  if (eps2 > 0.f) d = sqrtf(__fadd_rn(d2, eps2));
  const double dd = (double)d;
  const float sc = (float)((double)G * (double)s.w / (dd * dd * dd));
  ax = __fadd_rn(ax, __fmul_rn(__fsub_rn(s.x, p.x), sc));
  ay = __fadd_rn(ay, __fmul_rn(__fsub_rn(s.y, p.y), sc));
  az = __fadd_rn(az, __fmul_rn(__fsub_rn(s.z, p.z), sc));
  count++;
}

__global__ void __launch_bounds__(128)
bh_walk_body_kernel(const float4* __restrict__ posm, const float4* __restrict__ node_com, const int4* __restrict__ node_meta,
                    Counters* __restrict__ c, const float4* __restrict__ root, const float theta, const float eps2,
                    const float G, const int t0, const int t1, float4* __restrict__ acc) {
  const int i = t0 + blockIdx.x * blockDim.x + threadIdx.x;
  int count = 0;
  if (i < t1) {
    const float4 p = posm[i];
    const float root_half = root[0].w;
    float ax = 0.f, ay = 0.f, az = 0.f;
    int stack[kMaxLevel * 7 + 16];
    int sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
      const int e = stack[--sp];
      const float4 cm = node_com[e];
      const int4 m = node_meta[e];
      const float dx = __fsub_rn(p.x, cm.x), dy = __fsub_rn(p.y, cm.y), dz = __fsub_rn(p.z, cm.z);
      const float d = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
      if (d == 0.f) continue;                                                 // h:102
      const bool leaf = (m.z & kLeafFlag) != 0;
      const float size = root_half * __int_as_float((127 - (m.z & 255)) << 23);
      if (__fdiv_rn(size, d) < theta || (leaf && m.y == 1)) {                 // h:103
        ref_pair(cm, p, G, eps2, ax, ay, az, count);                          // h:104
      } else if (leaf) {
        for (int b = m.x; b < m.x + m.y; b++) ref_pair(posm[b], p, G, eps2, ax, ay, az, count);
      } else {
        for (int k = m.y - 1; k >= 0; k--) stack[sp++] = m.x + k;             // popped in octant order (h:105-107)
      }
    }
    acc[i] = make_float4(ax, ay, az, 0.f);
  }
  // interaction count: warp-reduce then one atomic per warp
  unsigned long long v = (unsigned long long)count;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(&c->interactions, v);
}

// ---- read-back of what DrawOctreeBoxes draws (OctreeSearch.cpp:36-45): one box per occupied leaf ---------------------
__global__ void __launch_bounds__(256)
leaf_boxes_kernel(const uint64_t* __restrict__ keys, const int4* __restrict__ meta, const Counters* __restrict__ c,
                  const float4* __restrict__ root, float* __restrict__ boxes7, const int cap, int* __restrict__ nout) {
  const int nn = c->nnodes;
  const float4 rc = root[0];
  for (int node = blockIdx.x * blockDim.x + threadIdx.x; node < nn; node += gridDim.x * blockDim.x) {
    const int4 m = meta[node];
    if (!(m.z & kLeafFlag) || m.y <= 0) continue;
    // the cell this leaf hangs in: one level below its parent (the reference's leaf cell), the root itself if alone
    const int lvl = m.w < 0 ? 0 : min((meta[m.w].z & 255) + 1, kMaxLevel);
    const uint64_t key = keys[m.x];
    const int drop = kMaxLevel - lvl;
    const uint32_t qx = compact21(key >> 2) >> drop, qy = compact21(key >> 1) >> drop, qz = compact21(key) >> drop;
    const float half = rc.w * __int_as_float((127 - lvl) << 23);
    const int slot = atomicAdd(nout, 1);
    if (slot < cap) {
      float* o = boxes7 + (size_t)slot * 7;
      o[0] = rc.x - rc.w + (2.f * (float)qx + 1.f) * half;
      o[1] = rc.y - rc.w + (2.f * (float)qy + 1.f) * half;
      o[2] = rc.z - rc.w + (2.f * (float)qz + 1.f) * half;
      o[3] = half; o[4] = half; o[5] = half;
      o[6] = (float)m.y;
    }
  }
}

__global__ void iota_kernel(int32_t* ids, int n, int first) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ids[i] = first + i;
}

// m = source tree; g = the tree that owns the targets' walk groups (g == m for an ordinary walk).
template <int B, bool EPS0>
int launch_walk(Impl* m, Impl* g, const BHParams& p, const float4* posm, const float4* tgt, float4* acc, int t0, int t1,
                bool accumulate, cudaStream_t s) {
  // 0 = one warp per walk group traverses and evaluates (default); 1 = warp-specialised (profiles/r2_walk_warp_specialised.md)
  static const int mode = getenv("NBODY_WALK") ? atoi(getenv("NBODY_WALK")) : 0;
  static const int split = getenv("NBODY_WS_SPLIT") ? atoi(getenv("NBODY_WS_SPLIT")) : 2;   // register split: 0 = 64/64, 1 = 40/88, 2 = 48/80
  auto ws = split == 1 ? bh_walk_ws_kernel<B, EPS0, (B <= 2 ? 1 : 0)> : split == 2 ? bh_walk_ws_kernel<B, EPS0, (B <= 2 ? 2 : 0)> : bh_walk_ws_kernel<B, EPS0, 0>;
  int per_sm = 0;
  if (mode == 1) NB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ws, kWsThreads, 0));
  else NB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bh_walk_group_kernel<B, EPS0>, kWalkThreads, 0));
  const int grid = sm_count() * std::max(1, std::min(per_sm, 8));
  const int64_t need = (int64_t)grid * (mode == 1 ? kWsPairs : kWalkWarps) * kStackCap;
  if (need > m->cap_stacks) {
    NB_CUDA(cudaStreamSynchronize(s));
    NB_TRY(realloc_dev(&m->stacks, (size_t)need));
    m->cap_stacks = need;
  }
  NB_CUDA(cudaMemsetAsync(&g->counters->next_group, 0, sizeof(int), s));
  if (mode == 1)
    ws<<<grid, kWsThreads, 0, s>>>(posm, m->node_com, m->node_meta, tgt, g->groups, g->counters, m->root, p.theta * p.theta, p.eps2, p.G, t0, t1,
                                   accumulate ? 1 : 0, m->stacks, acc, g->group_cost, g->bin_cost, g->n);
  else
    bh_walk_group_kernel<B, EPS0><<<grid, kWalkThreads, 0, s>>>(posm, m->node_com, m->node_meta, tgt, g->groups, g->counters, m->root,
                                                                p.theta * p.theta, p.eps2, p.G, t0, t1, accumulate ? 1 : 0,
                                                                m->stacks, acc, g->group_cost, g->bin_cost, g->n);
  return 0;
}

}  // namespace

void bh_reset(BHState& st, cudaStream_t s) {
  st.n_nodes_host = st.depth_host = st.n_groups_host = 0;
  st.root_mass_host = 0;
  for (int k = 0; k < 3; k++) st.root_com_host[k] = 0;
  if (st.impl) {
    Impl* m = static_cast<Impl*>(st.impl);
    if (m->root) cudaMemsetAsync(m->root, 0, 2 * sizeof(float4), s);
    m->n = 0;
  }
}

void bh_free(BHState& st) {
  if (!st.impl) return;
  Impl* m = static_cast<Impl*>(st.impl);
  for (int k = 0; k < 2; k++) { cudaFree(m->sort.keys[k]); cudaFree(m->sort.idx[k]); }
  cudaFree(m->sort.work);
  cudaFree(m->node_com); cudaFree(m->node_meta); cudaFree(m->node_range); cudaFree(m->node_ready); cudaFree(m->node_bmin); cudaFree(m->node_bmax); cudaFree(m->groups); cudaFree(m->group_cost);
  cudaFree(m->counters); cudaFree(m->root); cudaFree(m->stacks); cudaFree(m->boxes);
  cudaFree(m->samples); cudaFree(m->splitters); cudaFree(m->send_off); cudaFree(m->all_off); cudaFree(m->peer_pub); cudaFree(m->pub_node);
  cudaFree(m->visit); cudaFree(m->let_out); cudaFree(m->bin_cost); cudaFree(m->ret); if (m->h_counts) cudaFreeHost(m->h_counts); if (m->ev_counts) cudaEventDestroy(m->ev_counts); cudaFree(m->let_in); cudaFree(m->let_sorted); cudaFree(m->all_pos);
  delete m;
  st.impl = nullptr;
}

void bh_iota(int32_t* ids, int n, int first, cudaStream_t s) {
  if (n > 0) iota_kernel<<<(n + 255) / 256, 256, 0, s>>>(ids, n, first);
}

int bh_build(BHState& st, const BHParams& p, const float4* posm_in, const float4* vel_in, const int32_t* ids_in,
             float4* posm, float4* vel, int32_t* ids, int n, const uint32_t* box, cudaStream_t s, double* launches, const BodySegs* segs_in) {
  if (n <= 0) { set_error("Barnes-Hut: no bodies"); return -1; }
  if (n > (1 << 28)) { set_error("Barnes-Hut: at most 2^28 bodies per GPU"); return -1; }
  Impl* m = impl_of(st);
  NB_TRY(ensure(m, n, s));
  m->n = n;
  const unsigned nb = (unsigned)ceil_div(n, 256);
  const BodySegs segs = segs_in ? *segs_in : BodySegs{0, n, 0};
  if (!p.keep_root) {
    root_cube_kernel<<<1, 1, 0, s>>>(box, p.reference_root ? 1 : (p.sticky_root ? 2 : 0), m->root);
    *launches += 1;
  }
  // Sort only as many levels as the tree needs: the last known depth + 2 (a cell at the last sorted level is a leaf
  // whatever it holds, so a too-small hint costs accuracy nothing, only walk efficiency, and corrects itself through
  // the depth statistic). The parity configurations (one-body leaves / per-body walk) always sort all 63 bits.
  int levels = kMaxLevel;
  if (p.leaf_size > 1 && p.mac == kMacGroup && p.depth_hint > 0 && !getenv("NBODY_FULL_SORT")) levels = std::min(kMaxLevel, std::max(10, p.depth_hint + 2));
  const RadixPlan plan = radix_sort_begin(m->sort, n, 3 * kMaxLevel, s, 3 * (kMaxLevel - levels));
  const unsigned nbk = (unsigned)std::min<int64_t>(nb, (int64_t)sm_count() * 16);   // grid-stride: few histogram flushes per CTA
  morton_kernel<<<nbk, 256, 0, s>>>(posm_in, n, m->root, m->sort.keys[0], plan.ghist0, plan.shift0, segs);
  *launches += 1;
  m->sorted = radix_sort_run(m->sort, plan, n, s, launches);
  st.sort_passes_host = plan.passes - plan.p0;
  gather_bodies_kernel<<<nb, 256, 0, s>>>(m->sort.idx[m->sorted], n, segs, posm_in, vel_in, ids_in, posm, vel, ids);
  const uint64_t* keys = m->sort.keys[m->sorted];
  if (p.group_size != 32 && p.group_size != 64 && p.group_size != 128) { set_error("Barnes-Hut: group_size must be 32, 64 or 128"); return -1; }
  m->built_group_size = p.group_size;
  const int super = p.group_size * std::max(1, p.group_pack);
  tree_init_kernel<<<1, 1, 0, s>>>(keys, n, p.group_size, super, m->node_range, m->node_meta, m->node_ready, m->groups, m->counters);
  *launches += 2;
  {
    int leaf = std::max(1, p.leaf_size), gs = p.group_size, sup = super, lv = levels;
    void* args[] = {(void*)&keys, &leaf, &gs, &sup, &lv, &m->node_range, &m->node_meta, &m->node_ready, &m->groups, &m->counters};
    // as many CTAs as are co-resident (the split is latency bound: dependent key probes), 1 per SM when another
    // stream's walk shares the SMs
    static int split_ctas = 0;
    if (!split_ctas) {
      NB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&split_ctas, tree_split_kernel, 256, 0));
      split_ctas = std::max(1, std::min(split_ctas, 8));
    }
    // a small tree (the one over the received locally-essential points) has too few nodes per generation to occupy every
    // block, and each extra block makes the barrier slower: 2 per SM measured 146 us against 159 us at 6 for 250k points
    static const int small_ctas = getenv("NBODY_SPLIT_CTAS_SMALL") ? atoi(getenv("NBODY_SPLIT_CTAS_SMALL")) : 2;
    const int ctas = n < 400000 ? std::max(1, std::min(split_ctas, small_ctas)) : split_ctas;
    NB_CUDA(cudaLaunchCooperativeKernel((void*)tree_split_kernel, dim3(sm_count() * ctas), dim3(256), args, 0, s));
  }
  if (p.node_boxes) {
    if (m->cap_boxes_nodes < m->cap_nodes) {
      NB_CUDA(cudaStreamSynchronize(s));
      NB_TRY(realloc_dev(&m->node_bmin, (size_t)m->cap_nodes));
      NB_TRY(realloc_dev(&m->node_bmax, (size_t)m->cap_nodes));
      m->cap_boxes_nodes = m->cap_nodes;
    }
    monopole_kernel<true><<<sm_count() * 8, 256, 0, s>>>(posm, m->node_range, m->node_meta, m->node_ready, m->node_com, m->counters, m->root, m->node_bmin, m->node_bmax);
  } else {
    monopole_kernel<false><<<sm_count() * 8, 256, 0, s>>>(posm, m->node_range, m->node_meta, m->node_ready, m->node_com, m->counters, m->root, nullptr, nullptr);
  }
  *launches += 2;
  NB_CUDA(cudaGetLastError());
  return 0;
}

int bh_forces(BHState& st, const BHParams& p, const float4* posm, float4* acc, int n, int t0, int t1, cudaStream_t s,
              double* launches) {
  return bh_forces_from(st, st, p, posm, posm, acc, n, t0, t1, false, s, launches);
}

int bh_forces_from(BHState& src, BHState& tgt_tree, const BHParams& p, const float4* posm, const float4* tgt, float4* acc, int n,
                   int t0, int t1, bool accumulate, cudaStream_t s, double* launches) {
  Impl* m = impl_of(src);
  Impl* g = impl_of(tgt_tree);
  if (m->n != n || !m->counters || !g->counters) { set_error("Barnes-Hut: forces requested without a tree for these bodies"); return -5; }
  if (t1 <= t0) return 0;
  const bool eps0 = !(p.eps2 > 0.f);
  if (p.mac == kMacBody) {
    if (m != g || accumulate) { set_error("Barnes-Hut: the per-body walk (mac = 1) runs on a single tree"); return -1; }
    bh_walk_body_kernel<<<(unsigned)ceil_div(t1 - t0, 128), 128, 0, s>>>(posm, m->node_com, m->node_meta, m->counters, m->root,
                                                                          p.theta, p.eps2, p.G, t0, t1, acc);
  } else {
    if (g->built_group_size != p.group_size) { set_error("Barnes-Hut: group_size changed since the tree was built"); return -5; }
    const int B = p.group_size / 32;
    int rc = 0;
    if (eps0) rc = B == 1 ? launch_walk<1, true>(m, g, p, posm, tgt, acc, t0, t1, accumulate, s) : B == 2 ? launch_walk<2, true>(m, g, p, posm, tgt, acc, t0, t1, accumulate, s)
                                                                                      : launch_walk<4, true>(m, g, p, posm, tgt, acc, t0, t1, accumulate, s);
    else rc = B == 1 ? launch_walk<1, false>(m, g, p, posm, tgt, acc, t0, t1, accumulate, s) : B == 2 ? launch_walk<2, false>(m, g, p, posm, tgt, acc, t0, t1, accumulate, s)
                                                                                   : launch_walk<4, false>(m, g, p, posm, tgt, acc, t0, t1, accumulate, s);
    NB_TRY(rc);
  }
  *launches += 1;
  NB_CUDA(cudaGetLastError());
  return 0;
}

int bh_fetch_stats(BHState& st, cudaStream_t s, double* interactions) {
  if (!st.impl) return 0;
  Impl* m = static_cast<Impl*>(st.impl);
  if (!m->counters) return 0;
  Counters h;
  float4 root[2];
  float4 com0 = make_float4(0, 0, 0, 0);
  NB_CUDA(cudaMemcpyAsync(&h, m->counters, sizeof(h), cudaMemcpyDeviceToHost, s));
  NB_CUDA(cudaMemcpyAsync(root, m->root, sizeof(root), cudaMemcpyDeviceToHost, s));
  if (m->node_com) NB_CUDA(cudaMemcpyAsync(&com0, m->node_com, sizeof(com0), cudaMemcpyDeviceToHost, s));
  NB_CUDA(cudaStreamSynchronize(s));
  st.n_nodes_host = h.nnodes; st.depth_host = h.depth; st.n_groups_host = h.ngroups;
  st.root_com_host[0] = com0.x; st.root_com_host[1] = com0.y; st.root_com_host[2] = com0.z;
  st.root_mass_host = com0.w;
  st.root_cube_host[0] = root[0].x; st.root_cube_host[1] = root[0].y; st.root_cube_host[2] = root[0].z; st.root_cube_host[3] = root[0].w;
  if (interactions) *interactions = (double)h.interactions;
  if (h.overflow) { set_error("Barnes-Hut: a walk stack overflowed (leaf_size too large for kStackCap)"); return -5; }
  if ((int64_t)h.nnodes > m->cap_nodes) { set_error("Barnes-Hut: node pool exhausted"); return -5; }
  return 0;
}

int bh_leaf_boxes(BHState& st, const float4* posm, int n, float* boxes7, int64_t cap, int64_t* n_boxes, cudaStream_t s) {
  (void)posm;
  Impl* m = impl_of(st);
  if (m->n != n || n <= 0 || !m->counters) { set_error("Barnes-Hut: no tree has been built (call CreateOctree / Tick first)"); return -5; }
  const int64_t want = std::min<int64_t>(std::max<int64_t>(cap, 1), n);
  if (want * 7 + 1 > m->cap_boxes) {
    NB_CUDA(cudaStreamSynchronize(s));
    NB_TRY(realloc_dev(&m->boxes, (size_t)(want * 7 + 1)));
    m->cap_boxes = want * 7 + 1;
  }
  int* d_count = reinterpret_cast<int*>(m->boxes + want * 7);
  NB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), s));
  leaf_boxes_kernel<<<sm_count() * 4, 256, 0, s>>>(m->sort.keys[m->sorted], m->node_meta, m->counters, m->root, m->boxes,
                                                     (int)want, d_count);
  NB_CUDA(cudaGetLastError());
  int count = 0;
  NB_CUDA(cudaMemcpyAsync(&count, d_count, sizeof(int), cudaMemcpyDeviceToHost, s));
  NB_CUDA(cudaStreamSynchronize(s));
  const int64_t k = std::min<int64_t>(count, want);
  if (boxes7 && k > 0) NB_CUDA(cudaMemcpy(boxes7, m->boxes, (size_t)k * 7 * sizeof(float), cudaMemcpyDeviceToHost));
  *n_boxes = count;
  return 0;
}


// =====================================================================================================================
// K9 - multi-GPU Barnes-Hut: Morton domain split, body migration and locally-essential-tree (LET) exchange.
//
// Every rank owns a contiguous range of the GLOBAL Morton order: keys are taken in one root cube shared by all ranks
// (sticky: it only changes when bodies leave it or it becomes far too large, so keys and splitters stay comparable from
// step to step). splitters[r] = first key of rank r's domain.
//   when the bodies are set (bh_let_redistribute, eager): keys -> regular samples -> all-gather -> splitters (kept ones
//       are reused when the caller sets the bodies again) -> bodies bucketed by destination and exchanged.
//   per step:
//     (1) bodies that drifted out of the rank's key range are sent to their new owners (bh_let_redistribute again: one
//         bucket pass by destination + all-to-all-v; one host read for the counts), so no rank ever holds strays inside a
//         neighbour's domain - a few scattered strays would make the neighbours export everything around them;
//     (2) local sort + tree (bh_build);
//     (3) bh_let_plan: the rank publishes its BOUNDARY TREE (every cell whose parent holds more than n / 1024 bodies, with
//         the bounding box of its bodies; all-gathered); the local tree is descended for all peers at once with the
//         reference's acceptance rule taken against the peer's boundary tree (a published cell whose box is far enough
//         settles all its bodies, a published leaf that is too close opens the cell): accepted cells are exported as
//         point masses, opened leaves as bodies; the export counts are all-gathered and copied to pinned host memory;
//     (4) the local walk is enqueued; the host waits only for the counts (bh_let_plan_wait), then enqueues the exchange
//         of the export lists (bh_let_import, all-to-all-v), the tree over the received points and the walk through it;
//     (5) bh_let_finish, after the kick-drift: new splitters = equal-WORK quantiles (the walks add each group's
//         interaction count into 1024 bins along the sorted bodies; every rank contributes 256 equal-work samples),
//         damped by one half; they take effect in (1) of the next step.
// The reference has no counterpart (it is single threaded); forces equal the single-GPU walk up to the (stricter)
// acceptance of remote cells and summation order.
// =====================================================================================================================
namespace {

// A rank's message to the splitter selection: kLetSamples keys + its weight (float bits: work or body count) + its body count.
constexpr int kLetMsg = kLetSamples + 2;

// Regular positions of the (unsorted or sorted) local keys, weight = body count.
__global__ void let_sample_regular_kernel(const uint64_t* __restrict__ keys, const int n, uint64_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kLetSamples) {
    const int j = min((int)(((long long)i * n + n / 2) / kLetSamples), n - 1);
    out[i] = n > 0 ? keys[j] : ~0ull;
  } else if (i == kLetSamples) {
    out[i] = (uint64_t)__float_as_uint((float)n);
  } else if (i == kLetSamples + 1) {
    out[i] = (uint64_t)n;
  }
}

// Equal-WORK positions of the sorted local keys: bin_cost[b] = interactions evaluated for the bodies of bin b (equal-count
// bins along the sorted order). One CTA of kCostBins threads. Falls back to regular positions when nothing was counted.
__global__ void __launch_bounds__(kCostBins)
let_sample_cost_kernel(const uint64_t* __restrict__ keys, const int n, const uint32_t* __restrict__ bin_cost, uint64_t* __restrict__ out) {
  __shared__ float pre[kCostBins + 1];
  __shared__ float wsum[kCostBins / 32];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  float v = (float)bin_cost[t] + 1.0f;        // + 1: empty bins keep a little weight, positions stay strictly increasing
  float inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const float x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    float s = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float x = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += x; }
    wsum[lane] = s;
  }
  __syncthreads();
  pre[t + 1] = inc + (w ? wsum[w - 1] : 0.f);
  if (t == 0) pre[0] = 0.f;
  __syncthreads();
  const float total = pre[kCostBins];
  if (t < kLetSamples) {
    uint64_t key = ~0ull;
    if (n > 0) {
      const float target = total * ((float)t + 0.5f) / (float)kLetSamples;
      int lo = 0, hi = kCostBins;                     // first bin whose inclusive prefix reaches the target
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (pre[mid + 1] < target) lo = mid + 1; else hi = mid; }
      const int b = min(lo, kCostBins - 1);
      const float frac = fminf(fmaxf((target - pre[b]) / fmaxf(pre[b + 1] - pre[b], 1e-20f), 0.f), 1.f);
      const long long pos = (long long)(((double)b + (double)frac) * (double)n / (double)kCostBins);
      key = keys[min(max(pos, 0ll), (long long)n - 1)];
    }
    out[t] = key;
  } else if (t == kLetSamples) {
    out[t] = (uint64_t)__float_as_uint(n > 0 ? total : 0.f);
  } else if (t == kLetSamples + 1) {
    out[t] = (uint64_t)n;
  }
}

// One CTA: sort the gathered samples by key (bitonic, shared memory) and cut them into `world` pieces of equal WEIGHT: a
// sample of rank q stands for 1 / kLetSamples of that rank's weight. splitters[0] = 0, splitters[world] = ~0; the inner
// ones move a fraction `alpha` of the way from their previous value to the new quantile (alpha = 1: jump).
__global__ void __launch_bounds__(1024)
let_splitters_kernel(const uint64_t* __restrict__ msgs, const int world, const float alpha, uint64_t* __restrict__ splitters) {
  extern __shared__ uint64_t sk[];                 // [pow2] keys, then [pow2] floats (weights -> inclusive prefix)
  const int total = world * kLetSamples;
  int pow2 = 1;
  while (pow2 < total) pow2 <<= 1;
  float* sw = reinterpret_cast<float*>(sk + pow2);
  __shared__ float wq[kMaxWorld];
  if (threadIdx.x < world) wq[threadIdx.x] = __uint_as_float((uint32_t)msgs[(size_t)threadIdx.x * kLetMsg + kLetSamples]) / (float)kLetSamples;
  __syncthreads();
  for (int i = threadIdx.x; i < pow2; i += blockDim.x) {
    const int q = i / kLetSamples;
    sk[i] = i < total ? msgs[(size_t)q * kLetMsg + (i % kLetSamples)] : ~0ull;
    sw[i] = i < total ? wq[q] : 0.f;
  }
  __syncthreads();
  for (int k = 2; k <= pow2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < pow2; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const uint64_t a = sk[i], b = sk[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { sk[i] = b; sk[l] = a; const float t = sw[i]; sw[i] = sw[l]; sw[l] = t; }
        }
      }
      __syncthreads();
    }
  // inclusive prefix of the weights (Hillis-Steele; pow2 <= 4096 elements, 1024 threads)
  for (int off = 1; off < pow2; off <<= 1) {
    float v[4];
    int cnt = 0;
    for (int i = threadIdx.x; i < pow2; i += blockDim.x) v[cnt++] = i >= off ? sw[i - off] : 0.f;
    __syncthreads();
    cnt = 0;
    for (int i = threadIdx.x; i < pow2; i += blockDim.x) sw[i] += v[cnt++];
    __syncthreads();
  }
  const float all = sw[pow2 - 1];
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const float hi = sw[i], lo = i ? sw[i - 1] : 0.f;
    for (int r = 1; r < world; r++) {
      const float target = all * (float)r / (float)world;
      if (lo < target && target <= hi) {
        const uint64_t tgt = sk[i], old = splitters[r];
        uint64_t nw = tgt;
        if (alpha < 1.f) nw = tgt >= old ? old + (uint64_t)((double)(tgt - old) * (double)alpha) : old - (uint64_t)((double)(old - tgt) * (double)alpha);
        splitters[r] = nw;
      }
    }
  }
  if (threadIdx.x == 0) { splitters[0] = 0ull; splitters[world] = ~0ull; }
}

// keys[i] <- destination rank of body i = number of inner splitters <= key (bodies with equal keys stay together).
__global__ void __launch_bounds__(256)
let_dest_kernel(uint64_t* __restrict__ keys, const int n, const uint64_t* __restrict__ splitters, const int world) {
  __shared__ uint64_t sp[kMaxWorld];
  if (threadIdx.x < world - 1) sp[threadIdx.x] = splitters[threadIdx.x + 1];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t k = keys[i];
  int d = 0;
  for (int r = 0; r < world - 1; r++) d += sp[r] <= k ? 1 : 0;
  keys[i] = (uint64_t)d;
}

// send_off[r] = first position of the sorted values whose value is >= bound_r (bound_r = r for destination-sorted bodies,
// splitters[r] for key-sorted bodies); send_off[world] = n.
__global__ void let_offsets_kernel(const uint64_t* __restrict__ sorted, const int n, const int world, const uint64_t* __restrict__ splitters,
                                   int* __restrict__ send_off) {
  const int r = threadIdx.x;
  if (r > world) return;
  const uint64_t bound = splitters ? splitters[r] : (uint64_t)r;
  int lo = 0, hi = n;
  if (r == world) lo = n;
  else while (lo < hi) { const int mid = (lo + hi) >> 1; if (sorted[mid] < bound) lo = mid + 1; else hi = mid; }
  send_off[r] = lo;
}

// A rank describes its domain to its peers by the TOP OF ITS TREE ("boundary tree") with the bounding box of every
// published cell's bodies. Tree cells are spatially compact (a Morton range is not: it can jump across the whole cube).
// Layout of the message (kPubBytes per rank):
//   int header[32]: [0] = number of published cells
//   int2 child[kLetPub]: (first published child, number of children), (0, 0) for the leaves of the published tree
//   float4 box[kLetPub][2]: (min xyz, 0), (max xyz, 0) of the cell's bodies - two aligned 128-bit loads per cell
constexpr int kLetPub = 8192;
constexpr size_t kPubBytes = 32 * 4 + (size_t)kLetPub * 8 + (size_t)kLetPub * 32;
__host__ __device__ inline const int* pub_header(const void* msg) { return reinterpret_cast<const int*>(msg); }
__host__ __device__ inline const int2* pub_child(const void* msg) { return reinterpret_cast<const int2*>(reinterpret_cast<const char*>(msg) + 128); }
__host__ __device__ inline const float4* pub_box(const void* msg) { return reinterpret_cast<const float4*>(reinterpret_cast<const char*>(msg) + 128 + (size_t)kLetPub * 8); }

// One CTA, rounds of refinement: in round r every published leaf that is an inner cell of the tree with more than
// kPubMinBodies bodies and a bounding box wider than E_r = root width / 2^(r + 3) publishes its children (contiguous), as
// long as they fit. What counts is the EXTENT of a published leaf, not its body count: a peer must open everything within
// (cell size / theta) of the box, so a wide box around a few scattered bodies (the thin end of a Morton range) would make
// the peer export a large part of its domain. Boxes come from the tree (monopole pass).
constexpr int kPubMinBodies = 32;
__global__ void __launch_bounds__(1024)
let_publish_kernel(const int4* __restrict__ meta, const int2* __restrict__ range, const float4* __restrict__ bmin,
                   const float4* __restrict__ bmax, const float4* __restrict__ root, const int n, int* __restrict__ pub_node,
                   void* __restrict__ msg) {
  int* header = reinterpret_cast<int*>(msg);
  int2* child = reinterpret_cast<int2*>(reinterpret_cast<char*>(msg) + 128);
  float4* box = reinterpret_cast<float4*>(reinterpret_cast<char*>(msg) + 128 + (size_t)kLetPub * 8);
  __shared__ int s_count, s_valid, s_grew;
  if (threadIdx.x == 0) {
    s_count = n > 0 ? 1 : 0; s_valid = s_count; pub_node[0] = 0; child[0] = make_int2(0, 0);
    if (n > 0) { box[0] = bmin[0]; box[1] = bmax[0]; }
  }
  __syncthreads();
  float E = root[0].w * 0.25f;            // root width / 8
  for (int round = 0; round < 14; round++, E *= 0.5f) {
    // a round may need several sweeps: a cell published in this round can itself be wider than E
    for (int sweep = 0; sweep < kMaxLevel + 1; sweep++) {
      if (threadIdx.x == 0) s_grew = 0;
      __syncthreads();
      const int end = s_valid;
      for (int i = threadIdx.x; i < end; i += blockDim.x) {
        if (child[i].y != 0) continue;
        const int node = pub_node[i];
        const int4 m = meta[node];
        const int2 r = range[node];
        const float4 blo = box[2 * i], bhi = box[2 * i + 1];
        const float w = fmaxf(fmaxf(bhi.x - blo.x, bhi.y - blo.y), bhi.z - blo.z);
        if ((m.z & kLeafFlag) || r.y - r.x <= kPubMinBodies || !(w > E)) continue;
        const int slot = atomicAdd(&s_count, m.y);
        if (slot + m.y > kLetPub) { atomicSub(&s_count, m.y); continue; }     // no room: stays a leaf of the published tree
        for (int k = 0; k < m.y; k++) {
          const int c = m.x + k, j = slot + k;
          pub_node[j] = c;
          child[j] = make_int2(0, 0);
          box[2 * j] = bmin[c];
          box[2 * j + 1] = bmax[c];
        }
        child[i] = make_int2(slot, m.y);
        atomicMax(&s_valid, slot + m.y);
        s_grew = 1;
      }
      __syncthreads();
      if (!s_grew) break;
      __syncthreads();
    }
    if (s_valid > kLetPub - 8) break;
  }
  if (threadIdx.x == 0) { header[0] = s_valid; header[1] = 0; }
}

// Is the point mass (cm, cell half-width^2 / theta^2 = need) acceptable for EVERY body of a peer? Descends the peer's
// published tree: a published cell whose box is farther than the acceptance distance settles all bodies inside it; a
// published leaf that is too close settles the answer (no). The nearest child is visited first, so a "no" is found fast.
__device__ __forceinline__ bool let_accept_for_peer(const float4 cm, const float need, const void* __restrict__ msg) {
  const int npub = pub_header(msg)[0];
  if (npub <= 0) return true;                       // the peer holds no bodies
  const int2* child = pub_child(msg);
  const float4* box = pub_box(msg);
  auto dist2 = [&](const int b) {
    const float4 lo = box[2 * b], hi = box[2 * b + 1];
    const float dx = fmaxf(fmaxf(lo.x - cm.x, cm.x - hi.x), 0.f), dy = fmaxf(fmaxf(lo.y - cm.y, cm.y - hi.y), 0.f),
                dz = fmaxf(fmaxf(lo.z - cm.z, cm.z - hi.z), 0.f);
    return dx * dx + dy * dy + dz * dz;
  };
  if (dist2(0) > need) return true;
  int stack[96];
  int sp = 0;
  stack[sp++] = 0;
  while (sp > 0) {
    const int b = stack[--sp];
    const int2 ch = child[b];
    if (ch.y == 0) return false;                    // too close to a cell the peer does not describe any finer
    // children that are still too close; the nearest one goes on top (popped first). One child after the other on purpose:
    // the boxes hit L1 (81 %), so testing all eight at once (16 independent loads) only adds instructions to warps whose
    // lanes already diverge - measured 682 us against 303 us for this loop at 2M bodies per rank, 8 ranks.
    int nk = 0, idx[8], kmin = 0;
    float dmin = 3.0e38f;
    for (int k = 0; k < ch.y; k++) {
      const float d = dist2(ch.x + k);
      if (d > need) continue;
      if (d < dmin) { dmin = d; kmin = nk; }
      idx[nk++] = ch.x + k;
    }
    if (nk > 0) { const int t = idx[kmin]; idx[kmin] = idx[nk - 1]; idx[nk - 1] = t; }
    if (sp + nk > 96) return false;                 // cannot happen for a 21-level tree; stay conservative
    for (int k = 0; k < nk; k++) stack[sp++] = idx[k];
  }
  return true;
}

// The export descent, all peers at once, in one cooperative launch: a frontier of (cell, mask of the peers that still
// descend through it) per generation, so the work is proportional to the cells actually visited. frontier = 2 x 2 x cap
// words (cell, mask; current and next generation); fr_count[2] = their lengths (fr_count[0] = 1, frontier[0] = root before
// the launch).
__global__ void __launch_bounds__(256)
let_export_kernel(const float4* __restrict__ posm, const float4* __restrict__ node_com,
                  const int4* __restrict__ node_meta, const float4* __restrict__ root,
                  const char* __restrict__ peer_pub, const int world, const int rank, const float theta2,
                  uint32_t* __restrict__ frontier, const int cap, int* __restrict__ fr_count, float4* __restrict__ let_out,
                  int* __restrict__ let_cnt, const int cap_let) {
  const float root_half = root[0].w;
  const uint32_t all_peers = ((1u << world) - 1u) & ~(1u << rank);
  int wpad = 1;
  while (wpad < world) wpad <<= 1;
  // fr_count = three rotating frontier counters (a generation reads [gen % 3], appends to [(gen + 1) % 3] and clears
  // [(gen + 2) % 3], which nobody touches meanwhile) + the arrival counter and release word of the grid barrier: ONE
  // hand-written barrier per generation (cf. tree_split_kernel) instead of two cg::grid::sync().
  int* bar_count = fr_count + 32;    // each in a cache line of its own: the release word is polled
  int* bar_gen = fr_count + 64;
  for (int gen = 0; gen <= kMaxLevel + 1; gen++) {
    const int cur = gen & 1, nxt = cur ^ 1;
    const int ccur = gen % 3, cnxt = (gen + 1) % 3, cold = (gen + 2) % 3;
    const uint32_t* fnode = frontier + (size_t)cur * 2 * cap;
    const uint32_t* fmask = fnode + cap;
    uint32_t* nnode = frontier + (size_t)nxt * 2 * cap;
    uint32_t* nmask = nnode + cap;
    const int count = *((volatile int*)&fr_count[ccur]);
    if (count <= 0) break;
    if (blockIdx.x == 0 && threadIdx.x == 0) fr_count[cold] = 0;
    // one thread per (frontier cell, peer): wpad = world rounded up to a power of two lanes share a cell, so the descents
    // of one cell through its peers' boundary trees run side by side and are combined with shuffles
    const long long total = (long long)count * wpad;
    for (long long t0 = (long long)(blockIdx.x * blockDim.x + (threadIdx.x & ~31)); t0 < total; t0 += (long long)gridDim.x * blockDim.x) {
      const long long t = t0 + (threadIdx.x & 31);
      const int i = (int)(t / wpad), p = (int)(t % wpad);
      const bool live = t < total;
      int node = 0;
      uint32_t mask = 0;
      if (live) { node = (int)fnode[i]; mask = gen == 0 ? all_peers : fmask[i]; }
      const bool mine = live && p < world && (mask >> p & 1u);
      int4 m = make_int4(0, 0, 0, 0);
      bool leaf = false;
      uint32_t down = 0;
      if (live) { m = node_meta[node]; leaf = (m.z & kLeafFlag) != 0; }
      if (mine) {
        const float4 cm = node_com[node];
        const float size = root_half * __int_as_float((127 - (m.z & 255)) << 23);
        const float need = size * size / fmaxf(theta2, 1e-30f);      // accepted  <=>  distance^2 > need
        const bool single = leaf && m.y == 1;
        const bool accept = single || (theta2 > 0.f && let_accept_for_peer(cm, need, peer_pub + (size_t)p * kPubBytes));
        if (accept) {
          const int slot = atomicAdd(let_cnt + p, 1);
          if (slot < cap_let) let_out[(size_t)p * cap_let + slot] = cm;
        } else if (leaf) {
          const int slot = atomicAdd(let_cnt + p, m.y);
          for (int k = 0; k < m.y; k++) if (slot + k < cap_let) let_out[(size_t)p * cap_let + slot + k] = posm[m.x + k];
        } else {
          down = 1u << p;
        }
      }
      __syncwarp();
      for (int o = 1; o < wpad; o <<= 1) down |= __shfl_xor_sync(0xffffffffu, down, o);
      if (live && p == 0 && down) {
        const int slot = atomicAdd(&fr_count[cnxt], m.y);
        for (int k = 0; k < m.y; k++) if (slot + k < cap) { nnode[slot + k] = (uint32_t)(m.x + k); nmask[slot + k] = down; }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const int arrived = atomicAdd(bar_count, 1) + 1;
      if (arrived == (int)gridDim.x * (gen + 1)) {
        __threadfence();
        atomicExch(bar_gen, gen + 1);
      } else {
        while (*((volatile int*)bar_gen) < gen + 1) __nanosleep(100);
      }
      __threadfence();
    }
    __syncthreads();
  }
}

// Records of the bodies in owner order -> one staging array; out[i] = rec[order[i]] (rec_words floats each) and the id.
__global__ void __launch_bounds__(256)
let_pack_records_kernel(const uint32_t* __restrict__ order, const int n, const float* __restrict__ rec, const int rec_words,
                        const int32_t* __restrict__ ids, float* __restrict__ out, int32_t* __restrict__ ids_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t j = order[i];
  for (int k = 0; k < rec_words; k++) out[(size_t)i * rec_words + k] = rec[(size_t)j * rec_words + k];
  ids_out[i] = ids[j];
}
__global__ void __launch_bounds__(256)
let_owner_kernel(const int32_t* __restrict__ ids, const int n, const int64_t n_per, const int world, uint64_t* __restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = (uint64_t)min((int64_t)world - 1, (int64_t)ids[i] / n_per);
}
__global__ void __launch_bounds__(256)
let_unpack_records_kernel(const float* __restrict__ rec, const int32_t* __restrict__ ids, const int n, const int rec_words,
                          const int64_t first, const int n_slice, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t row = (int64_t)ids[i] - first;
  if (row < 0 || row >= n_slice) return;
  for (int k = 0; k < rec_words; k++) out[(size_t)row * rec_words + k] = rec[(size_t)i * rec_words + k];
}

int let_ensure(Impl* m, int world, int64_t cap_local, cudaStream_t s) {
  if (m->let_world != world) {
    NB_CUDA(cudaStreamSynchronize(s));
    NB_TRY(realloc_dev(&m->samples, (size_t)(world + 1) * kLetMsg));
    NB_TRY(realloc_dev(&m->splitters, (size_t)world + 1));
    NB_TRY(realloc_dev(&m->send_off, (size_t)2 * world + 2));
    NB_TRY(realloc_dev(&m->all_off, (size_t)world * (2 * world + 2)));
    NB_TRY(realloc_dev(&m->peer_pub, (size_t)world * kPubBytes));
    NB_TRY(realloc_dev(&m->pub_node, (size_t)kLetPub));
    NB_TRY(realloc_dev(&m->bin_cost, (size_t)kCostBins));
    if (m->h_counts) NB_CUDA(cudaFreeHost(m->h_counts));
    NB_CUDA(cudaHostAlloc((void**)&m->h_counts, (size_t)world * (2 * world + 2) * sizeof(int), cudaHostAllocDefault));
    if (!m->ev_counts) NB_CUDA(cudaEventCreateWithFlags(&m->ev_counts, cudaEventDisableTiming));
    NB_CUDA(cudaMemsetAsync(m->bin_cost, 0, kCostBins * sizeof(uint32_t), s));
    m->let_cnt = m->send_off + world + 1;      // the per-step count message: [world + 1] send_off | [world] export counts
    m->let_world = world;
    m->splitters_valid = false;
  }
  if (cap_local > m->cap_let) {
    NB_CUDA(cudaStreamSynchronize(s));
    NB_TRY(realloc_dev(&m->let_out, (size_t)world * (size_t)cap_local));
    m->cap_let = cap_local;
  }
  return 0;
}

int launch_splitters(Impl* m, int world, float alpha, cudaStream_t s) {
  int pow2 = 1;
  while (pow2 < world * kLetSamples) pow2 <<= 1;
  NB_CUDA(cudaFuncSetAttribute(let_splitters_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 12));   // world 9..16: 48 KB + static
  let_splitters_kernel<<<1, 1024, (size_t)pow2 * 12, s>>>(m->samples, world, alpha, m->splitters);
  NB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

void bh_let_forget_domains(BHState& st) { if (st.impl) static_cast<Impl*>(st.impl)->splitters_valid = false; }

// When the bodies are set: on return the first *n_local entries of posm_a / vel_a / ids_a hold the bodies of this rank's
// domain (`world` runs received from the peers, not yet sorted). posm_b / vel_b / ids_b are scratch (send staging).
// Splitters from an earlier run of this handle are reused (the caller uploads the same system again: the domains are
// already balanced for it); otherwise they are equal-count quantiles of regular key samples.
int bh_let_redistribute(BHState& st, Comm* comm, const BHParams& p, float4* posm_a, float4* vel_a, int32_t* ids_a, float4* posm_b,
                        float4* vel_b, int32_t* ids_b, int n, int64_t cap, const uint32_t* box_global, int* n_local, int* n_stay,
                        cudaStream_t s, double* launches) {
  Impl* m = impl_of(st);
  const int world = comm->world(), rank = comm->rank();
  if (world > kMaxWorld) { set_error("Barnes-Hut LET: at most 16 ranks"); return -1; }
  NB_TRY(ensure(m, (int)std::max<int64_t>(cap, 1), s));
  NB_TRY(let_ensure(m, world, cap, s));
  const unsigned nb = (unsigned)ceil_div(std::max(n, 1), 256);
  root_cube_kernel<<<1, 1, 0, s>>>(box_global, p.reference_root ? 1 : (p.sticky_root ? 2 : 0), m->root);
  if (n > 0) morton_kernel<<<nb, 256, 0, s>>>(posm_a, n, m->root, m->sort.keys[0], nullptr, 0);
  *launches += 2;
  if (!m->splitters_valid) {
    let_sample_regular_kernel<<<1, 288, 0, s>>>(m->sort.keys[0], n, m->samples + (size_t)world * kLetMsg);
    NB_TRY(comm->all_gather_bytes(m->samples + (size_t)world * kLetMsg, m->samples, (size_t)kLetMsg * 8, s));
    NB_TRY(launch_splitters(m, world, 1.f, s));
    *launches += 2;
    m->splitters_valid = true;
  }
  int sorted = 0;
  if (n > 0) {
    let_dest_kernel<<<nb, 256, 0, s>>>(m->sort.keys[0], n, m->splitters, world);
    sorted = radix_sort_pairs(m->sort, n, 8, s, launches);     // one pass: stable bucket by destination rank
    gather_bodies_kernel<<<nb, 256, 0, s>>>(m->sort.idx[sorted], n, BodySegs{0, n, 0}, posm_a, vel_a, ids_a, posm_b, vel_b, ids_b);
    *launches += 2;
  }
  let_offsets_kernel<<<1, 32, 0, s>>>(m->sort.keys[sorted], n, world, nullptr, m->send_off);
  *launches += 1;
  NB_TRY(comm->all_gather_bytes(m->send_off, m->all_off, (size_t)(world + 1) * 4, s));
  std::vector<int> off((size_t)world * (world + 1));
  NB_CUDA(cudaMemcpyAsync(off.data(), m->all_off, off.size() * 4, cudaMemcpyDeviceToHost, s));
  NB_CUDA(cudaStreamSynchronize(s));
  size_t sb[kMaxWorld], so[kMaxWorld], rb[kMaxWorld], ro[kMaxWorld];
  int64_t total = 0;
  for (int q = 0; q < world; q++) {
    const int* mine = off.data() + (size_t)rank * (world + 1);
    const int* theirs = off.data() + (size_t)q * (world + 1);
    sb[q] = (size_t)(mine[q + 1] - mine[q]); so[q] = (size_t)mine[q];
    rb[q] = (size_t)(theirs[rank + 1] - theirs[rank]); ro[q] = (size_t)total;
    total += (int64_t)rb[q];
  }
  // every rank sees every rank's offsets and the buffers have the same size everywhere, so all ranks take this exit
  // together (a rank-local return in front of the exchange would leave the others blocked in the collective)
  for (int d = 0; d < world; d++) {
    int64_t total_d = 0;
    for (int q = 0; q < world; q++) total_d += off[(size_t)q * (world + 1) + d + 1] - off[(size_t)q * (world + 1) + d];
    if (total_d > cap) {
      set_error("Barnes-Hut LET: the domain of rank " + std::to_string(d) + " outgrew the body buffers (" + std::to_string(total_d) + " > " +
                std::to_string(cap) + ")");
      m->splitters_valid = false;
      return -5;
    }
  }
  // destination-sorted bodies sit in *_b; every rank receives its new bodies into *_a (one grouped exchange)
  {
    const void* sendv[3] = {posm_b, vel_b, ids_b};
    void* recvv[3] = {posm_a, vel_a, ids_a};
    const size_t elem[3] = {16, 16, 4};
    NB_TRY(comm->all_to_all_v_multi(3, sendv, recvv, elem, sb, so, rb, ro, s));
  }
  *n_local = (int)total;
  *n_stay = (int)rb[rank];
  NB_CUDA(cudaGetLastError());
  return 0;
}

// Step phase (2): `local` holds the tree of the n sorted local bodies posm. Fills plan (host) after the step's one
// host synchronisation.
int bh_let_plan(BHState& local, Comm* comm, const BHParams& p, const float4* posm, int n, int64_t cap, LetPlan* plan, cudaStream_t s,
                double* launches) {
  Impl* m = impl_of(local);
  const int world = comm->world(), rank = comm->rank();
  NB_TRY(let_ensure(m, world, std::max<int64_t>(cap, 1024), s));
  const int64_t need_nodes = std::max<int64_t>(m->cap_nodes, 16);
  if (need_nodes > m->cap_visit) {
    NB_CUDA(cudaStreamSynchronize(s));
    NB_TRY(realloc_dev(&m->visit, (size_t)need_nodes * 4 + 96));
    m->cap_visit = need_nodes;
  }
  NB_CUDA(cudaMemsetAsync(m->send_off, 0, (size_t)(world + 1) * 4, s));
  // this rank's boundary tree: the top of its tree, refined until the published cells are small (or 8192 are used)
  char* my_pub = m->peer_pub + (size_t)rank * kPubBytes;
  if (!m->node_bmin) { set_error("Barnes-Hut LET: the local tree was built without node boxes"); return -5; }
  let_publish_kernel<<<1, 1024, 0, s>>>(m->node_meta, m->node_range, m->node_bmin, m->node_bmax, m->root, n, m->pub_node, my_pub);
  *launches += 1;
  NB_TRY(comm->all_gather_bytes(my_pub, m->peer_pub, kPubBytes, s));
  NB_CUDA(cudaMemsetAsync(m->let_cnt, 0, (size_t)world * 4, s));
  if (n > 0) {
    int w = world, r = rank, cap_let = (int)m->cap_let, cap_fr = (int)m->cap_visit;
    float theta2 = p.theta * p.theta;
    int* fr_count = reinterpret_cast<int*>(m->visit + (size_t)m->cap_visit * 4);
    static int init_count[96] = {1};   // frontier counters (the root is in), barrier words (lines of their own)
    NB_CUDA(cudaMemsetAsync(m->visit, 0, 4, s));                                    // frontier[0] = the root
    NB_CUDA(cudaMemcpyAsync(fr_count, init_count, sizeof(init_count), cudaMemcpyHostToDevice, s));
    void* args[] = {(void*)&posm, &m->node_com, &m->node_meta, &m->root, &m->peer_pub, &w, &r, &theta2, &m->visit, &cap_fr, &fr_count,
                    &m->let_out, &m->let_cnt, &cap_let};
    int per_sm = 1;
    NB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, let_export_kernel, 256, 0));
    static const int export_ctas = getenv("NBODY_EXPORT_CTAS") ? atoi(getenv("NBODY_EXPORT_CTAS")) : 4;   // measured: 396 / 350 / 329 us at 1 / 2 / 4 per SM
    NB_CUDA(cudaLaunchCooperativeKernel((void*)let_export_kernel, dim3(sm_count() * std::max(1, std::min(per_sm, export_ctas))), dim3(256), args, 0, s));
    *launches += 1;
  }
  // one message per rank: [world + 1] migration offsets | [world] export counts; one all-gather, one host read
  const int msg = 2 * world + 1;
  NB_TRY(comm->all_gather_bytes(m->send_off, m->all_off, (size_t)msg * 4, s));
  NB_CUDA(cudaMemcpyAsync(m->h_counts, m->all_off, (size_t)world * msg * 4, cudaMemcpyDeviceToHost, s));   // pinned: truly asynchronous
  NB_CUDA(cudaEventRecord(m->ev_counts, s));
  plan->world = world; plan->rank = rank; plan->n = n;
  return 0;
}

// Second half of the plan: wait for the counts only (work enqueued on the stream after bh_let_plan - the local walk - keeps
// running), then fill the host-side plan.
int bh_let_plan_wait(BHState& local, int64_t cap, LetPlan* plan) {
  Impl* m = impl_of(local);
  const int world = plan->world, rank = plan->rank;
  const int msg = 2 * world + 1;
  NB_CUDA(cudaEventSynchronize(m->ev_counts));
  const int* all = m->h_counts;
  if (getenv("NBODY_LET_DEBUG")) {   // development aid: size of every rank's boundary tree and this rank's export counts
    std::string line = "[let rank " + std::to_string(rank) + "] n=" + std::to_string(plan->n) + " published:";
    for (int q = 0; q < world; q++) {
      int hdr[2] = {0, 0};
      cudaMemcpy(hdr, m->peer_pub + (size_t)q * kPubBytes, 8, cudaMemcpyDeviceToHost);   // (waits for the whole device)
      line += " " + std::to_string(hdr[0]) + "/" + std::to_string(hdr[1]);
    }
    line += " export:";
    for (int q = 0; q < world; q++) line += " " + std::to_string(all[(size_t)rank * msg + world + 1 + q]);
    fprintf(stderr, "%s\n", line.c_str());
  }
  // cap_let is the same on every rank (it derives from the body capacity), so an overflow anywhere fails everywhere
  for (int q = 0; q < world; q++)
    for (int d = 0; d < world; d++)
      if (all[(size_t)q * msg + world + 1 + d] > m->cap_let) {
        set_error("Barnes-Hut LET: export list overflow (" + std::to_string(all[(size_t)q * msg + world + 1 + d]) + " > " + std::to_string(m->cap_let) + ")");
        return -5;
      }
  int64_t let_total = 0;
  for (int q = 0; q < world; q++) {
    plan->let_send[q] = all[(size_t)rank * msg + world + 1 + q];
    plan->let_recv[q] = all[(size_t)q * msg + world + 1 + rank];
    let_total += plan->let_recv[q];
  }
  plan->let_total = let_total;
  (void)cap;
  return 0;
}

// Step phase (3): exchange the export lists; `let` gets the tree over the points received (n_let of them; 0 is possible).
int bh_let_import(BHState& local, BHState& let, Comm* comm, const BHParams& p, const LetPlan& plan, const uint32_t* box_global, int* n_let,
                  cudaStream_t s, double* launches) {
  Impl* m = impl_of(local);
  const int world = plan.world;
  size_t sb[kMaxWorld], so[kMaxWorld], rb[kMaxWorld], ro[kMaxWorld];
  int64_t total = 0;
  for (int q = 0; q < world; q++) {
    sb[q] = (size_t)plan.let_send[q] * 16; so[q] = (size_t)q * (size_t)m->cap_let * 16;
    rb[q] = (size_t)plan.let_recv[q] * 16; ro[q] = (size_t)total * 16;
    total += plan.let_recv[q];
  }
  if (total > m->cap_let_in) {
    const int64_t c = total + total / 4 + 4096;
    NB_CUDA(cudaStreamSynchronize(s));
    NB_TRY(realloc_dev(&m->let_in, (size_t)c));
    NB_TRY(realloc_dev(&m->let_sorted, (size_t)c));
    m->cap_let_in = c;
  }
  NB_TRY(comm->all_to_all_v(m->let_out, sb, so, m->let_in, rb, ro, s));
  *n_let = (int)total;
  if (total > 0) {
    BHParams q = p;
    q.sticky_root = false;           // same cube as the local tree: copy it instead of deriving it again
    q.node_boxes = false;
    Impl* l = impl_of(let);
    NB_TRY(ensure(l, (int)total, s));
    NB_CUDA(cudaMemcpyAsync(l->root, m->root, sizeof(float4), cudaMemcpyDeviceToDevice, s));
    q.keep_root = true;
    NB_TRY(bh_build(let, q, m->let_in, nullptr, nullptr, m->let_sorted, nullptr, nullptr, (int)total, box_global, s, launches));
  }
  NB_CUDA(cudaGetLastError());
  return 0;
}

const float4* bh_let_sources(BHState& local) { return impl_of(local)->let_sorted; }

// Step phase (4), after the kick-drift: new splitters from equal-work samples (damped); they take effect when the bodies
// are sent to their domains at the start of the next step.
int bh_let_finish(BHState& local, Comm* comm, const BHParams& p, const LetPlan& plan, cudaStream_t s, double* launches) {
  Impl* m = impl_of(local);
  const int world = plan.world, n = plan.n;
  let_sample_cost_kernel<<<1, kCostBins, 0, s>>>(m->sort.keys[m->sorted], n, m->bin_cost, m->samples + (size_t)world * kLetMsg);
  NB_TRY(comm->all_gather_bytes(m->samples + (size_t)world * kLetMsg, m->samples, (size_t)kLetMsg * 8, s));
  NB_TRY(launch_splitters(m, world, p.let_damping, s));
  NB_CUDA(cudaMemsetAsync(m->bin_cost, 0, kCostBins * sizeof(uint32_t), s));
  *launches += 2;
  return 0;
}

// Read-back in the domain-split mode: every body's record travels to the rank that owns its ORIGINAL index slice
// [first, first + n_slice) (n_per bodies per rank), which receives its rows in order: out[(id - first) * rec_words ...].
// rec = n records of rec_words floats in the current local order. A collective.
int bh_let_return(BHState& local, Comm* comm, const int32_t* ids, int n, int64_t n_per, const float* rec, int rec_words, float* out,
                  int n_slice, int64_t first, cudaStream_t s, double* launches) {
  Impl* m = impl_of(local);
  const int world = comm->world(), rank = comm->rank();
  const int64_t cap = std::max<int64_t>(std::max<int64_t>(n, n_slice), 1);
  NB_TRY(ensure(m, (int)cap, s));
  NB_TRY(let_ensure(m, world, 1024, s));
  const size_t need = (size_t)cap * (size_t)(rec_words + 1) * 2;      // send staging + receive staging (records + ids)
  if ((int64_t)need > m->cap_ret) {
    NB_CUDA(cudaStreamSynchronize(s));
    NB_TRY(realloc_dev(&m->ret, need));
    m->cap_ret = (int64_t)need;
  }
  float* send_rec = m->ret;
  int32_t* send_ids = reinterpret_cast<int32_t*>(send_rec + (size_t)cap * rec_words);
  float* recv_rec = reinterpret_cast<float*>(send_ids + cap);
  int32_t* recv_ids = reinterpret_cast<int32_t*>(recv_rec + (size_t)cap * rec_words);
  const unsigned nb = (unsigned)ceil_div(std::max(n, 1), 256);
  int sorted = 0;
  if (n > 0) {
    let_owner_kernel<<<nb, 256, 0, s>>>(ids, n, n_per, world, m->sort.keys[0]);
    sorted = radix_sort_pairs(m->sort, n, 8, s, launches);
    let_pack_records_kernel<<<nb, 256, 0, s>>>(m->sort.idx[sorted], n, rec, rec_words, ids, send_rec, send_ids);
    *launches += 2;
  }
  let_offsets_kernel<<<1, 32, 0, s>>>(m->sort.keys[sorted], n, world, nullptr, m->send_off);
  NB_TRY(comm->all_gather_bytes(m->send_off, m->all_off, (size_t)(world + 1) * 4, s));
  std::vector<int> off((size_t)world * (world + 1));
  NB_CUDA(cudaMemcpyAsync(off.data(), m->all_off, off.size() * 4, cudaMemcpyDeviceToHost, s));
  NB_CUDA(cudaStreamSynchronize(s));
  size_t sb[kMaxWorld], so[kMaxWorld], rb[kMaxWorld], ro[kMaxWorld];
  int64_t total = 0;
  for (int q = 0; q < world; q++) {
    const int* mine = off.data() + (size_t)rank * (world + 1);
    const int* theirs = off.data() + (size_t)q * (world + 1);
    sb[q] = (size_t)(mine[q + 1] - mine[q]); so[q] = (size_t)mine[q];
    rb[q] = (size_t)(theirs[rank + 1] - theirs[rank]); ro[q] = (size_t)total;
    total += (int64_t)rb[q];
  }
  // all ranks see all offsets: a body count that does not add up fails everywhere, before the exchange
  for (int d = 0; d < world; d++) {
    int64_t t = 0;
    for (int q = 0; q < world; q++) t += off[(size_t)q * (world + 1) + d + 1] - off[(size_t)q * (world + 1) + d];
    const int64_t want = std::max<int64_t>(0, std::min<int64_t>(n_per, n_per * world - (int64_t)d * n_per));
    if (t > want) { set_error("Barnes-Hut LET: read-back found " + std::to_string(t) + " bodies for the slice of rank " + std::to_string(d)); return -5; }
  }
  size_t a[kMaxWorld], b[kMaxWorld], c2[kMaxWorld], d2[kMaxWorld];
  for (int q = 0; q < world; q++) { a[q] = sb[q] * rec_words * 4; b[q] = so[q] * rec_words * 4; c2[q] = rb[q] * rec_words * 4; d2[q] = ro[q] * rec_words * 4; }
  NB_TRY(comm->all_to_all_v(send_rec, a, b, recv_rec, c2, d2, s));
  for (int q = 0; q < world; q++) { a[q] = sb[q] * 4; b[q] = so[q] * 4; c2[q] = rb[q] * 4; d2[q] = ro[q] * 4; }
  NB_TRY(comm->all_to_all_v(send_ids, a, b, recv_ids, c2, d2, s));
  if (total > 0) {
    let_unpack_records_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, s>>>(recv_rec, recv_ids, (int)total, rec_words, first, n_slice, out);
    *launches += 1;
  }
  NB_CUDA(cudaGetLastError());
  return 0;
}

// Diagnostics: every rank's bodies gathered into one array (n_global float4) - the energy sum needs all sources.
int bh_let_gather_all(BHState& local, Comm* comm, const float4* posm, int n, int64_t n_global, const float4** out, int64_t* first,
                      cudaStream_t s) {
  Impl* m = impl_of(local);
  const int world = comm->world(), rank = comm->rank();
  NB_TRY(let_ensure(m, world, std::max<int64_t>(m->cap_n, 1024), s));
  if (n_global > m->cap_all_pos) {
    NB_CUDA(cudaStreamSynchronize(s));
    NB_TRY(realloc_dev(&m->all_pos, (size_t)n_global));
    m->cap_all_pos = n_global;
  }
  NB_CUDA(cudaMemcpyAsync(m->send_off, &n, 4, cudaMemcpyHostToDevice, s));
  NB_TRY(comm->all_gather_bytes(m->send_off, m->all_off, 4, s));
  std::vector<int> cnt((size_t)world);
  NB_CUDA(cudaMemcpyAsync(cnt.data(), m->all_off, (size_t)world * 4, cudaMemcpyDeviceToHost, s));
  NB_CUDA(cudaStreamSynchronize(s));
  size_t sb[kMaxWorld], so[kMaxWorld], rb[kMaxWorld], ro[kMaxWorld];
  int64_t total = 0;
  for (int q = 0; q < world; q++) {
    sb[q] = (size_t)n * 16; so[q] = 0;
    rb[q] = (size_t)cnt[(size_t)q] * 16; ro[q] = (size_t)total * 16;
    if (q == rank) *first = total;
    total += cnt[(size_t)q];
  }
  if (total != n_global) { set_error("Barnes-Hut LET: ranks hold " + std::to_string(total) + " bodies, expected " + std::to_string(n_global)); return -5; }
  NB_TRY(comm->all_to_all_v(posm, sb, so, m->all_pos, rb, ro, s));
  *out = m->all_pos;
  return 0;
}

// ---- inspection (tests / tools): tree read-back and the stand-alone sort --------------------------------------------
int bh_read_tree(BHState& st, float* com4, int32_t* meta4, int32_t* range2, uint64_t* keys, int64_t cap_nodes, int64_t cap_keys,
                 int64_t* n_nodes, cudaStream_t s) {
  Impl* m = impl_of(st);
  if (m->n <= 0 || !m->counters) { set_error("Barnes-Hut: no tree has been built (call CreateOctree / Tick first)"); return -5; }
  Counters h;
  NB_CUDA(cudaMemcpyAsync(&h, m->counters, sizeof(h), cudaMemcpyDeviceToHost, s));
  NB_CUDA(cudaStreamSynchronize(s));
  if (n_nodes) *n_nodes = h.nnodes;
  const size_t k = (size_t)std::min<int64_t>(h.nnodes, cap_nodes);
  if (com4 && k) NB_CUDA(cudaMemcpyAsync(com4, m->node_com, k * sizeof(float4), cudaMemcpyDeviceToHost, s));
  if (meta4 && k) NB_CUDA(cudaMemcpyAsync(meta4, m->node_meta, k * sizeof(int4), cudaMemcpyDeviceToHost, s));
  if (range2 && k) NB_CUDA(cudaMemcpyAsync(range2, m->node_range, k * sizeof(int2), cudaMemcpyDeviceToHost, s));
  const size_t kk = (size_t)std::min<int64_t>(m->n, cap_keys);
  if (keys && kk) NB_CUDA(cudaMemcpyAsync(keys, m->sort.keys[m->sorted], kk * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
  NB_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int sort_pairs_host(const uint64_t* keys_in, int64_t n, int key_bits, uint64_t* keys_out, uint32_t* idx_out, float* ms) {
  if (n <= 0 || n > (1ll << 30)) { set_error("sort: n out of range"); return -1; }
  RadixSortBuffers b;
  int rc = 0;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  auto body = [&]() -> int {
    for (int k = 0; k < 2; k++) { NB_TRY(realloc_dev(&b.keys[k], (size_t)n)); NB_TRY(realloc_dev(&b.idx[k], (size_t)n)); }
    b.work_words = radix_work_words(n);
    NB_TRY(realloc_dev(&b.work, b.work_words));
    NB_CUDA(cudaEventCreate(&e0));
    NB_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    int out = 0;
    for (int rep = 0; rep < (ms ? 3 : 1); rep++) {
      NB_CUDA(cudaMemcpy(b.keys[0], keys_in, (size_t)n * 8, cudaMemcpyHostToDevice));
      NB_CUDA(cudaEventRecord(e0, 0));
      out = radix_sort_pairs(b, (int)n, key_bits, 0, nullptr);
      NB_CUDA(cudaEventRecord(e1, 0));
      NB_CUDA(cudaEventSynchronize(e1));
      float t;
      NB_CUDA(cudaEventElapsedTime(&t, e0, e1));
      best = std::min(best, t);
    }
    NB_CUDA(cudaGetLastError());
    if (ms) *ms = best;
    if (keys_out) NB_CUDA(cudaMemcpy(keys_out, b.keys[out], (size_t)n * 8, cudaMemcpyDeviceToHost));
    if (idx_out) NB_CUDA(cudaMemcpy(idx_out, b.idx[out], (size_t)n * 4, cudaMemcpyDeviceToHost));
    return 0;
  };
  rc = body();
  for (int k = 0; k < 2; k++) { cudaFree(b.keys[k]); cudaFree(b.idx[k]); }
  cudaFree(b.work);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  return rc;
}

}  // namespace nbody
