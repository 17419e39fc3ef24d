// Host runtime + C ABI (include/nbody.h) of the B200-native N-body hot path.
//
// `nbody_sim` is the device-resident counterpart of the reference's simulation actor, class AOctreeSearch
// (/root/reference/Source/NBody/OctreeSearch.h:111-149, OctreeSearch.cpp:1-97): same verbs, but the bodies live
// in HBM as float4 SoA (posm = x,y,z,mass | vel | acc), every step is a short sequence of kernel launches on one
// stream with no host synchronisation in between, and the multi-GPU paths exchange data with NCCL.
//
// There is deliberately no CPU path in this file: if no sm_100 device is usable every entry point fails.
#include "../../include/nbody.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "bh.cuh"
#include "comm.h"
#include "common.cuh"
#include "direct_kernels.cuh"
#include "integrate.cuh"

namespace nbody {
static thread_local std::string g_err;
void set_error(const std::string& m) { g_err = m; }
}  // namespace nbody

using namespace nbody;

namespace {

int invalid(const std::string& m) { set_error(m); return NBODY_ERR_INVALID; }

template <class T>
int dev_reserve(T** p, int64_t* cap, int64_t need, cudaStream_t s) {
  if (need <= *cap && *p) return 0;
  if (*p) { NB_CUDA(cudaStreamSynchronize(s)); NB_CUDA(cudaFree(*p)); *p = nullptr; *cap = 0; }
  const int64_t want = std::max<int64_t>(need, 256);
  NB_CUDA(cudaMalloc((void**)p, (size_t)want * sizeof(T)));
  *cap = want;
  return 0;
}

struct DirectPlan {
  int variant = 0;  // 0 = packed I=8 (1 CTA/SM), 1 = packed I=2 (3 CTAs/SM)
  int i_per_thread = 8;
  int n_itiles = 0, jsplit = 1, chunk = 0;
  int64_t n_src_pad = 0, n_tgt_pad = 0;
};

DirectPlan plan_direct(int64_t n_tgt, int64_t n_src) {
  DirectPlan p;
  p.variant = n_tgt >= 32768 ? 0 : 1;
  p.i_per_thread = p.variant == 0 ? 8 : 2;
  const int itile = kDirectTPB * p.i_per_thread;
  p.n_itiles = (int)ceil_div(n_tgt, itile);
  // j-split: the grid (i-tiles x j-splits) runs in waves of `slots` CTAs; pick the split whose last wave is fullest and
  // whose source padding is smallest. Every split costs one float4 partial plane of HBM traffic per target, so among
  // splits within 0.3 % of the best the smallest wins (N = 1M on 148 SMs: 13 splits = 44.97 waves, not 16 = 55.35).
  const int slots = sm_count() * (p.variant == 0 ? 1 : 3);
  int best = 1;
  double best_cost = 1e30;
  for (int js = 1; js <= 64; js++) {
    const int64_t chunk = round_up(ceil_div(n_src, js), kDirectTJ);
    if (js > 1 && chunk < 512) break;
    const int64_t ctas = (int64_t)p.n_itiles * js;
    const double waves = (double)ctas / slots;
    const double cost = std::ceil(waves) / waves * ((double)chunk * js / (double)std::max<int64_t>(n_src, 1)) + 0.02 / std::max(1.0, std::floor(waves));
    if (cost < best_cost - 0.003) { best_cost = cost; best = js; }
  }
  p.jsplit = best;
  p.chunk = (int)round_up(ceil_div(n_src, best), kDirectTJ);
  p.n_src_pad = (int64_t)p.chunk * best;
  p.n_tgt_pad = round_up(n_tgt, 32);
  return p;
}

}  // namespace

struct nbody_sim {
  nbody_config cfg;
  bool initialized = false;
  bool show_octree = false;
  int64_t n_global = 0, n_local = 0, local_begin = 0, n_per = 0, slice_begin = 0;
  int64_t steps = 0;
  double launches = 0;

  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<cudaEvent_t> ev_pool;

  // state in HBM
  float4* d_posm = nullptr;  int64_t cap_posm = 0;   // gathered sources (direct) / local bodies (BH)
  float4* d_vel = nullptr;   int64_t cap_vel = 0;
  float4* d_acc = nullptr;   int64_t cap_acc = 0;
  float4* d_partial = nullptr; int64_t cap_partial = 0;
  int32_t* d_ids = nullptr;  int64_t cap_ids = 0;     // original index of body i (BH reorders bodies every step)
  float4* d_posm2 = nullptr; int64_t cap_posm2 = 0;   // BH: destination of the next Morton reordering (swapped each build)
  float4* d_vel2 = nullptr;  int64_t cap_vel2 = 0;
  int32_t* d_ids2 = nullptr; int64_t cap_ids2 = 0;
  uint8_t* d_stage = nullptr; int64_t cap_stage = 0;
  float4* d_acc2 = nullptr; int64_t cap_acc2 = 0;      // domain split: destination of the accelerations when the bodies are compacted
  float* d_ret = nullptr; int64_t cap_ret = 0;         // domain split: the rows of this rank's slice after a read-back
  uint32_t* d_box = nullptr;   // 8 words: absmax, min xyz, max xyz
  double* d_energy = nullptr;  // 2 doubles
  bool ids_identity = true;
  bool emulated = false;   // world > 1 without a communicator (all-zero NCCL id): slices only, exchanges are the caller's job

  Comm* comm = nullptr;
  DirectPlan plan;
  bool equal_mass = false;    // direct sum: every source has the same (finite, non-zero) mass -> 11-lane-op kernel
  float body_mass = 0.f;
  BHState tree;       // Barnes-Hut: tree over the bodies held by this rank
  BHState tree_let;   // multi-GPU LET mode: tree over the points received from the peers
  int n_let = 0;
  int n_migrated = 0;
  cudaEvent_t ev_ph[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // domain split: phase marks of the last step
  float ms_ph[5] = {0, 0, 0, 0, 0};
  bool drifted = false;   // domain split: the bodies moved since they were last sent to their domains
  BodySegs segs;      // domain split: where this rank's bodies sit in d_posm / d_vel / d_ids before the next build

  // timing of the last call
  float ms_call = 0, ms_force = 0, ms_build = 0, ms_integrate = 0, ms_comm = 0;
  double interactions = 0;
  float cube_size = 0;

  // Direct sum: posm holds all N sources (this rank's slice at local_begin), vel / acc hold the local slice only.
  // Barnes-Hut: every array holds all N bodies in Morton order; this rank integrates the slice at local_begin.
  bool bh() const { return cfg.method == NBODY_BARNES_HUT; }
  // multi-GPU Barnes-Hut with domain decomposition: the arrays hold only the bodies this rank owns (n_local varies)
  int exchange = 1;   // effective multi-GPU Barnes-Hut mode (cfg.bh_exchange, or chosen by size when it is -1)
  bool let_mode() const { return bh() && comm != nullptr && exchange == 0; }
  float4* posm_local() { return d_posm + local_begin; }
  float4* vel_local() { return d_vel + (bh() ? local_begin : 0); }
  float4* acc_local() { return d_acc + (bh() ? local_begin : 0); }
  int32_t* ids_local() { return d_ids + (bh() ? local_begin : 0); }
  // which bodies a set_* call uploads to this rank, and where they land
  int64_t load_begin() const { return bh() && !let_mode() ? 0 : slice_begin; }
  int64_t load_count() const { return bh() && !let_mode() ? n_global : n_local; }
  float4* posm_load() { return d_posm + (bh() ? 0 : slice_begin); }
  // original index of local body 0 while the bodies still sit in the order they were set in (ids_identity): in the
  // domain-split mode the arrays start at 0 but hold the global slice [slice_begin, slice_begin + n_local)
  int64_t identity_begin() const { return let_mode() ? slice_begin : local_begin; }
};

namespace {

int check_device(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    set_error(std::string("CUDA: no usable device (") + cudaGetErrorString(e) + "); this library has no CPU fallback");
    cudaGetLastError();
    return NBODY_ERR_CUDA;
  }
  if (device < 0 || device >= count) return invalid("device ordinal out of range");
  cudaDeviceProp prop;
  NB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("CUDA: device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
              "; kernels are built for sm_100a only");
    return NBODY_ERR_CUDA;
  }
  NB_CUDA(cudaSetDevice(device));
  return 0;
}

cudaEvent_t pool_event(nbody_sim* s, size_t k) {
  while (s->ev_pool.size() <= k) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    s->ev_pool.push_back(e);
  }
  return s->ev_pool[k];
}

// Split n bodies over the ranks: contiguous slices of n_per = ceil(n / world) (direct sum: fixed for the run).
void partition(nbody_sim* s, int64_t n) {
  s->n_global = n;
  // measured (profiles/): at 16M bodies the domain split wins on 4 and 8 GPUs (138 vs 117, 167 vs 159 steps/s), at 2M
  // bodies on 2 GPUs the replicated tree wins (537 vs 432): switch at 8M
  s->exchange = s->cfg.bh_exchange >= 0 ? s->cfg.bh_exchange : (n > ((int64_t)1 << 23) ? 0 : 1);
  s->n_per = ceil_div(n, s->cfg.world);
  s->slice_begin = std::min<int64_t>(n, (int64_t)s->cfg.rank * s->n_per);
  s->n_local = std::min<int64_t>(s->n_per, n - s->slice_begin);
  s->local_begin = s->let_mode() ? 0 : s->slice_begin;   // offset of the rank's bodies inside the device arrays
}

int reserve_state(nbody_sim* s) {
  const int64_t n = s->n_global;
  if (s->cfg.method == NBODY_DIRECT) {
    s->plan = plan_direct(std::max<int64_t>(s->n_local, 1), s->n_per * s->cfg.world);
    const int64_t need_src = std::max<int64_t>(s->plan.n_src_pad, s->n_per * s->cfg.world);
    NB_TRY(dev_reserve(&s->d_posm, &s->cap_posm, need_src, s->stream));
    NB_TRY(dev_reserve(&s->d_partial, &s->cap_partial, (int64_t)s->plan.jsplit * s->plan.n_tgt_pad, s->stream));
    NB_TRY(dev_reserve(&s->d_vel, &s->cap_vel, std::max<int64_t>(s->n_local, 1), s->stream));
    NB_TRY(dev_reserve(&s->d_acc, &s->cap_acc, std::max<int64_t>(s->n_local, 1), s->stream));
    NB_TRY(dev_reserve(&s->d_ids, &s->cap_ids, std::max<int64_t>(s->n_local, 1), s->stream));
  } else {
    // replicated: all bodies, padded so the slices all-gather in place; LET: the rank's domain with room to grow
    const int64_t total = s->let_mode() ? 2 * s->n_per + 4096 : std::max<int64_t>(s->n_per * s->cfg.world, 1);
    NB_TRY(dev_reserve(&s->d_posm, &s->cap_posm, total, s->stream));
    NB_TRY(dev_reserve(&s->d_vel, &s->cap_vel, total, s->stream));
    NB_TRY(dev_reserve(&s->d_acc, &s->cap_acc, total, s->stream));
    NB_TRY(dev_reserve(&s->d_ids, &s->cap_ids, total, s->stream));
    NB_TRY(dev_reserve(&s->d_posm2, &s->cap_posm2, total, s->stream));
    NB_TRY(dev_reserve(&s->d_vel2, &s->cap_vel2, total, s->stream));
    NB_TRY(dev_reserve(&s->d_ids2, &s->cap_ids2, total, s->stream));
  }
  (void)n;
  return 0;
}

// After the local slice of posm is in place: pad entries = zero-mass bodies at the origin, then make every rank's
// slice visible everywhere (direct sum).
int publish_positions(nbody_sim* s) {
  if (s->cfg.method != NBODY_DIRECT) return 0;
  const int64_t total = std::max<int64_t>(s->plan.n_src_pad, s->n_per * s->cfg.world);
  // zero the tail of this rank's slot and everything past the real bodies
  const int64_t slot_end = s->local_begin + s->n_per;
  const int64_t pad0 = s->local_begin + s->n_local;
  if (slot_end > pad0) {
    fill_float4_kernel<<<(unsigned)ceil_div(slot_end - pad0, 256), 256, 0, s->stream>>>(s->d_posm + pad0, slot_end - pad0, make_float4(kPadCoord, kPadCoord, kPadCoord, 0.f));
    s->launches++;
  }
  const int64_t gathered = s->n_per * s->cfg.world;
  if (total > gathered) {
    fill_float4_kernel<<<(unsigned)ceil_div(total - gathered, 256), 256, 0, s->stream>>>(s->d_posm + gathered, total - gathered, make_float4(kPadCoord, kPadCoord, kPadCoord, 0.f));
    s->launches++;
  }
  NB_CUDA(cudaGetLastError());
  if (s->comm) NB_TRY(s->comm->all_gather_f32_inplace(reinterpret_cast<float*>(s->d_posm), (size_t)s->n_per * 4, s->stream));
  return 0;
}

template <int V, bool E0, bool EQM>
void launch_direct_variant(const DirectPlan& p, const float4* src, const float4* tgt, int n_tgt, float eps2, float4* partial, cudaStream_t st) {
  dim3 grid(p.n_itiles, p.jsplit);
  if (V == 0) direct_packed_kernel<8, E0, 1, kDirectTPB, EQM><<<grid, kDirectTPB, 0, st>>>(src, p.chunk, tgt, n_tgt, eps2, partial, (int)p.n_tgt_pad);
  else direct_packed_kernel<2, E0, 3, kDirectTPB, EQM><<<grid, kDirectTPB, 0, st>>>(src, p.chunk, tgt, n_tgt, eps2, partial, (int)p.n_tgt_pad);
}

int launch_direct(nbody_sim* s) {
  if (s->n_local <= 0) return 0;
  const float eps2 = s->cfg.eps * s->cfg.eps;
  const float4* tgt = s->posm_local();
  const bool e0 = !(eps2 > 0.f);
  const int n = (int)s->n_local;
  const int sel = (s->plan.variant == 0 ? 0 : 4) + (e0 ? 2 : 0) + (s->equal_mass ? 1 : 0);
  switch (sel) {
    case 0: launch_direct_variant<0, false, false>(s->plan, s->d_posm, tgt, n, eps2, s->d_partial, s->stream); break;
    case 1: launch_direct_variant<0, false, true>(s->plan, s->d_posm, tgt, n, eps2, s->d_partial, s->stream); break;
    case 2: launch_direct_variant<0, true, false>(s->plan, s->d_posm, tgt, n, eps2, s->d_partial, s->stream); break;
    case 3: launch_direct_variant<0, true, true>(s->plan, s->d_posm, tgt, n, eps2, s->d_partial, s->stream); break;
    case 4: launch_direct_variant<1, false, false>(s->plan, s->d_posm, tgt, n, eps2, s->d_partial, s->stream); break;
    case 5: launch_direct_variant<1, false, true>(s->plan, s->d_posm, tgt, n, eps2, s->d_partial, s->stream); break;
    case 6: launch_direct_variant<1, true, false>(s->plan, s->d_posm, tgt, n, eps2, s->d_partial, s->stream); break;
    default: launch_direct_variant<1, true, true>(s->plan, s->d_posm, tgt, n, eps2, s->d_partial, s->stream); break;
  }
  s->launches++;
  NB_CUDA(cudaGetLastError());
  s->interactions = (double)s->n_local * (double)s->n_global;
  return 0;
}

// Direct sum: do all N sources carry one mass? (Reduced over the gathered source array, so every rank decides alike.)
int detect_equal_mass(nbody_sim* s) {
  s->equal_mass = false;
  s->body_mass = 0.f;
  if (s->cfg.method != NBODY_DIRECT || s->n_global <= 0) return 0;
  if (getenv("NBODY_NO_EQUAL_MASS")) return 0;   // development knob: always the general kernel
  static const uint32_t init[2] = {~0u, 0u};
  uint32_t* d_mm = s->d_box;   // scratch: 2 words (min, max of the order-preserving mass keys)
  NB_CUDA(cudaMemcpyAsync(d_mm, init, sizeof(init), cudaMemcpyHostToDevice, s->stream));
  const int blocks = (int)std::min<int64_t>(ceil_div(s->n_global, 256), sm_count() * 8);
  mass_range_kernel<<<blocks, 256, 0, s->stream>>>(s->d_posm, s->n_global, d_mm);
  s->launches++;
  NB_CUDA(cudaGetLastError());
  uint32_t h[2];
  NB_CUDA(cudaMemcpyAsync(h, d_mm, sizeof(h), cudaMemcpyDeviceToHost, s->stream));
  NB_CUDA(cudaStreamSynchronize(s->stream));
  const float lo = ordered_to_float(h[0]), hi = ordered_to_float(h[1]);
  if (h[0] == h[1] && std::isfinite(lo) && lo != 0.f) { s->equal_mass = true; s->body_mass = lo; }
  (void)hi;
  return 0;
}

int launch_cube_size(nbody_sim* s) {
  static const uint32_t init[8] = {0u, ~0u, ~0u, ~0u, 0u, 0u, 0u, 0u};
  NB_CUDA(cudaMemcpyAsync(s->d_box, init, sizeof(init), cudaMemcpyHostToDevice, s->stream));
  // replicated Barnes-Hut keeps all N bodies on every rank: reduce over them locally, no collective
  const bool all_here = s->bh() && !s->let_mode();
  // domain split after a migration: the bodies that stayed and the ones that arrived are two stretches of the arrays
  const bool two = s->let_mode() && s->segs.n0 < (int)s->n_local;
  const float4* src[2] = {all_here ? s->d_posm : s->let_mode() ? s->d_posm + s->segs.b0 : s->posm_local(), s->d_posm + s->segs.b1};
  const int64_t cnt[2] = {all_here ? s->n_global : two ? s->segs.n0 : s->n_local, two ? s->n_local - s->segs.n0 : 0};
  for (int k = 0; k < 2; k++) {
    if (cnt[k] <= 0) continue;
    const int blocks = (int)std::min<int64_t>(ceil_div(cnt[k], 256), sm_count() * 8);
    cube_size_kernel<<<blocks, 256, 0, s->stream>>>(src[k], (int)cnt[k], s->d_box);
    s->launches++;
    NB_CUDA(cudaGetLastError());
  }
  if (s->comm && !all_here) {
    NB_TRY(s->comm->all_reduce_u32_max(s->d_box, 1, s->stream));
    NB_TRY(s->comm->all_reduce_u32_min(s->d_box + 1, 3, s->stream));
    NB_TRY(s->comm->all_reduce_u32_max(s->d_box + 4, 3, s->stream));
  }
  return 0;
}

// One force evaluation (+ integration when dt > 0) enqueued on the stream. ev: optional 6 events
// (0 start, 1 after build, 5 after the pre-force exchange, 2 after force, 3 after integrate, 4 after the post-step exchange).
int enqueue_step(nbody_sim* s, float dt, bool integrate, cudaEvent_t* ev) {
  if (ev) NB_CUDA(cudaEventRecord(ev[0], s->stream));
  if (s->cfg.method == NBODY_DIRECT) {
    if (ev) { NB_CUDA(cudaEventRecord(ev[1], s->stream)); NB_CUDA(cudaEventRecord(ev[5], s->stream)); }
    NB_TRY(launch_direct(s));
    if (ev) NB_CUDA(cudaEventRecord(ev[2], s->stream));
    if (s->n_local > 0) {
      const unsigned blocks = (unsigned)ceil_div(s->n_local, 256);
      const float gscale = s->equal_mass ? s->cfg.G * s->body_mass : s->cfg.G;   // equal-mass kernel leaves the mass to K2
      if (integrate)
        reduce_kick_drift_kernel<true><<<blocks, 256, 0, s->stream>>>(s->d_partial, s->plan.jsplit, s->plan.n_tgt_pad, (int)s->n_local, gscale, dt, s->posm_local(), s->d_vel, s->d_acc);
      else
        reduce_kick_drift_kernel<false><<<blocks, 256, 0, s->stream>>>(s->d_partial, s->plan.jsplit, s->plan.n_tgt_pad, (int)s->n_local, gscale, dt, s->posm_local(), s->d_vel, s->d_acc);
      s->launches++;
      NB_CUDA(cudaGetLastError());
    }
    if (ev) NB_CUDA(cudaEventRecord(ev[3], s->stream));
    if (integrate && s->comm) NB_TRY(s->comm->all_gather_f32_inplace(reinterpret_cast<float*>(s->d_posm), (size_t)s->n_per * 4, s->stream));
    if (ev) NB_CUDA(cudaEventRecord(ev[4], s->stream));
  } else {
    NB_TRY(launch_cube_size(s));   // OctreeSearch.cpp:26
    BHParams bp;
    bp.G = s->cfg.G; bp.eps2 = s->cfg.eps * s->cfg.eps; bp.theta = s->cfg.theta;
    bp.leaf_size = std::max(1, s->cfg.leaf_size); bp.reference_root = s->cfg.reference_root != 0;
    bp.mac = s->cfg.mac;
    bp.group_size = s->cfg.group_size;
    bp.group_pack = s->cfg.group_pack;
    bp.depth_hint = s->tree.depth_host;
    if (const char* dm = getenv("NBODY_LET_DAMPING")) bp.let_damping = std::min(1.f, std::max(0.f, (float)atof(dm)));   // development knob
    double launches = 0;
    if (s->let_mode()) {
      // (1) local sort + tree, (2) plan: migration ranges, domain boxes, export descent, count exchange (the step's one host
      // synchronisation), (3) LET exchange + tree beside the local walk, then the walk over the received points,
      // (4) kick-drift, new splitters, lazy migration; see bh.cu K9
      // NBODY_LET_TRACE=1: synchronise after every phase and print host-clock phase times (development aid)
      // NBODY_LET_TRACE=2: no synchronisation, host-clock time at which each phase has been ENQUEUED (is the host ahead of the GPU?)
      static const int trace = getenv("NBODY_LET_TRACE") ? std::max(1, atoi(getenv("NBODY_LET_TRACE"))) : 0;
      auto t_prev = std::chrono::steady_clock::now();
      auto lap = [&](const char* what) {
        if (!trace) return;
        if (trace == 1) cudaStreamSynchronize(s->stream);
        else if (s->steps % 50 != 7) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[let rank %d step %lld] %-14s %8.3f ms  n_local=%lld n_let=%d\n", s->cfg.rank, (long long)s->steps, what,
                std::chrono::duration<double, std::milli>(now - t_prev).count(), (long long)s->n_local, s->n_let);
        t_prev = now;
      };
      bp.sticky_root = true;
      bp.node_boxes = true;
      lap("cube");
      if (!s->ev_ph[0]) for (int q = 0; q < 8; q++) NB_CUDA(cudaEventCreate(&s->ev_ph[q]));
      auto phase = [&](int k) { if (ev) cudaEventRecord(s->ev_ph[k], s->stream); };
      phase(0);
      if (s->drifted) {
        // bodies that crossed into another rank's key range move there BEFORE the tree is built: a rank that kept such
        // strays for a step would have to describe a handful of bodies scattered through its neighbours' domains, and the
        // neighbours would export everything around them (measured: 6x larger imports). The splitters were updated at the
        // end of the last step (equal-work quantiles, damped).
        int n_new = 0, n_stay = 0;
        NB_TRY(bh_let_redistribute(s->tree, s->comm, bp, s->d_posm, s->d_vel, s->d_ids, s->d_posm2, s->d_vel2, s->d_ids2, (int)s->n_local,
                                   std::min(std::min(s->cap_posm, s->cap_posm2), s->cap_acc), s->d_box, &n_new, &n_stay, s->stream, &launches));
        s->n_migrated = n_new - n_stay;
        s->n_local = n_new;
        s->segs = BodySegs{0, n_new, 0};
        s->drifted = false;
        lap("migrate");
      }
      phase(1);
      if (s->n_local <= 0) { set_error("Barnes-Hut domain split: a rank holds no bodies"); return NBODY_ERR_STATE; }
      NB_TRY(bh_build(s->tree, bp, s->d_posm, s->d_vel, s->d_ids, s->d_posm2, s->d_vel2, s->d_ids2, (int)s->n_local, s->d_box,
                      s->stream, &launches, &s->segs));
      std::swap(s->d_posm, s->d_posm2); std::swap(s->cap_posm, s->cap_posm2);
      std::swap(s->d_vel, s->d_vel2);   std::swap(s->cap_vel, s->cap_vel2);
      std::swap(s->d_ids, s->d_ids2);   std::swap(s->cap_ids, s->cap_ids2);
      s->segs = BodySegs{0, (int)s->n_local, 0};
      s->ids_identity = false;
      lap("local build");
      if (ev) { NB_CUDA(cudaEventRecord(ev[1], s->stream)); NB_CUDA(cudaEventRecord(ev[5], s->stream)); }
      // One stream, full occupancy: publish + export + count exchange, then the local walk; the host waits only for the
      // counts (they sit in front of the walk), and enqueues the exchange of the export lists, the tree over the received
      // points and the walk through it behind the local walk.
      LetPlan plan;
      const int64_t cap = std::min(std::min(s->cap_posm, s->cap_posm2), s->cap_acc);
      phase(2);
      NB_TRY(bh_let_plan(s->tree, s->comm, bp, s->d_posm, (int)s->n_local, cap, &plan, s->stream, &launches));
      lap("plan + export");
      phase(3);
      NB_TRY(bh_forces(s->tree, bp, s->d_posm, s->d_acc, (int)s->n_local, 0, (int)s->n_local, s->stream, &launches));
      phase(4);
      if (trace == 2) lap("walk enqueued");
      NB_TRY(bh_let_plan_wait(s->tree, cap, &plan));
      lap("walk local");
      NB_TRY(bh_let_import(s->tree, s->tree_let, s->comm, bp, plan, s->d_box, &s->n_let, s->stream, &launches));
      lap("let import");
      phase(5);
      if (s->n_let > 0)
        NB_TRY(bh_forces_from(s->tree_let, s->tree, bp, bh_let_sources(s->tree), s->d_posm, s->d_acc, s->n_let, 0, (int)s->n_local,
                              true, s->stream, &launches));
      lap("walk let");
      phase(6);
      if (ev) NB_CUDA(cudaEventRecord(ev[2], s->stream));
      if (integrate) {
        kick_drift_kernel<<<(unsigned)ceil_div(s->n_local, 256), 256, 0, s->stream>>>((int)s->n_local, dt, s->d_posm, s->d_vel, s->d_acc);
        s->launches++;
        NB_CUDA(cudaGetLastError());
      }
      if (ev) NB_CUDA(cudaEventRecord(ev[3], s->stream));
      if (integrate) {
        // the walks recorded their work along the sorted bodies: new splitters (equal-work quantiles, damped) for the next step
        NB_TRY(bh_let_finish(s->tree, s->comm, bp, plan, s->stream, &launches));
        s->drifted = true;
      }
      s->launches += launches;
      if (ev) NB_CUDA(cudaEventRecord(ev[4], s->stream));
      if (integrate) s->steps++;
      return 0;
    }
    // Morton reordering of ALL bodies (identical on every rank: same data, stable sort), tree + monopoles
    NB_TRY(bh_build(s->tree, bp, s->d_posm, s->d_vel, s->d_ids, s->d_posm2, s->d_vel2, s->d_ids2, (int)s->n_global, s->d_box,
                    s->stream, &launches));
    std::swap(s->d_posm, s->d_posm2); std::swap(s->cap_posm, s->cap_posm2);
    std::swap(s->d_vel, s->d_vel2);   std::swap(s->cap_vel, s->cap_vel2);
    std::swap(s->d_ids, s->d_ids2);   std::swap(s->cap_ids, s->cap_ids2);
    s->ids_identity = false;
    if (ev) { NB_CUDA(cudaEventRecord(ev[1], s->stream)); NB_CUDA(cudaEventRecord(ev[5], s->stream)); }
    // this rank walks and integrates its slice of the Morton order (a compact spatial domain)
    const int t0 = (int)s->local_begin, t1 = (int)(s->local_begin + s->n_local);
    NB_TRY(bh_forces(s->tree, bp, s->d_posm, s->d_acc, (int)s->n_global, t0, t1, s->stream, &launches));
    s->launches += launches;
    if (ev) NB_CUDA(cudaEventRecord(ev[2], s->stream));
    if (integrate && s->n_local > 0) {
      kick_drift_kernel<<<(unsigned)ceil_div(s->n_local, 256), 256, 0, s->stream>>>((int)s->n_local, dt, s->posm_local(), s->vel_local(), s->acc_local());
      s->launches++;
      NB_CUDA(cudaGetLastError());
    }
    if (ev) NB_CUDA(cudaEventRecord(ev[3], s->stream));
    if (integrate && s->comm) {
      NB_TRY(s->comm->all_gather_f32_inplace(reinterpret_cast<float*>(s->d_posm), (size_t)s->n_per * 4, s->stream));
      NB_TRY(s->comm->all_gather_f32_inplace(reinterpret_cast<float*>(s->d_vel), (size_t)s->n_per * 4, s->stream));
    }
    if (ev) NB_CUDA(cudaEventRecord(ev[4], s->stream));
  }
  if (integrate) s->steps++;
  return 0;
}

int run_steps(nbody_sim* s, float dt, int nsteps, bool integrate, bool sync) {
  if (!s->initialized) { set_error("not initialised: set bodies first (reference: Initialized == false)"); return NBODY_ERR_STATE; }
  NB_CUDA(cudaSetDevice(s->cfg.device));
  const int timed = sync ? std::min(nsteps, 128) : 0;
  NB_CUDA(cudaEventRecord(s->ev0, s->stream));
  for (int k = 0; k < nsteps; k++) {
    cudaEvent_t ev[6];
    bool use = k < timed;
    if (use) for (int q = 0; q < 6; q++) { ev[q] = pool_event(s, (size_t)k * 6 + q); if (!ev[q]) use = false; }
    NB_TRY(enqueue_step(s, dt, integrate, use ? ev : nullptr));
  }
  NB_CUDA(cudaEventRecord(s->ev1, s->stream));
  if (!sync) return 0;
  NB_CUDA(cudaStreamSynchronize(s->stream));
  NB_CUDA(cudaEventElapsedTime(&s->ms_call, s->ev0, s->ev1));
  s->ms_build = s->ms_force = s->ms_integrate = s->ms_comm = 0;
  for (int k = 0; k < timed; k++) {
    float a = 0, b = 0, c = 0, d = 0, x = 0;
    cudaEvent_t* e = &s->ev_pool[(size_t)k * 6];
    NB_CUDA(cudaEventElapsedTime(&a, e[0], e[1]));
    NB_CUDA(cudaEventElapsedTime(&x, e[1], e[5]));
    NB_CUDA(cudaEventElapsedTime(&b, e[5], e[2]));
    NB_CUDA(cudaEventElapsedTime(&c, e[2], e[3]));
    NB_CUDA(cudaEventElapsedTime(&d, e[3], e[4]));
    s->ms_build += a; s->ms_force += b; s->ms_integrate += c; s->ms_comm += d + x;
  }
  if (timed > 0 && timed < nsteps) {  // scale the sampled phases to the whole call
    const float f = (float)nsteps / (float)timed;
    s->ms_build *= f; s->ms_force *= f; s->ms_integrate *= f; s->ms_comm *= f;
  }
  if (s->let_mode() && timed > 0 && s->ev_ph[0]) {
    const int a[5] = {0, 2, 3, 4, 5}, b[5] = {1, 3, 4, 5, 6};
    for (int q = 0; q < 5; q++) if (cudaEventElapsedTime(&s->ms_ph[q], s->ev_ph[a[q]], s->ev_ph[b[q]]) != cudaSuccess) { s->ms_ph[q] = 0; cudaGetLastError(); }
  }
  if (s->cfg.method == NBODY_BARNES_HUT) {
    NB_TRY(bh_fetch_stats(s->tree, s->stream, &s->interactions));
    // AOctreeSearch::Size as the last Tick left it: max |coordinate| at the START of that step (OctreeSearch.cpp:26,47-56)
    uint32_t bits = 0;
    NB_CUDA(cudaMemcpyAsync(&bits, s->d_box, 4, cudaMemcpyDeviceToHost, s->stream));
    NB_CUDA(cudaStreamSynchronize(s->stream));
    memcpy(&s->cube_size, &bits, 4);
  }
  return 0;
}

int stage_reserve(nbody_sim* s, int64_t bytes) { return dev_reserve(&s->d_stage, &s->cap_stage, bytes, s->stream); }

int finish_set(nbody_sim* s) {
  s->ids_identity = true;
  if (s->bh()) {
    if (s->let_mode()) bh_iota(s->d_ids, (int)s->n_local, (int)s->slice_begin, s->stream);
    else bh_iota(s->d_ids, (int)s->n_global, 0, s->stream);
    s->launches++;
  }
  s->n_let = 0;
  s->n_migrated = 0;
  s->segs = BodySegs{0, (int)s->n_local, 0};
  if (s->let_mode()) {
    // Domain split: the rank uploaded a slice of the caller's order; send every body to the rank that owns its stretch of
    // the Morton curve now (a collective: all ranks set their bodies together). The root cube and the splitters of an
    // earlier run are kept when they still fit - a caller that uploads the same system every frame pays one exchange,
    // not a re-balancing from scratch.
    const int depth_hint = s->tree.depth_host;
    bh_reset(s->tree_let, s->stream);
    s->tree.depth_host = depth_hint;
    NB_TRY(launch_cube_size(s));
    BHParams bp;
    bp.sticky_root = true;
    double launches = 0;
    int n_new = 0, n_stay = 0;
    NB_TRY(bh_let_redistribute(s->tree, s->comm, bp, s->d_posm, s->d_vel, s->d_ids, s->d_posm2, s->d_vel2, s->d_ids2, (int)s->n_local,
                               std::min(std::min(s->cap_posm, s->cap_posm2), s->cap_acc), s->d_box, &n_new, &n_stay, s->stream, &launches));
    s->drifted = false;
    s->launches += launches;
    s->n_local = n_new;
    s->segs = BodySegs{0, n_new, 0};
    s->ids_identity = false;
    if (n_new > 0) NB_CUDA(cudaMemsetAsync(s->d_acc, 0, (size_t)n_new * 16, s->stream));   // the uploaded accelerations stayed with the slice
  } else {
    bh_reset(s->tree, s->stream);
    bh_reset(s->tree_let, s->stream);
  }
  NB_TRY(publish_positions(s));
  NB_TRY(detect_equal_mass(s));
  NB_CUDA(cudaStreamSynchronize(s->stream));
  s->initialized = true;
  s->steps = 0;
  return 0;
}

// Domain split: after a lazy migration the bodies sit in two stretches; everything that looks at the whole local set
// (read-backs, energy, device pointers) first makes them contiguous again.
int let_make_contiguous(nbody_sim* s) {
  if (!s->let_mode() || (s->segs.b0 == 0 && s->segs.n0 >= (int)s->n_local)) return 0;
  NB_TRY(dev_reserve(&s->d_acc2, &s->cap_acc2, s->cap_acc, s->stream));
  if (s->n_local > 0) {
    compact_segments_kernel<<<(unsigned)ceil_div(s->n_local, 256), 256, 0, s->stream>>>(s->segs.b0, s->segs.n0, s->segs.b1, (int)s->n_local, s->d_posm, s->d_vel,
                                                                                         s->d_acc, s->d_ids, s->d_posm2, s->d_vel2, s->d_acc2, s->d_ids2);
    s->launches++;
    NB_CUDA(cudaGetLastError());
  }
  std::swap(s->d_posm, s->d_posm2); std::swap(s->cap_posm, s->cap_posm2);
  std::swap(s->d_vel, s->d_vel2);   std::swap(s->cap_vel, s->cap_vel2);
  std::swap(s->d_ids, s->d_ids2);   std::swap(s->cap_ids, s->cap_ids2);
  std::swap(s->d_acc, s->d_acc2);   std::swap(s->cap_acc, s->cap_acc2);
  s->segs = BodySegs{0, (int)s->n_local, 0};
  return 0;
}

// Domain-split read-back: every rank receives the rows of its slice of the caller's order (return to owner) into the
// staging buffer, [n_slice][rec_words] floats. rec = the local records in their current order. A collective.
int let_read_back(nbody_sim* s, const float* rec, int rec_words, float** out_dev, int64_t* n_slice_out) {
  const int64_t n_slice = std::max<int64_t>(0, std::min<int64_t>(s->n_per, s->n_global - s->slice_begin));
  NB_TRY(let_make_contiguous(s));
  NB_TRY(dev_reserve(&s->d_ret, &s->cap_ret, std::max<int64_t>(n_slice, 1) * rec_words, s->stream));
  double launches = 0;
  NB_TRY(bh_let_return(s->tree, s->comm, s->d_ids, (int)s->n_local, s->n_per, rec, rec_words, s->d_ret, (int)n_slice, s->slice_begin,
                       s->stream, &launches));
  s->launches += launches;
  *out_dev = s->d_ret;
  *n_slice_out = n_slice;
  return 0;
}

// Copies this rank's share back to the host at the bodies' original indices. what: 0 posm, 1 vel, 2 acc.
int stage_reserve(nbody_sim* s, int64_t bytes);

int get_array(nbody_sim* s, int what, float* out4, int64_t n) {
  if (!s->initialized) { set_error("not initialised"); return NBODY_ERR_STATE; }
  if (!out4 || n < s->n_global) return invalid("output buffer is NULL or smaller than n_global bodies");
  NB_CUDA(cudaSetDevice(s->cfg.device));
  NB_TRY(let_make_contiguous(s));
  const float4* src = what == 0 ? s->posm_local() : what == 1 ? s->vel_local() : s->acc_local();
  if (s->let_mode()) {   // bodies live wherever their domain is: bring the rows of this rank's slice home, one copy out
    float* rows = nullptr;
    int64_t n_slice = 0;
    NB_TRY(let_read_back(s, reinterpret_cast<const float*>(src), 4, &rows, &n_slice));
    if (n_slice > 0) NB_CUDA(cudaMemcpyAsync(out4 + 4 * s->slice_begin, rows, (size_t)n_slice * 16, cudaMemcpyDeviceToHost, s->stream));
    NB_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
  }
  if (s->n_local == 0) return 0;
  if (s->ids_identity) {
    NB_CUDA(cudaMemcpyAsync(out4 + 4 * s->identity_begin(), src, (size_t)s->n_local * 16, cudaMemcpyDeviceToHost, s->stream));
    NB_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
  }
  if (s->cfg.world == 1) {  // all bodies are here: undo the Morton order on the device, one contiguous copy out
    NB_TRY(stage_reserve(s, s->n_global * 16));
    scatter_float4_kernel<<<(unsigned)ceil_div(s->n_local, 256), 256, 0, s->stream>>>(src, s->d_ids, (int)s->n_local, reinterpret_cast<float4*>(s->d_stage));
    s->launches++;
    NB_CUDA(cudaGetLastError());
    NB_CUDA(cudaMemcpyAsync(out4, s->d_stage, (size_t)s->n_global * 16, cudaMemcpyDeviceToHost, s->stream));
    NB_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
  }
  std::vector<float4> tmp((size_t)s->n_local);
  std::vector<int32_t> ids((size_t)s->n_local);
  NB_CUDA(cudaMemcpyAsync(tmp.data(), src, (size_t)s->n_local * 16, cudaMemcpyDeviceToHost, s->stream));
  NB_CUDA(cudaMemcpyAsync(ids.data(), s->ids_local(), (size_t)s->n_local * 4, cudaMemcpyDeviceToHost, s->stream));
  NB_CUDA(cudaStreamSynchronize(s->stream));
  for (int64_t i = 0; i < s->n_local; i++) memcpy(out4 + 4 * (size_t)ids[(size_t)i], &tmp[(size_t)i], 16);
  return 0;
}

}  // namespace

// =============================================================================================== C ABI
extern "C" {

int nbody_abi_version(void) { return NBODY_ABI_VERSION; }
const char* nbody_last_error(void) { return g_err.c_str(); }

int nbody_config_default(nbody_config* cfg) {
  if (!cfg) return invalid("cfg is NULL");
  memset(cfg, 0, sizeof(*cfg));
  cfg->struct_size = sizeof(nbody_config);
  cfg->method = NBODY_BARNES_HUT;
  cfg->G = 1e4f;
  cfg->eps = 0.f;
  cfg->theta = 1.0f;
  cfg->ph_delta_time = 0.01f;
  cfg->device = 0;
  cfg->rank = 0;
  cfg->world = 1;
  cfg->leaf_size = 16;
  cfg->reference_root = 0;
  cfg->mac = 0;
  cfg->group_size = 32;   // measured (profiles/): 32-body groups, cells of <= 64 bodies cut in two, 16-body leaves = shortest step at N = 1M
  cfg->group_pack = 2;
  cfg->bh_exchange = -1;
  return NBODY_OK;
}

int nbody_comm_unique_id(uint8_t out128[128]) {
  if (!out128) return invalid("out128 is NULL");
  return Comm::unique_id(out128);
}

int nbody_comm_loopback_id(uint8_t out128[128]) {
  if (!out128) return invalid("out128 is NULL");
  return Comm::loopback_id(out128);
}

int nbody_create(nbody_sim** out, const nbody_config* cfg) {
  if (!out || !cfg) return invalid("NULL argument");
  *out = nullptr;
  if (cfg->struct_size != sizeof(nbody_config)) return invalid("nbody_config.struct_size mismatch (use nbody_config_default)");
  if (cfg->method != NBODY_DIRECT && cfg->method != NBODY_BARNES_HUT) return invalid("unknown method");
  if (cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world) return invalid("bad rank/world");
  if (!(cfg->eps >= 0.f) || !(cfg->theta >= 0.f)) return invalid("eps and theta must be >= 0");
  if (cfg->mac != 0 && cfg->mac != 1) return invalid("mac must be 0 (group) or 1 (per body, reference rule)");
  if (cfg->group_size != 32 && cfg->group_size != 64 && cfg->group_size != 128) return invalid("group_size must be 32, 64 or 128");
  if (cfg->leaf_size < 1 || cfg->leaf_size > 64) return invalid("leaf_size must be in [1, 64]");
  if (cfg->group_pack < 1 || cfg->group_pack > 64) return invalid("group_pack must be in [1, 64]");
  if (cfg->bh_exchange < -1 || cfg->bh_exchange > 1) return invalid("bh_exchange must be -1 (auto), 0 (domain split + LET) or 1 (replicated tree)");
  if (cfg->method == NBODY_BARNES_HUT && cfg->world > 16) return invalid("Barnes-Hut runs on at most 16 ranks");
  NB_TRY(check_device(cfg->device));
  nbody_sim* s = new nbody_sim();
  s->cfg = *cfg;
  auto fail = [&](int code) { nbody_destroy(s); return code; };
#define NB_C(expr) do { int _c = [&]() -> int { expr; return 0; }(); if (_c) return fail(_c); } while (0)
  NB_C(if (cfg->stream) { s->stream = (cudaStream_t)cfg->stream; } else { NB_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)); s->own_stream = true; });
  NB_C(NB_CUDA(cudaEventCreate(&s->ev0)));
  NB_C(NB_CUDA(cudaEventCreate(&s->ev1)));
  NB_C(NB_CUDA(cudaMalloc((void**)&s->d_box, 8 * sizeof(uint32_t))));
  NB_C(NB_CUDA(cudaMalloc((void**)&s->d_energy, 2 * sizeof(double))));
  if (cfg->world > 1) {
    bool zero_id = true;
    for (int k = 0; k < 128; k++) zero_id = zero_id && cfg->nccl_unique_id[k] == 0;
    if (zero_id) s->emulated = true;
    else NB_C(NB_TRY(Comm::create(&s->comm, cfg->nccl_unique_id, cfg->rank, cfg->world)));
  }
#undef NB_C
  *out = s;
  return NBODY_OK;
}

void nbody_destroy(nbody_sim* s) {
  if (!s) return;
  cudaSetDevice(s->cfg.device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  delete s->comm;
  bh_free(s->tree);
  bh_free(s->tree_let);
  cudaFree(s->d_posm2); cudaFree(s->d_vel2); cudaFree(s->d_ids2);
  cudaFree(s->d_posm); cudaFree(s->d_vel); cudaFree(s->d_acc); cudaFree(s->d_partial); cudaFree(s->d_ids);
  cudaFree(s->d_stage); cudaFree(s->d_ret); cudaFree(s->d_acc2); cudaFree(s->d_box); cudaFree(s->d_energy);
  for (cudaEvent_t e : s->ev_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : s->ev_ph) if (e) cudaEventDestroy(e);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

int nbody_create_space_points(nbody_sim* s, int64_t n, float size, uint64_t seed) {
  if (!s) return invalid("sim is NULL");
  if (n < 1 || n > (int64_t)1 << 30) return invalid("N must be in [1, 2^30] (the reference indexes Particles[0], OctreeSearch.cpp:68)");
  if (s->emulated) return invalid("an emulated rank (world > 1, all-zero NCCL id) takes its bodies from nbody_set_bodies only");
  NB_CUDA(cudaSetDevice(s->cfg.device));
  partition(s, n);
  NB_TRY(reserve_state(s));
  if (s->load_count() > 0) {
    space_points_kernel<<<(unsigned)ceil_div(s->load_count(), 256), 256, 0, s->stream>>>(seed, s->load_begin(), (int)s->load_count(), size, s->posm_load(), s->d_vel, s->d_acc);
    s->launches++;
    NB_CUDA(cudaGetLastError());
  }
  s->cube_size = size;  // OctreeSearch.cpp:60
  return finish_set(s);
}

int nbody_set_particles_aos(nbody_sim* s, const void* particles, int64_t n, size_t stride) {
  if (!s) return invalid("sim is NULL");
  if (n < 1 || n > (int64_t)1 << 30) return invalid("n must be in [1, 2^30]");
  if (!particles) return invalid("particles is NULL");
  if (stride < sizeof(nbody_particle) || stride % 4) return invalid("stride must be >= 40 and a multiple of 4");
  if (s->emulated) return invalid("an emulated rank (world > 1, all-zero NCCL id) takes its bodies from nbody_set_bodies only");
  NB_CUDA(cudaSetDevice(s->cfg.device));
  partition(s, n);
  NB_TRY(reserve_state(s));
  if (s->load_count() > 0) {
    const size_t bytes = (size_t)s->load_count() * stride;
    NB_TRY(stage_reserve(s, (int64_t)bytes));
    NB_CUDA(cudaMemcpyAsync(s->d_stage, (const uint8_t*)particles + (size_t)s->load_begin() * stride, bytes, cudaMemcpyHostToDevice, s->stream));
    aos_to_soa_kernel<<<(unsigned)ceil_div(s->load_count(), 256), 256, 0, s->stream>>>(s->d_stage, stride, 0, (int)s->load_count(), s->posm_load(), s->d_vel, s->d_acc);
    s->launches++;
    NB_CUDA(cudaGetLastError());
  }
  return finish_set(s);
}

int nbody_set_bodies(nbody_sim* s, const float* posm4, const float* vel4, int64_t n) {
  if (!s) return invalid("sim is NULL");
  if (n < 1 || n > (int64_t)1 << 30) return invalid("n must be in [1, 2^30]");
  if (!posm4) return invalid("posm4 is NULL");
  NB_CUDA(cudaSetDevice(s->cfg.device));
  partition(s, n);
  NB_TRY(reserve_state(s));
  if (s->load_count() > 0) {
    const size_t bytes = (size_t)s->load_count() * 16;
    NB_CUDA(cudaMemcpyAsync(s->posm_load(), posm4 + 4 * s->load_begin(), bytes, cudaMemcpyHostToDevice, s->stream));
    if (vel4) NB_CUDA(cudaMemcpyAsync(s->d_vel, vel4 + 4 * s->load_begin(), bytes, cudaMemcpyHostToDevice, s->stream));
    else NB_CUDA(cudaMemsetAsync(s->d_vel, 0, bytes, s->stream));
    NB_CUDA(cudaMemsetAsync(s->d_acc, 0, bytes, s->stream));
    // an emulated rank has no communicator to gather the other slices: take all sources from the caller
    if (s->emulated && !s->bh()) {
      NB_CUDA(cudaMemcpyAsync(s->d_posm, posm4, (size_t)n * 16, cudaMemcpyHostToDevice, s->stream));
      // no all-gather will ever fill the padding of the other ranks' slots
      if (s->cap_posm > n) {
        fill_float4_kernel<<<(unsigned)ceil_div(s->cap_posm - n, 256), 256, 0, s->stream>>>(s->d_posm + n, s->cap_posm - n, make_float4(kPadCoord, kPadCoord, kPadCoord, 0.f));
        s->launches++;
      }
    }
  }
  return finish_set(s);
}

int nbody_clean_particles(nbody_sim* s) {
  if (!s) return invalid("sim is NULL");
  NB_CUDA(cudaSetDevice(s->cfg.device));
  NB_CUDA(cudaStreamSynchronize(s->stream));
  s->initialized = false;  // OctreeSearch.cpp:93
  s->n_global = s->n_local = 0;
  s->steps = 0;
  bh_reset(s->tree, s->stream);
  bh_let_forget_domains(s->tree);
  return NBODY_OK;
}

int nbody_compute_cube_size(nbody_sim* s, float* size_out) {
  if (!s) return invalid("sim is NULL");
  if (!s->initialized) { set_error("not initialised"); return NBODY_ERR_STATE; }  // OctreeSearch.cpp:49 returns silently
  NB_CUDA(cudaSetDevice(s->cfg.device));
  NB_TRY(launch_cube_size(s));
  uint32_t bits = 0;
  NB_CUDA(cudaMemcpyAsync(&bits, s->d_box, 4, cudaMemcpyDeviceToHost, s->stream));
  NB_CUDA(cudaStreamSynchronize(s->stream));
  memcpy(&s->cube_size, &bits, 4);
  if (size_out) *size_out = s->cube_size;
  return NBODY_OK;
}

int nbody_create_octree(nbody_sim* s) {
  if (!s) return invalid("sim is NULL");
  return run_steps(s, 0.f, 1, false, true);
}

int nbody_tick(nbody_sim* s) {
  if (!s) return invalid("sim is NULL");
  if (!(s->cfg.ph_delta_time > 0.f)) return NBODY_OK;  // OctreeSearch.cpp:25: paused
  if (!s->initialized) return NBODY_OK;                // OctreeSearch.cpp:49,76: silently nothing to do
  return run_steps(s, s->cfg.ph_delta_time, 1, true, true);
}

int nbody_step(nbody_sim* s, float dt, int32_t nsteps) {
  if (!s) return invalid("sim is NULL");
  if (nsteps < 0) return invalid("nsteps < 0");
  if (!(dt > 0.f) || nsteps == 0) return NBODY_OK;
  return run_steps(s, dt, nsteps, true, true);
}

int nbody_step_async(nbody_sim* s, float dt, int32_t nsteps) {
  if (!s) return invalid("sim is NULL");
  if (nsteps < 0) return invalid("nsteps < 0");
  if (!(dt > 0.f) || nsteps == 0) return NBODY_OK;
  return run_steps(s, dt, nsteps, true, false);
}

int nbody_synchronize(nbody_sim* s) {
  if (!s) return invalid("sim is NULL");
  NB_CUDA(cudaSetDevice(s->cfg.device));
  NB_CUDA(cudaStreamSynchronize(s->stream));
  return NBODY_OK;
}

int nbody_get_positions(nbody_sim* s, float* out, int64_t n) { return s ? get_array(s, 0, out, n) : invalid("sim is NULL"); }
int nbody_get_velocities(nbody_sim* s, float* out, int64_t n) { return s ? get_array(s, 1, out, n) : invalid("sim is NULL"); }
int nbody_get_accelerations(nbody_sim* s, float* out, int64_t n) { return s ? get_array(s, 2, out, n) : invalid("sim is NULL"); }

int nbody_get_particles_aos(nbody_sim* s, void* particles, int64_t n, size_t stride) {
  if (!s) return invalid("sim is NULL");
  if (!s->initialized) { set_error("not initialised"); return NBODY_ERR_STATE; }
  if (!particles || n < s->n_global) return invalid("output buffer is NULL or smaller than n_global bodies");
  if (stride < sizeof(nbody_particle) || stride % 4) return invalid("stride must be >= 40 and a multiple of 4");
  NB_CUDA(cudaSetDevice(s->cfg.device));
  if (s->let_mode()) {
    NB_TRY(let_make_contiguous(s));
    const size_t lbytes = (size_t)std::max<int64_t>(s->n_local, 1) * 40;
    NB_TRY(stage_reserve(s, (int64_t)lbytes));
    if (s->n_local > 0) {
      soa_to_aos_kernel<<<(unsigned)ceil_div(s->n_local, 256), 256, 0, s->stream>>>(s->d_posm, s->d_vel, s->d_acc, (int)s->n_local, nullptr, reinterpret_cast<float*>(s->d_stage));
      s->launches++;
    }
    float* rows = nullptr;
    int64_t n_slice = 0;
    NB_TRY(let_read_back(s, reinterpret_cast<const float*>(s->d_stage), 10, &rows, &n_slice));
    uint8_t* dst0 = (uint8_t*)particles + (size_t)s->slice_begin * stride;
    if (n_slice > 0) {
      if (stride == 40) NB_CUDA(cudaMemcpyAsync(dst0, rows, (size_t)n_slice * 40, cudaMemcpyDeviceToHost, s->stream));
      else NB_CUDA(cudaMemcpy2DAsync(dst0, stride, rows, 40, 40, (size_t)n_slice, cudaMemcpyDeviceToHost, s->stream));
    }
    NB_CUDA(cudaStreamSynchronize(s->stream));
    return NBODY_OK;
  }
  if (s->n_local == 0) return NBODY_OK;
  const size_t bytes = (size_t)s->n_local * 40;
  NB_TRY(stage_reserve(s, (int64_t)bytes));
  // record i of the staging array = local body i, or (single rank, reordered bodies) the body whose original index is i
  const bool unpermute = !s->ids_identity && s->cfg.world == 1;
  soa_to_aos_kernel<<<(unsigned)ceil_div(s->n_local, 256), 256, 0, s->stream>>>(s->posm_local(), s->vel_local(), s->acc_local(), (int)s->n_local, unpermute ? s->d_ids : nullptr, reinterpret_cast<float*>(s->d_stage));
  s->launches++;
  NB_CUDA(cudaGetLastError());
  uint8_t* dst = (uint8_t*)particles;
  if ((s->ids_identity || unpermute) && stride == 40) {
    NB_CUDA(cudaMemcpyAsync(dst + (size_t)(unpermute ? 0 : s->identity_begin()) * 40, s->d_stage, bytes, cudaMemcpyDeviceToHost, s->stream));
    NB_CUDA(cudaStreamSynchronize(s->stream));
    return NBODY_OK;
  }
  std::vector<uint8_t> tmp(bytes);
  std::vector<int32_t> ids;
  NB_CUDA(cudaMemcpyAsync(tmp.data(), s->d_stage, bytes, cudaMemcpyDeviceToHost, s->stream));
  if (!s->ids_identity && !unpermute) {
    ids.resize((size_t)s->n_local);
    NB_CUDA(cudaMemcpyAsync(ids.data(), s->ids_local(), (size_t)s->n_local * 4, cudaMemcpyDeviceToHost, s->stream));
  }
  NB_CUDA(cudaStreamSynchronize(s->stream));
  for (int64_t i = 0; i < s->n_local; i++) {
    const int64_t g = unpermute ? i : s->ids_identity ? s->identity_begin() + i : ids[(size_t)i];
    memcpy(dst + (size_t)g * stride, tmp.data() + (size_t)i * 40, 40);
  }
  return NBODY_OK;
}

int nbody_get_local_ids(nbody_sim* s, int64_t* ids, int64_t cap, int64_t* n_local) {
  if (!s) return invalid("sim is NULL");
  if (!s->initialized) { set_error("not initialised"); return NBODY_ERR_STATE; }
  if (s->let_mode()) {   // the rows this rank's read-backs fill: its slice of the caller's order (return to owner)
    const int64_t n_slice = std::max<int64_t>(0, std::min<int64_t>(s->n_per, s->n_global - s->slice_begin));
    if (n_local) *n_local = n_slice;
    if (!ids) return NBODY_OK;
    if (cap < n_slice) return invalid("ids capacity too small");
    for (int64_t i = 0; i < n_slice; i++) ids[i] = s->slice_begin + i;
    return NBODY_OK;
  }
  if (n_local) *n_local = s->n_local;
  if (!ids) return NBODY_OK;
  if (cap < s->n_local) return invalid("ids capacity too small");
  if (s->ids_identity) {
    for (int64_t i = 0; i < s->n_local; i++) ids[i] = s->identity_begin() + i;
    return NBODY_OK;
  }
  NB_CUDA(cudaSetDevice(s->cfg.device));
  std::vector<int32_t> h((size_t)s->n_local);
  NB_CUDA(cudaMemcpyAsync(h.data(), s->ids_local(), (size_t)s->n_local * 4, cudaMemcpyDeviceToHost, s->stream));
  NB_CUDA(cudaStreamSynchronize(s->stream));
  for (int64_t i = 0; i < s->n_local; i++) ids[i] = h[(size_t)i];
  return NBODY_OK;
}

int nbody_set_param(nbody_sim* s, int32_t which, double v) {
  if (!s) return invalid("sim is NULL");
  switch (which) {
    case NBODY_PARAM_G: s->cfg.G = (float)v; return NBODY_OK;
    case NBODY_PARAM_EPS: if (!(v >= 0)) return invalid("eps must be >= 0"); s->cfg.eps = (float)v; return NBODY_OK;
    case NBODY_PARAM_THETA: if (!(v >= 0)) return invalid("theta must be >= 0"); s->cfg.theta = (float)v; return NBODY_OK;
    case NBODY_PARAM_PH_DELTA_TIME: s->cfg.ph_delta_time = (float)v; return NBODY_OK;
    case NBODY_PARAM_LEAF_SIZE: if (v < 1 || v > 64) return invalid("leaf_size must be in [1, 64]"); s->cfg.leaf_size = (int)v; return NBODY_OK;
    case NBODY_PARAM_REFERENCE_ROOT: s->cfg.reference_root = v != 0; return NBODY_OK;
    case NBODY_PARAM_SHOW_OCTREE: s->show_octree = v != 0; return NBODY_OK;
    case NBODY_PARAM_MAC: if (v != 0 && v != 1) return invalid("mac must be 0 or 1"); s->cfg.mac = (int)v; return NBODY_OK;
    case NBODY_PARAM_GROUP_SIZE: if (v != 32 && v != 64 && v != 128) return invalid("group_size must be 32, 64 or 128"); s->cfg.group_size = (int)v; return NBODY_OK;
    case NBODY_PARAM_GROUP_PACK: if (v < 1 || v > 64) return invalid("group_pack must be in [1, 64]"); s->cfg.group_pack = (int)v; return NBODY_OK;
    case NBODY_PARAM_METHOD: return invalid("method is fixed at nbody_create (device layout depends on it)");
    default: return invalid("unknown or read-only parameter");
  }
}

int nbody_get_param(nbody_sim* s, int32_t which, double* v) {
  if (!s || !v) return invalid("NULL argument");
  switch (which) {
    case NBODY_PARAM_G: *v = s->cfg.G; return NBODY_OK;
    case NBODY_PARAM_EPS: *v = s->cfg.eps; return NBODY_OK;
    case NBODY_PARAM_THETA: *v = s->cfg.theta; return NBODY_OK;
    case NBODY_PARAM_PH_DELTA_TIME: *v = s->cfg.ph_delta_time; return NBODY_OK;
    case NBODY_PARAM_METHOD: *v = s->cfg.method; return NBODY_OK;
    case NBODY_PARAM_LEAF_SIZE: *v = s->cfg.leaf_size; return NBODY_OK;
    case NBODY_PARAM_REFERENCE_ROOT: *v = s->cfg.reference_root; return NBODY_OK;
    case NBODY_PARAM_SHOW_OCTREE: *v = s->show_octree; return NBODY_OK;
    case NBODY_PARAM_INITIALIZED: *v = s->initialized; return NBODY_OK;
    case NBODY_PARAM_MAC: *v = s->cfg.mac; return NBODY_OK;
    case NBODY_PARAM_GROUP_SIZE: *v = s->cfg.group_size; return NBODY_OK;
    case NBODY_PARAM_GROUP_PACK: *v = s->cfg.group_pack; return NBODY_OK;
    default: return invalid("unknown parameter");
  }
}

int nbody_energy(nbody_sim* s, double* ke, double* pe) {
  if (!s) return invalid("sim is NULL");
  if (!s->initialized) { set_error("not initialised"); return NBODY_ERR_STATE; }
  NB_CUDA(cudaSetDevice(s->cfg.device));
  NB_TRY(let_make_contiguous(s));
  NB_CUDA(cudaMemsetAsync(s->d_energy, 0, 2 * sizeof(double), s->stream));
  // sources = all N bodies; local body i sits at source index first + i (what the self-pair exclusion needs). Direct and
  // replicated Barnes-Hut keep them in d_posm; in LET mode they are gathered from the ranks (a collective: every rank calls).
  const float4* src = s->d_posm;
  int64_t first = s->local_begin;
  if (s->let_mode()) NB_TRY(bh_let_gather_all(s->tree, s->comm, s->d_posm, (int)s->n_local, s->n_global, &src, &first, s->stream));
  if (s->n_local > 0) {
    energy_kernel<<<(unsigned)ceil_div(s->n_local, 256), 256, 0, s->stream>>>(src, (int)s->n_global, s->posm_local(), s->vel_local(), (int)s->n_local, first, s->cfg.G, s->cfg.eps * s->cfg.eps, s->d_energy);
    s->launches++;
    NB_CUDA(cudaGetLastError());
  }
  if (s->comm) NB_TRY(s->comm->all_reduce_f64_sum(s->d_energy, 2, s->stream));
  double h[2];
  NB_CUDA(cudaMemcpyAsync(h, s->d_energy, sizeof(h), cudaMemcpyDeviceToHost, s->stream));
  NB_CUDA(cudaStreamSynchronize(s->stream));
  if (ke) *ke = h[0];
  if (pe) *pe = h[1];
  return NBODY_OK;
}

int nbody_stats_get(nbody_sim* s, nbody_stats* out) {
  if (!s || !out) return invalid("NULL argument");
  memset(out, 0, sizeof(*out));
  out->struct_size = sizeof(nbody_stats);
  out->method = s->cfg.method;
  out->n_global = s->n_global; out->n_local = s->n_local; out->steps = s->steps;
  out->interactions = s->interactions; out->kernel_launches = s->launches;
  out->ms_last_call = s->ms_call; out->ms_force = s->ms_force; out->ms_build = s->ms_build;
  out->ms_integrate = s->ms_integrate; out->ms_comm = s->ms_comm;
  out->cube_size = s->cube_size;
  out->jsplit = s->plan.jsplit; out->i_per_thread = s->plan.i_per_thread;
  out->tree_nodes = s->tree.n_nodes_host; out->tree_depth = s->tree.depth_host; out->walk_groups = s->tree.n_groups_host; out->let_points = s->n_let;
  out->equal_mass = s->equal_mass ? 1 : 0;
  out->sort_passes = s->tree.sort_passes_host;
  out->migrated = s->n_migrated;
  out->ms_let_migrate = s->ms_ph[0]; out->ms_let_plan = s->ms_ph[1]; out->ms_let_walk_local = s->ms_ph[2]; out->ms_let_import = s->ms_ph[3];
  out->ms_let_walk_let = s->ms_ph[4];
  memcpy(out->root_com, s->tree.root_com_host, sizeof(out->root_com));
  out->root_mass = s->tree.root_mass_host;
  return NBODY_OK;
}

int nbody_octree_boxes(nbody_sim* s, float* boxes7, int64_t cap, int64_t* n_boxes) {
  if (!s || !n_boxes) return invalid("NULL argument");
  if (!s->initialized) { set_error("not initialised"); return NBODY_ERR_STATE; }
  if (s->cfg.method != NBODY_BARNES_HUT) return invalid("octree boxes exist only for the Barnes-Hut method");
  NB_CUDA(cudaSetDevice(s->cfg.device));
  return bh_leaf_boxes(s->tree, s->d_posm, (int)s->n_global, boxes7, cap, n_boxes, s->stream);
}

int nbody_device_ptrs(nbody_sim* s, void** posm4, void** vel4, void** acc4) {
  if (!s) return invalid("sim is NULL");
  if (!s->initialized) { set_error("not initialised"); return NBODY_ERR_STATE; }
  NB_TRY(let_make_contiguous(s));
  if (posm4) *posm4 = s->posm_local();
  if (vel4) *vel4 = s->vel_local();
  if (acc4) *acc4 = s->acc_local();
  return NBODY_OK;
}

// ---- snapshot (checkpoint / resume): flat little-endian file, bodies in the caller's original order ------------------
namespace {
struct SnapshotHeader {
  char magic[8];           // "NBODYB2\0"
  uint32_t version, header_bytes;
  int64_t n, steps;
  float G, eps, theta, ph_delta_time;
  int32_t method, reserved[7];
};
}  // namespace

int nbody_save_snapshot(nbody_sim* s, const char* path) {
  if (!s || !path) return invalid("NULL argument");
  if (!s->initialized) { set_error("not initialised"); return NBODY_ERR_STATE; }
  if (s->cfg.world != 1) return invalid("snapshots are written by single-rank handles (gather the ranks' shares first)");
  std::vector<float> posm((size_t)s->n_global * 4), vel((size_t)s->n_global * 4);
  NB_TRY(get_array(s, 0, posm.data(), s->n_global));
  NB_TRY(get_array(s, 1, vel.data(), s->n_global));
  SnapshotHeader h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, "NBODYB2", 8);
  h.version = 1; h.header_bytes = sizeof(h); h.n = s->n_global; h.steps = s->steps;
  h.G = s->cfg.G; h.eps = s->cfg.eps; h.theta = s->cfg.theta; h.ph_delta_time = s->cfg.ph_delta_time; h.method = s->cfg.method;
  FILE* f = fopen(path, "wb");
  if (!f) return invalid(std::string("cannot open ") + path + " for writing");
  const bool ok = fwrite(&h, sizeof(h), 1, f) == 1 && fwrite(posm.data(), 16, (size_t)h.n, f) == (size_t)h.n &&
                  fwrite(vel.data(), 16, (size_t)h.n, f) == (size_t)h.n;
  if (fclose(f) != 0 || !ok) return invalid(std::string("short write to ") + path);
  return NBODY_OK;
}

int nbody_load_snapshot(nbody_sim* s, const char* path) {
  if (!s || !path) return invalid("NULL argument");
  FILE* f = fopen(path, "rb");
  if (!f) return invalid(std::string("cannot open ") + path);
  SnapshotHeader h;
  if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "NBODYB2", 8) != 0 || h.version != 1 || h.header_bytes != sizeof(h) ||
      h.n < 1 || h.n > (int64_t)1 << 30) { fclose(f); return invalid(std::string(path) + " is not an nbody snapshot"); }
  std::vector<float> posm((size_t)h.n * 4), vel((size_t)h.n * 4);
  const bool ok = fread(posm.data(), 16, (size_t)h.n, f) == (size_t)h.n && fread(vel.data(), 16, (size_t)h.n, f) == (size_t)h.n;
  fclose(f);
  if (!ok) return invalid(std::string(path) + " is truncated");
  s->cfg.G = h.G; s->cfg.eps = h.eps; s->cfg.theta = h.theta; s->cfg.ph_delta_time = h.ph_delta_time;   // method stays the handle's
  NB_TRY(nbody_set_bodies(s, posm.data(), vel.data(), h.n));
  s->steps = h.steps;
  return NBODY_OK;
}

int nbody_octree_nodes(nbody_sim* s, float* com4, int32_t* meta4, int32_t* range2, uint64_t* keys, int64_t cap_nodes,
                       int64_t cap_keys, int64_t* n_nodes) {
  if (!s) return invalid("sim is NULL");
  if (!s->initialized) { set_error("not initialised"); return NBODY_ERR_STATE; }
  if (s->cfg.method != NBODY_BARNES_HUT) return invalid("the octree exists only for the Barnes-Hut method");
  NB_CUDA(cudaSetDevice(s->cfg.device));
  return bh_read_tree(s->tree, com4, meta4, range2, keys, cap_nodes, cap_keys, n_nodes, s->stream);
}

int nbody_sort_pairs_u64(int32_t device, const uint64_t* keys_in, int64_t n, int32_t key_bits, uint64_t* keys_out,
                         uint32_t* idx_out, float* ms) {
  if (!keys_in) return invalid("keys_in is NULL");
  if (key_bits < 1 || key_bits > 64) return invalid("key_bits must be in [1, 64]");
  NB_TRY(check_device(device));
  return sort_pairs_host(keys_in, n, key_bits, keys_out, idx_out, ms);
}

// ---- FP32 peak probe ----------------------------------------------------------------------------------
}  // extern "C"

namespace {
// 16 independent FFMA chains per thread with register-only operands: the sustained FP32 issue rate.
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, const float* in, int iters, long long* cyc) {
  float acc[16], x[16], y[16];
#pragma unroll
  for (int k = 0; k < 16; k++) { acc[k] = in[k]; x[k] = in[16 + k] + threadIdx.x; y[k] = in[32 + k]; }
  const long long c0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int k = 0; k < 16; k++) acc[k] = fmaf(x[k], y[(k + u) & 15], acc[k]);
  }
  const long long c1 = clock64();
  float sum = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) sum += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
  if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}
}  // namespace

extern "C" int nbody_measure_fp32_peak(int32_t device, double* tflops, double* sm_mhz) {
  NB_TRY(check_device(device));
  const int blocks = sm_count() * 4, iters = 40000;
  float *d_out = nullptr, *d_in = nullptr;
  long long* d_cyc = nullptr;
  NB_CUDA(cudaMalloc((void**)&d_out, (size_t)blocks * 256 * 4));
  NB_CUDA(cudaMalloc((void**)&d_in, 48 * 4));
  NB_CUDA(cudaMalloc((void**)&d_cyc, (size_t)blocks * 8));
  float h[48];
  for (int i = 0; i < 48; i++) h[i] = 0.001f * (float)(i + 1);
  NB_CUDA(cudaMemcpy(d_in, h, sizeof(h), cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  NB_CUDA(cudaEventCreate(&e0));
  NB_CUDA(cudaEventCreate(&e1));
  fp32_peak_kernel<<<blocks, 256>>>(d_out, d_in, iters / 4, d_cyc);
  NB_CUDA(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    NB_CUDA(cudaEventRecord(e0));
    fp32_peak_kernel<<<blocks, 256>>>(d_out, d_in, iters, d_cyc);
    NB_CUDA(cudaEventRecord(e1));
    NB_CUDA(cudaEventSynchronize(e1));
    float ms;
    NB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, ms);
  }
  std::vector<long long> cyc((size_t)blocks);
  NB_CUDA(cudaMemcpy(cyc.data(), d_cyc, (size_t)blocks * 8, cudaMemcpyDeviceToHost));
  double cmax = 0;
  for (long long c : cyc) cmax = std::max(cmax, (double)c);
  const double fma = 128.0 * (double)iters * (double)blocks * 256.0;
  if (tflops) *tflops = 2.0 * fma / ((double)best * 1e-3) * 1e-12;
  if (sm_mhz) *sm_mhz = cmax / ((double)best * 1e-3) * 1e-6;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d_out); cudaFree(d_in); cudaFree(d_cyc);
  return NBODY_OK;
}
