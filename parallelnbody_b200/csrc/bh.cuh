// Barnes-Hut path (K4-K8) - interface used by nbody_sim.cu; implementation in bh.cu.
//
// Replaces class Octree of the reference (/root/reference/Source/NBody/OctreeSearch.h:21-109): Add (h:60-81) becomes
// Morton keys + radix sort + a top-down split of sorted key ranges; ComputeMass (h:83-97) a bottom-up monopole pass;
// ComputeForces (h:99-108) a warp-coherent stack walk (production) or a per-body depth-first walk with the reference's
// exact acceptance rule and visiting order (parity mode).
#pragma once
#include "common.cuh"

namespace nbody {

enum BHMac : int {
  kMacGroup = 0,  // one warp walks for a group of <= group_size neighbouring bodies; a cell is accepted when
                  // half-width / (distance from the group's bounding box to the cell's centre of mass) < theta.
                  // Never accepts a cell the reference's per-body test (OctreeSearch.h:103) would open.
  kMacBody = 1    // per body, exactly OctreeSearch.h:100-107: skip d == 0, accept when Size / d < Theta or one-body leaf,
                  // children visited in octant order
};

struct BHParams {
  float G = 1e4f, eps2 = 0.f, theta = 1.f;
  int leaf_size = 16;
  bool reference_root = false;
  int mac = kMacGroup;
  int group_size = 32;  // bodies per walk group: 32, 64 or 128 (1, 2 or 4 per lane)
  int depth_hint = 0;          // last known tree depth (0 = unknown): how many key levels the sort has to resolve
  int group_pack = 2;   // cells of <= group_pack * group_size bodies are cut into equal walk groups
  bool sticky_root = false;    // keep the previous root cube while it holds all bodies (multi-GPU domain split: keys stay comparable)
  bool node_boxes = false;     // the monopole pass also computes every node's bounding box (domain-split mode)
  bool keep_root = false;      // the root cube is already in place (tree over received points: same cube as the local tree)
  float let_damping = 0.5f;    // domain split: fraction of the way a splitter moves towards its new equal-work quantile per step
};

// Where the bodies of a build sit in the input arrays: logical body i is entry b0 + i for i < n0, else b1 + (i - n0).
struct BodySegs {
  int b0 = 0, n0 = 0x7fffffff, b1 = 0;
  __host__ __device__ int at(int i) const { return i < n0 ? b0 + i : b1 + (i - n0); }
};

// What one step of the domain-split mode exchanges, known on the host after the step's one synchronisation.
struct LetPlan {
  int world = 1, rank = 0, n = 0;
  int let_send[16] = {0}, let_recv[16] = {0};   // locally-essential points to / from rank q
  int64_t let_total = 0;
};

struct BHState {
  int n_nodes_host = 0, depth_host = 0, n_groups_host = 0, sort_passes_host = 0;
  float root_com_host[3] = {0, 0, 0};
  float root_mass_host = 0;
  float root_cube_host[4] = {0, 0, 0, 0};  // centre xyz, half-width
  void* impl = nullptr;
};

// Forget the previous tree (the next reference-mode root is centred on the origin, OctreeSearch.cpp:77).
void bh_reset(BHState& st, cudaStream_t s);
void bh_free(BHState& st);
void bh_iota(int32_t* ids, int n, int first, cudaStream_t s);
// Sorts the n bodies along the Morton curve (*_in -> posm / vel / ids, which must not alias the inputs), builds the
// tree over the sorted bodies and its monopoles. box = cube_size_kernel output (absmax, min xyz, max xyz).
int bh_build(BHState& st, const BHParams& p, const float4* posm_in, const float4* vel_in, const int32_t* ids_in,
             float4* posm, float4* vel, int32_t* ids, int n, const uint32_t* box, cudaStream_t s, double* launches,
             const BodySegs* segs = nullptr);
// Accelerations (G applied) of the sorted bodies [t0, t1) -> acc[t0 .. t1).
int bh_forces(BHState& st, const BHParams& p, const float4* posm, float4* acc, int n, int t0, int t1, cudaStream_t s,
              double* launches);
// Same walk with the sources taken from tree `src` (bodies posm, n of them) and the targets / walk groups from
// `tgt_tree` (bodies tgt); accumulate = add to acc instead of overwriting (second pass over received LET points).
int bh_forces_from(BHState& src, BHState& tgt_tree, const BHParams& p, const float4* posm, const float4* tgt, float4* acc, int n,
                   int t0, int t1, bool accumulate, cudaStream_t s, double* launches);

// ---- multi-GPU (K9): Morton domain split + body migration + locally-essential-tree exchange; see bh.cu
class Comm;
void bh_let_forget_domains(BHState& st);
int bh_let_redistribute(BHState& st, Comm* comm, const BHParams& p, float4* posm_a, float4* vel_a, int32_t* ids_a, float4* posm_b,
                        float4* vel_b, int32_t* ids_b, int n, int64_t cap, const uint32_t* box_global, int* n_local, int* n_stay,
                        cudaStream_t s, double* launches);
int bh_let_plan(BHState& local, Comm* comm, const BHParams& p, const float4* posm, int n, int64_t cap, LetPlan* plan, cudaStream_t s,
                double* launches);
int bh_let_plan_wait(BHState& local, int64_t cap, LetPlan* plan);
int bh_let_import(BHState& local, BHState& let, Comm* comm, const BHParams& p, const LetPlan& plan, const uint32_t* box_global, int* n_let,
                  cudaStream_t s, double* launches);
int bh_let_finish(BHState& local, Comm* comm, const BHParams& p, const LetPlan& plan, cudaStream_t s, double* launches);
int bh_let_return(BHState& local, Comm* comm, const int32_t* ids, int n, int64_t n_per, const float* rec, int rec_words, float* out,
                  int n_slice, int64_t first, cudaStream_t s, double* launches);
const float4* bh_let_sources(BHState& local);
int bh_let_gather_all(BHState& local, Comm* comm, const float4* posm, int n, int64_t n_global, const float4** out, int64_t* first,
                      cudaStream_t s);

// Synchronises; fills the *_host fields and the interaction count of the last bh_forces.
int bh_fetch_stats(BHState& st, cudaStream_t s, double* interactions);
int bh_leaf_boxes(BHState& st, const float4* posm, int n, float* boxes7, int64_t cap, int64_t* n_boxes, cudaStream_t s);

// Inspection: copies the first cap_nodes nodes (com float4, meta int4, range int2) and the sorted Morton keys to the host.
int bh_read_tree(BHState& st, float* com4, int32_t* meta4, int32_t* range2, uint64_t* keys, int64_t cap_nodes, int64_t cap_keys,
                 int64_t* n_nodes, cudaStream_t s);
// Stand-alone K5: host keys in, sorted keys + the stable permutation out; *ms (optional) = best-of-3 device time.
int sort_pairs_host(const uint64_t* keys_in, int64_t n, int key_bits, uint64_t* keys_out, uint32_t* idx_out, float* ms);

}  // namespace nbody
