// Barnes-Hut path (K4-K8) - interface used by nbody_sim.cu. (Implementation lands in bh.cu.)
#pragma once
#include "common.cuh"

namespace nbody {

struct BHParams {
  float G = 1e4f, eps2 = 0.f, theta = 1.f;
  int leaf_size = 16;
  bool reference_root = false;
};

struct BHState {
  int n_nodes_host = 0, depth_host = 0;
  float root_com_host[3] = {0, 0, 0};
  float root_mass_host = 0;
  void* impl = nullptr;
};

void bh_reset(BHState& st);
void bh_free(BHState& st);
void bh_iota(int32_t* ids, int n, int first, cudaStream_t s);
// Sorts the bodies along the Morton curve (posm/vel/ids are permuted; the pointers may be swapped with internal
// double buffers), builds the tree and its monopoles. box = launch_cube_size output.
int bh_build(BHState& st, const BHParams& p, float4** posm, float4** vel, int32_t** ids, int n, const uint32_t* box,
             cudaStream_t s, double* launches);
int bh_forces(BHState& st, const BHParams& p, const float4* posm, float4* acc, int n, cudaStream_t s, double* launches);
int bh_fetch_stats(BHState& st, cudaStream_t s, double* interactions);
int bh_leaf_boxes(BHState& st, const float4* posm, int n, float* boxes7, int64_t cap, int64_t* n_boxes, cudaStream_t s);

}  // namespace nbody
