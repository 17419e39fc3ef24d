#include "bh.cuh"
namespace nbody {
__global__ void iota_kernel(int32_t* ids, int n, int first) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) ids[i] = first + i; }
void bh_reset(BHState&) {}
void bh_free(BHState&) {}
void bh_iota(int32_t* ids, int n, int first, cudaStream_t s) { iota_kernel<<<(n + 255) / 256, 256, 0, s>>>(ids, n, first); }
int bh_build(BHState&, const BHParams&, float4**, float4**, int32_t**, int, const uint32_t*, cudaStream_t, double*) { set_error("Barnes-Hut not built yet"); return -1; }
int bh_forces(BHState&, const BHParams&, const float4*, float4*, int, cudaStream_t, double*) { set_error("Barnes-Hut not built yet"); return -1; }
int bh_fetch_stats(BHState&, cudaStream_t, double*) { return 0; }
int bh_leaf_boxes(BHState&, const float4*, int, float*, int64_t, int64_t*, cudaStream_t) { set_error("Barnes-Hut not built yet"); return -1; }
}
