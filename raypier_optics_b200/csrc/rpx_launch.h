// rpx_launch.h -- launchers of the specialised tracer kernels.  Each (gausslet, face-class)
// pair of k_shade lives in its own translation unit (rpx_shade_inst.cu compiled four times)
// so the variants build in parallel; rpx_api.cu only sees these prototypes.
#pragma once
#include <cuda_runtime.h>

#include "rpx_kernels.cuh"

namespace rpx {

struct ShadeArgs {
    DevScene S;
    Soa in, out;
    double max_length;
    unsigned long long* tile_state;
    uint32_t* tile_counter;
    unsigned long long* d_count;
    uint32_t* face_counts;
    uint32_t n_tiles;
    int smem_bytes;
};

// mm_idx: 0 LIGHT, 1 COATED, 2 FULLDIEL, 3 ALL (RPX_MM_* in rpx_materials.cuh)
cudaError_t launch_shade_g0_f0(int mm_idx, cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g0_f1(int mm_idx, cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g1_f0(int mm_idx, cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g1_f1(int mm_idx, cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_intersect(int fc, cudaStream_t st, unsigned n_tiles, int smem, const DevScene& S, const Soa& rays,
                             double max_length);

}  // namespace rpx
