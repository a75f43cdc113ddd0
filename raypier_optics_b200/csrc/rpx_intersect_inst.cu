// rpx_intersect_inst.cu -- the two face-class variants of k_intersect.
#include "rpx_launch.h"

namespace rpx {

cudaError_t launch_intersect(int fc, cudaStream_t st, unsigned n_tiles, int smem, const DevScene& S, const Soa& rays,
                             double max_length, int only_face) {
    if (fc == RPX_FC_SIMPLE)
        k_intersect<RPX_FC_SIMPLE><<<n_tiles, RPX_TILE, smem, st>>>(S, rays, max_length, smem, only_face);
    else
        k_intersect<RPX_FC_FULL><<<n_tiles, RPX_TILE, smem, st>>>(S, rays, max_length, smem, only_face);
    return cudaGetLastError();
}

}  // namespace rpx
