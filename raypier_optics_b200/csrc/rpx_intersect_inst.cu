// rpx_intersect_inst.cu -- the face-class variants of k_intersect and k_capture.
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

template <bool GAUSS, int FC>
static cudaError_t launch_capture_t(cudaStream_t st, unsigned n_tiles, const CaptureArgs& a) {
    k_capture<GAUSS, FC><<<n_tiles, RPX_TILE, a.smem_bytes, st>>>(a.S, a.in, a.out, a.tile_state, a.tile_counter,
                                                                    a.d_base, a.d_next, a.wl_offset, a.wl_map,
                                                                    a.face_ids, a.smem_bytes);
    return cudaGetLastError();
}

cudaError_t launch_capture(int gauss, int fc, cudaStream_t st, unsigned n_tiles, const CaptureArgs& a) {
    if (gauss) {
        return fc == RPX_FC_SIMPLE ? launch_capture_t<true, RPX_FC_SIMPLE>(st, n_tiles, a)
                                   : launch_capture_t<true, RPX_FC_FULL>(st, n_tiles, a);
    }
    return fc == RPX_FC_SIMPLE ? launch_capture_t<false, RPX_FC_SIMPLE>(st, n_tiles, a)
                               : launch_capture_t<false, RPX_FC_FULL>(st, n_tiles, a);
}

}  // namespace rpx
