// rpx_intersect_inst.cu -- the face-class variants of k_intersect and k_capture.
#include "rpx_launch.h"

namespace rpx {

cudaError_t launch_intersect(int fc, cudaStream_t st, unsigned n_tiles, int smem, const DevScene& S, const Soa& rays,
                             double max_length, int only_face) {
    if (smem <= 0)
        k_intersect<RPX_FC_MESH, false><<<n_tiles, RPX_TILE, 0, st>>>(S, rays, max_length, only_face);
    else if (fc == RPX_FC_SIMPLE)
        k_intersect<RPX_FC_SIMPLE, true><<<n_tiles, RPX_TILE, smem, st>>>(S, rays, max_length, only_face);
    else if (fc == RPX_FC_FULL)
        k_intersect<RPX_FC_FULL, true><<<n_tiles, RPX_TILE, smem, st>>>(S, rays, max_length, only_face);
    else
        k_intersect<RPX_FC_MESH, true><<<n_tiles, RPX_TILE, smem, st>>>(S, rays, max_length, only_face);
    return cudaGetLastError();
}

template <bool GAUSS, int FC, bool SS>
static cudaError_t launch_capture_t(cudaStream_t st, unsigned n_tiles, const CaptureArgs& a) {
    k_capture<GAUSS, FC, SS><<<n_tiles, RPX_TILE, SS ? a.smem_bytes : 0, st>>>(
        a.S, a.in, a.out, a.tile_state, a.tile_counter, a.d_base, a.d_next, a.wl_offset, a.wl_map, a.face_ids);
    return cudaGetLastError();
}

cudaError_t launch_capture(int gauss, int fc, cudaStream_t st, unsigned n_tiles, const CaptureArgs& a) {
    // capture faces are planes in practice: anything else runs the widest instantiation
    if (a.smem_bytes <= 0)
        return gauss ? launch_capture_t<true, RPX_FC_MESH, false>(st, n_tiles, a)
                     : launch_capture_t<false, RPX_FC_MESH, false>(st, n_tiles, a);
    if (gauss) {
        return fc == RPX_FC_SIMPLE ? launch_capture_t<true, RPX_FC_SIMPLE, true>(st, n_tiles, a)
                                   : launch_capture_t<true, RPX_FC_MESH, true>(st, n_tiles, a);
    }
    return fc == RPX_FC_SIMPLE ? launch_capture_t<false, RPX_FC_SIMPLE, true>(st, n_tiles, a)
                               : launch_capture_t<false, RPX_FC_MESH, true>(st, n_tiles, a);
}

}  // namespace rpx
