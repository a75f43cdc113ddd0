// rpx_shade_inst.cu -- instantiates k_shade<RPX_I_GAUSS, RPX_I_FC, MM> for the four material
// masks.  Compiled once per (RPX_I_GAUSS, RPX_I_FC) pair (see Makefile).
#include "rpx_launch.h"

#ifndef RPX_I_GAUSS
#error "compile with -DRPX_I_GAUSS=0|1 -DRPX_I_FC=0|1"
#endif

#define RPX_CAT_(a, b, c) launch_shade_g##a##_f##b
#define RPX_CAT(a, b) RPX_CAT_(a, b, 0)

namespace rpx {

template <uint32_t MM>
static cudaError_t go(cudaStream_t st, const ShadeArgs& a) {
    k_shade<(RPX_I_GAUSS != 0), RPX_I_FC, MM><<<a.n_tiles, RPX_TILE, a.smem_bytes, st>>>(
        a.S, a.in, a.out, a.max_length, a.tile_state, a.tile_counter, a.d_count, a.face_counts, a.n_tiles,
        a.smem_bytes);
    return cudaGetLastError();
}

cudaError_t RPX_CAT(RPX_I_GAUSS, RPX_I_FC)(int mm_idx, cudaStream_t st, const ShadeArgs& a) {
    switch (mm_idx) {
        case 0: return go<RPX_MM_LIGHT>(st, a);
        case 1: return go<RPX_MM_COATED>(st, a);
        case 2: return go<RPX_MM_FULLDIEL>(st, a);
        default: return go<RPX_MM_ALL>(st, a);
    }
}

}  // namespace rpx
