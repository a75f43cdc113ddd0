// rpx_launch.h -- launchers of the specialised tracer kernels.  Each (gausslet, face-class)
// pair of k_shade lives in its own translation unit (rpx_shade_inst.cu compiled four times)
// so the variants build in parallel; rpx_api.cu only sees these prototypes.
#pragma once
#include <cuda_runtime.h>

#include "rpx_kernels.cuh"

#define RPX_MAX_DEVICES 64

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
    int smem_bytes;  // staged scene tables (0: the scene does not fit -> SS=false kernels)
    int ahead_face;  // -1 all faces, >= 0 one face, -2 no trace-ahead (see k_shade)
    const unsigned long long* n_dev;  // device-resident parent count (pipelined launches) or NULL
    unsigned long long* h_count;      // mapped host slot for len(new_rays) or NULL
    const uint32_t* hits_in = nullptr;  // "some ray of this generation hit something" (NULL: unknown, run)
    uint32_t* hits_out = nullptr;       // the same flag for the generation built by this launch (NULL: not wanted)
    const uint32_t* miss_in = nullptr;  // number of rays of this generation that hit NOTHING (NULL: unknown, no compaction)
    uint32_t* miss_out = nullptr;       // the same count for the generation built by this launch
};

// one launcher per compiled variant: g = gausslets, f = face class, m = material mask index
cudaError_t launch_shade_g0_f0_m0(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g0_f0_m1(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g0_f0_m2(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g0_f0_m3(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g0_f1_m0(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g0_f1_m1(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g0_f1_m2(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g0_f1_m3(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g0_f2_m0(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g0_f2_m1(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g0_f2_m2(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g0_f2_m3(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g1_f0_m0(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g1_f0_m1(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g1_f0_m2(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g1_f0_m3(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g1_f1_m0(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g1_f1_m1(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g1_f1_m2(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g1_f1_m3(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g1_f2_m0(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g1_f2_m1(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g1_f2_m2(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g1_f2_m3(cudaStream_t st, const ShadeArgs& a);
// scene tables too large for shared memory: widest face class / material mask, tables in global memory
cudaError_t launch_shade_g0_f2_m3_nss(cudaStream_t st, const ShadeArgs& a);
cudaError_t launch_shade_g1_f2_m3_nss(cudaStream_t st, const ShadeArgs& a);
typedef cudaError_t (*ShadeLauncher)(cudaStream_t, const ShadeArgs&);
inline ShadeLauncher shade_launcher(int gauss, int fc, int mm_idx, bool scene_shared) {
    if (!scene_shared) return gauss ? launch_shade_g1_f2_m3_nss : launch_shade_g0_f2_m3_nss;
    static const ShadeLauncher table[2][3][4] = {
        {{launch_shade_g0_f0_m0, launch_shade_g0_f0_m1, launch_shade_g0_f0_m2, launch_shade_g0_f0_m3},
         {launch_shade_g0_f1_m0, launch_shade_g0_f1_m1, launch_shade_g0_f1_m2, launch_shade_g0_f1_m3},
         {launch_shade_g0_f2_m0, launch_shade_g0_f2_m1, launch_shade_g0_f2_m2, launch_shade_g0_f2_m3}},
        {{launch_shade_g1_f0_m0, launch_shade_g1_f0_m1, launch_shade_g1_f0_m2, launch_shade_g1_f0_m3},
         {launch_shade_g1_f1_m0, launch_shade_g1_f1_m1, launch_shade_g1_f1_m2, launch_shade_g1_f1_m3},
         {launch_shade_g1_f2_m0, launch_shade_g1_f2_m1, launch_shade_g1_f2_m2, launch_shade_g1_f2_m3}}};
    return table[gauss ? 1 : 0][fc < 0 ? 0 : (fc > 2 ? 2 : fc)][mm_idx & 3];
}
// smem == 0 selects the SS=false instantiation (face class FULL, tables in global memory)
cudaError_t launch_intersect(int fc, cudaStream_t st, unsigned n_tiles, int smem, const DevScene& S, const Soa& rays,
                             double max_length, int only_face);

// capture planes (rpx_capture): one launch per filtered collection
struct CaptureArgs {
    DevScene S;
    Soa in, out;
    unsigned long long* tile_state;
    uint32_t* tile_counter;
    const unsigned long long* d_base;
    unsigned long long* d_next;
    uint32_t wl_offset;
    const uint32_t* wl_map;
    const uint32_t* face_ids;
    int smem_bytes;
};
cudaError_t launch_capture(int gauss, int fc, cudaStream_t st, unsigned n_tiles, const CaptureArgs& a);

cudaError_t launch_unit_face_intersect(cudaStream_t st, const DevScene& S, int face, const double* p1,
                                       const double* p2, unsigned long long n, int is_base_ray, double* out);
cudaError_t launch_unit_face_normal(cudaStream_t st, const DevScene& S, int face, const double* pts,
                                    unsigned long long n, double* normal, double* tangent);
cudaError_t launch_unit_material_eval(cudaStream_t st, const DevScene& S, int mat, const uint32_t* rays,
                                      unsigned long long n, const double* point, const double* normal,
                                      const double* tangent, uint32_t* out, uint32_t* counts);
cudaError_t launch_unit_distortion(cudaStream_t st, const DevScene& S, int dist, const double* x, const double* y,
                                   unsigned long long n, double* z, double* grad);

}  // namespace rpx
