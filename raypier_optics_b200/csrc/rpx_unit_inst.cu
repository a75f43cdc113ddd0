// rpx_unit_inst.cu -- the batch "unit" kernels (full face class, all materials) and their
// launchers, in their own translation unit so they build in parallel with the tracer kernels.
#include "rpx_launch.h"
#include "rpx_unit.cuh"

namespace rpx {

cudaError_t launch_unit_face_intersect(cudaStream_t st, const DevScene& S, int face, const double* p1,
                                       const double* p2, unsigned long long n, int is_base_ray, double* out) {
    k_unit_face_intersect<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(S, face, p1, p2, n, is_base_ray, out);
    return cudaGetLastError();
}
cudaError_t launch_unit_face_normal(cudaStream_t st, const DevScene& S, int face, const double* pts,
                                    unsigned long long n, double* normal, double* tangent) {
    k_unit_face_normal<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(S, face, pts, n, normal, tangent);
    return cudaGetLastError();
}
cudaError_t launch_unit_material_eval(cudaStream_t st, const DevScene& S, int mat, const uint32_t* rays,
                                      unsigned long long n, const double* point, const double* normal,
                                      const double* tangent, uint32_t* out, uint32_t* counts) {
    k_unit_material_eval<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(S, mat, rays, n, point, normal, tangent, out,
                                                                   counts);
    return cudaGetLastError();
}
cudaError_t launch_unit_distortion(cudaStream_t st, const DevScene& S, int dist, const double* x, const double* y,
                                   unsigned long long n, double* z, double* grad) {
    k_unit_distortion<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(S, dist, x, y, n, z, grad);
    return cudaGetLastError();
}

}  // namespace rpx
