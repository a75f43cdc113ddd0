// rpx_unit.cuh -- batch "unit" kernels that expose single device functions through the
// C ABI, the GPU counterpart of the reference's Python-callable test wrappers
// (Face.intersect ctracer.pyx:1769, FaceList.compute_orientation :1955,
// InterfaceMaterial.eval_child_ray :1620, Distortion.z_offset[_and_gradient] :1703-1729;
// "mostly for testing", doc/source/creating_new_optics.rst:89-93).  Used by the
// function-level parity tests; not on the trace path.
#pragma once
#include "rpx_kernels.cuh"

namespace rpx {

RPX_DEV double ld_f64_u32(const uint32_t* w) { return __hiloint2double((int)w[1], (int)w[0]); }
RPX_DEV void st_f64_u32(uint32_t* w, double v) {
    w[0] = (uint32_t)__double2loint(v);
    w[1] = (uint32_t)__double2hiint(v);
}

static __global__ void k_unit_face_intersect(DevScene S, int face, const double* p1, const double* p2,
                                      unsigned long long n, int is_base_ray, double* out) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = face_intersect<RPX_FC_MESH>(S, &S.faces[face], ld3(p1 + 3 * i), ld3(p2 + 3 * i), is_base_ray);
}

static __global__ void k_unit_face_normal(DevScene S, int face, const double* pts, unsigned long long n,
                                   double* normal, double* tangent) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    vec3 nn, tt;
    compute_orientation<RPX_FC_MESH>(S, &S.faces[face], ld3(pts + 3 * i), &nn, &tt);
    normal[3 * i] = nn.x; normal[3 * i + 1] = nn.y; normal[3 * i + 2] = nn.z;
    tangent[3 * i] = tt.x; tangent[3 * i + 1] = tt.y; tangent[3 * i + 2] = tt.z;
}

RPX_DEV void unit_write_child(uint32_t* rec, const uint32_t* parent, const Kids& k, const Kid& c,
                              uint32_t idx) {
    for (int w = 0; w < RPX_WORDS_RAY; w++) rec[w] = parent[w];  // end_face_idx etc. carry over
    st_f64_u32(rec + 0, k.origin.x); st_f64_u32(rec + 2, k.origin.y); st_f64_u32(rec + 4, k.origin.z);
    st_f64_u32(rec + 6, c.dir.x); st_f64_u32(rec + 8, c.dir.y); st_f64_u32(rec + 10, c.dir.z);
    st_f64_u32(rec + 12, k.normal.x); st_f64_u32(rec + 14, k.normal.y); st_f64_u32(rec + 16, k.normal.z);
    st_f64_u32(rec + 18, k.evec.x); st_f64_u32(rec + 20, k.evec.y); st_f64_u32(rec + 22, k.evec.z);
    st_f64_u32(rec + 24, c.n.re); st_f64_u32(rec + 26, c.n.im);
    st_f64_u32(rec + 28, c.e1.re); st_f64_u32(rec + 30, c.e1.im);
    st_f64_u32(rec + 32, c.e2.re); st_f64_u32(rec + 34, c.e2.im);
    st_f64_u32(rec + 36, RPX_INF);
    st_f64_u32(rec + 38, k.phase);
    st_f64_u32(rec + 40, k.apath);
    rec[43] = idx;
    rec[46] = c.type;
}

static __global__ void k_unit_material_eval(DevScene S, int mat, const uint32_t* rays, unsigned long long n,
                                     const double* point, const double* normal, const double* tangent,
                                     uint32_t* out, uint32_t* counts) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t* rec = rays + i * RPX_WORDS_RAY;
    RayIn r;
    r.o = v3(ld_f64_u32(rec + 0), ld_f64_u32(rec + 2), ld_f64_u32(rec + 4));
    r.d = v3(ld_f64_u32(rec + 6), ld_f64_u32(rec + 8), ld_f64_u32(rec + 10));
    r.e = v3(ld_f64_u32(rec + 18), ld_f64_u32(rec + 20), ld_f64_u32(rec + 22));
    r.n = cx(ld_f64_u32(rec + 24), ld_f64_u32(rec + 26));
    r.e1 = cx(ld_f64_u32(rec + 28), ld_f64_u32(rec + 30));
    r.e2 = cx(ld_f64_u32(rec + 32), ld_f64_u32(rec + 34));
    r.len = ld_f64_u32(rec + 36);
    r.phase = ld_f64_u32(rec + 38);
    r.apath = ld_f64_u32(rec + 40);
    r.wl = rec[42];
    r.ident = rec[45];
    r.type = rec[46];
    Kids k;
    material_eval<RPX_MM_ALL>(S, &S.mats[mat], r, ld3(point + 3 * i), ld3(normal + 3 * i), ld3(tangent + 3 * i), k);
    uint32_t c = 0;
    if (k.has_a) unit_write_child(out + (2 * i + c++) * RPX_WORDS_RAY, rec, k, k.a, (uint32_t)i);
    if (k.has_b) unit_write_child(out + (2 * i + c++) * RPX_WORDS_RAY, rec, k, k.b, (uint32_t)i);
    counts[i] = c;
}

static __global__ void k_unit_distortion(DevScene S, int dist, const double* x, const double* y,
                                  unsigned long long n, double* z, double* grad) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    z[i] = distortion_z(S, &S.dists[dist], x[i], y[i]);
    vec3 g = distortion_zgrad(S, &S.dists[dist], x[i], y[i]);
    grad[3 * i] = g.x; grad[3 * i + 1] = g.y; grad[3 * i + 2] = g.z;
}

}  // namespace rpx
