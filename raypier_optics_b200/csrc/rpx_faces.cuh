// rpx_faces.cuh -- device ray x face intersection, surface normals, aperture shapes,
// implicit-surface bounds and Zernike distortions.
//
// One device function per reference Face class (raypier/core/cfaces.pyx), dispatched by
// a switch on rpx_face.type over the flat face table (staged in shared memory by the
// kernels).  Closed-form plane / sphere / conic / quadric roots, fp64 Newton iteration
// for aspheric and extended-polynomial faces, tangent-plane secant iteration for
// Zernike-distorted faces.  Semantics follow the cited reference lines, including its
// quirks; the arithmetic is free to contract to FMA.
#pragma once
#include "../../include/rpx.h"
#include "rpx_math.cuh"

namespace rpx {

// Device view of the flattened scene (pointers into device global memory, except
// faces/sets which the kernels re-point at their shared-memory copy).
struct DevScene {
    const rpx_face* faces;
    const rpx_face_set* sets;
    const rpx_material* mats;
    const rpx_shape_op* shape_ops;
    const rpx_implicit_op* impl_ops;
    const rpx_distortion* dists;
    const rpx_zcoef* zcoefs;
    const rpx_ztape_op* ztape;
    const double* wavelengths;
    const double* ntab;  // interleaved re, im
    const double* pool;
    // mesh / UV patch faces: the BVH repacked by rpx_scene_set for the device walk (mesh_intersect): 64-byte nodes
    // holding BOTH children's boxes in conservatively rounded fp32; mesh32_off[face] = first node of the face's
    // tree in bvh32 (in nodes), -1 = none (the walk then uses the fp64 nodes of the pool)
    const float4* bvh32;
    const int* mesh32_off;
    int n_traced, n_faces, n_sets, n_mats, n_wl, n_dists;
};
#define RPX_BVH32_NONE (-2147483647 - 1)   /* child slot without a child */

// Face-class specialisation of the kernels.  A scene made only of planes, spheres,
// extrusions and polygons runs kernels compiled WITHOUT the Newton / quadric / distortion
// code (RPX_FC_SIMPLE): fewer registers, no spills, a code footprint the instruction cache
// holds.  The host picks the variant from the face types present (rpx_scene_set).
#define RPX_FC_SIMPLE 0
#define RPX_FC_FULL 1
// ... and a third class for scenes with triangle-mesh / UV-patch faces: their BVH walk (explicit stacks in
// local memory) and Newton-on-a-patch code stay out of the kernels the BASELINE configs run (config3's
// k_shade: 2400 -> 1600 B of stack, fewer spills)
#define RPX_FC_MESH 2
RPX_DEV bool face_type_is_simple(int t) {
    return t == RPX_FACE_CIRCULAR || t == RPX_FACE_SHAPED_PLANAR || t == RPX_FACE_ELLIPTICAL_PLANE ||
           t == RPX_FACE_RECTANGULAR || t == RPX_FACE_SPHERICAL || t == RPX_FACE_SHAPED_SPHERICAL ||
           t == RPX_FACE_EXTRUDED_PLANAR || t == RPX_FACE_POLYGON || t == RPX_FACE_ORIENTED_POLYGON;
}

#define RPX_INF (__longlong_as_double(0x7ff0000000000000LL))
#define RPX_NO_HIT (-1.0)

// ------------------------------------------------------------------ shapes (cshapes.pyx)
RPX_DEV int shape_polygon_inside(const double* pts, int size, double X, double Y) {
    int ct = 0;  // cshapes.pyx:150-168: even-odd rule, half-open edges
    double y1 = pts[2 * (size - 1) + 1], x1 = pts[2 * (size - 1)];
    for (int i = 0; i < size; i++) {
        double y2 = pts[2 * i + 1], x2 = pts[2 * i];
        if ((y1 <= Y && Y < y2) || (y2 <= Y && Y < y1)) {
            if ((x1 + fdiv((Y - y1) * (x2 - x1), y2 - y1)) > X) ct = !ct;
        }
        y1 = y2;
        x1 = x2;
    }
    return ct;
}

// Shape tree as a postfix program on a bit stack (bit i of `stack` = i-th entry).
static __device__ __noinline__ int shape_inside(const DevScene& S, const rpx_face* f, double x, double y) {
    if (f->shape_off < 0) return 1;
    uint32_t stack = 0;
    int sp = 0;
    for (int i = 0; i < f->shape_len; i++) {
        const rpx_shape_op* op = &S.shape_ops[f->shape_off + i];
        uint32_t v;
        switch (op->type) {
            case RPX_SHAPE_CIRCLE: {  // cshapes.pyx:109-116 (strict <)
                double dx = x - op->p[0], dy = y - op->p[1];
                v = ((dx * dx) + (dy * dy) < (op->p[2] * op->p[2])) ? 1u : 0u;
                stack |= v << sp;
                sp++;
            } break;
            case RPX_SHAPE_RECT: {  // :128-136
                double dx = x - op->p[0], dy = y - op->p[1];
                v = ((2 * fabs(dx) < op->p[2]) && (2 * fabs(dy)) < op->p[3]) ? 1u : 0u;
                stack |= v << sp;
                sp++;
            } break;
            case RPX_SHAPE_POLYGON:
                v = (uint32_t)shape_polygon_inside(S.pool + op->aux_off, op->aux_n, x, y);
                stack |= v << sp;
                sp++;
                break;
            case RPX_SHAPE_NOT: stack ^= 1u << (sp - 1); break;
            case RPX_SHAPE_AND: {
                uint32_t b = (stack >> (sp - 1)) & 1u, a = (stack >> (sp - 2)) & 1u;
                sp--;
                stack &= ~(3u << (sp - 1));
                stack |= (a & b) << (sp - 1);
            } break;
            case RPX_SHAPE_OR: {
                uint32_t b = (stack >> (sp - 1)) & 1u, a = (stack >> (sp - 2)) & 1u;
                sp--;
                stack &= ~(3u << (sp - 1));
                stack |= (a | b) << (sp - 1);
            } break;
            case RPX_SHAPE_XOR: {
                uint32_t b = (stack >> (sp - 1)) & 1u, a = (stack >> (sp - 2)) & 1u;
                sp--;
                stack &= ~(3u << (sp - 1));
                stack |= (a ^ b) << (sp - 1);
            } break;
            default:  // RPX_SHAPE_TRUE
                stack |= 1u << sp;
                sp++;
                break;
        }
    }
    return (int)(stack & 1u);
}

// ------------------------------------------------ implicit surfaces (cimplicit_surfs.pyx)
static __device__ __noinline__ double implicit_eval(const DevScene& S, int off, int len, vec3 p) {
    double stack[8];
    int sp = 0;
    for (int i = 0; i < len; i++) {
        const rpx_implicit_op* op = &S.impl_ops[off + i];
        switch (op->type) {
            case RPX_IMPL_PLANE: stack[sp++] = dot(ld3(op->p + 3), p - ld3(op->p)); break;
            case RPX_IMPL_SPHERE: stack[sp++] = sep(p, ld3(op->p)) - op->p[3]; break;
            case RPX_IMPL_CYLINDER:
                stack[sp++] = mag(cross(p - ld3(op->p), ld3(op->p + 3))) - op->p[6];
                break;
            case RPX_IMPL_NEG: stack[sp - 1] = -stack[sp - 1]; break;
            case RPX_IMPL_MIN: sp--; if (stack[sp] < stack[sp - 1]) stack[sp - 1] = stack[sp]; break;
            case RPX_IMPL_MAX: sp--; if (stack[sp] > stack[sp - 1]) stack[sp - 1] = stack[sp]; break;
            case RPX_IMPL_SUB: sp--; stack[sp - 1] -= stack[sp]; break;
            default: stack[sp++] = -1.0; break;  // RPX_IMPL_NULL
        }
        if (sp > 7) sp = 7;
    }
    return stack[0];
}

// ------------------------------------------------------- distortions (cdistortions.pyx)
// Operand decode for the host-built Zernike tapes (see rpx.h, rpx_ztape_op).
RPX_DEV double zop(const double* ws, int op) {
    if (op < 2) return (double)op;
    int k = (op - 2) / 3, w = (op - 2) - 3 * k;
    return ws[w * RPX_ZERNIKE_MAX_K + k];
}

RPX_DEV void run_ztape(const DevScene& S, int off, int len, double r, double* ws) {
    for (int i = 0; i < len; i++) {
        const rpx_ztape_op t = S.ztape[off + i];
        double a = zop(ws, t.a), b = zop(ws, t.b), c = zop(ws, t.c);
        double val;
        if (t.kind == 0) {  // R: cdistortions.pyx:174-175
            val = r * (a + b);
            val -= c;
            ws[t.dst] = val;
        } else if (t.kind == 1) {  // R': :211-215
            val = a + b;
            val += r * (zop(ws, t.d) + zop(ws, t.e));
            val -= c;
            ws[RPX_ZERNIKE_MAX_K + t.dst] = val;
        } else if (t.kind == 2) {  // R/r: :312-313
            val = a + b;
            val -= c;
            ws[2 * RPX_ZERNIKE_MAX_K + t.dst] = val;
        } else {  // workspace[0, dst] = 1.0 (:433)
            ws[t.dst] = 1.0;
        }
    }
}

static __device__ __noinline__ double distortion_z(const DevScene& S, const rpx_distortion* D, double x, double y) {
    if (D->type == RPX_DIST_ZERNIKE_J7) {  // cdistortions.pyx:50-59
        x /= D->p[0];
        y /= D->p[0];
        double Z = sqrt(8.0) * (3 * (x * x + y * y) - 2) * y;
        return Z * D->p[1];
    }
    double ws[RPX_ZERNIKE_MAX_K];  // only row 0 is used by z_offset_c
    x /= D->p[0];
    y /= D->p[0];
    double r = sqrt_(x * x + y * y);
    double theta = atan2(y, x);
    run_ztape(S, D->tape_z_off, D->tape_z_len, r, ws);
    double Z = 0.0;
    for (int i = 0; i < D->n_coefs; i++) {
        const rpx_zcoef* c = &S.zcoefs[D->coef_off + i];
        double N = (c->m == 0) ? sqrt_((double)(c->n + 1)) : sqrt_((double)(2 * (c->n + 1)));
        N *= c->value;
        double PH = (c->m >= 0) ? cos(c->m * theta) : -sin(c->m * theta);  // Q11: signed m
        double R = zop(ws, c->opR_z);
        Z += N * R * PH;
    }
    return Z;
}

// -> (dz/dx, dz/dy, z)
static __device__ __noinline__ vec3 distortion_zgrad(const DevScene& S, const rpx_distortion* D, double x, double y) {
    if (D->type == RPX_DIST_ZERNIKE_J7) {  // cdistortions.pyx:61-81
        double root8 = sqrt(8.0) * D->p[1], R = D->p[0];
        x /= R;
        y /= R;
        return v3(root8 * 6 * x * y / R, root8 * (3 * x * x + 9 * y * y - 2) / R,
                  root8 * (3 * (x * x + y * y) - 2) * y);
    }
    double ws[3 * RPX_ZERNIKE_MAX_K];
    x /= D->p[0];
    y /= D->p[0];
    double r = sqrt_(x * x + y * y);
    double theta = atan2(y, x);
    run_ztape(S, D->tape_g_off, D->tape_g_len, r, ws);
    double st, ct;
    sincos(theta, &st, &ct);
    vec3 Z = v3(0.0, 0.0, 0.0);
    for (int i = 0; i < D->n_coefs; i++) {
        const rpx_zcoef* c = &S.zcoefs[D->coef_off + i];
        double sm, cm;
        sincos(c->m * theta, &sm, &cm);
        double PH, PHprime;
        if (c->m >= 0) {
            PH = cm;
            PHprime = -c->m * sm;
        } else {
            PH = -sm;
            PHprime = -c->m * cm;
        }
        double R = zop(ws, c->opR), Rprime = zop(ws, c->opRp), R_over_r = zop(ws, c->opRr);
        double N = (c->m == 0) ? sqrt_((double)(c->n + 1)) : sqrt_((double)(2 * (c->n + 1)));
        N *= c->value;
        Z.z += N * R * PH;
        Z.x += N * (Rprime * ct * PH + R_over_r * (-st) * PHprime);
        Z.y += N * (Rprime * st * PH + R_over_r * (ct)*PHprime);
    }
    Z.x /= D->p[0];
    Z.y /= D->p[0];
    return Z;
}

// ------------------------------------------------------------------ faces (cfaces.pyx)
RPX_DEV int point_in_polygon(double X, double Y, const double* pts, int size) {
    int ct = 0;  // cfaces.pyx:1050-1068
    double y1 = pts[2 * (size - 1) + 1], x1 = pts[2 * (size - 1)];
    for (int i = 0; i < size; i++) {
        double y2 = pts[2 * i + 1], x2 = pts[2 * i];
        double h = fdiv(Y - y1, y2 - y1);
        if (0 < h && h <= 1.0) {
            double x = x1 + h * (x2 - x1);
            if (x > X) ct = !ct;
        }
        y1 = y2;
        x1 = x2;
    }
    return ct;
}

// ---- ExtrudedBezierFace helpers (cfaces.pyx:717-845).  The reference evaluates the cubic's
// roots with x87 long double intermediates; fp64 agrees to ~1e-15 away from tangency.
struct flat2 {
    double x, y;
};
RPX_DEV flat2 f2(double x, double y) {
    flat2 r;
    r.x = x;
    r.y = y;
    return r;
}
RPX_DEV double eval_bezier(double t, double cp0, double cp1, double cp2, double cp3) {
    double u = 1 - t;
    return cp0 * (u * u * u) + 3 * cp1 * t * (u * u) + 3 * cp2 * u * (t * t) + cp3 * (t * t * t);
}
RPX_DEV double dif_bezier(double t, double cp0, double cp1, double cp2, double cp3) {
    double A = cp3 - 3 * cp2 + 3 * cp1 - cp0;
    double B = 3 * cp2 - 6 * cp1 + 3 * cp0;
    double C = 3 * cp1 - 3 * cp0;
    return 3 * A * (t * t) + 2 * B * t + C;
}
// roots_of_cubic, cfaces.pyx:735-787.  Returns the root count; the single-real-root branch of
// the reference divides by roots[0] == 0.0 (quirk Q8) and so never produces a usable root:
// it is reported as one non-finite root.
RPX_DEV int roots_of_cubic(double a, double b, double c, double d, double* roots) {
    if (fabs(a) <= 0.0000000001) {
        if (fabs(b) <= 0.0000000001) {
            roots[0] = (c == 0) ? 0.0 : -d / c;
            return 1;
        }
        double disc = sqrt(c * c - 4 * b * d);
        roots[0] = (-c + disc) / (2 * b);
        roots[1] = (-c - disc) / (2 * b);
        return 2;
    }
    double a1 = b / a, a2 = c / a, a3 = d / a;
    double Q = (a1 * a1 - 3.0 * a2) / 9.0;
    double R = (2.0 * a1 * a1 * a1 - 9.0 * a1 * a2 + 27.0 * a3) / 54.0;
    double R2_Q3 = R * R - Q * Q * Q;
    if (R2_Q3 < 0) {
        double theta = acos(R / sqrt(Q * Q * Q));
        double sq = -2.0 * sqrt(Q);
        roots[0] = sq * cos(theta / 3.0) - a1 / 3.0;
        roots[1] = sq * cos((theta + 2.0 * M_PI) / 3.0) - a1 / 3.0;
        roots[2] = sq * cos((theta + 4.0 * M_PI) / 3.0) - a1 / 3.0;
        return 3;
    }
    roots[0] = RPX_INF;
    return 1;
}
RPX_DEV flat2 rotate2D(double sn, double cs, flat2 p) { return f2(p.x * cs - p.y * sn, p.x * sn + p.y * cs); }
RPX_DEV bool bz_ccw(flat2 A, flat2 B, flat2 C) { return (C.y - A.y) * (B.x - A.x) > (B.y - A.y) * (C.x - A.x); }
RPX_DEV bool bz_seg_overlap(flat2 A, flat2 B, flat2 C, flat2 D) {
    return bz_ccw(A, C, D) != bz_ccw(B, C, D) && bz_ccw(A, B, C) != bz_ccw(A, B, D);
}
RPX_DEV bool bz_pnt_in_hull(flat2 p, flat2 A, flat2 B, flat2 C, flat2 D) {  // cfaces.pyx:815-845, float w
    bool i = p.x > A.x || p.x > B.x || p.x > C.x || p.x > D.x;
    bool j = p.x > A.x && p.x > B.x && p.x > C.x && p.x > D.x;
    bool k = i && !j;
    i = p.y > A.y || p.y > B.y || p.y > C.y || p.y > D.y;
    j = p.y > A.y && p.y > B.y && p.y > C.y && p.y > D.y;
    i = i && !j;
    float w = (float)(A.x - D.x);
    w = w * w;
    if (w <= .0005f) {
        w = (float)(B.x - A.x);
        w = w * w;
        if (w <= .0005f) i = k = true;
    } else {
        w = (float)(A.y - D.y);
        w = w * w;
        if (w <= .0005f) {
            w = (float)(B.y - A.y);
            w = w * w;
            if (w <= .0005f) i = k = true;
        }
    }
    return i && k;
}

// ExtrudedBezierFace.intersect_c, cfaces.pyx:867-971 (is_base_ray is ignored by the reference)
static __device__ __noinline__ double bezier_intersect(const DevScene& S, const rpx_face* f, vec3 p1, vec3 p2) {
    const double* P = f->p;
    const double* curves = S.pool + f->aux_off;
    const double z1 = P[0], z2 = P[1];
    const flat2 mincorner = f2(P[2], P[3]), maxcorner = f2(P[4], P[5]);
    if ((p1.z < z1 && p2.z < z1) || (p1.z > z2 && p2.z > z2)) return RPX_NO_HIT;
    const flat2 r = f2(p1.x, p1.y), q2 = f2(p2.x, p2.y);
    flat2 tv = f2(mincorner.x, maxcorner.y);
    if (!bz_seg_overlap(r, q2, mincorner, tv)) {
        if (!bz_seg_overlap(r, q2, tv, maxcorner)) {
            tv = f2(maxcorner.x, mincorner.y);
            if (!bz_seg_overlap(r, q2, maxcorner, tv)) {
                if (!bz_seg_overlap(r, q2, tv, mincorner)) return RPX_NO_HIT;
            }
        }
    }
    const double dZ = p2.z - p1.z;
    flat2 s = f2(p2.x - p1.x, p2.y - p1.y);
    const double theta = atan2(s.y, s.x);
    double sn, cs;
    sincos(-theta, &sn, &cs);
    s = rotate2D(sn, cs, s);
    const flat2 origin = f2(0, 0);
    double result = RPX_INF;
    for (int ci = 0; ci < f->aux_n; ci++) {
        flat2 cp[4];
#pragma unroll
        for (int q = 0; q < 4; q++)
            cp[q] = rotate2D(sn, cs, f2(curves[(ci * 4 + q) * 2] - p1.x, curves[(ci * 4 + q) * 2 + 1] - p1.y));
        if (bz_seg_overlap(origin, s, cp[0], cp[1]) || bz_seg_overlap(origin, s, cp[1], cp[2]) ||
            bz_seg_overlap(origin, s, cp[2], cp[3]) || bz_seg_overlap(origin, s, cp[3], cp[0])) {
            double A = cp[3].y - 3 * cp[2].y + 3 * cp[1].y - cp[0].y;
            double B = 3 * cp[2].y - 6 * cp[1].y + 3 * cp[0].y;
            double C = 3 * cp[1].y - 3 * cp[0].y;
            double D = cp[0].y;
            double roots[3];
            int n = roots_of_cubic(A, B, C, D, roots);
            while (n > 0) {
                n -= 1;
                double t = roots[n];
                if (0. < t && t < 1.) {
                    double b = eval_bezier(t, cp[0].x, cp[1].x, cp[2].x, cp[3].x);
                    if (0 < b && b < s.x) {
                        double c = dZ * b / s.x;
                        double a = c + p1.z;
                        if (z1 < a && a < z2) {
                            b = sqrt(c * c + b * b);
                            if (b < result && b > f->tolerance) result = b;
                        }
                    }
                }
            }
        }
    }
    if (result == RPX_INF) return RPX_NO_HIT;
    return result;
}

// ExtrudedBezierFace.compute_normal_c, cfaces.pyx:975-1046
static __device__ __noinline__ vec3 bezier_normal(const DevScene& S, const rpx_face* f, vec3 p) {
    const double* curves = S.pool + f->aux_off;
    flat2 ray = f2(p.x, p.y);
    const double theta = atan2(p.y, p.x);
    double sn, cs;
    sincos(-theta, &sn, &cs);
    for (int ci = 0; ci < f->aux_n; ci++) {
        flat2 cp[4];
#pragma unroll
        for (int q = 0; q < 4; q++) cp[q] = f2(curves[(ci * 4 + q) * 2], curves[(ci * 4 + q) * 2 + 1]);
        if (bz_pnt_in_hull(ray, cp[0], cp[1], cp[2], cp[3])) {
#pragma unroll
            for (int q = 0; q < 4; q++) cp[q] = rotate2D(sn, cs, cp[q]);
            double A = cp[3].y - 3 * cp[2].y + 3 * cp[1].y - cp[0].y;
            double B = 3 * cp[2].y - 6 * cp[1].y + 3 * cp[0].y;
            double C = 3 * cp[1].y - 3 * cp[0].y;
            double D = cp[0].y;
            double roots[3];
            int n = roots_of_cubic(A, B, C, D, roots);
            while (n > 0) {
                n -= 1;
                double t = roots[n];
                if (0 <= t && t <= 1) {
                    double tmp = eval_bezier(t, cp[0].x, cp[1].x, cp[2].x, cp[3].x);
                    if (tmp * tmp - (ray.x * ray.x + ray.y * ray.y) < .0001) {
                        flat2 dr = f2(dif_bezier(t, cp[0].x, cp[1].x, cp[2].x, cp[3].x),
                                      dif_bezier(t, cp[0].y, cp[1].y, cp[2].y, cp[3].y));
                        dr = rotate2D(-sn, cs, dr);  // rotate back by +theta
                        vec3 o = v3(0, 0, 0);
                        if (dr.y == 0) {
                            o.x = 0;
                            o.y = (dr.x > 0 ? 1 : -1);
                        } else if (dr.y > 0) {
                            o.x = -1;
                            o.y = dr.x / dr.y;
                        } else {
                            o.x = 1;
                            o.y = -dr.x / dr.y;
                        }
                        return norm(o);
                    }
                }
            }
        }
    }
    return v3(0, 0, 0);  // "Bezier normal not found": the reference prints and returns 0
}

// intersect_conic, cfaces.pyx:1695-1747
RPX_DEV double intersect_conic(vec3 a, vec3 d, double curvature, double conic_const) {
    double beta = 1 + conic_const;
    double R = -curvature;
    double b2 = beta * beta;
    double A = b2 * (d.z * d.z) + beta * (d.x * d.x) + beta * (d.y * d.y);
    double B = -2 * R * beta * d.z + 2 * a.x * beta * d.x + 2 * a.y * beta * d.y + 2 * a.z * b2 * d.z;
    double C = -2 * R * a.z * beta + (a.x * a.x) * beta + (a.y * a.y) * beta + (a.z * a.z) * b2;
    double D = B * B - 4 * A * C;
    if (D < 0) return -1;
    D = sqrt_(D);
    if (R * beta * d.z <= 0) return fdiv(-B + D, 2 * A);
    return fdiv(-B - D, 2 * A);
}

struct Aspheric {
    double R, beta, A4, A6, A8, A10, A12, A14, A16;
    vec3 a, d;
};

// eval_aspheric_impf + eval_aspheric_grad in one pass (cfaces.pyx:1854-1882); the even
// powers of r2 are built by repeated multiplication instead of libm pow().
RPX_DEV void aspheric_f_df(const Aspheric& A, double alpha, double* f, double* df) {
    double px = A.a.x + alpha * A.d.x, py = A.a.y + alpha * A.d.y;
    double r2 = px * px + py * py;
    double r4 = r2 * r2, r6 = r4 * r2, r8 = r4 * r4, r10 = r8 * r2, r12 = r8 * r4, r14 = r8 * r6,
           r16 = r8 * r8;
    double root = sqrt_(1 - A.beta * r2 / (A.R * A.R));
    double out = r2;
    out /= A.R * (1 + root);
    out -= A.a.z + alpha * A.d.z;
    out += A.A4 * r4 + A.A6 * r6 + A.A8 * r8 + A.A10 * r10 + A.A12 * r12 + A.A14 * r14 + A.A16 * r16;
    *f = out;
    if (df) {
        double dx = A.d.x * px, dy = A.d.y * py;
        double g = A.A10 * (10 * dx + 10 * dy) * r8;
        g += A.A12 * (12 * dx + 12 * dy) * r10;
        g += A.A14 * (14 * dx + 14 * dy) * r12;
        g += A.A16 * (16 * dx + 16 * dy) * r14;
        g += A.A4 * (4 * dx + 4 * dy) * (r2);
        g += A.A6 * (6 * dx + 6 * dy) * r4;
        g += A.A8 * (8 * dx + 8 * dy) * r6 - A.d.z;
        g += (2 * dx + 2 * dy) / (A.R * (root + 1));
        g += A.beta * (2 * dx + 2 * dy) * (r2) / (2 * (A.R * A.R * A.R) * root * ((root + 1) * (root + 1)));
        *df = g;
    }
}

// eval_extpoly_impf / eval_extpoly_grad, cfaces.pyx:2043-2126
RPX_DEV void extpoly_f_df(const rpx_face* f, const double* E, vec3 a, vec3 d, double alpha, double* fo,
                          double* dfo) {
    double R = f->p[0], beta = f->p[1], norm_radius = f->p[2], z_height = f->p[3];
    int Nx = f->aux_n, Ny = f->aux_m;
    double x = a.x + alpha * d.x, y = a.y + alpha * d.y;
    double r2 = x * x + y * y;
    double out = r2;
    if (R >= 0) out /= (R + sqrt_(R * R - beta * r2));
    else out /= (R - sqrt_(R * R - beta * r2));
    out -= a.z + alpha * d.z;
    double xn = x / norm_radius, yn = y / norm_radius;
    double xi = 1.0;
    for (int i = 0; i < Nx; i++) {
        double yj = 1.0;
        for (int j = 0; j < Ny; j++) {
            out += E[i * Ny + j] * xi * yj;
            yj *= yn;
        }
        xi *= xn;
    }
    out += z_height;
    *fo = out;
    if (dfo) {
        double R2 = R * R;
        double rt = sqrt_(1 - (beta * r2 / R2));
        double denom = R * (rt + 1);
        double nom = (2 * d.x * x + 2 * d.y * y);
        double inv_rad = 1. / norm_radius;
        double g = -d.z;
        g += nom / denom;
        g += beta * nom * r2 / (2 * R * rt * denom * denom);
        double xs = x * inv_rad, ys = y * inv_rad;
        double dEdx = 0.0, dEdy = 0.0;
        double xim1 = 1.0;  // xs^(i-1)
        for (int i = 1; i < Nx; i++) {
            double yj = 1.0;
            for (int j = 0; j < Ny; j++) {
                dEdx += (i)*E[i * Ny + j] * xim1 * yj;
                yj *= ys;
            }
            xim1 *= xs;
        }
        xi = 1.0;
        for (int i = 0; i < Nx; i++) {
            double yjm1 = 1.0;  // ys^(j-1)
            for (int j = 1; j < Ny; j++) {
                dEdy += (j)*E[i * Ny + j] * xi * yjm1;
                yjm1 *= ys;
            }
            xi *= xs;
        }
        dEdx *= inv_rad;
        dEdy *= inv_rad;
        g += dEdx * d.x;
        g += dEdy * d.y;
        *dfo = g;
    }
}

// Plane z = z0 with parametric test h in [tol, 1]; returns h or <0
RPX_DEV double plane_h(double z0, vec3 p1, vec3 p2, double tol, bool* ok) {
    double h = fdiv(z0 - p1.z, p2.z - p1.z);
    *ok = !((h < tol) || (h > 1.0));
    return h;
}


// The two roots a1 (+) and a2 (-) of a quadric with the sphere-style hemisphere and
// aperture culling shared by Spherical / ShapedSpherical faces.
RPX_DEV double sphere_hit(const DevScene& S, const rpx_face* f, vec3 r, vec3 p2, int is_base_ray,
                          double curvature, double z_height, double diameter, bool shaped) {
    vec3 s = p2 - r;  // cfaces.pyx:439-486 / 531-576
    double cz = z_height - curvature;
    vec3 d = r;
    d.z -= cz;
    double A = mag_sq(s);
    double B = 2 * dot(s, d);
    double C = mag_sq(d) - curvature * curvature;
    double D = B * B - 4 * A * C;
    if (D < 0) return RPX_NO_HIT;
    D = sqrt_(D);
    const double inv2A = rcp(2 * A);
    double a1 = (-B + D) * inv2A;
    vec3 pt1 = r + s * a1;
    double a2 = (-B - D) * inv2A;
    vec3 pt2 = r + s * a2;
    if (curvature >= 0) {
        if (pt1.z < cz) a1 = RPX_INF;
        if (pt2.z < cz) a2 = RPX_INF;
    } else {
        if (pt1.z > cz) a1 = RPX_INF;
        if (pt2.z > cz) a2 = RPX_INF;
    }
    if (is_base_ray) {
        if (!shaped) {
            double D4 = diameter * diameter / 4.;
            if ((pt1.x * pt1.x + pt1.y * pt1.y) > D4) a1 = RPX_INF;
            if ((pt2.x * pt2.x + pt2.y * pt2.y) > D4) a2 = RPX_INF;
        } else {
            if (!shape_inside(S, f, pt1.x, pt1.y)) a1 = RPX_INF;
            if (!shape_inside(S, f, pt2.x, pt2.y)) a2 = RPX_INF;
        }
    }
    if (a2 < a1) a1 = a2;
    if (a1 > 1.0 || a1 < f->tolerance) return RPX_NO_HIT;
    return a1 * sqrt_(A);  // sep(r, p2) == sqrt_(mag_sq(s))
}

// Face.intersect_c for the simple (non-wrapping) face classes.
// Both roots of  qa x^2 + qb x + qc = 0  from its discriminant root sd = sqrt(qb^2 - 4 qa qc), without
// the cancellation of the textbook form the reference uses ((-qb +- sd) / 2qa, e.g. cfaces.pyx:1262-1266):
// the root whose numerator does not cancel comes from q = -(qb + sign(qb) sd) / 2, the other one is
// qc / q.  For a ray that starts near the surface the reference's small root carries an error of
// ~1e3..1e6 ulp; this one is good to a few ulp, so the CUDA result differs from the reference's by the
// reference's own noise instead of adding to it.  *plus = (-qb + sd) / 2qa, *minus = (-qb - sd) / 2qa.
RPX_DEV void quad_roots(double qa, double qb, double qc, double sd, double* plus, double* minus) {
    const double q = -0.5 * (qb + copysign(sd, qb));
    const double big = q / qa, small_ = qc / q;
    if (q == 0.0) {  // qb == 0 and sd == 0: double root at 0 / 0 in this form -> textbook form
        *plus = (-qb + sd) / (2 * qa);
        *minus = (-qb - sd) / (2 * qa);
        return;
    }
    // qb >= 0: q = -(qb + sd)/2 -> big = (-qb - sd)/2qa is the "minus" root
    *plus = (qb >= 0.0) ? small_ : big;
    *minus = (qb >= 0.0) ? big : small_;
}

template <int FC>
__device__ double face_intersect_basic(const DevScene& S, const rpx_face* f, vec3 p1, vec3 p2,
                                       int is_base_ray) {
    const double* P = f->p;
    const double tol = f->tolerance;
    switch (f->type) {
        case RPX_FACE_CIRCULAR: {  // cfaces.pyx:151-178
            bool ok;
            double h = plane_h(P[2], p1, p2, tol, &ok);
            if (!ok) return RPX_NO_HIT;
            double X = p1.x + h * (p2.x - p1.x) - P[1];
            double Y = p1.y + h * (p2.y - p1.y);
            if (is_base_ray && (X * X + Y * Y) > (P[0] * P[0] / 4)) return RPX_NO_HIT;
            return h * sep(p1, p2);
        }
        case RPX_FACE_SHAPED_PLANAR: {  // :201-226
            bool ok;
            double h = plane_h(P[0], p1, p2, tol, &ok);
            if (!ok) return RPX_NO_HIT;
            double X = p1.x + h * (p2.x - p1.x);
            double Y = p1.y + h * (p2.y - p1.y);
            if (is_base_ray && !shape_inside(S, f, X, Y)) return RPX_NO_HIT;
            return h * sep(p1, p2);
        }
        case RPX_FACE_IMPLICIT_PLANAR: {  // :280-307
            if (FC == RPX_FC_SIMPLE) return RPX_NO_HIT;
            vec3 normal = ld3(P + 3), origin = ld3(P);
            vec3 dp = p2 - p1;
            vec3 po = origin - p1;
            double h = fdiv(dot(po, normal), dot(dp, normal));
            if ((h < tol) || (h > 1.0)) return RPX_NO_HIT;
            po = p1 + dp * h;
            if (is_base_ray && implicit_eval(S, f->aux_off, f->aux_n, po) > 0.0) return RPX_NO_HIT;
            return h * mag(dp);
        }
        case RPX_FACE_ELLIPTICAL_PLANE: {  // :322-341
            double gx = P[0], gy = P[1], d = P[2];
            double h = fdiv(gx * p1.x + gy * p1.y - p1.z,
                            (p2.z - p1.z) - gx * (p2.x - p1.x) - gy * (p2.y - p1.y));
            if ((h < tol) || (h > 1.0)) return RPX_NO_HIT;
            double X = p1.x + h * (p2.x - p1.x);
            double Y = p1.y + h * (p2.y - p1.y);
            if (is_base_ray && (X * X + Y * Y) > (d * d / 4)) return RPX_NO_HIT;
            return h * sep(p1, p2);
        }
        case RPX_FACE_RECTANGULAR: {  // :365-397
            bool ok;
            double h = plane_h(P[3], p1, p2, tol, &ok);
            if (!ok) return RPX_NO_HIT;
            if (is_base_ray) {
                double X = p1.x + h * (p2.x - p1.x) - P[2];
                double Y = p1.y + h * (p2.y - p1.y);
                if (X * X > P[0] * P[0] / 4) return RPX_NO_HIT;
                if (Y * Y > P[1] * P[1] / 4) return RPX_NO_HIT;
            }
            return h * sep(p1, p2);
        }
        case RPX_FACE_SPHERICAL: return sphere_hit(S, f, p1, p2, is_base_ray, P[1], P[2], P[0], false);
        case RPX_FACE_SHAPED_SPHERICAL:
            return sphere_hit(S, f, p1, p2, is_base_ray, P[0], P[1], 0.0, true);
        case RPX_FACE_EXTRUDED_PLANAR: {  // :665-699
            vec3 r = p1;
            double ux = P[0], uy = P[1];
            double vx = P[2] - ux, vy = P[3] - uy;
            vec3 s = p2 - r;
            double den = (s.x * vy - s.y * vx);
            if (is_base_ray) {
                double a = fdiv(s.y * (ux - r.x) - s.x * (uy - r.y), den);
                if (a < 0) return RPX_NO_HIT;
                if (a > 1) return RPX_NO_HIT;
            }
            double a = fdiv(vx * (r.y - uy) - vy * (r.x - ux), den);
            if (is_base_ray) {
                double dz = a * (p2.z - r.z);
                if (P[4] < (r.z + dz) && (r.z + dz) < P[5]) return a * mag(s);
                return RPX_NO_HIT;
            }
            return a * mag(s);
        }
        case RPX_FACE_POLYGON: {  // :1093-1108 (parabasal rays never hit: dist stays -1)
            bool ok;
            double h = plane_h(P[0], p1, p2, tol, &ok);
            if (!ok) return RPX_NO_HIT;
            double X = p1.x + h * (p2.x - p1.x);
            double Y = p1.y + h * (p2.y - p1.y);
            if (is_base_ray && point_in_polygon(X, Y, S.pool + f->aux_off, f->aux_n) == 1)
                return h * sep(p1, p2);
            return RPX_NO_HIT;
        }
        case RPX_FACE_ORIENTED_POLYGON: {  // :1189-1219 (h is an absolute distance here)
            vec3 n = ld3(P + 3), o = ld3(P);
            vec3 line = p2 - p1;
            double max_length = mag(line);
            line = norm(line);
            double h = dot(line, n);
            if (h == 0.0) return RPX_NO_HIT;
            h = fdiv(dot(o - p1, n), h);
            if ((h < tol) || (h > max_length)) return RPX_NO_HIT;
            if (is_base_ray) {
                line = (p1 + line * h) - o;
                double X = dot(line, ld3(P + 6));
                double Y = dot(line, ld3(P + 9));
                if (point_in_polygon(X, Y, S.pool + f->aux_off, f->aux_n) == 1) return h;
                return RPX_NO_HIT;
            }
            return h;
        }
        case RPX_FACE_OFFAXIS_PARABOLIC: {  // :1228-1298
            if (FC == RPX_FC_SIMPLE) return RPX_NO_HIT;
            double efl = P[0], diameter = P[1];
            double A = 1 / (2 * efl);
            vec3 s = p2 - p1;
            vec3 r = p1;
            r.z += efl / 2.;
            double a = A * (s.x * s.x + s.y * s.y);
            double b = 2 * A * (r.x * s.x + r.y * s.y) - s.z;
            double c = A * (r.x * r.x + r.y * r.y) - r.z;
            double d = b * b - 4 * a * c;
            if (d < 0) return RPX_NO_HIT;
            if (a < 1e-10) {
                double a1 = -c / b;
                vec3 pt1 = r + s * a1;
                pt1.x -= efl;
                if ((pt1.x * pt1.x + pt1.y * pt1.y) > (diameter / 2)) return RPX_NO_HIT;
                if (a1 > 1.0 || a1 < tol) return RPX_NO_HIT;
                return a1 * sep(p1, p2);
            }
            d = sqrt_(d);
            double a1, a2;
            quad_roots(a, b, c, d, &a1, &a2);
            vec3 pt1 = r + s * a1;
            vec3 pt2 = r + s * a2;
            pt1.x -= efl;
            pt2.x -= efl;
            if (is_base_ray) {
                d = diameter;
                d *= d / 4.;
                if ((pt1.x * pt1.x + pt1.y * pt1.y) > d) a1 = RPX_INF;
                if ((pt2.x * pt2.x + pt2.y * pt2.y) > d) a2 = RPX_INF;
            }
            if (a2 < a1) a1 = a2;
            if (a1 > 1.0 || a1 < tol) return RPX_NO_HIT;
            return a1 * sep(p1, p2);
        }
        case RPX_FACE_ELLIPSOIDAL: {  // :1344-1393
            if (FC == RPX_FC_SIMPLE) return RPX_NO_HIT;
            const double* T = S.pool + f->aux_off;
            vec3 Sv = p2 - p1;
            vec3 r = transform_pt(T, p1);
            vec3 s = transform_pt(T, p2);
            s = s - r;
            double B = P[1] * P[1], A = P[0] * P[0];
            double a = A * (s.z * s.z + s.y * s.y) + B * s.x * s.x;
            double b = 2 * (A * (r.z * s.z + r.y * s.y) + B * r.x * s.x);
            double c = A * (r.z * r.z + r.y * r.y) + B * r.x * r.x - A * B;
            double d = b * b - 4 * a * c;
            d = sqrt_(d);
            double root1, root2;
            if (d == d) quad_roots(a, b, c, d, &root1, &root2);
            else root1 = root2 = d;  // negative discriminant: NaN roots, as in the reference (no test there)
            vec3 q2 = p1 + Sv * root2;
            vec3 q1 = p1 + Sv * root1;
            if (is_base_ray) {
                if (!(P[2] < q2.x && q2.x < P[3])) root2 = 2;
                if (!(P[4] < q2.y && q2.y < P[5])) root2 = 2;
                if (!(P[6] < q2.z && q2.z < P[7])) root2 = 2;
                if (!(P[2] < q1.x && q1.x < P[3])) root1 = 2;
                if (!(P[4] < q1.y && q1.y < P[5])) root1 = 2;
                if (!(P[6] < q1.z && q1.z < P[7])) root1 = 2;
            }
            if (root1 < tol) root1 = 2;
            if (root2 < tol) root2 = 2;
            if (root1 > root2) root1 = root2;
            if (root1 > 1) return RPX_NO_HIT;
            return root1 * mag(Sv);
        }
        case RPX_FACE_SADDLE: {  // :1439-1495
            if (FC == RPX_FC_SIMPLE) return RPX_NO_HIT;
            double A = sqrt(6.0), root, denom, a1, a2;
            A *= P[1];
            vec3 p = p1;
            p.z -= P[0];
            vec3 d = p2 - p1;
            if (d.x == 0.0) {
                a1 = (-A * (p.x * p.y) + p.z) / (A * d.y * p.x - d.z);
                a2 = RPX_INF;
            } else if (d.y == 0.0) {
                a1 = (-A * (p.x * p.y) + p.z) / (A * d.x * p.y - d.z);
                a2 = RPX_INF;
            } else {
                double A2 = A * A;
                root = A2 * (d.x * d.x) * (p.y * p.y) - 2 * A2 * d.x * d.y * p.x * p.y +
                       A2 * (d.y * d.y) * (p.x * p.x) + 4 * A * d.x * d.y * p.z - 2 * A * d.x * d.z * p.y -
                       2 * A * d.y * d.z * p.x + d.z * d.z;
                if (root < 0) return RPX_NO_HIT;
                root = sqrt_(root);
                // A dx dy a^2 - t a + (A px py - pz) = 0 with t = -A dx py - A dy px + dz; the
                // reference's (t +- root) / denom, evaluated without cancellation
                const double t = -A * d.x * p.y - A * d.y * p.x + d.z;
                quad_roots(A * (d.x * d.y), -t, A * (p.x * p.y) - p.z, root, &a1, &a2);
            }
            vec3 pt1 = p1 + d * a1;
            vec3 pt2 = p1 + d * a2;
            if (a1 < 0.0) a1 = RPX_INF;
            if (a2 < 0.0) a2 = RPX_INF;
            if (is_base_ray) {
                if (!shape_inside(S, f, pt1.x, pt1.y)) a1 = RPX_INF;
                if (!shape_inside(S, f, pt2.x, pt2.y)) a2 = RPX_INF;
            }
            if (a2 < a1) a1 = a2;
            if (a1 > 1.0 || a1 < tol) return RPX_NO_HIT;
            return a1 * sep(p1, p2);
        }
        case RPX_FACE_CYLINDRICAL: {  // :1526-1584
            if (FC == RPX_FC_SIMPLE) return RPX_NO_HIT;
            double R = P[1];
            double R2 = R * R;
            vec3 o = p1;
            o.z -= P[0];
            vec3 d = p2 - p1;
            double ox2 = o.x * o.x, oz2 = o.z * o.z, dx2 = d.x * d.x, dz2 = d.z * d.z;
            double root = R2 * dz2 - 2 * R * dx2 * o.z + 2 * R * d.x * d.z * o.x - dx2 * oz2 +
                          2 * d.x * d.z * o.x * o.z - dz2 * ox2;
            if (root < 0) return RPX_NO_HIT;
            root = sqrt_(root);
            double denom = dx2 + dz2;
            double a1, a2;
            a1 = a2 = -R * d.z - d.x * o.x - d.z * o.z;
            a1 += root;
            a2 -= root;
            a1 /= denom;
            a2 /= denom;
            vec3 pt1 = p1 + d * a1;
            vec3 pt2 = p1 + d * a2;
            double cz = P[0] - P[1];
            if (R >= 0) {
                if (pt1.z < cz) a1 = RPX_INF;
                if (pt2.z < cz) a2 = RPX_INF;
            } else {
                if (pt1.z > cz) a1 = RPX_INF;
                if (pt2.z > cz) a2 = RPX_INF;
            }
            if (is_base_ray) {
                if (!shape_inside(S, f, pt1.x, pt1.y)) a1 = RPX_INF;
                if (!shape_inside(S, f, pt2.x, pt2.y)) a2 = RPX_INF;
            }
            if (a2 < a1) a1 = a2;
            if (a1 > 1.0 || a1 < tol) return RPX_NO_HIT;
            return a1 * sep(p1, p2);
        }
        case RPX_FACE_AXICON: {  // :1621-1675
            if (FC == RPX_FC_SIMPLE) return RPX_NO_HIT;
            double beta = P[1];
            vec3 d = p2 - p1;
            vec3 o = p1;
            o.z -= P[0];
            double beta2 = beta * beta;
            double ox2 = o.x * o.x, oy2 = o.y * o.y, oz2 = o.z * o.z;
            double dx2 = d.x * d.x, dy2 = d.y * d.y, dz2 = d.z * d.z;
            double root = -beta2 * dx2 * oy2 + 2 * beta2 * d.x * d.y * o.x * o.y - beta2 * dy2 * ox2 +
                          dx2 * oz2 - 2 * d.x * d.z * o.x * o.z + dy2 * oz2 - 2 * d.y * d.z * o.y * o.z +
                          dz2 * ox2 + dz2 * oy2;
            double denom = (beta2 * dx2 + beta2 * dy2 - dz2);
            if (root < 0) return RPX_NO_HIT;
            root = beta * sqrt_(root);
            double a1 = -beta2 * d.x * o.x - beta2 * d.y * o.y + d.z * o.z;
            double a2 = a1 + root;
            a1 -= root;
            a1 /= denom;
            a2 /= denom;
            vec3 pt1 = p1 + d * a1;
            vec3 pt2 = p1 + d * a2;
            if (pt1.z > P[0]) a1 = RPX_INF;
            if (pt2.z > P[0]) a2 = RPX_INF;
            if (is_base_ray) {
                if (!shape_inside(S, f, pt1.x, pt1.y)) a1 = RPX_INF;
                if (!shape_inside(S, f, pt2.x, pt2.y)) a2 = RPX_INF;
            }
            if (a2 < a1) a1 = a2;
            if (a1 > 1.0 || a1 < tol) return RPX_NO_HIT;
            return a1 * sep(p1, p2);
        }
        case RPX_FACE_CONIC: {  // :1767-1798
            if (FC == RPX_FC_SIMPLE) return RPX_NO_HIT;
            vec3 d = p2 - p1;
            vec3 a = p1;
            a.z -= P[1];
            double a1 = intersect_conic(a, d, P[0], P[2]);
            vec3 pt1 = a + d * a1;
            if (is_base_ray && !shape_inside(S, f, pt1.x, pt1.y)) return RPX_NO_HIT;
            if (a1 > 1.0 || a1 < tol) return RPX_NO_HIT;
            return a1 * sep(p1, p2);
        }
        case RPX_FACE_ASPHERIC: {  // :1909-1976, Newton on alpha
            if (FC == RPX_FC_SIMPLE) return RPX_NO_HIT;
            double atol2 = P[11] * P[11];
            vec3 d = p2 - p1;
            vec3 a = p1;
            a.z -= P[1];
            double a1 = intersect_conic(a, d, P[0], P[2]);
            Aspheric A;
            A.R = -P[0];
            A.beta = 1 + P[2];
            A.A4 = P[4]; A.A6 = P[5]; A.A8 = P[6]; A.A10 = P[7];
            A.A12 = P[8]; A.A14 = P[9]; A.A16 = P[10];
            A.a = a;
            A.d = d;
            double fv, f_last, g, dz;
            aspheric_f_df(A, a1, &fv, &g);
            f_last = fv;
            dz = -fv / g;
            bool converged = false;
            for (int i = 0; i < 100; i++) {
                a1 += dz;
                if (dz * dz < atol2) { converged = true; break; }
                aspheric_f_df(A, a1, &fv, &g);
                if (fabs(fv) > fabs(f_last)) return RPX_NO_HIT;
                f_last = fv;
                dz = -fv / g;
            }
            if (!converged) return RPX_NO_HIT;
            vec3 pt1 = a + d * a1;
            if (is_base_ray && !shape_inside(S, f, pt1.x, pt1.y)) return RPX_NO_HIT;
            if (a1 > 1.0 || a1 < tol) return RPX_NO_HIT;
            return a1 * sep(p1, p2);
        }
        case RPX_FACE_EXT_POLY: {  // :2186-2237
            if (FC == RPX_FC_SIMPLE) return RPX_NO_HIT;
            const double* E = S.pool + f->aux_off;
            double atol2 = P[4] * P[4];
            vec3 d = p2 - p1;
            vec3 a = p1;
            a.z -= P[3];
            double a1 = intersect_conic(a, d, -P[0], P[1] - 1.0);
            double fv, f_last, g, dz;
            extpoly_f_df(f, E, p1, d, a1, &fv, &g);
            f_last = fv;
            dz = -fv / g;
            bool converged = false;
            for (int i = 0; i < 100; i++) {
                a1 += dz;
                if (dz * dz < atol2) { converged = true; break; }
                extpoly_f_df(f, E, p1, d, a1, &fv, &g);
                if (fabs(fv) > fabs(f_last)) return RPX_NO_HIT;
                f_last = fv;
                dz = -fv / g;
            }
            if (!converged) return RPX_NO_HIT;
            vec3 pt1 = a + d * a1;
            if (is_base_ray && !shape_inside(S, f, pt1.x, pt1.y)) return RPX_NO_HIT;
            if (a1 > 1.0 || a1 < tol) return RPX_NO_HIT;
            return a1 * sep(p1, p2);
        }
        default: return RPX_NO_HIT;
    }
}

// Face.compute_normal_c (local coordinates) for the non-wrapping classes
template <int FC>
__device__ vec3 face_normal_basic(const DevScene& S, const rpx_face* f, vec3 p) {
    const double* P = f->p;
    switch (f->type) {
        case RPX_FACE_CIRCULAR: return v3(0, 0, P[3] != 0.0 ? 1 : -1);
        case RPX_FACE_SHAPED_PLANAR: return v3(0, 0, 1);
        case RPX_FACE_IMPLICIT_PLANAR: return ld3(P + 3);
        case RPX_FACE_ELLIPTICAL_PLANE: return norm(v3(P[0], P[1], -1));
        case RPX_FACE_RECTANGULAR: return v3(0, 0, -1);
        case RPX_FACE_SPHERICAL:
        case RPX_FACE_SHAPED_SPHERICAL: {
            double curvature = (f->type == RPX_FACE_SPHERICAL) ? P[1] : P[0];
            double z_height = (f->type == RPX_FACE_SPHERICAL) ? P[2] : P[1];
            p.z -= (z_height - curvature);
            if (curvature < 0) p = neg(p);
            return norm(p);
        }
        case RPX_FACE_EXTRUDED_PLANAR: return ld3(P + 6);
        case RPX_FACE_POLYGON: return v3(0, 0, -1);
        case RPX_FACE_ORIENTED_POLYGON: return ld3(P + 3);
        case RPX_FACE_OFFAXIS_PARABOLIC: {
            if (FC == RPX_FC_SIMPLE) return p;
            double A = 1 / (2 * P[0]);
            double m2 = p.x * p.x + p.y * p.y;
            double B = 4 * m2 * A * A;
            double dz = -sqrt_(B / (B + 1));
            m2 = sqrt_(m2);
            return v3(-(dz * p.x) / m2, -(dz * p.y) / m2, -1 / sqrt_(B + 1));
        }
        case RPX_FACE_ELLIPSOIDAL: {
            if (FC == RPX_FC_SIMPLE) return p;
            const double* T = S.pool + f->aux_off;
            p = transform_pt(T, p);
            vec3 n = v3(p.x / -(P[0] * P[0]), p.y / -(P[1] * P[1]), p.z / -(P[1] * P[1]));
            n = rotate_v(T + 12, n);
            return norm(n);
        }
        case RPX_FACE_SADDLE: {
            if (FC == RPX_FC_SIMPLE) return p;
            double rt6 = sqrt(6.0) * P[1];
            return norm(v3(-rt6 * p.y, -rt6 * p.x, 1.0));
        }
        case RPX_FACE_CYLINDRICAL: {
            if (FC == RPX_FC_SIMPLE) return p;
            p.z -= (P[0] - P[1]);
            if (P[1] < 0) { p.z = -p.z; p.x = -p.x; }
            p.y = 0;
            return norm(p);
        }
        case RPX_FACE_AXICON: {
            if (FC == RPX_FC_SIMPLE) return p;
            double beta = P[1];
            double r = sqrt_(p.x * p.x + p.y * p.y);
            return v3(beta * p.x / r, beta * p.y / r, 1.0);
        }
        case RPX_FACE_CONIC: {
            if (FC == RPX_FC_SIMPLE) return p;
            double R = -P[0], beta = 1 + P[2];
            int sign = (P[3] != 0.0) ? -1 : 1;
            p.z -= P[1];
            vec3 g = v3(-p.x * 2 * beta, -p.y * 2 * beta, 2 * beta * (R - beta * p.z));
            if ((R * beta) < 0) sign *= -1;
            return norm(g * (double)sign);
        }
        case RPX_FACE_ASPHERIC: {
            if (FC == RPX_FC_SIMPLE) return p;
            double R = -P[0], beta = 1 + P[2];
            int sign = (P[3] != 0.0) ? -1 : 1;
            p.z -= P[1];
            double r2 = p.x * p.x + p.y * p.y;
            double r4 = r2 * r2, r6 = r4 * r2, r8 = r4 * r4, r10 = r8 * r2, r12 = r8 * r4, r14 = r8 * r6;
            double root = sqrt_(1 - (beta * (r2) / (R * R)));
            double df = 10 * P[7] * r8 + 8 * P[6] * r6 + 6 * P[5] * r4 + 4 * P[4] * r2;
            df += 16 * P[10] * r14 + 14 * P[9] * r12 + 12 * P[8] * r10;
            df += 2 / (R * (1 + root));
            df += beta * (r2) / ((R * R * R) * root * ((1 + root) * (1 + root)));
            vec3 g = v3(-df * p.x, -df * p.y, 1.0);
            return norm(g * (double)sign);
        }
        case RPX_FACE_EXT_POLY: {
            if (FC == RPX_FC_SIMPLE) return p;
            const double* E = S.pool + f->aux_off;
            int Nx = f->aux_n, Ny = f->aux_m;
            double R = P[0], beta = P[1];
            bool inv = (P[5] != 0.0);
            int sign = inv ? -1 : 1;
            double inv_rad = 1. / P[2];
            double x = p.x * inv_rad, y = p.y * inv_rad;
            p.z -= P[3];
            vec3 g = v3(-p.x * 2 * beta, -p.y * 2 * beta, 2 * beta * (R - beta * p.z));
            if ((R * beta) < 0) sign *= -1;
            g = norm(g * (double)sign);
            double sx = 0.0, sy = 0.0;  // accumulated in the reference's term order
            double pm = inv ? 1.0 : -1.0;
            double xim1 = 1.0;
            for (int i = 1; i < Nx; i++) {
                double yj = 1.0;
                for (int j = 0; j < Ny; j++) {
                    g.x += pm * (i * E[i * Ny + j] * inv_rad * xim1 * yj);
                    yj *= y;
                }
                xim1 *= x;
            }
            double xi = 1.0;
            for (int i = 0; i < Nx; i++) {
                double yjm1 = 1.0;
                for (int j = 1; j < Ny; j++) {
                    g.y += pm * (j * E[i * Ny + j] * inv_rad * xi * yjm1);
                    yjm1 *= y;
                }
                xi *= x;
            }
            (void)sx; (void)sy;
            return norm(g);
        }
        default: return p;
    }
}

// DistortionFace.intersect_c, cfaces.pyx:2339-2416: base intersection, then a
// tangent-plane secant iteration on the distorted surface (<= 20 steps).
static __device__ __noinline__ double distortion_intersect(const DevScene& S, const rpx_face* f, vec3 p1, vec3 p2) {
    const rpx_face* base = &S.faces[f->base_face];
    const rpx_distortion* dist = &S.dists[f->aux_off];
    double h = sep(p2, p1);
    double tolerance = f->p[0];
    double a2 = face_intersect_basic<RPX_FC_FULL>(S, base, p1, p2, 0);
    if (a2 > h || a2 < f->tolerance) return RPX_NO_HIT;
    vec3 d = p2 - p1;
    vec3 pt1 = p1 + d * (a2 / h);
    vec3 dxdyz = distortion_zgrad(S, dist, pt1.x, pt1.y);
    vec3 n = face_normal_basic<RPX_FC_FULL>(S, base, pt1);
    pt1.z += dxdyz.z;
    n.x /= n.z;
    n.y /= n.z;
    n.x -= dxdyz.x;
    n.y -= dxdyz.y;
    vec3 o = p1 - pt1;
    double a1 = -h * dot(o, n) / dot(d, n);
    if (a1 < fabs(dxdyz.z)) return RPX_NO_HIT;
    for (int i = 0; i < 20; i++) {
        pt1 = p1 + d * (a1 / h);
        if (fabs(a1 - a2) < tolerance) break;
        double z_shift = distortion_z(S, dist, pt1.x, pt1.y);
        vec3 q1 = p1, q2 = p2;
        q1.z -= z_shift;
        q2.z -= z_shift;
        a2 = face_intersect_basic<RPX_FC_FULL>(S, base, q1, q2, 0);
        pt1 = q1 + d * (a2 / h);
        n = face_normal_basic<RPX_FC_FULL>(S, base, pt1);
        dxdyz = distortion_zgrad(S, dist, pt1.x, pt1.y);
        pt1.z += dxdyz.z;
        n.x /= n.z;
        n.y /= n.z;
        n.x -= dxdyz.x;
        n.y -= dxdyz.y;
        o = p1 - pt1;
        a2 = a1;
        a1 = -h * dot(o, n) / dot(d, n);
    }
    if (!shape_inside(S, f, pt1.x, pt1.y)) return RPX_NO_HIT;  // Q7: even for parabasal rays
    return a1;
}

// ------------------------------------------------------------------ triangle meshes (obbtree.pyx)
// OBBTreeFace.intersect_c (obbtree.pyx:913-932) over OBBTree.intersect_with_line_c (:367-400): the
// nearest triangle with tolerance / |p2 - p1| <= alpha < 1.  The reference prunes with its OBB tree
// (line_intersects_node_c, :271-296: a conservative interval test); here the pruning structure is a
// BVH of padded axis-aligned boxes built by the host (include/rpx.h, RPX_FACE_MESH): an explicit
// stack in local memory, slab test of the segment clipped to the best alpha so far, near child first.  The triangle test is line_intersects_cell_c
// (:310-343) on records that hold p1, v1, v2 and n = v1 x v2 ready-made.  Exactly equal alpha (a ray
// through a shared edge): the lowest cell id wins, whatever the traversal order.
// *piece = cell id of the hit (intersect_t.piece_idx), -1 on a miss.
static __device__ __noinline__ double mesh_intersect_f64(const DevScene& S, const rpx_face* f, vec3 p1, vec3 p2, int* piece,
                                                        int* rec_out = nullptr) {
    const double* H = S.pool + f->aux_off;
    const double* tris = H + (long long)H[5];
    const double* nodes = H + (long long)H[6];
    const vec3 d = p2 - p1;
    const double dmag = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);  // mag_(d): IEEE, it scales the result
    const double tol = H[7] / dmag;
    const double ix = 1.0 / d.x, iy = 1.0 / d.y, iz = 1.0 / d.z;  // +-inf for an axis-parallel segment
    double best = 1.0;
    long long best_id = -1;
    int best_rec = -1;  // position of the winning record in tris[] (leaf order)
    // Near child first (measured on B200, 71k-facet scene: k_intersect 3.58 -> 1.91 ms, k_shade 4.76 -> 3.77 ms
    // per 1e6 rays against the unordered walk; 55 -> 34 box tests and 10.6 -> 5.6 triangle tests per ray
    // through a closed 5120-facet ball): a node is tested when its parent is expanded, pushed with its entry
    // parameter, and skipped on pop when the best alpha has moved in front of it meanwhile.
    {
        auto slab = [&](const double* nd, double* t0) -> bool {
            const double ax = (nd[0] - p1.x) * ix, bx = (nd[3] - p1.x) * ix;
            const double ay = (nd[1] - p1.y) * iy, by = (nd[4] - p1.y) * iy;
            const double az = (nd[2] - p1.z) * iz, bz = (nd[5] - p1.z) * iz;
            const double tmin = fmax(fmax(fmin(ax, bx), fmin(ay, by)), fmax(fmin(az, bz), 0.0));
            const double tmax = fmin(fmin(fmax(ax, bx), fmax(ay, by)), fmin(fmax(az, bz), best));
            *t0 = tmin;
            return tmin <= tmax;
        };
        int sid[64];
        double st0[64];
        int sp = 0;
        double t0;
        if (slab(nodes, &t0)) {
            sid[0] = 0;
            st0[0] = t0;
            sp = 1;
        }
        while (sp > 0) {
            --sp;
            if (st0[sp] > best) continue;
            const double* nd = nodes + 8 * (long long)sid[sp];
            const double a = nd[6], b = nd[7];
            if (a >= 0.0) {
                double tl, tr;
                const bool hl = slab(nodes + 8 * (long long)a, &tl), hr = slab(nodes + 8 * (long long)b, &tr);
                if (sp > 61) continue;  // cannot happen below 2^60 triangles; never overrun the stack
                if (hl && hr) {
                    const bool left_first = tl <= tr;
                    sid[sp] = left_first ? (int)b : (int)a;
                    st0[sp++] = left_first ? tr : tl;
                    sid[sp] = left_first ? (int)a : (int)b;
                    st0[sp++] = left_first ? tl : tr;
                } else if (hl) {
                    sid[sp] = (int)a;
                    st0[sp++] = tl;
                } else if (hr) {
                    sid[sp] = (int)b;
                    st0[sp++] = tr;
                }
                continue;
            }
            const int first = (int)(-a - 1.0);
            const double* t = tris + 16 * (long long)first;
            const int count = (int)b;
            for (int c = 0; c < count; c++, t += 16) {
                const vec3 tp = v3(t[0], t[1], t[2]), v1 = v3(t[3], t[4], t[5]), v2 = v3(t[6], t[7], t[8]);
                const vec3 n = v3(t[9], t[10], t[11]);
                const double det = -dot(d, n);
                if (det == 0.0) continue;
                const double invdet = 1.0 / det;
                const vec3 a0 = p1 - tp;
                const vec3 da0 = cross(a0, d);
                const double u = dot(v2, da0) * invdet;
                const double v = -dot(v1, da0) * invdet;
                const double alpha = dot(a0, n) * invdet;
                if ((u + v > 1.0) | (u < 0) | (v < 0) | (alpha < 0)) continue;
                const long long id = (long long)t[12];
                if (alpha >= tol && (alpha < best || (alpha == best && id < best_id))) {
                    best = alpha;
                    best_id = id;
                    best_rec = first + c;
                }
            }
        }
        *piece = (int)best_id;
        if (rec_out) *rec_out = best_rec;
        if (best_id < 0) best = -1.0;
        return best * dmag;
    }
}

// The walk that ships (measured on B200, profiles/r02_notes.md): the tree repacked by rpx_scene_set into 64-byte
// nodes that hold BOTH children's boxes as fp32 rounded outwards -- one 64-byte load per step instead of three
// dependent 64-byte fp64 nodes, slab tests on the fp32 pipe, 8-byte stack entries.  The fp32 test is CONSERVATIVE:
// the segment is taken in fp32 (origin and direction rounded to nearest) and every box is widened by
// delta_c = 2^-19 (|p1_c| + |d_c| + the largest box coordinate), 32x the worst-case sum of the rounding errors of
// the conversion and of the six operations of the slab test, so a box the exact segment touches is never
// rejected; the running bound is the fp64 best alpha rounded UP.  The triangle test itself is the fp64 code of
// the fp64 walk, on the same records, with the same tie rule -- the result is bit-identical, only fewer / cheaper
// box tests are made.  Leaf reference: -(first * 8 + count - 1) - 1 (count <= 8).
static __device__ __noinline__ double mesh_intersect(const DevScene& S, const rpx_face* f, vec3 p1, vec3 p2, int* piece,
                                                    int* rec_out = nullptr) {
    const int root = S.mesh32_off ? S.mesh32_off[(int)(f - S.faces)] : -1;
    if (root < 0) return mesh_intersect_f64(S, f, p1, p2, piece, rec_out);
    const double* H = S.pool + f->aux_off;
    const double* tris = H + (long long)H[5];
    const float4* nodes = S.bvh32 + 4 * (long long)root;
    const vec3 d = p2 - p1;
    const double dmag = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);  // mag_(d): IEEE, it scales the result
    const double tol = H[7] / dmag;
    double best = 1.0;
    long long best_id = -1;
    int best_rec = -1;
    float best_f = 1.0f;
    // the fp32 segment and its padding
    const float ox = (float)p1.x, oy = (float)p1.y, oz = (float)p1.z;
    const float dx = (float)d.x, dy = (float)d.y, dz = (float)d.z;
    const float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;  // +-inf for an axis-parallel segment
    const float mx = __int_as_float(__float_as_int(nodes[3].z));  // largest |coordinate| of the root box
    const float k = 1.9073486328125e-6f;                           // 2^-19
    const float ex = k * (fabsf(ox) + fabsf(dx) + mx), ey = k * (fabsf(oy) + fabsf(dy) + mx), ez = k * (fabsf(oz) + fabsf(dz) + mx);
    const float omx = ox + ex, omy = oy + ey, omz = oz + ez;  // box.lo - delta - o = box.lo - om
    const float opx = ox - ex, opy = oy - ey, opz = oz - ez;  // box.hi + delta - o = box.hi - op
    auto slab = [&](float lx, float ly, float lz, float hx, float hy, float hz, float* t0) -> bool {
        const float ax = (lx - omx) * ix, bx = (hx - opx) * ix;
        const float ay = (ly - omy) * iy, by = (hy - opy) * iy;
        const float az = (lz - omz) * iz, bz = (hz - opz) * iz;
        const float tmin = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
        const float tmax = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), best_f));
        *t0 = tmin;
        return tmin <= tmax;
    };
    auto leaf = [&](int ref) {
        const int code = -ref - 1;
        const int first = code >> 3, count = (code & 7) + 1;
        const double* t = tris + 16 * (long long)first;
        for (int c = 0; c < count; c++, t += 16) {
            const vec3 tp = v3(t[0], t[1], t[2]), v1 = v3(t[3], t[4], t[5]), v2 = v3(t[6], t[7], t[8]);
            const vec3 n = v3(t[9], t[10], t[11]);
            const double det = -dot(d, n);
            if (det == 0.0) continue;
            const double invdet = 1.0 / det;
            const vec3 a0 = p1 - tp;
            const vec3 da0 = cross(a0, d);
            const double u = dot(v2, da0) * invdet;
            const double v = -dot(v1, da0) * invdet;
            const double alpha = dot(a0, n) * invdet;
            if ((u + v > 1.0) | (u < 0) | (v < 0) | (alpha < 0)) continue;
            const long long id = (long long)t[12];
            if (alpha >= tol && (alpha < best || (alpha == best && id < best_id))) {
                best = alpha;
                best_id = id;
                best_rec = first + c;
                best_f = __double2float_ru(alpha);
            }
        }
    };
    int2 stk[48];  // (child reference, entry parameter as float bits): one 8-byte local access per push / pop
    int sp = 0;
    // entries whose box starts behind the best hit so far are dropped on pop
    auto pop = [&]() -> int {
        while (sp > 0) {
            const int2 e = stk[--sp];
            if (__int_as_float(e.y) <= best_f) return e.x;
        }
        return RPX_BVH32_NONE;
    };
    // "while-while" walk: all lanes of a warp descend through inner nodes together until each holds a leaf (or is
    // finished), then all test their leaf together -- ONE site for the box code and ONE for the fp64 triangle code
    // (ncu of the first version, which tested leaves where it met them: 9.2 of 32 threads per instruction)
    int cur = 0;
    for (;;) {
        while (cur >= 0) {
            const float4* nd = nodes + 4 * (long long)cur;
            const float4 A = nd[0], B = nd[1], C = nd[2], D = nd[3];
            // A = lo0.xyz hi0.x | B = hi0.yz lo1.xy | C = lo1.z hi1.xyz | D = ref0 ref1 (bits) . .
            float ta, tb;
            int ra = __float_as_int(D.x), rb = __float_as_int(D.y);
            const bool ha = slab(A.x, A.y, A.z, A.w, B.x, B.y, &ta) && ra != RPX_BVH32_NONE;
            const bool hb = slab(B.z, B.w, C.x, C.y, C.z, C.w, &tb) && rb != RPX_BVH32_NONE;
            if (ha && hb) {
                const bool a_first = ta <= tb;  // near child next, far child on the stack
                if (sp < 48) stk[sp++] = a_first ? make_int2(rb, __float_as_int(tb)) : make_int2(ra, __float_as_int(ta));
                cur = a_first ? ra : rb;        // (depth checked by rpx_scene_set)
            } else if (ha) {
                cur = ra;
            } else if (hb) {
                cur = rb;
            } else {
                cur = pop();
            }
        }
        if (cur == RPX_BVH32_NONE) break;
        leaf(cur);
        cur = pop();
    }
    *piece = (int)best_id;
    if (rec_out) *rec_out = best_rec;
    if (best_id < 0) best = -1.0;
    return best * dmag;
}

// the triangle test of mesh_intersect on ONE known record (the facet the trace-ahead found): same arithmetic, so
// the same alpha; *piece = its cell id, -1 if the segment does not cross it (cannot happen for a recorded hit)
static __device__ __noinline__ double mesh_single(const DevScene& S, const rpx_face* f, vec3 p1, vec3 p2, int rec, int* piece) {
    const double* H = S.pool + f->aux_off;
    const double* t = H + (long long)H[5] + 16 * (long long)rec;
    const vec3 d = p2 - p1;
    const double dmag = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
    const double tol = H[7] / dmag;
    *piece = -1;
    const vec3 tp = v3(t[0], t[1], t[2]), v1 = v3(t[3], t[4], t[5]), v2 = v3(t[6], t[7], t[8]);
    const vec3 n = v3(t[9], t[10], t[11]);
    const double det = -dot(d, n);
    if (det == 0.0) return -dmag;
    const double invdet = 1.0 / det;
    const vec3 a0 = p1 - tp;
    const vec3 da0 = cross(a0, d);
    const double u = dot(v2, da0) * invdet;
    const double v = -dot(v1, da0) * invdet;
    const double alpha = dot(a0, n) * invdet;
    if ((u + v > 1.0) | (u < 0) | (v < 0) | (alpha < 0) | !(alpha >= tol) | !(alpha < 1.0)) return -dmag;
    *piece = (int)t[12];
    return alpha * dmag;
}

// OBBTreeFace.__cinit__ (:898-908) + compute_normal_c (:935-946): the flat normal of cell `piece`
static __device__ __noinline__ vec3 mesh_normal(const DevScene& S, const rpx_face* f, int piece) {
    const double* H = S.pool + f->aux_off;
    const double* pts = H + (long long)H[3];
    const double* c = H + (long long)H[4] + 3 * (long long)piece;
    const double* q1 = pts + 3 * (long long)c[0];
    const double* q2 = pts + 3 * (long long)c[1];
    const double* q3 = pts + 3 * (long long)c[2];
    const vec3 a = v3(q1[0], q1[1], q1[2]);
    return norm(cross(v3(q2[0], q2[1], q2[2]) - a, v3(q3[0], q3[1], q3[2]) - a));
}

// ------------------------------------------------------------------ UV patch faces (cbezier.pyx)
// UVPatchFace.intersect_c (cbezier.pyx:459-528): the nearest facet of the patch's tessellation (the mesh
// walk above) gives a seed (u, v) by barycentric interpolation of the vertex parameters
// (interpolate_cell_c, :421-456); a Newton iteration on the patch itself (<= 100 steps, both steps below
// atol) polishes it; the hit is the patch point at the final (u, v).  BezierPatch (:200-286) is evaluated
// with the reference's pow() products, BSplinePatch (:290-388) with the Cox - de Boor recursion
// (iterative here: a triangular table per direction instead of the reference's exponential recursion).
struct HitAux {
    int rec;      // position of the hit facet's record in the mesh block (-1: none); carried in the rays' side array
    int piece;    // intersect_t.piece_idx: the facet of a mesh face
    double u, v;  // intersect_t.uv: the patch parameters of a UVPatchFace hit
};

struct UVPatch {
    int kind, N, M, udeg, vdeg;
    const double *uvs, *ctrl, *a, *b;  // a, b: binomials (Bezier) or knots (B-spline)
};

RPX_DEV UVPatch uvpatch_of(const DevScene& S, const rpx_face* f) {
    UVPatch P;
    const double* H = S.pool + f->aux_off;
    const long long n_points = (long long)H[0];
    P.kind = (int)f->p[2];
    P.N = (int)f->p[3];
    P.M = (int)f->p[4];
    P.udeg = (int)f->p[5];
    P.vdeg = (int)f->p[6];
    P.uvs = S.pool + (long long)f->p[7];
    P.ctrl = P.uvs + 2 * n_points;
    P.a = P.ctrl + 3 * (P.N + 1) * (P.M + 1);
    P.b = P.a + (P.kind == 0 ? P.N + 1 : (long long)f->p[8]);
    return P;
}

// B-spline basis N_{idx,p}(t) and its derivative for ONE idx (cbezier.pyx:47-104), bottom-up: the
// recursion only ever touches N_{k,q} for idx <= k <= idx + p - q.
#define RPX_BSPLINE_MAX_DEG 8
static __device__ __noinline__ void bspline_basis(double t, int p, int idx, const double* knots, double* N_out,
                                                  double* dN_out) {
    double N[RPX_BSPLINE_MAX_DEG + 1], D[RPX_BSPLINE_MAX_DEG + 1];
    for (int k = 0; k <= p; k++) {
        N[k] = (knots[idx + k] <= t && t < knots[idx + k + 1]) ? 1.0 : 0.0;
        D[k] = 0.0;
    }
    for (int q = 1; q <= p; q++)
        for (int k = 0; k <= p - q; k++) {
            const int i = idx + k;
            double n = 0.0, d = 0.0;
            double denom = knots[i + q] - knots[i];
            if (denom != 0.0) {
                const double nom = (t - knots[i]) / denom;
                n = nom * N[k];
                d = (1.0 / denom) * N[k] + nom * D[k];
            }
            denom = knots[i + q + 1] - knots[i + 1];
            if (denom != 0.0) {
                const double nom = (knots[i + q + 1] - t) / denom;
                n += nom * N[k + 1];
                d += (-1.0 / denom) * N[k + 1] + nom * D[k + 1];
            }
            N[k] = n;
            D[k] = d;
        }
    *N_out = N[0];
    *dN_out = D[0];
}

// _eval_pt_and_grads (:259-286, :363-388); grads == false: _eval_pt (:232-254, :335-358)
static __device__ __noinline__ void uvpatch_eval(const UVPatch& P, double u, double v, bool grads, vec3* p_out,
                                                 vec3* du_out, vec3* dv_out) {
    vec3 p = v3(0, 0, 0), pu = v3(0, 0, 0), pv = v3(0, 0, 0);
    const int N = P.N, M = P.M;
    for (int i = 0; i <= N; i++) {
        double ut, dut = 0.0;
        if (P.kind == 0) {
            ut = P.a[i] * pow(u, (double)i) * pow(1 - u, (double)(N - i));
            if (grads) dut = P.a[i] * (i - N * u) * pow(u, (double)(i - 1)) * pow(1 - u, (double)(N - 1 - i));
        } else {
            bspline_basis(u, P.udeg, i, P.a, &ut, &dut);
        }
        for (int j = 0; j <= M; j++) {
            double vt, dvt = 0.0;
            if (P.kind == 0) {
                vt = P.b[j] * pow(v, (double)j) * pow(1 - v, (double)(M - j));
                if (grads) dvt = P.b[j] * (j - M * v) * pow(v, (double)(j - 1)) * pow(1 - v, (double)(M - 1 - j));
            } else {
                bspline_basis(v, P.vdeg, j, P.b, &vt, &dvt);
            }
            const vec3 c = ld3(P.ctrl + 3 * (i * (M + 1) + j));
            p = p + c * (ut * vt);
            if (grads) {
                pu = pu + c * (dut * vt);
                pv = pv + c * (ut * dvt);
            }
        }
    }
    *p_out = p;
    if (grads) {
        *du_out = pu;
        *dv_out = pv;
    }
}

// known_rec >= 0: the facet was found by the launch that traced this ray ahead (side array of the collection):
// its record gives the same alpha and cell without walking the BVH again
static __device__ __noinline__ double uvpatch_intersect(const DevScene& S, const rpx_face* f, vec3 p1, vec3 p2, HitAux* aux,
                                                        int known_rec = -1) {
    int cell = -1, rec = -1;
    double dist_mesh;
    if (known_rec >= 0) {
        dist_mesh = mesh_single(S, f, p1, p2, known_rec, &cell);
        rec = known_rec;
    } else {
        dist_mesh = mesh_intersect(S, f, p1, p2, &cell, &rec);
    }
    if (aux) aux->rec = rec;
    const vec3 seg = p2 - p1;
    const double dmag = sqrt(seg.x * seg.x + seg.y * seg.y + seg.z * seg.z);
    const double alpha = (cell < 0) ? -1.0 : dist_mesh / dmag;
    if (alpha < f->tolerance) return -1.0;
    const double* H = S.pool + f->aux_off;
    const double* pts = H + (long long)H[3];
    const double* c = H + (long long)H[4] + 3 * (long long)cell;
    const long long i0 = (long long)c[0], i1 = (long long)c[1], i2 = (long long)c[2];
    const UVPatch P = uvpatch_of(S, f);
    const double tol = f->p[0] * f->p[0];
    const vec3 d = seg * (1.0 / dmag);
    vec3 pt = p2 * alpha + p1 * (1.0 - alpha);
    const vec3 q1 = ld3(pts + 3 * i0);
    const vec3 edge1 = ld3(pts + 3 * i1) - q1, edge2 = ld3(pts + 3 * i2) - q1;
    const double x2 = sqrt(dot(edge1, edge1));
    const vec3 en1 = edge1 * (1.0 / x2);
    vec3 en2 = cross(en1, cross(edge1, edge2));
    en2 = en2 * (1.0 / sqrt(dot(en2, en2)));
    const double x3 = dot(edge2, en1), y3 = dot(edge2, en2);
    pt = pt - q1;
    const double px = dot(pt, en1), py = dot(pt, en2);
    double a2 = x2 * y3;
    double a1 = py * x3 - px * y3;
    const double a0 = (a1 - py * x2 + x2 * y3) / a2;
    a1 = -a1 / a2;
    a2 = py / y3;
    double u = a0 * P.uvs[2 * i0] + a1 * P.uvs[2 * i1] + a2 * P.uvs[2 * i2];
    double v = a0 * P.uvs[2 * i0 + 1] + a1 * P.uvs[2 * i1 + 1] + a2 * P.uvs[2 * i2 + 1];
    int it = 0;
    for (; it < 100; it++) {
        vec3 p, pu, pv;
        uvpatch_eval(P, u, v, true, &p, &pu, &pv);
        vec3 n = cross(pu, pv);
        n = n * (1.0 / sqrt(dot(n, n)));
        const double dist = dot(p - p1, n) / dot(d, n);
        const vec3 dp = (p1 + d * dist) - p;
        const double du = dot(pu, dp) / dot(pu, pu);
        const double dv = dot(pv, dp) / dot(pv, pv);
        u += du;
        v += dv;
        if (du * du < tol && dv * dv < tol) break;
    }
    if (it == 100) return -1.0;
    vec3 p, unused1, unused2;
    uvpatch_eval(P, u, v, false, &p, &unused1, &unused2);
    if (aux) {
        aux->u = u;
        aux->v = v;
    }
    const vec3 r = p - p1;
    return sqrt(dot(r, r));
}

// compute_normal_and_tangent_c (cbezier.pyx:533-550)
static __device__ __noinline__ void uvpatch_orientation(const DevScene& S, const rpx_face* f, const HitAux& aux, vec3* normal,
                                                        vec3* tangent) {
    const UVPatch P = uvpatch_of(S, f);
    vec3 p, pu, pv;
    uvpatch_eval(P, aux.u, aux.v, true, &p, &pu, &pv);
    vec3 n = cross(pu, pv);
    n = n * (1.0 / sqrt(dot(n, n)));
    *normal = (f->p[1] != 0.0) ? neg(n) : n;
    *tangent = pu * (1.0 / sqrt(dot(pu, pu)));
}

// aux (may be NULL): the part of intersect_t that travels from Face.intersect_c to
// compute_normal_and_tangent_c -- piece_idx for mesh faces, uv for UV patch faces

template <int FC>
RPX_DEV double face_intersect(const DevScene& S, const rpx_face* f, vec3 p1, vec3 p2, int is_base_ray,
                              HitAux* aux = nullptr) {
    if (FC == RPX_FC_MESH && f->type == RPX_FACE_MESH) {
        int pc, rec;
        const double dist = mesh_intersect(S, f, p1, p2, &pc, &rec);
        if (aux) {
            aux->piece = pc;
            aux->rec = rec;
        }
        return dist;
    }
    if (aux) {
        aux->piece = 0;
        aux->rec = -1;
    }
    if (FC == RPX_FC_MESH && f->type == RPX_FACE_UVPATCH) return uvpatch_intersect(S, f, p1, p2, aux);
    if (FC >= RPX_FC_FULL && f->type == RPX_FACE_DISTORTION) return distortion_intersect(S, f, p1, p2);
    if (FC >= RPX_FC_FULL && f->type == RPX_FACE_EXTRUDED_BEZIER) return bezier_intersect(S, f, p1, p2);
    return face_intersect_basic<FC>(S, f, p1, p2, is_base_ray);
}

// What Face.intersect_c left in intersect_t for compute_normal_and_tangent_c, rebuilt from the facet record the
// trace-ahead stored for this ray (no second BVH walk): the cell id of a mesh face, the converged (u, v) of a patch
// face.  p1, p2: the segment in the face's local frame, as the trace-ahead saw it.
RPX_DEV void face_aux_from_rec(const DevScene& S, const rpx_face* f, vec3 p1, vec3 p2, int rec, HitAux* aux) {
    aux->rec = rec;
    aux->piece = 0;
    aux->u = aux->v = 0.0;
    if (f->type == RPX_FACE_MESH) {
        const double* H = S.pool + f->aux_off;
        aux->piece = (int)H[(long long)H[5] + 16 * (long long)rec + 12];
    } else if (f->type == RPX_FACE_UVPATCH) {
        uvpatch_intersect(S, f, p1, p2, aux, rec);
    }
}

template <int FC>
__device__ vec3 face_normal(const DevScene& S, const rpx_face* f, vec3 p, int piece = 0) {
    if (FC == RPX_FC_MESH && f->type == RPX_FACE_MESH) return mesh_normal(S, f, piece);
    if (FC >= RPX_FC_FULL && f->type == RPX_FACE_DISTORTION) {  // cfaces.pyx:2418-2431
        const rpx_face* base = &S.faces[f->base_face];
        const rpx_distortion* dist = &S.dists[f->aux_off];
        vec3 dxdyz = distortion_zgrad(S, dist, p.x, p.y);
        vec3 p1 = p;
        p1.z -= dxdyz.z;
        vec3 n = face_normal_basic<RPX_FC_FULL>(S, base, p1);
        n.x /= n.z;
        n.y /= n.z;
        n.z = 1.0;
        n.x -= dxdyz.x;
        n.y -= dxdyz.y;
        return norm(n);
    }
    if (FC >= RPX_FC_FULL && f->type == RPX_FACE_EXTRUDED_BEZIER) return bezier_normal(S, f, p);
    return face_normal_basic<FC>(S, f, p);
}

RPX_DEV vec3 face_tangent(const rpx_face* f) {
    if (f->type == RPX_FACE_EXTRUDED_PLANAR) return v3(0.0, 0.0, 1.0);  // cfaces.pyx:704-709
    if (f->type == RPX_FACE_ORIENTED_POLYGON) return ld3(f->p + 6);     // :1186-1187
    return v3(1.0, 0.0, 0.0);                                           // ctracer.pyx:1786-1791
}

// FaceList.compute_orientation_c, ctracer.pyx:1939-1953
template <int FC>
RPX_DEV void compute_orientation(const DevScene& S, const rpx_face* f, vec3 point, vec3* normal,
                                 vec3* tangent, const HitAux* aux = nullptr) {
    const rpx_face_set* fs = &S.sets[f->face_set];
    point = transform_pt(fs->inv_trans.m, point);
    vec3 n, t;
    if (FC == RPX_FC_MESH && f->type == RPX_FACE_UVPATCH) {
        HitAux zero;
        zero.rec = -1;
        zero.piece = 0;
        zero.u = zero.v = 0.0;
        uvpatch_orientation(S, f, aux ? *aux : zero, &n, &t);
    } else {
        n = face_normal<FC>(S, f, point, aux ? aux->piece : 0);
        t = face_tangent(f);
    }
    if (f->invert_normal) {
        n = neg(n);
        t = neg(t);
    }
    *normal = rotate_v(fs->trans.m, n);
    *tangent = rotate_v(fs->trans.m, t);
}

}  // namespace rpx
