/*
 * rpx_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT.
 *
 * A scalar, single-threaded, plain-C99 restatement of the reference's
 * non-sequential trace (raypier/core: ctracer.pyx, cfaces.pyx, cmaterials.pyx,
 * cshapes.pyx, cdistortions.pyx, cimplicit_surfs.pyx, the OBBTreeFace of
 * obbtree.pyx) and of the consumers next to it (capture planes, sequential mode,
 * cfields.pyx + core/fields.py: the E-field summation with its gausslet and
 * plain-ray front ends) over the flattened scene tables of include/rpx.h.  It exists so the CUDA path can be checked on the GPU
 * box, where /root/reference does not exist.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * this library.  The product (raypier_optics_b200) never does.
 *
 * PARITY PIN: this oracle is pinned against the real reference (built unmodified
 * into oracle/_ref by oracle/build_ref.sh) by tests/test_oracle_vs_reference.py,
 * against the reference's own known-answer tests (tests/test_reference_kats.py)
 * and against committed golden vectors generated from the reference
 * (tests/golden/, made by tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (no FMA contraction: the op
 * order below is the reference's op order; every function cites the lines it
 * restates).  AoS records, per-ray face loop, append-with-order -- i.e. the
 * reference's own structure, deliberately NOT the wavefront/SoA design of the
 * CUDA product.
 */
#include <complex.h>
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/rpx.h"

typedef struct { double x, y, z; } vec3;
typedef struct { vec3 normal, tangent; } orient_t;
typedef double complex cplx;

#define ORACLE_INF (DBL_MAX + DBL_MAX) /* ctracer.pyx:18 */
#define NO_HIT (-1.0)                  /* NO_INTERSECTION.dist, cfaces.pyx:40-44 */
#define SP_TOL 1.0e-10                 /* cmaterials.pyx:32 */
/* Cython promotes a real operand to (r + 0i) before every mixed operation */
/* Cython builds every complex temporary with  x + y*_Complex_I  and gcc's complex
 * lowering then treats known-zero parts specially (signed zeros!).  Writing the same C
 * expression is the only way to get the same bits, so CX / cy_parts mirror the generated
 * code of the reference (oracle/_ref/build/cmaterials.c, __pyx_t_double_complex_from_parts). */
static inline double complex from_parts(double x, double y) {
    return x + y * (double complex)_Complex_I;
}
#define CX(r) from_parts((r), 0)
/* `x.real + 1.0j*x.imag` (cmaterials.pyx:783-784, 1057-1058, 1261-1262) */
static inline double complex cy_parts(double re, double im) {
    return CX(re) + (from_parts(0, 1.0) * CX(im));
}

/* ------------------------------------------------ vector maths, ctracer.pyx:83-262 */
static inline vec3 v3(double x, double y, double z) { vec3 v = {x, y, z}; return v; }
static inline vec3 ld3(const double* p) { return v3(p[0], p[1], p[2]); }
static inline void st3(double* p, vec3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
static inline vec3 addvv(vec3 a, vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline vec3 subvv(vec3 a, vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline vec3 multvs(vec3 a, double b) { return v3(a.x * b, a.y * b, a.z * b); }
static inline vec3 invert(vec3 a) { return v3(-a.x, -a.y, -a.z); }
static inline double dotprod(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline double mag_sq(vec3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
static inline double mag(vec3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
static inline vec3 cross(vec3 a, vec3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline vec3 norm(vec3 a) {
    double m = sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
    return v3(a.x / m, a.y / m, a.z / m);
}
static inline double sep(vec3 p1, vec3 p2) {
    double a = p2.x - p1.x, b = p2.y - p1.y, c = p2.z - p1.z;
    return sqrt((a * a) + (b * b) + (c * c));
}
/* transform_c / rotate_c, ctracer.pyx:83-95 */
static inline vec3 transform_pt(const rpx_transform* t, vec3 p) {
    return v3(p.x * t->m[0] + p.y * t->m[1] + p.z * t->m[2] + t->t[0],
              p.x * t->m[3] + p.y * t->m[4] + p.z * t->m[5] + t->t[1],
              p.x * t->m[6] + p.y * t->m[7] + p.z * t->m[8] + t->t[2]);
}
static inline vec3 rotate_v(const rpx_transform* t, vec3 p) {
    return v3(p.x * t->m[0] + p.y * t->m[1] + p.z * t->m[2],
              p.x * t->m[3] + p.y * t->m[4] + p.z * t->m[5],
              p.x * t->m[6] + p.y * t->m[7] + p.z * t->m[8]);
}

/* ------------------------------------------------------- shapes, cshapes.pyx */
/* point_in_polygon of PolygonShape, cshapes.pyx:150-168 */
static int shape_polygon_inside(const double* pts, int size, double X, double Y) {
    int ct = 0;
    double y1 = pts[2 * (size - 1) + 1], x1 = pts[2 * (size - 1)];
    for (int i = 0; i < size; i++) {
        double y2 = pts[2 * i + 1], x2 = pts[2 * i];
        if ((y1 <= Y && Y < y2) || (y2 <= Y && Y < y1)) {
            if ((x1 + (Y - y1) * (x2 - x1) / (y2 - y1)) > X) ct = !ct;
        }
        y1 = y2;
        x1 = x2;
    }
    return ct;
}

static int shape_inside(const rpx_scene* S, const rpx_face* f, double x, double y) {
    if (f->shape_off < 0) return 1; /* base Shape, ctracer.pyx:1659-1661 */
    int stack[32];
    int sp = 0;
    for (int i = 0; i < f->shape_len; i++) {
        const rpx_shape_op* op = &S->shape_ops[f->shape_off + i];
        switch (op->type) {
            case RPX_SHAPE_TRUE: stack[sp++] = 1; break;
            case RPX_SHAPE_CIRCLE: { /* cshapes.pyx:109-116 */
                double dx = x - op->p[0], dy = y - op->p[1];
                stack[sp++] = ((dx * dx) + (dy * dy) < (op->p[2] * op->p[2])) ? 1 : 0;
            } break;
            case RPX_SHAPE_RECT: { /* cshapes.pyx:128-136 */
                double dx = x - op->p[0], dy = y - op->p[1];
                stack[sp++] = ((2 * fabs(dx) < op->p[2]) && (2 * fabs(dy)) < op->p[3]) ? 1 : 0;
            } break;
            case RPX_SHAPE_POLYGON:
                stack[sp++] = shape_polygon_inside(S->pool + op->aux_off, op->aux_n, x, y);
                break;
            case RPX_SHAPE_NOT: stack[sp - 1] = 1 & (~stack[sp - 1]); break; /* :44 */
            case RPX_SHAPE_AND: sp--; stack[sp - 1] = stack[sp - 1] & stack[sp]; break;
            case RPX_SHAPE_OR: sp--; stack[sp - 1] = stack[sp - 1] | stack[sp]; break;
            case RPX_SHAPE_XOR: sp--; stack[sp - 1] = stack[sp - 1] ^ stack[sp]; break;
            default: return 0;
        }
    }
    return stack[0];
}

/* ------------------------------------- implicit surfaces, cimplicit_surfs.pyx */
static double implicit_eval(const rpx_scene* S, int off, int len, vec3 p) {
    double stack[32];
    int sp = 0;
    for (int i = 0; i < len; i++) {
        const rpx_implicit_op* op = &S->implicit_ops[off + i];
        switch (op->type) {
            case RPX_IMPL_NULL: stack[sp++] = -1.0; break; /* :25 */
            case RPX_IMPL_PLANE:                           /* :57 */
                stack[sp++] = dotprod(ld3(op->p + 3), subvv(p, ld3(op->p)));
                break;
            case RPX_IMPL_SPHERE: /* :80 */
                stack[sp++] = sep(p, ld3(op->p)) - op->p[3];
                break;
            case RPX_IMPL_CYLINDER: /* :119 */
                stack[sp++] = mag(cross(subvv(p, ld3(op->p)), ld3(op->p + 3))) - op->p[6];
                break;
            case RPX_IMPL_NEG: stack[sp - 1] = -stack[sp - 1]; break;
            case RPX_IMPL_MIN: sp--; if (stack[sp] < stack[sp - 1]) stack[sp - 1] = stack[sp]; break;
            case RPX_IMPL_MAX: sp--; if (stack[sp] > stack[sp - 1]) stack[sp - 1] = stack[sp]; break;
            case RPX_IMPL_SUB: sp--; stack[sp - 1] -= stack[sp]; break;
            default: return 1.0;
        }
    }
    return stack[0];
}

/* ------------------------------------------- distortions, cdistortions.pyx */
typedef struct { double* w[3]; } zws_t; /* workspace[3][k_max] */

/* zernike_R_c, cdistortions.pyx:149-178 (memoised recursion, C integer division) */
static double zernike_R(double r, int k, int n, int m, zws_t* ws) {
    if (n < m) return 0.0;
    if (n == 0) return 1.0;
    double val = ws->w[0][k];
    if (isnan(val)) {
        int nA = n - 1, mA = abs(m - 1);
        int half_n = nA / 2;
        int kA = half_n * (half_n + 1) + abs(mA);
        int nB = nA, mB = m + 1;
        int kB = half_n * (half_n + 1) + abs(mB);
        int nC = n - 2, mC = m;
        half_n = nC / 2;
        int kC = half_n * (half_n + 1) + abs(mC);
        val = r * (zernike_R(r, kA, nA, mA, ws) + zernike_R(r, kB, nB, mB, ws));
        val -= zernike_R(r, kC, nC, mC, ws);
        ws->w[0][k] = val;
        return val;
    }
    return val;
}

/* zernike_Rprime_c, cdistortions.pyx:189-219 (kC uses the stale half_n: quirk kept) */
static double zernike_Rprime(double r, int k, int n, int m, zws_t* ws) {
    if (n < m) return 0.0;
    if (n == 0) return 0.0;
    double val = ws->w[1][k];
    if (isnan(val)) {
        int nA = n - 1, mA = abs(m - 1);
        int half_n = nA / 2;
        int kA = half_n * (half_n + 1) + abs(mA);
        int nB = nA, mB = m + 1;
        int kB = half_n * (half_n + 1) + abs(mB);
        int nC = n - 2, mC = m;
        int kC = (half_n - 1) * half_n + abs(mC);
        val = zernike_R(r, kA, nA, mA, ws) + zernike_R(r, kB, nB, mB, ws);
        val += r * (zernike_Rprime(r, kA, nA, mA, ws) + zernike_Rprime(r, kB, nB, mB, ws));
        val -= zernike_Rprime(r, kC, nC, mC, ws);
        ws->w[1][k] = val;
        return val;
    }
    return val;
}

/* zernike_R_over_r_c, cdistortions.pyx:288-316 */
static double zernike_R_over_r(double r, int k, int n, int m, zws_t* ws) {
    if (n < m) return 0.0;
    double val = ws->w[2][k];
    if (isnan(val)) {
        int nA = n - 1, mA = abs(m - 1);
        int half_n = nA / 2;
        int kA = half_n * (half_n + 1) + abs(mA);
        int nB = nA, mB = m + 1;
        int kB = half_n * (half_n + 1) + abs(mB);
        int nC = n - 2, mC = m;
        half_n = nC / 2;
        int kC = half_n * (half_n + 1) + abs(mC);
        val = zernike_R(r, kA, nA, mA, ws) + zernike_R(r, kB, nB, mB, ws);
        val -= zernike_R_over_r(r, kC, nC, mC, ws);
        ws->w[2][k] = val;
    }
    return val;
}

#define ORACLE_ZK 256
/* Distortion.z_offset_c */
static double distortion_z(const rpx_scene* S, const rpx_distortion* D, double x, double y) {
    if (D->type == RPX_DIST_ZERNIKE_J7) { /* cdistortions.pyx:50-59 */
        x /= D->p[0];
        y /= D->p[0];
        double Z = sqrt(8.0) * (3 * (x * x + y * y) - 2) * y;
        return Z * D->p[1];
    }
    /* ZernikeDistortion.z_offset_c, cdistortions.pyx:416-450 */
    double w0[ORACLE_ZK], w1[ORACLE_ZK], w2[ORACLE_ZK];
    zws_t ws = {{w0, w1, w2}};
    x /= D->p[0];
    y /= D->p[0];
    double r = sqrt(x * x + y * y);
    double theta = atan2(y, x);
    double Z = 0.0;
    for (int i = 0; i < D->k_max; i++) w0[i] = NAN;
    w0[0] = 1.0;
    for (int i = 0; i < D->n_coefs; i++) {
        const rpx_zcoef* c = &S->zcoefs[D->coef_off + i];
        double N, PH, R;
        if (c->m == 0) N = sqrt(c->n + 1);
        else N = sqrt(2 * (c->n + 1));
        N *= c->value;
        if (c->m >= 0) PH = cos(c->m * theta);
        else PH = -sin(c->m * theta);
        R = zernike_R(r, c->k, c->n, abs(c->m), &ws);
        Z += N * R * PH;
    }
    return Z;
}

/* Distortion.z_offset_and_gradient_c -> (dz/dx, dz/dy, z) */
static vec3 distortion_zgrad(const rpx_scene* S, const rpx_distortion* D, double x, double y) {
    vec3 p;
    if (D->type == RPX_DIST_ZERNIKE_J7) { /* cdistortions.pyx:61-81 */
        double root8 = sqrt(8) * D->p[1], R = D->p[0];
        x /= R;
        y /= R;
        p.x = root8 * 6 * x * y / R;
        p.y = root8 * (3 * x * x + 9 * y * y - 2) / R;
        p.z = root8 * (3 * (x * x + y * y) - 2) * y;
        return p;
    }
    /* ZernikeDistortion.z_offset_and_gradient_c, cdistortions.pyx:456-514 */
    double w0[ORACLE_ZK], w1[ORACLE_ZK], w2[ORACLE_ZK];
    zws_t ws = {{w0, w1, w2}};
    x /= D->p[0];
    y /= D->p[0];
    double r = sqrt(x * x + y * y);
    double theta = atan2(y, x);
    for (int i = 0; i < D->k_max; i++) { w0[i] = NAN; w1[i] = NAN; w2[i] = NAN; }
    vec3 Z = {0.0, 0.0, 0.0};
    for (int i = 0; i < D->n_coefs; i++) {
        const rpx_zcoef* c = &S->zcoefs[D->coef_off + i];
        double PH, PHprime, N;
        if (c->m >= 0) {
            PH = cos(c->m * theta);
            PHprime = -c->m * sin(c->m * theta);
        } else {
            PH = -sin(c->m * theta);
            PHprime = -c->m * cos(c->m * theta);
        }
        double R = zernike_R(r, c->k, c->n, abs(c->m), &ws);
        double Rprime = zernike_Rprime(r, c->k, c->n, abs(c->m), &ws);
        double R_over_r = zernike_R_over_r(r, c->k, c->n, abs(c->m), &ws);
        if (c->m == 0) N = sqrt(c->n + 1);
        else N = sqrt(2 * (c->n + 1));
        N *= c->value;
        Z.z += N * R * PH;
        Z.x += N * (Rprime * cos(theta) * PH + R_over_r * (-sin(theta)) * PHprime);
        Z.y += N * (Rprime * sin(theta) * PH + R_over_r * (cos(theta)) * PHprime);
    }
    Z.x /= D->p[0];
    Z.y /= D->p[0];
    return Z;
}

/* ------------------------------------------------------ faces, cfaces.pyx */
/* point_in_polygon_c, cfaces.pyx:1050-1068 */
static int point_in_polygon(double X, double Y, const double* pts, int size) {
    int ct = 0;
    double y1 = pts[2 * (size - 1) + 1], x1 = pts[2 * (size - 1)];
    for (int i = 0; i < size; i++) {
        double y2 = pts[2 * i + 1], x2 = pts[2 * i];
        double h = (Y - y1) / (y2 - y1);
        if (0 < h && h <= 1.0) {
            double x = x1 + h * (x2 - x1);
            if (x > X) ct = !ct;
        }
        y1 = y2;
        x1 = x2;
    }
    return ct;
}

/* intersect_conic, cfaces.pyx:1695-1747 */
static double intersect_conic(vec3 a, vec3 d, double curvature, double conic_const) {
    double beta = 1 + conic_const;
    double R = -curvature;
    double A = pow(beta, 2.0) * pow(d.z, 2.0) + beta * pow(d.x, 2.0) + beta * pow(d.y, 2.0);
    double B = -2 * R * beta * d.z + 2 * a.x * beta * d.x + 2 * a.y * beta * d.y +
               2 * a.z * pow(beta, 2.0) * d.z;
    double C = -2 * R * a.z * beta + pow(a.x, 2.0) * beta + pow(a.y, 2.0) * beta +
               pow(a.z, 2.0) * pow(beta, 2.0);
    double D = B * B - 4 * A * C;
    if (D < 0) return -1;
    D = sqrt(D);
    if (R * beta * d.z <= 0) return (-B + D) / (2 * A);
    return (-B - D) / (2 * A);
}

typedef struct { double R, beta, A4, A6, A8, A10, A12, A14, A16; vec3 a, d; } aspheric_t;

/* eval_aspheric_impf, cfaces.pyx:1854-1863 */
static double aspheric_impf(const aspheric_t* A, double alpha) {
    double r2 = (pow(A->a.x + alpha * A->d.x, 2.0) + pow(A->a.y + alpha * A->d.y, 2.0));
    double out = r2;
    out /= A->R * (1 + sqrt(1 - A->beta * r2 / (pow(A->R, 2.0))));
    out -= A->a.z + alpha * A->d.z;
    out += A->A4 * pow(r2, 2.0) + A->A6 * pow(r2, 3.0) + A->A8 * pow(r2, 4.0) +
           A->A10 * pow(r2, 5.0) + A->A12 * pow(r2, 6.0) + A->A14 * pow(r2, 7.0) +
           A->A16 * pow(r2, 8.0);
    return out;
}

/* eval_aspheric_grad, cfaces.pyx:1866-1882 */
static double aspheric_grad(const aspheric_t* A, double alpha) {
    double r2 = (pow(A->a.x + alpha * A->d.x, 2.0) + pow(A->a.y + alpha * A->d.y, 2.0));
    double dx = A->d.x * (A->a.x + alpha * A->d.x);
    double dy = A->d.y * (A->a.y + alpha * A->d.y);
    double out = A->A10 * (10 * dx + 10 * dy) * pow(r2, 4.0);
    out += A->A12 * (12 * dx + 12 * dy) * pow(r2, 5.0);
    out += A->A14 * (14 * dx + 14 * dy) * pow(r2, 6.0);
    out += A->A16 * (16 * dx + 16 * dy) * pow(r2, 7.0);
    out += A->A4 * (4 * dx + 4 * dy) * (r2);
    out += A->A6 * (6 * dx + 6 * dy) * pow(r2, 2.0);
    out += A->A8 * (8 * dx + 8 * dy) * pow(r2, 3.0) - A->d.z;
    out += (2 * dx + 2 * dy) / (A->R * (sqrt(1 - A->beta * (r2) / pow(A->R, 2.0)) + 1));
    out += A->beta * (2 * dx + 2 * dy) * (r2) /
           (2 * pow(A->R, 3.0) * sqrt(1 - A->beta * r2 / pow(A->R, 2.0)) *
            pow(sqrt(1 - A->beta * r2 / pow(A->R, 2.0)) + 1, 2.0));
    return out;
}

/* eval_extpoly_impf, cfaces.pyx:2043-2081 */
static double extpoly_impf(const rpx_face* f, const double* E, vec3 a, vec3 d, double alpha) {
    double R = f->p[0], beta = f->p[1], norm_radius = f->p[2], z_height = f->p[3];
    int Nx = f->aux_n, Ny = f->aux_m;
    double x = a.x + alpha * d.x;
    double y = a.y + alpha * d.y;
    double r2 = pow(x, 2.0) + pow(y, 2.0);
    double out = r2;
    if (R >= 0) out /= (R + sqrt(R * R - beta * r2));
    else out /= (R - sqrt(R * R - beta * r2));
    out -= a.z + alpha * d.z;
    x /= norm_radius;
    y /= norm_radius;
    for (int i = 0; i < Nx; i++)
        for (int j = 0; j < Ny; j++) out += E[i * Ny + j] * pow(x, (double)i) * pow(y, (double)j);
    out += z_height;
    return out;
}

/* eval_extpoly_grad, cfaces.pyx:2084-2126 */
static double extpoly_grad(const rpx_face* f, const double* E, vec3 a, vec3 d, double alpha) {
    double R = f->p[0], beta = f->p[1], norm_radius = f->p[2];
    int Nx = f->aux_n, Ny = f->aux_m;
    double dEdx = 0.0, dEdy = 0.0;
    double x = a.x + alpha * d.x;
    double y = a.y + alpha * d.y;
    double r2 = pow(x, 2.0) + pow(y, 2.0);
    double R2 = R * R;
    double rt = sqrt(1 - (beta * r2 / R2));
    double denom = R * (rt + 1);
    double nom = (2 * d.x * x + 2 * d.y * y);
    double inv_rad = 1. / norm_radius;
    double out = -d.z;
    out += nom / denom;
    out += beta * nom * r2 / (2 * R * rt * denom * denom);
    x *= inv_rad;
    y *= inv_rad;
    for (int i = 1; i < Nx; i++)
        for (int j = 0; j < Ny; j++)
            dEdx += (i)*E[i * Ny + j] * pow(x, (double)(i - 1)) * pow(y, (double)j);
    for (int i = 0; i < Nx; i++)
        for (int j = 1; j < Ny; j++)
            dEdy += (j)*E[i * Ny + j] * pow(x, (double)i) * pow(y, (double)(j - 1));
    dEdx *= inv_rad;
    dEdy *= inv_rad;
    out += dEdx * d.x;
    out += dEdy * d.y;
    return out;
}

static vec3 face_normal(const rpx_scene* S, const rpx_face* f, vec3 p);

/* ---- ExtrudedBezierFace helpers, cfaces.pyx:717-845 ---- */
typedef struct { double x, y; } flat2;
typedef struct { double roots[3]; int n; } poly_roots;

/* eval_bezier, cfaces.pyx:717-719 */
static double eval_bezier(double t, double cp0, double cp1, double cp2, double cp3) {
    return cp0 * pow(1 - t, 3.0) + 3 * cp1 * t * pow(1 - t, 2.0) + 3 * cp2 * (1 - t) * pow(t, 2.0) +
           cp3 * pow(t, 3.0);
}
/* dif_bezier, cfaces.pyx:721-727 (long double coefficients) */
static double dif_bezier(double t, double cp0, double cp1, double cp2, double cp3) {
    long double A, B, C;
    A = cp3 - 3 * cp2 + 3 * cp1 - cp0;
    B = 3 * cp2 - 6 * cp1 + 3 * cp0;
    C = 3 * cp1 - 3 * cp0;
    return 3 * A * pow(t, 2.0) + 2 * B * t + C;
}
/* roots_of_cubic, cfaces.pyx:735-787: x87 long double intermediates; the single-real-root
 * branch divides by roots[0] == 0.0 and never yields a usable root (quirk Q8). */
static poly_roots roots_of_cubic(double a, double b, double c, double d) {
    long double a1 = b / a, a2 = c / a, a3 = d / a;
    long double Q = (a1 * a1 - 3.0 * a2) / 9.0;
    long double R = (2.0 * a1 * a1 * a1 - 9.0 * a1 * a2 + 27.0 * a3) / 54.0;
    long double R2_Q3 = R * R - Q * Q * Q;
    long double theta;
    poly_roots x = {{0.0, 0.0, 0.0}, 0};
    if (fabs(a) <= 0.0000000001) {
        if (fabs(b) <= 0.0000000001) {
            if (c == 0) {
                x.n = 1;
                x.roots[0] = 0;
            } else {
                x.n = 1;
                x.roots[0] = -d / c;
            }
        } else {
            a1 = pow(c, 2.0) - 4 * b * d;
            a1 = sqrt((double)a1);
            x.n = 2;
            x.roots[0] = (-c + a1) / (2 * b);
            x.roots[1] = (-c - a1) / (2 * b);
        }
    } else {
        if (R2_Q3 < 0) {
            x.n = 3;
            theta = acos((double)(R / sqrt((double)(Q * Q * Q))));
            x.roots[0] = -2.0 * sqrt((double)Q) * cos((double)(theta / 3.0)) - a1 / 3.0;
            x.roots[1] = -2.0 * sqrt((double)Q) * cos((double)((theta + 2.0 * M_PI) / 3.0)) - a1 / 3.0;
            x.roots[2] = -2.0 * sqrt((double)Q) * cos((double)((theta + 4.0 * M_PI) / 3.0)) - a1 / 3.0;
        } else {
            x.n = 1;
            a2 = pow(sqrt((double)R2_Q3) + fabs((double)R), 1 / 3.0);
            a2 += Q / x.roots[0];
            a2 *= ((R < 0.0) ? 1 : -1);
            a2 -= a1 / 3.0;
            x.roots[0] = a2;
        }
    }
    return x;
}
static flat2 rotate2D(double phi, flat2 p) { /* cfaces.pyx:789-793 */
    flat2 r;
    r.x = p.x * cos(phi) - p.y * sin(phi);
    r.y = p.x * sin(phi) + p.y * cos(phi);
    return r;
}
static int bz_ccw(flat2 A, flat2 B, flat2 C) { return (C.y - A.y) * (B.x - A.x) > (B.y - A.y) * (C.x - A.x); }
static int bz_seg_overlap(flat2 A, flat2 B, flat2 C, flat2 D) {
    return bz_ccw(A, C, D) != bz_ccw(B, C, D) && bz_ccw(A, B, C) != bz_ccw(A, B, D);
}
static int bz_pnt_in_hull(flat2 p, flat2 A, flat2 B, flat2 C, flat2 D) { /* cfaces.pyx:815-845 (float w) */
    int i, j, k;
    float w;
    i = p.x > A.x || p.x > B.x || p.x > C.x || p.x > D.x;
    j = p.x > A.x && p.x > B.x && p.x > C.x && p.x > D.x;
    k = i && !j;
    i = p.y > A.y || p.y > B.y || p.y > C.y || p.y > D.y;
    j = p.y > A.y && p.y > B.y && p.y > C.y && p.y > D.y;
    i = i && !j;
    w = A.x - D.x;
    w = w * w;
    if (w <= .0005) {
        w = B.x - A.x;
        w = w * w;
        if (w <= .0005) i = k = 1;
    } else {
        w = A.y - D.y;
        w = w * w;
        if (w <= .0005) {
            w = B.y - A.y;
            w = w * w;
            if (w <= .0005) i = k = 1;
        }
    }
    return i && k;
}

/* ---- triangle meshes: OBBTree / OBBTreeFace, raypier/core/obbtree.pyx -------------------------
 * intersect_t.piece_idx (ctracer.pxd:77-80) travels from Face.intersect_c to compute_normal_c in the
 * reference; every other face class ignores it.  This scalar restatement keeps it in two statics:
 * s_piece = piece of the last face_intersect call, s_hit_piece = piece of the accepted hit.
 * Thread-local: the full-size parity tests run independent shards of one trace on several host threads. */
static __thread int s_piece = 0, s_hit_piece = 0;
/* intersect_t.uv (ctracer.pxd:74-81) travels the same way for UVPatchFace (cbezier.pyx:514-516, 538-539) */
static __thread double s_u = 0, s_v = 0, s_hit_u = 0, s_hit_v = 0;

/* Mesh block in the pool (scene.py::_mesh_block): header of 8 doubles, then the raw points / cells
 * (what the reference object holds) and, for the CUDA path only, BVH-ordered triangle records and
 * nodes.  The oracle reads ONLY the raw points and cells.                                        */
#define MESH_HDR 8

/* OBBTree.line_intersects_cell_c, obbtree.pyx:310-343 */
static double mesh_line_intersects_cell(const double* points, const double* cells, long cell_idx, vec3 o, vec3 d) {
    const double* c = cells + 3 * cell_idx;
    vec3 p1 = ld3(points + 3 * (long)c[0]), p2 = ld3(points + 3 * (long)c[1]), p3 = ld3(points + 3 * (long)c[2]);
    vec3 v1 = subvv(p2, p1), v2 = subvv(p3, p1);
    vec3 n = cross(v1, v2);
    double det = -dotprod(d, n);
    if (det == 0.0) return -1;
    double invdet = 1.0 / det;
    vec3 a0 = subvv(o, p1);
    vec3 da0 = cross(a0, d);
    double u = dotprod(v2, da0) * invdet;
    double v = -dotprod(v1, da0) * invdet;
    double alpha = dotprod(a0, n) * invdet;
    if ((u + v > 1.0) | (u < 0) | (v < 0) | (alpha < 0)) return -1.0;
    return alpha;
}

/* OBBTree.intersect_with_line_c (obbtree.pyx:367-400) + OBBTreeFace.intersect_c (:913-932).  The OBB
 * tree only prunes (line_intersects_node_c, :271-296, is a conservative interval test with a positive
 * margin), so the result is the nearest cell with tol <= alpha < 1 over ALL cells; on exactly equal
 * alpha the reference keeps the cell its traversal meets first, this loop the lowest cell index.  */
static double mesh_nearest(const rpx_scene* S, const rpx_face* f, vec3 p1, vec3 p2, int* piece);
static double mesh_intersect(const rpx_scene* S, const rpx_face* f, vec3 p1, vec3 p2, int* piece) {
    return mesh_nearest(S, f, p1, p2, piece) * mag(subvv(p2, p1));
}

/* OBBTree.intersect_with_line_c itself: alpha of the nearest cell, -1 if there is none */
static double mesh_nearest(const rpx_scene* S, const rpx_face* f, vec3 p1, vec3 p2, int* piece) {
    const double* H = S->pool + f->aux_off;
    const long n_cells = (long)H[1];
    const double* points = H + (long)H[3];
    const double* cells = H + (long)H[4];
    vec3 d = subvv(p2, p1);
    double tol = H[7] / mag(d);
    double best = 1.0;
    long best_cell = -1;
    for (long i = 0; i < n_cells; i++) {
        double alpha = mesh_line_intersects_cell(points, cells, i, p1, d);
        if ((alpha >= tol) && (alpha < best)) {
            best = alpha;
            best_cell = i;
        }
    }
    if (best_cell < 0) best = -1.0;
    *piece = (int)best_cell;
    return best;
}

/* OBBTreeFace.__cinit__ (:898-908) + compute_normal_c (:935-946): the flat normal of cell `piece` */
static vec3 mesh_normal(const rpx_scene* S, const rpx_face* f, int piece) {
    const double* H = S->pool + f->aux_off;
    const double* points = H + (long)H[3];
    const double* c = H + (long)H[4] + 3 * (long)piece;
    vec3 p1 = ld3(points + 3 * (long)c[0]), p2 = ld3(points + 3 * (long)c[1]), p3 = ld3(points + 3 * (long)c[2]);
    return norm(cross(subvv(p2, p1), subvv(p3, p1)));
}

/* ---- UV patch faces: raypier/core/cbezier.pyx -------------------------------------------------
 * UVPatchFace (:391-550) over a BezierPatch (:200-286) or BSplinePatch (:290-388).  Face record:
 * p[0] atol, p[1] invert_normals, p[2] patch kind (0 Bezier, 1 B-spline), p[3] N, p[4] M (orders:
 * (N+1) x (M+1) control points), p[5] u_degree, p[6] v_degree, p[7] offset of the patch block in the
 * pool, p[8] / p[9] number of u / v knots.  aux_off = mesh block of the (u_res x v_res) tessellation
 * (get_mesh, :153-197), the same layout as RPX_FACE_MESH.  Patch block: uvs[n_points][2],
 * ctrl[N+1][M+1][3], then binom_n[N+1], binom_m[M+1] (Bezier) or u_knots, v_knots (B-spline).     */
typedef struct { vec3 p, dpdu, dpdv; } vec3x3;

/* _N_basis, cbezier.pyx:47-70 */
static double bs_basis(double t, int p, int idx, const double* knots) {
    if (p == 0) return ((knots[idx] <= t) && (t < knots[idx + 1])) ? 1.0 : 0.0;
    double out;
    double denom = knots[idx + p] - knots[idx];
    if (denom == 0.0) out = 0.0;
    else out = ((t - knots[idx]) / denom) * bs_basis(t, p - 1, idx, knots);
    denom = knots[idx + p + 1] - knots[idx + 1];
    if (denom != 0.0) out += ((knots[idx + p + 1] - t) / denom) * bs_basis(t, p - 1, idx + 1, knots);
    return out;
}

/* _N_basis_grad, cbezier.pyx:73-104 */
typedef struct { double N, dNdt; } basis_val;
static basis_val bs_basis_grad(double t, int p, int idx, const double* knots) {
    basis_val out, prev;
    if (p == 0) {
        out.N = ((knots[idx] <= t) && (t < knots[idx + 1])) ? 1.0 : 0.0;
        out.dNdt = 0.0;
        return out;
    }
    double denom = knots[idx + p] - knots[idx], nom;
    if (denom == 0.0) {
        out.N = 0.0;
        out.dNdt = 0.0;
    } else {
        prev = bs_basis_grad(t, p - 1, idx, knots);
        nom = ((t - knots[idx]) / denom);
        out.N = nom * prev.N;
        out.dNdt = (1. / denom) * prev.N + nom * prev.dNdt;
    }
    denom = knots[idx + p + 1] - knots[idx + 1];
    if (denom != 0.0) {
        prev = bs_basis_grad(t, p - 1, idx + 1, knots);
        nom = ((knots[idx + p + 1] - t) / denom);
        out.N += nom * prev.N;
        out.dNdt += (-1 / denom) * prev.N + nom * prev.dNdt;
    }
    return out;
}

typedef struct {
    int kind, N, M, udeg, vdeg;
    const double *uvs, *ctrl, *a, *b; /* a, b: binomials (Bezier) or knots (B-spline) */
} uvpatch_t;

static uvpatch_t uvpatch_of(const rpx_scene* S, const rpx_face* f) {
    uvpatch_t P;
    const double* H = S->pool + f->aux_off;
    const long n_points = (long)H[0];
    P.kind = (int)f->p[2];
    P.N = (int)f->p[3];
    P.M = (int)f->p[4];
    P.udeg = (int)f->p[5];
    P.vdeg = (int)f->p[6];
    P.uvs = S->pool + (long)f->p[7];
    P.ctrl = P.uvs + 2 * n_points;
    P.a = P.ctrl + 3 * (P.N + 1) * (P.M + 1);
    P.b = P.a + (P.kind == 0 ? P.N + 1 : (long)f->p[8]);
    return P;
}

/* BezierPatch._eval_pt (:232-254) / BSplinePatch._eval_pt (:335-358) */
static vec3 uvpatch_eval(const uvpatch_t* P, double u, double v) {
    vec3 out = v3(0, 0, 0);
    const int N = P->N, M = P->M;
    for (int i = 0; i < N + 1; i++)
        for (int j = 0; j < M + 1; j++) {
            const double* c = P->ctrl + 3 * (i * (M + 1) + j);
            if (P->kind == 0) {
                double coef = P->a[i] * pow(u, (double)i) * pow(1 - u, (double)(N - i)) * P->b[j] * pow(v, (double)j) *
                              pow(1 - v, (double)(M - j));
                out.x += coef * c[0];
                out.y += coef * c[1];
                out.z += coef * c[2];
            } else {
                double coef1 = bs_basis(u, P->udeg, i, P->a);
                double coef2 = bs_basis(v, P->vdeg, j, P->b);
                out.x += coef1 * coef2 * c[0];
                out.y += coef1 * coef2 * c[1];
                out.z += coef1 * coef2 * c[2];
            }
        }
    return out;
}

/* BezierPatch._eval_pt_and_grads (:259-286) / BSplinePatch._eval_pt_and_grads (:363-388) */
static vec3x3 uvpatch_eval_grads(const uvpatch_t* P, double u, double v) {
    vec3x3 out;
    out.p = out.dpdu = out.dpdv = v3(0, 0, 0);
    const int N = P->N, M = P->M;
    for (int i = 0; i < N + 1; i++)
        for (int j = 0; j < M + 1; j++) {
            const double* c = P->ctrl + 3 * (i * (M + 1) + j);
            if (P->kind == 0) {
                double u_term = P->a[i] * pow(u, (double)i) * pow(1 - u, (double)(N - i));
                double v_term = P->b[j] * pow(v, (double)j) * pow(1 - v, (double)(M - j));
                double coef = u_term * v_term;
                out.p.x += coef * c[0];
                out.p.y += coef * c[1];
                out.p.z += coef * c[2];
                coef = P->a[i] * (i - N * u) * pow(u, (double)(i - 1)) * pow(1 - u, (double)(N - 1 - i)) * v_term;
                out.dpdu.x += coef * c[0];
                out.dpdu.y += coef * c[1];
                out.dpdu.z += coef * c[2];
                coef = u_term * P->b[j] * (j - M * v) * pow(v, (double)(j - 1)) * pow(1 - v, (double)(M - 1 - j));
                out.dpdv.x += coef * c[0];
                out.dpdv.y += coef * c[1];
                out.dpdv.z += coef * c[2];
            } else {
                basis_val uc = bs_basis_grad(u, P->udeg, i, P->a);
                basis_val vc = bs_basis_grad(v, P->vdeg, j, P->b);
                out.p.x += uc.N * vc.N * c[0];
                out.p.y += uc.N * vc.N * c[1];
                out.p.z += uc.N * vc.N * c[2];
                out.dpdu.x += uc.dNdt * vc.N * c[0];
                out.dpdu.y += uc.dNdt * vc.N * c[1];
                out.dpdu.z += uc.dNdt * vc.N * c[2];
                out.dpdv.x += vc.dNdt * uc.N * c[0];
                out.dpdv.y += vc.dNdt * uc.N * c[1];
                out.dpdv.z += vc.dNdt * uc.N * c[2];
            }
        }
    return out;
}

/* UVPatchFace.intersect_c (:459-528) with interpolate_cell_c (:421-456) */
static double uvpatch_intersect(const rpx_scene* S, const rpx_face* f, vec3 p1, vec3 p2) {
    int cell_idx;
    const double* H = S->pool + f->aux_off;
    const double* points = H + (long)H[3];
    const double* cells = H + (long)H[4];
    double alpha = mesh_nearest(S, f, p1, p2, &cell_idx);
    if (alpha < f->tolerance) return NO_HIT;
    const uvpatch_t P = uvpatch_of(S, f);
    const double tol = f->p[0] * f->p[0];
    vec3 d = norm(subvv(p2, p1));
    vec3 pt = addvv(multvs(p2, alpha), multvs(p1, 1.0 - alpha));
    /* barycentric interpolation of the vertex uv's */
    const double* c = cells + 3 * (long)cell_idx;
    const long i0 = (long)c[0], i1 = (long)c[1], i2 = (long)c[2];
    vec3 q1 = ld3(points + 3 * i0), q2 = ld3(points + 3 * i1), q3 = ld3(points + 3 * i2);
    vec3 edge1 = subvv(q2, q1), edge2 = subvv(q3, q1);
    vec3 en1 = norm(edge1);
    vec3 en2 = norm(cross(en1, cross(edge1, edge2)));
    double x2 = mag(edge1), x3 = dotprod(edge2, en1), y3 = dotprod(edge2, en2);
    pt = subvv(pt, q1);
    double px = dotprod(pt, en1), py = dotprod(pt, en2);
    double alpha2 = (x2 * y3);
    double alpha1 = py * x3 - px * y3;
    double alpha0 = (alpha1 - py * x2 + x2 * y3) / alpha2;
    alpha1 = -alpha1 / alpha2;
    alpha2 = py / y3;
    double u = alpha0 * P.uvs[2 * i0] + alpha1 * P.uvs[2 * i1] + alpha2 * P.uvs[2 * i2];
    double v = alpha0 * P.uvs[2 * i0 + 1] + alpha1 * P.uvs[2 * i1 + 1] + alpha2 * P.uvs[2 * i2 + 1];
    int i;
    for (i = 0; i < 100; i++) {
        vec3x3 g = uvpatch_eval_grads(&P, u, v);
        vec3 normal = norm(cross(g.dpdu, g.dpdv));
        double dist = dotprod(subvv(g.p, p1), normal) / dotprod(d, normal);
        vec3 dp = subvv(addvv(p1, multvs(d, dist)), g.p);
        double du = dotprod(g.dpdu, dp) / mag_sq(g.dpdu);
        double dv = dotprod(g.dpdv, dp) / mag_sq(g.dpdv);
        u += du;
        v += dv;
        if ((du * du < tol) && (dv * dv < tol)) break;
    }
    if (i == 100) return NO_HIT;
    pt = uvpatch_eval(&P, u, v);
    s_u = u;
    s_v = v;
    return mag(subvv(pt, p1));
}

/* Face.intersect_c for every concrete class: distance along p1->p2, or <= 0 / -1 */
static double face_intersect(const rpx_scene* S, const rpx_face* f, vec3 p1, vec3 p2,
                             int is_base_ray) {
    s_piece = 0;
    if (f->type == RPX_FACE_MESH) return mesh_intersect(S, f, p1, p2, &s_piece);
    if (f->type == RPX_FACE_UVPATCH) return uvpatch_intersect(S, f, p1, p2);
    const double* P = f->p;
    const double tol = f->tolerance;
    switch (f->type) {
        case RPX_FACE_CIRCULAR: { /* cfaces.pyx:151-178 */
            double max_length = sep(p1, p2);
            double h = (P[2] - p1.z) / (p2.z - p1.z);
            double d = P[0];
            if ((h < tol) || (h > 1.0)) return NO_HIT;
            double X = p1.x + h * (p2.x - p1.x) - P[1];
            double Y = p1.y + h * (p2.y - p1.y);
            if (is_base_ray && (X * X + Y * Y) > (d * d / 4)) return NO_HIT;
            return h * max_length;
        }
        case RPX_FACE_SHAPED_PLANAR: { /* :201-226 */
            double max_length = sep(p1, p2);
            double h = (P[0] - p1.z) / (p2.z - p1.z);
            if ((h < tol) || (h > 1.0)) return NO_HIT;
            double X = p1.x + h * (p2.x - p1.x);
            double Y = p1.y + h * (p2.y - p1.y);
            if (is_base_ray && !shape_inside(S, f, X, Y)) return NO_HIT;
            return h * max_length;
        }
        case RPX_FACE_IMPLICIT_PLANAR: { /* :280-307 */
            vec3 normal = ld3(P + 3), origin = ld3(P);
            vec3 dp = subvv(p2, p1);
            vec3 po = subvv(origin, p1);
            double h = dotprod(po, normal) / dotprod(dp, normal);
            double max_length = mag(dp);
            if ((h < tol) || (h > 1.0)) return NO_HIT;
            po = addvv(p1, multvs(dp, h));
            if (is_base_ray && implicit_eval(S, f->aux_off, f->aux_n, po) > 0.0) return NO_HIT;
            return h * max_length;
        }
        case RPX_FACE_ELLIPTICAL_PLANE: { /* :322-341 */
            double max_length = sep(p1, p2);
            double gx = P[0], gy = P[1], d = P[2];
            double h = (gx * p1.x + gy * p1.y - p1.z) /
                       ((p2.z - p1.z) - gx * (p2.x - p1.x) - gy * (p2.y - p1.y));
            if ((h < tol) || (h > 1.0)) return NO_HIT;
            double X = p1.x + h * (p2.x - p1.x);
            double Y = p1.y + h * (p2.y - p1.y);
            if (is_base_ray && (X * X + Y * Y) > (d * d / 4)) return NO_HIT;
            return h * max_length;
        }
        case RPX_FACE_RECTANGULAR: { /* :365-397 */
            double max_length = sep(p1, p2);
            double h = (P[3] - p1.z) / (p2.z - p1.z);
            double lngth = P[0], wdth = P[1];
            if ((h < tol) || (h > 1.0)) return NO_HIT;
            if (is_base_ray) {
                double X = p1.x + h * (p2.x - p1.x) - P[2];
                double Y = p1.y + h * (p2.y - p1.y);
                if (X * X > lngth * lngth / 4) return NO_HIT;
                if (Y * Y > wdth * wdth / 4) return NO_HIT;
            }
            return h * max_length;
        }
        case RPX_FACE_SPHERICAL:          /* :421-486 */
        case RPX_FACE_SHAPED_SPHERICAL: { /* :513-576 */
            double curvature, z_height, diameter = 0;
            if (f->type == RPX_FACE_SPHERICAL) { diameter = P[0]; curvature = P[1]; z_height = P[2]; }
            else { curvature = P[0]; z_height = P[1]; }
            vec3 r = p1;
            vec3 s = subvv(p2, r);
            double cz = z_height - curvature;
            vec3 d = r;
            d.z -= cz;
            double A = mag_sq(s);
            double B = 2 * dotprod(s, d);
            double C = mag_sq(d) - pow(curvature, 2.0);
            double D = B * B - 4 * A * C;
            if (D < 0) return NO_HIT;
            D = sqrt(D);
            double a1 = (-B + D) / (2 * A);
            vec3 pt1 = addvv(r, multvs(s, a1));
            double a2 = (-B - D) / (2 * A);
            vec3 pt2 = addvv(r, multvs(s, a2));
            if (curvature >= 0) {
                if (pt1.z < cz) a1 = ORACLE_INF;
                if (pt2.z < cz) a2 = ORACLE_INF;
            } else {
                if (pt1.z > cz) a1 = ORACLE_INF;
                if (pt2.z > cz) a2 = ORACLE_INF;
            }
            if (f->type == RPX_FACE_SPHERICAL) {
                D = diameter * diameter / 4.;
                if (is_base_ray) {
                    if ((pt1.x * pt1.x + pt1.y * pt1.y) > D) a1 = ORACLE_INF;
                    if ((pt2.x * pt2.x + pt2.y * pt2.y) > D) a2 = ORACLE_INF;
                }
            } else if (is_base_ray) {
                if (!shape_inside(S, f, pt1.x, pt1.y)) a1 = ORACLE_INF;
                if (!shape_inside(S, f, pt2.x, pt2.y)) a2 = ORACLE_INF;
            }
            if (a2 < a1) a1 = a2;
            if (a1 > 1.0 || a1 < tol) return NO_HIT;
            return a1 * sep(r, p2);
        }
        case RPX_FACE_EXTRUDED_PLANAR: { /* :665-699 */
            vec3 r = p1;
            double ux = P[0], uy = P[1];
            double vx = P[2] - ux, vy = P[3] - uy;
            vec3 s = subvv(p2, r);
            double a;
            if (is_base_ray) {
                a = (s.y * (ux - r.x) - s.x * (uy - r.y)) / (s.x * vy - s.y * vx);
                if (a < 0) return NO_HIT;
                if (a > 1) return NO_HIT;
            }
            a = (vx * (r.y - uy) - vy * (r.x - ux)) / (s.x * vy - s.y * vx);
            if (is_base_ray) {
                double dz = a * (p2.z - r.z);
                if (P[4] < (r.z + dz) && (r.z + dz) < P[5]) return a * mag(s);
            } else {
                return a * mag(s);
            }
            return NO_HIT;
        }
        case RPX_FACE_POLYGON: { /* :1093-1108 */
            double max_length = sep(p1, p2);
            double h = (P[0] - p1.z) / (p2.z - p1.z);
            if ((h < tol) || (h > 1.0)) return NO_HIT;
            double X = p1.x + h * (p2.x - p1.x);
            double Y = p1.y + h * (p2.y - p1.y);
            if (is_base_ray && point_in_polygon(X, Y, S->pool + f->aux_off, f->aux_n) == 1)
                return h * max_length;
            return NO_HIT;
        }
        case RPX_FACE_ORIENTED_POLYGON: { /* :1189-1219 */
            vec3 n = ld3(P + 3), o = ld3(P);
            vec3 line = subvv(p2, p1);
            double max_length = mag(line);
            line = norm(line);
            double h = dotprod(line, n);
            if (h == 0.0) return NO_HIT;
            h = dotprod(subvv(o, p1), n) / h;
            if ((h < tol) || (h > max_length)) return NO_HIT;
            if (is_base_ray) {
                line = subvv(addvv(p1, multvs(line, h)), o);
                double X = dotprod(line, ld3(P + 6));
                double Y = dotprod(line, ld3(P + 9));
                if (point_in_polygon(X, Y, S->pool + f->aux_off, f->aux_n) == 1) return h;
                return NO_HIT;
            }
            return h;
        }
        case RPX_FACE_OFFAXIS_PARABOLIC: { /* :1228-1298 */
            double efl = P[0], diameter = P[1];
            double A = 1 / (2 * efl);
            vec3 s = subvv(p2, p1);
            vec3 r = p1;
            r.z += efl / 2.;
            double a = A * (pow(s.x, 2.0) + pow(s.y, 2.0));
            double b = 2 * A * (r.x * s.x + r.y * s.y) - s.z;
            double c = A * (pow(r.x, 2.0) + pow(r.y, 2.0)) - r.z;
            double d = pow(b, 2.0) - 4 * a * c;
            if (d < 0) return NO_HIT;
            if (a < 1e-10) {
                double a1 = -c / b;
                vec3 pt1 = addvv(r, multvs(s, a1));
                pt1.x -= efl;
                if ((pt1.x * pt1.x + pt1.y * pt1.y) > (diameter / 2)) return NO_HIT;
                if (a1 > 1.0 || a1 < tol) return NO_HIT;
                return a1 * sep(p1, p2);
            } else {
                d = sqrt(d);
                double a1 = (-b + d) / (2 * a);
                vec3 pt1 = addvv(r, multvs(s, a1));
                double a2 = (-b - d) / (2 * a);
                vec3 pt2 = addvv(r, multvs(s, a2));
                pt1.x -= efl;
                pt2.x -= efl;
                if (is_base_ray) {
                    d = diameter;
                    d *= d / 4.;
                    if ((pt1.x * pt1.x + pt1.y * pt1.y) > d) a1 = ORACLE_INF;
                    if ((pt2.x * pt2.x + pt2.y * pt2.y) > d) a2 = ORACLE_INF;
                }
                if (a2 < a1) a1 = a2;
                if (a1 > 1.0 || a1 < tol) return NO_HIT;
                return a1 * sep(p1, p2);
            }
        }
        case RPX_FACE_ELLIPSOIDAL: { /* :1344-1393 */
            const rpx_transform* T = (const rpx_transform*)(S->pool + f->aux_off);
            vec3 Sv = subvv(p2, p1);
            vec3 r = transform_pt(T, p1);
            vec3 s = transform_pt(T, p2);
            s = subvv(s, r);
            double B = pow(P[1], 2.0), A = pow(P[0], 2.0);
            double a = A * (s.z * s.z + s.y * s.y) + B * s.x * s.x;
            double b = 2 * (A * (r.z * s.z + r.y * s.y) + B * r.x * s.x);
            double c = A * (r.z * r.z + r.y * r.y) + B * r.x * r.x - A * B;
            double d = b * b - 4 * a * c;
            d = sqrt(d);
            double root1 = (-b + d) / (2 * a);
            double root2 = (-b - d) / (2 * a);
            vec3 q2 = addvv(p1, multvs(Sv, root2));
            vec3 q1 = addvv(p1, multvs(Sv, root1));
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
            if (root1 > 1) return NO_HIT;
            return root1 * mag(Sv);
        }
        case RPX_FACE_SADDLE: { /* :1439-1495 */
            double A = sqrt(6.0), root, denom, a1, a2;
            A *= P[1];
            vec3 p = p1;
            p.z -= P[0];
            vec3 d = subvv(p2, p1);
            if (d.x == 0.0) {
                a1 = (-A * (p.x * p.y) + p.z) / (A * d.y * p.x - d.z);
                a2 = ORACLE_INF;
            } else if (d.y == 0.0) {
                a1 = (-A * (p.x * p.y) + p.z) / (A * d.x * p.y - d.z);
                a2 = ORACLE_INF;
            } else {
                root = pow(A, 2.0) * pow(d.x, 2.0) * pow(p.y, 2.0) -
                       2 * pow(A, 2.0) * d.x * d.y * p.x * p.y +
                       pow(A, 2.0) * pow(d.y, 2.0) * pow(p.x, 2.0) + 4 * A * d.x * d.y * p.z -
                       2 * A * d.x * d.z * p.y - 2 * A * d.y * d.z * p.x + pow(d.z, 2.0);
                if (root < 0) return NO_HIT;
                root = sqrt(root);
                denom = 2 * A * (d.x * d.y);
                a1 = a2 = -A * d.x * p.y - A * d.y * p.x + d.z;
                a1 += root;
                a2 -= root;
                a1 /= denom;
                a2 /= denom;
            }
            vec3 pt1 = addvv(p1, multvs(d, a1));
            vec3 pt2 = addvv(p1, multvs(d, a2));
            if (a1 < 0.0) a1 = ORACLE_INF;
            if (a2 < 0.0) a2 = ORACLE_INF;
            if (is_base_ray) {
                if (!shape_inside(S, f, pt1.x, pt1.y)) a1 = ORACLE_INF;
                if (!shape_inside(S, f, pt2.x, pt2.y)) a2 = ORACLE_INF;
            }
            if (a2 < a1) a1 = a2;
            if (a1 > 1.0 || a1 < tol) return NO_HIT;
            return a1 * sep(p1, p2);
        }
        case RPX_FACE_CYLINDRICAL: { /* :1526-1584 */
            double R = P[1];
            double R2 = R * R;
            vec3 o = p1;
            o.z -= P[0];
            vec3 d = subvv(p2, p1);
            double ox2 = o.x * o.x, oz2 = o.z * o.z, dx2 = d.x * d.x, dz2 = d.z * d.z;
            double root = R2 * dz2 - 2 * R * dx2 * o.z + 2 * R * d.x * d.z * o.x - dx2 * oz2 +
                          2 * d.x * d.z * o.x * o.z - dz2 * ox2;
            if (root < 0) return NO_HIT;
            root = sqrt(root);
            double denom = dx2 + dz2;
            double a1, a2;
            a1 = a2 = -R * d.z - d.x * o.x - d.z * o.z;
            a1 += root;
            a2 -= root;
            a1 /= denom;
            a2 /= denom;
            vec3 pt1 = addvv(p1, multvs(d, a1));
            vec3 pt2 = addvv(p1, multvs(d, a2));
            double cz = P[0] - P[1];
            if (R >= 0) {
                if (pt1.z < cz) a1 = ORACLE_INF;
                if (pt2.z < cz) a2 = ORACLE_INF;
            } else {
                if (pt1.z > cz) a1 = ORACLE_INF;
                if (pt2.z > cz) a2 = ORACLE_INF;
            }
            if (is_base_ray) {
                if (!shape_inside(S, f, pt1.x, pt1.y)) a1 = ORACLE_INF;
                if (!shape_inside(S, f, pt2.x, pt2.y)) a2 = ORACLE_INF;
            }
            if (a2 < a1) a1 = a2;
            if (a1 > 1.0 || a1 < tol) return NO_HIT;
            return a1 * sep(p1, p2);
        }
        case RPX_FACE_AXICON: { /* :1621-1675 */
            double beta = P[1];
            vec3 d = subvv(p2, p1);
            vec3 o = p1;
            o.z -= P[0];
            double beta2 = beta * beta;
            double ox2 = o.x * o.x, oy2 = o.y * o.y, oz2 = o.z * o.z;
            double dx2 = d.x * d.x, dy2 = d.y * d.y, dz2 = d.z * d.z;
            double root = -beta2 * dx2 * oy2 + 2 * beta2 * d.x * d.y * o.x * o.y - beta2 * dy2 * ox2 +
                          dx2 * oz2 - 2 * d.x * d.z * o.x * o.z + dy2 * oz2 -
                          2 * d.y * d.z * o.y * o.z + dz2 * ox2 + dz2 * oy2;
            double denom = (beta2 * dx2 + beta2 * dy2 - dz2);
            if (root < 0) return NO_HIT;
            root = beta * sqrt(root);
            double a1 = -beta2 * d.x * o.x - beta2 * d.y * o.y + d.z * o.z;
            double a2 = a1 + root;
            a1 -= root;
            a1 /= denom;
            a2 /= denom;
            vec3 pt1 = addvv(p1, multvs(d, a1));
            vec3 pt2 = addvv(p1, multvs(d, a2));
            if (pt1.z > P[0]) a1 = ORACLE_INF;
            if (pt2.z > P[0]) a2 = ORACLE_INF;
            if (is_base_ray) {
                if (!shape_inside(S, f, pt1.x, pt1.y)) a1 = ORACLE_INF;
                if (!shape_inside(S, f, pt2.x, pt2.y)) a2 = ORACLE_INF;
            }
            if (a2 < a1) a1 = a2;
            if (a1 > 1.0 || a1 < tol) return NO_HIT;
            return a1 * sep(p1, p2);
        }
        case RPX_FACE_CONIC: { /* :1767-1798 */
            vec3 d = subvv(p2, p1);
            vec3 a = p1;
            a.z -= P[1];
            double a1 = intersect_conic(a, d, P[0], P[2]);
            vec3 pt1 = addvv(a, multvs(d, a1));
            if (is_base_ray && !shape_inside(S, f, pt1.x, pt1.y)) return NO_HIT;
            if (a1 > 1.0 || a1 < tol) return NO_HIT;
            return a1 * sep(p1, p2);
        }
        case RPX_FACE_ASPHERIC: { /* :1909-1976 */
            double atol = pow(P[11], 2.0);
            vec3 d = subvv(p2, p1);
            vec3 a = p1;
            a.z -= P[1];
            double a1 = intersect_conic(a, d, P[0], P[2]);
            aspheric_t A;
            A.R = -P[0];
            A.beta = 1 + P[2];
            A.A4 = P[4]; A.A6 = P[5]; A.A8 = P[6]; A.A10 = P[7];
            A.A12 = P[8]; A.A14 = P[9]; A.A16 = P[10];
            A.a = a;
            A.d = d;
            double f_, f_last, dz;
            f_ = f_last = aspheric_impf(&A, a1);
            dz = -f_ / aspheric_grad(&A, a1);
            int i, converged = 0;
            for (i = 0; i < 100; i++) {
                a1 += dz;
                if (dz * dz < atol) { converged = 1; break; }
                f_ = aspheric_impf(&A, a1);
                if (fabs(f_) > fabs(f_last)) return NO_HIT;
                f_last = f_;
                dz = -f_ / aspheric_grad(&A, a1);
            }
            if (!converged) return NO_HIT;
            vec3 pt1 = addvv(a, multvs(d, a1));
            if (is_base_ray && !shape_inside(S, f, pt1.x, pt1.y)) return NO_HIT;
            if (a1 > 1.0 || a1 < tol) return NO_HIT;
            return a1 * sep(p1, p2);
        }
        case RPX_FACE_EXT_POLY: { /* :2186-2237 */
            const double* E = S->pool + f->aux_off;
            double atol = pow(P[4], 2.0);
            vec3 d = subvv(p2, p1);
            vec3 a = p1;
            a.z -= P[3];
            double a1 = intersect_conic(a, d, -P[0], P[1] - 1.0);
            double f_, f_last, dz;
            f_ = f_last = extpoly_impf(f, E, p1, d, a1);
            dz = -f_ / extpoly_grad(f, E, p1, d, a1);
            int i, converged = 0;
            for (i = 0; i < 100; i++) {
                a1 += dz;
                if (dz * dz < atol) { converged = 1; break; }
                f_ = extpoly_impf(f, E, p1, d, a1);
                if (fabs(f_) > fabs(f_last)) return NO_HIT;
                f_last = f_;
                dz = -f_ / extpoly_grad(f, E, p1, d, a1);
            }
            if (!converged) return NO_HIT;
            vec3 pt1 = addvv(a, multvs(d, a1));
            if (is_base_ray && !shape_inside(S, f, pt1.x, pt1.y)) return NO_HIT;
            if (a1 > 1.0 || a1 < tol) return NO_HIT;
            return a1 * sep(p1, p2);
        }
        case RPX_FACE_DISTORTION: { /* :2339-2416 */
            const rpx_face* base = &S->faces[f->base_face];
            const rpx_distortion* dist = &S->distortions[f->aux_off];
            double h = sep(p2, p1);
            double tolerance = P[0];
            double a2 = face_intersect(S, base, p1, p2, 0);
            if (a2 > h || a2 < tol) return NO_HIT;
            vec3 d = subvv(p2, p1);
            vec3 pt1 = addvv(p1, multvs(d, a2 / h));
            vec3 dxdyz = distortion_zgrad(S, dist, pt1.x, pt1.y);
            vec3 n = face_normal(S, base, pt1);
            pt1.z += dxdyz.z;
            n.x /= n.z;
            n.y /= n.z;
            n.x -= dxdyz.x;
            n.y -= dxdyz.y;
            vec3 o = subvv(p1, pt1);
            double a1 = -h * dotprod(o, n) / dotprod(d, n);
            if (a1 < fabs(dxdyz.z)) return NO_HIT;
            for (int i = 0; i < 20; i++) {
                pt1 = addvv(p1, multvs(d, a1 / h));
                if (fabs(a1 - a2) < tolerance) break;
                double z_shift = distortion_z(S, dist, pt1.x, pt1.y);
                vec3 q1 = p1, q2 = p2;
                q1.z -= z_shift;
                q2.z -= z_shift;
                a2 = face_intersect(S, base, q1, q2, 0);
                pt1 = addvv(q1, multvs(d, a2 / h));
                n = face_normal(S, base, pt1);
                dxdyz = distortion_zgrad(S, dist, pt1.x, pt1.y);
                pt1.z += dxdyz.z;
                n.x /= n.z;
                n.y /= n.z;
                n.x -= dxdyz.x;
                n.y -= dxdyz.y;
                o = subvv(p1, pt1);
                a2 = a1;
                a1 = -h * dotprod(o, n) / dotprod(d, n);
            }
            if (!shape_inside(S, f, pt1.x, pt1.y)) return NO_HIT; /* even for parabasal rays */
            return a1;
        }
        case RPX_FACE_EXTRUDED_BEZIER: { /* cfaces.pyx:867-971 (is_base_ray is ignored) */
            const double* curves = S->pool + f->aux_off;
            double z1 = P[0], z2 = P[1];
            flat2 mincorner = {P[2], P[3]}, maxcorner = {P[4], P[5]};
            flat2 origin = {0, 0}, r, q2, s, tempvector;
            double result = ORACLE_INF;
            if ((p1.z < z1 && p2.z < z1) || (p1.z > z2 && p2.z > z2)) return NO_HIT;
            r.x = p1.x; r.y = p1.y;
            q2.x = p2.x; q2.y = p2.y;
            tempvector.x = mincorner.x;
            tempvector.y = maxcorner.y;
            if (!bz_seg_overlap(r, q2, mincorner, tempvector)) {
                if (!bz_seg_overlap(r, q2, tempvector, maxcorner)) {
                    tempvector.x = maxcorner.x;
                    tempvector.y = mincorner.y;
                    if (!bz_seg_overlap(r, q2, maxcorner, tempvector)) {
                        if (!bz_seg_overlap(r, q2, tempvector, mincorner)) return NO_HIT;
                    }
                }
            }
            vec3 tempv = subvv(p2, p1);
            double dZ = tempv.z;
            s.x = tempv.x;
            s.y = tempv.y;
            double theta = atan2(s.y, s.x);
            s = rotate2D(-theta, s);
            for (int ci = 0; ci < f->aux_n; ci++) {
                flat2 cp[4];
                for (int q = 0; q < 4; q++) {
                    cp[q].x = curves[(ci * 4 + q) * 2] - p1.x;
                    cp[q].y = curves[(ci * 4 + q) * 2 + 1] - p1.y;
                    cp[q] = rotate2D(-theta, cp[q]);
                }
                if (bz_seg_overlap(origin, s, cp[0], cp[1]) || bz_seg_overlap(origin, s, cp[1], cp[2]) ||
                    bz_seg_overlap(origin, s, cp[2], cp[3]) || bz_seg_overlap(origin, s, cp[3], cp[0])) {
                    double A = cp[3].y - 3 * cp[2].y + 3 * cp[1].y - cp[0].y;
                    double B = 3 * cp[2].y - 6 * cp[1].y + 3 * cp[0].y;
                    double C = 3 * cp[1].y - 3 * cp[0].y;
                    double D = cp[0].y;
                    poly_roots ts = roots_of_cubic(A, B, C, D);
                    while (ts.n > 0) {
                        ts.n -= 1;
                        double t = ts.roots[ts.n];
                        if (0. < t && t < 1.) {
                            double b = eval_bezier(t, cp[0].x, cp[1].x, cp[2].x, cp[3].x);
                            if (0 < b && b < s.x) {
                                double c = dZ * b / s.x;
                                double a = c + p1.z;
                                if (z1 < a && a < z2) {
                                    b = sqrt(pow(c, 2.0) + pow(b, 2.0));
                                    if (b < result && b > tol) result = b;
                                }
                            }
                        }
                    }
                }
            }
            if (result == ORACLE_INF) return NO_HIT;
            return result;
        }
        default: return NO_HIT;
    }
}

/* Face.compute_normal_c (local coordinates) */
static vec3 face_normal(const rpx_scene* S, const rpx_face* f, vec3 p) {
    const double* P = f->p;
    if (f->type == RPX_FACE_MESH) return mesh_normal(S, f, s_hit_piece);
    if (f->type == RPX_FACE_UVPATCH) { /* compute_normal_and_tangent_c, cbezier.pyx:533-550 */
        const uvpatch_t UP = uvpatch_of(S, f);
        vec3x3 g = uvpatch_eval_grads(&UP, s_hit_u, s_hit_v);
        vec3 n = norm(cross(g.dpdu, g.dpdv));
        return (P[1] != 0.0) ? invert(n) : n;
    }
    switch (f->type) {
        case RPX_FACE_CIRCULAR: return v3(0, 0, P[3] != 0.0 ? 1 : -1); /* :180-191 */
        case RPX_FACE_SHAPED_PLANAR: return v3(0, 0, 1);               /* :228-236 */
        case RPX_FACE_IMPLICIT_PLANAR: return ld3(P + 3);              /* :309-310 */
        case RPX_FACE_ELLIPTICAL_PLANE: return norm(v3(P[0], P[1], -1)); /* :343-351 */
        case RPX_FACE_RECTANGULAR: return v3(0, 0, -1);                /* :399-407 */
        case RPX_FACE_SPHERICAL:                                       /* :488-498 */
        case RPX_FACE_SHAPED_SPHERICAL: {                              /* :578-588 */
            double curvature = (f->type == RPX_FACE_SPHERICAL) ? P[1] : P[0];
            double z_height = (f->type == RPX_FACE_SPHERICAL) ? P[2] : P[1];
            p.z -= (z_height - curvature);
            if (curvature < 0) { p.z = -p.z; p.y = -p.y; p.x = -p.x; }
            return norm(p);
        }
        case RPX_FACE_EXTRUDED_PLANAR: return ld3(P + 6); /* :701-702 */
        case RPX_FACE_POLYGON: return v3(0, 0, -1);       /* :1110-1118 */
        case RPX_FACE_ORIENTED_POLYGON: return ld3(P + 3); /* :1183-1184 */
        case RPX_FACE_OFFAXIS_PARABOLIC: {                 /* :1300-1317 */
            double A = 1 / (2 * P[0]);
            double m2 = p.x * p.x + p.y * p.y;
            double B = 4 * m2 * A * A;
            double dz = -sqrt(B / (B + 1));
            m2 = sqrt(m2);
            return v3(-(dz * p.x) / m2, -(dz * p.y) / m2, -1 / sqrt(B + 1));
        }
        case RPX_FACE_ELLIPSOIDAL: { /* :1395-1405 */
            const rpx_transform* T = (const rpx_transform*)(S->pool + f->aux_off);
            p = transform_pt(T, p);
            vec3 n = v3(p.x / -(pow(P[0], 2.0)), p.y / -(pow(P[1], 2.0)), p.z / -(pow(P[1], 2.0)));
            n = rotate_v(T + 1, n);
            return norm(n);
        }
        case RPX_FACE_SADDLE: { /* :1497-1508 */
            double rt6 = sqrt(6) * P[1];
            return norm(v3(-rt6 * p.y, -rt6 * p.x, 1.0));
        }
        case RPX_FACE_CYLINDRICAL: { /* :1586-1596 */
            p.z -= (P[0] - P[1]);
            if (P[1] < 0) { p.z = -p.z; p.x = -p.x; }
            p.y = 0;
            return norm(p);
        }
        case RPX_FACE_AXICON: { /* :1677-1689 */
            double beta = P[1];
            double r = sqrt(p.x * p.x + p.y * p.y);
            return v3(beta * p.x / r, beta * p.y / r, 1.0);
        }
        case RPX_FACE_CONIC: { /* :1800-1823 */
            double R = -P[0], beta = 1 + P[2];
            int sign = (P[3] != 0.0) ? -1 : 1;
            p.z -= P[1];
            vec3 g;
            g.z = 2 * beta * (R - beta * p.z);
            g.x = -p.x * 2 * beta;
            g.y = -p.y * 2 * beta;
            if ((R * beta) < 0) sign *= -1;
            g.z *= sign;
            g.y *= sign;
            g.x *= sign;
            return norm(g);
        }
        case RPX_FACE_ASPHERIC: { /* :1978-2007 */
            double R = -P[0], beta = 1 + P[2];
            int sign = (P[3] != 0.0) ? -1 : 1;
            p.z -= P[1];
            double r2 = p.x * p.x + p.y * p.y;
            double root = sqrt(1 - (beta * (r2) / (R * R)));
            double df = 10 * P[7] * pow(r2, 4.0) + 8 * P[6] * pow(r2, 3.0) + 6 * P[5] * pow(r2, 2.0) +
                        4 * P[4] * r2;
            df += 16 * P[10] * pow(r2, 7.0) + 14 * P[9] * pow(r2, 6.0) + 12 * P[8] * pow(r2, 5.0);
            df += 2 / (R * (1 + root));
            df += beta * (r2) / (pow(R, 3.0) * root * pow(1 + root, 2.0));
            vec3 g;
            g.z = 1.0;
            g.x = -df * p.x;
            g.y = -df * p.y;
            g.z *= sign;
            g.y *= sign;
            g.x *= sign;
            return norm(g);
        }
        case RPX_FACE_EXT_POLY: { /* :2239-2291 */
            const double* E = S->pool + f->aux_off;
            int Nx = f->aux_n, Ny = f->aux_m;
            double R = P[0], beta = P[1];
            int inv = (P[5] != 0.0);
            int sign = inv ? -1 : 1;
            double inv_rad = 1. / P[2];
            double x = p.x * inv_rad, y = p.y * inv_rad;
            p.z -= P[3];
            vec3 g;
            g.z = 2 * beta * (R - beta * p.z);
            g.x = -p.x * 2 * beta;
            g.y = -p.y * 2 * beta;
            if ((R * beta) < 0) sign *= -1;
            g.z *= sign;
            g.y *= sign;
            g.x *= sign;
            g = norm(g);
            if (inv) {
                for (int i = 1; i < Nx; i++)
                    for (int j = 0; j < Ny; j++)
                        g.x += i * E[i * Ny + j] * inv_rad * pow(x, (double)(i - 1)) * pow(y, (double)j);
                for (int i = 0; i < Nx; i++)
                    for (int j = 1; j < Ny; j++)
                        g.y += j * E[i * Ny + j] * inv_rad * pow(x, (double)i) * pow(y, (double)(j - 1));
            } else {
                for (int i = 1; i < Nx; i++)
                    for (int j = 0; j < Ny; j++)
                        g.x -= i * E[i * Ny + j] * inv_rad * pow(x, (double)(i - 1)) * pow(y, (double)j);
                for (int i = 0; i < Nx; i++)
                    for (int j = 1; j < Ny; j++)
                        g.y -= j * E[i * Ny + j] * inv_rad * pow(x, (double)i) * pow(y, (double)(j - 1));
            }
            return norm(g);
        }
        case RPX_FACE_DISTORTION: { /* :2418-2431 */
            const rpx_face* base = &S->faces[f->base_face];
            const rpx_distortion* dist = &S->distortions[f->aux_off];
            vec3 dxdyz = distortion_zgrad(S, dist, p.x, p.y);
            vec3 p1 = p;
            p1.z -= dxdyz.z;
            vec3 n = face_normal(S, base, p1);
            n.x /= n.z;
            n.y /= n.z;
            n.z = 1.0;
            n.x -= dxdyz.x;
            n.y -= dxdyz.y;
            return norm(n);
        }
        case RPX_FACE_EXTRUDED_BEZIER: { /* cfaces.pyx:975-1046 */
            const double* curves = S->pool + f->aux_off;
            flat2 ray = {p.x, p.y};
            double theta = atan2(p.y, p.x);
            for (int ci = 0; ci < f->aux_n; ci++) {
                flat2 cp[4];
                for (int q = 0; q < 4; q++) {
                    cp[q].x = curves[(ci * 4 + q) * 2];
                    cp[q].y = curves[(ci * 4 + q) * 2 + 1];
                }
                if (bz_pnt_in_hull(ray, cp[0], cp[1], cp[2], cp[3])) {
                    for (int q = 0; q < 4; q++) cp[q] = rotate2D(-theta, cp[q]);
                    double A = cp[3].y - 3 * cp[2].y + 3 * cp[1].y - cp[0].y;
                    double B = 3 * cp[2].y - 6 * cp[1].y + 3 * cp[0].y;
                    double C = 3 * cp[1].y - 3 * cp[0].y;
                    double D = cp[0].y;
                    poly_roots ts = roots_of_cubic(A, B, C, D);
                    while (ts.n > 0) {
                        ts.n -= 1;
                        double t = ts.roots[ts.n];
                        if (0 <= t && t <= 1) {
                            double tmp = eval_bezier(t, cp[0].x, cp[1].x, cp[2].x, cp[3].x);
                            if (pow(tmp, 2.0) - (pow(ray.x, 2.0) + pow(ray.y, 2.0)) < .0001) {
                                ray.x = dif_bezier(t, cp[0].x, cp[1].x, cp[2].x, cp[3].x);
                                ray.y = dif_bezier(t, cp[0].y, cp[1].y, cp[2].y, cp[3].y);
                                ray = rotate2D(theta, ray);
                                p.z = 0;
                                if (ray.y == 0) {
                                    p.x = 0;
                                    p.y = (ray.x > 0 ? 1 : -1);
                                } else if (ray.y > 0) {
                                    p.x = -1;
                                    p.y = ray.x / ray.y;
                                } else if (ray.y < 0) {
                                    p.x = 1;
                                    p.y = -ray.x / ray.y;
                                }
                                return norm(p);
                            }
                        }
                    }
                }
            }
            return v3(0, 0, 0); /* "Bezier normal not found": the reference prints and returns 0 */
        }
        default: return p; /* Face.compute_normal_c base, ctracer.pyx:1783-1784 */
    }
}

/* Face.compute_tangent_c: default (1,0,0) ctracer.pyx:1786-1791 */
static vec3 face_tangent(const rpx_scene* S, const rpx_face* f) {
    if (f->type == RPX_FACE_UVPATCH) { /* tangent = norm(dpdu), cbezier.pyx:550 */
        const uvpatch_t UP = uvpatch_of(S, f);
        return norm(uvpatch_eval_grads(&UP, s_hit_u, s_hit_v).dpdu);
    }
    switch (f->type) {
        case RPX_FACE_EXTRUDED_PLANAR: return v3(0.0, 0.0, 1.0);    /* cfaces.pyx:704-709 */
        case RPX_FACE_ORIENTED_POLYGON: return ld3(f->p + 6);       /* :1186-1187 */
        default: return v3(1.0, 0.0, 0.0);
    }
}

/* FaceList.compute_orientation_c, ctracer.pyx:1939-1953 */
static orient_t compute_orientation(const rpx_scene* S, const rpx_face* f, vec3 point) {
    const rpx_face_set* fs = &S->face_sets[f->face_set];
    orient_t out;
    point = transform_pt(&fs->inv_trans, point);
    out.normal = face_normal(S, f, point);
    out.tangent = face_tangent(S, f);
    if (f->invert_normal) {
        out.normal = invert(out.normal);
        out.tangent = invert(out.tangent);
    }
    out.normal = rotate_v(&fs->trans, out.normal);
    out.tangent = rotate_v(&fs->trans, out.tangent);
    return out;
}

/* ------------------------------------------------ materials, cmaterials.pyx */
/* convert_to_sp, cmaterials.pyx:49-91 */
static rpx_ray convert_to_sp(rpx_ray ray, vec3 normal) {
    vec3 dir = ld3(ray.direction);
    vec3 E2_vector = norm(cross(dir, ld3(ray.E_vector)));
    vec3 E1_vector = norm(cross(E2_vector, dir));
    normal = norm(normal);
    vec3 S_vector = cross(dir, normal);
    if (fabs(S_vector.x) < SP_TOL && fabs(S_vector.y) < SP_TOL && fabs(S_vector.z) < SP_TOL)
        return ray;
    S_vector = norm(S_vector);
    vec3 v = cross(dir, S_vector);
    vec3 P_vector = norm(v);
    double A = dotprod(E1_vector, S_vector);
    double B = dotprod(E2_vector, S_vector);
    double S_re = ray.E1_amp[0] * A + ray.E2_amp[0] * B;
    double S_im = ray.E1_amp[1] * A + ray.E2_amp[1] * B;
    B = dotprod(E1_vector, P_vector);
    A = dotprod(E2_vector, P_vector);
    double P_re = ray.E1_amp[0] * B + ray.E2_amp[0] * A;
    double P_im = ray.E1_amp[1] * B + ray.E2_amp[1] * A;
    st3(ray.E_vector, S_vector);
    ray.E1_amp[0] = S_re; ray.E1_amp[1] = S_im;
    ray.E2_amp[0] = P_re; ray.E2_amp[1] = P_im;
    return ray;
}

typedef struct { rpx_ray* rays; int n; } childbuf_t;
static inline void add_ray(childbuf_t* out, const rpx_ray* r) { out->rays[out->n++] = *r; }

static inline cplx ntab_get(const rpx_scene* S, const rpx_material* M, int row, uint32_t wl) {
    const double* t = S->ntab + 2 * ((size_t)M->ntab_off + (size_t)row * S->n_wavelengths + wl);
    return CMPLX(t[0], t[1]);
}

/* Shared tail of FullDielectric*/ /* and coated materials: emit reflected then transmitted */
static void fresnel_emit(childbuf_t* out, rpx_ray* sp_ray, uint32_t idx, vec3 point, vec3 normal,
                         vec3 in_direction, vec3 cosThetaNormal, int flip, cplx n1, cplx n_t,
                         cplx R_s, cplx R_p, cplx T_s, cplx T_p, double P_in, double refl_thr,
                         double trans_thr) {
    /* reflected: cmaterials.pyx:826-840 / 967-981 / 1140-1154 / 1349-1363 */
    if ((creal(n1) * (pow(cabs(R_s), 2.0) + pow(cabs(R_p), 2.0)) / P_in) > refl_thr) {
        vec3 reflected = subvv(in_direction, multvs(cosThetaNormal, 2));
        st3(sp_ray->origin, point);
        st3(sp_ray->normal, normal);
        st3(sp_ray->direction, reflected);
        sp_ray->length = ORACLE_INF;
        sp_ray->E1_amp[0] = creal(R_s);
        sp_ray->E1_amp[1] = cimag(R_s);
        sp_ray->E2_amp[0] = -creal(R_p);
        sp_ray->E2_amp[1] = -cimag(R_p);
        sp_ray->parent_idx = idx;
        sp_ray->refractive_index[0] = creal(n1);
        sp_ray->refractive_index[1] = cimag(n1);
        sp_ray->ray_type_id |= RPX_REFL_RAY;
        add_ray(out, sp_ray);
    }
    /* transmitted direction: :843-847 (real-part approximation) */
    vec3 tangent = subvv(in_direction, cosThetaNormal);
    vec3 tg2 = multvs(tangent, creal(n1) / creal(n_t));
    double tan_mag_sq = mag_sq(tg2);
    double c2 = sqrt(1 - tan_mag_sq);
    vec3 transmitted = subvv(tg2, multvs(normal, c2 * flip));
    if ((creal(n_t) * (pow(cabs(T_s), 2.0) + pow(cabs(T_p), 2.0)) / P_in) > trans_thr) {
        st3(sp_ray->origin, point);
        st3(sp_ray->normal, normal);
        st3(sp_ray->direction, transmitted);
        sp_ray->length = ORACLE_INF;
        sp_ray->E1_amp[0] = creal(T_s);
        sp_ray->E1_amp[1] = cimag(T_s);
        sp_ray->E2_amp[0] = creal(T_p);
        sp_ray->E2_amp[1] = cimag(T_p);
        sp_ray->parent_idx = idx;
        sp_ray->refractive_index[0] = creal(n_t);
        sp_ray->refractive_index[1] = cimag(n_t);
        sp_ray->ray_type_id &= ~RPX_REFL_RAY;
        add_ray(out, sp_ray);
    }
}

/* InterfaceMaterial.eval_child_ray_c for every material class */
static void material_eval(const rpx_scene* S, const rpx_material* M, const rpx_ray* in_ray,
                          uint32_t idx, vec3 point, orient_t orient, childbuf_t* out) {
    const double* P = M->p;
    switch (M->type) {
        case RPX_MAT_OPAQUE: return; /* :245-251 */
        case RPX_MAT_TRANSPARENT: {  /* :260-278 */
            vec3 normal = norm(orient.normal);
            rpx_ray sp = convert_to_sp(*in_ray, normal);
            sp.accumulated_path += sp.length * sp.refractive_index[0];
            st3(sp.origin, point);
            st3(sp.normal, normal);
            sp.length = ORACLE_INF;
            sp.parent_idx = idx;
            sp.ray_type_id &= ~RPX_REFL_RAY;
            add_ray(out, &sp);
            return;
        }
        case RPX_MAT_PEC: { /* :285-319 */
            vec3 normal = norm(orient.normal);
            rpx_ray sp = convert_to_sp(*in_ray, normal);
            vec3 dir = ld3(in_ray->direction);
            double cosTheta = dotprod(normal, dir);
            vec3 cosThetaNormal = multvs(normal, cosTheta);
            vec3 reflected = subvv(dir, multvs(cosThetaNormal, 2));
            sp.accumulated_path += sp.length * sp.refractive_index[0];
            st3(sp.origin, point);
            st3(sp.normal, normal);
            st3(sp.direction, reflected);
            sp.length = ORACLE_INF;
            sp.E1_amp[0] = -sp.E1_amp[0];
            sp.E1_amp[1] = -sp.E1_amp[1];
            sp.parent_idx = idx;
            sp.ray_type_id |= RPX_REFL_RAY;
            add_ray(out, &sp);
            return;
        }
        case RPX_MAT_PARTIALLY_REFLECTIVE: /* :345-397 */
        case RPX_MAT_LINEAR_POLARISING: {  /* :404-455 */
            vec3 normal = norm(orient.normal);
            vec3 in_direction = norm(ld3(in_ray->direction));
            rpx_ray sp, sp2;
            sp = sp2 = convert_to_sp(*in_ray, normal);
            double cosTheta = dotprod(normal, in_direction);
            vec3 cosThetaNormal = multvs(normal, cosTheta);
            sp.accumulated_path += sp.length * sp.refractive_index[0];
            sp2.accumulated_path = sp.accumulated_path;
            vec3 reflected = subvv(in_direction, multvs(cosThetaNormal, 2));
            st3(sp.origin, point);
            st3(sp.normal, normal);
            st3(sp.direction, reflected);
            sp.length = ORACLE_INF;
            st3(sp2.origin, point);
            st3(sp2.normal, normal);
            st3(sp2.direction, in_direction);
            sp2.length = ORACLE_INF;
            if (M->type == RPX_MAT_PARTIALLY_REFLECTIVE) {
                double R = sqrt(P[0]);
                double T = sqrt(1 - P[0]);
                /* `E1_amp *= R` on a complex_t: complex * (R + 0i) */
                cplx e;
                e = CMPLX(sp.E1_amp[0], sp.E1_amp[1]) * CX(R); sp.E1_amp[0] = creal(e); sp.E1_amp[1] = cimag(e);
                e = CMPLX(sp.E2_amp[0], sp.E2_amp[1]) * CX(R); sp.E2_amp[0] = creal(e); sp.E2_amp[1] = cimag(e);
                e = CMPLX(sp2.E1_amp[0], sp2.E1_amp[1]) * CX(T); sp2.E1_amp[0] = creal(e); sp2.E1_amp[1] = cimag(e);
                e = CMPLX(sp2.E2_amp[0], sp2.E2_amp[1]) * CX(T); sp2.E2_amp[0] = creal(e); sp2.E2_amp[1] = cimag(e);
            } else {
                sp.E2_amp[0] = 0.0;
                sp.E2_amp[1] = 0.0;
                sp2.E1_amp[0] = 0.0;
                sp2.E1_amp[1] = 0.0;
            }
            sp.parent_idx = idx;
            sp.ray_type_id |= RPX_REFL_RAY;
            add_ray(out, &sp);
            sp2.parent_idx = idx;
            sp2.ray_type_id &= ~RPX_REFL_RAY;
            add_ray(out, &sp2);
            return;
        }
        case RPX_MAT_WAVEPLATE: { /* :520-551 */
            vec3 normal = norm(orient.normal);
            vec3 in_direction = norm(ld3(in_ray->direction));
            rpx_ray o = convert_to_sp(*in_ray, ld3(P + 2));
            o.accumulated_path += o.length * o.refractive_index[0];
            st3(o.origin, point);
            st3(o.normal, normal);
            st3(o.direction, in_direction);
            o.length = ORACLE_INF;
            o.parent_idx = idx;
            o.ray_type_id &= ~RPX_REFL_RAY;
            double e1r = o.E1_amp[0], e1i = o.E1_amp[1]; /* apply_retardance_c :498-504 */
            o.E1_amp[0] = e1r * P[0] - e1i * P[1];
            o.E1_amp[1] = e1i * P[0] + e1r * P[1];
            add_ray(out, &o);
            return;
        }
        case RPX_MAT_DIELECTRIC: { /* :587-680 */
            cplx n_inside = ntab_get(S, M, 0, in_ray->wavelength_idx);
            cplx n_outside = ntab_get(S, M, 1, in_ray->wavelength_idx);
            vec3 normal = norm(orient.normal);
            vec3 in_direction = norm(ld3(in_ray->direction));
            rpx_ray sp = convert_to_sp(*in_ray, normal);
            sp.accumulated_path += sp.length * sp.refractive_index[0];
            double cosTheta = dotprod(normal, in_direction);
            double cos1 = fabs(cosTheta);
            double n1, n2;
            int flip;
            if (cosTheta < 0.0) {
                n1 = creal(n_outside);
                n2 = creal(n_inside);
                sp.refractive_index[0] = creal(n_inside);
                sp.refractive_index[1] = cimag(n_inside);
                flip = 1;
            } else {
                n1 = creal(n_inside);
                n2 = creal(n_outside);
                sp.refractive_index[0] = creal(n_outside);
                sp.refractive_index[1] = cimag(n_outside);
                flip = -1;
            }
            double N2 = pow(n2 / n1, 2.0);
            double N2_sin2 = (cosTheta * cosTheta) + (N2 - 1);
            vec3 cosThetaNormal = multvs(normal, cosTheta);
            if (N2_sin2 < 0.0) {
                vec3 reflected = subvv(in_direction, multvs(cosThetaNormal, 2));
                st3(sp.origin, point);
                st3(sp.normal, normal);
                st3(sp.direction, reflected);
                sp.length = ORACLE_INF;
                sp.E1_amp[0] *= -1; sp.E1_amp[1] *= -1;
                sp.E2_amp[0] *= -1; sp.E2_amp[1] *= -1;
                sp.parent_idx = idx;
                sp.ray_type_id |= RPX_REFL_RAY;
            } else {
                vec3 tangent = subvv(in_direction, cosThetaNormal);
                vec3 tg2 = multvs(tangent, n1 / n2);
                double tan_mag_sq = mag_sq(tg2);
                double c2 = sqrt(1 - tan_mag_sq);
                vec3 transmitted = subvv(tg2, multvs(normal, c2 * flip));
                double cos2 = fabs(dotprod(transmitted, normal));
                double Two_n1_cos1 = (2 * n1) * cos1;
                double aspect = sqrt(cos2 / cos1) * Two_n1_cos1;
                double T_p = aspect / (n2 * cos1 + n1 * cos2);
                double T_s = aspect / (n2 * cos2 + n1 * cos1);
                st3(sp.origin, point);
                st3(sp.normal, normal);
                st3(sp.direction, transmitted);
                sp.length = ORACLE_INF;
                sp.E1_amp[0] *= T_s; sp.E1_amp[1] *= T_s;
                sp.E2_amp[0] *= T_p; sp.E2_amp[1] *= T_p;
                sp.parent_idx = idx;
                sp.ray_type_id &= ~RPX_REFL_RAY;
            }
            add_ray(out, &sp);
            return;
        }
        case RPX_MAT_FULL_DIELECTRIC: { /* :755-872 and :896-1013 */
            vec3 normal = norm(orient.normal);
            vec3 in_direction = norm(ld3(in_ray->direction));
            rpx_ray sp = convert_to_sp(*in_ray, normal);
            sp.accumulated_path += sp.length * sp.refractive_index[0];
            cplx E1_amp = cy_parts(sp.E1_amp[0], sp.E1_amp[1]);
            cplx E2_amp = cy_parts(sp.E2_amp[0], sp.E2_amp[1]);
            double cosTheta = dotprod(normal, in_direction);
            double cos1 = fabs(cosTheta);
            double sin1 = sqrt(fabs(1 - cos1 * cos1));
            cplx n1, n2;
            int flip;
            if (cosTheta < 0.0) {
                n1 = ntab_get(S, M, 1, in_ray->wavelength_idx);
                n2 = ntab_get(S, M, 0, in_ray->wavelength_idx);
                flip = 1;
            } else {
                n1 = ntab_get(S, M, 0, in_ray->wavelength_idx);
                n2 = ntab_get(S, M, 1, in_ray->wavelength_idx);
                flip = -1;
            }
            cplx sin2 = (n1 * CX(sin1)) / n2;
            cplx cos2 = csqrt(CX(1) - sin2 * sin2);
            vec3 cosThetaNormal = multvs(normal, cosTheta);
            double P_in = creal(n1) * (pow(creal(E1_amp), 2.0) + pow(cimag(E1_amp), 2.0) +
                                       pow(creal(E2_amp), 2.0) + pow(cimag(E2_amp), 2.0));
            if (P_in == 0.0) return;
            cplx R_p = (-(n2 * CX(cos1) - n1 * cos2)) / (n2 * CX(cos1) + n1 * cos2);
            cplx R_s = (-(n2 * cos2 - n1 * CX(cos1))) / (n2 * cos2 + n1 * CX(cos1));
            R_s = R_s * E1_amp;
            R_p = R_p * E2_amp;
            double aspect = sqrt(creal(cos2) / cos1);
            cplx T_p = (CX(aspect) * (CX(2.0 * cos1) * n1)) / (n2 * CX(cos1) + n1 * cos2);
            cplx T_s = (CX(aspect) * (CX(2.0 * cos1) * n1)) / (n2 * cos2 + n1 * CX(cos1));
            T_s = T_s * E1_amp;
            T_p = T_p * E2_amp;
            fresnel_emit(out, &sp, idx, point, normal, in_direction, cosThetaNormal, flip, n1, n2,
                         R_s, R_p, T_s, T_p, P_in, P[0], P[1]);
            return;
        }
        case RPX_MAT_COATED: { /* :1026-1182 and :1228-1391 */
            double wavelength = S->wavelengths[in_ray->wavelength_idx];
            vec3 normal = norm(orient.normal);
            vec3 in_direction = norm(ld3(in_ray->direction));
            rpx_ray sp = convert_to_sp(*in_ray, normal);
            sp.accumulated_path += sp.length * sp.refractive_index[0];
            cplx E1_amp = cy_parts(sp.E1_amp[0], sp.E1_amp[1]);
            cplx E2_amp = cy_parts(sp.E2_amp[0], sp.E2_amp[1]);
            double cosTheta = dotprod(normal, in_direction);
            double cos1 = fabs(cosTheta);
            double sin1 = sqrt(fabs(1 - cos1 * cos1));
            cplx n2 = ntab_get(S, M, 2, in_ray->wavelength_idx);
            cplx n1, n3;
            int flip;
            if (cosTheta < 0.0) {
                n1 = ntab_get(S, M, 1, in_ray->wavelength_idx);
                n3 = ntab_get(S, M, 0, in_ray->wavelength_idx);
                flip = 1;
            } else {
                n1 = ntab_get(S, M, 0, in_ray->wavelength_idx);
                n3 = ntab_get(S, M, 1, in_ray->wavelength_idx);
                flip = -1;
            }
            cplx sin2 = (n1 * CX(sin1)) / n2;
            cplx cos2 = csqrt(CX(1) - sin2 * sin2);
            cplx sin3 = (n1 * CX(sin1)) / n3;
            cplx cos3 = csqrt(CX(1) - sin3 * sin3);
            vec3 cosThetaNormal = multvs(normal, cosTheta);
            double P_in = creal(n1) * (pow(creal(E1_amp), 2.0) + pow(cimag(E1_amp), 2.0) +
                                       pow(creal(E2_amp), 2.0) + pow(cimag(E2_amp), 2.0));
            if (P_in == 0.0) return;
            cplx n1cos1 = n1 * CX(cos1);
            cplx n2cos2 = n2 * cos2;
            cplx n3cos3 = n3 * cos3;
            double dwc = 2 * M_PI * P[2] / wavelength;
            /* phi = -I*dwc*(n2 - sin2*sin2)/cos2, left-to-right */
            cplx phi = (((-_Complex_I) * CX(dwc)) * (n2 - sin2 * sin2)) / cos2;
            cplx ep1 = cexp(phi) / ((CX(4) * n2cos2) * n3cos3);
            cplx ep2 = cexp(CX(-2) * phi);
            cplx M00 = (-ep1) * ((n1cos1 - n2cos2) * (n2cos2 + n3cos3) +
                                 ((n1cos1 + n2cos2) * (n2cos2 - n3cos3)) * ep2);
            cplx M01 = ep1 * (((n1cos1 - n2cos2) * (n2cos2 - n3cos3)) * ep2 +
                              (n1cos1 + n2cos2) * (n2cos2 + n3cos3));
            cplx M10 = ep1 * ((n1cos1 - n2cos2) * (n2cos2 - n3cos3) +
                              ((n1cos1 + n2cos2) * (n2cos2 + n3cos3)) * ep2);
            cplx M11 = (-ep1) * (((n1cos1 - n2cos2) * (n2cos2 + n3cos3)) * ep2 +
                                 (n1cos1 + n2cos2) * (n2cos2 - n3cos3));
            cplx R_s = (-M00) / M01;
            cplx T_s = M10 + M11 * R_s;
            cplx n1cos2 = n1 * cos2;
            cplx n2cos1 = n2 * CX(cos1);
            cplx n2cos3 = n2 * cos3;
            cplx n3cos2 = n3 * cos2;
            M00 = (-ep1) * ((n1cos2 - n2cos1) * (n2cos3 + n3cos2) +
                            ((n1cos2 + n2cos1) * (n2cos3 - n3cos2)) * ep2);
            M01 = ep1 * (((n1cos2 - n2cos1) * (n2cos3 - n3cos2)) * ep2 +
                         (n1cos2 + n2cos1) * (n2cos3 + n3cos2));
            M10 = ep1 * ((n1cos2 - n2cos1) * (n2cos3 - n3cos2) +
                         ((n1cos2 + n2cos1) * (n2cos3 + n3cos2)) * ep2);
            M11 = (-ep1) * (((n1cos2 - n2cos1) * (n2cos3 + n3cos2)) * ep2 +
                            (n1cos2 + n2cos1) * (n2cos3 - n3cos2));
            cplx R_p = (-M00) / M01;
            cplx T_p = M10 + M11 * R_p;
            R_s = R_s * E1_amp;
            R_p = R_p * E2_amp;
            double aspect = sqrt(creal(cos3) / cos1);
            T_s = T_s * (E1_amp * CX(aspect));
            T_p = T_p * (E2_amp * CX(aspect));
            fresnel_emit(out, &sp, idx, point, normal, in_direction, cosThetaNormal, flip, n1, n3,
                         R_s, R_p, T_s, T_p, P_in, P[0], P[1]);
            return;
        }
        case RPX_MAT_GRATING: { /* :1472-1542 */
            vec3 normal = norm(orient.normal);
            vec3 tangent = norm(orient.tangent);
            vec3 tangent2 = cross(normal, tangent);
            double wavelen = S->wavelengths[in_ray->wavelength_idx];
            double line_spacing = 1000.0 / P[0];
            int order = (int)P[1];
            vec3 reflected = norm(ld3(in_ray->direction));
            double k_z = dotprod(normal, reflected);
            double k_y = dotprod(tangent2, reflected);
            double k_x = dotprod(tangent, reflected);
            int sign = (k_z < 0.0) ? 1 : -1;
            double n_ray_re = in_ray->refractive_index[0];
            k_x = k_x - order * wavelen / (line_spacing * n_ray_re);
            k_z = 1 - (k_x * k_x) - (k_y * k_y);
            if (k_z < 0) return;
            k_z = sign * sqrt(k_z);
            reflected = multvs(tangent, k_x);
            reflected = addvv(reflected, multvs(tangent2, k_y));
            reflected = addvv(reflected, multvs(normal, k_z));
            rpx_ray sp = convert_to_sp(*in_ray, normal);
            sp.accumulated_path += sp.length * sp.refractive_index[0];
            st3(sp.origin, point);
            st3(sp.normal, normal);
            st3(sp.direction, reflected);
            sp.E1_amp[0] = -sp.E1_amp[0] * P[2];
            sp.E1_amp[1] = -sp.E1_amp[1] * P[2];
            sp.E2_amp[0] = sp.E2_amp[0] * P[2];
            sp.E2_amp[1] = sp.E2_amp[1] * P[2];
            sp.parent_idx = idx;
            sp.ray_type_id |= RPX_REFL_RAY;
            sp.phase += 1000.0 * dotprod(subvv(ld3(P + 3), point), tangent) * order * 2 * M_PI /
                        line_spacing;
            add_ray(out, &sp);
            return;
        }
        case RPX_MAT_CIRC_APERTURE: { /* :1641-1674 */
            double width = P[2];
            double r = sqrt(mag_sq(subvv(ld3(P + 4), point)));
            if (r > P[0]) return;
            double atten = 0.5 + 0.5 * erf((P[1] - r) / width);
            if (P[3] != 0.0) atten = 1 - atten;
            vec3 normal = norm(orient.normal);
            rpx_ray sp = convert_to_sp(*in_ray, normal);
            sp.accumulated_path += sp.length * sp.refractive_index[0];
            st3(sp.origin, point);
            st3(sp.normal, normal);
            sp.parent_idx = idx;
            sp.ray_type_id &= ~RPX_REFL_RAY;
            sp.E1_amp[0] *= atten; sp.E1_amp[1] *= atten;
            sp.E2_amp[0] *= atten; sp.E2_amp[1] *= atten;
            add_ray(out, &sp);
            return;
        }
        case RPX_MAT_RECT_APERTURE: { /* :1718-1763 */
            double width = P[4];
            double x = P[2] / 2., y = P[3] / 2.;
            vec3 p = subvv(point, ld3(P + 6));
            double px = dotprod(p, orient.tangent);
            double py = dotprod(p, cross(orient.normal, orient.tangent));
            if (fabs(px) > P[0] / 2.) return;
            if (fabs(py) > P[1] / 2.) return;
            double atten = 0.5 - 0.5 * erf((px - x) / width);
            atten *= 0.5 - 0.5 * erf(-(px + x) / width);
            atten *= 0.5 - 0.5 * erf((py - y) / width);
            atten *= 0.5 - 0.5 * erf(-(py + y) / width);
            if (P[5] != 0.0) atten = 1 - atten;
            vec3 normal = norm(orient.normal);
            rpx_ray sp = convert_to_sp(*in_ray, normal);
            sp.accumulated_path += sp.length * sp.refractive_index[0];
            st3(sp.origin, point);
            st3(sp.normal, normal);
            sp.parent_idx = idx;
            sp.ray_type_id &= ~RPX_REFL_RAY;
            sp.E1_amp[0] *= atten; sp.E1_amp[1] *= atten;
            sp.E2_amp[0] *= atten; sp.E2_amp[1] *= atten;
            add_ray(out, &sp);
            return;
        }
        default: return;
    }
}

/* InterfaceMaterial.eval_parabasal_ray_c per para_model */
static rpx_para material_eval_para(const rpx_scene* S, const rpx_material* M, const rpx_ray* base_ray,
                                   vec3 direction, vec3 point, orient_t orient, uint32_t ray_type_id) {
    rpx_para po;
    vec3 normal = norm(orient.normal);
    if (M->para_model == RPX_PARA_SNELL) { /* cmaterials.pyx:683-724, 1393-1434 */
        direction = norm(direction);
        double cosTheta = dotprod(normal, direction);
        vec3 cosThetaNormal = multvs(normal, cosTheta);
        double n1, n2;
        int flip;
        if (cosTheta < 0.0) {
            n1 = creal(ntab_get(S, M, 1, base_ray->wavelength_idx));
            n2 = creal(ntab_get(S, M, 0, base_ray->wavelength_idx));
            flip = 1;
        } else {
            n1 = creal(ntab_get(S, M, 0, base_ray->wavelength_idx));
            n2 = creal(ntab_get(S, M, 1, base_ray->wavelength_idx));
            flip = -1;
        }
        vec3 out_dir;
        if (ray_type_id & RPX_REFL_RAY) {
            out_dir = subvv(direction, multvs(cosThetaNormal, 2));
        } else {
            vec3 tangent = subvv(direction, cosThetaNormal);
            vec3 tg2 = multvs(tangent, n1 / n2);
            double tan_mag_sq = mag_sq(tg2);
            double c2 = sqrt(1 - tan_mag_sq);
            out_dir = subvv(tg2, multvs(normal, c2 * flip));
        }
        st3(po.direction, out_dir);
    } else if (M->para_model == RPX_PARA_GRATING) { /* :1544-1599 */
        const double* P = M->p;
        vec3 tangent = norm(orient.tangent);
        vec3 tangent2 = cross(normal, tangent);
        double wavelen = S->wavelengths[base_ray->wavelength_idx];
        double line_spacing = 1000.0 / P[0];
        int order = (int)P[1];
        vec3 reflected = norm(direction);
        double k_z = dotprod(normal, reflected);
        double k_y = dotprod(tangent2, reflected);
        double k_x = dotprod(tangent, reflected);
        int sign = (k_z < 0.0) ? 1 : -1;
        k_x = k_x - order * wavelen / (line_spacing * base_ray->refractive_index[0]);
        k_z = 1 - (k_x * k_x) - (k_y * k_y);
        k_z = sign * sqrt(k_z); /* evanescent -> NaN, the reference only prints */
        reflected = multvs(tangent, k_x);
        reflected = addvv(reflected, multvs(tangent2, k_y));
        reflected = addvv(reflected, multvs(normal, k_z));
        st3(po.direction, reflected);
    } else { /* default, ctracer.pyx:1588-1610 */
        if (ray_type_id & RPX_REFL_RAY) {
            double cosTheta = dotprod(normal, direction);
            vec3 cosThetaNormal = multvs(normal, cosTheta);
            st3(po.direction, subvv(direction, multvs(cosThetaNormal, 2)));
        } else {
            st3(po.direction, direction);
        }
    }
    st3(po.origin, point);
    st3(po.normal, normal);
    po.length = ORACLE_INF;
    return po;
}

/* --------------------------------------------------- trace loops, ctracer.pyx */
/* FaceList.intersect_c for every face set, ctracer.pyx:1882-1904 + 2093-2104.
 * Mutates ray->length / ray->end_face_idx; returns nearest face index or -1. */
static int nearest_hit(const rpx_scene* S, rpx_ray* ray, vec3 point) {
    int nearest_idx = -1;
    for (int j = 0; j < S->n_face_sets; j++) {
        const rpx_face_set* fs = &S->face_sets[j];
        vec3 p1 = transform_pt(&fs->inv_trans, ld3(ray->origin));
        vec3 p2 = transform_pt(&fs->inv_trans, point);
        for (int i = fs->face_begin; i < fs->face_end; i++) {
            const rpx_face* f = &S->faces[i];
            double dist = face_intersect(S, f, p1, p2, 1);
            if (f->tolerance < dist && dist < ray->length) {
                ray->length = dist;
                ray->end_face_idx = (uint32_t)i;
                nearest_idx = i;
                s_hit_piece = s_piece, s_hit_u = s_u, s_hit_v = s_v;
            }
        }
    }
    return nearest_idx;
}

/* FaceList.intersect_one_face_c, ctracer.pyx:1861-1879 */
static int one_face_hit(const rpx_scene* S, rpx_ray* ray, vec3 point, int face_idx) {
    const rpx_face* f = &S->faces[face_idx];
    const rpx_face_set* fs = &S->face_sets[f->face_set];
    vec3 p1 = transform_pt(&fs->inv_trans, ld3(ray->origin));
    vec3 p2 = transform_pt(&fs->inv_trans, point);
    double dist = face_intersect(S, f, p1, p2, 1);
    if (f->tolerance < dist && dist < ray->length) {
        ray->length = dist;
        ray->end_face_idx = (uint32_t)face_idx;
        s_hit_piece = s_piece, s_hit_u = s_u, s_hit_v = s_v;
        return face_idx;
    }
    return -1;
}

/* trace_segment_c (ctracer.pyx:2062-2118) when only_face < 0, trace_one_face_segment_c
 * (ctracer.pyx:2121-2170) otherwise.  rays_out must hold 2*n records. */
uint64_t rpxo_trace_segment_ex(const rpx_scene* S, rpx_ray* rays, uint64_t n, double max_length_d,
                               rpx_ray* rays_out, uint32_t* face_counts, int only_face);

uint64_t rpxo_trace_segment(const rpx_scene* S, rpx_ray* rays, uint64_t n, double max_length_d,
                            rpx_ray* rays_out, uint32_t* face_counts) {
    return rpxo_trace_segment_ex(S, rays, n, max_length_d, rays_out, face_counts, -1);
}

uint64_t rpxo_trace_segment_ex(const rpx_scene* S, rpx_ray* rays, uint64_t n, double max_length_d,
                               rpx_ray* rays_out, uint32_t* face_counts, int only_face) {
    float max_length = (float)max_length_d; /* `float max_length`, ctracer.pyx:2066 */
    childbuf_t out = {rays_out, 0};
    uint64_t n_out = 0;
    for (uint64_t i = 0; i < n; i++) {
        rpx_ray* ray = &rays[i];
        ray->length = max_length;
        ray->end_face_idx = (uint32_t)-1;
        vec3 point = addvv(ld3(ray->origin), multvs(ld3(ray->direction), max_length));
        int nearest_idx = only_face < 0 ? nearest_hit(S, ray, point) : one_face_hit(S, ray, point, only_face);
        if (nearest_idx >= 0) {
            const rpx_face* face = &S->faces[nearest_idx];
            if (face_counts) face_counts[nearest_idx] += 1;
            point = addvv(ld3(ray->origin), multvs(ld3(ray->direction), ray->length));
            orient_t orient = compute_orientation(S, face, point);
            out.rays = rays_out + n_out;
            out.n = 0;
            material_eval(S, &S->materials[face->material], ray, (uint32_t)i, point, orient, &out);
            n_out += (uint64_t)out.n;
        }
    }
    return n_out;
}

/* trace_gausslet_c + trace_parabasal_rays (ctracer.pyx:2214-2281, 2350-2385) when only_face < 0,
 * trace_one_face_gausslet_c (ctracer.pyx:2284-2347) otherwise */
uint64_t rpxo_trace_gausslet_ex(const rpx_scene* S, rpx_gausslet* gs, uint64_t n, double max_length,
                                rpx_gausslet* gs_out, uint32_t* face_counts, int only_face);

uint64_t rpxo_trace_gausslet(const rpx_scene* S, rpx_gausslet* gs, uint64_t n, double max_length,
                             rpx_gausslet* gs_out, uint32_t* face_counts) {
    return rpxo_trace_gausslet_ex(S, gs, n, max_length, gs_out, face_counts, -1);
}

uint64_t rpxo_trace_gausslet_ex(const rpx_scene* S, rpx_gausslet* gs, uint64_t n, double max_length,
                                rpx_gausslet* gs_out, uint32_t* face_counts, int only_face) {
    uint64_t n_out = 0;
    rpx_ray child[2];
    for (uint64_t i = 0; i < n; i++) {
        rpx_gausslet* g = &gs[i];
        rpx_ray* ray = &g->base_ray;
        ray->end_face_idx = (uint32_t)-1;
        vec3 point = addvv(ld3(ray->origin), multvs(ld3(ray->direction), max_length));
        int nearest_idx = only_face < 0 ? nearest_hit(S, ray, point) : one_face_hit(S, ray, point, only_face);
        if (nearest_idx < 0) continue;
        const rpx_face* face = &S->faces[nearest_idx];
        const rpx_face_set* fs = &S->face_sets[face->face_set];
        const rpx_material* M = &S->materials[face->material];
        if (face_counts) face_counts[nearest_idx] += 1;
        point = addvv(ld3(ray->origin), multvs(ld3(ray->direction), ray->length));
        orient_t orient = compute_orientation(S, face, point);
        childbuf_t cb = {child, 0};
        material_eval(S, M, ray, (uint32_t)i, point, orient, &cb);
        /* trace_parabasal_rays */
        vec3 ppoint[6];
        orient_t porient[6];
        int ok = 1;
        for (int j = 0; j < 6; j++) {
            rpx_para* pr = &g->para[j];
            vec3 ray_end = addvv(ld3(pr->origin), multvs(ld3(pr->direction), max_length));
            /* FaceList.intersect_para_c, ctracer.pyx:1915-1928 */
            vec3 p1 = transform_pt(&fs->inv_trans, ld3(pr->origin));
            vec3 p2 = transform_pt(&fs->inv_trans, ray_end);
            double dist = face_intersect(S, face, p1, p2, 0);
            if (face->tolerance < dist && dist < pr->length) {
                pr->length = dist;
                s_hit_piece = s_piece, s_hit_u = s_u, s_hit_v = s_v; /* the parabasal ray's own piece / uv (intersect_para_c returns its intersect_t) */
            } else {
                ok = 0;
                break;
            }
            ppoint[j] = addvv(ld3(pr->origin), multvs(ld3(pr->direction), pr->length));
            porient[j] = compute_orientation(S, face, ppoint[j]);
        }
        if (!ok) continue;
        for (int c = 0; c < cb.n; c++) {
            rpx_gausslet* o = &gs_out[n_out];
            o->base_ray = child[c];
            for (int j = 0; j < 6; j++) {
                o->para[j] = material_eval_para(S, M, &child[c], ld3(g->para[j].direction), ppoint[j],
                                                porient[j], o->base_ray.ray_type_id);
            }
            n_out++;
        }
    }
    /* new_gausslets.reset_length_c(max_length), ctracer.pyx:2280 */
    for (uint64_t i = 0; i < n_out; i++) {
        gs_out[i].base_ray.length = max_length;
        for (int j = 0; j < 6; j++) gs_out[i].para[j].length = max_length;
    }
    return n_out;
}

/* --------------------------------------------------- capture planes, ctracer.pyx:1981-2058
 * Inner loop of select_ray_intersections (:1995-2009) / select_gausslet_intersections
 * (:2034-2048) over ONE collection: S holds only the capture FaceList.  A COPY of each ray is
 * intersected between its origin and origin + direction * length (FaceList.intersect_c mutates
 * the copy: length = distance, end_face_idx = face.idx); hits are appended with
 * wavelength_idx += wl_offset.  face_ids[i] stands for the Python attribute Face.idx of capture
 * face i (NULL = position).  The np.unique re-mapping of wavelength_idx (:2011-2016) is table
 * work done by the caller.  Returns the number of records appended to `out`.                  */
uint64_t rpxo_capture_rays(const rpx_scene* S, const uint32_t* face_ids, const rpx_ray* rays, uint64_t n,
                           uint32_t wl_offset, rpx_ray* out) {
    uint64_t n_out = 0;
    for (uint64_t i = 0; i < n; i++) {
        rpx_ray ray = rays[i];
        vec3 point = addvv(ld3(ray.origin), multvs(ld3(ray.direction), ray.length));
        int idx = nearest_hit(S, &ray, point);
        if (idx >= 0) {
            ray.end_face_idx = face_ids ? face_ids[idx] : (uint32_t)idx;
            ray.wavelength_idx += wl_offset;
            out[n_out++] = ray;
        }
    }
    return n_out;
}

uint64_t rpxo_capture_gausslets(const rpx_scene* S, const uint32_t* face_ids, const rpx_gausslet* gs, uint64_t n,
                                uint32_t wl_offset, rpx_gausslet* out) {
    uint64_t n_out = 0;
    for (uint64_t i = 0; i < n; i++) {
        rpx_gausslet g = gs[i];
        rpx_ray* ray = &g.base_ray;
        vec3 point = addvv(ld3(ray->origin), multvs(ld3(ray->direction), ray->length));
        int idx = nearest_hit(S, ray, point);
        if (idx >= 0) {
            ray->end_face_idx = face_ids ? face_ids[idx] : (uint32_t)idx;
            ray->wavelength_idx += wl_offset;
            out[n_out++] = g;
        }
    }
    return n_out;
}

/* -------------------------------------------- unit entry points for KAT pins */
double rpxo_face_intersect(const rpx_scene* S, int face, const double* p1, const double* p2,
                           int is_base_ray) {
    return face_intersect(S, &S->faces[face], ld3(p1), ld3(p2), is_base_ray);
}
void rpxo_face_normal(const rpx_scene* S, int face, const double* p, double* out) {
    st3(out, face_normal(S, &S->faces[face], ld3(p)));
}
void rpxo_orientation(const rpx_scene* S, int face, const double* point, double* normal,
                      double* tangent) {
    orient_t o = compute_orientation(S, &S->faces[face], ld3(point));
    st3(normal, o.normal);
    st3(tangent, o.tangent);
}
void rpxo_convert_to_sp(const rpx_ray* in, const double* normal, rpx_ray* out) {
    *out = convert_to_sp(*in, ld3(normal));
}
/* eval_child_ray with explicit normal/tangent, ctracer.pyx:1620-1632; returns child count */
int rpxo_material_eval(const rpx_scene* S, int material, const rpx_ray* in_ray, uint32_t idx,
                       const double* point, const double* normal, const double* tangent,
                       rpx_ray* out2) {
    orient_t o = {ld3(normal), ld3(tangent)};
    childbuf_t cb = {out2, 0};
    material_eval(S, &S->materials[material], in_ray, idx, ld3(point), o, &cb);
    return cb.n;
}
void rpxo_material_eval_para(const rpx_scene* S, int material, const rpx_ray* base_ray,
                             const double* direction, const double* point, const double* normal,
                             const double* tangent, uint32_t ray_type_id, rpx_para* out) {
    orient_t o = {ld3(normal), ld3(tangent)};
    *out = material_eval_para(S, &S->materials[material], base_ray, ld3(direction), ld3(point), o,
                              ray_type_id);
}
double rpxo_distortion_z(const rpx_scene* S, int dist, double x, double y) {
    return distortion_z(S, &S->distortions[dist], x, y);
}
void rpxo_distortion_zgrad(const rpx_scene* S, int dist, double x, double y, double* out) {
    st3(out, distortion_zgrad(S, &S->distortions[dist], x, y));
}
int rpxo_shape_inside(const rpx_scene* S, int face, double x, double y) {
    return shape_inside(S, &S->faces[face], x, y);
}
double rpxo_implicit_eval(const rpx_scene* S, int off, int len, const double* p) {
    return implicit_eval(S, off, len, ld3(p));
}
double rpxo_zernike_R(double r, int k, int n, int m, double* ws3k, int kmax) {
    zws_t ws = {{ws3k, ws3k + kmax, ws3k + 2 * kmax}};
    return zernike_R(r, k, n, m, &ws);
}
double rpxo_zernike_Rprime(double r, int k, int n, int m, double* ws3k, int kmax) {
    zws_t ws = {{ws3k, ws3k + kmax, ws3k + 2 * kmax}};
    return zernike_Rprime(r, k, n, m, &ws);
}
double rpxo_zernike_R_over_r(double r, int k, int n, int m, double* ws3k, int kmax) {
    zws_t ws = {{ws3k, ws3k + kmax, ws3k + 2 * kmax}};
    return zernike_R_over_r(r, k, n, m, &ws);
}
/* =============================================================== E-field summation
 * raypier/core/cfields.pyx + the gausslet front end of raypier/core/fields.py.             */

/* evaluate_neighbours_gc, core/fields.py:114-137 (numpy there, a loop here): the six parabasal
 * rays of every gausslet in the (E, H) basis of its base ray.  x, y, dx, dy are n x 6.      */
void rpxo_evaluate_neighbours_gc(const rpx_gausslet* gs, uint64_t n, double* x, double* y, double* dx,
                                 double* dy) {
    for (uint64_t i = 0; i < n; i++) {
        const rpx_ray* b = &gs[i].base_ray;
        vec3 origin = ld3(b->origin), direction = ld3(b->direction), E = ld3(b->E_vector);
        vec3 H = cross(E, direction); /* numpy.cross(E, direction) */
        for (int j = 0; j < RPX_NPARA; j++) {
            const rpx_para* p = &gs[i].para[j];
            vec3 off = subvv(ld3(p->origin), origin), nd = ld3(p->direction);
            x[i * 6 + j] = dotprod(off, E);
            y[i * 6 + j] = dotprod(off, H);
            double dz = dotprod(nd, direction);
            dx[i * 6 + j] = dotprod(nd, E) / dz;
            dy[i * 6 + j] = dotprod(nd, H) / dz;
        }
    }
}

/* evaluate_one_mode, cfields.pyx:156-214: closed-form 3x3 normal equations of the least-squares
 * fit; out[0..2] = the complex (A, B, C) of one gausslet.                                   */
static void evaluate_one_mode(cplx* out, const double* x, const double* y, const double* dx, const double* dy,
                              double blending, int size) {
    double a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0, b0 = 0, b1 = 0, b2 = 0;
    double im[3], re[3];
    for (int i = 0; i < size; i++) { /* imaginary parts, :170-181 */
        double xi2 = x[i] * x[i], yi2 = y[i] * y[i];
        a00 += xi2 * xi2;
        a01 += ((2 * xi2) * x[i]) * y[i];
        a02 += xi2 * yi2;
        a11 += (4 * xi2) * yi2;
        a12 += ((2 * x[i]) * y[i]) * yi2;
        a22 += yi2 * yi2;
        b0 += xi2;
        b1 += (2 * x[i]) * y[i];
        b2 += yi2;
    }
    {
        double den = (((((-a00) * a11) * a22) + (a00 * (a12 * a12))) + ((a01 * a01) * a22)) -
                     (((2.0 * a01) * a02) * a12) + ((a02 * a02) * a11);
        im[0] = (blending * ((((-b0) * ((a11 * a22) - (a12 * a12))) + (b1 * ((a01 * a22) - (a02 * a12)))) -
                             (b2 * ((a01 * a12) - (a02 * a11))))) / den;
        im[1] = ((-blending) * (((b0 * ((a01 * a22) - (a02 * a12))) - (b1 * ((a00 * a22) - (a02 * a02)))) +
                                (b2 * ((a00 * a12) - (a01 * a02))))) / den;
        im[2] = (blending * ((((-b0) * ((a01 * a12) - (a02 * a11))) + (b1 * ((a00 * a12) - (a01 * a02)))) -
                             (b2 * ((a00 * a11) - (a01 * a01))))) / den;
    }
    a00 = a01 = a02 = a11 = a12 = a22 = b0 = b1 = b2 = 0;
    for (int i = 0; i < size; i++) { /* real parts, :196-206 */
        double xi2 = x[i] * x[i], yi2 = y[i] * y[i];
        a00 += xi2;
        a01 += x[i] * y[i];
        a11 += xi2 + yi2;
        a22 += yi2;
        b0 += dx[i] * x[i];
        b1 += (dx[i] * y[i]) + (dy[i] * x[i]);
        b2 += dy[i] * y[i];
    }
    a12 = a01 * a01;
    {
        double den = ((a00 * a12) - ((a00 * a11) * a22)) + (a12 * a22);
        re[0] = ((((-a12) * b2) + ((a01 * a22) * b1)) + (b0 * (a12 - (a11 * a22)))) / den;
        re[1] = (-((((a00 * a01) * b2) - ((a00 * a22) * b1)) + ((a01 * a22) * b0))) / den;
        re[2] = ((((a00 * a01) * b1) - (a12 * b0)) - (b2 * ((a00 * a11) - a12))) / den;
    }
    for (int k = 0; k < 3; k++) out[k] = from_parts(re[k], im[k]);
}

/* evaluate_modes, cfields.pyx:217-228; modes is n x 3 complex */
void rpxo_evaluate_modes(const double* x, const double* y, const double* dx, const double* dy, uint64_t n,
                         int row_size, double blending, double* modes) {
    cplx* out = (cplx*)modes;
    for (uint64_t i = 0; i < n; i++)
        evaluate_one_mode(out + 3 * i, x + i * row_size, y + i * row_size, dx + i * row_size, dy + i * row_size,
                          blending, row_size);
}

/* calc_mode_U, cfields.pyx:118-153.  Mixed real/complex operations are written the way Cython
 * lowers them (real promoted to r + 0i, native C complex * and /; oracle/_ref/build/cfields.c). */
static cplx calc_mode_U(cplx A, cplx B, cplx C, cplx detG0, vec3 pt, vec3 E, vec3 H, vec3 direction, cplx k,
                        double phase, double inv_root_area, cplx rootI) {
    double x = dotprod(pt, E), y = dotprod(pt, H), z = dotprod(pt, direction);
    cplx denom = ((CX(1) + (CX(z) * (A + C))) + (CX(z * z) * detG0)) * CX(2);
    cplx AA = (A + (CX(z) * detG0)) / denom;
    cplx CC = (C + (CX(z) * detG0)) / denom;
    cplx t1 = B * CX((2.0 * x) * y);
    cplx U = cexp(_Complex_I *
                  (CX(phase) + (k * (((CX(z) + (AA * CX(x * x))) + (t1 / denom)) + (CC * CX(y * y))))));
    cplx q = csqrt((((CX(1) + (CX(z) * A)) * (CX(1) + (CX(z) * C))) - ((CX(z) * B) * (CX(z) * B))) * _Complex_I);
    U = U / q;
    U = U * (CX(inv_root_area) * rootI);
    return U;
}

/* sum_gaussian_modes, cfields.pyx:51-115: E-field at npt points as the sum over rays of
 * general astigmatic Gaussian modes.  out is npt x 3 complex, zeroed here (np.zeros, :68).
 * Accumulation order = the reference's: rays outermost, in order.                         */
void rpxo_sum_gaussian_modes(const rpx_ray* rays, uint64_t n_rays, const double* modes_, const double* wavelengths,
                             const double* points, uint64_t npt, double time_ps, double* out_) {
    const cplx* modes = (const cplx*)modes_;
    cplx* out = (cplx*)out_;
    const cplx rootI = csqrt(_Complex_I); /* module-level rootI, :45 */
    double c = 0.299792458;
    c *= time_ps;
    for (uint64_t i = 0; i < 3 * npt; i++) out[i] = CX(0);
    for (uint64_t iray = 0; iray < n_rays; iray++) {
        rpx_ray ray = rays[iray];
        vec3 E = norm(ld3(ray.E_vector));
        vec3 dir = ld3(ray.direction);
        vec3 H = norm(cross(dir, E));
        double k = (2000.0 * M_PI) / wavelengths[ray.wavelength_idx];
        double n_re = ray.refractive_index[0];
        double phase = (ray.phase + (ray.accumulated_path * k)) - ((c * k) / n_re);
        cplx kz = from_parts(ray.refractive_index[0], ray.refractive_index[1]);
        kz = kz * CX(k);
        double invk = 2. / creal(kz);
        cplx A = modes[3 * iray], B = modes[3 * iray + 1], C = modes[3 * iray + 2];
        double inv_root_area = sqrt(sqrt((cimag(A) * cimag(C)) - (cimag(B) * cimag(B))) * (2.0 / M_PI));
        A = from_parts(creal(A), cimag(A) * invk); /* __Pyx_SET_CIMAG */
        B = from_parts(creal(B), cimag(B) * invk);
        C = from_parts(creal(C), cimag(C) * invk);
        cplx detG0 = (A * C) - (B * B);
        cplx E1a = from_parts(ray.E1_amp[0], ray.E1_amp[1]), E2a = from_parts(ray.E2_amp[0], ray.E2_amp[1]);
        vec3 origin = ld3(ray.origin);
        for (uint64_t ipt = 0; ipt < npt; ipt++) {
            vec3 pt = subvv(ld3(points + 3 * ipt), origin);
            cplx U = calc_mode_U(A, B, C, detG0, pt, E, H, dir, kz, phase, inv_root_area, rootI);
            cplx E1 = E1a * U, E2 = E2a * U;
            out[3 * ipt + 0] = out[3 * ipt + 0] + ((E1 * CX(E.x)) + (E2 * CX(H.x)));
            out[3 * ipt + 1] = out[3 * ipt + 1] + ((E1 * CX(E.y)) + (E2 * CX(H.y)));
            out[3 * ipt + 2] = out[3 * ipt + 2] + ((E1 * CX(E.z)) + (E2 * CX(H.z)));
        }
    }
}


/* ---- the plain-ray front end of the E-field summation: fields.py eval_Efield_from_rays (:206-229) */

/* project_to_sphere, core/fields.py:50-77 (numpy there, a loop here with numpy's own association:
 * .sum(axis=1) over three components is ((a0 + a1) + a2); a = 1 is folded as in the source).
 * Rays are updated IN PLACE where the ray line meets the sphere (selector[i] = 1); rays with a negative
 * discriminant are left untouched (selector[i] = 0; the reference returns rays[selector]).
 * Returns the number of selected rays.                                                         */
uint64_t rpxo_project_to_sphere(rpx_ray* rays, uint64_t n, const double* centre, double radius, uint8_t* selector) {
    uint64_t kept = 0;
    for (uint64_t i = 0; i < n; i++) {
        rpx_ray* r = &rays[i];
        double ox = r->origin[0] - centre[0], oy = r->origin[1] - centre[1], oz = r->origin[2] - centre[2];
        double dx = r->direction[0], dy = r->direction[1], dz = r->direction[2];
        double c = ((ox * ox + oy * oy) + oz * oz) - (radius * radius);
        double b = 2 * ((dx * ox + dy * oy) + dz * oz);
        double d = b * b - 4 * c;
        selector[i] = (d >= 0.0);
        if (!(d >= 0.0)) continue;
        d = sqrt(d);
        double root1 = (-b + d) / 2, root2 = (-b - d) / 2;
        double alpha = root2 < root1 ? root2 : root1; /* .min(axis=1): most negative path */
        r->origin[0] += alpha * dx;
        r->origin[1] += alpha * dy;
        r->origin[2] += alpha * dz;
        r->accumulated_path += alpha * r->refractive_index[0];
        kept++;
    }
    return kept;
}

/* evaluate_neighbours, core/fields.py:80-111: every ray with all `row` neighbours present (mask) gets the
 * (x, y) of its neighbours projected along THEIR direction onto the plane through its own origin, in its
 * (E, H = E x direction) basis, and the direction differences (dx, dy).  x, y, dx, dy are n_kept x row,
 * written for the kept rays in order; mask is n bytes.  Returns n_kept.                         */
uint64_t rpxo_evaluate_neighbours(const rpx_ray* rays, uint64_t n, const int32_t* nb, int row, double* x, double* y,
                                  double* dx, double* dy, uint8_t* mask) {
    uint64_t k = 0;
    for (uint64_t i = 0; i < n; i++) {
        int all = 1;
        for (int j = 0; j < row; j++) all = all && (nb[i * row + j] >= 0);
        mask[i] = (uint8_t)all;
        if (!all) continue;
        vec3 origin = ld3(rays[i].origin), direction = ld3(rays[i].direction), E = ld3(rays[i].E_vector);
        vec3 H = cross(E, direction);
        for (int j = 0; j < row; j++) {
            const rpx_ray* q = &rays[nb[i * row + j]];
            vec3 off = subvv(ld3(q->origin), origin), nd = ld3(q->direction);
            double alpha = -dotprod(off, direction) / dotprod(nd, direction);
            vec3 proj = addvv(off, multvs(nd, alpha));
            x[k * row + j] = dotprod(proj, E);
            y[k * row + j] = dotprod(proj, H);
            double dz = dotprod(nd, direction);
            dx[k * row + j] = dotprod(nd, E) / dz;
            dy[k * row + j] = dotprod(nd, H) / dz;
        }
        k++;
    }
    return k;
}

int rpxo_sizeof_ray(void) { return (int)sizeof(rpx_ray); }
int rpxo_sizeof_gausslet(void) { return (int)sizeof(rpx_gausslet); }
