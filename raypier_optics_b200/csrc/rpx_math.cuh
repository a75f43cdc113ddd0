// rpx_math.cuh -- fp64 vector / complex helpers for the device code of librpx.
//
// These are the building blocks the reference keeps in raypier/core/ctracer.pyx:83-262
// (vector_t maths) and gets from <complex.h> (csqrt / cexp / cabs, complex * and /).
// Written for sm_100a; everything is __forceinline__ so the face / material code
// compiles into straight-line fp64 (DFMA/DMUL/DADD) with no calls.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define RPX_DEV __device__ __forceinline__

namespace rpx {

struct vec3 {
    double x, y, z;
};

RPX_DEV vec3 v3(double x, double y, double z) {
    vec3 v;
    v.x = x;
    v.y = y;
    v.z = z;
    return v;
}
RPX_DEV vec3 ld3(const double* p) { return v3(p[0], p[1], p[2]); }
RPX_DEV vec3 operator+(vec3 a, vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
RPX_DEV vec3 operator-(vec3 a, vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
RPX_DEV vec3 operator*(vec3 a, double b) { return v3(a.x * b, a.y * b, a.z * b); }
RPX_DEV vec3 neg(vec3 a) { return v3(-a.x, -a.y, -a.z); }
RPX_DEV double dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RPX_DEV double mag_sq(vec3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
RPX_DEV vec3 cross(vec3 a, vec3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// ---- branch-free fp64 primitives ---------------------------------------------------
// CUDA's IEEE double division / sqrt / rsqrt / __drcp_rn each expand to 35-60 executed SASS
// instructions with a guarded slow path (BSSY/BRA/BSYNC, exponent tests).  On this path they
// dominated the instruction stream (ncu: ~2000 of 3050 warp instructions per ray were not
// fp64 arithmetic).  The versions below are MUFU seed + two Newton steps: 5-7 instructions,
// no branches, <= 2 ulp -- seven orders of magnitude inside the 1e-9 / 1e-10 parity
// tolerances.  They assume finite, normal, non-zero operands; the helpers that can meet a
// zero (sqrt_, fdiv) special-case it so IEEE results (0, +-inf, NaN) are kept there.
RPX_DEV double rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
RPX_DEV double rsqrt_(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double h = 0.5 * x;
    double e = fma(-h * r, r, 0.5);
    r = fma(r, e, r);
    e = fma(-h * r, r, 0.5);
    return fma(r, e, r);
}
// sqrt(x) = x * rsqrt(x); sqrt(0) = 0 exactly (normal incidence gives sin = sqrt(0)).
RPX_DEV double sqrt_(double x) {
    double r = x * rsqrt_(x);
    return (x == 0.0) ? 0.0 : r;
}
// a / b with the IEEE result when b == 0 (+-inf or NaN decide hit/miss for rays parallel to
// a face, which is the COMMON case for axis-aligned sources); one multiply otherwise.
RPX_DEV double fdiv(double a, double b) {
    // a * (+-inf) reproduces IEEE a / (+-0): +-inf with the right sign, NaN for 0 / 0 -- no branch
    const double q = a * rcp(b);
    const double z = a * copysign(__longlong_as_double(0x7ff0000000000000LL), b);
    return (b == 0.0) ? z : q;
}
// norm_ of the reference divides each component by the magnitude (ctracer.pyx:251-256):
// one sqrt + three divisions.  Here: one rsqrt + three multiplies.  A zero vector still
// yields NaN (0 * inf), as 0/0 does in the reference.
RPX_DEV vec3 norm(vec3 a) {
    double inv = rsqrt_(a.x * a.x + a.y * a.y + a.z * a.z);
    return v3(a.x * inv, a.y * inv, a.z * inv);
}
RPX_DEV double mag(vec3 a) { return sqrt_(a.x * a.x + a.y * a.y + a.z * a.z); }
RPX_DEV double sep(vec3 p1, vec3 p2) {
    double a = p2.x - p1.x, b = p2.y - p1.y, c = p2.z - p1.z;
    return sqrt_((a * a) + (b * b) + (c * c));
}

// transform_t (ctracer.pxd:67-69) : row-major 3x3 + translation, 12 doubles
RPX_DEV vec3 transform_pt(const double* t, vec3 p) {
    return v3(p.x * t[0] + p.y * t[1] + p.z * t[2] + t[9], p.x * t[3] + p.y * t[4] + p.z * t[5] + t[10],
              p.x * t[6] + p.y * t[7] + p.z * t[8] + t[11]);
}
RPX_DEV vec3 rotate_v(const double* t, vec3 p) {
    return v3(p.x * t[0] + p.y * t[1] + p.z * t[2], p.x * t[3] + p.y * t[4] + p.z * t[5],
              p.x * t[6] + p.y * t[7] + p.z * t[8]);
}

// ------------------------------------------------------------------ complex
struct cplx {
    double re, im;
};
RPX_DEV cplx cx(double re, double im) {
    cplx c;
    c.re = re;
    c.im = im;
    return c;
}
RPX_DEV cplx operator+(cplx a, cplx b) { return cx(a.re + b.re, a.im + b.im); }
RPX_DEV cplx operator-(cplx a, cplx b) { return cx(a.re - b.re, a.im - b.im); }
RPX_DEV cplx operator-(cplx a) { return cx(-a.re, -a.im); }
RPX_DEV cplx operator*(cplx a, cplx b) {
    return cx(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
RPX_DEV cplx operator*(cplx a, double b) { return cx(a.re * b, a.im * b); }
// a / b = a * conj(b) / |b|^2 with ONE reciprocal (libgcc's __divdc3 uses Smith's scheme
// with three divisions to survive |b| near DBL_MAX/DBL_MIN; refractive indices and
// Fresnel terms are O(1), where both agree to ~2 ulp).
RPX_DEV cplx operator/(cplx a, cplx b) {
    double inv = rcp(b.re * b.re + b.im * b.im);
    return cx((a.re * b.re + a.im * b.im) * inv, (a.im * b.re - a.re * b.im) * inv);
}
RPX_DEV cplx crcp(cplx b) {
    double inv = rcp(b.re * b.re + b.im * b.im);
    return cx(b.re * inv, -b.im * inv);
}
RPX_DEV double cabs2(cplx a) { return a.re * a.re + a.im * a.im; }  // cabs(a)**2
// C99 csqrt (Annex G branch cut along the negative real axis, sign of the imaginary
// part follows the sign of z.im including -0.0) -- the TIR branch of the Fresnel
// materials depends on csqrt(negative + 0i) = +i*sqrt(|x|)  (cmaterials.pyx:805)
RPX_DEV cplx csqrt_(cplx z) {
    double x = z.re, y = z.im;
    if (y == 0.0) {
        double q = sqrt_(fabs(x));
        if (x < 0.0) return cx(0.0, copysign(q, y));
        return cx(q, copysign(0.0, y));
    }
    if (x == 0.0) {
        double r = sqrt_(0.5 * fabs(y));
        return cx(r, copysign(r, y));
    }
    double d = sqrt_(x * x + y * y);  // |z| is O(1) here: no need for hypot's rescaling
    // t = sqrt((|z| + |x|) / 2) is the larger of the two parts; the other is y / (2 t)
    double t = sqrt_(0.5 * (d + fabs(x)));
    double u = 0.5 * (y * rcp(t));
    if (x > 0.0) return cx(t, u);
    return cx(fabs(u), copysign(t, y));
}
// sin and cos of a large-magnitude phase: quadrant reduction with an exact-product FMA step
// (pi/2 = hi + lo; |error| < 3e-16 rad for |n| < 2^30), then the fdlibm minimax kernels on
// |r| <= pi/4 with the coefficients in the constant bank (no UMOV pairs, no Payne-Hanek path).
static __constant__ double c_sin[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
                                2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10};
static __constant__ double c_cos[6] = {4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                                -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11};
RPX_DEV void sincos_phase(double w, double* sn, double* cs) {
    const double n = rint(w * 0.63661977236758138);  // 2 / pi
    double r = fma(-n, 1.5707963267948966, w);
    r = fma(-n, 6.123233995736766e-17, r);
    const double z = r * r;
    double ps = c_sin[5];
    double pc = c_cos[5];
#pragma unroll
    for (int k = 4; k >= 0; k--) {
        ps = fma(ps, z, c_sin[k]);
        pc = fma(pc, z, c_cos[k]);
    }
    const double s = fma(r * z, ps, r);               // r + r^3 (S1 + ...)
    const double c = fma(z * z, pc, fma(z, -0.5, 1.0));  // 1 - z/2 + z^2 (C1 + ...)
    const int q = (int)(long long)n;                  // |n| < 2^31 for any physical phase
    const double a = (q & 1) ? c : s;
    const double b = (q & 1) ? s : c;
    *sn = (q & 2) ? -a : a;
    *cs = ((q + 1) & 2) ? -b : b;
}


RPX_DEV cplx cexp_(cplx z) {
    double s, c;
    sincos_phase(z.im, &s, &c);
    if (z.re == 0.0) return cx(c, s);  // lossless media: the film phase is purely imaginary
    double e = exp(z.re);
    return cx(e * c, e * s);
}

}  // namespace rpx
