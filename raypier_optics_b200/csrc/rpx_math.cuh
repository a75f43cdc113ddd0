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
RPX_DEV double mag(vec3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
RPX_DEV vec3 cross(vec3 a, vec3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// norm_ of the reference divides each component by the magnitude (ctracer.pyx:251-256);
// three IEEE divisions, kept (not one reciprocal) so a unit vector stays a unit vector to
// the same ulp as the reference.
RPX_DEV vec3 norm(vec3 a) {
    double m = sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
    return v3(a.x / m, a.y / m, a.z / m);
}
RPX_DEV double sep(vec3 p1, vec3 p2) {
    double a = p2.x - p1.x, b = p2.y - p1.y, c = p2.z - p1.z;
    return sqrt((a * a) + (b * b) + (c * c));
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
// Smith's algorithm, the same scheme libgcc's __divdc3 uses for in-range operands
RPX_DEV cplx operator/(cplx a, cplx b) {
    cplx r;
    if (fabs(b.re) < fabs(b.im)) {
        double ratio = b.re / b.im;
        double denom = (b.re * ratio) + b.im;
        r.re = ((a.re * ratio) + a.im) / denom;
        r.im = ((a.im * ratio) - a.re) / denom;
    } else {
        double ratio = b.im / b.re;
        double denom = (b.im * ratio) + b.re;
        r.re = ((a.im * ratio) + a.re) / denom;
        r.im = (a.im - (a.re * ratio)) / denom;
    }
    return r;
}
RPX_DEV double cabs2(cplx a) {  // cabs(a)**2 as the reference writes it
    double h = hypot(a.re, a.im);
    return h * h;
}
// C99 csqrt (Annex G branch cut along the negative real axis, sign of the imaginary
// part follows the sign of z.im including -0.0) -- the TIR branch of the Fresnel
// materials depends on csqrt(negative + 0i) = +i*sqrt(|x|)  (cmaterials.pyx:805)
RPX_DEV cplx csqrt_(cplx z) {
    double x = z.re, y = z.im;
    if (y == 0.0) {
        if (x < 0.0) return cx(0.0, copysign(sqrt(-x), y));
        return cx(fabs(sqrt(x)), copysign(0.0, y));
    }
    if (x == 0.0) {
        double r = sqrt(0.5 * fabs(y));
        return cx(r, copysign(r, y));
    }
    double d = hypot(x, y);
    double r, s;
    if (x > 0.0) {
        r = sqrt(0.5 * (d + x));
        s = 0.5 * (y / r);
    } else {
        s = sqrt(0.5 * (d - x));
        r = fabs(0.5 * (y / s));
    }
    return cx(r, copysign(s, y));
}
RPX_DEV cplx cexp_(cplx z) {
    double e = exp(z.re);
    double s, c;
    sincos(z.im, &s, &c);
    return cx(e * c, e * s);
}

}  // namespace rpx
