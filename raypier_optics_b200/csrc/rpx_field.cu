// rpx_field.cu -- E-field at a detector as a sum of general astigmatic Gaussian modes
// (SURVEY 8f.1): raypier/core/cfields.pyx sum_gaussian_modes (:51-115), calc_mode_U (:118-153),
// evaluate_modes / evaluate_one_mode (:156-228) and the gausslet front end of
// raypier/core/fields.py (evaluate_neighbours_gc :114-137, EFieldSummation :206-249).
//
// B200 design.  The work is N_ray x N_pt independent complex exponentials -- the one compute-dense
// (FP64-pipe bound) kernel of the product, no dense contraction, so no tensor cores.
//   k_field_prepare  one thread per ray: parabasal rays -> (x, y, dx, dy) -> closed-form
//                    least-squares mode (A, B, C) -> a 33-double "mode record" holding every
//                    per-ray constant of the inner loop (basis, k, phase, A, B, C, detG0, A + C and
//                    the polarisation vector V = E1*W*E + E2*W*H with W = inv_root_area*sqrt(i)).
//   k_field_sum      grid = point tiles x ray slices; a CTA stages 32 records at a time in shared
//                    memory (broadcast LDS), every thread owns 2 points (two independent
//                    dependency chains), accumulates 6 doubles per point in registers and adds them
//                    to the output with fp64 atomics.
// Algebra (all exact identities of the reference's formulae):
//   (1 + zA)(1 + zC) - (zB)^2 = 1 + z(A + C) + z^2 detG0 = M, so denom = 2M and the csqrt argument
//   is i*M: ONE |M|^2, two rsqrt, no division;  AA, B/denom, CC share the factor 1/(2M);
//   U /= csqrt(iM) is U * conj(s) / |M| with s = csqrt(iM).
// Rounding: the optical phase is ~1e6 rad, so one ulp of the path term is ~1e-10 rad.  The terms
// that carry that magnitude (z = pt . direction, the sum z + AA x^2 + ..., k * arg, + phase) are
// evaluated with the reference's own association and WITHOUT FMA contraction (__dmul_rn /
// __dadd_rn); everything small is free to contract.  The argument of sin/cos is reduced mod 2 pi
// with a two-constant FMA Cody-Waite step (|error| < 4e-16 rad for |x| < 2^40) so the library
// sincos stays on its fast path instead of Payne-Hanek.
#include <cstdio>
#include <new>
#include <vector>

#include "rpx_internal.h"

using namespace rpx;

namespace {

enum {
    M_EX = 0, M_EY, M_EZ, M_HX, M_HY, M_HZ, M_DX, M_DY, M_DZ, M_OX, M_OY, M_OZ,
    M_KR, M_KI, M_PH, M_K, M_NRE,
    M_AR, M_AI, M_BR, M_BI, M_CR, M_CI, M_GR, M_GI, M_SR, M_SI,
    M_V0R, M_V0I, M_V1R, M_V1I, M_V2R, M_V2I,
    M_NF
};
static_assert(M_NF == 33, "mode record");

#ifndef FIELD_LIBM
#define FIELD_LIBM 0  // 1: library sincos / exp instead of the constant-bank versions below
#endif
#define FIELD_THREADS 128
#define FIELD_PTS 2    // points per thread
#define FIELD_STAGE 32 // mode records per shared-memory stage

// ------------------------------------------------------------------ prepare
// evaluate_one_mode, cfields.pyx:156-214
__device__ void evaluate_one_mode(const double* x, const double* y, const double* dx, const double* dy,
                                  double blending, cplx* out) {
    double a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0, b0 = 0, b1 = 0, b2 = 0;
#pragma unroll
    for (int i = 0; i < RPX_NPARA; i++) {
        double xi2 = x[i] * x[i], yi2 = y[i] * y[i];
        a00 += xi2 * xi2;
        a01 += 2 * xi2 * x[i] * y[i];
        a02 += xi2 * yi2;
        a11 += 4 * xi2 * yi2;
        a12 += 2 * x[i] * y[i] * yi2;
        a22 += yi2 * yi2;
        b0 += xi2;
        b1 += 2 * x[i] * y[i];
        b2 += yi2;
    }
    double den = -a00 * a11 * a22 + a00 * a12 * a12 + a01 * a01 * a22 - 2 * a01 * a02 * a12 + a02 * a02 * a11;
    double im0 = blending * (-b0 * (a11 * a22 - a12 * a12) + b1 * (a01 * a22 - a02 * a12) - b2 * (a01 * a12 - a02 * a11)) / den;
    double im1 = -blending * (b0 * (a01 * a22 - a02 * a12) - b1 * (a00 * a22 - a02 * a02) + b2 * (a00 * a12 - a01 * a02)) / den;
    double im2 = blending * (-b0 * (a01 * a12 - a02 * a11) + b1 * (a00 * a12 - a01 * a02) - b2 * (a00 * a11 - a01 * a01)) / den;
    a00 = a01 = a02 = a11 = a12 = a22 = b0 = b1 = b2 = 0;
#pragma unroll
    for (int i = 0; i < RPX_NPARA; i++) {
        double xi2 = x[i] * x[i], yi2 = y[i] * y[i];
        a00 += xi2;
        a01 += x[i] * y[i];
        a11 += xi2 + yi2;
        a22 += yi2;
        b0 += dx[i] * x[i];
        b1 += dx[i] * y[i] + dy[i] * x[i];
        b2 += dy[i] * y[i];
    }
    a12 = a01 * a01;
    den = a00 * a12 - a00 * a11 * a22 + a12 * a22;
    out[0] = cx((-a12 * b2 + a01 * a22 * b1 + b0 * (a12 - a11 * a22)) / den, im0);
    out[1] = cx(-(a00 * a01 * b2 - a00 * a22 * b1 + a01 * a22 * b0) / den, im1);
    out[2] = cx((a00 * a01 * b1 - a12 * b0 - b2 * (a00 * a11 - a12)) / den, im2);
}

// The per-ray part of sum_gaussian_modes (cfields.pyx:74-97): everything the inner loop needs of ray i
// (already loaded: o, d, e) and its mode M, written as record `io`.
__device__ void write_mode_record(const Soa& in, unsigned long long i, unsigned long long io, vec3 o, vec3 d, vec3 e,
                                  double apath, const cplx* M, const double* wavelengths, int n_wl, double* rec,
                                  double* modes_out) {
    const unsigned long long cap = in.cap;
    for (int k = 0; k < 3; k++) {
        modes_out[io * 6 + 2 * k] = M[k].re;
        modes_out[io * 6 + 2 * k + 1] = M[k].im;
    }
    // IEEE sqrt / division here: per-ray work, accuracy over speed
    const double einv = 1.0 / sqrt(mag_sq(e));
    const vec3 E = v3(e.x * einv, e.y * einv, e.z * einv);  // norm_(ray.E_vector)
    vec3 H = cross(d, E);                                   // norm_(cross_(ray.direction, E))
    const double hinv = 1.0 / sqrt(mag_sq(H));
    H = v3(H.x * hinv, H.y * hinv, H.z * hinv);
    const uint32_t wl = in.u[U_WL * cap + i];
    const double k = (wl < (uint32_t)n_wl) ? (2000.0 * M_PI) / wavelengths[wl] : __longlong_as_double(0x7ff8000000000000LL);
    const double n_re = in.f[F_NR * cap + i], n_im = in.f[F_NI * cap + i];
    // phase = ray.phase + accumulated_path*k  (the - c*k/n term carries time_ps: applied per evaluate)
    const double ph0 = __dadd_rn(in.f[F_PHASE * cap + i], __dmul_rn(apath, k));
    const double kr = n_re * k, ki = n_im * k;
    const double invk = 2.0 / kr;
    const double inv_root_area = sqrt(sqrt(M[0].im * M[2].im - M[1].im * M[1].im) * (2.0 / M_PI));
    const cplx A = cx(M[0].re, M[0].im * invk), B = cx(M[1].re, M[1].im * invk), C = cx(M[2].re, M[2].im * invk);
    const cplx G = A * C - B * B;  // detG0
    const cplx S = A + C;
    const double rt = sqrt(0.5);   // rootI = csqrt(i) = (sqrt(1/2), sqrt(1/2))
    const cplx W = cx(inv_root_area * rt, inv_root_area * rt);
    const cplx W1 = cx(in.f[F_E1R * cap + i], in.f[F_E1I * cap + i]) * W;
    const cplx W2 = cx(in.f[F_E2R * cap + i], in.f[F_E2I * cap + i]) * W;
    double* r = rec + io * M_NF;
    r[M_EX] = E.x; r[M_EY] = E.y; r[M_EZ] = E.z;
    r[M_HX] = H.x; r[M_HY] = H.y; r[M_HZ] = H.z;
    r[M_DX] = d.x; r[M_DY] = d.y; r[M_DZ] = d.z;
    r[M_OX] = o.x; r[M_OY] = o.y; r[M_OZ] = o.z;
    r[M_KR] = kr; r[M_KI] = ki; r[M_PH] = ph0; r[M_K] = k; r[M_NRE] = n_re;
    r[M_AR] = A.re; r[M_AI] = A.im; r[M_BR] = B.re; r[M_BI] = B.im; r[M_CR] = C.re; r[M_CI] = C.im;
    r[M_GR] = G.re; r[M_GI] = G.im; r[M_SR] = S.re; r[M_SI] = S.im;
    r[M_V0R] = W1.re * E.x + W2.re * H.x; r[M_V0I] = W1.im * E.x + W2.im * H.x;
    r[M_V1R] = W1.re * E.y + W2.re * H.y; r[M_V1I] = W1.im * E.y + W2.im * H.y;
    r[M_V2R] = W1.re * E.z + W2.re * H.z; r[M_V2I] = W1.im * E.z + W2.im * H.z;
}

// Mode records of a whole collection; for gausslets the mode is fitted first (fields.py:114-137 +
// cfields.pyx:217-228).
template <bool FROM_PARA>
__global__ void k_field_prepare(Soa in, const double* modes_in, const double* wavelengths, int n_wl, double blending,
                                double* rec, double* modes_out) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= in.n) return;
    const unsigned long long cap = in.cap;
    const vec3 o = v3(in.f[F_OX * cap + i], in.f[F_OY * cap + i], in.f[F_OZ * cap + i]);
    const vec3 d = v3(in.f[F_DX * cap + i], in.f[F_DY * cap + i], in.f[F_DZ * cap + i]);
    const vec3 e = v3(in.f[F_EX * cap + i], in.f[F_EY * cap + i], in.f[F_EZ * cap + i]);
    cplx M[3];
    if (FROM_PARA) {
        const vec3 H0 = cross(e, d);  // numpy.cross(E, direction), fields.py:120
        double x[RPX_NPARA], y[RPX_NPARA], dx[RPX_NPARA], dy[RPX_NPARA];
#pragma unroll
        for (int j = 0; j < RPX_NPARA; j++) {
            const double* pp = in.p + (unsigned long long)(j * NPF) * cap + i;
            vec3 off = v3(pp[(P_OX + 0) * cap], pp[(P_OX + 1) * cap], pp[(P_OX + 2) * cap]) - o;
            vec3 nd = v3(pp[(P_DX + 0) * cap], pp[(P_DX + 1) * cap], pp[(P_DX + 2) * cap]);
            x[j] = dot(off, e);
            y[j] = dot(off, H0);
            double dz = dot(nd, d);
            dx[j] = dot(nd, e) / dz;
            dy[j] = dot(nd, H0) / dz;
        }
        evaluate_one_mode(x, y, dx, dy, blending, M);
    } else {
        for (int k = 0; k < 3; k++) M[k] = cx(modes_in[i * 6 + 2 * k], modes_in[i * 6 + 2 * k + 1]);
    }
    write_mode_record(in, i, i, o, d, e, in.f[F_APATH * cap + i], M, wavelengths, n_wl, rec, modes_out);
}

// ---- plain rays with neighbour lists: eval_Efield_from_rays (core/fields.py:206-229) -------------
// project_to_sphere (fields.py:50-77), in place on the SoA collection: origin += alpha * direction,
// accumulated_path += alpha * n.real with alpha the most negative root of the ray-line / sphere
// quadratic; rays that miss the sphere are left alone and flagged 0 in `selector`.
__global__ void k_project_to_sphere(Soa rays, double cx_, double cy_, double cz_, double radius, unsigned char* selector,
                                    unsigned long long* n_selected) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rays.n) return;
    const unsigned long long cap = rays.cap;
    const double ox = rays.f[F_OX * cap + i] - cx_, oy = rays.f[F_OY * cap + i] - cy_, oz = rays.f[F_OZ * cap + i] - cz_;
    const double dx = rays.f[F_DX * cap + i], dy = rays.f[F_DY * cap + i], dz = rays.f[F_DZ * cap + i];
    const double c = ((ox * ox + oy * oy) + oz * oz) - radius * radius;
    const double b = 2 * ((dx * ox + dy * oy) + dz * oz);
    double disc = b * b - 4 * c;
    const bool ok = disc >= 0.0;
    selector[i] = ok ? 1 : 0;
    if (!ok) return;
    disc = sqrt(disc);
    const double alpha = fmin((-b + disc) / 2, (-b - disc) / 2);
    rays.f[F_OX * cap + i] += alpha * dx;
    rays.f[F_OY * cap + i] += alpha * dy;
    rays.f[F_OZ * cap + i] += alpha * dz;
    rays.f[F_APATH * cap + i] += alpha * rays.f[F_NR * cap + i];
    atomicAdd(n_selected, 1ull);
}

// evaluate_neighbours (fields.py:80-111) -> evaluate_modes (cfields.pyx:217-228) -> mode record, one
// thread per ray; pos[i] = index among the rays that have all six neighbours, or -1 (dropped, :97).
// xy (may be NULL) receives x, y, dx, dy of the kept rays as four n_kept x 6 blocks.
__global__ void k_field_prepare_nb(Soa in, const int* nb, const long long* pos, unsigned long long n_kept,
                                   const double* wavelengths, int n_wl, double blending, double* rec, double* modes_out,
                                   double* xy) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= in.n) return;
    const long long io = pos[i];
    if (io < 0) return;
    const unsigned long long cap = in.cap;
    const vec3 o = v3(in.f[F_OX * cap + i], in.f[F_OY * cap + i], in.f[F_OZ * cap + i]);
    const vec3 d = v3(in.f[F_DX * cap + i], in.f[F_DY * cap + i], in.f[F_DZ * cap + i]);
    const vec3 e = v3(in.f[F_EX * cap + i], in.f[F_EY * cap + i], in.f[F_EZ * cap + i]);
    const vec3 H0 = cross(e, d);  // numpy.cross(E, direction), fields.py:101
    double x[RPX_NPARA], y[RPX_NPARA], dx[RPX_NPARA], dy[RPX_NPARA];
#pragma unroll
    for (int j = 0; j < RPX_NPARA; j++) {
        const unsigned long long q = (unsigned long long)nb[i * RPX_NPARA + j];
        const vec3 off = v3(in.f[F_OX * cap + q], in.f[F_OY * cap + q], in.f[F_OZ * cap + q]) - o;
        const vec3 nd = v3(in.f[F_DX * cap + q], in.f[F_DY * cap + q], in.f[F_DZ * cap + q]);
        const double dz = dot(nd, d);
        const double alpha = -dot(off, d) / dz;
        const vec3 proj = off + nd * alpha;
        x[j] = dot(proj, e);
        y[j] = dot(proj, H0);
        dx[j] = dot(nd, e) / dz;
        dy[j] = dot(nd, H0) / dz;
        if (xy) {
            const unsigned long long blk = n_kept * RPX_NPARA, at = (unsigned long long)io * RPX_NPARA + j;
            xy[at] = x[j];
            xy[blk + at] = y[j];
            xy[2 * blk + at] = dx[j];
            xy[3 * blk + at] = dy[j];
        }
    }
    cplx M[3];
    evaluate_one_mode(x, y, dx, dy, blending, M);
    write_mode_record(in, i, (unsigned long long)io, o, d, e, in.f[F_APATH * cap + i], M, wavelengths, n_wl, rec, modes_out);
}

// cfields.evaluate_modes (cfields.pyx:217-228) on explicit n x 6 arrays
__global__ void k_evaluate_modes(const double* x, const double* y, const double* dx, const double* dy, unsigned long long n,
                                 double blending, double* modes_out) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double xs[RPX_NPARA], ys[RPX_NPARA], dxs[RPX_NPARA], dys[RPX_NPARA];
#pragma unroll
    for (int j = 0; j < RPX_NPARA; j++) {
        xs[j] = x[i * RPX_NPARA + j];
        ys[j] = y[i * RPX_NPARA + j];
        dxs[j] = dx[i * RPX_NPARA + j];
        dys[j] = dy[i * RPX_NPARA + j];
    }
    cplx M[3];
    evaluate_one_mode(xs, ys, dxs, dys, blending, M);
    for (int k = 0; k < 3; k++) {
        modes_out[i * 6 + 2 * k] = M[k].re;
        modes_out[i * 6 + 2 * k + 1] = M[k].im;
    }
}

// ------------------------------------------------------------------ the N_ray x N_pt sum
// ---- sin / cos / exp of the inner loop ----------------------------------------------------------
// The library sincos + exp are 30 % of k_field_sum (measured by stubbing them out), a good part of
// it UMOV traffic that materialises their polynomial coefficients.  These versions keep the
// coefficients in the constant bank (DFMA takes a c[][] operand directly) and fold the argument
// reduction of the optical phase (|w| up to ~1e7 rad) into the quadrant reduction:
//   n = rint(w * 2/pi);  r = w - n*pi/2 with pi/2 = hi + lo and exact products in the FMAs
//   (|error| < 3e-16 rad for |n| < 2^30);  sin / cos on |r| <= pi/4 by the fdlibm minimax kernels
//   (__kernel_sin / __kernel_cos coefficients, < 1 ulp);  exp(x) = 2^n * P13(x - n ln2), Taylor to
//   degree 13 on |r| <= ln2 / 2 (truncation 4e-18), n clamped so underflow -> 0 and overflow -> inf.
__constant__ double c_exp[12] = {1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320,
                                 1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0};

__device__ __forceinline__ double exp_fast(double x) {
    x = fmin(fmax(x, -746.0), 710.0);
    const double n = rint(x * 1.4426950408889634);  // log2(e)
    double r = fma(-n, 6.93147180369123816490e-01, x);  // ln2 hi (fdlibm split: low 32 bits zero)
    r = fma(-n, 1.90821492927058770002e-10, r);     // ln2 lo
    double p = c_exp[11];
#pragma unroll
    for (int k = 10; k >= 0; k--) p = fma(p, r, c_exp[k]);
    p = fma(r * r, p, r + 1.0);                        // 1 + r + r^2 (1/2 + r/6 + ...)
    // 2^n in two factors so that the result may be subnormal or overflow to inf like exp() does
    const int ni = (int)n, h = ni / 2;
    const double f1 = __hiloint2double((1023 + h) << 20, 0), f2 = __hiloint2double((1023 + ni - h) << 20, 0);
    return p * f1 * f2;
}

// One (ray, point) pair: calc_mode_U (cfields.pyx:118-153) + the accumulation of :104-110.
__device__ __forceinline__ void field_pair(const double* __restrict__ R, double Px, double Py, double Pz, double* acc) {
    const double px = Px - R[M_OX], py = Py - R[M_OY], pz = Pz - R[M_OZ];
    const double x = px * R[M_EX] + py * R[M_EY] + pz * R[M_EZ];
    const double y = px * R[M_HX] + py * R[M_HY] + pz * R[M_HZ];
    // dotprod_ association and roundings of the reference: the k*z term is ~1e5..1e6 rad
    const double z = __dadd_rn(__dadd_rn(__dmul_rn(px, R[M_DX]), __dmul_rn(py, R[M_DY])), __dmul_rn(pz, R[M_DZ]));
    const double z2 = z * z;
    const double Gr = R[M_GR], Gi = R[M_GI];
    // M = 1 + z (A + C) + z^2 detG0 = denom / 2 = (1 + zA)(1 + zC) - (zB)^2
    const double Mr = fma(z2, Gr, fma(z, R[M_SR], 1.0));
    const double Mi = fma(z2, Gi, z * R[M_SI]);
    const double m2 = Mr * Mr + Mi * Mi;
    const double rinv = rsqrt_(m2);  // 1 / |M|
    const double inv_m2 = rinv * rinv;
    const double hr = 0.5 * Mr * inv_m2, hi = -0.5 * Mi * inv_m2;  // 1 / denom
    // AA = (A + z detG0) / denom, CC likewise, BB = B / denom
    const double Anr = fma(z, Gr, R[M_AR]), Ani = fma(z, Gi, R[M_AI]);
    const double Cnr = fma(z, Gr, R[M_CR]), Cni = fma(z, Gi, R[M_CI]);
    const double xx = x * x, yy = y * y, xy2 = 2.0 * x * y;
    const double t1r = (Anr * hr - Ani * hi) * xx, t1i = (Anr * hi + Ani * hr) * xx;
    const double t2r = (R[M_BR] * hr - R[M_BI] * hi) * xy2, t2i = (R[M_BR] * hi + R[M_BI] * hr) * xy2;
    const double t3r = (Cnr * hr - Cni * hi) * yy, t3i = (Cnr * hi + Cni * hr) * yy;
    // arg = ((z + AA x^2) + B 2xy / denom) + CC y^2 in the reference's order (:141)
    const double ar = __dadd_rn(__dadd_rn(__dadd_rn(z, t1r), t2r), t3r);
    const double ai = t1i + t2i + t3i;
    // w = phase + k * arg; U = cexp(i w) = exp(-w.im) (cos w.re, sin w.re)
    const double kr = R[M_KR], ki = R[M_KI];
    const double wr = __dadd_rn(R[M_PH], __dsub_rn(__dmul_rn(kr, ar), __dmul_rn(ki, ai)));
    const double wi = fma(kr, ai, ki * ar);
    double sn, cs;
#if FIELD_LIBM
    const double nrot = rint(wr * 0.15915494309189535);  // 1 / (2 pi)
    double red = fma(-nrot, 6.283185307179586, wr);      // 2 pi = hi + lo, exact products in the FMAs
    red = fma(-nrot, 2.4492935982947064e-16, red);
    sincos(red, &sn, &cs);
#else
    rpx::sincos_phase(wr, &sn, &cs);
#endif
    // U /= csqrt(i M):  s = csqrt(w), w = (-Mi, Mr), |w| = |M|;  1/s = conj(s) / |M|
    const double rabs = m2 * rinv;
    const double a = -Mi, b = Mr;
    const double q = 0.5 * (rabs + fabs(a));
    const double ti = rsqrt_(q);
    const double t = q * ti, u = 0.5 * b * ti;
    const double sr = (a >= 0.0) ? t : fabs(u);
    const double si = (a >= 0.0) ? u : copysign(t, b);
#if FIELD_LIBM
    const double g = exp(-wi) * rinv;
#else
    const double g = exp_fast(-wi) * rinv;
#endif
    // U' = g (cs + i sn) (sr - i si)
    const double ur = g * (cs * sr + sn * si), ui = g * (sn * sr - cs * si);
    // out[c] += U' * V[c]   (V = E1*W*E + E2*W*H: the E1*U*E.x + E2*U*H.x of :108-110, factored)
    acc[0] = fma(ur, R[M_V0R], fma(-ui, R[M_V0I], acc[0]));
    acc[1] = fma(ur, R[M_V0I], fma(ui, R[M_V0R], acc[1]));
    acc[2] = fma(ur, R[M_V1R], fma(-ui, R[M_V1I], acc[2]));
    acc[3] = fma(ur, R[M_V1I], fma(ui, R[M_V1R], acc[3]));
    acc[4] = fma(ur, R[M_V2R], fma(-ui, R[M_V2I], acc[4]));
    acc[5] = fma(ur, R[M_V2I], fma(ui, R[M_V2R], acc[5]));
}

__global__ void __launch_bounds__(FIELD_THREADS)
k_field_sum(const double* __restrict__ rec, unsigned long long n_rays, const double* __restrict__ points,
            unsigned long long npt, double ctime, double* out, unsigned long long rays_per_slice) {
    __shared__ double sm[FIELD_STAGE * M_NF];
    const unsigned long long p0 = (unsigned long long)blockIdx.x * (FIELD_THREADS * FIELD_PTS) + threadIdx.x;
    double P[FIELD_PTS][3], acc[FIELD_PTS][6];
#pragma unroll
    for (int q = 0; q < FIELD_PTS; q++) {
        const unsigned long long p = p0 + (unsigned long long)q * FIELD_THREADS;
        const bool ok = p < npt;
        P[q][0] = ok ? points[3 * p] : 0.0;
        P[q][1] = ok ? points[3 * p + 1] : 0.0;
        P[q][2] = ok ? points[3 * p + 2] : 0.0;
#pragma unroll
        for (int c = 0; c < 6; c++) acc[q][c] = 0.0;
    }
    const unsigned long long r_begin = (unsigned long long)blockIdx.y * rays_per_slice;
    const unsigned long long r_end = min(r_begin + rays_per_slice, n_rays);
    for (unsigned long long r0 = r_begin; r0 < r_end; r0 += FIELD_STAGE) {
        const int cnt = (int)min((unsigned long long)FIELD_STAGE, r_end - r0);
        __syncthreads();
        for (int w = threadIdx.x; w < cnt * M_NF; w += FIELD_THREADS) sm[w] = rec[r0 * M_NF + w];
        __syncthreads();
        if (ctime != 0.0 && threadIdx.x < cnt) {  // phase -= (c*k)/n.real, cfields.pyx:80
            double* R = sm + threadIdx.x * M_NF;
            R[M_PH] = __dsub_rn(R[M_PH], __ddiv_rn(__dmul_rn(ctime, R[M_K]), R[M_NRE]));
        }
        if (ctime != 0.0) __syncthreads();
        for (int j = 0; j < cnt; j++) {
            const double* R = sm + j * M_NF;
#pragma unroll
            for (int q = 0; q < FIELD_PTS; q++) field_pair(R, P[q][0], P[q][1], P[q][2], acc[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < FIELD_PTS; q++) {
        const unsigned long long p = p0 + (unsigned long long)q * FIELD_THREADS;
        if (p < npt) {
#pragma unroll
            for (int c = 0; c < 6; c++) atomicAdd(out + 6 * p + c, acc[q][c]);
        }
    }
}

}  // namespace

// ------------------------------------------------------------------ C ABI
struct rpx_field {
    double* rec;    // n x 33 mode records
    double* modes;  // n x 3 complex (A, B, C) as fitted / as given
    uint64_t n;
    float last_ms;
};

extern "C" int rpx_field_prepare(rpx_ctx* ctx, const rpx_rays* rays, const double* modes, const double* wavelengths,
                                 int n_wavelengths, double blending, rpx_field** out) {
    if (!ctx || !rays || !wavelengths || !out || n_wavelengths <= 0) return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (!modes && !rays->is_gausslet)
        return fail(ctx, RPX_ERR_INVALID, "plain rays need explicit modes (only gausslets carry parabasal rays to fit)");
    CU(ctx, cudaSetDevice(ctx->device));
    const uint64_t n = rays->soa.n;
    rpx_field* f = new (std::nothrow) rpx_field();
    if (!f) return fail(ctx, RPX_ERR_NOMEM, "out of host memory");
    f->n = n;
    f->rec = nullptr;
    f->modes = nullptr;
    f->last_ms = 0.f;
    const size_t nn = n ? n : 1;
    double* d_wl = nullptr;
    double* d_modes_in = nullptr;
    cudaError_t e;
    auto bail = [&](cudaError_t err, const char* what) {
        if (f->rec) cudaFreeAsync(f->rec, ctx->stream);
        if (f->modes) cudaFreeAsync(f->modes, ctx->stream);
        if (d_wl) cudaFreeAsync(d_wl, ctx->stream);
        if (d_modes_in) cudaFreeAsync(d_modes_in, ctx->stream);
        delete f;
        return fail(ctx, err == cudaErrorMemoryAllocation ? RPX_ERR_NOMEM : RPX_ERR_CUDA, "%s failed: %s", what,
                    cudaGetErrorString(err));
    };
    if ((e = cudaMallocAsync((void**)&f->rec, nn * M_NF * sizeof(double), ctx->stream)) != cudaSuccess ||
        (e = cudaMallocAsync((void**)&f->modes, nn * 6 * sizeof(double), ctx->stream)) != cudaSuccess ||
        (e = cudaMallocAsync((void**)&d_wl, sizeof(double) * (size_t)n_wavelengths, ctx->stream)) != cudaSuccess)
        return bail(e, "cudaMallocAsync");
    if ((e = cudaMemcpyAsync(d_wl, wavelengths, sizeof(double) * (size_t)n_wavelengths, cudaMemcpyHostToDevice,
                             ctx->stream)) != cudaSuccess)
        return bail(e, "cudaMemcpyAsync(wavelengths)");
    if (modes && n) {
        if ((e = cudaMallocAsync((void**)&d_modes_in, n * 6 * sizeof(double), ctx->stream)) != cudaSuccess ||
            (e = cudaMemcpyAsync(d_modes_in, modes, n * 6 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) !=
                cudaSuccess)
            return bail(e, "modes upload");
    }
    if (n) {
        const unsigned T = 128, G = (unsigned)((n + T - 1) / T);
        if (modes)
            k_field_prepare<false><<<G, T, 0, ctx->stream>>>(rays->soa, d_modes_in, d_wl, n_wavelengths, blending, f->rec,
                                                             f->modes);
        else
            k_field_prepare<true><<<G, T, 0, ctx->stream>>>(rays->soa, nullptr, d_wl, n_wavelengths, blending, f->rec,
                                                            f->modes);
        if ((e = cudaGetLastError()) != cudaSuccess) return bail(e, "k_field_prepare launch");
    }
    cudaFreeAsync(d_wl, ctx->stream);
    d_wl = nullptr;
    if (d_modes_in) cudaFreeAsync(d_modes_in, ctx->stream);
    d_modes_in = nullptr;
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return bail(e, "k_field_prepare");
    *out = f;
    return RPX_OK;
}

extern "C" int rpx_rays_project_to_sphere(rpx_ctx* ctx, rpx_rays* rays, const double* centre, double radius,
                                          uint8_t* selector, uint64_t* n_selected) {
    if (!ctx || !rays || !centre) return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    if (rays->is_gausslet) return fail(ctx, RPX_ERR_INVALID, "project_to_sphere takes plain rays (ray_t)");
    CU(ctx, cudaSetDevice(ctx->device));
    const uint64_t n = rays->soa.n;
    if (n_selected) *n_selected = 0;
    if (!n) return RPX_OK;
    unsigned char* d_sel = nullptr;
    unsigned long long* d_cnt = nullptr;
    CU(ctx, cudaMallocAsync((void**)&d_sel, n + sizeof(unsigned long long) + 8, ctx->stream));
    d_cnt = reinterpret_cast<unsigned long long*>(d_sel + ((n + 7) / 8) * 8);
    cudaError_t e = cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), ctx->stream);
    if (e == cudaSuccess) {
        k_project_to_sphere<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(rays->soa, centre[0], centre[1], centre[2],
                                                                                  radius, d_sel, d_cnt);
        e = cudaGetLastError();
    }
    unsigned long long cnt = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && selector) e = cudaMemcpyAsync(selector, d_sel, n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFreeAsync(d_sel, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, RPX_ERR_CUDA, "project_to_sphere failed: %s", cudaGetErrorString(e));
    if (n_selected) *n_selected = cnt;
    return RPX_OK;
}

extern "C" int rpx_field_prepare_neighbours(rpx_ctx* ctx, const rpx_rays* rays, const int32_t* neighbours, int row_size,
                                            const double* wavelengths, int n_wavelengths, double blending,
                                            double* xy_out, rpx_field** out) {
    if (!ctx || !rays || !wavelengths || !out || n_wavelengths <= 0 || (!neighbours && rays->soa.n))
        return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (rays->is_gausslet) return fail(ctx, RPX_ERR_INVALID, "neighbour lists belong to plain rays (ray_t)");
    if (row_size != RPX_NPARA)
        return fail(ctx, RPX_ERR_UNSUPPORTED, "neighbour lists of %d entries per ray (only 6 are supported)", row_size);
    CU(ctx, cudaSetDevice(ctx->device));
    const uint64_t n = rays->soa.n;
    // mask = (neighbours_idx >= 0).all(axis=1) and the position of every kept ray (fields.py:97)
    std::vector<long long> pos(n ? n : 1);
    uint64_t kept = 0;
    for (uint64_t i = 0; i < n; i++) {
        bool all = true;
        for (int j = 0; j < RPX_NPARA; j++) {
            const int32_t q = neighbours[i * RPX_NPARA + j];
            if (q >= 0 && (uint64_t)q >= n)
                return fail(ctx, RPX_ERR_INVALID, "neighbour index %d of ray %llu is out of bounds for %llu rays", q,
                            (unsigned long long)i, (unsigned long long)n);
            all = all && q >= 0;
        }
        pos[i] = all ? (long long)kept++ : -1;
    }
    rpx_field* f = new (std::nothrow) rpx_field();
    if (!f) return fail(ctx, RPX_ERR_NOMEM, "out of host memory");
    f->n = kept;
    f->rec = nullptr;
    f->modes = nullptr;
    f->last_ms = 0.f;
    const size_t nk = kept ? kept : 1, nn = n ? n : 1;
    double* d_wl = nullptr;
    double* d_xy = nullptr;
    int* d_nb = nullptr;
    long long* d_pos = nullptr;
    cudaError_t e;
    auto release = [&]() {
        if (d_wl) cudaFreeAsync(d_wl, ctx->stream);
        if (d_xy) cudaFreeAsync(d_xy, ctx->stream);
        if (d_nb) cudaFreeAsync(d_nb, ctx->stream);
        if (d_pos) cudaFreeAsync(d_pos, ctx->stream);
    };
    auto bail = [&](cudaError_t err, const char* what) {
        if (f->rec) cudaFreeAsync(f->rec, ctx->stream);
        if (f->modes) cudaFreeAsync(f->modes, ctx->stream);
        release();
        delete f;
        return fail(ctx, err == cudaErrorMemoryAllocation ? RPX_ERR_NOMEM : RPX_ERR_CUDA, "%s failed: %s", what,
                    cudaGetErrorString(err));
    };
    if ((e = cudaMallocAsync((void**)&f->rec, nk * M_NF * sizeof(double), ctx->stream)) != cudaSuccess ||
        (e = cudaMallocAsync((void**)&f->modes, nk * 6 * sizeof(double), ctx->stream)) != cudaSuccess ||
        (e = cudaMallocAsync((void**)&d_wl, sizeof(double) * (size_t)n_wavelengths, ctx->stream)) != cudaSuccess ||
        (e = cudaMallocAsync((void**)&d_nb, nn * RPX_NPARA * sizeof(int), ctx->stream)) != cudaSuccess ||
        (e = cudaMallocAsync((void**)&d_pos, nn * sizeof(long long), ctx->stream)) != cudaSuccess ||
        (xy_out && (e = cudaMallocAsync((void**)&d_xy, nk * 4 * RPX_NPARA * sizeof(double), ctx->stream)) != cudaSuccess))
        return bail(e, "cudaMallocAsync");
    if ((e = cudaMemcpyAsync(d_wl, wavelengths, sizeof(double) * (size_t)n_wavelengths, cudaMemcpyHostToDevice,
                             ctx->stream)) != cudaSuccess)
        return bail(e, "cudaMemcpyAsync(wavelengths)");
    if (n) {
        if ((e = cudaMemcpyAsync(d_nb, neighbours, n * RPX_NPARA * sizeof(int), cudaMemcpyHostToDevice, ctx->stream)) !=
                cudaSuccess ||
            (e = cudaMemcpyAsync(d_pos, pos.data(), n * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess)
            return bail(e, "neighbour upload");
        k_field_prepare_nb<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(rays->soa, d_nb, d_pos, kept, d_wl,
                                                                                 n_wavelengths, blending, f->rec, f->modes,
                                                                                 d_xy);
        if ((e = cudaGetLastError()) != cudaSuccess) return bail(e, "k_field_prepare_nb launch");
        if (xy_out && kept &&
            (e = cudaMemcpyAsync(xy_out, d_xy, kept * 4 * RPX_NPARA * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) !=
                cudaSuccess)
            return bail(e, "neighbour coordinates download");
    }
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return bail(e, "k_field_prepare_nb");  // pos / nb stay alive
    release();
    *out = f;
    return RPX_OK;
}

extern "C" int rpx_unit_evaluate_modes(rpx_ctx* ctx, const double* x, const double* y, const double* dx, const double* dy,
                                       uint64_t n, int row_size, double blending, double* modes_out) {
    if (!ctx || ((!x || !y || !dx || !dy || !modes_out) && n)) return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    if (row_size != RPX_NPARA)
        return fail(ctx, RPX_ERR_UNSUPPORTED, "%d neighbours per ray (only 6 are supported)", row_size);
    if (!n) return RPX_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    double* d_in = nullptr;
    const size_t blk = n * RPX_NPARA;
    CU(ctx, cudaMallocAsync((void**)&d_in, (4 * blk + n * 6) * sizeof(double), ctx->stream));
    double* d_modes = d_in + 4 * blk;
    const double* src[4] = {x, y, dx, dy};
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < 4 && e == cudaSuccess; k++)
        e = cudaMemcpyAsync(d_in + k * blk, src[k], blk * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        k_evaluate_modes<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(d_in, d_in + blk, d_in + 2 * blk, d_in + 3 * blk,
                                                                               n, blending, d_modes);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(modes_out, d_modes, n * 6 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFreeAsync(d_in, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, RPX_ERR_CUDA, "evaluate_modes failed: %s", cudaGetErrorString(e));
    return RPX_OK;
}

extern "C" uint64_t rpx_field_count(const rpx_field* f) { return f ? f->n : 0; }

extern "C" int rpx_field_modes(rpx_ctx* ctx, const rpx_field* f, double* modes_out) {
    if (!ctx || !f || (!modes_out && f->n)) return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    CU(ctx, cudaSetDevice(ctx->device));
    if (!f->n) return RPX_OK;
    CU(ctx, cudaMemcpyAsync(modes_out, f->modes, f->n * 6 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return RPX_OK;
}

// d_points / d_out are DEVICE pointers; d_out (npt x 3 complex) is accumulated into.
static int field_launch(rpx_ctx* ctx, rpx_field* f, const double* d_points, uint64_t npt, double time_ps, double* d_out,
                        cudaEvent_t ev0, cudaEvent_t ev1) {
    if (!npt || !f->n) return RPX_OK;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const unsigned long long ptiles = (npt + FIELD_THREADS * FIELD_PTS - 1) / (FIELD_THREADS * FIELD_PTS);
    // enough CTAs for ~8 per SM; a slice is a whole number of shared-memory stages
    unsigned long long slices = ((unsigned long long)sms * 8 + ptiles - 1) / ptiles;
    const unsigned long long stages = (f->n + FIELD_STAGE - 1) / FIELD_STAGE;
    if (slices > stages) slices = stages;
    if (slices < 1) slices = 1;
    if (slices > 65535) slices = 65535;
    const unsigned long long per = ((stages + slices - 1) / slices) * FIELD_STAGE;
    slices = (f->n + per - 1) / per;
    if (ptiles > 0x7fffffffull) return fail(ctx, RPX_ERR_INVALID, "too many evaluation points");
    dim3 grid((unsigned)ptiles, (unsigned)slices);
    if (ev0) CU(ctx, cudaEventRecord(ev0, ctx->stream));
    k_field_sum<<<grid, FIELD_THREADS, 0, ctx->stream>>>(f->rec, f->n, d_points, npt, 0.299792458 * time_ps, d_out, per);
    CU(ctx, cudaGetLastError());
    if (ev1) CU(ctx, cudaEventRecord(ev1, ctx->stream));
    return RPX_OK;
}

extern "C" int rpx_field_evaluate_device(rpx_ctx* ctx, rpx_field* f, const double* d_points, uint64_t npt,
                                         double time_ps, double* d_out) {
    if (!ctx || !f || ((!d_points || !d_out) && npt)) return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    CU(ctx, cudaSetDevice(ctx->device));
    cudaEvent_t ev0, ev1;
    CU(ctx, cudaEventCreate(&ev0));
    CU(ctx, cudaEventCreate(&ev1));
    int rc = field_launch(ctx, f, d_points, npt, time_ps, d_out, ev0, ev1);
    if (rc == RPX_OK) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(ctx, RPX_ERR_CUDA, "k_field_sum failed: %s", cudaGetErrorString(e));
        else if (npt && f->n) cudaEventElapsedTime(&f->last_ms, ev0, ev1);
    }
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    return rc;
}

extern "C" int rpx_field_evaluate(rpx_ctx* ctx, rpx_field* f, const double* points, uint64_t npt, double time_ps,
                                  double* out) {
    if (!ctx || !f || ((!points || !out) && npt)) return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    if (!npt) return RPX_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    double* d_pts = nullptr;
    double* d_out = nullptr;
    CU(ctx, cudaMallocAsync((void**)&d_pts, npt * 3 * sizeof(double), ctx->stream));
    cudaError_t e = cudaMallocAsync((void**)&d_out, npt * 6 * sizeof(double), ctx->stream);
    if (e != cudaSuccess) {
        cudaFreeAsync(d_pts, ctx->stream);
        return fail(ctx, RPX_ERR_NOMEM, "field output (%llu points): %s", (unsigned long long)npt, cudaGetErrorString(e));
    }
    int rc = RPX_OK;
    if ((e = cudaMemcpyAsync(d_pts, points, npt * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
        (e = cudaMemsetAsync(d_out, 0, npt * 6 * sizeof(double), ctx->stream)) != cudaSuccess)
        rc = fail(ctx, RPX_ERR_CUDA, "field input upload: %s", cudaGetErrorString(e));
    if (rc == RPX_OK) rc = rpx_field_evaluate_device(ctx, f, d_pts, npt, time_ps, d_out);
    if (rc == RPX_OK &&
        ((e = cudaMemcpyAsync(out, d_out, npt * 6 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess ||
         (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess))
        rc = fail(ctx, RPX_ERR_CUDA, "field download: %s", cudaGetErrorString(e));
    cudaFreeAsync(d_pts, ctx->stream);
    cudaFreeAsync(d_out, ctx->stream);
    return rc;
}

extern "C" double rpx_field_last_ms(const rpx_field* f) { return f ? (double)f->last_ms : 0.0; }

extern "C" void rpx_field_free(rpx_ctx* ctx, rpx_field* f) {
    if (!f) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        if (f->rec) cudaFreeAsync(f->rec, ctx->stream);
        if (f->modes) cudaFreeAsync(f->modes, ctx->stream);
    }
    delete f;
}

// ------------------------------------------------------------------ detector (accumulating field)
// EFieldSummation (fields.py:206-249) as a long-lived device object: points and wavelengths uploaded once,
// gausslet collections summed into one field buffer collection after collection, nothing synchronised
// until the field is read -- the consumer at the end of rpx_trace_consume's chunk pipeline.
struct rpx_detector {
    double* d_points;
    double* d_field;  // npt x 6
    double* d_wl;
    int n_wl;
    uint64_t npt;
    double blending, time_ps;
    uint64_t modes;
    double ms_done;
    std::vector<cudaEvent_t> ev;  // start / stop pairs not yet harvested
};

static void detector_harvest(rpx_detector* det, bool wait) {
    size_t keep = 0;
    for (size_t i = 0; i + 1 < det->ev.size(); i += 2) {
        const bool done = wait ? (cudaEventSynchronize(det->ev[i + 1]) == cudaSuccess)
                               : (cudaEventQuery(det->ev[i + 1]) == cudaSuccess);
        if (done) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, det->ev[i], det->ev[i + 1]) == cudaSuccess) det->ms_done += ms;
            cudaEventDestroy(det->ev[i]);
            cudaEventDestroy(det->ev[i + 1]);
        } else {
            det->ev[keep++] = det->ev[i];
            det->ev[keep++] = det->ev[i + 1];
        }
    }
    det->ev.resize(keep);
}

extern "C" int rpx_detector_create(rpx_ctx* ctx, const double* points, uint64_t npt, const double* wavelengths,
                                   int n_wavelengths, double blending, double time_ps, rpx_detector** out) {
    if (!ctx || !out || (!points && npt) || !wavelengths || n_wavelengths <= 0)
        return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    *out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    rpx_detector* det = new (std::nothrow) rpx_detector();
    if (!det) return fail(ctx, RPX_ERR_NOMEM, "out of host memory");
    det->d_points = det->d_field = det->d_wl = nullptr;
    det->n_wl = n_wavelengths;
    det->npt = npt;
    det->blending = blending;
    det->time_ps = time_ps;
    det->modes = 0;
    det->ms_done = 0.0;
    const size_t np_ = npt ? npt : 1;
    cudaError_t e;
    if ((e = cudaMalloc((void**)&det->d_points, np_ * 3 * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc((void**)&det->d_field, np_ * 6 * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc((void**)&det->d_wl, sizeof(double) * (size_t)n_wavelengths)) != cudaSuccess ||
        (npt && (e = cudaMemcpyAsync(det->d_points, points, npt * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) ||
        (e = cudaMemcpyAsync(det->d_wl, wavelengths, sizeof(double) * (size_t)n_wavelengths, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
        (e = cudaMemsetAsync(det->d_field, 0, np_ * 6 * sizeof(double), ctx->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) {
        if (det->d_points) cudaFree(det->d_points);
        if (det->d_field) cudaFree(det->d_field);
        if (det->d_wl) cudaFree(det->d_wl);
        delete det;
        return fail(ctx, e == cudaErrorMemoryAllocation ? RPX_ERR_NOMEM : RPX_ERR_CUDA, "detector set-up: %s", cudaGetErrorString(e));
    }
    *out = det;
    return RPX_OK;
}

extern "C" int rpx_detector_reset(rpx_ctx* ctx, rpx_detector* det) {
    if (!ctx || !det) return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    CU(ctx, cudaSetDevice(ctx->device));
    detector_harvest(det, true);
    CU(ctx, cudaMemsetAsync(det->d_field, 0, (det->npt ? det->npt : 1) * 6 * sizeof(double), ctx->stream));
    det->modes = 0;
    det->ms_done = 0.0;
    return RPX_OK;
}

// No host synchronisation: the mode records live in the stream-ordered pool for exactly two kernels.
extern "C" int rpx_detector_accumulate(rpx_ctx* ctx, rpx_detector* det, const rpx_rays* rays) {
    if (!ctx || !det || !rays) return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    if (!rays->is_gausslet)
        return fail(ctx, RPX_ERR_INVALID, "a detector sums gausslets (plain rays carry no parabasal rays to fit a mode to)");
    const uint64_t n = rays->soa.n;
    if (!n || !det->npt) return RPX_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    detector_harvest(det, false);
    rpx_field f;
    f.n = n;
    f.rec = f.modes = nullptr;
    f.last_ms = 0.f;
    cudaError_t e;
    if ((e = cudaMallocAsync((void**)&f.rec, n * M_NF * sizeof(double), ctx->stream)) != cudaSuccess ||
        (e = cudaMallocAsync((void**)&f.modes, n * 6 * sizeof(double), ctx->stream)) != cudaSuccess) {
        if (f.rec) cudaFreeAsync(f.rec, ctx->stream);
        return fail(ctx, RPX_ERR_NOMEM, "mode records of %llu gausslets: %s", (unsigned long long)n, cudaGetErrorString(e));
    }
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, ctx->stream);
    const unsigned T = 128, G = (unsigned)((n + T - 1) / T);
    k_field_prepare<true><<<G, T, 0, ctx->stream>>>(rays->soa, nullptr, det->d_wl, det->n_wl, det->blending, f.rec, f.modes);
    int rc = (e = cudaGetLastError()) == cudaSuccess ? field_launch(ctx, &f, det->d_points, det->npt, det->time_ps, det->d_field, nullptr, nullptr)
                                                     : fail(ctx, RPX_ERR_CUDA, "k_field_prepare launch: %s", cudaGetErrorString(e));
    cudaEventRecord(ev1, ctx->stream);
    det->ev.push_back(ev0);
    det->ev.push_back(ev1);
    cudaFreeAsync(f.rec, ctx->stream);
    cudaFreeAsync(f.modes, ctx->stream);
    if (rc == RPX_OK) det->modes += n;
    return rc;
}

extern "C" int rpx_detector_read(rpx_ctx* ctx, rpx_detector* det, double* field_out) {
    if (!ctx || !det || (!field_out && det->npt)) return fail(ctx, RPX_ERR_INVALID, "NULL argument");
    CU(ctx, cudaSetDevice(ctx->device));
    if (det->npt) CU(ctx, cudaMemcpyAsync(field_out, det->d_field, det->npt * 6 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return RPX_OK;
}

extern "C" void* rpx_detector_field_device(rpx_detector* det) { return det ? (void*)det->d_field : nullptr; }
extern "C" uint64_t rpx_detector_npoints(const rpx_detector* det) { return det ? det->npt : 0; }
extern "C" uint64_t rpx_detector_modes(const rpx_detector* det) { return det ? det->modes : 0; }

extern "C" double rpx_detector_ms(rpx_ctx* ctx, rpx_detector* det) {
    if (!ctx || !det) return 0.0;
    cudaSetDevice(ctx->device);
    detector_harvest(det, true);
    return det->ms_done;
}

extern "C" void rpx_detector_free(rpx_ctx* ctx, rpx_detector* det) {
    if (!det) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        detector_harvest(det, true);
        cudaFree(det->d_points);
        cudaFree(det->d_field);
        cudaFree(det->d_wl);
    }
    delete det;
}
