// rpx_materials.cuh -- device evaluation of every InterfaceMaterial of the reference
// (raypier/core/cmaterials.pyx): S/P projection, polarised Fresnel coefficients with
// complex indices, Snell refraction, total internal reflection, single-layer thin-film
// transfer matrices, reflective grating orders, soft apertures, wave plates.
//
// A parent ray yields up to two children, always in the reference's emission order
// (reflected first, then transmitted): slot a, then slot b.  Fields that both children
// share are stored once (Kids::origin/normal/evec/phase/apath).
#pragma once
#include "rpx_faces.cuh"

namespace rpx {

#define RPX_SP_TOL 1.0e-10  // cmaterials.pyx:32

// One parent ray in registers (what the shade kernels load from the SoA buffers).
struct RayIn {
    vec3 o, d, e;     // origin, direction, E_vector
    cplx n, e1, e2;   // refractive_index, E1_amp, E2_amp
    double len, phase, apath;
    uint32_t wl, ident, type;
};

struct Kid {
    vec3 dir;
    cplx n, e1, e2;
    uint32_t type;
};

struct Kids {
    bool has_a, has_b;
    vec3 origin, normal, evec;
    double phase, apath;
    Kid a, b;
};

// convert_to_sp, cmaterials.pyx:49-91.  Outputs the new E_vector and (E1, E2) =
// (S, P) amplitudes; returns them unchanged when direction x normal is ~0.
RPX_DEV void convert_to_sp(const RayIn& r, vec3 normal, vec3* evec, cplx* s_amp, cplx* p_amp) {
    vec3 E2_vector = norm(cross(r.d, r.e));
    vec3 E1_vector = norm(cross(E2_vector, r.d));
    normal = norm(normal);
    vec3 S_vector = cross(r.d, normal);
    if (fabs(S_vector.x) < RPX_SP_TOL && fabs(S_vector.y) < RPX_SP_TOL && fabs(S_vector.z) < RPX_SP_TOL) {
        *evec = r.e;
        *s_amp = r.e1;
        *p_amp = r.e2;
        return;
    }
    S_vector = norm(S_vector);
    vec3 P_vector = norm(cross(r.d, S_vector));
    double A = dot(E1_vector, S_vector);
    double B = dot(E2_vector, S_vector);
    *s_amp = cx(r.e1.re * A + r.e2.re * B, r.e1.im * A + r.e2.im * B);
    B = dot(E1_vector, P_vector);
    A = dot(E2_vector, P_vector);
    *p_amp = cx(r.e1.re * B + r.e2.re * A, r.e1.im * B + r.e2.im * A);
    *evec = S_vector;
}

RPX_DEV cplx ntab_get(const DevScene& S, const rpx_material* M, int row, uint32_t wl) {
    const double* t = S.ntab + 2 * ((size_t)M->ntab_off + (size_t)row * S.n_wl + wl);
    return cx(t[0], t[1]);
}

// Emission tail shared by the uncoated and coated Fresnel materials
// (cmaterials.pyx:826-872, 967-1013, 1140-1182, 1349-1391).
RPX_DEV void fresnel_emit(Kids& k, const RayIn& r, vec3 normal, vec3 in_direction, double cosTheta,
                          int flip, cplx n1, cplx n_t, cplx R_s, cplx R_p, cplx T_s, cplx T_p,
                          double P_in, double refl_thr, double trans_thr) {
    vec3 cosThetaNormal = normal * cosTheta;
    const double invP = rcp(P_in);
    if ((n1.re * (cabs2(R_s) + cabs2(R_p)) * invP) > refl_thr) {
        k.has_a = true;
        k.a.dir = in_direction - cosThetaNormal * 2.0;
        k.a.e1 = R_s;
        k.a.e2 = -R_p;
        k.a.n = n1;
        k.a.type = r.type | RPX_REFL_RAY;
    }
    if ((n_t.re * (cabs2(T_s) + cabs2(T_p)) * invP) > trans_thr) {
        vec3 tangent = in_direction - cosThetaNormal;
        vec3 tg2 = tangent * (n1.re * rcp(n_t.re));  // real-part approximation, :844
        double tan_mag_sq = mag_sq(tg2);
        double c2 = sqrt_(1 - tan_mag_sq);
        k.has_b = true;
        k.b.dir = tg2 - normal * (c2 * flip);
        k.b.e1 = T_s;
        k.b.e2 = T_p;
        k.b.n = n_t;
        k.b.type = r.type & ~RPX_REFL_RAY;
    }
}

// Material-class specialisation: MM is a bit mask of the rpx_material_type values a kernel
// variant is compiled for (bit t set = type t supported); the host picks the smallest
// compiled mask that covers the scene.  Unsupported types cannot occur (checked on the host).
#define RPX_MBIT(t) (1u << (t))
#define RPX_MM_LIGHT                                                                             \
    (RPX_MBIT(RPX_MAT_OPAQUE) | RPX_MBIT(RPX_MAT_TRANSPARENT) | RPX_MBIT(RPX_MAT_PEC) |          \
     RPX_MBIT(RPX_MAT_PARTIALLY_REFLECTIVE) | RPX_MBIT(RPX_MAT_LINEAR_POLARISING) |              \
     RPX_MBIT(RPX_MAT_WAVEPLATE) | RPX_MBIT(RPX_MAT_DIELECTRIC) | RPX_MBIT(RPX_MAT_GRATING) |    \
     RPX_MBIT(RPX_MAT_CIRC_APERTURE) | RPX_MBIT(RPX_MAT_RECT_APERTURE))
#define RPX_MM_COATED (RPX_MBIT(RPX_MAT_COATED) | RPX_MBIT(RPX_MAT_OPAQUE) | RPX_MBIT(RPX_MAT_PEC))
#define RPX_MM_FULLDIEL                                                                          \
    (RPX_MBIT(RPX_MAT_FULL_DIELECTRIC) | RPX_MBIT(RPX_MAT_OPAQUE) | RPX_MBIT(RPX_MAT_PEC) |      \
     RPX_MBIT(RPX_MAT_PARTIALLY_REFLECTIVE) | RPX_MBIT(RPX_MAT_LINEAR_POLARISING) |              \
     RPX_MBIT(RPX_MAT_TRANSPARENT))
#define RPX_MM_ALL (RPX_MM_LIGHT | RPX_MBIT(RPX_MAT_FULL_DIELECTRIC) | RPX_MBIT(RPX_MAT_COATED))
#define RPX_N_MM 4  // compiled variants, in this order: LIGHT, COATED, FULLDIEL, ALL

// InterfaceMaterial.eval_child_ray_c for every material class.
//   point           hit point, global coordinates
//   onormal/otangent  FaceList.compute_orientation_c output (not yet normalised)
#ifndef RPX_MAT_INLINE
#define RPX_MAT_INLINE 0
#endif
#if RPX_MAT_INLINE
#define RPX_MAT_ATTR __device__ __forceinline__
#else
#define RPX_MAT_ATTR __device__
#endif
template <uint32_t MM>
RPX_MAT_ATTR void material_eval(const DevScene& S, const rpx_material* M, const RayIn& r, vec3 point,
                              vec3 onormal, vec3 otangent, Kids& k) {
    const double* P = M->p;
    k.has_a = false;
    k.has_b = false;
    if (M->type == RPX_MAT_OPAQUE) return;  // :245-251

    vec3 normal = norm(onormal);
    k.origin = point;
    k.normal = normal;
    k.phase = r.phase;
    k.apath = r.apath + r.len * r.n.re;  // accumulated_path += length * n.real

    cplx s_amp, p_amp;
    if ((MM & RPX_MBIT(RPX_MAT_WAVEPLATE)) && M->type == RPX_MAT_WAVEPLATE)
        convert_to_sp(r, ld3(P + 2), &k.evec, &s_amp, &p_amp);  // :541
    else convert_to_sp(r, normal, &k.evec, &s_amp, &p_amp);

    switch (M->type) {
        case RPX_MAT_TRANSPARENT: {  // :260-278
            if (!(MM & RPX_MBIT(RPX_MAT_TRANSPARENT))) break;
            k.has_a = true;
            k.a.dir = r.d;
            k.a.n = r.n;
            k.a.e1 = s_amp;
            k.a.e2 = p_amp;
            k.a.type = r.type & ~RPX_REFL_RAY;
        } break;
        case RPX_MAT_PEC: {  // :285-319 (uses the un-normalised incoming direction)
            if (!(MM & RPX_MBIT(RPX_MAT_PEC))) break;
            double cosTheta = dot(normal, r.d);
            k.has_a = true;
            k.a.dir = r.d - (normal * cosTheta) * 2.0;
            k.a.n = r.n;
            k.a.e1 = -s_amp;
            k.a.e2 = p_amp;
            k.a.type = r.type | RPX_REFL_RAY;
        } break;
        case RPX_MAT_PARTIALLY_REFLECTIVE:  // :345-397
        case RPX_MAT_LINEAR_POLARISING: {   // :404-455
            if (!(MM & RPX_MBIT(RPX_MAT_PARTIALLY_REFLECTIVE))) break;
            vec3 in_direction = norm(r.d);
            double cosTheta = dot(normal, in_direction);
            k.has_a = true;
            k.has_b = true;
            k.a.dir = in_direction - (normal * cosTheta) * 2.0;
            k.b.dir = in_direction;
            k.a.n = r.n;
            k.b.n = r.n;
            k.a.type = r.type | RPX_REFL_RAY;
            k.b.type = r.type & ~RPX_REFL_RAY;
            if (M->type == RPX_MAT_PARTIALLY_REFLECTIVE) {
                double R = sqrt_(P[0]);
                double T = sqrt_(1 - P[0]);
                k.a.e1 = s_amp * R;
                k.a.e2 = p_amp * R;
                k.b.e1 = s_amp * T;
                k.b.e2 = p_amp * T;
            } else {
                k.a.e1 = s_amp;
                k.a.e2 = cx(0.0, 0.0);
                k.b.e1 = cx(0.0, 0.0);
                k.b.e2 = p_amp;
            }
        } break;
        case RPX_MAT_WAVEPLATE: {  // :520-551
            if (!(MM & RPX_MBIT(RPX_MAT_WAVEPLATE))) break;
            k.has_a = true;
            k.a.dir = norm(r.d);
            k.a.n = r.n;
            k.a.e1 = cx(s_amp.re * P[0] - s_amp.im * P[1], s_amp.im * P[0] + s_amp.re * P[1]);
            k.a.e2 = p_amp;
            k.a.type = r.type & ~RPX_REFL_RAY;
        } break;
        case RPX_MAT_DIELECTRIC: {  // :587-680
            if (!(MM & RPX_MBIT(RPX_MAT_DIELECTRIC))) break;
            cplx n_inside = ntab_get(S, M, 0, r.wl), n_outside = ntab_get(S, M, 1, r.wl);
            vec3 in_direction = norm(r.d);
            double cosTheta = dot(normal, in_direction);
            double cos1 = fabs(cosTheta);
            double n1, n2;
            int flip;
            k.has_a = true;
            if (cosTheta < 0.0) {
                n1 = n_outside.re;
                n2 = n_inside.re;
                k.a.n = n_inside;  // assigned before the TIR test (quirk Q15)
                flip = 1;
            } else {
                n1 = n_inside.re;
                n2 = n_outside.re;
                k.a.n = n_outside;
                flip = -1;
            }
            double N2 = (n2 / n1) * (n2 / n1);
            double N2_sin2 = (cosTheta * cosTheta) + (N2 - 1);
            vec3 cosThetaNormal = normal * cosTheta;
            if (N2_sin2 < 0.0) {  // total internal reflection
                k.a.dir = in_direction - cosThetaNormal * 2.0;
                k.a.e1 = -s_amp;
                k.a.e2 = -p_amp;
                k.a.type = r.type | RPX_REFL_RAY;
            } else {
                vec3 tangent = in_direction - cosThetaNormal;
                vec3 tg2 = tangent * (n1 / n2);
                double tan_mag_sq = mag_sq(tg2);
                double c2 = sqrt_(1 - tan_mag_sq);
                vec3 transmitted = tg2 - normal * (c2 * flip);
                double cos2 = fabs(dot(transmitted, normal));
                double Two_n1_cos1 = (2 * n1) * cos1;
                double aspect = sqrt_(cos2 / cos1) * Two_n1_cos1;
                double T_p = aspect / (n2 * cos1 + n1 * cos2);
                double T_s = aspect / (n2 * cos2 + n1 * cos1);
                k.a.dir = transmitted;
                k.a.e1 = s_amp * T_s;
                k.a.e2 = p_amp * T_p;
                k.a.type = r.type & ~RPX_REFL_RAY;
            }
        } break;
        case RPX_MAT_FULL_DIELECTRIC: {  // :755-872, :896-1013
            if (!(MM & RPX_MBIT(RPX_MAT_FULL_DIELECTRIC))) break;
            vec3 in_direction = norm(r.d);
            double cosTheta = dot(normal, in_direction);
            double cos1 = fabs(cosTheta);
            double sin1 = sqrt_(fabs(1 - cos1 * cos1));
            cplx n1, n2;
            int flip;
            if (cosTheta < 0.0) {
                n1 = ntab_get(S, M, 1, r.wl);
                n2 = ntab_get(S, M, 0, r.wl);
                flip = 1;
            } else {
                n1 = ntab_get(S, M, 0, r.wl);
                n2 = ntab_get(S, M, 1, r.wl);
                flip = -1;
            }
            double P_in = n1.re * (s_amp.re * s_amp.re + s_amp.im * s_amp.im + p_amp.re * p_amp.re +
                                   p_amp.im * p_amp.im);
            if (P_in == 0.0) return;
            cplx R_s, R_p, T_s, T_p;
            // Lossless media below the critical angle (the common case): every Fresnel term is
            // REAL, so the complex products and reciprocals of the general form collapse to real
            // ones.  Same formulae, same values to rounding; absorbing media and TIR take the
            // general complex path below.
            const double sin2r = (n1.re * sin1) * rcp(n2.re);
            const double c2sq = 1.0 - sin2r * sin2r;
            if (n1.im == 0.0 && n2.im == 0.0 && c2sq > 0.0) {
                const double cos2 = sqrt_(c2sq);
                const double n2c1 = n2.re * cos1, n1c2 = n1.re * cos2, n2c2 = n2.re * cos2, n1c1 = n1.re * cos1;
                const double idp = rcp(n2c1 + n1c2), ids = rcp(n2c2 + n1c1);
                R_p = p_amp * ((n1c2 - n2c1) * idp);
                R_s = s_amp * ((n1c1 - n2c2) * ids);
                const double num = (n1.re * (2.0 * cos1)) * sqrt_(cos2 * rcp(cos1));
                T_p = p_amp * (num * idp);
                T_s = s_amp * (num * ids);
            } else {
                cplx sin2 = (n1 * sin1) / n2;
                cplx cos2 = csqrt_(cx(1.0, 0.0) - sin2 * sin2);
                cplx n2c1 = n2 * cos1, n1c2 = n1 * cos2, n2c2 = n2 * cos2, n1c1 = n1 * cos1;
                cplx idp = crcp(n2c1 + n1c2), ids = crcp(n2c2 + n1c1);  // the two Fresnel denominators
                R_p = (-(n2c1 - n1c2)) * idp;
                R_s = (-(n2c2 - n1c1)) * ids;
                R_s = R_s * s_amp;
                R_p = R_p * p_amp;
                double aspect = sqrt_(cos2.re * rcp(cos1));
                cplx num = (n1 * (2.0 * cos1)) * aspect;
                T_p = (num * idp) * p_amp;
                T_s = (num * ids) * s_amp;
            }
            fresnel_emit(k, r, normal, in_direction, cosTheta, flip, n1, n2, R_s, R_p, T_s, T_p, P_in,
                         P[0], P[1]);
        } break;
        case RPX_MAT_COATED: {  // :1026-1182, :1228-1391: single-layer thin film
            if (!(MM & RPX_MBIT(RPX_MAT_COATED))) break;
            double wavelength = S.wavelengths[r.wl];
            vec3 in_direction = norm(r.d);
            double cosTheta = dot(normal, in_direction);
            double cos1 = fabs(cosTheta);
            double sin1 = sqrt_(fabs(1 - cos1 * cos1));
            cplx n2 = ntab_get(S, M, 2, r.wl);
            cplx n1, n3;
            int flip;
            if (cosTheta < 0.0) {
                n1 = ntab_get(S, M, 1, r.wl);
                n3 = ntab_get(S, M, 0, r.wl);
                flip = 1;
            } else {
                n1 = ntab_get(S, M, 0, r.wl);
                n3 = ntab_get(S, M, 1, r.wl);
                flip = -1;
            }
            double P_in = n1.re * (s_amp.re * s_amp.re + s_amp.im * s_amp.im + p_amp.re * p_amp.re +
                                   p_amp.im * p_amp.im);
            if (P_in == 0.0) return;
            const double dwc = 2 * M_PI * P[2] * rcp(wavelength);
            cplx R_s, T_s, R_p, T_p;
            double aspect;
            // ep1 = exp(phi) / (4 n2cos2 n3cos3),  ep2 = exp(-2 phi) = 1 / exp(phi)^2.
            // The reference multiplies every transfer-matrix entry by +-ep1 and then forms
            // R = -M00/M01 and T = M10 + M11*R (:1101-1114): ep1 cancels in R and is a common
            // factor of T, so it is applied once.  Same value to a few ulp, ~40% fewer flops.
            //   M00 = -ep1*X, M01 = ep1*Y, M10 = ep1*U, M11 = -ep1*V
            const double n1sr = n1.re * sin1;
            const double sin2r = n1sr * rcp(n2.re), sin3r = n1sr * rcp(n3.re);
            const double c2sq = 1.0 - sin2r * sin2r, c3sq = 1.0 - sin3r * sin3r;
            if (n1.im == 0.0 && n2.im == 0.0 && n3.im == 0.0 && c2sq > 0.0 && c3sq > 0.0) {
                // Lossless film and media, no evanescent wave in the film or the substrate (the
                // common case): cos2, cos3 and all n*cos products are REAL and phi is purely
                // imaginary, so exp(phi) has unit modulus, ep2 = conj(exp(phi))^2 needs no
                // reciprocal, and the matrix entries are real + real * ep2.
                const double cos2 = sqrt_(c2sq), cos3 = sqrt_(c3sq);
                const double n2c2 = n2.re * cos2, n3c3 = n3.re * cos3;
                const double a = dwc * (n2.re - sin2r * sin2r) * rcp(cos2);  // phi = -i a
                double sa, ca;
                sincos_phase(a, &sa, &ca);
                const double g = rcp((n2c2 * 4.0) * n3c3);
                const cplx ep1 = cx(ca * g, -sa * g);
                const double e2r = ca * ca - sa * sa, e2i = 2.0 * sa * ca;
                auto film = [&](double am, double ap, double bm, double bp, cplx& R, cplx& T) {
                    const double ambp = am * bp, apbm = ap * bm, ambm = am * bm, apbp = ap * bp;
                    const cplx X = cx(ambp + apbm * e2r, apbm * e2i), Y = cx(ambm * e2r + apbp, ambm * e2i);
                    const cplx U = cx(ambm + apbp * e2r, apbp * e2i), V = cx(ambp * e2r + apbm, ambp * e2i);
                    R = X / Y;
                    T = ep1 * (U - V * R);
                };
                const double n1c1 = n1.re * cos1;
                film(n1c1 - n2c2, n1c1 + n2c2, n2c2 - n3c3, n2c2 + n3c3, R_s, T_s);
                const double n1c2 = n1.re * cos2, n2c1 = n2.re * cos1, n2c3 = n2.re * cos3, n3c2 = n3.re * cos2;
                film(n1c2 - n2c1, n1c2 + n2c1, n2c3 - n3c2, n2c3 + n3c2, R_p, T_p);
                aspect = sqrt_(cos3 * rcp(cos1));
            } else {
                cplx n1s = n1 * sin1;
                cplx sin2 = n1s / n2;
                cplx cos2 = csqrt_(cx(1.0, 0.0) - sin2 * sin2);
                cplx sin3 = n1s / n3;
                cplx cos3 = csqrt_(cx(1.0, 0.0) - sin3 * sin3);
                cplx n1cos1 = n1 * cos1;
                cplx n2cos2 = n2 * cos2;
                cplx n3cos3 = n3 * cos3;
                // phi = -I*dwc*(n2 - sin2*sin2)/cos2   (cmaterials.pyx:1098)
                cplx phi = (cx(0.0, -dwc) * (n2 - sin2 * sin2)) / cos2;
                cplx ephi = cexp_(phi);
                cplx ep1 = ephi / ((n2cos2 * 4.0) * n3cos3);
                cplx ep2 = crcp(ephi * ephi);
                auto film = [&](cplx am, cplx ap, cplx bm, cplx bp, cplx& R, cplx& T) {
                    cplx ambp = am * bp, apbm = ap * bm, ambm = am * bm, apbp = ap * bp;
                    cplx X = ambp + apbm * ep2, Y = ambm * ep2 + apbp;
                    cplx U = ambm + apbp * ep2, V = ambp * ep2 + apbm;
                    R = X / Y;
                    T = ep1 * (U - V * R);
                };
                film(n1cos1 - n2cos2, n1cos1 + n2cos2, n2cos2 - n3cos3, n2cos2 + n3cos3, R_s, T_s);
                cplx n1cos2 = n1 * cos2, n2cos1 = n2 * cos1, n2cos3 = n2 * cos3, n3cos2 = n3 * cos2;
                film(n1cos2 - n2cos1, n1cos2 + n2cos1, n2cos3 - n3cos2, n2cos3 + n3cos2, R_p, T_p);
                aspect = sqrt_(cos3.re * rcp(cos1));
            }
            R_s = R_s * s_amp;
            R_p = R_p * p_amp;
            T_s = T_s * (s_amp * aspect);
            T_p = T_p * (p_amp * aspect);
            fresnel_emit(k, r, normal, in_direction, cosTheta, flip, n1, n3, R_s, R_p, T_s, T_p, P_in,
                         P[0], P[1]);
        } break;
        case RPX_MAT_GRATING: {  // :1472-1542
            if (!(MM & RPX_MBIT(RPX_MAT_GRATING))) break;
            vec3 tangent = norm(otangent);
            vec3 tangent2 = cross(normal, tangent);
            double wavelen = S.wavelengths[r.wl];
            double line_spacing = 1000.0 / P[0];
            double order = (double)(int)P[1];
            vec3 reflected = norm(r.d);
            double k_z = dot(normal, reflected);
            double k_y = dot(tangent2, reflected);
            double k_x = dot(tangent, reflected);
            int sign = (k_z < 0.0) ? 1 : -1;
            k_x = k_x - order * wavelen / (line_spacing * r.n.re);
            k_z = 1 - (k_x * k_x) - (k_y * k_y);
            if (k_z < 0) return;  // evanescent order
            k_z = sign * sqrt_(k_z);
            reflected = tangent * k_x;
            reflected = reflected + tangent2 * k_y;
            reflected = reflected + normal * k_z;
            k.has_a = true;
            k.a.dir = reflected;
            k.a.n = r.n;
            k.a.e1 = cx(-s_amp.re * P[2], -s_amp.im * P[2]);
            k.a.e2 = cx(p_amp.re * P[2], p_amp.im * P[2]);
            k.a.type = r.type | RPX_REFL_RAY;
            k.phase = r.phase + 1000.0 * dot(ld3(P + 3) - point, tangent) * order * 2 * M_PI / line_spacing;
        } break;
        case RPX_MAT_CIRC_APERTURE: {  // :1641-1674
            if (!(MM & RPX_MBIT(RPX_MAT_CIRC_APERTURE))) break;
            double rr = sqrt_(mag_sq(ld3(P + 4) - point));
            if (rr > P[0]) return;
            double atten = 0.5 + 0.5 * erf((P[1] - rr) / P[2]);
            if (P[3] != 0.0) atten = 1 - atten;
            k.has_a = true;
            k.a.dir = r.d;
            k.a.n = r.n;
            k.a.e1 = s_amp * atten;
            k.a.e2 = p_amp * atten;
            k.a.type = r.type & ~RPX_REFL_RAY;
        } break;
        case RPX_MAT_RECT_APERTURE: {  // :1718-1763 (uses the un-normalised orientation)
            if (!(MM & RPX_MBIT(RPX_MAT_RECT_APERTURE))) break;
            double width = P[4];
            double x = P[2] / 2., y = P[3] / 2.;
            vec3 p = point - ld3(P + 6);
            double px = dot(p, otangent);
            double py = dot(p, cross(onormal, otangent));
            if (fabs(px) > P[0] / 2.) return;
            if (fabs(py) > P[1] / 2.) return;
            double atten = 0.5 - 0.5 * erf((px - x) / width);
            atten *= 0.5 - 0.5 * erf(-(px + x) / width);
            atten *= 0.5 - 0.5 * erf((py - y) / width);
            atten *= 0.5 - 0.5 * erf(-(py + y) / width);
            if (P[5] != 0.0) atten = 1 - atten;
            k.has_a = true;
            k.a.dir = r.d;
            k.a.n = r.n;
            k.a.e1 = s_amp * atten;
            k.a.e2 = p_amp * atten;
            k.a.type = r.type & ~RPX_REFL_RAY;
        } break;
        default: break;
    }
}

// InterfaceMaterial.eval_parabasal_ray_c -> outgoing parabasal direction.
// (origin = point, normal = norm(orient.normal), length = INF are set by the caller.)
// What the Snell model needs of the material per (thread, wavelength): the two index ratios n1 / n2 it can form.
// The twelve calls of one gausslet (six parabasal rays x two children) share them, so the caller reads the two
// table entries (global memory) and does the two IEEE divisions ONCE instead of twelve times each -- the loads sat
// between the parabasal stores, where ptxas cannot hoist them itself (ncu: 8.7 % of the stall samples of a
// two-children generation waited on them).  Same operands, same IEEE division: bit-identical directions.
struct ParaSnell {
    double ratio_neg;  // cosTheta <  0: n1 = row 1, n2 = row 0
    double ratio_pos;  // cosTheta >= 0: n1 = row 0, n2 = row 1
};
RPX_DEV ParaSnell para_snell_setup(const DevScene& S, const rpx_material* M, uint32_t wl) {
    ParaSnell ps;
    ps.ratio_neg = ps.ratio_pos = 0.0;
    if (M->para_model == RPX_PARA_SNELL) {
        const double a = ntab_get(S, M, 0, wl).re, b = ntab_get(S, M, 1, wl).re;
        ps.ratio_neg = b / a;
        ps.ratio_pos = a / b;
    }
    return ps;
}

RPX_DEV vec3 material_eval_para(const DevScene& S, const rpx_material* M, uint32_t wl, double base_n_re,
                                vec3 direction, vec3 point, vec3 onormal, vec3 otangent,
                                uint32_t ray_type_id, const ParaSnell& ps) {
    vec3 normal = norm(onormal);
    if (M->para_model == RPX_PARA_SNELL) {  // cmaterials.pyx:683-724, 1393-1434
        direction = norm(direction);
        double cosTheta = dot(normal, direction);
        vec3 cosThetaNormal = normal * cosTheta;
        const double n1_over_n2 = (cosTheta < 0.0) ? ps.ratio_neg : ps.ratio_pos;
        const int flip = (cosTheta < 0.0) ? 1 : -1;
        if (ray_type_id & RPX_REFL_RAY) return direction - cosThetaNormal * 2.0;
        vec3 tangent = direction - cosThetaNormal;
        vec3 tg2 = tangent * n1_over_n2;
        double tan_mag_sq = mag_sq(tg2);
        double c2 = sqrt_(1 - tan_mag_sq);
        return tg2 - normal * (c2 * flip);
    }
    if (M->para_model == RPX_PARA_GRATING) {  // :1544-1599
        const double* P = M->p;
        vec3 tangent = norm(otangent);
        vec3 tangent2 = cross(normal, tangent);
        double wavelen = S.wavelengths[wl];
        double line_spacing = 1000.0 / P[0];
        double order = (double)(int)P[1];
        vec3 reflected = norm(direction);
        double k_z = dot(normal, reflected);
        double k_y = dot(tangent2, reflected);
        double k_x = dot(tangent, reflected);
        int sign = (k_z < 0.0) ? 1 : -1;
        k_x = k_x - order * wavelen / (line_spacing * base_n_re);
        k_z = 1 - (k_x * k_x) - (k_y * k_y);
        k_z = sign * sqrt_(k_z);  // evanescent -> NaN, as in the reference (it only prints)
        reflected = tangent * k_x;
        reflected = reflected + tangent2 * k_y;
        reflected = reflected + normal * k_z;
        return reflected;
    }
    // default, ctracer.pyx:1588-1610
    if (ray_type_id & RPX_REFL_RAY) {
        double cosTheta = dot(normal, direction);
        return direction - (normal * cosTheta) * 2.0;
    }
    return direction;
}

}  // namespace rpx
