// BlueROV2 Heavy 6DoF manoeuvring model, one environment per thread, all
// intermediate state in registers.  Restates (does not copy) the algorithm of
// dynamicsModel_BlueROV2_Heavy_6DoF.py:220-442 and resources.py:98-143; the
// reference's quirks are kept on purpose ("bug-compatibility").
#pragma once
#include "mvrl_math.cuh"

namespace mvrl {

// Constants of one vehicle, converted to the compute type on the host and
// passed to every kernel as a __grid_constant__ argument (constant bank, one
// copy per launch, handle-specific, CUDA-graph friendly).
template <typename T> struct Rov6Dev {
    T m, xg, yg, zg;
    T Ixx, Iyy, Izz, Ixy, Ixz, Iyz;
    T Xud, Yvd, Zwd, Kpd, Mqd, Nrd;
    T Xu, Yv, Yp, Yr, Zw, Zq, Kv, Kp, Kr, Mw, Mq, Nv, Np, Nr;
    T Xuu, Yvv, Ypp, Yrr, Zww, Zqq, Kvv, Kpp, Krr, Mww, Mqq, Nvv, Npp, Nrr;
    T WmB, gx, gy, gz;        // W-B and (xg W - xb B), (yg W - yb B), (zg W - zb B)
    T A[6][8];
    T Ainv[8][6];
    T Minv[6][6];
    T thrust_k;               // thrust_coef / 3600: F = thrust_k * rpm * |rpm|
    T inv_thrust_coef;        // 1 / (rho D^4 Kt)
    T rpm_max, rpm_db;
    T f_max, f_db;            // thruster force at rpm_max / at the deadband edge
    T pKp[6], pKi[6], pKd[6], pWind[6], pMax[6];
    T inv_3L, act_pos, act_ang, inv_ang;  // 1/(3 Length), 2 Length, pi/4, 4/pi
    // Crb + Ca folded for the default sparsity (see body_accel): effective masses
    // m - Xudot.., m*zg, and the differences that multiply the velocity products
    T mX, mY, mZ, mzg, cVW, cUW, cUV, cQR, cPR, cPQ;
    int thrusters_on;
};

template <typename T> struct Trig6 { T sph, cph, sth, cth, sps, cps; };

// fp32 sin/cos without libm's branches: Cody-Waite reduction by pi/2 (three
// FMAs, exact for |x| < 2^16) + the minimax polynomials of the Cephes sinf /
// cosf kernels (|r| <= pi/4, ~1 ulp) + branch-free quadrant fix-up.  Angles are
// wrapped to [0, 2 pi) every env step, so |x| stays far below 2^16; the step
// kernel flags any environment whose unwrapped angle leaves that range in its
// non-finite/out-of-range counter instead of paying a branch per evaluation.
#define MVRL_SINCOS_F32_MAX_ARG 65536.0f

__device__ __forceinline__ void sincos_f32(float x, float* sn, float* cs) {
    const float j = fmaf(x, 0.636619772367581343f, 12582912.0f);   // round(x * 2/pi) in the low mantissa bits
    const int q = __float_as_int(j);
    const float k = j - 12582912.0f;
    float r = fmaf(k, -1.5703125f, x);
    r = fmaf(k, -4.837512969970703125e-4f, r);
    r = fmaf(k, -7.549789954891882e-8f, r);
    const float z = r * r;
    float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(ps, z, -1.6666654611e-1f);
    const float s0 = fmaf(ps * z, r, r);
    float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(pc, z, 4.166664568298827e-2f);
    const float c0 = fmaf(pc * z, z, fmaf(z, -0.5f, 1.0f));
    const bool swap = (q & 1) != 0;
    const float s1 = swap ? c0 : s0;
    const float c1 = swap ? s0 : c0;
    *sn = __int_as_float(__float_as_int(s1) ^ ((q & 2) << 30));
    *cs = __int_as_float(__float_as_int(c1) ^ (((q + 1) & 2) << 30));
}

template <typename T, bool FAST>
__device__ __forceinline__ void sincos_t(T x, T* s, T* c) {
    if constexpr (sizeof(T) == 4) {
        if constexpr (FAST) { *s = __sinf(x); *c = __cosf(x); }   // MUFU.SIN / MUFU.COS
        else sincos_f32(x, s, c);
    } else {
        Real<T>::sincos(x, s, c);
    }
}

template <typename T, bool FAST>
__device__ __forceinline__ Trig6<T> trig6(T phi, T theta, T psi) {
    Trig6<T> g;
    sincos_t<T, FAST>(phi, &g.sph, &g.cph);
    sincos_t<T, FAST>(theta, &g.sth, &g.cth);
    sincos_t<T, FAST>(psi, &g.sps, &g.cps);
    return g;
}

// 6DoF.py:271-275 + 233-236: saturate, deadband, static thrust.
template <typename T> __device__ __forceinline__ T thruster_force(const Rov6Dev<T>& P, T rpm) {
    T r = tmax(-P.rpm_max, tmin(P.rpm_max, rpm));
    if (tabs(r) < P.rpm_db) r = T(0);
    return P.thrust_k * r * tabs(r);
}

// H = sum_i F_i A[:, i]  (6DoF.py:278-282).  SP: the reference's default
// allocation pattern - horizontal thrusters 0-3 produce no heave force,
// vertical thrusters 4-7 produce heave, roll and pitch only.
template <typename T, bool SP>
__device__ __forceinline__ void thrust_wrench(const Rov6Dev<T>& P, const T (&F)[8], T (&H)[6]) {
    if (!P.thrusters_on) {
#pragma unroll
        for (int k = 0; k < 6; ++k) H[k] = T(0);
        return;
    }
    if constexpr (SP) {
        H[0] = P.A[0][0] * F[0] + P.A[0][1] * F[1] + P.A[0][2] * F[2] + P.A[0][3] * F[3];
        H[1] = P.A[1][0] * F[0] + P.A[1][1] * F[1] + P.A[1][2] * F[2] + P.A[1][3] * F[3];
        H[2] = P.A[2][4] * F[4] + P.A[2][5] * F[5] + P.A[2][6] * F[6] + P.A[2][7] * F[7];
#pragma unroll
        for (int k = 3; k < 5; ++k) {
            T s = T(0);
#pragma unroll
            for (int i = 0; i < 8; ++i) s += P.A[k][i] * F[i];
            H[k] = s;
        }
        H[5] = P.A[5][0] * F[0] + P.A[5][1] * F[1] + P.A[5][2] * F[2] + P.A[5][3] * F[3];
    } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            T s = T(0);
#pragma unroll
            for (int i = 0; i < 8; ++i) s += P.A[k][i] * F[i];
            H[k] = s;
        }
    }
}

// 6DoF.py:220-231 (+ 238-248): earth-frame demand -> body frame (forces AND
// moments rotated like vectors through the intrinsic-XYZ axes) -> Ainv -> rpm.
// Returns the allocated per-thruster force demand c_i (newtons); rpm_i =
// sign(c_i) sqrt(|c_i| / (rho D^4 Kt)) 60.
template <typename T, bool SP>
__device__ __forceinline__ void allocate_demand(const Rov6Dev<T>& P, const Trig6<T>& g, const T (&gcf)[6], T (&c)[8]) {
    const T ix = g.cth * g.cps, iy = g.cph * g.sps + g.sph * g.sth * g.cps, iz = g.sph * g.sps - g.cph * g.sth * g.cps;
    const T jx = -g.cth * g.sps, jy = g.cph * g.cps - g.sph * g.sth * g.sps, jz = g.sph * g.cps + g.cph * g.sth * g.sps;
    const T kx = g.sth, ky = -g.sph * g.cth, kz = g.cph * g.cth;
    T b[6];
    b[0] = gcf[0] * ix + gcf[1] * iy + gcf[2] * iz;
    b[1] = gcf[0] * jx + gcf[1] * jy + gcf[2] * jz;
    b[2] = gcf[0] * kx + gcf[1] * ky + gcf[2] * kz;
    b[3] = gcf[3] * ix + gcf[4] * iy + gcf[5] * iz;
    b[4] = gcf[3] * jx + gcf[4] * jy + gcf[5] * jz;
    b[5] = gcf[3] * kx + gcf[4] * ky + gcf[5] * kz;
    if constexpr (SP) {
        // pinv of the default A: horizontals see (X, Y, N) only, verticals see (Z, K, M) only
#pragma unroll
        for (int i = 0; i < 4; ++i) c[i] = P.Ainv[i][0] * b[0] + P.Ainv[i][1] * b[1] + P.Ainv[i][5] * b[5];
#pragma unroll
        for (int i = 4; i < 8; ++i) {
            T s = T(0);
#pragma unroll
            for (int k = 0; k < 5; ++k) s += P.Ainv[i][k] * b[k];
            c[i] = s;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            T s = T(0);
#pragma unroll
            for (int k = 0; k < 6; ++k) s += P.Ainv[i][k] * b[k];
            c[i] = s;
        }
    }
}

template <typename T> __device__ __forceinline__ T demand_to_rpm(const Rov6Dev<T>& P, T c) {
    return sgn(c) * Real<T>::sqrt(tabs(c) * P.inv_thrust_coef) * T(60);
}

// Thruster force from an allocated demand.  EXACT follows the reference
// literally (rpm via sqrt, limit, thrust law).  Otherwise the algebraic
// identity F(rpm(c)) = c is used with the limits moved to force space
// (|c| > F(rpm_max) saturates, |c| < F(deadband) is zero): same result up to
// rounding, no sqrt.
template <typename T, bool EXACT>
__device__ __forceinline__ T demand_to_force(const Rov6Dev<T>& P, T c) {
    if constexpr (EXACT) {
        return thruster_force(P, demand_to_rpm(P, c));
    } else {
        T a = tabs(c);
        T f = tmin(a, P.f_max);
        f = a < P.f_db ? T(0) : f;
        return copysign(f, c);
    }
}

// RHS = -Crb v - (Ca + D) v - G + H ; acc = Minv RHS   (6DoF.py:284-396, 428)
// comp (nullable): -Crb v, -Ca v, -D v, G as in forceModel(retComp=True).
template <typename T, bool SP>
__device__ __forceinline__ void body_accel(const Rov6Dev<T>& P, const Trig6<T>& g, const T (&nu)[6], const T (&H)[6],
                                           T (&acc)[6], T (&rhs)[6], T* comp = nullptr, long comp_ld = 0) {
    const T u = nu[0], v = nu[1], w = nu[2], p = nu[3], q = nu[4], r = nu[5];
    if constexpr (SP) {
        if (comp == nullptr) {
            // Default sparsity, no component dump: Crb v + Ca v folded analytically
            // (xg = yg = 0, diagonal inertia; the m w v - m v w pairs of 6DoF.py:313-331
            // cancel identically) and every product accumulated straight into RHS by FMA.
            const T pr = p * r, qr = q * r;
            rhs[0] = fmaf_t(fmaf_t(P.Xuu, tabs(u), P.Xu), u, fmaf_t(-P.mzg, pr, fmaf_t(-P.mZ, w * q, fmaf_t(P.mY, v * r, H[0]))));
            rhs[1] = fmaf_t(fmaf_t(P.Yvv, tabs(v), P.Yv), v, fmaf_t(-P.mzg, qr, fmaf_t(-P.mX, u * r, fmaf_t(P.mZ, w * p, H[1]))));
            rhs[2] = fmaf_t(fmaf_t(P.Zww, tabs(w), P.Zw), w, fmaf_t(P.mzg, fmaf_t(q, q, p * p), fmaf_t(P.mX, u * q, fmaf_t(-P.mY, v * p, H[2]))));
            rhs[3] = fmaf_t(fmaf_t(P.Kpp, tabs(p), P.Kp), p,
                            fmaf_t(-P.gz, g.cth * g.sph, fmaf_t(-P.cQR, qr, fmaf_t(-P.cVW, v * w, fmaf_t(-P.mzg, fmaf_t(-r, u, p * w), H[3])))));
            rhs[4] = fmaf_t(fmaf_t(P.Mqq, tabs(q), P.Mq), q,
                            fmaf_t(P.Mww * tabs(w), w,
                                   fmaf_t(-P.gz, g.sth, fmaf_t(-P.cPR, pr, fmaf_t(-P.cUW, u * w, fmaf_t(-P.mzg, fmaf_t(-r, v, q * w), H[4]))))));
            rhs[5] = fmaf_t(fmaf_t(P.Nrr, tabs(r), P.Nr), r, fmaf_t(-P.cPQ, p * q, fmaf_t(-P.cUV, u * v, H[5])));
            acc[0] = fmaf_t(P.Minv[0][4], rhs[4], P.Minv[0][0] * rhs[0]);
            acc[1] = fmaf_t(P.Minv[1][3], rhs[3], P.Minv[1][1] * rhs[1]);
            acc[2] = P.Minv[2][2] * rhs[2];
            acc[3] = fmaf_t(P.Minv[3][1], rhs[1], P.Minv[3][3] * rhs[3]);
            acc[4] = fmaf_t(P.Minv[4][0], rhs[0], P.Minv[4][4] * rhs[4]);
            acc[5] = P.Minv[5][5] * rhs[5];
            return;
        }
    }
    const T m = P.m;
    T crb[6], ca[6], dv[6], G[6];

    // ---- Crb(v) v, 6DoF.py:303-332
    if constexpr (SP) {  // xg = yg = 0, diagonal inertia
        const T zr = P.zg * r, zp = P.zg * p, zq = P.zg * q;
        const T a1 = m * zr, b1 = m * w, b2 = m * zr;
        const T c1 = m * (zp - v), c2 = m * (zq + u);
        const T mw = m * w, mv = m * v, mu = m * u;
        crb[0] = a1 * p + mw * q - mv * r;
        crb[1] = -b1 * p + b2 * q + mu * r;
        crb[2] = -c1 * p - c2 * q;
        const T i1 = P.Izz * r, i2 = -P.Iyy * q, i3 = P.Ixx * p;
        crb[3] = -a1 * u + b1 * v + c1 * w + i1 * q + i2 * r;
        crb[4] = -mw * u - b2 * v + c2 * w - i1 * p + i3 * r;
        crb[5] = mv * u - mu * v - i2 * p - i3 * q;
    } else {
        const T a1 = m * (P.yg * q + P.zg * r), a2 = m * (P.xg * q - w), a3 = m * (P.xg * r + v);
        const T b1 = m * (P.yg * p + w), b2 = m * (P.zg * r + P.xg * p), b3 = m * (P.yg * r - u);
        const T c1 = m * (P.zg * p - v), c2 = m * (P.zg * q + u), c3 = m * (P.xg * p + P.yg * q);
        const T i1 = -P.Iyz * q - P.Ixz * p + P.Izz * r;
        const T i2 = P.Iyz * r + P.Ixy * p - P.Iyy * q;
        const T i3 = -P.Ixz * r - P.Ixy * q + P.Ixx * p;
        crb[0] = a1 * p - a2 * q - a3 * r;
        crb[1] = -b1 * p + b2 * q - b3 * r;
        crb[2] = -c1 * p - c2 * q + c3 * r;
        crb[3] = -a1 * u + b1 * v + c1 * w + i1 * q + i2 * r;
        crb[4] = a2 * u - b2 * v + c2 * w - i1 * p + i3 * r;
        crb[5] = a3 * u + b3 * v - c3 * w - i2 * p - i3 * q;
    }

    // ---- Ca(v) v, 6DoF.py:334-341 (Zwdot here although Ma carries Zvdot)
    {
        const T xu = P.Xud * u, yv = P.Yvd * v, zw = P.Zwd * w, kp = P.Kpd * p, mq = P.Mqd * q, nr = P.Nrd * r;
        ca[0] = -zw * q + yv * r;
        ca[1] = zw * p - xu * r;
        ca[2] = -yv * p + xu * q;
        ca[3] = -zw * v + yv * w - nr * q + mq * r;
        ca[4] = zw * u - xu * w + nr * p - kp * r;
        ca[5] = -yv * u + xu * v - mq * p + kp * q;
    }

    // ---- -D(v) v, 6DoF.py:345-370: D = -(Dl + Dq |v|)
    {
        const T au = tabs(u), av = tabs(v), aw = tabs(w), ap = tabs(p), aq = tabs(q), ar = tabs(r);
        dv[0] = (P.Xu + P.Xuu * au) * u;
        if constexpr (SP) {  // only Mww couples
            dv[1] = (P.Yv + P.Yvv * av) * v;
            dv[2] = (P.Zw + P.Zww * aw) * w;
            dv[3] = (P.Kp + P.Kpp * ap) * p;
            dv[4] = (P.Mww * aw) * w + (P.Mq + P.Mqq * aq) * q;
            dv[5] = (P.Nr + P.Nrr * ar) * r;
        } else {
            dv[1] = (P.Yv + P.Yvv * av) * v + (P.Yp + P.Ypp * ap) * p + (P.Yr + P.Yrr * ar) * r;
            dv[2] = (P.Zw + P.Zww * aw) * w + (P.Zq + P.Zqq * aq) * q;
            dv[3] = (P.Kv + P.Kvv * av) * v + (P.Kp + P.Kpp * ap) * p + (P.Kr + P.Krr * ar) * r;
            dv[4] = (P.Mw + P.Mww * aw) * w + (P.Mq + P.Mqq * aq) * q;
            dv[5] = (P.Nv + P.Nvv * av) * v + (P.Np + P.Npp * ap) * p + (P.Nr + P.Nrr * ar) * r;
        }
    }

    // ---- G(phi, theta), 6DoF.py:374-388
    if constexpr (SP) {  // neutrally buoyant, CG/CB on the z axis
        G[0] = T(0); G[1] = T(0); G[2] = T(0);
        G[3] = P.gz * g.cth * g.sph;
        G[4] = P.gz * g.sth;
        G[5] = T(0);
    } else {
        G[0] = P.WmB * g.sth;
        G[1] = -P.WmB * g.cth * g.sph;
        G[2] = -P.WmB * g.cth * g.cph;
        G[3] = -P.gy * g.cth * g.cph + P.gz * g.cth * g.sph;
        G[4] = P.gz * g.sth + P.gx * g.cth * g.cph;
        G[5] = -P.gx * g.cth * g.sph - P.gy * g.sth;
    }

#pragma unroll
    for (int k = 0; k < 6; ++k) rhs[k] = -crb[k] - ca[k] + dv[k] - G[k] + H[k];

    if (comp != nullptr) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            comp[(0 * 6 + k) * comp_ld] = -crb[k];
            comp[(1 * 6 + k) * comp_ld] = -ca[k];
            comp[(2 * 6 + k) * comp_ld] = dv[k];
            comp[(3 * 6 + k) * comp_ld] = G[k];
            comp[(4 * 6 + k) * comp_ld] = H[k];
        }
    }

    // ---- acc = M^-1 RHS (M is state independent, 6DoF.py:286-299, 428)
    if constexpr (SP) {  // couplings (0,4) and (1,3) only
        acc[0] = P.Minv[0][0] * rhs[0] + P.Minv[0][4] * rhs[4];
        acc[1] = P.Minv[1][1] * rhs[1] + P.Minv[1][3] * rhs[3];
        acc[2] = P.Minv[2][2] * rhs[2];
        acc[3] = P.Minv[3][1] * rhs[1] + P.Minv[3][3] * rhs[3];
        acc[4] = P.Minv[4][0] * rhs[0] + P.Minv[4][4] * rhs[4];
        acc[5] = P.Minv[5][5] * rhs[5];
    } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            T s = T(0);
#pragma unroll
            for (int j = 0; j < 6; ++j) s += P.Minv[k][j] * rhs[j];
            acc[k] = s;
        }
    }
}

// eta_dot = J(phi, theta, psi) v  (resources.py:115-141, 6DoF.py:432-435).
// J1[0][2] keeps the reference's sin(phi) in its second term; cos(theta) in
// J2 is clamped exactly as resources.py:116-120.
template <typename T, bool FAST>
__device__ __forceinline__ void kinematics6(const Trig6<T>& g, const T (&nu)[6], T (&ed)[6]) {
    const T u = nu[0], v = nu[1], w = nu[2], p = nu[3], q = nu[4], r = nu[5];
    // J1 v factored through psi: x' = c(psi) A1 - s(psi) B, y' = s(psi) A2 + c(psi) B with
    // A1 = c(th) u + s(th)s(ph) (v + w)   <- the reference's J1[0][2] (sin(phi), not cos(phi))
    // A2 = c(th) u + s(th)s(ph) v + s(th)c(ph) w,  B = c(ph) v - s(ph) w
    const T ss = g.sth * g.sph, sc = g.sth * g.cph, cu = g.cth * u;
    const T A1 = fmaf_t(ss, v + w, cu);
    const T A2 = fmaf_t(sc, w, fmaf_t(ss, v, cu));
    const T B = fmaf_t(g.cph, v, -(g.sph * w));
    ed[0] = fmaf_t(g.cps, A1, -(g.sps * B));
    ed[1] = fmaf_t(g.sps, A2, g.cps * B);
    ed[2] = fmaf_t(g.cth, fmaf_t(g.sph, v, g.cph * w), -(g.sth * u));
    // resources.py:116-120, branch-free: |c| < 1e-12 -> 1e-6, |c| < 1e-6 -> 1e-6 sign(c)
    const T ad = tabs(g.cth);
    const T tiny = ad < T(1e-12) ? T(1e-6) : copysign(T(1e-6), g.cth);
    const T den = ad < T(1e-6) ? tiny : g.cth;
    T inv;
    if constexpr (sizeof(T) == 4) {
        // |den| >= 1e-6 after the clamp: MUFU.RCP needs no special-case path; one Newton step
        // brings it to <= 1 ulp in the accurate mode.
        inv = __fdividef(1.0f, den);
        if constexpr (!FAST) inv = fmaf(inv, fmaf(-den, inv, 1.0f), inv);
    } else {
        inv = T(1) / den;
    }
    const T a = fmaf_t(g.sph, q, g.cph * r);  // shared by rows 0 and 2 of J2
    ed[5] = inv * a;
    ed[3] = fmaf_t(g.sth, ed[5], p);
    ed[4] = fmaf_t(g.cph, q, -(g.sph * r));
}

// BlueROV2Heavy6DoF_PID_controller.computeControlForces, 6DoF.py:43-73.
// dtc = t - tOld.  e_old[0] = NaN encodes eOld is None.
template <typename T>
__device__ __forceinline__ void pid6(const Rov6Dev<T>& P, T (&e_old)[6], T (&e_int)[6], const T (&sp)[6],
                                     const T (&pose)[6], T dtc, T (&out)[6]) {
    T e[6];
#pragma unroll
    for (int k = 0; k < 5; ++k) e[k] = sp[k] - pose[k];   // roll/pitch: raw differences (6DoF.py:59-60)
    e[5] = angle_error(sp[5], pose[5]);
    const bool none = e_old[0] != e_old[0];
    const T inv_dt = T(1) / tmax(T(1e-9), dtc);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const T eo = none ? e[k] : e_old[k];
        const T dedt = (e[k] - eo) * inv_dt;
        T ei = e_int[k] + T(0.5) * (eo + e[k]) * dtc;
        if (tabs(e[k]) > P.pWind[k]) ei = T(0);
        T cvl = P.pKp[k] * e[k] + P.pKd[k] * dedt + P.pKi[k] * ei;
        out[k] = tmax(-P.pMax[k], tmin(P.pMax[k], cvl));
        e_int[k] = ei;
        e_old[k] = e[k];
    }
}

}  // namespace mvrl
