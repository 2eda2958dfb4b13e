/*
 * CPU ORACLE (plain C) for the batched BlueROV2 6DoF env-step path.
 *
 * TEST INFRASTRUCTURE ONLY - never linked into, loaded by, or called from the
 * product (marinevehiclereinforcementlearning_b200/).  Users: tests/, the
 * cpu_baseline / --impl reference legs of bench.py, __graft_entry__.smoke().
 *
 * It restates, in scalar double-precision C, the algorithm of the reference's
 * dynamicsModel_BlueROV2_Heavy_6DoF.py (cited per function below as 6DoF.py:
 * line) and resources.py, one environment at a time exactly like the
 * reference, with OpenMP only across environments.  Compiled with
 * -ffp-contract=off so that, like numpy, no FMA contraction happens.
 *
 * Parity pin: checked in tests/test_c_oracle.py against the golden vectors
 * written by running the unmodified reference (tests/golden/golden_*.npz) and
 * against the numpy oracle; the reference itself ships no tests.
 *
 * Layout: array-of-structs, row i = environment i (numpy's natural layout).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TWO_PI (2.0 * M_PI)

typedef struct OrcRov6Params {
    /* 6DoF.py:83-218 */
    double rho_f, m, Length, dispVol;
    double CG[3], CB[3], I[9];
    double Xudot, Yvdot, Zwdot, Kpdot, Mqdot, Nrdot, Zvdot;
    double Xu, Yv, Yp, Yr, Zw, Zq, Kv, Kp, Kr, Mw, Mq, Nv, Np, Nr;
    double Xuu, Yvv, Ypp, Yrr, Zww, Zqq, Kvv, Kpp, Krr, Mww, Mqq, Nvv, Npp, Nrr;
    double D_thruster, Kt_thruster;
    double A[48];    /* 6x8 row-major, resources.py:19-35 (filled by the caller with numpy's result) */
    double Ainv[48]; /* 8x6 row-major */
} OrcRov6Params;

/* Python float % (CPython float_rem) */
static double pymod(double a, double b) {
    double r = fmod(a, b);
    if (r != 0.0) { if ((b < 0.0) != (r < 0.0)) r += b; }
    else r = copysign(0.0, b);
    return r;
}

/* resources.py:75-95 */
double orc_angle_error(double psi_d, double psi) {
    double a = pymod(psi_d - psi, TWO_PI);
    double b = pymod(psi - psi_d, TWO_PI);
    return a < b ? a : -b;
}

static double sign_(double x) { return (x > 0.0) - (x < 0.0); }

/* resources.py:115-141 applied to vel: eta_dot = J vel */
static void eta_dot(const double* ang, const double* vel, double* out) {
    const double phi = ang[0], theta = ang[1], psi = ang[2];
    double c = cos(theta);
    if (fabs(c) < 1e-12) c = 1e-6;
    else if (fabs(c) < 1e-6) c = 1e-6 * sign_(c);
    const double J1[3][3] = {
        {cos(psi) * cos(theta), -sin(psi) * cos(phi) + cos(psi) * sin(theta) * sin(phi), sin(psi) * sin(phi) + cos(psi) * sin(theta) * sin(phi)},
        {sin(psi) * cos(theta), cos(psi) * cos(phi) + sin(psi) * sin(theta) * sin(phi), -cos(psi) * sin(phi) + sin(psi) * sin(theta) * cos(phi)},
        {-sin(theta), cos(theta) * sin(phi), cos(theta) * cos(phi)}};
    const double J2[3][3] = {
        {1., sin(phi) * sin(theta) / c, cos(phi) * sin(theta) / c},
        {0., cos(phi), -sin(phi)},
        {0., sin(phi) / c, cos(phi) / c}};
    for (int i = 0; i < 3; ++i) {
        double s = 0., t = 0.;
        for (int j = 0; j < 3; ++j) { s += J1[i][j] * vel[j]; t += J2[i][j] * vel[3 + j]; }
        out[i] = s; out[3 + i] = t;
    }
}

/* 6DoF.py:238-242: rows iHat, jHat, kHat = columns of Rx(phi) Ry(theta) Rz(psi) transposed */
static void body_axes(const double* ang, double ax[3][3]) {
    const double sph = sin(ang[0]), cph = cos(ang[0]), sth = sin(ang[1]), cth = cos(ang[1]), sps = sin(ang[2]), cps = cos(ang[2]);
    ax[0][0] = cth * cps; ax[0][1] = cph * sps + sph * sth * cps; ax[0][2] = sph * sps - cph * sth * cps;
    ax[1][0] = -cth * sps; ax[1][1] = cph * cps - sph * sth * sps; ax[1][2] = sph * cps + cph * sth * sps;
    ax[2][0] = sth; ax[2][1] = -sph * cth; ax[2][2] = cph * cth;
}

/* 6DoF.py:220-231 */
static void allocate_thrust(const OrcRov6Params* p, const double* ang, const double* gcf, double* cv) {
    double ax[3][3], b[6];
    body_axes(ang, ax);
    for (int i = 0; i < 3; ++i) {
        b[i] = gcf[0] * ax[i][0] + gcf[1] * ax[i][1] + gcf[2] * ax[i][2];
        b[3 + i] = gcf[3] * ax[i][0] + gcf[4] * ax[i][1] + gcf[5] * ax[i][2];
    }
    const double k = p->rho_f * pow(p->D_thruster, 4.) * p->Kt_thruster;
    for (int i = 0; i < 8; ++i) {
        double s = 0.;
        for (int j = 0; j < 6; ++j) s += p->Ainv[i * 6 + j] * b[j];
        cv[i] = sign_(s) * sqrt(fabs(s) / k) * 60.;
    }
}

/* NOT part of the algorithm - conditioning diagnostic for the parity tests: the smallest relative distance of a thruster
 * demand from the dead-band edge (| |rpm| - 300 | / 300) seen by the calling thread since it was last reset.  The thrust
 * jumps from 0 to 0.29 N there, so an environment that comes within rounding distance of the edge cannot agree
 * between two precisions. */
static _Thread_local double g_dbmargin = 1e300;

/* 6DoF.py:271-275 */
static double limit_rpm(double x) {
    double r = fmax(-3500., fmin(3500., x));
    const double m = fabs(fabs(r) - 300.) / 300.;
    if (m < g_dbmargin) g_dbmargin = m;
    if (fabs(r) < 300) r = 0.;
    return r;
}

/* 6DoF.py:233-236 */
static double thruster_model(const OrcRov6Params* p, double rpm) {
    return p->rho_f * pow(rpm / 60., 2.) * sign_(rpm) * pow(p->D_thruster, 4.) * p->Kt_thruster;
}

/* 6DoF.py:286-299 */
static void mass_matrix(const OrcRov6Params* p, double M[6][6]) {
    const double m = p->m, xg = p->CG[0], yg = p->CG[1], zg = p->CG[2];
    const double Mrb[6][6] = {
        {m, 0., 0., 0., m * zg, -m * yg}, {0., m, 0., -m * zg, 0., m * xg}, {0., 0., m, m * yg, -m * xg, 0.},
        {0., -m * zg, m * yg, p->I[0], p->I[1], p->I[2]}, {m * zg, 0., -m * xg, p->I[3], p->I[4], p->I[5]},
        {-m * yg, m * xg, 0., p->I[6], p->I[7], p->I[8]}};
    const double Ma[6] = {-p->Xudot, -p->Yvdot, -p->Zvdot, -p->Kpdot, -p->Mqdot, -p->Nrdot};
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) M[i][j] = Mrb[i][j] + (i == j ? Ma[i] : 0.);
}

/* 6DoF.py:253-404: RHS (and M) of the force model */
static void force_model(const OrcRov6Params* p, const double* ang, const double* vel, const double* rpms, double* RHS) {
    const double phi = ang[0], theta = ang[1];
    const double u = vel[0], v = vel[1], w = vel[2], pp = vel[3], q = vel[4], r = vel[5];
    const double m = p->m, xg = p->CG[0], yg = p->CG[1], zg = p->CG[2];
    const double* I = p->I;
    double H[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 8; ++i) {
        const double F = thruster_model(p, limit_rpm(rpms[i]));
        for (int k = 0; k < 6; ++k) H[k] += F * p->A[k * 8 + i];
    }
    const double Crb[6][6] = {
        {0., 0., 0., m * (yg * q + zg * r), -m * (xg * q - w), -m * (xg * r + v)},
        {0., 0., 0., -m * (yg * pp + w), m * (zg * r + xg * pp), -m * (yg * r - u)},
        {0., 0., 0., -m * (zg * pp - v), -m * (zg * q + u), m * (xg * pp + yg * q)},
        {-m * (yg * q + zg * r), m * (yg * pp + w), m * (zg * pp - v), 0., -I[5] * q - I[2] * pp + I[8] * r, I[5] * r + I[1] * pp - I[4] * q},
        {m * (xg * q - w), -m * (zg * r + xg * pp), m * (zg * q + u), I[5] * q + I[2] * pp - I[8] * r, 0., -I[2] * r - I[1] * q + I[0] * pp},
        {m * (xg * r + v), m * (yg * r - u), -m * (xg * pp + yg * q), -I[5] * r - I[1] * pp + I[4] * q, I[2] * r + I[1] * q - I[0] * pp, 0.}};
    const double Ca[6][6] = {
        {0., 0., 0., 0., -p->Zwdot * w, p->Yvdot * v},
        {0., 0., 0., p->Zwdot * w, 0., -p->Xudot * u},
        {0., 0., 0., -p->Yvdot * v, p->Xudot * u, 0.},
        {0., -p->Zwdot * w, p->Yvdot * v, 0., -p->Nrdot * r, p->Mqdot * q},
        {p->Zwdot * w, 0., -p->Xudot * u, p->Nrdot * r, 0., -p->Kpdot * pp},
        {-p->Yvdot * v, p->Xudot * u, 0., -p->Mqdot * q, p->Kpdot * pp, 0.}};
    const double Dl[6][6] = {
        {p->Xu, 0., 0., 0., 0., 0.}, {0., p->Yv, 0., p->Yp, 0., p->Yr}, {0., 0., p->Zw, 0., p->Zq, 0.},
        {0., p->Kv, 0., p->Kp, 0., p->Kr}, {0., 0., p->Mw, 0., p->Mq, 0.}, {0., p->Nv, 0., p->Np, 0., p->Nr}};
    const double Dq[6][6] = {
        {p->Xuu, 0., 0., 0., 0., 0.}, {0., p->Yvv, 0., p->Ypp, 0., p->Yrr}, {0., 0., p->Zww, 0., p->Zqq, 0.},
        {0., p->Kvv, 0., p->Kpp, 0., p->Krr}, {0., 0., p->Mww, 0., p->Mqq, 0.}, {0., p->Nvv, 0., p->Npp, 0., p->Nrr}};
    const double av[6] = {fabs(u), fabs(v), fabs(w), fabs(pp), fabs(q), fabs(r)};
    const double W = p->m * 9.81, B = p->dispVol * p->rho_f * 9.81;
    const double xb = p->CB[0], yb = p->CB[1], zb = p->CB[2];
    const double G[6] = {
        (W - B) * sin(theta), -(W - B) * cos(theta) * sin(phi), -(W - B) * cos(theta) * cos(phi),
        -(yg * W - yb * B) * cos(theta) * cos(phi) + (zg * W - zb * B) * cos(theta) * sin(phi),
        (zg * W - zb * B) * sin(theta) + (xg * W - xb * B) * cos(theta) * cos(phi),
        -(xg * W - xb * B) * cos(theta) * sin(phi) - (yg * W - yb * B) * sin(theta)};
    for (int i = 0; i < 6; ++i) {
        double crb = 0., cad = 0.;
        for (int j = 0; j < 6; ++j) {
            crb += Crb[i][j] * vel[j];
            const double D = -1. * Dl[i][j] + -1. * Dq[i][j] * av[j];
            cad += (Ca[i][j] + D) * vel[j];
        }
        RHS[i] = -crb - cad - G[i] + H[i];
    }
}

/* numpy.linalg.solve for one 6x6 system: LU with partial pivoting (LAPACK dgesv) */
static void solve6(double A[6][6], double* b) {
    for (int c = 0; c < 6; ++c) {
        int piv = c;
        for (int r = c + 1; r < 6; ++r) if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
        if (piv != c) {
            for (int j = 0; j < 6; ++j) { double t = A[c][j]; A[c][j] = A[piv][j]; A[piv][j] = t; }
            double t = b[c]; b[c] = b[piv]; b[piv] = t;
        }
        for (int r = c + 1; r < 6; ++r) {
            const double f = A[r][c] / A[c][c];
            if (f != 0.) { for (int j = c; j < 6; ++j) A[r][j] -= f * A[c][j]; b[r] -= f * b[c]; }
        }
    }
    for (int r = 5; r >= 0; --r) {
        double s = b[r];
        for (int j = r + 1; j < 6; ++j) s -= A[r][j] * b[j];
        b[r] = s / A[r][r];
    }
}

/* controller state, 6DoF.py:37-41 */
/* margin is NOT part of the algorithm: a conditioning diagnostic for the parity tests.  It keeps the smallest non-zero
 * |e - eOld| seen by a call with t - tOld < 1e-9, where dedt = (e - eOld) / 1e-9 turns the sign of that difference
 * into a saturated demand (RK4 stages 1 and 3) - the distance of the environment from a sign flip. */
typedef struct OrcPid6 { double eOld[6], eInt[6], tOld, margin; int has_old; } OrcPid6;

static const double PID_WINDUP[6] = {2., 2., 2., 90. / 180. * M_PI, 90. / 180. * M_PI, 90. / 180. * M_PI};
static const double PID_MAX[6] = {50., 50., 50., 1., 1., 2.};
static const double PID_KP[6] = {25., 25., 25., 10., 10., 1.};
static const double PID_KI[6] = {2., 2., 2., 0.1, 0.1, 0.2};
static const double PID_KD[6] = {20., 20., 20., 5., 5., 0.65};

/* 6DoF.py:43-73 */
static void pid_control(OrcPid6* c, const double* sp, const double* pose, double t, double* out) {
    double e[6];
    for (int k = 0; k < 5; ++k) e[k] = sp[k] - pose[k];
    e[5] = orc_angle_error(sp[5], pose[5]);
    if (!c->has_old) { memcpy(c->eOld, e, sizeof(e)); c->has_old = 1; }
    const double dtc = t - c->tOld;
    if (dtc < 1e-9) {
        for (int k = 0; k < 6; ++k) {
            const double d = fabs(e[k] - c->eOld[k]);
            if (d > 0. && d < c->margin) c->margin = d;
        }
    }
    for (int k = 0; k < 6; ++k) {
        const double dedt = (e[k] - c->eOld[k]) / fmax(1e-9, dtc);
        c->eInt[k] += 0.5 * (c->eOld[k] + e[k]) * dtc;
        if (fabs(e[k]) > PID_WINDUP[k]) c->eInt[k] = 0.;
        double u = PID_KP[k] * e[k] + PID_KD[k] * dedt + PID_KI[k] * c->eInt[k];
        out[k] = fmax(-PID_MAX[k], fmin(PID_MAX[k], u));
        c->eOld[k] = e[k];
    }
    c->tOld = t;
}

/* 6DoF.py:406-442.  mode 0: act = rpm[8]; 1: act = earth-frame forces[6]; 2: PID (sp, ctrl). */
static void derivs6(const OrcRov6Params* p, int mode, double t, const double* s, const double* act, const double* sp,
                    OrcPid6* ctrl, double* out, double* gcf_out, double* cv_out) {
    double gcf[6] = {0, 0, 0, 0, 0, 0}, cv[8], M[6][6], RHS[6];
    if (mode == 0) memcpy(cv, act, sizeof(cv));
    else {
        if (mode == 1) memcpy(gcf, act, sizeof(gcf));
        else pid_control(ctrl, sp, s, t, gcf);
        allocate_thrust(p, s + 3, gcf, cv);
    }
    force_model(p, s + 3, s + 6, cv, RHS);
    mass_matrix(p, M);
    solve6(M, RHS);
    eta_dot(s + 3, s + 6, out);
    memcpy(out + 6, RHS, 6 * sizeof(double));
    if (gcf_out) memcpy(gcf_out, gcf, sizeof(gcf));
    if (cv_out) memcpy(cv_out, cv, sizeof(cv));
}

void orc_rov6_derivs(const OrcRov6Params* p, int mode, long n, const double* state, const double* act, const double* t,
                     const double* sp, OrcPid6* ctrl, double* out, double* gcf, double* cv) {
    const int na = mode == 0 ? 8 : 6;
    for (long i = 0; i < n; ++i)
        derivs6(p, mode, t ? t[i] : 0., state + 12 * i, act ? act + na * i : 0, sp ? sp + 6 * i : 0, ctrl ? ctrl + i : 0,
                out + 12 * i, gcf ? gcf + 6 * i : 0, cv ? cv + 8 * i : 0);
}

/* ---- Philox4x32-10 (new in the build; mirrors oracle_np.philox_uniform) ---- */
static void philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int i = 0; i < 10; ++i) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
static void draw_uniform(uint64_t seed, uint64_t env, uint32_t episode, int nvals, double* u) {
    for (int blk = 0; blk * 4 < nvals; ++blk) {
        uint32_t c[4] = {(uint32_t)env, (uint32_t)(env >> 32), episode, (uint32_t)blk};
        philox(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        for (int j = 0; j < 4 && blk * 4 + j < nvals; ++j) u[blk * 4 + j] = (double)(c[j] >> 8) * (1.0 / 16777216.0);
    }
}

/* 6DoF.py:467-483 */
static void observe(const OrcRov6Params* p, const double* s, const double* path, const double* sp, double* obs) {
    const double L3 = p->Length * 3.;
    for (int k = 0; k < 3; ++k) {
        obs[k] = (path[k] - s[k]) / L3;
        obs[3 + k] = (path[3 + k] - s[k]) / L3;
        obs[6 + k] = orc_angle_error(sp[3 + k], s[3 + k]) / (45. / 180. * M_PI);
    }
    for (int k = 0; k < 9; ++k) obs[k] = fmax(-1., fmin(1., obs[k]));
}

typedef struct OrcRov6Env {
    int mode, n_sub, max_steps, auto_reset, fixed_sp, threads;
    double dt;
    uint64_t seed, env_id0;
} OrcRov6Env;

/* One env step for n environments, 6DoF.py:531-594 with the integrator fixed to RK4 x n_sub.
 * state [n][12], action [n][8|6], setpoint [n][6], path [n][6], ctrl [n], istep [n], time [n],
 * episode [n] -> obs [n][9], done [n], term_obs [n][9] (nullable), aux [n][14] (nullable).
 * mincos [n] (nullable, in/out): running minimum of |cos(theta)| over every RK4 stage - a conditioning
 * diagnostic for the tests (distance from the 1/cos(theta) pole of J2), not part of the algorithm.
 * dbmargin [n] (nullable, in/out): running minimum of the dead-band distance (see g_dbmargin), likewise. */
void orc_rov6_step(const OrcRov6Params* p, const OrcRov6Env* e, long n, double* state, const double* action, double* setpoint,
                   double* path, OrcPid6* ctrl, int32_t* istep, double* time, uint32_t* episode, double* obs, uint8_t* done,
                   double* term_obs, double* aux, double* mincos, double* dbmargin) {
    const int na = e->mode == 0 ? 8 : 6;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(e->threads > 0 ? e->threads : omp_get_max_threads())
#endif
    for (long i = 0; i < n; ++i) {
        double* y = state + 12 * i;
        double* sp = setpoint + 6 * i;
        const double* act = action + na * i;
        OrcPid6* c = ctrl ? ctrl + i : 0;
        double gcf[6] = {0, 0, 0, 0, 0, 0}, cv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        istep[i] += 1;
        time[i] += e->dt;
        g_dbmargin = 1e300;
        if (e->mode == 2 && !e->fixed_sp) { /* 6DoF.py:545-552 */
            for (int k = 0; k < 3; ++k) {
                sp[k] = act[k] * (2. * p->Length) + y[k];
                sp[3 + k] = act[3 + k] * (45. / 180. * M_PI) + y[3 + k];
            }
        }
        const double h = e->dt / e->n_sub, t0 = time[i] - e->dt;
        for (int j = 0; j < e->n_sub; ++j) {
            const double t = t0 + j * h;
            double k1[12], k2[12], k3[12], k4[12], yt[12];
            derivs6(p, e->mode, t, y, act, sp, c, k1, gcf, cv);
            for (int k = 0; k < 12; ++k) yt[k] = y[k] + 0.5 * h * k1[k];
            derivs6(p, e->mode, t + 0.5 * h, yt, act, sp, c, k2, gcf, cv);
            for (int k = 0; k < 12; ++k) yt[k] = y[k] + 0.5 * h * k2[k];
            derivs6(p, e->mode, t + 0.5 * h, yt, act, sp, c, k3, gcf, cv);
            for (int k = 0; k < 12; ++k) yt[k] = y[k] + h * k3[k];
            derivs6(p, e->mode, t + h, yt, act, sp, c, k4, gcf, cv);
            if (mincos) {
                const double th[4] = {y[4], y[4] + 0.5 * h * k1[4], y[4] + 0.5 * h * k2[4], yt[4]};
                for (int q = 0; q < 4; ++q) mincos[i] = fmin(mincos[i], fabs(cos(th[q])));
            }
            for (int k = 0; k < 12; ++k) y[k] = y[k] + (h / 6.0) * (k1[k] + 2.0 * k2[k] + 2.0 * k3[k] + k4[k]);
        }
        if (dbmargin && g_dbmargin < dbmargin[i]) dbmargin[i] = g_dbmargin;
        for (int k = 3; k < 6; ++k) y[k] = pymod(y[k], TWO_PI); /* 6DoF.py:560 */
        observe(p, y, path + 6 * i, sp, obs + 9 * i);
        done[i] = istep[i] >= e->max_steps;
        if (aux) { memcpy(aux + 14 * i, gcf, sizeof(gcf)); memcpy(aux + 14 * i + 6, cv, sizeof(cv)); }
        if (done[i] && e->auto_reset) {
            if (term_obs) memcpy(term_obs + 9 * i, obs + 9 * i, 9 * sizeof(double));
            episode[i] += 1;
            istep[i] = 0; time[i] = 0.;
            memset(y, 0, 12 * sizeof(double));
            if (!e->fixed_sp) {
                double u[9];
                draw_uniform(e->seed, e->env_id0 + (uint64_t)i, episode[i], 9, u);
                for (int k = 0; k < 6; ++k) path[6 * i + k] = (u[k] - 0.5) * 10.;
                for (int k = 0; k < 3; ++k) { sp[k] = path[6 * i + k]; sp[3 + k] = u[6 + k] * TWO_PI; }
            }
            if (c) { const double mg = c->margin; memset(c, 0, sizeof(*c)); c->margin = mg; }
            observe(p, y, path + 6 * i, sp, obs + 9 * i);
        }
    }
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
