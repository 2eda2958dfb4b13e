// C ABI of the 3DoF path (K3): host-side argument checks, constant conversion, launches.
#include <cmath>
#include <cstring>
#include <new>

#include "mvrl_host.h"
#include "rov3_kernels.cuh"

using namespace mvrl;

struct MvrlRov3 {
    MvrlRov3Params p;
    MvrlRov3Config c;
    Rov3Dev<float> pf;
    Rov3Dev<double> pd;
};

extern "C" MVRL_API int mvrl_rov3_default_params(MvrlRov3Params* p) {
    if (!p) return mvrl_fail(MVRL_EINVAL, "mvrl_rov3_default_params: null output");
    memset(p, 0, sizeof(*p));
    const double pi = 3.14159265358979323846;
    // dynamicsModel_BlueROV2_Heavy_3DoF.py:39-95
    p->rho_f = 1000.; p->m = 11.4; p->Length = 0.457; p->dispVol = p->m / p->rho_f;
    p->xg = 0.; p->yg = 0.; p->Izz = 0.16;
    p->Xudot = -5.5; p->Yvdot = -12.7; p->Nrdot = -0.12;
    p->Xuu = -18.18; p->Yvv = -21.66; p->Nrr = -1.55;
    p->Xu = -4.03; p->Yv = -6.22; p->Nr = -0.07;
    p->D_thruster = 0.1;
    const double Kt = 40. / (1000. * pow(3500. / 60., 2.) * pow(p->D_thruster, 4.));
    p->thrust_coef = p->rho_f * pow(p->D_thruster, 4.) * Kt;
    p->alphaThruster = 45. / 180. * pi; p->l_x = 0.156; p->l_y = 0.111;
    p->rpm_max = 3500.; p->rpm_deadband = 300.;
    const double m = p->m;
    const double M[9] = {m - p->Xudot, 0., -m * p->yg, 0., m - p->Yvdot, m * p->xg, -m * p->yg, m * p->xg, p->Izz - p->Nrdot};
    memcpy(p->M, M, sizeof(M));
    if (!mvrl_invert_n(p->M, p->Minv, 3)) return mvrl_fail(MVRL_EINVAL, "singular mass matrix");
    // allocation matrix with Length/2 arms (3DoF.py:104-112); full row rank -> pinv = A^T (A A^T)^-1
    const double ca = cos(p->alphaThruster), sa = sin(p->alphaThruster), arm = sa * p->Length / 2.;
    const double A[3][4] = {{ca, ca, -ca, -ca}, {sa, -sa, sa, -sa}, {arm, arm, arm, arm}};
    double AAt[9], AAtInv[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += A[i][k] * A[j][k]; AAt[i * 3 + j] = s; }
    if (!mvrl_invert_n(AAt, AAtInv, 3)) return mvrl_fail(MVRL_EINVAL, "rank-deficient allocation matrix");
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += A[k][i] * AAtInv[k * 3 + j]; p->Ainv[i * 3 + j] = s; }
    const double wind[3] = {2., 2., 90. / 180. * pi}, kp[3] = {20., 20., 20.}, ki[3] = {0.1, 0.1, 0.1}, kd[3] = {5., 5., 0.5}, mx[3] = {150., 150., 100.};
    for (int i = 0; i < 3; ++i) { p->pid_windup[i] = wind[i]; p->pid_Kp[i] = kp[i]; p->pid_Ki[i] = ki[i]; p->pid_Kd[i] = kd[i]; p->pid_max[i] = mx[i]; }
    return MVRL_OK;
}

template <typename T> static void to_dev3(const MvrlRov3Params& p, Rov3Dev<T>& d) {
    const double pi = 3.14159265358979323846;
    d.m = T(p.m); d.xg = T(p.xg); d.yg = T(p.yg); d.Xud = T(p.Xudot); d.Yvd = T(p.Yvdot);
    d.Xu = T(p.Xu); d.Yv = T(p.Yv); d.Yr = T(p.Yr); d.Nv = T(p.Nv); d.Nr = T(p.Nr);
    d.Xuu = T(p.Xuu); d.Yvv = T(p.Yvv); d.Yrr = T(p.Yrr); d.Nvv = T(p.Nvv); d.Nrr = T(p.Nrr);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) d.Minv[i][j] = T(p.Minv[i * 3 + j]);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 3; ++j) d.Ainv[i][j] = T(p.Ainv[i * 3 + j]);
    d.thrust_k = T(p.thrust_coef / 3600.); d.inv_thrust_coef = T(1. / p.thrust_coef);
    d.rpm_max = T(p.rpm_max); d.rpm_db = T(p.rpm_deadband);
    d.inv_jet_area = T(1. / (0.5 * p.rho_f * pi * p.D_thruster * p.D_thruster));
    d.drag_k = T(-0.5 * p.rho_f * pow(p.dispVol, 2. / 3.));
    d.cos_a = T(cos(p.alphaThruster)); d.sin_a = T(sin(p.alphaThruster)); d.arm = T(sqrt(p.l_x * p.l_x + p.l_y * p.l_y));
    for (int i = 0; i < 3; ++i) {
        d.pKp[i] = T(p.pid_Kp[i]); d.pKi[i] = T(p.pid_Ki[i]); d.pKd[i] = T(p.pid_Kd[i]); d.pWind[i] = T(p.pid_windup[i]); d.pMax[i] = T(p.pid_max[i]);
    }
    d.inv_3L = T(1. / (p.Length * 3.)); d.act_pos = T(2. * p.Length); d.act_ang = T(45. / 180. * pi); d.inv_ang = T(1. / (45. / 180. * pi));
}

extern "C" MVRL_API int mvrl_rov3_create(MvrlRov3** out, const MvrlRov3Params* params, const MvrlRov3Config* cfg) {
    if (!out || !params || !cfg) return mvrl_fail(MVRL_EINVAL, "mvrl_rov3_create: null argument");
    if (cfg->dtype != MVRL_F32 && cfg->dtype != MVRL_F64) return mvrl_fail(MVRL_EINVAL, "dtype must be MVRL_F32 or MVRL_F64");
    if (cfg->action_mode != MVRL_ACT_RPM && cfg->action_mode != MVRL_ACT_SETPOINT) return mvrl_fail(MVRL_EINVAL, "3DoF action_mode must be MVRL_ACT_RPM or MVRL_ACT_SETPOINT");
    if (cfg->n_sub < 1 || !(cfg->dt > 0) || !(params->thrust_coef > 0)) return mvrl_fail(MVRL_EINVAL, "mvrl_rov3_create: n_sub >= 1, dt > 0, thrust_coef > 0 required");
    { const int rc = mvrl_require_device(cfg->device); if (rc != MVRL_OK) return rc; }
    MvrlRov3* h = new (std::nothrow) MvrlRov3();
    if (!h) return mvrl_fail(MVRL_EINVAL, "out of host memory");
    h->p = *params; h->c = *cfg;
    to_dev3(*params, h->pf); to_dev3(*params, h->pd);
    *out = h;
    return MVRL_OK;
}

extern "C" MVRL_API int mvrl_rov3_destroy(MvrlRov3* h) { delete h; return MVRL_OK; }

template <typename T>
static int step3_impl(const MvrlRov3* h, const Rov3Dev<T>& P, int64_t n, int64_t ld, const MvrlRov3Buffers* b, cudaStream_t s) {
    Rov3StepArgs<T> a;
    a.P = P; a.n = n; a.ld = ld;
    a.state = (T*)b->state; a.action = (const T*)b->action; a.obs = (T*)b->obs; a.reward = (T*)b->reward; a.done = b->done;
    a.istep = b->istep; a.setpoint = (T*)b->setpoint; a.path = (T*)b->path; a.ctrl = (T*)b->ctrl; a.episode = b->episode;
    a.term_obs = (T*)b->terminal_obs; a.aux = (T*)b->aux; a.stats = b->ep_stats;
    a.dt = T(h->c.dt); a.h = T(h->c.dt / h->c.n_sub); a.hh = T(0.5) * a.h; a.h6 = a.h / T(6); a.h3 = a.h / T(3); a.n_sub = h->c.n_sub; a.max_steps = h->c.max_steps;
    a.seed = h->c.seed; a.env_id0 = h->c.env_id0; a.auto_reset = h->c.auto_reset; a.fixed_sp = h->c.fixed_sp;
    const unsigned g = mvrl_grid_for(n, 128);
    constexpr bool F32 = sizeof(T) == 4;
    const bool fast = F32 && h->c.fast_math;
    if (h->c.action_mode == MVRL_ACT_RPM) {
        if (fast) rov3_step_kernel<T, ACT_RPM, F32><<<g, 128, 0, s>>>(a); else rov3_step_kernel<T, ACT_RPM, false><<<g, 128, 0, s>>>(a);
    } else {
        if (fast) rov3_step_kernel<T, ACT_SETPOINT, F32><<<g, 128, 0, s>>>(a); else rov3_step_kernel<T, ACT_SETPOINT, false><<<g, 128, 0, s>>>(a);
    }
    return mvrl_check_launch("rov3_step");
}

extern "C" MVRL_API int mvrl_rov3_step(MvrlRov3* h, int64_t n, int64_t ld, const MvrlRov3Buffers* b, mvrl_stream_t stream) {
    if (!h || !b) return mvrl_fail(MVRL_EINVAL, "mvrl_rov3_step: null argument");
    if (n < 0 || ld < n) return mvrl_fail(MVRL_EINVAL, "mvrl_rov3_step: need 0 <= n <= ld");
    if (!b->state || !b->action || !b->obs || !b->reward || !b->done || !b->istep || !b->setpoint || !b->path)
        return mvrl_fail(MVRL_EINVAL, "mvrl_rov3_step: state/action/obs/reward/done/istep/setpoint/path are required");
    if (h->c.action_mode == MVRL_ACT_SETPOINT && !b->ctrl) return mvrl_fail(MVRL_EINVAL, "mvrl_rov3_step: ctrl is required in set-point mode");
    if (h->c.auto_reset && !b->episode) return mvrl_fail(MVRL_EINVAL, "mvrl_rov3_step: episode is required with auto_reset");
    if (n == 0) return MVRL_OK;
    MVRL_ON_DEVICE(h->c.device);
    if (h->c.dtype == MVRL_F64) return step3_impl<double>(h, h->pd, n, ld, b, (cudaStream_t)stream);
    return step3_impl<float>(h, h->pf, n, ld, b, (cudaStream_t)stream);
}

template <typename T>
static int derivs3_impl(const MvrlRov3* h, const Rov3Dev<T>& P, int64_t n, int64_t ld, const void* state, const void* act, const void* t,
                        const void* sp, void* ctrl, void* dstate, void* aux, cudaStream_t s) {
    Rov3DerivArgs<T> a;
    a.P = P; a.n = n; a.ld = ld; a.state = (const T*)state; a.act = (const T*)act; a.t = (const T*)t; a.setpoint = (const T*)sp;
    a.ctrl = (T*)ctrl; a.dstate = (T*)dstate; a.aux = (T*)aux;
    if (h->c.action_mode == MVRL_ACT_RPM) rov3_derivs_kernel<T, ACT_RPM><<<mvrl_grid_for(n, 128), 128, 0, s>>>(a);
    else rov3_derivs_kernel<T, ACT_SETPOINT><<<mvrl_grid_for(n, 128), 128, 0, s>>>(a);
    return mvrl_check_launch("rov3_derivs");
}

extern "C" MVRL_API int mvrl_rov3_derivs(MvrlRov3* h, int64_t n, int64_t ld, const void* state, const void* act, const void* t,
                                         const void* setpoint, void* ctrl, void* dstate, void* aux, mvrl_stream_t stream) {
    if (!h || !state || !dstate) return mvrl_fail(MVRL_EINVAL, "mvrl_rov3_derivs: null argument");
    if (n < 0 || ld < n) return mvrl_fail(MVRL_EINVAL, "mvrl_rov3_derivs: need 0 <= n <= ld");
    if (h->c.action_mode == MVRL_ACT_SETPOINT ? (!t || !setpoint || !ctrl) : !act) return mvrl_fail(MVRL_EINVAL, "mvrl_rov3_derivs: missing inputs for the action mode");
    if (n == 0) return MVRL_OK;
    MVRL_ON_DEVICE(h->c.device);
    if (h->c.dtype == MVRL_F64) return derivs3_impl<double>(h, h->pd, n, ld, state, act, t, setpoint, ctrl, dstate, aux, (cudaStream_t)stream);
    return derivs3_impl<float>(h, h->pf, n, ld, state, act, t, setpoint, ctrl, dstate, aux, (cudaStream_t)stream);
}

template <typename T>
static int reset3_impl(const MvrlRov3* h, const Rov3Dev<T>& P, int64_t n, int64_t ld, const MvrlRov3Buffers* b, const uint8_t* mask, const double* sp, cudaStream_t s) {
    Rov3ResetArgs<T> a;
    a.P = P; a.n = n; a.ld = ld; a.state = (T*)b->state; a.obs = (T*)b->obs; a.istep = b->istep; a.setpoint = (T*)b->setpoint;
    a.path = (T*)b->path; a.ctrl = (T*)b->ctrl; a.episode = b->episode; a.aux = (T*)b->aux; a.mask = mask;
    a.has_init_sp = sp ? 1 : 0;
    for (int k = 0; k < 3; ++k) a.init_sp[k] = sp ? T(sp[k]) : T(0);
    a.seed = h->c.seed; a.env_id0 = h->c.env_id0;
    rov3_reset_kernel<T><<<mvrl_grid_for(n, 128), 128, 0, s>>>(a);
    return mvrl_check_launch("rov3_reset");
}

extern "C" MVRL_API int mvrl_rov3_reset(MvrlRov3* h, int64_t n, int64_t ld, const MvrlRov3Buffers* b, const uint8_t* mask,
                                        const double* initial_setpoint_host, mvrl_stream_t stream) {
    if (!h || !b) return mvrl_fail(MVRL_EINVAL, "mvrl_rov3_reset: null argument");
    if (n < 0 || ld < n) return mvrl_fail(MVRL_EINVAL, "mvrl_rov3_reset: need 0 <= n <= ld");
    if (!b->state || !b->obs || !b->istep || !b->setpoint || !b->path) return mvrl_fail(MVRL_EINVAL, "mvrl_rov3_reset: state/obs/istep/setpoint/path are required");
    if (n == 0) return MVRL_OK;
    MVRL_ON_DEVICE(h->c.device);
    if (h->c.dtype == MVRL_F64) return reset3_impl<double>(h, h->pd, n, ld, b, mask, initial_setpoint_host, (cudaStream_t)stream);
    return reset3_impl<float>(h, h->pf, n, ld, b, mask, initial_setpoint_host, (cudaStream_t)stream);
}

extern "C" MVRL_API int mvrl_rov3_thruster_model(MvrlRov3* h, int64_t n, const void* u, const void* rpm, void* F, void* X, mvrl_stream_t stream) {
    if (!h || !u || !rpm || !F || !X || n < 0) return mvrl_fail(MVRL_EINVAL, "mvrl_rov3_thruster_model: bad argument");
    if (n == 0) return MVRL_OK;
    MVRL_ON_DEVICE(h->c.device);
    cudaStream_t s = (cudaStream_t)stream;
    if (h->c.dtype == MVRL_F64) rov3_thruster_kernel<double><<<mvrl_grid_for(n, 128), 128, 0, s>>>(h->pd, n, (const double*)u, (const double*)rpm, (double*)F, (double*)X);
    else rov3_thruster_kernel<float><<<mvrl_grid_for(n, 128), 128, 0, s>>>(h->pf, n, (const float*)u, (const float*)rpm, (float*)F, (float*)X);
    return mvrl_check_launch("rov3_thruster_model");
}

// LOSNavigation.predict for n observations (3DoF.py:586-607)
extern "C" MVRL_API int mvrl_los_navigation(int dtype, int64_t n, int64_t ld, const void* obs, void* action, double rnav, mvrl_stream_t stream) {
    if (!obs || !action || n < 0 || ld < n) return mvrl_fail(MVRL_EINVAL, "mvrl_los_navigation: bad argument");
    if (n == 0) return MVRL_OK;
    MVRL_ON_DEVICE_OF(action, obs, "mvrl_los_navigation");
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == MVRL_F64) los_navigation_kernel<double><<<mvrl_grid_for(n, 128), 128, 0, s>>>(n, ld, (const double*)obs, (double*)action, rnav);
    else if (dtype == MVRL_F32) los_navigation_kernel<float><<<mvrl_grid_for(n, 128), 128, 0, s>>>(n, ld, (const float*)obs, (float*)action, (float)rnav);
    else return mvrl_fail(MVRL_EINVAL, "bad dtype");
    return mvrl_check_launch("los_navigation");
}
