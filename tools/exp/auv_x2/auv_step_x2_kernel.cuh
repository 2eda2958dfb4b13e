// RECORD of a rejected variant - not compiled into libmvrl.so (round 2, GPU calls J and K; tools/exp/r2j.sh, r2k.sh).
//
// auv_step with TWO fp32 environments per thread on the packed FADD2 / FMUL2 / FFMA2 path, 8-byte row accesses.
// It does what it was written for - 10.7 M -> 8.2 M warp instructions per 262 144-env launch - and is still slower than
// the one-environment kernel on every launch shape, because K4 is bound by dependent latency, not by instruction issue:
//
//   one environment per thread (shipped), 102 registers, 20 warps / SM     26.7 us   9.82e9 env-steps/s
//   x2, 181 registers,  8 warps / SM (__launch_bounds__(128, 2))           33.2 us   7.89e9
//   x2, 166 registers, 12 warps / SM (128, 3)                               31.9 us   8.22e9
//   x2, 128 registers, 16 warps / SM (128, 4; 120 B of spills)              34.5 us   7.60e9
//
// ncu of the 8-warp shape (gpurun_out/r2j): 38 us, issue slots 19 % busy, long_scoreboard 2.0 + wait 1.4 per issue, 9.7 % of the
// warp slots occupied.  Two environments per thread halve the number of independent dependency chains per SM at the same
// register budget; the chains (state -> flow cell -> gather -> sincos / exp / sqrt -> stores) are what the kernel waits on.
// The parity test that went with it (packed vs one-environment kernel, 4097 envs x 40 steps with auto-reset: observations
// 1e-5, rewards 1e-4, same Philox draws) is test_auv_x2_variant.py next to this file.
//
// To rebuild: paste the block below back into csrc/auv_kernels.cuh before `struct AuvResetArgs` and add the dispatch that
// `git show dc4adc6 -- marinevehiclereinforcementlearning_b200/csrc/mvrl_auv.cu` shows.
// ---------------------------------------------------------------------------
// K4, packed: TWO fp32 environments per thread (2t, 2t + 1) for the plain AuvEnv.  The one-environment kernel is bound by
// instruction issue, not by bandwidth (profiles/r2_f_auv_step_ncu_full_summary.txt: 1310 instructions per environment, of
// which ~500 are integer / address arithmetic and ~120 loads and stores, IPC 0.43, DRAM 26 %).  Here every row access is one
// 8-byte vector (half the memory instructions and address arithmetic per environment) and the floating-point work runs on
// the packed FADD2 / FMUL2 / FFMA2 instructions of sm_100 (value type F2, mvrl_math.cuh); what stays per lane is what has
// no packed form: the flow-cell lookup and gather, sqrt / exp / reciprocal (MUFU), min / max, and the rare auto-reset.
// Rows must be 8-byte aligned with an even leading dimension; the second lane of an unpaired last thread reads the row's
// padding and stores nothing.
// ---------------------------------------------------------------------------
// auto-reset of ONE environment after the regular stores (same thread: program order makes these stores win)
template <typename T>
__device__ MVRL_NOINLINE void auv_auto_reset_env(const AuvStepArgs<T>& a, long i, uint32_t episode_in) {
    const AuvDev<T>& P = a.P;
    const long ld = a.ld;
    if (a.term_obs != nullptr) {
#pragma unroll
        for (int k = 0; k < 11; ++k) a.term_obs[k * ld + i] = a.obs[k * ld + i];
    }
    const uint32_t ep = episode_in + 1u;
    a.episode[i] = ep;
    T mm[11], x, y, psi, heading_target = a.target[i], t_offset;
    draw_reset_auv(P, a.seed, a.env_id0 + (unsigned long long)i, ep, a.apply_noise != 0, mm, &x, &y, &psi, &heading_target, &t_offset);
#pragma unroll
    for (int k = 0; k < 11; ++k) a.mults[k * ld + i] = mm[k];
    a.target[i] = heading_target;
    a.target[ld + i] = t_offset;
    a.state[i] = x; a.state[ld + i] = y; a.state[2 * ld + i] = psi;
    a.state[3 * ld + i] = T(0); a.state[4 * ld + i] = T(0); a.state[5 * ld + i] = T(0);
    a.istep[i] = 0;
    a.ep_return[i] = T(0);
    const T perr_x = -x, perr_y = -y, herr = angle_error(heading_target, psi);
    a.err_o[i] = perr_x; a.err_o[ld + i] = perr_y; a.err_o[2 * ld + i] = herr;
    T obs[11];
    observe_auv<false>(T(0), T(0), x, y, psi, T(0), T(0), T(0), heading_target, perr_x, perr_y, herr, obs);
#pragma unroll
    for (int k = 0; k < 11; ++k) a.obs[k * ld + i] = obs[k];
}

#ifndef MVRL_AUV_X2_MINB
#define MVRL_AUV_X2_MINB 2
#endif
__global__ void __launch_bounds__(MVRL_AUV_BLOCK, MVRL_AUV_X2_MINB)
auv_step_x2_kernel(const __grid_constant__ AuvStepArgs<float> a) {
    __shared__ float2 stage[2][8][MVRL_AUV_BLOCK];
    const long i0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i0 >= a.n) return;
    const bool pair = i0 + 1 < a.n;
    const AuvDev<float>& P = a.P;
    const long ld = a.ld, row_bytes = ld * 4;
    auto row2 = [&](const float* base, int k) {
        return f2_from(*reinterpret_cast<const float2*>(reinterpret_cast<const char*>(base + i0) + k * row_bytes));
    };
    // ---- every input, before anything waits on one of them
    F2 x = row2(a.state, 0), y = row2(a.state, 1), psi = row2(a.state, 2);
    const F2 heading_target = row2(a.target, 0), t_offset = row2(a.target, 1);
    const int2 istep_in = *reinterpret_cast<const int2*>(a.istep + i0);
    F2 u = row2(a.state, 3), v = row2(a.state, 4), r = row2(a.state, 5);
    const F2 a0 = row2(a.action, 0), a1 = row2(a.action, 1), a2 = row2(a.action, 2);
    F2 mm[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) mm[k] = row2(a.mults, k);
    const F2 err_o0 = row2(a.err_o, 0), err_o1 = row2(a.err_o, 1), err_o2 = row2(a.err_o, 2);
    const F2 ep_return_in = row2(a.ep_return, 0);
    uint2 episode_in = make_uint2(0u, 0u);
    if (a.auto_reset) episode_in = *reinterpret_cast<const uint2*>(a.episode + i0);
    F2 ring[10][3];
#pragma unroll
    for (int s = 0; s < 10; ++s) {
#pragma unroll
        for (int c = 0; c < 3; ++c) ring[s][c] = row2(a.recent, s * 3 + c);
    }
    const int istep[2] = {istep_in.x + 1, istep_in.y + 1};
    const F2 time = F2(float(istep[0]) * a.dt, float(istep[1]) * a.dt);
    const F2 tq = time + t_offset;
    // ---- flow cell and gather, per lane (each lane has its own cell); the copies land while the ring statistics are set up
    FlowCell<float> cell[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        cell[l] = flow_locate<true>(a.flow, lane_get(tq, l), lane_get(x, l), lane_get(y, l));
        flow_stage_issue(a.flow, cell[l], stage[l]);
    }
    // recentActions.appendleft(action): ring slot, then statistics over the valid entries
    const int slot[2] = {(istep[0] - 1) % 10, (istep[1] - 1) % 10};
    const int cnt[2] = {istep[0] < 10 ? istep[0] : 10, istep[1] < 10 ? istep[1] : 10};
#pragma unroll
    for (int s = 0; s < 10; ++s) {
        const B2 here = B2{s == slot[0], s == slot[1]};
        ring[s][0] = vsel(here, a0, ring[s][0]); ring[s][1] = vsel(here, a1, ring[s][1]); ring[s][2] = vsel(here, a2, ring[s][2]);
    }
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        if (l == 0 || pair) {
            a.recent[(slot[l] * 3 + 0) * ld + i0 + l] = lane_get(a0, l);
            a.recent[(slot[l] * 3 + 1) * ld + i0 + l] = lane_get(a1, l);
            a.recent[(slot[l] * 3 + 2) * ld + i0 + l] = lane_get(a2, l);
        }
    }
    const F2 inv_cnt = F2(__frcp_rn(float(cnt[0])), __frcp_rn(float(cnt[1])));
    F2 rms = F2(0.0f);
    {
        F2 valid[10];
#pragma unroll
        for (int s = 0; s < 10; ++s) valid[s] = F2(s < cnt[0] ? 1.0f : 0.0f, s < cnt[1] ? 1.0f : 0.0f);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            F2 mean = F2(0.0f);
#pragma unroll
            for (int s = 0; s < 10; ++s) mean = fmaf_t(valid[s], ring[s][c], mean);
            mean = mean * inv_cnt;
            F2 var = F2(0.0f);
#pragma unroll
            for (int s = 0; s < 10; ++s) { const F2 d = ring[s][c] - mean; var = fmaf_t(valid[s], d * d, var); }
            const F2 q = var * inv_cnt;
            rms += F2(sqrtf(q.v.x), sqrtf(q.v.y));
        }
        rms = rms * F2(float(1. / 3.));
    }
    const F2 Fx_set = a0 * F2(P.maxForce) * mm[8], Fy_set = a1 * F2(P.maxForce) * mm[9], N_set = a2 * F2(P.maxMoment) * mm[10];
    F2 sn, cs;
    sincos_f32(psi, &sn, &cs);     // psi is wrapped to [0, 2 pi) every step: far inside the exact range of the reduction
    cp_async_wait_all();           // each thread reads back only what it copied itself: no barrier needed
    float cur0[2], cur1[2];
    flow_stage_blend(cell[0], stage[0], cur0);
    flow_stage_blend(cell[1], stage[1], cur1);
    const F2 curx = F2(cur0[0], cur1[0]), cury = F2(cur0[1], cur1[1]);
    const F2 dxv = u - curx, dyv = v - cury;
    const F2 vr0 = fmaf_t(cs, dxv, sn * dyv), vr1 = fmaf_t(cs, dyv, -(sn * dxv));
    const F2 fh0 = fmaf_t(F2(P.Xuu) * mm[2], tabs(vr0), F2(P.Xu) * mm[5]) * vr0;
    const F2 fh1 = fmaf_t(F2(P.Yvv) * mm[3], tabs(vr1), F2(P.Yv) * mm[6]) * vr1;
    const F2 fh2 = fmaf_t(F2(P.Nrr) * mm[4], tabs(r), F2(P.Nr) * mm[7]) * r;
    const F2 Fx = fmaf_t(cs, fh0, -(sn * fh1)), Fy = fmaf_t(sn, fh0, cs * fh1);
    const F2 den_m = F2(P.m) * mm[0], den_i = F2(P.Izz) * mm[1];
    const F2 inv_m = F2(__frcp_rn(den_m.v.x), __frcp_rn(den_m.v.y)), inv_i = F2(__frcp_rn(den_i.v.x), __frcp_rn(den_i.v.y));
    const F2 ax = (Fx + Fx_set) * inv_m, ay = (Fy + Fy_set) * inv_m, ar = (fh2 + N_set) * inv_i;
    // explicit Euler, position advanced with the OLD velocity (verySimpleAuv.py:321-326)
    const F2 dt2 = F2(a.dt);
    x = fmaf_t(u, dt2, x);
    y = fmaf_t(v, dt2, y);
    {
        const F2 raw = fmaf_t(r, dt2, psi);
        const F2 tp = F2(float(MVRL_TWO_PI));
        F2 w = fmaf_t(tp, vmask_lt(raw, F2(0.0f)) - vmask_ge(raw, tp), raw);   // Python % on (-2 pi, 4 pi): one turn at most
        const bool far0 = !(raw.v.x > -float(MVRL_TWO_PI) && raw.v.x < 2.0f * float(MVRL_TWO_PI));
        const bool far1 = !(raw.v.y > -float(MVRL_TWO_PI) && raw.v.y < 2.0f * float(MVRL_TWO_PI));
        if (far0 || far1) {   // a yaw rate beyond 300 rad/s (or NaN): the exact out-of-line modulo
            if (far0) w.v.x = pymod_general(raw.v.x, float(MVRL_TWO_PI));
            if (far1) w.v.y = pymod_general(raw.v.y, float(MVRL_TWO_PI));
        }
        psi = w;
    }
    u = fmaf_t(ax, dt2, u);
    v = fmaf_t(ay, dt2, v);
    r = fmaf_t(ar, dt2, r);

    // dataToState V3 (target at the origin, no scaling) and the reward's errors
    const F2 perr_x = -x, perr_y = -y;
    const F2 herr = angle_error_v(heading_target, psi);
    const F2 one = F2(1.0f), mone = F2(-1.0f);
    auto clip = [&](F2 q) { return tmax(mone, tmin(one, q)); };
    F2 obs[9];
    obs[0] = clip(perr_x); obs[1] = clip(perr_y);
    obs[2] = clip(herr * F2(float(1. / (45. / 180. * 3.14159265358979323846))));
    obs[3] = clip(herr - err_o2); obs[4] = clip(perr_x - err_o0); obs[5] = clip(perr_y - err_o1);
    obs[6] = clip(u); obs[7] = clip(v); obs[8] = clip(r);

    const F2 pn2 = fmaf_t(perr_x, perr_x, perr_y * perr_y);
    const F2 perr_norm = F2(sqrtf(pn2.v.x), sqrtf(pn2.v.y));
    const F2 herr_deg = tabs(herr * F2(float(180. / 3.14159265358979323846)));
    const F2 t3 = F2(float(-0.1 / 3.)) * fmaf_t(a0, a0, fmaf_t(a1, a1, a2 * a2));
    float rew[2], ep_ret[2];
    bool is_done[2], lane_on[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        lane_on[l] = (l == 0) || pair;
        const float xl = lane_get(x, l), yl = lane_get(y, l), hl = lane_get(herr, l), hd = lane_get(herr_deg, l);
        float bonus = 0.0f;
        bool done = istep[l] >= a.max_steps;
        if (xl < P.xmin || xl > P.xmax) { if (a.stop_on_bounds) done = true; bonus += -100.0f; }
        if (yl < P.ymin || yl > P.ymax) { if (a.stop_on_bounds) done = true; bonus += -100.0f; }
        const float t0 = expf(-5.0f * lane_get(perr_norm, l));
        const bool facing = fabsf(hl) < 3.14159265358979323846f * 0.5f;
        const float t1e = expf(-0.1f * (facing ? hd : 180.0f - hd));
        const float t1 = facing ? t1e : -t1e;
        const float t2 = expf(-0.6f * lane_get(rms, l));
        rew[l] = t0 + t1 + t2 + lane_get(t3, l) + bonus;
        ep_ret[l] = lane_get(ep_return_in, l) + rew[l];
        is_done[l] = done;
        if (a.aux != nullptr && lane_on[l]) {  // the per-step log columns of verySimpleAuv.py:389-401 that are not state/obs
            const float vals[14] = {lane_get(Fx, l), lane_get(Fy, l), lane_get(fh2, l), lane_get(Fx_set, l), lane_get(Fy_set, l), lane_get(N_set, l),
                                    lane_get(curx, l), lane_get(cury, l), lane_get(rms, l), t0, t1, t2, lane_get(t3, l), bonus};
#pragma unroll
            for (int k = 0; k < 14; ++k) a.aux[k * ld + i0 + l] = vals[k];
        }
        const bool bad = !(finite_t(xl) && finite_t(yl) && finite_t(lane_get(psi, l)) && finite_t(lane_get(u, l)) && finite_t(lane_get(v, l)) && finite_t(lane_get(r, l)));
        if (a.stats != nullptr) stats_accumulate(a.stats, lane_on[l] && done && a.auto_reset, (double)istep[l], (double)ep_ret[l], lane_on[l] && bad);
    }
    // ---- regular stores (8-byte vectors; the unpaired thread stores one element), then the rare auto-reset fix-ups
    auto st2 = [&](float* base, int k, F2 q) {
        char* p = reinterpret_cast<char*>(base + i0) + k * row_bytes;
        if (pair) *reinterpret_cast<float2*>(p) = q.v; else *reinterpret_cast<float*>(p) = q.v.x;
    };
    st2(a.state, 0, x); st2(a.state, 1, y); st2(a.state, 2, psi); st2(a.state, 3, u); st2(a.state, 4, v); st2(a.state, 5, r);
    st2(a.err_o, 0, perr_x); st2(a.err_o, 1, perr_y); st2(a.err_o, 2, herr);
#pragma unroll
    for (int k = 0; k < 9; ++k) st2(a.obs, k, obs[k]);
    st2(a.obs, 9, F2(0.0f)); st2(a.obs, 10, F2(0.0f));
    st2(a.reward, 0, F2(rew[0], rew[1]));
    st2(a.ep_return, 0, F2(ep_ret[0], ep_ret[1]));
    if (pair) {
        *reinterpret_cast<uchar2*>(a.done + i0) = make_uchar2(is_done[0] ? 1 : 0, is_done[1] ? 1 : 0);
        *reinterpret_cast<int2*>(a.istep + i0) = make_int2(istep[0], istep[1]);
    } else {
        a.done[i0] = is_done[0] ? 1 : 0;
        a.istep[i0] = istep[0];
    }
    if (a.auto_reset) {
        if (is_done[0]) auv_auto_reset_env<float>(a, i0, episode_in.x);
        if (pair && is_done[1]) auv_auto_reset_env<float>(a, i0 + 1, episode_in.y);
    }
}

