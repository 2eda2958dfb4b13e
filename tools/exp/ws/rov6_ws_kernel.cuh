// K1, warp-specialised variant (fp32, two environments per thread, direct-rpm actions, default sparsity).
//
// The plain step kernel runs its RK4 loop at the FMA-pipe limit (1114 cycles per warp-tile and sub-step = 2 cycles per
// packed instruction) while three warps per scheduler are inside the loop, but every warp spends a third of its life
// outside it - waiting for its CTA to be scheduled, for its loads to come back from DRAM, in the epilogue - and a single
// warp issues only ~0.3 instructions per clock (ptxas schedules for few registers, not for ILP), so the pipe idles
// whenever fewer than three warps of a scheduler are in the loop.  Here one persistent 512-thread
// CTA per SM splits the work by kind:
//   * warps 4..15 (three per scheduler) are COMPUTE warps: they take a tile's inputs from shared memory, turn the rpm
//     action into the thrust wrench, run the RK4 sub-steps in registers, wrap the angles, build the observation and put
//     the results back into shared memory.  They never touch global memory, so nothing but arithmetic keeps them from
//     the loop.
//   * warps 0..3 (one per scheduler) are IO warps: cp.async the next tile's inputs into shared memory, and after the
//     compute warp is done store state / observation / reward / done / counter, accumulate the episode statistics and
//     run the (rare) auto-reset.
// A tile is 32 lanes x 2 environments; every compute warp owns two stages (double buffer) that it and its IO warp hand
// back and forth through two mbarriers per stage (full: inputs staged; done: results staged).  setmaxnreg moves
// registers from the IO warpgroup (56) to the three compute warpgroups (152).
// Per-environment arithmetic is exactly that of rov6_step_kernel<F2, ACT_RPM, true, false>: results are bitwise equal
// (tools/exp/ws_check.py, tests/test_rov6_gpu.py).
#pragma once
#include "rov6_kernels.cuh"

namespace mvrl {

constexpr int WS_IO_WARPS = 4, WS_COMPUTE_WARPS = 12, WS_SERVED = WS_COMPUTE_WARPS / WS_IO_WARPS;
constexpr int WS_THREADS = 32 * (WS_IO_WARPS + WS_COMPUTE_WARPS), WS_STAGES = 2;
// input slots (float2 per lane): state 12, rpm action 8, way-points 6, target angles 3;
// output slots: the state in place, the 9 observations over the (by then dead) action and first way-point slots
constexpr int WS_SLOT_Y = 0, WS_SLOT_ACT = 12, WS_SLOT_PATH = 20, WS_SLOT_SP = 26, WS_SLOTS = 29, WS_SLOT_OBS = 12;
enum { WS_DONE0 = 1, WS_DONE1 = 2, WS_RESET0 = 4, WS_RESET1 = 8, WS_BAD0 = 16, WS_BAD1 = 32 };

struct WsStage {
    float2 v[WS_SLOTS][32];
    int istep[32][2];      // in: step counters; out: incremented
    int flags[32];         // out: WS_* bits of the two environments of the lane
};
struct WsShared {
    WsStage stage[WS_COMPUTE_WARPS][WS_STAGES];
    unsigned long long full[WS_COMPUTE_WARPS][WS_STAGES], done[WS_COMPUTE_WARPS][WS_STAGES];
};

__device__ __forceinline__ unsigned ws_saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ws_saddr(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ws_saddr(b)) : "memory");
}
// Waits for the phase with the given parity to complete; traps instead of hanging if it never does.  SLEEP_NS > 0: back
// off between polls - an IO warp that spins on try_wait executes hundreds of IMAD / ISETP / BRA per tile on the very
// pipes the compute warps need (r1_ws profile: the spinning IO warps issued 40 % as many instructions as the compute warps).
template <unsigned SLEEP_NS>
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    unsigned ok = 0, spins = 0;
    unsigned long long t0 = 0, t1;
    for (;;) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(ws_saddr(b)), "r"(parity) : "memory");
        if (ok) return;
        if (SLEEP_NS) __nanosleep(SLEEP_NS);
        if ((++spins & (SLEEP_NS ? 63u : 1023u)) == 0u) {   // a protocol error must end as a launch failure, not as a hung GPU
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t0 == 0) t0 = t1;
            else if (t1 - t0 > 2000000000ull) __trap();
        }
    }
}

// n_sub classic RK4 sub-steps of the direct-rpm model: rov6_step_kernel's loop for MODE = ACT_RPM, SP, !FAST
// (same operations in the same order, so the results are bitwise equal)
__device__ __forceinline__ void ws_rk4(const Rov6StepArgs<float>& a, F2 (&y)[12], const F2 (&H)[6]) {
    using V = F2;
    const Rov6Dev<float>& P = a.P;
    const float h = a.h, hh = a.hh, h6 = a.h6, h3 = a.h3;
    V carry[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) carry[j] = V(0.0f);
    int sub = a.n_sub;
#pragma unroll 1
    do {
        V k[12], acc[12], yt[12];
        Trig6<V> g0;
#pragma unroll
        for (int j = 0; j < 12; ++j) { acc[j] = (j < 6) ? V(0.0f) : y[j]; yt[j] = y[j]; }
#pragma unroll
        for (int st = 0; st < 4; ++st) {
            Trig6<V> g;
            if (st == 0) {
                g0 = trig6<V, false>(y[3], y[4], y[5]);
                g = g0;
            } else {
                const float cp = (st == 3) ? h : hh;
                const V d0 = V(cp) * k[3], d1 = V(cp) * k[4], d2 = V(cp) * k[5];
                const V z0 = d0 * d0, z1 = d1 * d1, z2 = d2 * d2;
                const V zs = z0 + z1 + z2;
                sincos_delta(g0.sph, g0.cph, d0, z0, &g.sph, &g.cph);
                sincos_delta(g0.sth, g0.cth, d1, z1, &g.sth, &g.cth);
                sincos_delta(g0.sps, g0.cps, d2, z2, &g.sps, &g.cps);
                const auto big = vgt(zs, V(MVRL_TRIG_DELTA_MAX2));
                if (vany(big)) {
                    const Trig6<V> gf = trig6<V, false>(yt[3], yt[4], yt[5]);
                    g.sph = vsel(big, gf.sph, g.sph); g.cph = vsel(big, gf.cph, g.cph);
                    g.sth = vsel(big, gf.sth, g.sth); g.cth = vsel(big, gf.cth, g.cth);
                    g.sps = vsel(big, gf.sps, g.sps); g.cps = vsel(big, gf.cps, g.cps);
                }
            }
            const V nu[6] = {yt[6], yt[7], yt[8], yt[9], yt[10], yt[11]};
            V ed[6], ac[6], rhs[6];
            kinematics6<V, false>(g, nu, ed);
            body_accel<V, true>(P, g, nu, H, ac, rhs);
#pragma unroll
            for (int j = 0; j < 6; ++j) { k[j] = ed[j]; k[6 + j] = ac[j]; }
            const float wk = (st == 0 || st == 3) ? h6 : h3;
            const float ck = (st == 2) ? h : hh;
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                acc[j] = fmaf_t(V(wk), k[j], acc[j]);
                if (st < 3) yt[j] = fmaf_t(V(ck), k[j], y[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            if (j < 6) rk4_pose_update(y[j], carry[j], acc[j]);
            else y[j] = acc[j];
        }
    } while (--sub > 0);
}

// wrap + dataToState + done / reset / non-finite flags of one thread's two environments: rov6_step_kernel's epilogue
// (MODE = ACT_RPM) up to, but without, the global stores
__device__ __forceinline__ int ws_observe(const Rov6StepArgs<float>& a, bool pair, F2 (&y)[12], const F2 (&path_v)[6], const F2 (&sp_ang)[3],
                                          int (&istep)[2], F2 (&obs_v)[9]) {
    using V = F2;
    using T = float;
    constexpr int L = 2;
    const Rov6Dev<T>& P = a.P;
    int flags = 0;
    V nonfinite = V(T(0));
#pragma unroll
    for (int k = 0; k < 12; ++k) nonfinite = fmaf_t(y[k], V(T(0)), nonfinite);
    bool lane_on[L], ok[L];
    T ang_raw[L][3], spa[L][3];
    V pos_v[6];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        pos_v[k] = (path_v[k] - y[k]) * V(P.inv_3L);
        pos_v[3 + k] = (path_v[3 + k] - y[k]) * V(P.inv_3L);
    }
#pragma unroll
    for (int l = 0; l < L; ++l) {
        lane_on[l] = (l == 0) || pair;
        T ys[12], path[6], obs[9], wrapped[3];
#pragma unroll
        for (int k = 0; k < 12; ++k) ys[k] = lane_get(y[k], l);
#pragma unroll
        for (int k = 0; k < 6; ++k) path[k] = lane_get(pos_v[k], l);
#pragma unroll
        for (int k = 0; k < 3; ++k) { spa[l][k] = lane_get(sp_ang[k], l); ang_raw[l][k] = ys[3 + k]; }
        istep[l] = istep[l] + 1;
        bool bad = lane_get(nonfinite, l) != T(0);
        bad = bad || tmax(tmax(tabs(ys[3]), tabs(ys[4])), tabs(ys[5])) > T(MVRL_SINCOS_F32_MAX_ARG);
        ok[l] = wrap_observe6_fast(P, ys, path, spa[l], wrapped, obs) || !lane_on[l];
#pragma unroll
        for (int k = 0; k < 3; ++k) lane_set(y[3 + k], l, wrapped[k]);
#pragma unroll
        for (int k = 0; k < 9; ++k) lane_set(obs_v[k], l, obs[k]);
        const bool is_done = istep[l] >= a.max_steps;
        if (lane_on[l] && is_done) flags |= (WS_DONE0 << l);
        if (lane_on[l] && is_done && a.auto_reset) flags |= (WS_RESET0 << l);
        if (lane_on[l] && bad) flags |= (WS_BAD0 << l);
    }
    bool all_ok = true;
#pragma unroll
    for (int l = 0; l < L; ++l) all_ok = all_ok && ok[l];
    if (!all_ok) {
#pragma unroll
        for (int l = 0; l < L; ++l) {
            if (ok[l]) continue;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const WrapObs<T> r = wrap_angle_exact(P.inv_ang, ang_raw[l][k], spa[l][k]);
                lane_set(y[3 + k], l, r.wrapped);
                lane_set(obs_v[6 + k], l, r.obs);
            }
        }
    }
    return flags;
}

__global__ void __launch_bounds__(WS_THREADS, 1)
rov6_step_ws_kernel(const __grid_constant__ Rov6StepArgs<float> a) {
    extern __shared__ __align__(16) unsigned char ws_smem[];
    WsShared& sh = *reinterpret_cast<WsShared*>(ws_smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long ld = a.ld;
    const long n_tiles = (a.n + 63) / 64;                              // warp-tiles of 32 lanes x 2 environments
    const long tile_stride = (long)gridDim.x * WS_COMPUTE_WARPS;

    // zeroed stages: lanes past the end of the batch integrate zeros instead of whatever shared memory held
    for (int e = tid; e < (int)(sizeof(WsShared) / 4); e += WS_THREADS) reinterpret_cast<unsigned*>(ws_smem)[e] = 0u;
    __syncthreads();
    if (tid < WS_COMPUTE_WARPS * WS_STAGES) {
        mbar_init(&sh.full[tid / WS_STAGES][tid % WS_STAGES], 1);
        mbar_init(&sh.done[tid / WS_STAGES][tid % WS_STAGES], 1);
    }
    __syncthreads();

    if (warp >= WS_IO_WARPS) {
        // ------------------------------------------------------------------ compute warps
        asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
        const int c = warp - WS_IO_WARPS;
        long tile = (long)blockIdx.x * WS_COMPUTE_WARPS + c;
        for (int k = 0; tile < n_tiles; ++k, tile += tile_stride) {
            WsStage& st = sh.stage[c][k & 1];
            mbar_wait<0>(&sh.full[c][k & 1], (unsigned)(k >> 1) & 1u);
            F2 y[12], H[6];
#pragma unroll
            for (int j = 0; j < 12; ++j) y[j] = f2_from(st.v[WS_SLOT_Y + j][lane]);
            {   // rpm -> thruster forces -> wrench (6DoF.py:271-282): constant over the env step
                F2 F[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) F[j] = thruster_force(a.P, f2_from(st.v[WS_SLOT_ACT + j][lane]));
                thrust_wrench<F2, true>(a.P, F, H);
            }
            ws_rk4(a, y, H);
            F2 path_v[6], sp_ang[3], obs_v[9];
#pragma unroll
            for (int j = 0; j < 6; ++j) path_v[j] = f2_from(st.v[WS_SLOT_PATH + j][lane]);
#pragma unroll
            for (int j = 0; j < 3; ++j) sp_ang[j] = f2_from(st.v[WS_SLOT_SP + j][lane]);
            int istep[2] = {st.istep[lane][0], st.istep[lane][1]};
            const bool pair = (tile * 32 + lane) * 2 + 1 < a.n;
            const int flags = ws_observe(a, pair, y, path_v, sp_ang, istep, obs_v);
#pragma unroll
            for (int j = 0; j < 12; ++j) st.v[WS_SLOT_Y + j][lane] = y[j].v;
#pragma unroll
            for (int j = 0; j < 9; ++j) st.v[WS_SLOT_OBS + j][lane] = obs_v[j].v;
            st.istep[lane][0] = istep[0]; st.istep[lane][1] = istep[1];
            st.flags[lane] = flags;
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.done[c][k & 1]);
        }
    } else {
        // ------------------------------------------------------------------ IO warps: each serves WS_SERVED compute warps
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        static_assert(WS_SERVED == 3, "the cp.async group waits below assume three served warps");
        const long cta_first = (long)blockIdx.x * WS_COMPUTE_WARPS;
        const long rounds = cta_first < n_tiles ? (n_tiles - cta_first + tile_stride - 1) / tile_stride : 0;   // tiles of the busiest compute warp
        for (long k = 0; k < rounds + 2; ++k) {
            const int s = (int)(k & 1);
            // (1) per served warp: store the results of tile k - 2 (its stage is the one tile k reuses), queue the loads of tile k
#pragma unroll 1
            for (int q = 0; q < WS_SERVED; ++q) {
                const int c = warp + WS_IO_WARPS * q;
                WsStage& st = sh.stage[c][s];
                const long t_old = cta_first + c + (k - 2) * tile_stride;
                if (k >= 2 && t_old < n_tiles) {
                    mbar_wait<300>(&sh.done[c][s], (unsigned)((k - 2) >> 1) & 1u);
                    const long i0 = (t_old * 32 + lane) * 2;
                    const bool on = i0 < a.n, pair = i0 + 1 < a.n;
                    int n_done = 0, len = 0, n_bad = 0, flags = 0;
                    if (on) {
                        flags = st.flags[lane];
#pragma unroll
                        for (int j = 0; j < 12; ++j) store_v<F2>(a.state + j * ld, i0, pair, f2_from(st.v[WS_SLOT_Y + j][lane]));
#pragma unroll
                        for (int j = 0; j < 9; ++j) store_v<F2>(a.obs + j * ld, i0, pair, f2_from(st.v[WS_SLOT_OBS + j][lane]));
                        store_v<F2>(a.reward, i0, pair, F2(0.0f));   // 6DoF.py:575
#pragma unroll
                        for (int l = 0; l < 2; ++l) {
                            if (l == 0 || pair) {
                                a.done[i0 + l] = (flags & (WS_DONE0 << l)) ? 1 : 0;
                                a.istep[i0 + l] = st.istep[lane][l];
                            }
                            if (flags & (WS_RESET0 << l)) { n_done += 1; len += st.istep[lane][l]; }
                            if (flags & (WS_BAD0 << l)) n_bad += 1;
                        }
                    }
                    if (a.stats != nullptr) stats_accumulate_counts(a.stats, n_done, len, n_bad);
#pragma unroll
                    for (int l = 0; l < 2; ++l) {
                        if (flags & (WS_RESET0 << l)) rov6_auto_reset_env<float, ACT_RPM>(a, i0 + l, a.episode[i0 + l]);
                    }
                }
                const long t_new = cta_first + c + k * tile_stride;
                if (k < rounds && t_new < n_tiles) {
                    const long i0 = (t_new * 32 + lane) * 2;
                    if (i0 + 1 < a.n) {
                        auto row = [&](int slot, const float* src) { cp_async_v<F2>(reinterpret_cast<F2*>(&st.v[slot][lane]), src + i0, true); };
#pragma unroll
                        for (int j = 0; j < 12; ++j) row(WS_SLOT_Y + j, a.state + j * ld);
#pragma unroll
                        for (int j = 0; j < 8; ++j) row(WS_SLOT_ACT + j, a.action + j * ld);
#pragma unroll
                        for (int j = 0; j < 6; ++j) row(WS_SLOT_PATH + j, a.path + j * ld);
#pragma unroll
                        for (int j = 0; j < 3; ++j) row(WS_SLOT_SP + j, a.setpoint + (3 + j) * ld);
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(ws_saddr(&st.istep[lane][0])), "l"(a.istep + i0));
                    } else if (i0 < a.n) {   // the one unpaired environment of an odd batch
                        auto row = [&](int slot, const float* src) { st.v[slot][lane] = make_float2(src[i0], 0.0f); };
#pragma unroll
                        for (int j = 0; j < 12; ++j) row(WS_SLOT_Y + j, a.state + j * ld);
#pragma unroll
                        for (int j = 0; j < 8; ++j) row(WS_SLOT_ACT + j, a.action + j * ld);
#pragma unroll
                        for (int j = 0; j < 6; ++j) row(WS_SLOT_PATH + j, a.path + j * ld);
#pragma unroll
                        for (int j = 0; j < 3; ++j) row(WS_SLOT_SP + j, a.setpoint + (3 + j) * ld);
                        st.istep[lane][0] = a.istep[i0]; st.istep[lane][1] = 0;
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            // (2) publish each stage as its inputs land
#pragma unroll 1
            for (int q = 0; q < WS_SERVED; ++q) {
                const int c = warp + WS_IO_WARPS * q;
                const long t_new = cta_first + c + k * tile_stride;
                if (q == 0) asm volatile("cp.async.wait_group 2;" ::: "memory");
                else if (q == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
                else asm volatile("cp.async.wait_group 0;" ::: "memory");
                if (k < rounds && t_new < n_tiles) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sh.full[c][s]);
                }
            }
        }
    }
}

}  // namespace mvrl
