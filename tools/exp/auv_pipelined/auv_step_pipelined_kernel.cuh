// Round-2 experiment, measured and REJECTED: persistent auv_step kernel with double-buffered cp.async prefetch of the next
// tile's inputs.  8.15e9 env-steps/s against 9.79e9 for the plain one-pass kernel on the same box (gpurun_out/r2d): with 59
// 4-byte LDGSTS + 59 LDS per environment the MIO queue throttles (mio_throttle 1.85 per issue), only 12 warps per SM are
// resident (68 KB of shared memory per CTA), and the plain kernel was never bandwidth-bound to begin with - it is bound by
// instruction issue (18 k issue slots per scheduler = 9 us at IPC 1; measured IPC 0.4).  Kept as the record of the
// protocol; it plugs into csrc/auv_kernels.cuh after auv_step_kernel (uses auv_step_env, GatherStaged<1>, AUV_W_*).
// Pipelined kernel (fp32, plain AuvEnv, interleaved 2-component field): the plain kernel runs its ~3 waves in lock step -
// every CTA of a wave loads (22 MB requested at once), then computes, then stores - so DRAM idles while the SMs compute
// and the other way round: 0.48 of the HBM roofline (profiles/r2_c_auv_step_ncu_full_summary.txt: long_scoreboard 4.9 per
// issue, issue slots 38 % busy).  Here a persistent CTA walks over 128-environment tiles with TWO shared-memory stages:
// while it computes tile t out of one stage, cp.async (LDGSTS, no registers) is filling the other with the 59 input words
// per environment of tile t + gridDim.x.  Per tile: wait for the stage -> locate the flow cell -> issue the gather ->
// issue the prefetch of the next tile -> arithmetic (identical code: auv_step_env) -> stores.
template <bool CYL>
__global__ void __launch_bounds__(MVRL_AUV_BLOCK, 3)
auv_step_pipelined_kernel(const __grid_constant__ AuvStepArgs<float> a) {
    extern __shared__ float auv_smem[];
    float (*in_stage)[AUV_IN_WORDS][MVRL_AUV_BLOCK] = reinterpret_cast<float (*)[AUV_IN_WORDS][MVRL_AUV_BLOCK]>(auv_smem);
    float2 (*gather_slot)[MVRL_AUV_BLOCK] = reinterpret_cast<float2 (*)[MVRL_AUV_BLOCK]>(auv_smem + 2 * AUV_IN_WORDS * MVRL_AUV_BLOCK);
    const int tid = threadIdx.x;
    const long ld = a.ld, row_bytes = ld * 4;
    const long tiles = (a.n + MVRL_AUV_BLOCK - 1) / MVRL_AUV_BLOCK;
    auto cp4 = [](void* smem, const void* gmem) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
    };
    // one cp.async group = the 59 input words of this thread's environment of `tile`
    auto prefetch = [&](long tile, int stage) {
        const long i = tile * MVRL_AUV_BLOCK + tid;
        if (tile < tiles && i < a.n) {
            float (*s)[MVRL_AUV_BLOCK] = in_stage[stage];
            auto rows = [&](const void* base, int w0, int count) {
                const char* p = reinterpret_cast<const char*>(base) + i * 4;
#pragma unroll
                for (int k = 0; k < count; ++k, p += row_bytes) cp4(&s[w0 + k][tid], p);
            };
            rows(a.state, AUV_W_STATE, 6); rows(a.action, AUV_W_ACTION, 3); rows(a.mults, AUV_W_MULTS, 11);
            rows(a.target, AUV_W_TARGET, 2); rows(a.err_o, AUV_W_ERR, 3); rows(a.ep_return, AUV_W_RET, 1);
            rows(a.recent, AUV_W_RING, 30); rows(a.istep, AUV_W_ISTEP, 1);
            if (a.auto_reset) rows(a.episode, AUV_W_EPISODE, 1);
            if constexpr (CYL) rows(a.iwp, AUV_W_IWP, 1);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");   // committed even when empty: the group count stays uniform
    };
    prefetch(blockIdx.x, 0);
    int it = 0;
    for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int stage = it & 1;
        const long i = tile * MVRL_AUV_BLOCK + tid;
        asm volatile("cp.async.wait_group 0;" ::: "memory");   // this tile's inputs (each thread reads only its own column: no barrier)
        const bool live = i < a.n;
        AuvIn<float> in;
        const float (*s)[MVRL_AUV_BLOCK] = in_stage[stage];
        auto w = [&](int k) { return s[k][tid]; };
        in.x = w(0); in.y = w(1); in.psi = w(2); in.u = w(3); in.v = w(4); in.r = w(5);
        in.a0 = w(AUV_W_ACTION); in.a1 = w(AUV_W_ACTION + 1); in.a2 = w(AUV_W_ACTION + 2);
#pragma unroll
        for (int k = 0; k < 11; ++k) in.mm[k] = w(AUV_W_MULTS + k);
        in.heading_target = w(AUV_W_TARGET); in.t_offset = w(AUV_W_TARGET + 1);
        in.err_o0 = w(AUV_W_ERR); in.err_o1 = w(AUV_W_ERR + 1); in.err_o2 = w(AUV_W_ERR + 2);
        in.ep_return = w(AUV_W_RET);
#pragma unroll
        for (int q = 0; q < 10; ++q) {
#pragma unroll
            for (int c = 0; c < 3; ++c) in.ring[q][c] = w(AUV_W_RING + q * 3 + c);
        }
        in.istep = __float_as_int(w(AUV_W_ISTEP));
        in.episode = a.auto_reset ? (uint32_t)__float_as_int(w(AUV_W_EPISODE)) : 0u;
        in.iwp = CYL ? __float_as_int(w(AUV_W_IWP)) : 0;
        const long next = tile + gridDim.x;
        auto prefetch_next = [&] { prefetch(next, stage ^ 1); };
        if (live) {
            auv_step_env<float, CYL>(a, i, in, GatherStaged<1>{a.flow, gather_slot}, prefetch_next);
        } else {   // a thread beyond the batch keeps the group count in step with its warp
            asm volatile("cp.async.commit_group;" ::: "memory");
            prefetch_next();
        }
    }
}

