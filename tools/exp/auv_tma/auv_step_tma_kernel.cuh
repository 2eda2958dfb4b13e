// RECORD of a rejected variant - not compiled into libmvrl.so (round 2, GPU calls N and O; tools/exp/r2n.sh, r2o.sh).
//
// auv_step as a persistent kernel fed by the bulk copy engine: cp.async.bulk (UBLKCP) of the next tile's 57-59 input rows
// into ONE shared-memory stage, completion on an mbarrier, 5 CTAs = 20 warps per SM like the plain kernel.  Bitwise equal
// to the plain kernel (tests passed on B200), and slower:
//
//   plain one-pass kernel (shipped)                          26.7-26.9 us   9.74-9.82e9 env-steps/s
//   TMA-fed, one issuing thread per CTA                      30.6 us        8.57e9
//   TMA-fed, rows dealt out to lane 0 of each warp (below)   32.5 us        8.07e9
//
// ncu of the first version (gpurun_out/r2o): 35.6 us, 11.2 M warp instructions (plain 10.7 M), stall samples: 13 % at the
// __syncthreads that frees the stage, 11 % waiting for the first tile, 15 % on the flow-field gather, the remaining 61 %
// spread over the arithmetic (wait 1.5, mio_throttle 1.4, short_scoreboard 0.8 per issue).  The tile time of a persistent CTA
// (10.2-10.8 us) is no shorter than a wave of the plain kernel (9.7 us): hiding the input loads buys nothing, because K4
// is bound by the dependent-issue rate of its ~1300-instruction chain at 5 warps per scheduler (IPC 0.4), not by the load
// phase - and the persistent grid quantises 2.77 waves to 3 tile times.  Third negative result for this kernel after the
// cp.async pipeline (auv_pipelined/) and two environments per thread (auv_x2/).
//
// To rebuild: paste the block below into csrc/auv_kernels.cuh before `struct AuvResetArgs` and launch it with
// grid = min(tiles, 5 * SMs) when ld % 4 == 0 and every row base is 16-byte aligned.
// ---------------------------------------------------------------------------
// K4, TMA-fed: persistent CTAs, the 59 input rows of the NEXT 128-environment tile copied global -> shared by the bulk
// copy engine (cp.async.bulk, one instruction per row issued by one thread, completion counted on an mbarrier) while the
// threads compute the current tile out of registers.  The plain kernel runs its ~3 waves in lock step - every CTA loads,
// then computes, then stores, so DRAM idles while the SMs compute and the other way round; here the next tile's DRAM
// traffic is in flight during the arithmetic.  ONE shared-memory stage is enough: a thread moves its 59 words into
// registers first thing (the arithmetic lives there anyway), and after one __syncthreads the stage is free to be refilled.
// 30 KB stage + 8 KB gather slots per CTA: 5 CTAs = 20 warps per SM like the plain kernel, no per-thread LDGSTS for the
// inputs (an earlier cp.async version of the idea drowned in 59 of them per environment: tools/exp/auv_pipelined/).
// Needs 16-byte aligned rows (ld % 4 == 0); results are bitwise those of the plain kernel (same auv_step_env).
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <bool CYL>
__global__ void __launch_bounds__(MVRL_AUV_BLOCK, 5)
auv_step_tma_kernel(const __grid_constant__ AuvStepArgs<float> a) {
    __shared__ __align__(128) float in_stage[AUV_IN_WORDS][MVRL_AUV_BLOCK];
    __shared__ float2 gather_slot[8][MVRL_AUV_BLOCK];
    __shared__ __align__(8) unsigned long long full_bar;
    __shared__ const char* row_ptr[AUV_IN_WORDS];             // global address of every input row at environment 0
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = MVRL_AUV_BLOCK / 32;
    const long ld = a.ld;
    const long tiles = (a.n + MVRL_AUV_BLOCK - 1) / MVRL_AUV_BLOCK;
    const unsigned bar = smem_u32(&full_bar);
    if (tid < AUV_IN_WORDS) {
        const void* base; int k;
        if (tid < AUV_W_ACTION) { base = a.state; k = tid - AUV_W_STATE; }
        else if (tid < AUV_W_MULTS) { base = a.action; k = tid - AUV_W_ACTION; }
        else if (tid < AUV_W_TARGET) { base = a.mults; k = tid - AUV_W_MULTS; }
        else if (tid < AUV_W_ERR) { base = a.target; k = tid - AUV_W_TARGET; }
        else if (tid < AUV_W_RET) { base = a.err_o; k = tid - AUV_W_ERR; }
        else if (tid < AUV_W_RING) { base = a.ep_return; k = 0; }
        else if (tid < AUV_W_ISTEP) { base = a.recent; k = tid - AUV_W_RING; }
        else if (tid == AUV_W_ISTEP) { base = a.istep; k = 0; }
        else if (tid == AUV_W_EPISODE) { base = a.episode; k = 0; }
        else { base = a.iwp; k = 0; }
        row_ptr[tid] = reinterpret_cast<const char*>(base) + (long)k * ld * 4;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(NW));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // lane 0 of every warp: arm the barrier with its share of the tile's bytes, then one bulk copy per row r = warp, warp + NW, ...
    // (one issuing thread for all 57 rows put ~650 serial instructions per tile on one warp's critical path and the other
    // warps waited for it at the next barrier)
    auto issue = [&](long tile) {
        const long i0 = tile * MVRL_AUV_BLOCK;
        const unsigned cnt = (unsigned)((ld - i0) < MVRL_AUV_BLOCK ? (ld - i0) : MVRL_AUV_BLOCK);   // rows are padded to ld: the tail tile copies what exists
        const unsigned bytes = cnt * 4u;
        // rows of this warp: r in {warp, warp + NW, ...} below n_rows, skipping the episode row when it is not read
        int mine = 0;
        for (int r = warp; r < AUV_IN_WORDS; r += NW) mine += (r <= AUV_W_ISTEP || (r == AUV_W_EPISODE && a.auto_reset) || (r == AUV_W_IWP && CYL)) ? 1 : 0;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes * (unsigned)mine) : "memory");
        for (int r = warp; r < AUV_IN_WORDS; r += NW) {
            if (!(r <= AUV_W_ISTEP || (r == AUV_W_EPISODE && a.auto_reset) || (r == AUV_W_IWP && CYL))) continue;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(&in_stage[r][0])), "l"(row_ptr[r] + i0 * 4), "r"(bytes), "r"(bar) : "memory");
        }
    };
    if (lane == 0 && (long)blockIdx.x < tiles) issue(blockIdx.x);
    unsigned parity = 0;
    for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x, parity ^= 1u) {
        {   // wait for this tile's rows
            unsigned done = 0;
            while (!done)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        }
        const long i = tile * MVRL_AUV_BLOCK + tid;
        AuvIn<float> in;
        auto w = [&](int k) { return in_stage[k][tid]; };
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
        __syncthreads();                                     // every thread holds its words: the stage may be overwritten
        const long next = tile + gridDim.x;
        if (lane == 0 && next < tiles) issue(next);          // in flight during the arithmetic below
        if (i < a.n) auv_step_env<float, CYL>(a, i, in, GatherStaged<0>{a.flow, gather_slot}, [] {});
    }
}

