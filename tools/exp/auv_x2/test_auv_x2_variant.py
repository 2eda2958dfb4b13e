"""Parity test of the rejected two-environments-per-thread auv_step (see auv_step_x2_kernel.cuh); needs the helpers of
tests/test_auv_gpu.py (load_golden, make_flows, AuvVecEnv, DEV) and a library built with the variant."""


def test_auv_two_envs_per_thread_kernel_vs_one_env_kernel(monkeypatch):
    """fp32 plain env: the packed kernel (two environments per thread, FADD2 / FMUL2 / FFMA2, 8-byte row accesses) against
    the one-environment kernel (MVRL_AUV_NO_X2=1) - odd batch (an unpaired last thread), auto-reset with Philox draws,
    terminal observations, statistics.  Not bitwise (branch-free sincos instead of libm, explicit FMAs): 1e-5."""
    g = load_golden("legacy")
    n, steps = 4097, 40
    flow, _ = make_flows(g, torch.float32, smooth=True)
    kw = dict(dtype=torch.float32, noiseMagCoeffs=0.1, noiseMagActuation=0.1, maxSteps=15, auto_reset=True, seed=11)
    monkeypatch.setenv("MVRL_AUV_NO_X2", "0")
    packed = AuvVecEnv(n, flow, **kw)
    o_p = packed.reset().clone()
    monkeypatch.setenv("MVRL_AUV_NO_X2", "1")
    single = AuvVecEnv(n, flow, **kw)
    assert torch.equal(o_p, single.reset())
    rng = np.random.default_rng(12)
    agree = []
    for k in range(steps):
        a = torch.as_tensor(rng.uniform(-1, 1, (n, 3)), dtype=torch.float32, device=DEV)
        single._state.copy_(packed._state); single._err_o.copy_(packed._err_o); single._recent.copy_(packed._recent)
        single._mults.copy_(packed._mults); single._target.copy_(packed._target); single._istep.copy_(packed._istep)
        single._ep_return.copy_(packed._ep_return); single._episode.copy_(packed._episode)
        op, rp, dp, ip = packed.step(a)
        os_, rs, ds, is_ = single.step(a)
        same = dp == ds                       # an environment within rounding of a boundary may terminate on one side only
        agree.append(float(same.float().mean()))
        assert float((op - os_).abs()[same].max()) < 1e-5 and float((rp - rs).abs()[same].max()) < 1e-4, k
        both = same & dp
        if bool(both.any()):
            assert float((ip["terminal_observation"] - is_["terminal_observation"]).abs()[both].max()) < 1e-5
            assert torch.equal(packed._mults[:, :n][:, both], single._mults[:, :n][:, both])      # same Philox draws
    assert min(agree) > 0.999
    assert packed.episode_stats()["episodes"] > 0
