"""Bitwise test of the rejected TMA-fed auv_step (see auv_step_tma_kernel.cuh); needs the helpers of tests/test_auv_gpu.py
and a library built with the variant and its MVRL_AUV_TMA switch."""


@pytest.mark.parametrize("n", [4097, 100000])
def test_auv_tma_fed_kernel_is_bitwise_the_plain_kernel(monkeypatch, n):
    """MVRL_AUV_TMA=1: the persistent kernel whose inputs arrive by bulk copy (cp.async.bulk + mbarrier) runs the same
    per-environment arithmetic (auv_step_env) as the plain kernel - every array bitwise equal after 40 free-running steps
    with auto-reset, an odd batch (partial last tile, rows padded to ld) and more tiles than resident CTAs."""
    g = load_golden("legacy")
    flow, _ = make_flows(g, torch.float32, smooth=True)
    kw = dict(dtype=torch.float32, noiseMagCoeffs=0.1, noiseMagActuation=0.1, maxSteps=15, auto_reset=True, seed=11)
    monkeypatch.setenv("MVRL_AUV_TMA", "1")
    fed = AuvVecEnv(n, flow, **kw)
    monkeypatch.setenv("MVRL_AUV_TMA", "0")
    plain = AuvVecEnv(n, flow, **kw)
    assert torch.equal(fed.reset(), plain.reset())
    gen = torch.Generator(device=DEV).manual_seed(5)
    for k in range(40):
        a = torch.rand((n, 3), generator=gen, device=DEV) * 2 - 1
        of, rf, df, inf = fed.step(a)
        op, rp, dp, inp = plain.step(a)
        assert torch.equal(of, op) and torch.equal(rf, rp) and torch.equal(df, dp), k
        if bool(df.any()):
            assert torch.equal(inf["terminal_observation"], inp["terminal_observation"])
    for key in AuvVecEnv._STATE_KEYS:
        assert torch.equal(getattr(fed, key), getattr(plain, key)), key
    assert fed.episode_stats() == plain.episode_stats()
