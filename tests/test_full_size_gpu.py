"""Parity at BASELINE.json's FULL size (config 3: 1 048 576 environments).

The element-wise trajectory tests run at sizes the oracle finishes in seconds
(test_rov6_gpu.py).  At the full batch the CUDA path is checked
  * element-wise against the C oracle for a few env steps (the C port does
    ~1e6 env-steps/s, so 1 Mi envs x 3 steps is a few seconds), fp64 and fp32;
  * through size-independent properties: the batch is a 256-fold tiling of a
    4096-env problem, so every replica must equal replica 0 BITWISE (an
    indexing / tail / alignment error anywhere in the grid breaks it), the
    4096-env run itself is bitwise equal to the stand-alone 4096-env launch,
    and two half-size shards concatenate to the full launch bitwise, including
    the auto-reset draws keyed on the global environment id.
"""
import numpy as np
import pytest
import torch

from oracle import oracle_np as o

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from marinevehiclereinforcementlearning_b200 import BlueROV2Heavy6DoFVecEnv

DEV = "cuda"
N_FULL = 1 << 20


def make_env(n, dtype, **kw):
    kw.setdefault("auto_reset", False)
    kw.setdefault("maxSteps", 10 ** 9)
    return BlueROV2Heavy6DoFVecEnv(n, action_mode="rpm", dtype=dtype, device=DEV, **kw)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 1e-4)])
def test_full_batch_three_steps_vs_c_oracle(dtype, tol):
    from oracle import c_oracle as c
    n, steps = N_FULL, 3
    gen = torch.Generator(device="cpu").manual_seed(4321)
    env = make_env(n, dtype, maxSteps=2, auto_reset=True, seed=5)
    ref = c.Rov6EnvC(n, mode=o.MODE_RPM, max_steps=2, auto_reset=True, seed=5)
    o0 = env.reset().cpu().numpy()
    r0 = ref.reset()
    assert np.abs(o0 - r0).max() < (1e-12 if dtype == torch.float64 else 1e-6)
    for k in range(steps):
        a = (torch.rand((n, 8), generator=gen, dtype=torch.float64) * 2 - 1) * 3500.0
        obs, rew, done, _ = env.step(a.to(DEV, dtype))
        # the oracle sees the actions the kernel saw (fp32 rounding of the rpm is part of the input, not of the error)
        ro, rr, rd, _ = ref.step(a.to(dtype).to(torch.float64).numpy())
        d = np.abs(env.systemState.cpu().numpy().astype(np.float64) - ref.state)
        d[:, 3:6] = np.abs((d[:, 3:6] + np.pi) % (2 * np.pi) - np.pi)
        scale = 1.0 + np.abs(ref.state)
        assert (d / scale).max() < tol, (k, (d / scale).max())
        assert np.array_equal(done.cpu().numpy(), rd)
        assert np.abs(obs.cpu().numpy().astype(np.float64) - ro).max() < max(tol, 1e-6 if dtype == torch.float32 else 0)
        assert float(rew.abs().max()) == 0.0
    st = env.episode_stats()
    assert st["episodes"] == n and st["mean_length"] == 2.0 and st["nonfinite"] == 0


def test_full_batch_is_a_bitwise_tiling_of_4096_envs():
    base, reps, steps = 4096, N_FULL // 4096, 20
    rng = np.random.default_rng(11)
    s0 = np.zeros((base, 12))
    s0[:, 3:6] = rng.uniform(0, 2 * np.pi, (base, 3))
    s0[:, 6:12] = rng.uniform(-1, 1, (base, 6))
    small = make_env(base, torch.float32)
    full = make_env(N_FULL, torch.float32)
    small.reset(initialSetpoint=np.zeros(6))
    full.reset(initialSetpoint=np.zeros(6))
    s0_fm = torch.as_tensor(np.ascontiguousarray(s0.T), device=DEV, dtype=torch.float32)      # feature-major [12, base]
    small._state[:, :base].copy_(s0_fm)
    full._state[:, :N_FULL].copy_(s0_fm.repeat(1, reps))
    for k in range(steps):
        a = torch.as_tensor(rng.uniform(-3500, 3500, (base, 8)), device=DEV, dtype=torch.float32)
        os_, _, _, _ = small.step(a)
        of, _, _, _ = full.step(a.repeat(reps, 1))
    sf = full.systemState.reshape(reps, base, 12)
    assert torch.equal(sf[0], small.systemState)
    assert bool((sf == sf[0:1]).all())
    assert torch.equal(of.reshape(reps, base, 9)[reps - 1], os_)
    assert bool(torch.isfinite(sf).all())


def test_full_batch_two_shards_bitwise():
    n, steps = N_FULL, 6
    gen = torch.Generator(device=DEV).manual_seed(77)
    full = make_env(n, torch.float32, maxSteps=4, auto_reset=True, seed=3)
    lo = make_env(n // 2, torch.float32, maxSteps=4, auto_reset=True, seed=3, env_id0=0)
    hi = make_env(n // 2, torch.float32, maxSteps=4, auto_reset=True, seed=3, env_id0=n // 2)
    of = full.reset(); ol = lo.reset(); oh = hi.reset()
    assert torch.equal(of, torch.cat([ol, oh]))
    for k in range(steps):
        a = (torch.rand((n, 8), generator=gen, device=DEV) * 2 - 1) * 3500.0
        of, _, df, _ = full.step(a)
        ol, _, dl, _ = lo.step(a[: n // 2])
        oh, _, dh, _ = hi.step(a[n // 2:])
        assert torch.equal(of, torch.cat([ol, oh])) and torch.equal(df, torch.cat([dl, dh]))
    assert torch.equal(full.systemState, torch.cat([lo.systemState, hi.systemState]))
    assert torch.equal(full.path, torch.cat([lo.path, hi.path]))
    sf, sl, sh = full.episode_stats(), lo.episode_stats(), hi.episode_stats()
    assert sf["episodes"] == sl["episodes"] + sh["episodes"] == n
