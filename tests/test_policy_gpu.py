"""GPU: the fused rollout actor (csrc/mvrl_policy.cu) against a plain PyTorch fp32 restatement of the same network."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from marinevehiclereinforcementlearning_b200 import BlueROV2Heavy6DoFVecEnv, MlpGaussianPolicy

DEV = "cuda"


def reference_forward(pol, obs_nk, round_bf16=True):
    """Plain PyTorch restatement of the network (the test's reference; the product has no PyTorch path): mean [N, act_dim].
    ``round_bf16=True`` mirrors the kernel's arithmetic (bf16 operands, tanh-form GELU); ``False`` is the reference's
    network as SB3 builds it: fp32 with ``torch.nn.GELU()`` (the erf form, legacy/main_00_sbl.py:100-105)."""
    r = (lambda t: t.to(torch.bfloat16).to(torch.float32)) if round_bf16 else (lambda t: t)
    approx = "tanh" if round_bf16 else "none"
    h = r(obs_nk.to(torch.float32))
    for i in range(3):
        h = r(torch.nn.functional.gelu(h @ r(pol.weights[i].to(h.device)).T + pol.biases[i].to(h.device), approximate=approx))
    return torch.tanh(h @ r(pol.weights[3].to(h.device)).T + pol.biases[3].to(h.device))


def _buffers(pol, n, seed=0):
    ld = (n + 31) // 32 * 32
    g = torch.Generator(device=DEV).manual_seed(seed)
    obs = torch.zeros((pol.obs_dim, ld), device=DEV)
    obs[:, :n] = torch.rand((pol.obs_dim, n), generator=g, device=DEV) * 2 - 1
    z = lambda k: torch.full((k, ld), 7.0, device=DEV)
    return obs, z(pol.act_dim), z(pol.act_dim), z(pol.act_dim), torch.full((ld,), 7.0, device=DEV)


@pytest.mark.parametrize("n", [1, 37, 128, 5000, 131072])
def test_actor_matches_pytorch_reference(n):
    pol = MlpGaussianPolicy(9, 6, device=DEV, seed=3)
    for b in pol.biases:   # non-zero biases so that they are exercised
        b.uniform_(-0.3, 0.3, generator=torch.Generator().manual_seed(5))
    pol.log_std = torch.tensor([-0.5, -0.7, -0.2, -1.0, 0.1, -0.4])
    pol.sync_weights()
    obs, act, mean, eps, logp = _buffers(pol, n)
    pol.act_into(obs, act, n, logp=logp, mean=mean, eps=eps, env_id0=11, step=4)
    x = obs[:, :n].T
    ref_bf16 = reference_forward(pol, x, round_bf16=True)       # same operand rounding as the kernel: bf16 inputs, fp32 accumulate
    ref_fp32 = reference_forward(pol, x, round_bf16=False)      # the reference's network: plain fp32, erf-form GELU (torch.nn.GELU())
    got = mean[:, :n].T
    assert float((got - ref_bf16).abs().max()) < 6e-3          # tanh.approx (2^-11) + accumulation order
    assert float((got - ref_fp32).abs().max()) < 4e-2          # bf16 operands (3 significant digits) + tanh-form GELU
    e = eps[:, :n].T
    std = pol.log_std.exp().to(DEV)
    assert torch.allclose(act[:, :n].T, (got + std * e).clamp(-1, 1), atol=1e-6)
    assert torch.allclose(logp[:n], -0.5 * (e * e).sum(1) + pol.logp_const, atol=1e-5)
    # nothing written beyond n (rows n .. ld keep the fill value)
    assert bool((act[:, n:] == 7.0).all()) and bool((logp[n:] == 7.0).all())
    if n >= 5000:   # the draws are standard normal
        assert abs(float(e.mean())) < 0.02 and abs(float(e.std()) - 1.0) < 0.02
        assert abs(float((e ** 3).mean())) < 0.08 and abs(float((e ** 4).mean()) - 3.0) < 0.2


def test_actor_draws_do_not_depend_on_sharding_and_deterministic_mode():
    pol = MlpGaussianPolicy(9, 6, device=DEV, seed=9)
    n = 1000
    obs, act, mean, eps, logp = _buffers(pol, n, seed=2)
    pol.act_into(obs, act, n, eps=eps, env_id0=0, step=7)
    lo = 512                                                    # second shard: environments 512 .. 999 with env_id0 = 512
    obs2 = obs[:, lo:lo + 512].contiguous()                     # ld 512 (488 real environments)
    act2, eps2 = torch.zeros_like(obs2[:6]), torch.zeros_like(obs2[:6])
    pol.act_into(obs2, act2, n - lo, eps=eps2, env_id0=lo, step=7)
    assert torch.equal(eps2[:, :n - lo], eps[:, lo:n]) and torch.equal(act2[:, :n - lo], act[:, lo:n])
    pol.act_into(obs, act, n, mean=mean, step=8, deterministic=True)
    assert torch.equal(act[:, :n], mean[:, :n])
    a, _ = pol.predict(obs[:, :5].T.cpu().numpy(), deterministic=True)
    assert a.shape == (5, 6) and np.allclose(a, mean[:, :5].T.cpu().numpy(), atol=1e-6)
    a1, _ = pol.predict(np.zeros(9, dtype=np.float32), deterministic=True)
    assert a1.shape == (6,)


def test_rollout_step_is_actor_plus_env_step():
    """Closed loop on the env's own buffers: actor writes env._action, the fused step reads it; 20 steps, finite, actions in range."""
    n = 4096
    env = BlueROV2Heavy6DoFVecEnv(n, action_mode="setpoint", dtype=torch.float32, device=DEV, maxSteps=8, auto_reset=True, seed=1)
    pol = MlpGaussianPolicy(env.lenObs, env.lenAction, device=DEV, seed=1)
    env.reset()
    logp = torch.empty(env.ld, device=DEV)
    for k in range(20):
        pol.act(env, logp=logp)
        obs, rew, done, _ = env.step()
        assert float(env.actions_fm.abs().max()) <= 1.0
    assert bool(torch.isfinite(env.systemState).all()) and bool(torch.isfinite(logp[:n]).all())
    assert env.episode_stats()["episodes"] == 2 * n


@pytest.mark.parametrize("obs_dim,act_dim", [(9, 6), (16, 8), (5, 3), (1, 1)])
def test_tcgen05_actor_agrees_with_the_warp_level_mma_kernel(monkeypatch, obs_dim, act_dim):
    """Two independent implementations of the same network in the library - the default tcgen05 kernel (operands through
    shared-memory descriptors, accumulator in tensor memory, biases added by the tensor core) and the mma.sync kernel
    (register fragments; MVRL_POLICY_MMA_SYNC=1) - on the same inputs: identical Philox draws, means within the fp32
    summation-order difference, log-probs equal; sizes around the 128-row tile and more tiles than tile groups."""
    for n in (1, 127, 129, 1000, 80000):
        pols = []
        for flag in ("0", "1"):
            monkeypatch.setenv("MVRL_POLICY_MMA_SYNC", flag)
            p = MlpGaussianPolicy(obs_dim, act_dim, device=DEV, seed=3)
            for b in p.biases:
                b.uniform_(-0.3, 0.3, generator=torch.Generator().manual_seed(5))
            p.log_std = torch.linspace(-1.0, 0.1, act_dim)
            p.sync_weights()
            pols.append(p)
        out = []
        for p in pols:
            obs, act, mean, eps, logp = _buffers(p, n, seed=1)
            p.act_into(obs, act, n, logp=logp, mean=mean, eps=eps, env_id0=5, step=9)
            out.append((act, mean, eps, logp))
        (a5, m5, e5, l5), (as_, ms, es, ls) = out
        assert torch.equal(e5, es)                                             # same noise, bit for bit (also the untouched padding)
        assert float((m5[:, :n] - ms[:, :n]).abs().max()) < 2e-3               # bf16 re-rounding of activations that differ in the last fp32 bits
        assert float((m5[:, :n] - ms[:, :n]).abs().mean()) < 1e-4
        assert torch.allclose(l5[:n], ls[:n], atol=1e-6)
        assert bool((a5[:, n:] == 7.0).all()) and bool((l5[n:] == 7.0).all())  # nothing written beyond n
