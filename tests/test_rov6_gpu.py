"""GPU parity tests of the 6DoF path: CUDA kernels (through the C ABI) vs the
numpy oracle and vs the golden vectors produced by the unmodified reference.

Tolerances (BASELINE.json north_star): derivatives 1e-10 relative in fp64;
1000-step trajectories 1e-8 in fp64 and 1e-4 in fp32.  "Relative" is measured
per environment against |ref| + max|ref| so that components which cancel to
~0 do not blow the ratio up.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import oracle_np as o

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from marinevehiclereinforcementlearning_b200 import BlueROV2Heavy6DoFVecEnv, Rov6Constants, Rov6Derivs
    from marinevehiclereinforcementlearning_b200 import resources as res

DEV = "cuda"


def rel_err(a, ref):
    ref = np.asarray(ref)
    scale = np.abs(ref) + np.abs(ref).max(axis=-1, keepdims=True)
    return (np.abs(np.asarray(a) - ref) / np.maximum(scale, 1e-300)).max()


def fm(x, dtype=torch.float64):
    """[N, k] numpy -> feature-major [k, N] CUDA tensor."""
    return torch.as_tensor(np.ascontiguousarray(np.asarray(x).T), dtype=dtype, device=DEV)


def sample_states(rng, n):
    s = np.empty((n, 12))
    s[:, 0:3] = rng.uniform(-5, 5, (n, 3))
    s[:, 3:6] = rng.uniform(-np.pi, np.pi, (n, 3))
    s[:, 6:9] = rng.uniform(-2, 2, (n, 3))
    s[:, 9:12] = rng.uniform(-5, 5, (n, 3))
    return s


# ------------------------------------------------------------------ K2 ------
def test_derivs_rpm_fp64_vs_golden_and_oracle():
    g = load_golden("rov6")
    f = Rov6Derivs(dtype=torch.float64, action_mode="rpm")
    d, aux = f(fm(g["rpm_states"]), fm(g["rpm_rpms"]), want_aux=True)
    assert rel_err(d.T.cpu().numpy(), g["rpm_derivs"]) < 1e-10
    aux = aux.T.cpu().numpy()
    assert rel_err(aux[:, 0:6], g["rpm_RHS"]) < 1e-10
    comp = aux[:, 20:50].reshape(-1, 5, 6).transpose(0, 2, 1)  # [N, 6, 5] like forceModel(retComp=True)
    assert np.abs(comp - g["rpm_retComp"]).max() < 1e-9
    # config 2: 1e5 random states, eta ~ U(-pi, pi), nu ~ U(-2,2) / U(-5,5)
    rng = np.random.default_rng(11)
    s = sample_states(rng, 100_000)
    rpm = rng.uniform(-4000, 4000, (100_000, 8))
    ref = o.derivs6_rpm(o.Rov6Params(), s, rpm)
    got = f(fm(s), fm(rpm)).T.cpu().numpy()
    assert rel_err(got, ref) < 1e-10


def test_derivs_force_fp64():
    g = load_golden("rov6")
    f = Rov6Derivs(dtype=torch.float64, action_mode="force")
    d, aux = f(fm(g["force_states"]), fm(g["force_forces"]), want_aux=True)
    assert rel_err(d.T.cpu().numpy(), g["force_derivs"]) < 1e-10
    assert np.abs(aux[12:20].T.cpu().numpy() - g["force_cv"]).max() < 1e-8
    rng = np.random.default_rng(12)
    s = sample_states(rng, 50_000)
    frc = rng.uniform(-1, 1, (50_000, 6)) * np.array([50., 50., 50., 2., 2., 2.])
    ref = o.derivs6_force(o.Rov6Params(), s, frc)
    assert rel_err(f(fm(s), fm(frc)).T.cpu().numpy(), ref) < 1e-10


def test_derivs_pid_sequences_fp64():
    g = load_golden("rov6")
    f = Rov6Derivs(dtype=torch.float64, action_mode="setpoint")
    n_env, n_call = g["pid_t"].shape
    ctrl = Rov6Derivs.new_ctrl(n_env)
    sp = fm(g["pid_sp"])
    for c in range(n_call):
        d, aux = f(fm(g["pid_states"][:, c]), t=torch.as_tensor(g["pid_t"][:, c], device=DEV), setpoint=sp, ctrl=ctrl, want_aux=True)
        assert rel_err(d.T.cpu().numpy(), g["pid_derivs"][:, c]) < 1e-9, c
        assert np.abs(aux[6:12].T.cpu().numpy() - g["pid_gcf"][:, c]).max() < 1e-9, c
        assert np.abs(aux[12:20].T.cpu().numpy() - g["pid_cv"][:, c]).max() < 1e-6, c
        assert np.abs(ctrl[6:12].T.cpu().numpy() - g["pid_eint"][:, c]).max() < 1e-12, c


def test_derivs_fp32_close_to_fp64():
    rng = np.random.default_rng(13)
    s = sample_states(rng, 20_000)
    rpm = rng.uniform(-4000, 4000, (20_000, 8))
    ref = o.derivs6_rpm(o.Rov6Params(), s, rpm)
    got = Rov6Derivs(dtype=torch.float32, action_mode="rpm")(fm(s, torch.float32), fm(rpm, torch.float32)).T.cpu().numpy()
    # avoid the 1/cos(theta) pole for the fp32 statement
    ok = np.abs(np.cos(s[:, 4])) > 0.05
    assert rel_err(got[ok], ref[ok]) < 2e-5


def test_generic_kernel_matches_specialised():
    c = Rov6Constants()
    c.CG = np.array([1e-300, 0., 0.05])  # breaks the default-sparsity test, changes nothing numerically
    g = load_golden("rov6")
    fg = Rov6Derivs(consts=c, dtype=torch.float64, action_mode="rpm")
    fs = Rov6Derivs(dtype=torch.float64, action_mode="rpm")
    dg = fg(fm(g["rpm_states"]), fm(g["rpm_rpms"]))
    ds = fs(fm(g["rpm_states"]), fm(g["rpm_rpms"]))
    assert not fg._get_handle().specialised and fs._get_handle().specialised
    assert rel_err(dg.T.cpu().numpy(), g["rpm_derivs"]) < 1e-10
    assert rel_err(dg.T.cpu().numpy(), ds.T.cpu().numpy()) < 1e-13
    # a genuinely different vehicle: off-axis CG, cross damping, inertia products, net buoyancy
    p = o.Rov6Params(CG=np.array([0.01, -0.02, 0.05]), Yr=-0.3, Nv=-0.2, Kvv=-0.4, Zq=-0.1, Mw=0.2, m=11.0,
                     I=np.array([[0.16, 0.01, -0.02], [0.01, 0.17, 0.005], [-0.02, 0.005, 0.18]]))
    p.dispVol = 11.4 / 1000.
    c = Rov6Constants()
    c.CG, c.Yr, c.Nv, c.Kvv, c.Zq, c.Mw, c.m, c.I = p.CG, p.Yr, p.Nv, p.Kvv, p.Zq, p.Mw, p.m, p.I
    got = Rov6Derivs(consts=c, dtype=torch.float64, action_mode="rpm")(fm(g["rpm_states"]), fm(g["rpm_rpms"]))
    assert rel_err(got.T.cpu().numpy(), o.derivs6_rpm(p, g["rpm_states"], g["rpm_rpms"])) < 1e-10


# ------------------------------------------------------------------ K1 ------
def make_env(n, mode, dtype=torch.float64, **kw):
    kw.setdefault("auto_reset", False)
    kw.setdefault("maxSteps", 10 ** 9)
    return BlueROV2Heavy6DoFVecEnv(n, action_mode=mode, dtype=dtype, device=DEV, **kw)


def test_trajectory_1000_steps_fp64_vs_reference_golden():
    t = load_golden("traj6")
    env = make_env(4, "rpm", n_sub=int(t["n_sub"]), dt=float(t["dt"]))
    env.reset(initialSetpoint=np.zeros(6))
    worst = 0.0
    for k in range(t["actions"].shape[0]):
        env.step(torch.as_tensor(t["actions"][k], device=DEV))
        worst = max(worst, np.abs(env.systemState.cpu().numpy() - t["traj"][k]).max())
    assert worst < 1e-8, worst
    envf = make_env(2, "force", n_sub=int(t["n_sub"]), dt=float(t["dt"]))
    envf.reset(initialSetpoint=np.zeros(6))
    worst = 0.0
    for k in range(t["force_actions"].shape[0]):
        envf.step(torch.as_tensor(t["force_actions"][k], device=DEV))
        worst = max(worst, np.abs(envf.systemState.cpu().numpy() - t["force_traj"][k]).max())
    assert worst < 1e-8, worst


def test_trajectory_4096_envs_1000_steps_fp64_vs_oracle():
    """BASELINE config 2: 4096 envs, fp64, nSub 8, rpm ~ U(-3500, 3500) from
    torch.Generator(seed 1234) regenerated every step, initial state 0, 1000
    steps, element-wise against the C oracle (tolerance 1e-8).

    Conditioning: random full-throttle thrusters drive some vehicles through
    the pitch pole of J2 (1/cos(theta), resources.py:116-131).  Within
    |cos(theta)| < 1e-2 of it round-off is amplified by >1e4 per stage and no
    two fp64 implementations agree afterwards - measured here between the
    numpy and the C oracle, which are bit-faithful restatements of the same
    reference code: 1.4e-13 worst disagreement over 250 steps for envs that
    stay outside that band, up to 1e-1 inside it.  Each environment is
    therefore compared up to its first stage inside the band (the oracle
    tracks min |cos(theta)| over every RK4 stage); a large majority must stay
    outside for the whole run."""
    from oracle import c_oracle as c
    n, steps = 4096, 1000
    gen = torch.Generator(device="cpu").manual_seed(1234)
    env = make_env(n, "rpm")
    env.reset(initialSetpoint=np.zeros(6))
    ref = c.Rov6EnvC(n, mode=o.MODE_RPM, max_steps=10 ** 9)
    ref.reset(initial_setpoint=np.zeros(6))

    def adiff(a, b):
        d = np.abs(a - b)
        d[:, 3:6] = np.abs((d[:, 3:6] + np.pi) % (2 * np.pi) - np.pi)
        return d.max(axis=1)

    worst = 0.0
    for k in range(steps):
        a = (torch.rand((n, 8), generator=gen, dtype=torch.float64) * 2 - 1) * 3500.0
        obs, rew, done, _ = env.step(a.to(DEV))
        ro, _, _, _ = ref.step(a.numpy())
        if k % 10 == 9 or k == steps - 1:
            good = ref.mincos >= 1e-2
            worst = max(worst, adiff(env.systemState.cpu().numpy(), ref.state)[good].max())
            assert np.abs(obs.cpu().numpy() - ro)[good].max() < 1e-8, k
    good = ref.mincos >= 1e-2
    print("envs compared for all %d steps: %d of %d; worst error: %.3e" % (steps, good.sum(), n, worst))
    assert good.sum() >= 0.75 * n
    assert worst < 1e-8, worst
    assert float(rew.abs().max()) == 0.0 and not bool(done.any())
    assert bool(torch.isfinite(env.systemState).all())


def test_trajectory_4096_envs_1000_steps_fp32_vs_oracle():
    """The fp32 kernel against the fp64 C oracle on the config-2 workload (4096 envs, 1000 steps, rpm ~ U(-3500, 3500),
    the oracle sees the fp32-rounded actions): tolerance 1e-4 on the state scaled by 1 + |ref| (BASELINE north_star).

    Conditioning as in the fp64 test, with a wider band because fp32 round-off is 1e9 times larger: an environment is
    compared up to its first RK4 stage with |cos(theta)| < band (inside, 1/cos(theta) amplifies a 6e-8 rounding of
    theta).  Full-throttle random thrusters tumble the vehicles, so the set that never enters a band shrinks with time -
    that is the workload, not the kernel (the oracle alone decides it): 75 % / 23 % / 5.4 % of the 4096 environments
    stay outside |cos| < 0.3 for 100 / 500 / 1000 steps, 94 % / 72 % / 51 % outside |cos| < 0.1.  Both fractions are
    asserted, with 1e-4 for the first band and 1e-3 for the second (measured on B200: 5.5e-5 and 4e-4).  Every
    environment, tumbling or not, is covered step by step by test_parity_modes_gpu.py::test_rov6_one_step_local_error_all_envs."""
    from oracle import c_oracle as c
    n, steps = 4096, 1000
    gen = torch.Generator(device="cpu").manual_seed(1234)
    env = make_env(n, "rpm", dtype=torch.float32)
    env.reset(initialSetpoint=np.zeros(6))
    ref = c.Rov6EnvC(n, mode=o.MODE_RPM, max_steps=10 ** 9)
    ref.reset(initial_setpoint=np.zeros(6))
    worst = {0.3: 0.0, 0.1: 0.0}
    frac = {}
    for k in range(steps):
        a = ((torch.rand((n, 8), generator=gen, dtype=torch.float64) * 2 - 1) * 3500.0).to(torch.float32)
        env.step(a.to(DEV))
        ref.step(a.to(torch.float64).numpy())
        if k % 10 == 9 or k == steps - 1:
            d = np.abs(env.systemState.cpu().numpy().astype(np.float64) - ref.state)
            d[:, 3:6] = np.abs((d[:, 3:6] + np.pi) % (2 * np.pi) - np.pi)
            d = (d / (1.0 + np.abs(ref.state))).max(axis=1)
            for band in worst:
                good = ref.mincos >= band
                worst[band] = max(worst[band], d[good].max())
                if k + 1 in (100, 500, 1000):
                    frac[(band, k + 1)] = float(good.mean())
    print("fp32 rpm trajectories: fraction of %d envs compared for 100 / 500 / 1000 steps and worst scaled error - outside |cos| < 0.3: "
          "%.3f / %.3f / %.3f, %.3e; outside |cos| < 0.1: %.3f / %.3f / %.3f, %.3e" %
          (n, frac[(0.3, 100)], frac[(0.3, 500)], frac[(0.3, 1000)], worst[0.3], frac[(0.1, 100)], frac[(0.1, 500)], frac[(0.1, 1000)], worst[0.1]))
    assert frac[(0.3, 100)] >= 0.70 and frac[(0.3, 500)] >= 0.20 and frac[(0.3, 1000)] >= 0.045, frac
    assert frac[(0.1, 100)] >= 0.90 and frac[(0.1, 500)] >= 0.65 and frac[(0.1, 1000)] >= 0.45, frac
    assert worst[0.3] < 1e-4, worst
    assert worst[0.1] < 1e-3, worst


def test_trajectory_fp32_within_1e4():
    """fp32 kernel vs the fp64 oracle over 1000 steps (tolerance 1e-4 on the
    state, angles compared modulo 2 pi)."""
    t = load_golden("traj6")
    for fast in (False, True):
        env = make_env(4, "rpm", dtype=torch.float32, n_sub=int(t["n_sub"]), dt=float(t["dt"]), fast_math=fast)
        env.reset(initialSetpoint=np.zeros(6))
        worst = 0.0
        for k in range(t["actions"].shape[0]):
            env.step(torch.as_tensor(t["actions"][k], device=DEV, dtype=torch.float32))
            d = env.systemState.cpu().numpy().astype(np.float64) - t["traj"][k]
            d[:, 3:6] = (d[:, 3:6] + np.pi) % (2 * np.pi) - np.pi
            scale = 1.0 + np.abs(t["traj"][k])
            worst = max(worst, (np.abs(d) / scale).max())
        print("fp32 fast=%s worst scaled error over 1000 steps: %.3e" % (fast, worst))
        assert worst < (1e-4 if not fast else 1e-3), (fast, worst)


def test_env_semantics_fixed_setpoint_vs_reference_env():
    e6 = load_golden("env6")
    env = make_env(1, "setpoint", maxSteps=60, record_aux=True)
    obs = [env.reset(initialSetpoint=e6["fixed_sp"]).cpu().numpy()[0]]
    hist = []
    for k in range(60):
        ob, r, d, _ = env.step(torch.zeros((1, 6), dtype=torch.float64, device=DEV))
        obs.append(ob.cpu().numpy()[0])
        hist.append(np.concatenate([[float(env.time[0])], env.systemState.cpu().numpy()[0], env._aux[:, 0].cpu().numpy(),
                                    env.setPoint.cpu().numpy()[0]]))
        assert bool(d[0]) == bool(e6["fixed_done"][k]) and float(r[0]) == 0.0
    assert np.abs(np.array(obs) - e6["fixed_obs"]).max() < 1e-8
    ref = e6["fixed_history"][1:]
    got = np.array(hist)
    assert np.abs(got[:, :13] - ref[:, :13]).max() < 1e-8          # t + 12 states
    assert np.abs(got[:, 13:19] - ref[:, 13:19]).max() < 1e-5      # controller forces of the last stage
    assert np.abs(got[:, 27:] - ref[:, 27:]).max() < 1e-12         # set-point


def test_env_semantics_action_driven_vs_reference_env():
    e6 = load_golden("env6")
    env = make_env(1, "setpoint", maxSteps=40)
    env.reset(initialSetpoint=np.append(e6["act_path"][0], e6["act_orient"]))
    env._path[:, 0] = torch.as_tensor(e6["act_path"].reshape(-1), device=DEV)
    env.fixedSp = False
    for k in range(40):
        ob, r, d, _ = env.step(torch.as_tensor(e6["act_actions"][k:k + 1], device=DEV))
        assert np.abs(ob.cpu().numpy()[0] - e6["act_obs"][k + 1]).max() < 1e-7, k
        assert np.abs(env.systemState.cpu().numpy()[0] - e6["act_history"][k + 1, 1:13]).max() < 1e-7, k
        assert bool(d[0]) == bool(e6["act_done"][k])


def test_auto_reset_and_sharding_bitwise():
    """Auto-reset (terminal_observation, Philox draws) vs the oracle, and
    N envs in one launch == two shards with env_id0 offsets, bitwise."""
    n, steps = 512, 12
    rng = np.random.default_rng(3)
    acts = rng.uniform(-3500, 3500, (steps, n, 8))
    full = make_env(n, "rpm", maxSteps=5, auto_reset=True, seed=99)
    a = make_env(n // 2, "rpm", maxSteps=5, auto_reset=True, seed=99, env_id0=0)
    b = make_env(n // 2, "rpm", maxSteps=5, auto_reset=True, seed=99, env_id0=n // 2)
    ref = o.Rov6EnvOracle(n, mode=o.MODE_RPM, max_steps=5, auto_reset=True, seed=99)
    o0 = full.reset().cpu().numpy()
    a.reset(); b.reset()
    r0 = ref.reset()
    assert np.abs(o0 - r0).max() < 1e-12
    for k in range(steps):
        act = torch.as_tensor(acts[k], device=DEV)
        obs, rew, done, info = full.step(act)
        oa, _, da, _ = a.step(act[: n // 2])
        ob, _, db, _ = b.step(act[n // 2:])
        assert torch.equal(obs, torch.cat([oa, ob])) and torch.equal(done, torch.cat([da, db]))
        assert torch.equal(full.systemState, torch.cat([a.systemState, b.systemState]))
        ro, rr, rd, rinfo = ref.step(acts[k])
        assert np.array_equal(done.cpu().numpy(), rd)
        assert np.abs(obs.cpu().numpy() - ro).max() < 1e-9
        if rd.any():
            assert np.abs(info["terminal_observation"].cpu().numpy()[rd] - rinfo["terminal_observation"][rd]).max() < 1e-9
            assert np.abs(full.path.cpu().numpy().reshape(n, 6) - ref.path).max() < 1e-12
            assert (full.iStep == 0).all()
    st = full.episode_stats()
    assert st["episodes"] == n * (steps // 5) and st["mean_length"] == 5.0 and st["nonfinite"] == 0


def test_resources_helpers():
    r = load_golden("resources")
    got = res.angleError(r["angle_pairs"][:, 0], r["angle_pairs"][:, 1])
    assert np.abs(got - r["angle_err"]).max() < 1e-14
    assert res.angleError(0.1, 6.2) == pytest.approx(0.1831853071795857, abs=1e-15)
    assert np.signbit(res.angleError(1.0, 1.0)) and res.angleError(0.0, np.pi) == -np.pi
    a = r["ct_angles"]
    J6 = res.coordinateTransform(a[:, 0], a[:, 1], a[:, 2], dof=6)
    assert np.abs(J6 - r["ct_J6"]).max() / np.abs(r["ct_J6"]).max() < 1e-12
    # away from the clamped pole the agreement is at round-off level
    ok = np.abs(np.cos(a[:, 1])) > 1e-3
    assert np.abs(J6[ok] - r["ct_J6"][ok]).max() < 1e-12
    assert np.abs(res.coordinateTransform(a[:, 0], a[:, 1], a[:, 2], dof=3) - r["ct_J3"]).max() < 1e-15
    assert res.coordinateTransform(0.1, 0.2, 0.3, dof=6).shape == (6, 6)
    assert res.coordinateTransform(0.1, 0.2, 0.3).shape == (3, 3)


def test_step_host_and_step_range_match_step_bitwise():
    """The host-buffer pipeline (chunked upload / transpose / step / download) and the range launch
    give exactly what one whole-batch launch gives, including auto-reset draws (global env ids)."""
    n, steps = 5000, 7   # not a multiple of the chunk or tile size
    for dtype, mode, na in ((torch.float32, "rpm", 8), (torch.float64, "setpoint", 6)):
        rng = np.random.default_rng(41)
        scale = 3500.0 if mode == "rpm" else 1.0
        acts = torch.as_tensor(rng.uniform(-scale, scale, (steps, n, na)), dtype=dtype)
        kw = dict(action_mode=mode, dtype=dtype, device=DEV, maxSteps=3, auto_reset=True, seed=5)
        a, b, c = (BlueROV2Heavy6DoFVecEnv(n, **kw) for _ in range(3))
        a.reset(); b.reset(); c.reset()
        for k in range(steps):
            obs, rew, done, _ = a.step(acts[k].to(DEV))
            if k % 3 == 2:   # pageable host memory: copy-engine path with device staging
                outs = (torch.empty((n, 9), dtype=dtype), torch.empty(n, dtype=dtype), torch.empty(n, dtype=torch.uint8))
                h_obs, h_rew, h_done = b.step_host(acts[k].clone(), *outs, chunks=3)
            else:            # pinned: overlapping copies; CUDA-graph replay / direct streams
                h_obs, h_rew, h_done = b.step_host(acts[k].pin_memory(), chunks=3 if k % 2 else -3)
            c.set_actions(acts[k].to(DEV))
            for first in range(0, n, 1111):
                c.step_range_async(first, min(1111, n - first))
            assert torch.equal(obs.cpu(), h_obs) and torch.equal(rew.cpu(), h_rew) and torch.equal(done.cpu(), h_done.bool())
            assert torch.equal(a._state, b._state) and torch.equal(a._state, c._state) and torch.equal(a._obs, c._obs)
            assert torch.equal(a._path, b._path) and torch.equal(a._path, c._path)


def test_two_envs_per_thread_kernel_matches_one_env_kernel_bitwise(monkeypatch):
    """fp32: the packed (F2, FFMA2) instantiation and the one-env-per-thread instantiation run the same
    arithmetic per environment - bitwise equal states / observations, odd batch size, every action mode,
    with auto-reset, for the accurate and the fast-math variants."""
    n, steps = 4097, 12
    for mode, na, scale in (("rpm", 8, 3500.0), ("force", 6, 40.0), ("setpoint", 6, 1.0)):
        for fast in (False, True):
            rng = np.random.default_rng(51)
            acts = torch.as_tensor(rng.uniform(-scale, scale, (steps, n, na)), dtype=torch.float32, device=DEV)
            kw = dict(action_mode=mode, dtype=torch.float32, device=DEV, maxSteps=5, auto_reset=True, seed=3, fast_math=fast)
            monkeypatch.setenv("MVRL_NO_X2", "0")
            packed = BlueROV2Heavy6DoFVecEnv(n, **kw)
            packed.reset()
            monkeypatch.setenv("MVRL_NO_X2", "1")
            single = BlueROV2Heavy6DoFVecEnv(n, **kw)
            single.reset()
            for k in range(steps):
                op, _, dp, _ = packed.step(acts[k])
                os_, _, ds, _ = single.step(acts[k])
                assert torch.equal(op, os_) and torch.equal(dp, ds), (mode, fast, k)
                nn = lambda t: torch.nan_to_num(t, nan=12345.0)   # ctrl[0] = NaN marks a fresh controller
                assert torch.equal(packed._state, single._state) and torch.equal(nn(packed._ctrl), nn(single._ctrl)), (mode, fast, k)
            assert packed.episode_stats() == single.episode_stats()
