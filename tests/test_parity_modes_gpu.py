"""GPU parity of EVERY action mode and precision of the fused step kernels against the CPU oracles
(VERDICT r1, items 1-3): rpm / earth-frame force / PID set-point (the reference's Gym semantics,
dynamicsModel_BlueROV2_Heavy_6DoF.py:43-73, 545-557; 3DoF.py:141-157, 466-475), fp32 and fp64, 6DoF and 3DoF.

Two kinds of test:

* ONE-STEP LOCAL ERROR on all environments: every step the oracle's state AND controller state are copied into the
  CUDA env, both step once, all 4096 environments are compared (fp32 <= 1e-4, fp64 <= 1e-10 on the state scaled by
  1 + |ref|).  No trajectory filter.
* FREE-RUNNING 1000-step trajectories, reporting and asserting the FRACTION of environments inside the tolerance at
  steps 100 / 500 / 1000.

Three places of the reference's model are discontinuous, and an environment that comes within rounding distance of one
of them cannot agree between two precisions; the oracles report the distance as diagnostics (never used by the algorithm):

* `mincos` - min |cos(theta)| over the RK4 stages: the 1/cos(theta) pole of J2 (resources.py:116-131);
* `margin` - the smallest non-zero |e - eOld| seen by a PID call at the same t as the call before it, where
  dedt = (e - eOld) / 1e-9 turns the SIGN of the difference into a saturated demand (6DoF.py:64, RK4 stages 1 and 3);
* `dbmargin` - relative distance of a thruster demand from the 300 rpm dead-band edge, where the thrust jumps from 0
  to 0.29 N (6DoF.py:271-275).

Measured on B200 (tools/exp/r2_parity_probe.py, gpurun_out/r2a): with the literal e - eOld the fp32 set-point kernel
flips the sign in 3.6e-3 of the env-steps and 5 % of the trajectories survive 500 steps; forming e - eOld from the pose
increments (pid6_core_dp) leaves 1.2e-4 sign flips (all at margin < 1e-8) plus 5e-5 dead-band flips per env-step; free
running, 97 % / 80 % / 50 % of the fp32 set-point trajectories stay inside 1e-4 for 100 / 500 / 1000 steps (literal form:
70 % / 5 % / 0 %), 98 % / 89 % / 77 % in force mode, 99.5 % of the fp64 set-point trajectories inside 1e-8 for 1000 steps.
"""
import numpy as np
import pytest
import torch

from oracle import c_oracle as c
from oracle import oracle_np as o

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from marinevehiclereinforcementlearning_b200 import BlueROV2Heavy3DoFVecEnv, BlueROV2Heavy6DoFVecEnv

DEV = "cuda"
MODES = {"rpm": o.MODE_RPM, "force": o.MODE_FORCE, "setpoint": o.MODE_PID}
SCALE6 = {"rpm": np.full(8, 3500.0), "force": np.array([50., 50., 50., 1., 1., 2.]), "setpoint": np.ones(6)}
TOL = {torch.float32: 1e-4, torch.float64: 1e-10}


def scaled_err(got, ref, ang):
    d = np.abs(got.astype(np.float64) - ref)
    d[:, ang] = np.abs((d[:, ang] + np.pi) % (2 * np.pi) - np.pi)
    return (d / (1.0 + np.abs(ref))).max(axis=1)


def fm(x, dtype):
    return torch.as_tensor(np.ascontiguousarray(np.asarray(x).T), dtype=dtype, device=DEV)


def sync6(env, ref, mode):
    """oracle -> CUDA env: state, set-point, way-points, counters and (set-point mode) the controller state."""
    n, dt = env.num_envs, env.dtype
    env._state[:, :n] = fm(ref.state, dt)
    env._setpoint[:, :n] = fm(ref.set_point, dt)
    env._path[:, :n] = fm(ref.path, dt)
    env._istep[:n] = torch.as_tensor(ref.i_step, dtype=torch.int32, device=DEV)
    if mode == "setpoint":
        cn = ref.ctrl_np
        e_old = cn["eOld"].copy()
        e_old[cn["has_old"] == 0, 0] = np.nan            # eOld is None
        env._ctrl[0:6, :n] = fm(e_old, dt)
        env._ctrl[6:12, :n] = fm(cn["eInt"], dt)
        env._ctrl[12, :n] = torch.as_tensor(cn["tOld"], dtype=dt, device=DEV)


def sync3(env, ref, mode):
    n, dt = env.num_envs, env.dtype
    env._state[:, :n] = fm(ref.state, dt)
    env._setpoint[:, :n] = fm(ref.set_point, dt)
    env._path[:, :n] = fm(ref.path, dt)
    env._istep[:n] = torch.as_tensor(ref.i_step, dtype=torch.int32, device=DEV)
    if mode == "setpoint":
        e_old = ref.ctrl["eOld"].copy()
        e_old[~ref.ctrl["has_old"], 0] = np.nan
        env._ctrl[0:3, :n] = fm(e_old, dt)
        env._ctrl[3:6, :n] = fm(ref.ctrl["eInt"], dt)
        env._ctrl[6, :n] = torch.as_tensor(ref.ctrl["tOld"], dtype=dt, device=DEV)


def report(name, err, tol, well, parts):
    bad = err > tol
    msg = "%s: %d env-steps, max %.3e, median %.3e; outside tol %d (%.2e)" % (name, err.size, err.max(), np.median(err), bad.sum(), bad.mean())
    for label, m in parts:
        msg += "; %s: %d env-steps, %d of them outside tol" % (label, m.sum(), (bad & m).sum())
    msg += "; well-conditioned: %d env-steps (%.5f), max %.3e" % (well.sum(), well.mean(), err[well].max())
    print(msg)


# ----------------------------------------------------------------- 6DoF: one-step local error, all environments --------
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["fp32", "fp64"])
@pytest.mark.parametrize("mode", ["rpm", "force", "setpoint"])
def test_rov6_one_step_local_error_all_envs(mode, dtype):
    n, steps, tol = 4096, 120, TOL[dtype]
    ref = c.Rov6EnvC(n, mode=MODES[mode], max_steps=10 ** 9)
    ref.reset()
    env = BlueROV2Heavy6DoFVecEnv(n, action_mode=mode, dtype=dtype, device=DEV, auto_reset=False, maxSteps=10 ** 9)
    env.reset()
    rng = np.random.default_rng(1)
    errs, oerrs, mcs, mgs, dbs = [], [], [], [], []
    for k in range(steps):
        a = torch.as_tensor(rng.uniform(-1, 1, (n, len(SCALE6[mode]))) * SCALE6[mode], dtype=dtype)
        sync6(env, ref, mode)
        ref.mincos[:] = 1.0
        ref.dbmargin[:] = np.inf
        ref.ctrl_np["margin"] = np.inf
        obs, rew, done, _ = env.step(a.to(DEV))
        ro, _, _, _ = ref.step(a.to(torch.float64).numpy())     # the oracle sees the action the kernel saw (fp32-rounded)
        errs.append(scaled_err(env.systemState.cpu().numpy(), ref.state, slice(3, 6)))
        od = np.abs(obs.cpu().numpy() - ro)
        # the heading-error observation is clipped to +-1 and changes sign where the error passes +-pi (resources.py:92-95):
        # within 1e-3 of that wrap the two precisions may legitimately sit on opposite sides
        od[:, 6:9][np.abs(np.abs(o.angle_error(ref.set_point[:, 3:6], ref.state[:, 3:6])) - np.pi) < 1e-3] = 0.0
        oerrs.append(od.max(axis=1))
        mcs.append(ref.mincos.copy()); mgs.append(ref.ctrl_np["margin"].copy()); dbs.append(ref.dbmargin.copy())
    err, oerr, mc, mg, db = map(np.concatenate, (errs, oerrs, mcs, mgs, dbs))
    near_pole = mc < 1e-2
    sign_risk = (mg < 1e-7) if mode == "setpoint" else np.zeros_like(near_pole)
    # rpm mode: both sides see the same rpm.  Otherwise the demand is a difference of O(50 N) terms compared with the
    # 0.29 N dead-band edge: 1e-4 relative (in rpm) is ~10 fp32 roundings of the terms
    db_risk = (db < 1e-4) if mode != "rpm" else np.zeros_like(near_pole)
    well = ~(near_pole | sign_risk | db_risk)
    report("6DoF %s %s" % (mode, dtype), err, tol, well,
           [("|cos theta| < 1e-2", near_pole), ("PID margin < 1e-7", sign_risk), ("dead-band distance < 1e-4", db_risk)])
    # every well-conditioned env-step agrees - and they are (almost) all of them
    assert err[well].max() <= tol, err[well].max()
    assert oerr[well].max() <= max(10 * tol, 1e-9)     # observations: differences of up to ~6 m / 2 pi scaled by 1 / (3 L), 4 / pi
    assert well.mean() >= 0.985, well.mean()
    assert np.isfinite(err).all()
    if dtype == torch.float64:
        # fp64 needs no dead-band / sign exclusion in practice: everything away from the pole agrees to 1e-8 even where
        # |e - eOld| ~ 1e-9 leaves dedt unsaturated (gain Kd 1e9 on a 1e-16 rounding)
        assert err[~near_pole].max() <= 1e-8, err[~near_pole].max()
    else:
        # fp32: the flips are rare and all sit where the diagnostics say they must
        assert (err > tol).mean() <= 6e-4, (err > tol).mean()
        assert ((err > tol) & well).sum() == 0


# ----------------------------------------------------------------- 6DoF: free-running 1000-step trajectories ------------
# asserted lower bounds on the fraction of the 4096 environments inside the tolerance at steps (100, 500, 1000);
# measured on B200 in round 2: set-point fp32 .976 / .797 / .503, force fp32 .984 / .886 / .771, fp64 >= .9946
FRACTION_FLOOR = {("setpoint", torch.float32): (0.95, 0.70, 0.40), ("force", torch.float32): (0.96, 0.82, 0.68),
                  ("setpoint", torch.float64): (0.995, 0.99, 0.98), ("force", torch.float64): (0.995, 0.99, 0.98)}


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["fp32", "fp64"])
@pytest.mark.parametrize("mode", ["force", "setpoint"])
def test_rov6_trajectory_4096_envs_1000_steps_vs_oracle(mode, dtype):
    """Tolerance 1e-8 (fp64) / 1e-4 (fp32) on the state scaled by 1 + |ref| (BASELINE north_star).  An environment
    counts as inside only while its WORST error so far is inside (checked every 10 steps)."""
    n, steps = 4096, 1000
    tol = 1e-8 if dtype == torch.float64 else 1e-4
    ref = c.Rov6EnvC(n, mode=MODES[mode], max_steps=10 ** 9)
    ref.reset()
    env = BlueROV2Heavy6DoFVecEnv(n, action_mode=mode, dtype=dtype, device=DEV, auto_reset=False, maxSteps=10 ** 9)
    env.reset()
    sync6(env, ref, mode)
    gen = torch.Generator(device="cpu").manual_seed(1236)
    scale = torch.as_tensor(SCALE6[mode])
    worst = np.zeros(n)
    fracs = {}
    for k in range(steps):
        a = ((torch.rand((n, len(SCALE6[mode])), generator=gen, dtype=torch.float64) * 2 - 1) * scale).to(dtype)
        env.step(a.to(DEV))
        ref.step(a.to(torch.float64).numpy())
        if k % 10 == 9:
            worst = np.maximum(worst, scaled_err(env.systemState.cpu().numpy(), ref.state, slice(3, 6)))
        if k + 1 in (100, 500, 1000):
            fracs[k + 1] = float((worst <= tol).mean())
    print("6DoF %s %s free-running: fraction of %d envs within %.0e at steps 100 / 500 / 1000: %.4f / %.4f / %.4f; median worst error %.2e; "
          "envs staying outside |cos theta| < 0.3: %.4f" % (mode, dtype, n, tol, fracs[100], fracs[500], fracs[1000], np.median(worst), (ref.mincos >= 0.3).mean()))
    floor = FRACTION_FLOOR[(mode, dtype)]
    assert fracs[100] >= floor[0] and fracs[500] >= floor[1] and fracs[1000] >= floor[2], (fracs, floor)
    assert bool(torch.isfinite(env.systemState).all())


def test_rov6_fixed_setpoint_fp32_tracks_oracle():
    """reset(initialSetpoint=...) - the only branch of the reference's reset that runs (6DoF.py:502-511), fixedSp = True:
    64 environments hold different fixed set-points for 250 steps; fp32 against the fp64 oracle."""
    n, steps = 64, 250
    rng = np.random.default_rng(5)
    sps = rng.uniform(-1, 1, (n, 6)) * np.array([1.0, 1.0, 1.0, 0.3, 0.3, 2.0])
    ref = c.Rov6EnvC(n, mode=o.MODE_PID, max_steps=10 ** 9)
    ref.reset(initial_setpoint=np.zeros(6))
    ref.set_point[:] = sps
    ref.path[:] = np.concatenate([sps[:, :3], sps[:, :3]], axis=1)
    env = BlueROV2Heavy6DoFVecEnv(n, action_mode="setpoint", dtype=torch.float32, device=DEV, auto_reset=False, maxSteps=10 ** 9)
    env.reset(initialSetpoint=np.zeros(6))
    sync6(env, ref, "setpoint")
    worst, frac25 = np.zeros(n), None
    for k in range(steps):
        env.step(torch.zeros((n, 6), dtype=torch.float32, device=DEV))
        ref.step(np.zeros((n, 6)))
        worst = np.maximum(worst, scaled_err(env.systemState.cpu().numpy(), ref.state, slice(3, 6)))
        if k == 24:
            frac25 = (worst <= 1e-4).mean()
    pose_gap = scaled_err(env.systemState.cpu().numpy()[:, :6], ref.state[:, :6], slice(3, 6))
    print("fixed set-point fp32: fraction of %d envs within 1e-4 after 25 / %d steps: %.3f / %.3f; median worst %.2e; final pose gap max %.2e" %
          (n, steps, frac25, (worst <= 1e-4).mean(), np.median(worst), pose_gap.max()))
    # approaching the set-point the sign of e - eOld is well defined; near rest it is rounding noise in either precision,
    # so the bang-bang stages chatter differently - there the bound is on the tracked pose, not on the chatter
    assert frac25 >= 0.5, frac25
    assert pose_gap.max() < 0.1, pose_gap.max()


# ----------------------------------------------------------------- 3DoF -------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["fp32", "fp64"])
@pytest.mark.parametrize("mode", ["rpm", "setpoint"])
def test_rov3_one_step_local_error_all_envs(mode, dtype):
    n, steps, tol = 2048, 100, TOL[dtype]
    ref = o.Rov3EnvOracle(n, mode=MODES[mode], max_steps=10 ** 9)
    ref.reset()
    env = BlueROV2Heavy3DoFVecEnv(n, action_mode=mode, dtype=dtype, device=DEV, auto_reset=False, maxSteps=10 ** 9)
    env.reset()
    rng = np.random.default_rng(3)
    na, sc = (4, 3500.0) if mode == "rpm" else (3, 1.0)
    errs, mgs, dbs = [], [], []
    try:
        for k in range(steps):
            a = torch.as_tensor(rng.uniform(-1, 1, (n, na)) * sc, dtype=dtype)
            sync3(env, ref, mode)
            ref.ctrl["margin"] = np.full(n, np.inf)
            o.DIAG["dbmargin"] = np.full(n, np.inf)
            env.step(a.to(DEV))
            ref.step(a.to(torch.float64).numpy())
            errs.append(scaled_err(env.systemState.cpu().numpy(), ref.state, slice(2, 3)))
            mgs.append(ref.ctrl["margin"].copy()); dbs.append(o.DIAG["dbmargin"].copy())
    finally:
        o.DIAG["dbmargin"] = None
    err, mg, db = map(np.concatenate, (errs, mgs, dbs))
    db_risk = (db < 1e-4) if mode != "rpm" else np.zeros(err.shape, dtype=bool)
    # |e - eOld| < 3e-8 leaves Kd dedt unsaturated (pMax / (Kd 1e9)): a 1e-16 rounding then reaches the state at 1e-10
    sign_risk = (mg < 1e-7) if mode == "setpoint" else np.zeros(err.shape, dtype=bool)
    well = ~(db_risk | sign_risk)
    report("3DoF %s %s" % (mode, dtype), err, tol, well, [("dead-band distance < 1e-4", db_risk), ("PID margin < 1e-7", sign_risk)])
    assert err[well].max() <= tol, err[well].max()
    assert well.mean() >= (0.75 if mode == "setpoint" else 0.995), well.mean()   # the 3DoF vehicle coasts a lot: small margins are common
    if dtype == torch.float64:
        assert err.max() <= 1e-8
    else:
        assert (err > tol).mean() <= 2e-4, (err > tol).mean()


@pytest.mark.parametrize("mode,dtype,steps,floors", [
    ("setpoint", torch.float32, 1000, (0.99, 0.98, 0.93)),     # measured .9985 / .9907 (step 300) / .9565
    ("rpm", torch.float32, 1000, (0.999, 0.999, 0.99)),        # measured 1.0 / 1.0 / .9985
    ("setpoint", torch.float64, 300, (0.999, 0.999, 0.999)),
], ids=["setpoint-fp32", "rpm-fp32", "setpoint-fp64"])
def test_rov3_trajectory_vs_oracle(mode, dtype, steps, floors):
    """3DoF free-running trajectories (2048 envs): fraction inside 1e-4 (fp32) / 1e-8 (fp64) at steps 100, 300 and the last."""
    n = 2048
    tol = 1e-8 if dtype == torch.float64 else 1e-4
    ref = o.Rov3EnvOracle(n, mode=MODES[mode], max_steps=10 ** 9)
    ref.reset()
    env = BlueROV2Heavy3DoFVecEnv(n, action_mode=mode, dtype=dtype, device=DEV, auto_reset=False, maxSteps=10 ** 9)
    env.reset()
    sync3(env, ref, mode)
    rng = np.random.default_rng(3)
    na, sc = (4, 3500.0) if mode == "rpm" else (3, 1.0)
    worst, fracs = np.zeros(n), {}
    for k in range(steps):
        a = torch.as_tensor(rng.uniform(-1, 1, (n, na)) * sc, dtype=dtype)
        env.step(a.to(DEV))
        ref.step(a.to(torch.float64).numpy())
        if k % 5 == 4:
            worst = np.maximum(worst, scaled_err(env.systemState.cpu().numpy(), ref.state, slice(2, 3)))
        if k + 1 in (100, 300, steps):
            fracs[k + 1] = float((worst <= tol).mean())
    print("3DoF %s %s free-running: fraction of %d envs within %.0e at steps 100 / 300 / %d: %.4f / %.4f / %.4f; median worst %.2e" %
          (mode, dtype, n, tol, steps, fracs[100], fracs[300], fracs[steps], np.median(worst)))
    assert fracs[100] >= floors[0] and fracs[300] >= floors[1] and fracs[steps] >= floors[2], (fracs, floors)


# ----------------------------------------------------------------- generic (non-default) vehicle ------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["fp32", "fp64"])
@pytest.mark.parametrize("mode", ["rpm", "setpoint"])
def test_rov6_generic_vehicle_one_step_local_error(mode, dtype):
    """The dense (non-specialised) instantiation that serves domain-randomised vehicles - off-axis CG, cross damping,
    inertia products, net buoyancy - against the C oracle with the same parameters, step by step on all environments."""
    from marinevehiclereinforcementlearning_b200 import Rov6Constants
    n, steps, tol = 1024, 40, TOL[dtype]
    p = o.Rov6Params(CG=np.array([0.01, -0.02, 0.05]), Yr=-0.3, Nv=-0.2, Kvv=-0.4, Zq=-0.1, Mw=0.2, m=11.0,
                     I=np.array([[0.16, 0.01, -0.02], [0.01, 0.17, 0.005], [-0.02, 0.005, 0.18]]))
    p.dispVol = 11.4 / 1000.
    veh = Rov6Constants()
    veh.CG, veh.Yr, veh.Nv, veh.Kvv, veh.Zq, veh.Mw, veh.m, veh.I = p.CG, p.Yr, p.Nv, p.Kvv, p.Zq, p.Mw, p.m, p.I
    ref = c.Rov6EnvC(n, params=p, mode=MODES[mode], max_steps=10 ** 9)
    ref.reset()
    env = BlueROV2Heavy6DoFVecEnv(n, action_mode=mode, dtype=dtype, device=DEV, auto_reset=False, maxSteps=10 ** 9, vehicle=veh)
    env.reset()
    assert not env._get_handle().specialised
    rng = np.random.default_rng(7)
    errs, mcs, mgs, dbs = [], [], [], []
    for k in range(steps):
        a = torch.as_tensor(rng.uniform(-1, 1, (n, len(SCALE6[mode]))) * SCALE6[mode], dtype=dtype)
        sync6(env, ref, mode)
        ref.mincos[:] = 1.0
        ref.dbmargin[:] = np.inf
        ref.ctrl_np["margin"] = np.inf
        env.step(a.to(DEV))
        ref.step(a.to(torch.float64).numpy())
        errs.append(scaled_err(env.systemState.cpu().numpy(), ref.state, slice(3, 6)))
        mcs.append(ref.mincos.copy()); mgs.append(ref.ctrl_np["margin"].copy()); dbs.append(ref.dbmargin.copy())
    err, mc, mg, db = map(np.concatenate, (errs, mcs, mgs, dbs))
    well = ~((mc < 1e-2) | ((mg < 1e-7) if mode == "setpoint" else False) | ((db < 1e-4) if mode != "rpm" else False))
    report("6DoF generic vehicle %s %s" % (mode, dtype), err, tol, well, [])
    assert err[well].max() <= tol, err[well].max()
    assert well.mean() >= 0.98
