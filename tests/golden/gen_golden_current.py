"""Generate golden vectors by EXECUTING the unmodified reference (current
3DoF/6DoF generation) in this container.

    python tests/golden/gen_golden_current.py

Needs /root/reference (read-only) + numpy/scipy/pandas; writes small ``.npz``
fixtures next to this file.  The fixtures are committed; this script is the
provenance record.  Nothing here is imported by the product or by the tests.

The reference ships no tests, so these executed outputs
are what pins the oracle.  Reference entry points exercised (file:line):

* resources.computeThrustAllocation   resources.py:19-35
* resources.angleError                resources.py:75-95
* resources.coordinateTransform       resources.py:98-143
* BlueROV2Heavy6DoF.forceModel/derivs dynamicsModel_BlueROV2_Heavy_6DoF.py:253-442
* BlueROV2Heavy6DoF_PID_controller    dynamicsModel_BlueROV2_Heavy_6DoF.py:27-73
* BlueROV2Heavy6DoFEnv.reset/step     dynamicsModel_BlueROV2_Heavy_6DoF.py:485-594
* BlueROV2Heavy3DoF.derivs            dynamicsModel_BlueROV2_Heavy_3DoF.py:128-296
* BlueROV2Heavy3DoFEnv.reset/step     dynamicsModel_BlueROV2_Heavy_3DoF.py:411-514

The env ``step`` functions integrate with scipy's adaptive RK45; the north
star defines parity against the reference's derivative function driven by a
FIXED-STEP classic RK4, so ``scipy.integrate.solve_ivp`` is swapped for the
RK4 driver below while the reference env code runs (everything else in
``step`` - set-point scaling, wrap, observation, done, timeHistory - is the
reference's own code).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_shims import import_current  # noqa: E402

ref_res, ref6, ref3 = import_current()
import scipy.integrate  # noqa: E402


# --------------------------------------------------------------------------
# fixed-step RK4 around a reference derivative function
# --------------------------------------------------------------------------
def rk4_substep(f, t, y, h):
    k1 = f(t, y)
    k2 = f(t + 0.5 * h, y + 0.5 * h * k1)
    k3 = f(t + 0.5 * h, y + 0.5 * h * k2)
    k4 = f(t + h, y + h * k3)
    return y + (h / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)


def rk4_advance(f, t0, y, dt, n_sub):
    h = dt / n_sub
    for j in range(n_sub):
        y = rk4_substep(f, t0 + j * h, y, h)
    return y


class _IvpResult:
    def __init__(self, t, y):
        self.t = np.array([t])
        self.y = np.array(y, dtype=float).reshape(-1, 1)


def make_fixed_step_ivp(n_sub):
    def solve_ivp(fun, t_span, y0, method=None, t_eval=None, max_step=None, rtol=None, atol=None, **kw):
        y = rk4_advance(fun, t_span[0], np.array(y0, dtype=float), t_span[1] - t_span[0], n_sub)
        return _IvpResult(t_span[1], y)
    return solve_ivp


class ConstController:
    """Stateless controller injected through the reference's own seam
    (dynamicsModel_BlueROV2_Heavy_6DoF.py:76-78, 418)."""

    def __init__(self, f):
        self.f = np.array(f, dtype=float)
        self.setPoint = np.zeros(6)

    def reset(self):
        pass

    def computeControlForces(self, x, y, z, phi, theta, psi, t):
        return self.f.copy()


def sample_states6(rng, n):
    s = np.empty((n, 12))
    s[:, 0:3] = rng.uniform(-5.0, 5.0, (n, 3))
    s[:, 3:6] = rng.uniform(-np.pi, np.pi, (n, 3))
    s[:, 6:9] = rng.uniform(-2.0, 2.0, (n, 3))
    s[:, 9:12] = rng.uniform(-5.0, 5.0, (n, 3))
    return s


def ref_f6_rpm(rov, state, rpm):
    """dynamicsModel_BlueROV2_Heavy_6DoF.py:424-442 minus the controller."""
    pos, angles, vel = state[0:3], state[3:6], state[6:12]
    rov.updateMovingCoordSystem(angles)
    M, RHS = rov.forceModel(pos, angles, vel, rpm)
    acc = np.linalg.solve(M, RHS)
    J = ref_res.coordinateTransform(angles[0], angles[1], angles[2], dof=6)
    return np.append(np.dot(J, vel), acc), M, RHS


# --------------------------------------------------------------------------
def gen_resources(out):
    rng = np.random.default_rng(101)
    pairs = np.array([(0.1, 6.2), (6.2, 0.1), (3.0, 0.0), (0.0, np.pi), (1.0, 1.0), (-0.5, 7.0),
                      (np.pi, np.pi), (0.0, 0.0), (2 * np.pi, 0.0), (0.0, 2 * np.pi), (-7.0, 9.0)]
                     + [tuple(v) for v in rng.uniform(-10, 10, (117, 2))])
    out["angle_pairs"] = pairs
    out["angle_err"] = np.array([ref_res.angleError(a, b) for a, b in pairs])

    ang = rng.uniform(-np.pi, np.pi, (61, 3))
    # edge cases of the cos(theta) clamp, resources.py:116-120
    edge = np.array([
        [0.3, np.pi / 2, 0.7], [0.3, -np.pi / 2, 0.7], [0.3, np.pi / 2 - 1e-7, 0.7],
        [0.3, np.pi / 2 + 1e-7, 0.7], [0.3, np.pi / 2 - 1e-13, 0.7], [0.0, 0.0, 0.0],
        [-0.2, 3 * np.pi / 2 + 5e-7, 1.0]])
    ang = np.vstack([ang, edge])
    out["ct_angles"] = ang
    out["ct_J6"] = np.array([ref_res.coordinateTransform(a[0], a[1], a[2], dof=6) for a in ang])
    out["ct_J3"] = np.array([ref_res.coordinateTransform(a[0], a[1], a[2], dof=3) for a in ang])
    out["ct_J3_default"] = np.array([ref_res.coordinateTransform(a[0], a[1], a[2]) for a in ang])

    rov = ref6.BlueROV2Heavy6DoF(ConstController(np.zeros(6)))
    out["thrusterPositions"] = rov.thrusterPositions
    out["thrusterNormals"] = rov.thrusterNormals
    out["A6"] = rov.A
    out["Ainv6"] = rov.Ainv
    x0 = np.array([0.01, -0.02, 0.0375])
    A2, Ainv2 = ref_res.computeThrustAllocation(rov.thrusterPositions, rov.thrusterNormals, x0=x0)
    out["alloc_x0"] = x0
    out["A6_x0"] = A2
    out["Ainv6_x0"] = Ainv2


def gen_rov6(out):
    rng = np.random.default_rng(202)
    rov = ref6.BlueROV2Heavy6DoF(ConstController(np.zeros(6)))
    out["Kt_thruster"] = np.array(rov.Kt_thruster)
    out["rhoD4Kt"] = np.array(rov.rho_f * rov.D_thruster ** 4. * rov.Kt_thruster)

    # --- KAT-1 of SURVEY.md + random direct-rpm derivative evaluations
    n = 256
    states = sample_states6(rng, n)
    states[0] = [0.1, -0.2, 0.3, 0.2, -0.1, 1.0, 0.3, -0.1, 0.05, 0.02, -0.03, 0.1]
    rpms = rng.uniform(-4200.0, 4200.0, (n, 8))
    rpms[0] = [1000, -2000, 500, 3600, -250, 1500, -1500, 800]
    rpms[1] = [300, -300, 299.999, -299.999, 3500, -3500, 3500.001, 0.0]
    d = np.empty((n, 12)); rhs = np.empty((n, 6)); axes = np.empty((n, 3, 3)); comp = np.empty((n, 6, 5))
    for i in range(n):
        d[i], M, rhs[i] = ref_f6_rpm(rov, states[i], rpms[i])
        axes[i] = np.array([rov.iHat, rov.jHat, rov.kHat])
        comp[i] = rov.forceModel(states[i, 0:3], states[i, 3:6], states[i, 6:12], rpms[i], retComp=True)
    out["M"] = M
    out["Minv"] = np.linalg.inv(M)
    out["rpm_states"] = states
    out["rpm_rpms"] = rpms
    out["rpm_derivs"] = d
    out["rpm_RHS"] = rhs
    out["rpm_axes"] = axes
    out["rpm_retComp"] = comp
    out["thruster_rpm"] = np.linspace(-3500, 3500, 15)
    out["thruster_F"] = rov.thrusterModel(out["thruster_rpm"])

    # --- generalised-force mode through the injected stateless controller
    n = 128
    states = sample_states6(rng, n)
    forces = rng.uniform(-1.0, 1.0, (n, 6)) * np.array([50., 50., 50., 2., 2., 2.])
    forces[0] = 0.0
    forces[1] = [1e-9, -1e-9, 0., 0., 1e-12, 0.]
    d = np.empty((n, 12)); cv = np.empty((n, 8))
    for i in range(n):
        r = ref6.BlueROV2Heavy6DoF(ConstController(forces[i]))
        d[i] = r.derivs(0.3, states[i])
        cv[i] = r.controlVector
    out["force_states"] = states
    out["force_forces"] = forces
    out["force_derivs"] = d
    out["force_cv"] = cv

    # --- stateful PID controller: call sequences in RK4 stage-time order
    n_env, n_call = 6, 48
    h = 0.025
    stage_dt = [0.0, 0.5 * h, 0.5 * h, h]
    sp = rng.uniform(-1, 1, (n_env, 6)) * np.array([1., 1., 1., 0.5, 0.5, 3.0])
    sp[0] = 0.0
    seq_states = np.empty((n_env, n_call, 12)); seq_t = np.empty((n_env, n_call))
    seq_d = np.empty((n_env, n_call, 12)); seq_gcf = np.empty((n_env, n_call, 6)); seq_cv = np.empty((n_env, n_call, 8))
    seq_eint = np.empty((n_env, n_call, 6))
    for e in range(n_env):
        r = ref6.BlueROV2Heavy6DoF(ref6.BlueROV2Heavy6DoF_PID_controller(sp[e].copy()))
        base = sample_states6(rng, 1)[0] * 0.2
        if e == 0:
            base = np.array([0.1, -0.2, 0.3, 0.2, -0.1, 1.0, 0.3, -0.1, 0.05, 0.02, -0.03, 0.1])
        for c in range(n_call):
            t = (0.1 if e == 0 else 0.0) + (c // 4) * h + stage_dt[c % 4]
            s = base + (0.0 if c == 0 else 1.0) * rng.normal(0, 1e-3, 12) + 0.002 * c
            seq_states[e, c] = s; seq_t[e, c] = t
            seq_d[e, c] = r.derivs(t, s)
            seq_gcf[e, c] = r.generalisedControlForces
            seq_cv[e, c] = r.controlVector
            seq_eint[e, c] = r.controller.eInt
    out["pid_sp"] = sp
    out["pid_states"] = seq_states
    out["pid_t"] = seq_t
    out["pid_derivs"] = seq_d
    out["pid_gcf"] = seq_gcf
    out["pid_cv"] = seq_cv
    out["pid_eint"] = seq_eint


def gen_traj6(out, n_env=4, n_steps=1000, n_sub=8, dt=0.2):
    """1000-step trajectories, direct-rpm mode, reference forceModel under RK4
    (config 2 of BASELINE.json at a size the reference finishes in a minute)."""
    rng = np.random.default_rng(1234)
    rov = ref6.BlueROV2Heavy6DoF(ConstController(np.zeros(6)))
    actions = rng.uniform(-3500.0, 3500.0, (n_steps, n_env, 8))
    traj = np.empty((n_steps, n_env, 12))
    t_start = time.time()
    for e in range(n_env):
        y = np.zeros(12)
        for k in range(n_steps):
            rpm = actions[k, e]
            f = lambda t, s: ref_f6_rpm(rov, s, rpm)[0]
            y = rk4_advance(f, k * dt, y, dt, n_sub)
            y[3:6] = y[3:6] % (2. * np.pi)  # dynamicsModel_BlueROV2_Heavy_6DoF.py:560
            traj[k, e] = y
    out["n_sub"] = np.array(n_sub); out["dt"] = np.array(dt)
    out["actions"] = actions
    out["traj"] = traj
    print("traj6 rpm: %.1f s" % (time.time() - t_start))

    # generalised-force mode (earth-frame force held over each step)
    n_env_f, n_steps_f = 2, 250
    forces = rng.uniform(-1.0, 1.0, (n_steps_f, n_env_f, 6)) * np.array([50., 50., 50., 1., 1., 2.])
    trajf = np.empty((n_steps_f, n_env_f, 12))
    for e in range(n_env_f):
        y = np.zeros(12)
        for k in range(n_steps_f):
            r = ref6.BlueROV2Heavy6DoF(ConstController(forces[k, e]))
            y = rk4_advance(r.derivs, k * dt, y, dt, n_sub)
            y[3:6] = y[3:6] % (2. * np.pi)
            trajf[k, e] = y
    out["force_actions"] = forces
    out["force_traj"] = trajf


def gen_env6(out, n_sub=8):
    """Reference env semantics with the integrator swapped for fixed-step RK4."""
    orig = scipy.integrate.solve_ivp
    scipy.integrate.solve_ivp = make_fixed_step_ivp(n_sub)
    try:
        # (a) fixed set-point (the reference's own __main__ scenario, 6DoF.py:757-761)
        sp = [0.5, -0.3, 0.2, 10. / 180. * np.pi, -5. / 180. * np.pi, 280. / 180. * np.pi]
        env = ref6.BlueROV2Heavy6DoFEnv(maxSteps=60)
        obs0 = env.reset(initialSetpoint=sp)
        obs = [obs0]; dones = []; rewards = []
        for k in range(60):
            o, r, d, info = env.step(np.zeros(6))
            obs.append(o); dones.append(d); rewards.append(r)
        out["fixed_sp"] = np.array(sp)
        out["fixed_obs"] = np.array(obs)
        out["fixed_done"] = np.array(dones)
        out["fixed_reward"] = np.array(rewards)
        out["fixed_history"] = env.timeHistory.values  # 33 columns, 6DoF.py:578-587
        out["fixed_history_cols"] = np.array(list(env.timeHistory.columns))

        # (b) action-driven set-points.  reset() without initialSetpoint raises in
        # the reference (6DoF.py:497, shape (2,3)-(2,)), so the random branch's
        # state is installed by hand: path (2,3), targetOrientation, fixedSp=False.
        rng = np.random.default_rng(77)
        env = ref6.BlueROV2Heavy6DoFEnv(maxSteps=40)
        path = (rng.random((2, 3)) - 0.5) * 10.
        orient = rng.random(3) * 2. * np.pi
        env.reset(initialSetpoint=np.append(path[0], orient))
        env.path = path.copy(); env.targetOrientation = orient.copy(); env.fixedSp = False
        env.state = env.dataToState(env.systemState)
        actions = rng.uniform(-1, 1, (40, 6))
        obs = [env.state]; dones = []
        for k in range(40):
            o, r, d, info = env.step(actions[k])
            obs.append(o); dones.append(d)
        out["act_path"] = path
        out["act_orient"] = orient
        out["act_actions"] = actions
        out["act_obs"] = np.array(obs)
        out["act_done"] = np.array(dones)
        out["act_history"] = env.timeHistory.values
    finally:
        scipy.integrate.solve_ivp = orig


def gen_rov3(out, n_sub=8):
    rng = np.random.default_rng(303)
    r = ref3.BlueROV2Heavy3DoF(np.array([1., -1., 280. / 180. * np.pi]))
    s = np.array([0.1, -0.2, 1.0, 0.3, -0.1, 0.1])
    out["kat3_state"] = s
    out["kat3_derivs"] = r.derivs(0.1, s)
    out["kat3_cv"] = r.controlVector
    out["Ainv3"] = r.Ainv

    # thruster model incl. jet-drag augment, 3DoF.py:114-126
    uvr = rng.uniform(-2, 2, (64, 2)); rpm = rng.uniform(-3500, 3500, 64)
    rpm[:3] = [0.0, 1e-3, -3500.0]; uvr[3] = 0.0
    FX = np.array([r.thrusterModel(uvr[i, 0], uvr[i, 1], rpm[i]) for i in range(64)])
    out["thr3_uv"] = uvr; out["thr3_rpm"] = rpm; out["thr3_FX"] = FX

    # PID call sequences in RK4 stage order
    n_env, n_call = 6, 48
    h = 0.025
    stage_dt = [0.0, 0.5 * h, 0.5 * h, h]
    sp = rng.uniform(-1, 1, (n_env, 3)) * np.array([1., 1., 3.])
    seq_s = np.empty((n_env, n_call, 6)); seq_t = np.empty((n_env, n_call)); seq_d = np.empty((n_env, n_call, 6))
    seq_gcf = np.empty((n_env, n_call, 3)); seq_cv = np.empty((n_env, n_call, 4))
    for e in range(n_env):
        r = ref3.BlueROV2Heavy3DoF(sp[e].copy())
        base = rng.uniform(-1, 1, 6) * np.array([0.5, 0.5, 3., 1., 1., 1.])
        for c in range(n_call):
            t = (c // 4) * h + stage_dt[c % 4]
            s = base + (0.0 if c == 0 else 1.0) * rng.normal(0, 1e-3, 6) + 0.002 * c
            seq_s[e, c] = s; seq_t[e, c] = t
            seq_d[e, c] = r.derivs(t, s)
            seq_gcf[e, c] = r.generalisedControlForces
            seq_cv[e, c] = r.controlVector
    out["pid3_sp"] = sp; out["pid3_states"] = seq_s; out["pid3_t"] = seq_t
    out["pid3_derivs"] = seq_d; out["pid3_gcf"] = seq_gcf; out["pid3_cv"] = seq_cv

    # env semantics with fixed-step RK4
    orig = scipy.integrate.solve_ivp
    scipy.integrate.solve_ivp = make_fixed_step_ivp(n_sub)
    try:
        sp = [0.5, -0.3, 280. / 180. * np.pi]
        env = ref3.BlueROV2Heavy3DoFEnv(maxSteps=50)
        obs = [env.reset(initialSetpoint=sp)]; dones = []
        for k in range(50):
            o, rwd, d, info = env.step(np.zeros(3))
            obs.append(o); dones.append(d)
        out["env3_fixed_sp"] = np.array(sp)
        out["env3_fixed_obs"] = np.array(obs)
        out["env3_fixed_done"] = np.array(dones)
        out["env3_fixed_history"] = env.timeHistory.values  # 17 columns, 3DoF.py:498-507

        np.random.seed(5)  # the reference draws from the global RNG, 3DoF.py:423-424
        env = ref3.BlueROV2Heavy3DoFEnv(maxSteps=40)
        obs = [env.reset()]
        out["env3_act_path"] = env.path.copy(); out["env3_act_heading"] = np.array(env.targetHeading)
        actions = rng.uniform(-1, 1, (40, 3)); dones = []
        for k in range(40):
            o, rwd, d, info = env.step(actions[k])
            obs.append(o); dones.append(d)
        out["env3_act_actions"] = actions
        out["env3_act_obs"] = np.array(obs)
        out["env3_act_done"] = np.array(dones)
        out["env3_act_history"] = env.timeHistory.values
    finally:
        scipy.integrate.solve_ivp = orig


def main():
    for name, fn in (("resources", gen_resources), ("rov6", gen_rov6), ("traj6", gen_traj6),
                     ("env6", gen_env6), ("rov3", gen_rov3)):
        out = {}
        t0 = time.time()
        fn(out)
        path = os.path.join(HERE, "golden_%s.npz" % name)
        np.savez_compressed(path, **out)
        print("%-10s %6.1f s  %7.1f KiB  %s" % (name, time.time() - t0, os.path.getsize(path) / 1024., path))


if __name__ == "__main__":
    main()
