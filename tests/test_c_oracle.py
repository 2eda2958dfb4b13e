"""CPU: the C oracle against the reference-generated golden vectors and the
numpy oracle."""
import numpy as np

from conftest import load_golden
from oracle import c_oracle as c
from oracle import oracle_np as o


def rel_err(a, ref):
    ref = np.asarray(ref)
    scale = np.abs(ref) + np.abs(ref).max(axis=-1, keepdims=True)
    return (np.abs(np.asarray(a) - ref) / np.maximum(scale, 1e-300)).max()


def test_angle_error():
    g = load_golden("resources")
    lib = c.load()
    got = np.array([lib.orc_angle_error(a, b) for a, b in g["angle_pairs"]])
    assert np.array_equal(got, g["angle_err"])
    assert np.signbit(lib.orc_angle_error(1.0, 1.0))


def test_derivs_all_modes_vs_golden():
    g = load_golden("rov6")
    assert rel_err(c.derivs6(g["rpm_states"], g["rpm_rpms"], mode=0), g["rpm_derivs"]) < 1e-13
    d, gcf, cv = c.derivs6(g["force_states"], g["force_forces"], mode=1, want_aux=True)
    assert rel_err(d, g["force_derivs"]) < 1e-13 and np.abs(cv - g["force_cv"]).max() < 1e-9
    for e in range(g["pid_sp"].shape[0]):
        ctrl = c.OrcPid6()
        for k in range(g["pid_t"].shape[1]):
            d, gcf, cv = c.derivs6(g["pid_states"][e, k], mode=2, t=g["pid_t"][e, k], sp=g["pid_sp"][e], ctrl=ctrl, want_aux=True)
            assert rel_err(d[0], g["pid_derivs"][e, k]) < 1e-12
            assert np.abs(gcf[0] - g["pid_gcf"][e, k]).max() < 1e-11
            assert np.abs(np.array(ctrl.eInt) - g["pid_eint"][e, k]).max() < 1e-14


def test_trajectory_1000_steps_vs_golden():
    t = load_golden("traj6")
    env = c.Rov6EnvC(4, mode=o.MODE_RPM, max_steps=10 ** 9, n_sub=int(t["n_sub"]), dt=float(t["dt"]))
    env.reset(initial_setpoint=np.zeros(6))
    for k in range(t["actions"].shape[0]):
        env.step(t["actions"][k])
        assert np.abs(env.state - t["traj"][k]).max() < 1e-10, k
    envf = c.Rov6EnvC(2, mode=o.MODE_FORCE, max_steps=10 ** 9, n_sub=int(t["n_sub"]), dt=float(t["dt"]))
    envf.reset(initial_setpoint=np.zeros(6))
    for k in range(t["force_actions"].shape[0]):
        envf.step(t["force_actions"][k])
        assert np.abs(envf.state - t["force_traj"][k]).max() < 1e-10, k


def test_env_pid_vs_reference_env_golden():
    e6 = load_golden("env6")
    env = c.Rov6EnvC(1, mode=o.MODE_PID, max_steps=60)
    obs = [env.reset(initial_setpoint=e6["fixed_sp"])[0]]
    for k in range(60):
        ob, r, d, _ = env.step(np.zeros((1, 6)))
        obs.append(ob[0].copy())
        assert np.abs(env.state[0] - e6["fixed_history"][k + 1, 1:13]).max() < 1e-9
        assert np.abs(env.aux[0, :6] - e6["fixed_history"][k + 1, 13:19]).max() < 1e-7
        assert bool(d[0]) == bool(e6["fixed_done"][k])
    assert np.abs(np.array(obs) - e6["fixed_obs"]).max() < 1e-10


def test_auto_reset_matches_numpy_oracle():
    n = 64
    rng = np.random.default_rng(4)
    a = c.Rov6EnvC(n, mode=o.MODE_RPM, max_steps=4, auto_reset=True, seed=21, env_id0=1000)
    b = o.Rov6EnvOracle(n, mode=o.MODE_RPM, max_steps=4, auto_reset=True, seed=21, env_id0=1000)
    a.reset(); b.reset()
    for k in range(9):
        act = rng.uniform(-3500, 3500, (n, 8))
        oa, _, da, ia = a.step(act)
        ob, _, db, ib = b.step(act)
        assert np.array_equal(da, db) and np.abs(oa - ob).max() < 1e-12 and np.abs(a.state - b.state).max() < 1e-12
        if db.any():
            assert np.abs(a.path - b.path).max() == 0.0
            assert np.abs(ia["terminal_observation"][db] - ib["terminal_observation"][db]).max() < 1e-12
