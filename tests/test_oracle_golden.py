"""CPU: the numpy oracle against the golden vectors produced by running the
unmodified reference (tests/golden/gen_golden_current.py)."""
import numpy as np
import pytest

from oracle import oracle_np as o
from conftest import load_golden


def rel_err(a, ref):
    ref = np.asarray(ref)
    scale = np.abs(ref) + np.abs(ref).max(axis=-1, keepdims=True)
    return (np.abs(np.asarray(a) - ref) / np.maximum(scale, 1e-300)).max()


def test_angle_error_kat():
    g = load_golden("resources")
    got = o.angle_error(g["angle_pairs"][:, 0], g["angle_pairs"][:, 1])
    assert np.array_equal(got, g["angle_err"])
    # KAT-4 of SURVEY.md 8(c): sign of zero and the a == b == pi case
    assert np.signbit(o.angle_error(1.0, 1.0)) and o.angle_error(1.0, 1.0) == 0.0
    assert o.angle_error(0.0, np.pi) == -np.pi
    assert o.angle_error(0.1, 6.2) == pytest.approx(0.1831853071795857, abs=1e-15)


def test_coordinate_transform_and_allocation():
    g = load_golden("resources")
    a = g["ct_angles"]
    assert rel_err(o.coordinate_transform6(a[:, 0], a[:, 1], a[:, 2]).reshape(len(a), -1), g["ct_J6"].reshape(len(a), -1)) < 1e-15
    assert np.array_equal(o.coordinate_transform3(a[:, 2]), g["ct_J3"])
    assert np.array_equal(g["ct_J3"], g["ct_J3_default"])
    A, Ainv = o.compute_thrust_allocation(g["thrusterPositions"], g["thrusterNormals"])
    assert np.array_equal(A, g["A6"]) and np.abs(Ainv - g["Ainv6"]).max() < 1e-15
    A2, Ainv2 = o.compute_thrust_allocation(g["thrusterPositions"], g["thrusterNormals"], x0=g["alloc_x0"])
    assert np.array_equal(A2, g["A6_x0"]) and np.abs(Ainv2 - g["Ainv6_x0"]).max() < 1e-15


def test_survey_kats():
    """The frozen numbers of SURVEY.md 8(c) (KAT-A, KAT-M, KAT-1, KAT-2)."""
    g = load_golden("rov6")
    p = o.Rov6Params()
    assert float(g["rhoD4Kt"]) == pytest.approx(0.011755102040816326, rel=1e-15)
    assert p.A[:, 0] == pytest.approx([0.838671, -0.544639, 0, 0.037035, 0.05703, -0.16504], abs=1e-6)
    assert p.Ainv[4] == pytest.approx([-0.141667, -0.077273, -0.25, -1.136364, 2.083333, 0], abs=1e-6)
    Minv = np.linalg.inv(p.mass_matrix())
    assert Minv[0, 0] == pytest.approx(0.06353384311678882, rel=1e-13)
    assert Minv[0, 4] == pytest.approx(-0.12933675205917725, rel=1e-13)
    assert Minv[3, 3] == pytest.approx(3.7520823278479236, rel=1e-13)
    d = o.derivs6_rpm(p, g["rpm_states"][0], g["rpm_rpms"][0])[0]
    assert d[6:] == pytest.approx([-2.3458113182380824, 0.14117300847952502, -0.22319560687432868,
                                   -16.501583727295348, -10.265217360246886, -32.028225894606095], rel=1e-12)
    assert d[:6] == pytest.approx([0.25264520959187564, 0.19041218341361893, 0.05894086018873085,
                                   0.01076453679380168, -0.04926893041474337, 0.09250873621674992], rel=1e-12)
    assert g["pid_gcf"][0, 0] == pytest.approx([-2.52, 5.04, -7.56, -1, 1, -1.02], abs=1e-12)
    assert g["pid_derivs"][0, 0, 6:] == pytest.approx([-0.37836252889058264, 0.12903386553378404, -0.7717309882328138,
                                                     -3.214521163590409, 8.018270531747259, -3.309355392109698], rel=1e-12)


def test_example_temp_golden():
    """The only numeric vector inside the reference repo (example_temp.py:19-28):
    acc = solve(M, RHS) to the 7 digits it was printed with; its M pins
    M[2,2] = m (the Zvdot slip) and the +-m*zg couplings (older zg = 0.025)."""
    RHS = np.array([-1.366025e+01, 3.660254e+00, 5.0, 0.0, 0.0, 0.0])
    acc = np.array([-8.224159e-01, 1.537282e-01, 4.385965e-01, 1.564733e-01, 8.371019e-01, 0.0])
    p = o.Rov6Params(CG=np.array([0., 0., 0.025]))
    M = p.mass_matrix()
    assert M[0, 0] == pytest.approx(16.9) and M[1, 1] == pytest.approx(24.1) and M[2, 2] == pytest.approx(11.4)
    assert M[0, 4] == pytest.approx(0.285) and M[1, 3] == pytest.approx(-0.285) and M[3, 3] == pytest.approx(0.28)
    assert np.abs(np.linalg.solve(M, RHS) - acc).max() < 5e-7


def test_derivs_rpm_and_components():
    g = load_golden("rov6")
    p = o.Rov6Params()
    s = g["rpm_states"]
    assert rel_err(o.derivs6_rpm(p, s, g["rpm_rpms"]), g["rpm_derivs"]) < 1e-14
    comps = o.force_components6(p, s[:, 3:6], s[:, 6:12], g["rpm_rpms"])
    for k in range(5):
        assert np.abs(comps[k] - g["rpm_retComp"][:, :, k]).max() < 1e-12
    ih, jh, kh = o.body_axes(s[:, 3:6])
    assert np.abs(np.stack([ih, jh, kh], axis=1) - g["rpm_axes"]).max() < 1e-15
    assert np.abs(o.thruster_force6(p, g["thruster_rpm"]) - g["thruster_F"]).max() < 1e-13
    assert np.abs(np.linalg.inv(p.mass_matrix()) - g["Minv"]).max() == 0.0


def test_derivs_force_mode():
    g = load_golden("rov6")
    p = o.Rov6Params()
    d, cv = o.derivs6_force(p, g["force_states"], g["force_forces"], return_cv=True)
    assert rel_err(d, g["force_derivs"]) < 1e-14
    assert np.abs(cv - g["force_cv"]).max() < 1e-10


def test_pid_sequences():
    g = load_golden("rov6")
    p = o.Rov6Params()
    for e in range(g["pid_sp"].shape[0]):
        ctrl = o.pid6_new_state(1)
        for c in range(g["pid_t"].shape[1]):
            d, gcf, cv = o.derivs6_pid(p, g["pid_t"][e, c], g["pid_states"][e, c], ctrl, g["pid_sp"][e:e + 1], True)
            assert rel_err(d[0], g["pid_derivs"][e, c]) < 1e-13
            assert np.abs(gcf[0] - g["pid_gcf"][e, c]).max() < 1e-12
            assert np.abs(cv[0] - g["pid_cv"][e, c]).max() < 1e-9
            assert np.abs(ctrl["eInt"][0] - g["pid_eint"][e, c]).max() < 1e-14


def test_trajectories_1000_steps():
    t = load_golden("traj6")
    env = o.Rov6EnvOracle(4, mode=o.MODE_RPM, max_steps=10 ** 9, n_sub=int(t["n_sub"]), dt=float(t["dt"]))
    env.reset(initial_setpoint=np.zeros(6))
    for k in range(0, 300):  # the full 1000 steps run in the gpu-marked test against the same file
        env.step(t["actions"][k])
        assert np.abs(env.state - t["traj"][k]).max() < 1e-11
    envf = o.Rov6EnvOracle(2, mode=o.MODE_FORCE, max_steps=10 ** 9, n_sub=int(t["n_sub"]), dt=float(t["dt"]))
    envf.reset(initial_setpoint=np.zeros(6))
    for k in range(t["force_actions"].shape[0]):
        envf.step(t["force_actions"][k])
        assert np.abs(envf.state - t["force_traj"][k]).max() < 1e-11


def test_env_semantics_fixed_and_action_driven():
    e6 = load_golden("env6")
    env = o.Rov6EnvOracle(1, mode=o.MODE_PID, max_steps=60)
    obs = [env.reset(initial_setpoint=e6["fixed_sp"])[0]]
    hist = [env.history_row()[0]]
    dones = []
    for k in range(60):
        ob, r, d, _ = env.step(np.zeros((1, 6)))
        obs.append(ob[0]); hist.append(env.history_row()[0]); dones.append(d[0])
        assert r[0] == 0.0
    assert np.abs(np.array(obs) - e6["fixed_obs"]).max() < 1e-12
    assert np.abs(np.array(hist) - e6["fixed_history"]).max() < 1e-8
    assert np.array_equal(np.array(dones), e6["fixed_done"])
    assert list(e6["fixed_history_cols"][:4]) == ["t", "x", "y", "z"] and e6["fixed_history"].shape[1] == 33

    env = o.Rov6EnvOracle(1, mode=o.MODE_PID, max_steps=40)
    env.reset(initial_setpoint=np.append(e6["act_path"][0], e6["act_orient"]))
    env.path[0] = e6["act_path"].reshape(-1)
    env.fixed_sp = False
    obs = [env.observe()[0]]
    hist = [env.history_row()[0]]
    for k in range(40):
        ob, r, d, _ = env.step(e6["act_actions"][k:k + 1])
        obs.append(ob[0]); hist.append(env.history_row()[0])
    assert np.abs(np.array(obs) - e6["act_obs"]).max() < 1e-12
    assert np.abs(np.array(hist) - e6["act_history"]).max() < 1e-8


def test_philox_reference_vector():
    """Philox4x32-10 known-answer test (Random123 kat_vectors: zero counter/key
    and the all-ones vector)."""
    z = o.philox4x32(np.zeros((1, 4), dtype=np.uint32), (0, 0))[0]
    assert [int(v) for v in z] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = o.philox4x32(np.full((1, 4), 0xffffffff, dtype=np.uint32), (0xffffffff, 0xffffffff))[0]
    assert [int(v) for v in f] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    u = o.philox_uniform(7, np.arange(1000), np.zeros(1000), 9)
    assert u.min() >= 0.0 and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.02


def test_auto_reset_oracle():
    env = o.Rov6EnvOracle(8, mode=o.MODE_RPM, max_steps=3, auto_reset=True, seed=5)
    env.reset()
    p0 = env.path.copy()
    rng = np.random.default_rng(0)
    for k in range(3):
        obs, r, d, info = env.step(rng.uniform(-3500, 3500, (8, 8)))
    assert d.all() and (env.i_step == 0).all() and (env.state == 0).all()
    assert "terminal_observation" in info and not np.allclose(env.path, p0)
    assert np.all(np.abs(env.path) <= 5.0) and np.all((env.set_point[:, 3:] >= 0) & (env.set_point[:, 3:] < 2 * np.pi))
