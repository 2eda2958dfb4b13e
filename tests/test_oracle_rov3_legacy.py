"""Pins the 3DoF and legacy (AuvEnv / ReconstructedFlow) numpy oracles to the
golden vectors produced by executing the unmodified reference
(tests/golden/gen_golden_current.py, gen_golden_legacy.py).  CPU only."""
import numpy as np

from conftest import load_golden
from oracle import oracle_np as o


def rel_err(a, ref):
    ref = np.asarray(ref)
    scale = np.abs(ref) + np.abs(ref).max(axis=-1, keepdims=True)
    return (np.abs(np.asarray(a) - ref) / np.maximum(scale, 1e-300)).max()


# ------------------------------------------------------------------ 3DoF ----
def test_rov3_kat_and_thruster_model():
    g = load_golden("rov3")
    p = o.Rov3Params()
    assert np.abs(p.Ainv - g["Ainv3"]).max() < 1e-15
    ctrl = o.pid3_new_state(1)
    sp = np.array([[1., -1., 280. / 180. * np.pi]])
    d, gcf, cv = o.derivs3_pid(p, 0.1, g["kat3_state"], ctrl, sp, return_aux=True)
    assert rel_err(d[0], g["kat3_derivs"]) < 1e-13
    assert np.abs(cv[0] - g["kat3_cv"]).max() < 1e-9
    # SURVEY.md KAT-3
    assert abs(d[0, 5] - (-108.71474492988926)) < 1e-9
    F, X = o.thruster_model3(p, g["thr3_uv"][:, 0], g["thr3_rpm"])
    assert rel_err(np.stack([F, X], axis=1), g["thr3_FX"]) < 1e-13


def test_rov3_pid_sequences():
    g = load_golden("rov3")
    p = o.Rov3Params()
    for e in range(g["pid3_sp"].shape[0]):
        ctrl = o.pid3_new_state(1)
        for c in range(g["pid3_t"].shape[1]):
            d, gcf, cv = o.derivs3_pid(p, g["pid3_t"][e, c], g["pid3_states"][e, c], ctrl, g["pid3_sp"][e:e + 1], True)
            assert rel_err(d[0], g["pid3_derivs"][e, c]) < 1e-12, (e, c)
            assert np.abs(gcf[0] - g["pid3_gcf"][e, c]).max() < 1e-10
            assert np.abs(cv[0] - g["pid3_cv"][e, c]).max() < 1e-7


def test_rov3_env_semantics():
    g = load_golden("rov3")
    env = o.Rov3EnvOracle(1, mode=o.MODE_PID, max_steps=50)
    obs = [env.reset(initial_setpoint=g["env3_fixed_sp"])[0]]
    hist = [env.history_row()[0]]
    dones = []
    for k in range(50):
        ob, r, d, _ = env.step(np.zeros((1, 3)))
        obs.append(ob[0]); hist.append(env.history_row()[0]); dones.append(d[0])
        assert r[0] == 0.0
    assert np.abs(np.array(obs) - g["env3_fixed_obs"]).max() < 1e-11
    assert np.abs(np.array(hist) - g["env3_fixed_history"]).max() < 1e-7
    assert np.array_equal(np.array(dones), g["env3_fixed_done"]) and g["env3_fixed_history"].shape[1] == 17

    env = o.Rov3EnvOracle(1, mode=o.MODE_PID, max_steps=40)
    env.reset(initial_setpoint=np.append(g["env3_act_path"][0], g["env3_act_heading"]))
    env.path[0] = g["env3_act_path"].reshape(-1)
    env.fixed_sp = False
    obs = [env.observe()[0]]
    hist = [env.history_row()[0]]
    for k in range(40):
        ob, r, d, _ = env.step(g["env3_act_actions"][k:k + 1])
        obs.append(ob[0]); hist.append(env.history_row()[0])
    assert np.abs(np.array(obs) - g["env3_act_obs"]).max() < 1e-11
    assert np.abs(np.array(hist) - g["env3_act_history"]).max() < 1e-7


def test_rov3_auto_reset_oracle():
    env = o.Rov3EnvOracle(64, mode=o.MODE_RPM, max_steps=3, auto_reset=True, seed=5)
    env.reset()
    rng = np.random.default_rng(0)
    for k in range(7):
        ob, r, d, info = env.step(rng.uniform(-3500, 3500, (64, 4)))
        assert d.all() == ((k + 1) % 3 == 0)
        if d.all():
            assert (env.state == 0).all() and "terminal_observation" in info
    assert np.abs(env.path).max() <= 5.0 and (env.episode == 2).all()


# ---------------------------------------------------------------- legacy ----
def make_flow(g):
    base = g["ltm"][None] + 0.05 * np.random.default_rng(7).standard_normal((int(g["nt"]),) + g["ltm"].shape)
    assert np.array_equal(base[::7, ::5, ::6, :], g["base_field_sample"])  # the regenerated synthetic field is the generator's
    flow = o.FlowOracle(base, float(g["base_dx"]), float(g["base_dy"]), float(g["base_dt"]))
    flow.scale(11., 1.0, 2.0, translate=(-1.65, -1.1))
    return flow


def test_flow_scale_and_interp():
    g = load_golden("legacy")
    flow = make_flow(g)
    assert flow.dx == float(g["scaled_dx"]) and flow.dy == float(g["scaled_dy"]) and flow.dt == float(g["scaled_dt"])
    assert np.abs(flow.flowData[::7, ::5, ::6, :] - g["scaled_field_sample"]).max() < 1e-15
    res = flow.interp(g["interp_t"], g["interp_xy"])
    assert rel_err(res, g["interp_res"]) < 1e-12
    assert flow.time[flow.time.shape[0] // 4] == float(g["flow_time_quarter"])


def test_spod_reconstruction_oracle_vs_reference_constructor():
    """oracle.FlowOracle.reconstruct / intensity against the golden file written by the UNMODIFIED
    ReconstructedFlow.__init__ (flowGenerator.py:14-51) on synthetic blobs (tests/golden/gen_golden_spod.py)."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from _spod_blobs import spod_blobs
    g = load_golden("spod")
    for tag, cplx in (("c", True), ("r", False)):
        modes, coeffs = spod_blobs(g["ltm"].shape, int(g["n_modes"]), int(g["nt"]), int(g["seed"]), cplx)
        base = o.FlowOracle.reconstruct(modes, coeffs, g["ltm"])
        assert np.abs(base[::2, ::3, ::4, :] - g[tag + "_base_sample"]).max() < 1e-13
        assert np.abs(base.sum(axis=(1, 2)) - g[tag + "_base_plane_sum"]).max() < 1e-9
        assert np.abs(base[-1] - g[tag + "_base_last"]).max() < 1e-13
        flow = o.FlowOracle(base, float(g[tag + "_baseDx"]), float(g[tag + "_baseDy"]), float(g[tag + "_baseDt"]))
        up, vp, ti, base_ti = flow.intensity()
        assert np.abs(up - g[tag + "_uPrime"]).max() < 1e-13 and np.abs(vp - g[tag + "_vPrime"]).max() < 1e-13
        assert np.abs(ti - g[tag + "_TI"]).max() < 1e-13 and abs(base_ti - float(g[tag + "_baseTI"])) < 1e-13
        flow.scale(11., 1.0, 2.0, translate=(-1.65, -1.1))
        assert rel_err(flow.interp(g[tag + "_interp_t"], g[tag + "_interp_xy"]), g[tag + "_interp_res"]) < 1e-12


def test_replay_buffer_oracle_vs_reference_class():
    """oracle.ReplayBufferOracle against the arrays left behind by the UNMODIFIED CustomReplayBuffer
    (main_02_sbl_contrib_customBuffer.py:57-160; tests/golden/gen_golden_replay.py), add by add."""
    g = load_golden("replay")
    buf = o.ReplayBufferOracle(int(g["buffer_size_arg"]), int(g["n_envs"]))
    assert buf.buffer_size == int(g["slots"])
    for k in range(g["in_obs"].shape[0]):
        buf.add(g["in_obs"][k], g["in_next_obs"][k], g["in_act"][k], g["in_rew"][k], g["in_done"][k], g["in_timeout"][k])
        assert (buf.pos, int(buf.full), buf.nRollovers) == tuple(g["trace"][k])
    for name in ("observations", "next_observations", "actions", "rewards", "dones", "timeouts"):
        assert np.array_equal(getattr(buf, name), g["buf_" + name]), name


def test_heading_error_legacy():
    g = load_golden("legacy")
    got = o.angle_error(g["heading_pairs"][:, 0], g["heading_pairs"][:, 1])
    assert np.abs(got - g["heading_err"]).max() < 1e-15


def test_auv_episodes():
    g = load_golden("legacy")
    flow = make_flow(g)
    for e in range(g["ep_actions"].shape[0]):
        env = o.AuvEnvOracle(1, flow, noiseMagCoeffs=0.1, noiseMagActuation=0.1, stopOnBoundsExceeded=(e != 1),
                             max_steps=50 if e == 2 else 250)
        env.reset()
        ob0 = env.set_initial(g["ep_mults"][e:e + 1], g["ep_pos0"][e:e + 1], g["ep_heading0"][e:e + 1],
                              g["ep_heading_target"][e:e + 1], g["ep_t_offset"][e:e + 1])
        assert np.abs(ob0[0] - g["ep_obs0"][e]).max() < 1e-14
        for k in range(g["ep_actions"].shape[1]):
            ob, r, d, _ = env.step(g["ep_actions"][e, k:k + 1])
            assert np.abs(ob[0] - g["ep_obs"][e, k]).max() < 1e-11, (e, k)
            assert abs(r[0] - g["ep_reward"][e, k]) < 1e-10, (e, k)
            assert bool(d[0]) == bool(g["ep_done"][e, k])
            h = g["ep_history"][e, k]
            # Fx Fy N Fx_set Fy_set N_set (cols 9-14), u_current v_current rmsAc (18-20), reward terms (21-25)
            last = env.last
            assert rel_err(np.concatenate([last["Fhydro"][0], last["Fset"][0], [last["Nset"][0]]]), h[9:15]) < 1e-12  # episode 1 diverges (Euler, no bounds stop)
            assert np.abs(np.concatenate([last["vel_current"][0], [last["rmsAc"][0]]]) - h[18:21]).max() < 1e-11
            assert np.abs(last["terms"][0] - h[21:26]).max() < 1e-10
            if d[0]:
                break


def test_los_navigation_oracle_vs_reference_golden():
    g = load_golden("agents")
    t = o.line_of_sight(g["los_p0"], g["los_p1"], g["los_rnav"])
    assert np.abs(t - g["los_target"]).max() < 1e-14 and not np.isnan(t).any()
    assert np.abs(o.los_navigation_predict(g["nav_obs"]) - g["nav_action"]).max() < 1e-14


def make_modes_flow(nt=48):
    g = load_golden("legacy")
    rng = np.random.default_rng(3)
    ny, nx, _ = g["ltm"].shape
    t, y, x = np.meshgrid(np.arange(nt), np.arange(ny), np.arange(nx), indexing="ij")
    f = np.repeat(g["ltm"][None], nt, axis=0).copy()
    for c in range(3):
        for _ in range(4):
            kt, ky, kx = rng.uniform(0.05, 0.3), rng.uniform(0.02, 0.12), rng.uniform(0.02, 0.12)
            f[..., c] += 0.02 * np.sin(kt * t + ky * y + kx * x + rng.uniform(0, 2 * np.pi))
    flow = o.FlowOracle(f, float(g["base_dx"]), float(g["base_dy"]), float(g["base_dt"]))
    flow.scale(11., 1.0, 2.0, translate=(-1.65, -1.1))
    return flow


def test_auv_cyl_episodes_vs_reference_golden():
    """AuvEnvCyl (legacy/verySimpleAuv_cyl.py): way-point switch, iWp carried across episodes, V0 observation scaling."""
    g = load_golden("legacy_cyl")
    flow = make_modes_flow()
    assert abs(flow.dt - float(g["flow_dt"])) < 1e-18
    wps, thr = o.cyl_waypoints()
    assert np.abs(wps - g["waypoints"]).max() < 1e-15 and thr == float(g["wp_threshold"])
    env = o.AuvCylEnvOracle(1, flow, noiseMagCoeffs=0.1, noiseMagActuation=0.1)
    env.reset()
    for e in range(g["ep_actions"].shape[0]):
        assert env.i_wp[0] == g["ep_iwp0"][e]          # never reset between episodes, as upstream
        ob0 = env.set_initial(g["ep_mults"][e:e + 1], g["ep_pos0"][e:e + 1], g["ep_heading0"][e:e + 1], None, g["ep_t_offset"][e:e + 1])
        assert np.abs(ob0[0] - g["ep_obs0"][e]).max() < 1e-13
        for k in range(g["ep_actions"].shape[1]):
            ob, r, d, _ = env.step(g["ep_actions"][e, k:k + 1])
            assert np.abs(ob[0] - g["ep_obs"][e, k]).max() < 1e-9, (e, k)
            assert abs(r[0] - g["ep_reward"][e, k]) < 1e-9, (e, k)
            assert bool(d[0]) == bool(g["ep_done"][e, k]) and env.i_wp[0] == g["ep_iwp"][e, k]
            if d[0]:
                break
    assert g["ep_iwp"].max() >= 1   # the golden episodes do contain a way-point switch
