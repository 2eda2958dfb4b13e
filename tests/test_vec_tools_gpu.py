"""GPU: the callers / data formats either side of the step path: trajectory recorder
and episode CSVs, evaluate_agent, the SB3 VecEnv adapter with VecMonitor CSV, the symmetry replay buffer."""
import csv
import os

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import oracle_np as o

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from marinevehiclereinforcementlearning_b200 import AuvVecEnv, BlueROV2Heavy6DoFVecEnv, resources, vec_tools
    from marinevehiclereinforcementlearning_b200.tag_00_Dec2023_simpleControlTurbulence import flowGenerator, verySimpleAuv

DEV = "cuda"


def make_flow(g, dtype):
    base = g["ltm"][None] + 0.05 * np.random.default_rng(7).standard_normal((int(g["nt"]),) + g["ltm"].shape)
    return flowGenerator.ReconstructedFlow.from_base_field(base, float(g["base_dx"]), float(g["base_dy"]), float(g["base_dt"]), dtype=dtype, device=DEV)


def test_recorder_reproduces_reference_time_history(tmp_path):
    """6DoF: the recorder's 33-column rows for env 0 equal the reference env's timeHistory (golden)."""
    e6 = load_golden("env6")
    env = BlueROV2Heavy6DoFVecEnv(3, action_mode="setpoint", dtype=torch.float64, device=DEV, maxSteps=60, auto_reset=False, record_aux=True)
    rec = vec_tools.TrajectoryRecorder(env, num_record=2)
    env.reset(initialSetpoint=e6["fixed_sp"])
    rec.on_reset()
    for k in range(60):
        env.step(torch.zeros((3, 6), dtype=torch.float64, device=DEV))
        rec.on_step()
    df = rec.dataframe(0)
    assert list(df.columns) == list(e6["fixed_history_cols"]) and df.shape == (61, 33)
    ref = e6["fixed_history"]
    assert np.abs(df.values[:, :13] - ref[:, :13]).max() < 1e-8 and np.abs(df.values[:, 27:] - ref[:, 27:]).max() < 1e-12
    assert np.abs(df.values[1:, 13:19] - ref[1:, 13:19]).max() < 1e-5
    paths = rec.to_csv(str(tmp_path / "eval"))
    assert [os.path.basename(p) for p in paths] == ["ep_0.csv", "ep_1.csv"]
    back = np.loadtxt(paths[1], delimiter=",", skiprows=1)
    assert back.shape == (61, 33) and np.abs(back - rec.dataframe(1).values).max() < 1e-12


def test_recorder_legacy_columns_and_evaluate_agent(tmp_path):
    g = load_golden("legacy")
    flow = make_flow(g, torch.float64)
    flow.scale(11., 1.0, 2.0, translate=(-1.65, -1.1))
    e = 3
    env = AuvVecEnv(2, flow, dtype=torch.float64, maxSteps=250, auto_reset=False, record_aux=True)
    env.reset(applyNoise=False, fixedInitialValues=[g["ep_pos0"][e], float(g["ep_heading0"][e]), float(g["ep_heading_target"][e])])
    env._target[1, :] = float(g["ep_t_offset"][e])
    rec = vec_tools.TrajectoryRecorder(env, num_record=1)
    rec.on_reset()
    for k in range(20):
        a = torch.as_tensor(np.repeat(g["ep_actions"][e, k:k + 1], 2, axis=0), device=DEV)
        env.step(a)
        rec.on_step(a)
    df = rec.dataframe(0)
    assert list(df.columns) == verySimpleAuv.HISTORY_COLUMNS and df.shape == (20, 40)
    ref = g["ep_history"][e, :20]
    scale = np.abs(ref) + np.abs(ref).max(axis=1, keepdims=True)
    assert (np.abs(df.values - ref) / scale).max() < 1e-8
    # evaluate_agent with the reference-shaped single env + PD controller (tag_00.../verySimpleAuv.py:437-447)
    single = verySimpleAuv.AuvEnv(flow=make_flow(g, torch.float64))
    single._max_episode_steps = 30
    pd = verySimpleAuv.PDController(single.dt)
    mean, median, rewards = resources.evaluate_agent(pd, single, num_episodes=2, saveDir=str(tmp_path / "pd"),
                                                     init=[np.array([0.3, -0.2]), 1.0, 4.0])
    assert len(rewards) == 2 and np.isfinite(mean) and os.path.exists(str(tmp_path / "pd" / "ep_1.csv"))
    with open(str(tmp_path / "pd" / "ep_0.csv")) as fh:
        assert next(csv.reader(fh)) == verySimpleAuv.HISTORY_COLUMNS
    # batched: every environment runs the episode at once
    benv = AuvVecEnv(64, flow, dtype=torch.float64, maxSteps=30, auto_reset=False)
    mean_b, _, rewards_b = resources.evaluate_agent(verySimpleAuv.PDController(benv.dt), benv, num_episodes=1)
    assert len(rewards_b) == 64 and np.isfinite(mean_b)


def test_sb3_vecenv_protocol_and_monitor_csv(tmp_path):
    env = BlueROV2Heavy6DoFVecEnv(16, action_mode="setpoint", dtype=torch.float32, device=DEV, maxSteps=4, auto_reset=True, seed=1)
    venv = vec_tools.Sb3VecEnv(env, monitor_file=str(tmp_path / "run" / "agent_0"))
    assert venv.num_envs == 16 and venv.observation_space.shape == (9,) and venv.action_space.shape == (6,)
    obs = venv.reset()
    assert obs.shape == (16, 9) and obs.dtype == np.float32
    n_ep = 0
    for k in range(9):
        venv.step_async(np.random.default_rng(k).uniform(-1, 1, (16, 6)).astype(np.float32))
        obs, rew, done, infos = venv.step_wait()
        assert obs.shape == (16, 9) and rew.shape == (16,) and done.dtype == np.bool_ and len(infos) == 16
        if (k + 1) % 4 == 0:
            assert done.all()
            assert all(i["episode"]["l"] == 4 and i["episode"]["r"] == 0.0 and i["terminal_observation"].shape == (9,) for i in infos)
            assert all("TimeLimit.truncated" not in i for i in infos)   # like the reference envs: no TimeLimit wrapper upstream
            n_ep += 16
        else:
            assert not done.any() and all(i == {} for i in infos)
    venv.close()
    flagged = vec_tools.Sb3VecEnv(BlueROV2Heavy6DoFVecEnv(4, action_mode="setpoint", dtype=torch.float32, device=DEV, maxSteps=2, auto_reset=True),
                                  flag_time_limit=True)
    flagged.reset()
    for k in range(2):
        _, _, done, infos = flagged.step(np.zeros((4, 6), dtype=np.float32))
    assert done.all() and all(i["TimeLimit.truncated"] is True for i in infos)   # opt-in
    lines = open(str(tmp_path / "run" / "agent_0.monitor.csv")).read().splitlines()
    assert lines[0].startswith("#{") and lines[1] == "r,l,t" and len(lines) == 2 + n_ep
    assert venv.env_is_wrapped(object) == [False] * 16 and venv.get_attr("num_envs") == [16] * 16


def test_symmetry_replay_buffer_vs_reference_class():
    """Against the arrays the UNMODIFIED reference class ``CustomReplayBuffer`` (main_02_sbl_contrib_customBuffer.py:57-160)
    left behind on the same sequence of ``add`` calls (tests/golden/gen_golden_replay.py: 34 adds into 23 slots of 5 envs,
    past the third roll-over, with ``TimeLimit.truncated`` infos): slot bookkeeping after every add, final contents bitwise."""
    g = load_golden("replay")
    n, ld = int(g["n_envs"]), 32
    buf = vec_tools.SymmetryReplayBuffer(int(g["buffer_size_arg"]), n, dtype=torch.float32, device=DEV)
    assert buf.buffer_size == int(g["slots"])            # SB3: transitions // n_envs
    fm = lambda x: torch.as_tensor(np.pad(np.atleast_2d(x.T), ((0, 0), (0, ld - n))), device=DEV).contiguous()
    for k in range(g["in_obs"].shape[0]):
        rew = torch.as_tensor(np.pad(g["in_rew"][k], (0, ld - n)), device=DEV)
        done = torch.as_tensor(np.pad(g["in_done"][k].astype(np.uint8), (0, ld - n)), device=DEV)
        infos = [({"TimeLimit.truncated": True} if t else {}) for t in g["in_timeout"][k]]      # as VecEnv.step returns them
        buf.add(fm(g["in_obs"][k]), fm(g["in_next_obs"][k]), fm(g["in_act"][k]), rew, done, infos)
        assert (buf.pos, int(buf.full), buf.nRollovers) == tuple(g["trace"][k]), k
    for name in ("observations", "next_observations", "actions", "rewards", "dones", "timeouts"):
        got = getattr(buf, name).cpu().numpy().astype(np.float32)
        assert np.array_equal(got, g["buf_" + name]), name


def test_symmetry_replay_buffer_vs_oracle_larger_batch():
    """A batch that is not a multiple of anything (37 envs, padded rows), fp64, no timeouts, against oracle.ReplayBufferOracle
    (itself pinned to the reference class on the CPU side)."""
    n, slots, ld = 37, 23, 64
    ref = o.ReplayBufferOracle(slots * n, n, dtype=np.float64)
    buf = vec_tools.SymmetryReplayBuffer(slots * n, n, dtype=torch.float64, device=DEV)
    rng = np.random.default_rng(0)
    for k in range(30):
        ob, no, a = rng.uniform(-1, 1, (n, 11)), rng.uniform(-1, 1, (n, 11)), rng.uniform(-1, 1, (n, 3))
        r, d = rng.uniform(-3, 3, n), (rng.uniform(size=n) < 0.1).astype(np.uint8)
        ref.add(ob, no, a, r, d)
        fm = lambda x: torch.as_tensor(np.pad(x.T, ((0, 0), (0, ld - n))), device=DEV).contiguous()
        buf.add(fm(ob), fm(no), fm(a), torch.as_tensor(np.pad(r, (0, ld - n)), device=DEV), torch.as_tensor(np.pad(d, (0, ld - n)), device=DEV))
        assert (buf.pos, buf.full, buf.nRollovers) == (ref.pos, ref.full, ref.nRollovers)
    assert buf.nRollovers > 2   # both regimes (with and without mirror images) were exercised
    for name in ("observations", "next_observations", "actions", "rewards", "dones", "timeouts"):
        assert np.array_equal(getattr(buf, name).cpu().numpy().astype(np.float64), getattr(ref, name)), name
    s = buf.sample(256)
    assert s["observations"].shape == (256, 11) and s["actions"].shape == (256, 3) and s["dones"].dtype == torch.uint8


@pytest.mark.parametrize("mode,groups,n", [("rpm", 3, 10007), ("setpoint", 4, 4096), ("force", 2, 130)])
def test_env_blocks_on_their_own_streams_match_whole_batch_launches_bitwise(mode, groups, n):
    """EnvBlocks (independent blocks of environments, each a chain of launches on its own stream - the SubprocVecEnv workers
    of tag_00.../main_00_sbl.py:145-146 on the device) against one launch per step: same kernels, random draws keyed on the
    global environment id, so state / observation / done / episode counters / statistics agree bit for bit - eagerly and
    as parallel branches of one captured CUDA graph; auto-reset fires several times (maxSteps = 4)."""
    na = 8 if mode == "rpm" else 6
    scale = {"rpm": 3500.0, "force": 40.0, "setpoint": 1.0}[mode]
    K = 11
    g = torch.Generator(device=DEV).manual_seed(5)
    envs = [BlueROV2Heavy6DoFVecEnv(n, action_mode=mode, dtype=torch.float32, device=DEV, maxSteps=4, auto_reset=True, seed=3) for _ in range(3)]
    acts = [(torch.rand((na, envs[0].ld), generator=g, device=DEV) * 2 - 1) * scale for _ in range(3)]
    for e in envs:
        e.reset()
    whole, eager, graphed = envs
    for k in range(K):
        whole._bufs.action = acts[k % 3].data_ptr()
        whole.step_async()
    blocks = vec_tools.EnvBlocks(eager, groups)
    assert len(blocks) == groups and sum(c for _, c in blocks.blocks) == n and all(lo % 2 == 0 for lo, _ in blocks.blocks)
    with pytest.raises(RuntimeError):
        blocks.step_async()                     # outside fork() / join()
    with blocks:
        for k in range(K):
            eager._bufs.action = acts[k % 3].data_ptr()
            blocks.step_async()
    gb = vec_tools.EnvBlocks(graphed, groups)
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(device=DEV)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            with gb:
                for k in range(K):
                    graphed._bufs.action = acts[k % 3].data_ptr()
                    gb.step_async()
    torch.cuda.current_stream().wait_stream(side)
    graph.replay()
    torch.cuda.synchronize()
    stats = whole.episode_stats()               # (reading resets the accumulators)
    assert stats["episodes"] == 2 * n           # 11 steps with maxSteps = 4: every environment finished two episodes
    for e in (eager, graphed):
        assert torch.equal(e.systemState, whole.systemState) and torch.equal(e._obs, whole._obs)
        assert torch.equal(e._done, whole._done) and torch.equal(e._istep, whole._istep) and torch.equal(e._episode, whole._episode)
        assert e.episode_stats() == stats


def test_env_shards_on_their_own_streams_match_one_env_bitwise():
    """EnvShards: separate env objects with consecutive env_id0 (the reference's list of SubprocVecEnv workers), each a chain of
    launches on its own stream, against one env over all environments - legacy env, fp32, noise and auto-reset on."""
    g = load_golden("legacy")
    flow = make_flow(g, torch.float32)
    n, K = 3000, 12
    kw = dict(noiseMagCoeffs=0.1, noiseMagActuation=0.1, maxSteps=5, auto_reset=True, seed=9, dtype=torch.float32)
    whole = AuvVecEnv(n, flow, **kw)
    parts = [AuvVecEnv(n // 3, flow, env_id0=i * (n // 3), **kw) for i in range(3)]
    whole.reset()
    for p in parts:
        p.reset()
    gen = torch.Generator(device=DEV).manual_seed(2)
    acts = [torch.rand((n, 3), generator=gen, device=DEV) * 2 - 1 for _ in range(K)]
    for k in range(K):
        whole.step_async(acts[k])
    shards = vec_tools.EnvShards(parts)
    assert len(shards) == 3
    with pytest.raises(RuntimeError):
        shards.step_async()
    for k in range(K):                         # the actions go in on the caller's stream, before the fork
        for i, p in enumerate(parts):
            p.set_actions(acts[k][i * (n // 3):(i + 1) * (n // 3)])
        with shards:
            shards.step_async()
    torch.cuda.synchronize()
    m = n // 3
    for i, p in enumerate(parts):
        assert torch.equal(p._obs[:, :m], whole._obs[:, i * m:(i + 1) * m]) and torch.equal(p._state[:, :m], whole._state[:, i * m:(i + 1) * m])
        assert torch.equal(p._done[:m], whole._done[i * m:(i + 1) * m]) and torch.equal(p._reward[:m], whole._reward[i * m:(i + 1) * m])
    episodes = whole.episode_stats()["episodes"]
    assert sum(p.episode_stats()["episodes"] for p in parts) == episodes and episodes >= 2 * n   # maxSteps = 5, 12 steps (+ bounds terminations)
