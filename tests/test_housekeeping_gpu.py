"""GPU: checkpoint round trip, device handling of the C ABI, config-4-sized legacy parity, launch-shape equivalence."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import oracle_np as o

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from marinevehiclereinforcementlearning_b200 import (AuvVecEnv, BlueROV2Heavy3DoFVecEnv, BlueROV2Heavy6DoFVecEnv)
    from marinevehiclereinforcementlearning_b200 import resources as res
    from marinevehiclereinforcementlearning_b200.tag_00_Dec2023_simpleControlTurbulence import flowGenerator

DEV = "cuda"


def _nan_eq(a, b):
    return torch.equal(torch.nan_to_num(a, nan=12345.0), torch.nan_to_num(b, nan=12345.0))


@pytest.mark.parametrize("kind", ["rov6-setpoint", "rov6-rpm", "rov3-setpoint", "auv"])
def test_state_dict_round_trip_is_bitwise(kind):
    """step k, save, step m, restore, step m again => bitwise equal (state, controller, counters, way-points, auto-reset draws)."""
    n, k, m = 777, 6, 9
    rng = np.random.default_rng(8)
    if kind == "auv":
        ltm = load_golden("legacy")["ltm"]
        flow = flowGenerator.ReconstructedFlow.synthetic(lt_mean=ltm, nt=64, seed=7, sigma=0.05, kind="modes", dtype=torch.float32, device=DEV)
        flow.scale(11., 1.0, 2.0, translate=(-1.65, -1.1))
        env = AuvVecEnv(n, flow, seed=3, noiseMagCoeffs=0.1, noiseMagActuation=0.1, auto_reset=True, dtype=torch.float32)
        na, sc = 3, 1.0
    elif kind.startswith("rov3"):
        env = BlueROV2Heavy3DoFVecEnv(n, action_mode="setpoint", dtype=torch.float32, device=DEV, maxSteps=4, auto_reset=True, seed=3)
        na, sc = 3, 1.0
    else:
        mode = kind.split("-")[1]
        env = BlueROV2Heavy6DoFVecEnv(n, action_mode=mode, dtype=torch.float32, device=DEV, maxSteps=4, auto_reset=True, seed=3)
        na, sc = (8, 3500.0) if mode == "rpm" else (6, 1.0)
    acts = torch.as_tensor(rng.uniform(-sc, sc, (k + m, n, na)), dtype=torch.float32, device=DEV)
    env.reset()
    for i in range(k):
        env.step(acts[i])
    saved = env.state_dict()
    first = []
    for i in range(k, k + m):
        obs, rew, done, _ = env.step(acts[i])
        first.append((obs.clone(), rew.clone(), done.clone()))
    end = {key: v.clone() if torch.is_tensor(v) else v for key, v in env.state_dict().items()}
    env.load_state_dict(saved)
    for j, i in enumerate(range(k, k + m)):
        obs, rew, done, _ = env.step(acts[i])
        assert torch.equal(obs, first[j][0]) and torch.equal(rew, first[j][1]) and torch.equal(done, first[j][2]), (kind, j)
    again = env.state_dict()
    for key, v in end.items():
        assert _nan_eq(again[key], v) if torch.is_tensor(v) else again[key] == v, (kind, key)


def test_compile_time_constant_kernels_match_runtime_constant_kernels_bitwise(monkeypatch):
    """The default vehicle's fp32 step kernels carry its constants as literals (CONSTP, csrc/rov6_default_consts.h) - chosen
    only when the handle's constants equal the compiled-in ones bit for bit; MVRL_NO_CONSTP=1 keeps the kernels that read them
    from the kernel argument.  Same values, same operations: bitwise equal, every action mode, with auto-reset."""
    n, steps = 20001, 8
    for mode, na, sc in (("rpm", 8, 3500.0), ("setpoint", 6, 1.0), ("force", 6, 40.0)):
        rng = np.random.default_rng(4)
        acts = torch.as_tensor(rng.uniform(-sc, sc, (steps, n, na)), dtype=torch.float32, device=DEV)
        kw = dict(action_mode=mode, dtype=torch.float32, device=DEV, maxSteps=3, auto_reset=True, seed=9)
        monkeypatch.setenv("MVRL_NO_CONSTP", "0")
        lit = BlueROV2Heavy6DoFVecEnv(n, **kw)
        o_lit = lit.reset().clone()          # the handle is created (and reads the environment variable) at first use
        monkeypatch.setenv("MVRL_NO_CONSTP", "1")
        arg = BlueROV2Heavy6DoFVecEnv(n, **kw)
        assert torch.equal(o_lit, arg.reset())
        assert lit._get_handle().specialisation == 2 and arg._get_handle().specialisation == 1
        for k in range(steps):
            ol, _, dl, _ = lit.step(acts[k])
            oa, _, da, _ = arg.step(acts[k])
            assert torch.equal(ol, oa) and torch.equal(dl, da), (mode, k)
            assert torch.equal(lit._state, arg._state) and _nan_eq(lit._ctrl, arg._ctrl), (mode, k)
        assert lit.episode_stats() == arg.episode_stats()
    monkeypatch.delenv("MVRL_NO_CONSTP")
    # a vehicle that differs in one coefficient keeps the run-time constants
    from marinevehiclereinforcementlearning_b200 import Rov6Constants
    veh = Rov6Constants()
    veh.Xuu = -18.0
    other = BlueROV2Heavy6DoFVecEnv(64, action_mode="setpoint", dtype=torch.float32, device=DEV, vehicle=veh)
    other.reset()
    assert other._get_handle().specialisation == 1


def test_calls_leave_the_callers_current_device_alone():
    """a handle's entry points run on the handle's device and restore the caller's; the stateless helpers run
    where their pointers live.  With one GPU the guard is exercised with the same ordinal; with two, across devices."""
    n_dev = torch.cuda.device_count()
    target = n_dev - 1
    torch.cuda.set_device(0)
    env = BlueROV2Heavy6DoFVecEnv(64, action_mode="rpm", dtype=torch.float32, device="cuda:%d" % target, auto_reset=True)
    env.reset()
    env.step(torch.zeros((64, 8), device="cuda:%d" % target))
    assert torch.cuda.current_device() == 0
    a = torch.tensor([0.1, 3.0], dtype=torch.float64, device="cuda:%d" % target)
    b = torch.tensor([6.2, 0.0], dtype=torch.float64, device="cuda:%d" % target)
    e = res.angleError(a, b)
    assert torch.cuda.current_device() == 0 and e.device.index == target
    assert np.allclose(e.cpu().numpy(), [0.1831853071795857, 3.0], atol=1e-15)
    torch.cuda.synchronize(target)
    assert bool(torch.isfinite(env.systemState).all())


def test_auv_config4_size_three_steps_vs_oracle():
    """BASELINE config 4 at its full size: 262 144 legacy environments, 3 steps, against the numpy oracle (fp64: 1e-9)."""
    n = 262144
    ltm = load_golden("legacy")["ltm"]
    base = flowGenerator.synthetic_base_field(ltm, 96, seed=7, sigma=0.05, kind="modes")
    flow = flowGenerator.ReconstructedFlow.from_base_field(base, lt_mean=ltm, dtype=torch.float64, device=DEV)
    flow.scale(11., 1.0, 2.0, translate=(-1.65, -1.1))
    fo = o.FlowOracle(np.asarray(base, dtype=float))
    fo.scale(11., 1.0, 2.0, translate=(-1.65, -1.1))
    env = AuvVecEnv(n, flow, seed=1234, noiseMagCoeffs=0.1, noiseMagActuation=0.1, auto_reset=True, dtype=torch.float64)
    ref = o.AuvEnvOracle(n, fo, noiseMagCoeffs=0.1, noiseMagActuation=0.1, auto_reset=True, seed=1234)
    o0, r0 = env.reset().cpu().numpy(), ref.reset()
    assert np.abs(o0 - r0).max() < 1e-12
    rng = np.random.default_rng(2)
    for k in range(3):
        a = rng.uniform(-1, 1, (n, 3))
        obs, rew, done, _ = env.step(torch.as_tensor(a, device=DEV))
        ro, rr, rd, _ = ref.step(a)
        assert np.array_equal(done.cpu().numpy(), rd), k
        assert np.abs(obs.cpu().numpy() - ro).max() < 1e-9, k
        assert np.abs(rew.cpu().numpy() - rr).max() < 1e-9, k
