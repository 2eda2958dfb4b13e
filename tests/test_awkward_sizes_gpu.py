"""GPU: every kernel family on small, awkwardly sized batches (n = 1, 31, 130, 257; odd leading dimensions
and offsets; masks; NaN states; both fp32 instantiations; the host-buffer pipeline) must run without a
CUDA fault.  The script is tools/sanitize_smoke.py (also usable under compute-sanitizer where that is open)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_all_kernels_on_awkward_sizes():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_smoke.py")], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "sanitize smoke ok" in res.stdout


def test_empty_batch():
    """num_envs = 0 is a valid (ragged-shard) batch: every env call is a no-op with correctly shaped empty results."""
    import numpy as np
    import torch
    from marinevehiclereinforcementlearning_b200 import AuvVecEnv, BlueROV2Heavy3DoFVecEnv, BlueROV2Heavy6DoFVecEnv
    from marinevehiclereinforcementlearning_b200.tag_00_Dec2023_simpleControlTurbulence import flowGenerator
    dev = "cuda"
    for mode, na in (("rpm", 8), ("setpoint", 6)):
        env = BlueROV2Heavy6DoFVecEnv(0, action_mode=mode, device=dev, auto_reset=True)
        assert tuple(env.reset().shape) == (0, 9)
        obs, rew, done, info = env.step(torch.zeros((0, na), device=dev))
        assert tuple(obs.shape) == (0, 9) and tuple(rew.shape) == (0,) and tuple(done.shape) == (0,)
        h_obs, h_rew, h_done = env.step_host(torch.zeros((0, na)))
        assert tuple(h_obs.shape) == (0, 9)
        assert env.episode_stats()["episodes"] == 0
    e3 = BlueROV2Heavy3DoFVecEnv(0, action_mode="setpoint", device=dev)
    assert tuple(e3.reset().shape) == (0, 5)
    assert tuple(e3.step(torch.zeros((0, 3), device=dev))[0].shape) == (0, 5)
    ltm = np.load(os.path.join(ROOT, "tests", "golden", "golden_legacy.npz"))["ltm"]
    flow = flowGenerator.ReconstructedFlow.synthetic(lt_mean=ltm, nt=8, kind="modes", dtype=torch.float32, device=dev)
    flow.scale(11., 1., 2., translate=(-1.65, -1.1))
    ea = AuvVecEnv(0, flow, dtype=torch.float32)
    assert tuple(ea.reset().shape) == (0, 11)
    assert tuple(ea.step(torch.zeros((0, 3), device=dev))[0].shape) == (0, 11)
    torch.cuda.synchronize()
