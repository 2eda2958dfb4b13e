"""Exercises every kernel family of libmvrl once on small, awkwardly sized batches (odd n, n < warp, masks,
both fp32 instantiations, host-buffer pipeline, NaN states).  Written for
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
(compute-sanitizer is closed on the round-1 pool, so it runs as a plain GPU test: tests/test_awkward_sizes_gpu.py)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from marinevehiclereinforcementlearning_b200 import (AuvVecEnv, BlueROV2Heavy3DoFVecEnv, BlueROV2Heavy6DoFVecEnv, Rov3Derivs, Rov6Derivs,
                                                     resources, vec_tools)
from marinevehiclereinforcementlearning_b200 import dynamicsModel_BlueROV2_Heavy_3DoF as m3
from marinevehiclereinforcementlearning_b200.tag_00_Dec2023_simpleControlTurbulence import flowGenerator

dev = "cuda"
rng = np.random.default_rng(0)
for n in (1, 31, 130, 257):
    for dtype in (torch.float32, torch.float64):
        for mode, na, sc in (("rpm", 8, 3500.), ("force", 6, 40.), ("setpoint", 6, 1.)):
            for x2 in ("0", "1"):
                os.environ["MVRL_NO_X2"] = x2
                env = BlueROV2Heavy6DoFVecEnv(n, action_mode=mode, dtype=dtype, device=dev, maxSteps=3, auto_reset=True, seed=1, record_aux=True)
                env.reset()
                for k in range(5):
                    env.step(torch.as_tensor(rng.uniform(-sc, sc, (n, na)), dtype=dtype, device=dev))
                env.reset(mask=torch.arange(n, device=dev) % 2 == 0)
                env.step_host(torch.as_tensor(rng.uniform(-sc, sc, (n, na)), dtype=dtype).pin_memory(), chunks=3)
                env.step_host(torch.as_tensor(rng.uniform(-sc, sc, (n, na)), dtype=dtype).pin_memory(), chunks=-2)
                if n > 40:
                    env.step_range_async(33, n - 40)
                env.episode_stats()
        s = torch.as_tensor(rng.uniform(-1, 1, (12, n)), dtype=dtype, device=dev)
        Rov6Derivs(dtype=dtype, action_mode="rpm")(s, torch.zeros((8, n), dtype=dtype, device=dev), want_aux=True)
        e3 = BlueROV2Heavy3DoFVecEnv(n, action_mode="setpoint", dtype=dtype, device=dev, maxSteps=3, auto_reset=True, record_aux=True)
        ob = e3.reset()
        for k in range(5):
            a, _ = m3.LOSNavigation().predict(ob)
            ob, _, _, _ = e3.step(a)
        Rov3Derivs(dtype=dtype, action_mode="rpm")(torch.zeros((6, n), dtype=dtype, device=dev), torch.ones((4, n), dtype=dtype, device=dev) * 900.)
        resources.angleError(torch.zeros(n, dtype=dtype, device=dev), torch.ones(n, dtype=dtype, device=dev))
        resources.coordinateTransform(*(torch.zeros(n, dtype=dtype, device=dev),) * 3, dof=6)
        ltm = np.load(os.path.join(ROOT, "tests", "golden", "golden_legacy.npz"))["ltm"]
        flow = flowGenerator.ReconstructedFlow.synthetic(lt_mean=ltm, nt=16, kind="modes", dtype=dtype, device=dev)
        flow.scale(11., 1., 2., translate=(-1.65, -1.1))
        for stage in ("0", "1"):
            os.environ["MVRL_AUV_NO_STAGE"] = stage
            ea = AuvVecEnv(n, flow, dtype=dtype, maxSteps=4, auto_reset=True, noiseMagCoeffs=0.1, record_aux=True)
            prev = ea.reset().T.contiguous()
            buf = vec_tools.SymmetryReplayBuffer(7 * n, n, dtype=dtype, device=dev)   # 7 slots of n transitions
            ea._state[0, : max(1, n // 3)] = float("nan")
            for k in range(6):
                act = torch.as_tensor(rng.uniform(-1, 1, (n, 3)), dtype=dtype, device=dev)
                ea.step(act)
                buf.add(prev, ea._obs, ea._action, ea._reward, ea._done)
                prev = ea._obs.clone()
        flow.interp(torch.linspace(-1, 2, n, dtype=dtype, device=dev), torch.zeros((n, 2), dtype=dtype, device=dev))
        flow.interpField(0.01)
torch.cuda.synchronize()
print("sanitize smoke ok")
