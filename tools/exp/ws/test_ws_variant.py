"""Bitwise test of the warp-specialised persistent variant (experiment record, see tools/exp/ws/README.md).
Not collected by the test suite: the variant is no longer linked into libmvrl.so."""
import numpy as np
import torch

from marinevehiclereinforcementlearning_b200 import BlueROV2Heavy6DoFVecEnv

DEV = "cuda"


def test_warp_specialised_kernel_matches_plain_kernel_bitwise(monkeypatch):
    """The opt-in warp-specialised persistent variant (csrc/rov6_ws_kernel.cuh: IO warps + compute warps handing
    tiles over through mbarriers, MVRL_WS=1) runs the same arithmetic per environment as the plain fused kernel:
    bitwise equal observations, dones, states, way-points, counters and statistics, with auto-reset, for batch sizes
    that leave lanes, warps and whole CTAs of the persistent grid empty."""
    kw = dict(action_mode="rpm", dtype=torch.float32, device=DEV, maxSteps=3, auto_reset=True, seed=5)
    for n in (1, 65, 777, 100001):
        rng = np.random.default_rng(n)
        acts = torch.as_tensor(rng.uniform(-3500, 3500, (7, n, 8)), dtype=torch.float32, device=DEV)
        monkeypatch.setenv("MVRL_WS", "0")
        plain = BlueROV2Heavy6DoFVecEnv(n, **kw)
        monkeypatch.setenv("MVRL_WS", "1")
        ws = BlueROV2Heavy6DoFVecEnv(n, **kw)
        assert torch.equal(plain.reset(), ws.reset())
        for k in range(acts.shape[0]):
            op, _, dp, ip = plain.step(acts[k])
            ow, _, dw, iw = ws.step(acts[k])
            assert torch.equal(op, ow) and torch.equal(dp, dw), (n, k)
            assert torch.equal(ip["terminal_observation"], iw["terminal_observation"]), (n, k)
            assert torch.equal(plain._state, ws._state) and torch.equal(plain._path, ws._path) and torch.equal(plain._istep, ws._istep), (n, k)
        assert plain.episode_stats() == ws.episode_stats()
