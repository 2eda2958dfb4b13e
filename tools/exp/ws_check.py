"""Warp-specialised step kernel vs the plain one: bitwise comparison over several batch sizes incl. auto-reset."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from marinevehiclereinforcementlearning_b200 import BlueROV2Heavy6DoFVecEnv


def run(n, ws, steps=7):
    os.environ["MVRL_WS"] = "1" if ws else "0"
    env = BlueROV2Heavy6DoFVecEnv(n, action_mode="rpm", dtype=torch.float32, device="cuda", maxSteps=3, auto_reset=True, seed=5)
    env.reset()
    gen = torch.Generator(device="cuda").manual_seed(1)
    outs = []
    for k in range(steps):
        a = (torch.rand((n, 8), generator=gen, device="cuda") * 2 - 1) * 3500.0
        obs, rew, done, info = env.step(a)
        outs.append((obs.clone(), done.clone(), env.systemState.clone(), info["terminal_observation"].clone(), env.path.clone(), env.iStep.clone()))
    torch.cuda.synchronize()
    return outs, env.episode_stats()


for n in (1, 2, 63, 64, 65, 777, 4096, 100001, 1 << 20):
    a, sa = run(n, False)
    b, sb = run(n, True)
    for k, (x, y) in enumerate(zip(a, b)):
        for u, v in zip(x, y):
            assert torch.equal(u, v), (n, k)
    assert sa == sb, (sa, sb)
    print("n=%d ok" % n, flush=True)
print("ws check ok")
