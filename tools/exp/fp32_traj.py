"""Experiment: fp32 kernel vs the fp64 C oracle over 1000 steps at 4096 envs (error distribution vs conditioning)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import oracle_np as o, c_oracle as c
from marinevehiclereinforcementlearning_b200 import BlueROV2Heavy6DoFVecEnv
n, steps = 4096, 1000
gen = torch.Generator(device="cpu").manual_seed(1234)
env = BlueROV2Heavy6DoFVecEnv(n, action_mode="rpm", dtype=torch.float32, device="cuda", auto_reset=False, maxSteps=10**9)
env.reset(initialSetpoint=np.zeros(6))
ref = c.Rov6EnvC(n, mode=o.MODE_RPM, max_steps=10 ** 9)
ref.reset(initial_setpoint=np.zeros(6))
worst = np.zeros(n)
first_bad = np.full(n, -1)
for k in range(steps):
    a = (torch.rand((n, 8), generator=gen, dtype=torch.float64) * 2 - 1) * 3500.0
    a32 = a.to(torch.float32)
    env.step(a32.to("cuda"))
    ref.step(a32.to(torch.float64).numpy())
    if k % 10 == 9:
        d = np.abs(env.systemState.cpu().numpy().astype(np.float64) - ref.state)
        d[:, 3:6] = np.abs((d[:, 3:6] + np.pi) % (2 * np.pi) - np.pi)
        e = (d / (1.0 + np.abs(ref.state))).max(axis=1)
        worst = np.maximum(worst, e)
        nb = (e > 1e-4) & (first_bad < 0)
        first_bad[nb] = k
for thr in (0.3, 0.1, 0.05, 0.02, 0.01, 0.0):
    good = ref.mincos >= thr
    print("mincos >= %.2f: %4d envs, worst %.3e, >1e-4: %d" % (thr, good.sum(), worst[good].max() if good.any() else 0, (worst[good] > 1e-4).sum()))
print("quantiles of worst:", np.quantile(worst, [0.5, 0.9, 0.99, 0.999, 1.0]))
