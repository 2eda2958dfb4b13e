"""Round-2 diagnostic (not a test): error tables of the CUDA step kernels against the CPU oracles, used to choose the
asserted bounds of tests/test_parity_modes_gpu.py.  Run on the GPU box:  python tools/exp/r2_parity_probe.py [tag]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import c_oracle as c          # noqa: E402
from oracle import oracle_np as o         # noqa: E402
from marinevehiclereinforcementlearning_b200 import BlueROV2Heavy6DoFVecEnv, BlueROV2Heavy3DoFVecEnv  # noqa: E402

DEV = "cuda"
MODES = {"rpm": o.MODE_RPM, "force": o.MODE_FORCE, "setpoint": o.MODE_PID}
SCALE6 = {"rpm": np.full(8, 3500.0), "force": np.array([50., 50., 50., 1., 1., 2.]), "setpoint": np.ones(6)}


def adiff(a, b, ang):
    d = np.abs(a - b)
    d[:, ang] = np.abs((d[:, ang] + np.pi) % (2 * np.pi) - np.pi)
    return d


def sync6(env, ref, mode):
    n, dt = env.num_envs, env.dtype
    t = lambda x: torch.as_tensor(np.ascontiguousarray(x.T), dtype=dt, device=DEV)
    env._state[:, :n] = t(ref.state)
    env._setpoint[:, :n] = t(ref.set_point)
    env._path[:, :n] = t(ref.path)
    env._istep[:n] = torch.as_tensor(ref.i_step, dtype=torch.int32, device=DEV)
    if mode == "setpoint":
        cn = ref.ctrl_np
        e_old = cn["eOld"].copy()
        e_old[cn["has_old"] == 0, 0] = np.nan
        env._ctrl[0:6, :n] = t(e_old)
        env._ctrl[6:12, :n] = t(cn["eInt"])
        env._ctrl[12, :n] = torch.as_tensor(cn["tOld"], dtype=dt, device=DEV)


def table(name, err, key, edges):
    print("  %s: max err by %s" % (name, key[0]))
    k = key[1]
    for lo, hi in zip(edges[:-1], edges[1:]):
        m = (k >= lo) & (k < hi)
        if m.any():
            print("    [%8.1e, %8.1e): n=%8d  max=%.3e  p99=%.3e  frac>1e-4=%.2e" % (lo, hi, m.sum(), err[m].max(), np.percentile(err[m], 99), (err[m] > 1e-4).mean()))


def local6(mode, dtype, n=4096, steps=120, seed=1):
    ref = c.Rov6EnvC(n, mode=MODES[mode], max_steps=10 ** 9)
    ref.reset()
    env = BlueROV2Heavy6DoFVecEnv(n, action_mode=mode, dtype=dtype, device=DEV, auto_reset=False, maxSteps=10 ** 9)
    env.reset()
    rng = np.random.default_rng(seed)
    errs, mcs, mgs, oerr = [], [], [], []
    for k in range(steps):
        a = torch.as_tensor(rng.uniform(-1, 1, (n, len(SCALE6[mode]))) * SCALE6[mode], dtype=dtype)
        sync6(env, ref, mode)
        ref.mincos[:] = 1.0
        ref.ctrl_np["margin"] = np.inf
        obs, _, _, _ = env.step(a.to(DEV))
        ro, _, _, _ = ref.step(a.to(torch.float64).numpy())
        d = adiff(env.systemState.cpu().numpy().astype(np.float64), ref.state, slice(3, 6)) / (1.0 + np.abs(ref.state))
        errs.append(d.max(axis=1)); mcs.append(ref.mincos.copy()); mgs.append(ref.ctrl_np["margin"].copy())
        oerr.append(np.abs(obs.cpu().numpy() - ro).max(axis=1))
    err, mc, mg, oe = map(np.concatenate, (errs, mcs, mgs, oerr))
    print("local6 %s %s: env-steps=%d  max=%.3e median=%.3e  obs max=%.3e" % (mode, dtype, err.size, err.max(), np.median(err), oe.max()))
    table("state", err, ("mincos", mc), [0, 1e-3, 1e-2, 3e-2, 0.1, 0.3, 1.01])
    if mode == "setpoint":
        table("state", err, ("margin", mg), [0, 1e-12, 1e-11, 1e-10, 1e-9, 1e-8, 1e-7, 1e-6, 1e-5, 1e-3, np.inf])


def traj6(mode, dtype, n=4096, steps=1000, seed=2):
    ref = c.Rov6EnvC(n, mode=MODES[mode], max_steps=10 ** 9)
    ref.reset()
    env = BlueROV2Heavy6DoFVecEnv(n, action_mode=mode, dtype=dtype, device=DEV, auto_reset=False, maxSteps=10 ** 9)
    env.reset()
    sync6(env, ref, mode)
    gen = torch.Generator(device="cpu").manual_seed(1234 + seed)
    na = len(SCALE6[mode])
    worst = np.zeros(n)
    for k in range(steps):
        a = ((torch.rand((n, na), generator=gen, dtype=torch.float64) * 2 - 1) * torch.as_tensor(SCALE6[mode])).to(dtype)
        env.step(a.to(DEV))
        ref.step(a.to(torch.float64).numpy())
        if k % 10 == 9:
            d = adiff(env.systemState.cpu().numpy().astype(np.float64), ref.state, slice(3, 6)) / (1.0 + np.abs(ref.state))
            worst = np.maximum(worst, d.max(axis=1))
        if k + 1 in (100, 250, 500, 1000):
            msg = "  traj6 %s %s step %4d:" % (mode, dtype, k + 1)
            for tol in (1e-8, 1e-6, 1e-5, 1e-4, 1e-3):
                msg += "  frac<=%.0e: %.4f" % (tol, (worst <= tol).mean())
            msg += "  | mincos>=0.3: %.3f, of those <=1e-4: %.4f; median %.2e" % ((ref.mincos >= 0.3).mean(), (worst[ref.mincos >= 0.3] <= 1e-4).mean() if (ref.mincos >= 0.3).any() else -1, np.median(worst))
            print(msg)


def sync3(env, ref, mode):
    n, dt = env.num_envs, env.dtype
    t = lambda x: torch.as_tensor(np.ascontiguousarray(x.T), dtype=dt, device=DEV)
    env._state[:, :n] = t(ref.state)
    env._setpoint[:, :n] = t(ref.set_point)
    env._path[:, :n] = t(ref.path)
    env._istep[:n] = torch.as_tensor(ref.i_step, dtype=torch.int32, device=DEV)
    if mode == "setpoint":
        e_old = ref.ctrl["eOld"].copy()
        e_old[~ref.ctrl["has_old"], 0] = np.nan
        env._ctrl[0:3, :n] = t(e_old)
        env._ctrl[3:6, :n] = t(ref.ctrl["eInt"])
        env._ctrl[6, :n] = torch.as_tensor(ref.ctrl["tOld"], dtype=dt, device=DEV)


def run3(mode, dtype, local, n=2048, steps=200, seed=3):
    ref = o.Rov3EnvOracle(n, mode=MODES[mode], max_steps=10 ** 9)
    ref.reset()
    env = BlueROV2Heavy3DoFVecEnv(n, action_mode=mode, dtype=dtype, device=DEV, auto_reset=False, maxSteps=10 ** 9)
    env.reset()
    sync3(env, ref, mode)
    rng = np.random.default_rng(seed)
    na, sc = (4, 3500.0) if mode == "rpm" else (3, 1.0)
    errs, mgs = [], []
    worst = np.zeros(n)
    for k in range(steps):
        a = torch.as_tensor(rng.uniform(-1, 1, (n, na)) * sc, dtype=dtype)
        if local:
            sync3(env, ref, mode)
        ref.ctrl["margin"] = np.full(n, np.inf)
        env.step(a.to(DEV))
        ref.step(a.to(torch.float64).numpy())
        d = (adiff(env.systemState.cpu().numpy().astype(np.float64), ref.state, slice(2, 3)) / (1.0 + np.abs(ref.state))).max(axis=1)
        if local:
            errs.append(d); mgs.append(ref.ctrl["margin"].copy())
        else:
            worst = np.maximum(worst, d)
            if k + 1 in (100, 300, 1000):
                print("  traj3 %s %s step %4d: " % (mode, dtype, k + 1) + "  ".join("frac<=%.0e: %.4f" % (t, (worst <= t).mean()) for t in (1e-8, 1e-6, 1e-5, 1e-4, 1e-3)) + "  median %.2e max %.2e" % (np.median(worst), worst.max()))
    if local:
        err, mg = np.concatenate(errs), np.concatenate(mgs)
        print("local3 %s %s: env-steps=%d max=%.3e median=%.3e" % (mode, dtype, err.size, err.max(), np.median(err)))
        if mode == "setpoint":
            table("state", err, ("margin", mg), [0, 1e-12, 1e-11, 1e-10, 1e-9, 1e-8, 1e-7, 1e-6, 1e-5, 1e-3, np.inf])


if __name__ == "__main__":
    print("lib:", os.environ.get("MVRL_LIB", "default"))
    quick = len(sys.argv) > 1 and sys.argv[1] == "nodp"
    t0 = time.time()
    if quick:   # the literal e - eOld build: only the set-point fp32 cases differ
        local6("setpoint", torch.float32)
        traj6("setpoint", torch.float32, steps=500)
        run3("setpoint", torch.float32, True)
        run3("setpoint", torch.float32, False, steps=300)
    else:
        for dtype in (torch.float32, torch.float64):
            for mode in ("rpm", "force", "setpoint"):
                local6(mode, dtype)
        for mode, dtype in (("setpoint", torch.float32), ("setpoint", torch.float64), ("force", torch.float32), ("force", torch.float64)):
            traj6(mode, dtype)
        for dtype in (torch.float32, torch.float64):
            for mode in ("rpm", "setpoint"):
                run3(mode, dtype, True)
        run3("setpoint", torch.float32, False, steps=1000)
        run3("setpoint", torch.float64, False, steps=300)
        run3("rpm", torch.float32, False, steps=1000)
    print("probe time %.1f s" % (time.time() - t0))
