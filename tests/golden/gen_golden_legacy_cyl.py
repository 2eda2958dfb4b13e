"""Golden vectors for the legacy way-point variant AuvEnvCyl
(tag_00_Dec2023_simpleControlTurbulence/verySimpleAuv_cyl.py:22-345), produced by executing the unmodified
reference in this container (own interpreter: legacy/resources.py clashes with /resources.py by name).

    python tests/golden/gen_golden_legacy_cyl.py

The SPOD blobs are absent upstream, so ReconstructedFlow's data-loading constructor is replaced by the
synthetic long-wave field (mean + travelling modes); scale(), interp() and all of AuvEnvCyl are the reference's code."""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_shims import LEGACY_ROOT, import_legacy  # noqa: E402

ref_flow, ref_auv, ref_lres = import_legacy()
ref_cyl = importlib.import_module("verySimpleAuv_cyl")
NT = 48


def modes_field(ltm, nt, seed=3, amp=0.02):
    """Same long-wave stand-in as tests/test_auv_gpu.py::smooth_base_field."""
    rng = np.random.default_rng(seed)
    ny, nx, _ = ltm.shape
    t, y, x = np.meshgrid(np.arange(nt), np.arange(ny), np.arange(nx), indexing="ij")
    f = np.repeat(ltm[None], nt, axis=0).copy()
    for c in range(3):
        for _ in range(4):
            kt, ky, kx = rng.uniform(0.05, 0.3), rng.uniform(0.02, 0.12), rng.uniform(0.02, 0.12)
            f[..., c] += amp * np.sin(kt * t + ky * y + kx * x + rng.uniform(0, 2 * np.pi))
    return f


class SyntheticFlow(ref_flow.ReconstructedFlow):
    def __init__(self, dataDir=None):
        d = os.path.join(LEGACY_ROOT, "turbulenceData")
        self.lt_mean = np.load(os.path.join(d, "ltm.npy"))
        self.baseFlowData = modes_field(self.lt_mean, NT)
        self.baseDt = 0.002
        self.baseTime = np.array([i * self.baseDt for i in range(NT)])
        self.baseCoords = np.load(os.path.join(d, "turbulence_coords.npy"))
        self.baseDx = (self.baseCoords[0, 1:, 0] - self.baseCoords[0, :-1, 0])[0]
        self.baseDy = (self.baseCoords[1:, 0, 1] - self.baseCoords[:-1, 0, 1])[0]
        self.scale(1., 1., 1.)


def main():
    ref_flow.ReconstructedFlow = SyntheticFlow
    ref_cyl.flowGenerator.ReconstructedFlow = SyntheticFlow
    out = {"nt": np.array(NT)}
    n_steps = 80
    ep = {k: [] for k in ("mults", "pos0", "heading0", "t_offset", "iwp0", "actions", "obs0", "obs", "reward", "done", "history", "iwp")}
    np.random.seed(11)
    env = ref_cyl.AuvEnvCyl(noiseMagCoeffs=0.1, noiseMagActuation=0.1)
    out["waypoints"] = env.waypoints.copy(); out["wp_threshold"] = np.array(env.wpThreshold)
    pd = ref_auv.PDController(env.dt)
    for e in range(4):   # ONE env object: iWp carries over between episodes exactly as upstream (it is set in __init__ only)
        if e == 1:       # start on top of the current way-point: the switch triggers in the first steps
            wp = env.waypoints[env.iWp]
            obs0 = env.reset(fixedInitialValues=[wp[:2] + np.array([0.03, -0.02]), 0.4, None])
        elif e == 3:     # next to the boundary
            obs0 = env.reset(applyNoise=False, fixedInitialValues=[np.array([1.97, -1.9]), 2.0, None])
        else:
            obs0 = env.reset()
        ep["iwp0"].append(env.iWp)
        ep["mults"].append([env.mMult, env.IMult, env.XuuMult, env.YvvMult, env.NrrMult, env.XuMult, env.YvMult, env.NrMult,
                            env.XactMult, env.YactMult, env.NactMult])
        ep["pos0"].append(np.array(env.position, dtype=float)); ep["heading0"].append(env.heading); ep["t_offset"].append(env.flowDataTimeOffset)
        ep["obs0"].append(obs0)
        arng = np.random.default_rng(700 + e)
        acts = np.zeros((n_steps, 3)); obs = np.full((n_steps, 11), np.nan); rew = np.full(n_steps, np.nan)
        done = np.zeros(n_steps, dtype=bool); hist = np.full((n_steps, 40), np.nan); iwp = np.full(n_steps, -1)
        o = obs0
        pd.oldObs = None
        for k in range(n_steps):
            a = pd.predict(o)[0] * 0.6 + 0.2 * arng.uniform(-1, 1, 3) if e != 3 else np.array([1.0, -1.0, 0.2])
            acts[k] = a
            o, r, d, _ = env.step(a)
            obs[k], rew[k], done[k], iwp[k] = o, r, d, env.iWp
            if d:
                hist[:k + 1] = env.timeHistory.values
                break
        else:
            hist[:] = np.array([list(row.values()) for row in env.timeHistory])
        ep["actions"].append(acts); ep["obs"].append(obs); ep["reward"].append(rew); ep["done"].append(done); ep["history"].append(hist); ep["iwp"].append(iwp)
        print("episode %d: %d steps, done %s, iWp %d -> %d, return %.3f" % (e, k + 1, np.nonzero(done)[0], ep["iwp0"][-1], env.iWp, np.nansum(rew)))
    for k, v in ep.items():
        out["ep_" + k] = np.array(v)
    out["flow_dt"] = np.array(env.flow.dt)
    path = os.path.join(HERE, "golden_legacy_cyl.npz")
    np.savez_compressed(path, **out)
    print("%s  %.1f KiB" % (path, os.path.getsize(path) / 1024.))


if __name__ == "__main__":
    main()
