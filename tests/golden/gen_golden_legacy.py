"""Generate golden vectors for the LEGACY path by executing the unmodified
reference (tag_00_Dec2023_simpleControlTurbulence/) in this container.

    python tests/golden/gen_golden_legacy.py

Reference entry points exercised (file:line, legacy/ = tag_00_...):
* ReconstructedFlow.scale / interp        legacy/flowGenerator.py:53-136
* AuvEnv.reset / step / dataToState       legacy/verySimpleAuv.py:147-410
* headingError                            legacy/resources.py:26-46

The SPOD blobs coeffs.npy / modes_r.npy are absent from the reference
checkout (.MISSING_LARGE_BLOBS), so ReconstructedFlow.__init__ cannot run.
Its reconstruction step (modes @ coeffs + mean) is replaced by a synthetic
field = ltm.npy mean + seeded noise installed on an instance created without
__init__; scale(), interp() and all of AuvEnv are the reference's own code.
Runs in its own interpreter because legacy/resources.py and /resources.py
clash by module name.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_shims import LEGACY_ROOT, import_legacy  # noqa: E402

ref_flow, ref_auv, ref_lres = import_legacy()

NT = 48  # time levels of the synthetic field kept small for the fixture


def synthetic_base_field(ltm, nt, seed=7, sigma=0.05):
    rng = np.random.default_rng(seed)
    return ltm[None] + sigma * rng.standard_normal((nt,) + ltm.shape)


class SyntheticFlow(ref_flow.ReconstructedFlow):
    """The reference class with only the data-loading constructor replaced."""

    def __init__(self, dataDir=None):
        d = os.path.join(LEGACY_ROOT, "turbulenceData")
        self.lt_mean = np.load(os.path.join(d, "ltm.npy"))
        self.baseFlowData = synthetic_base_field(self.lt_mean, NT)
        self.baseDt = 0.002  # params_coeffs.yaml: time_step
        self.baseTime = np.array([i * self.baseDt for i in range(NT)])
        self.baseCoords = np.load(os.path.join(d, "turbulence_coords.npy"))
        self.baseDx = (self.baseCoords[0, 1:, 0] - self.baseCoords[0, :-1, 0])[0]
        self.baseDy = (self.baseCoords[1:, 0, 1] - self.baseCoords[:-1, 0, 1])[0]
        self.scale(1., 1., 1.)


def main():
    out = {}
    flow = SyntheticFlow()
    out["ltm"] = flow.lt_mean
    out["nt"] = np.array(NT)
    out["base_dx"] = np.array(flow.baseDx); out["base_dy"] = np.array(flow.baseDy); out["base_dt"] = np.array(flow.baseDt)
    out["base_field_sample"] = flow.baseFlowData[::7, ::5, ::6, :].copy()  # spot check of the regenerated field

    # --- scale + interp, legacy/flowGenerator.py:53-136 (scaling as legacy/verySimpleAuv.py:104)
    flow.scale(11., 1.0, 2.0, translate=(-1.65, -1.1))
    out["scaled_dx"] = np.array(flow.dx); out["scaled_dy"] = np.array(flow.dy); out["scaled_dt"] = np.array(flow.dt)
    out["scaled_field_sample"] = flow.flowData[::7, ::5, ::6, :].copy()
    rng = np.random.default_rng(11)
    nq = 400
    tq = rng.uniform(-0.05, flow.time[-1] * 1.1, nq)
    xq = rng.uniform(-1.2, 3.6, nq)   # negative x/y extrapolate from cell 0 (translate is ignored by interp)
    yq = rng.uniform(-1.2, 2.5, nq)
    tq[:4] = [0.0, flow.time[-1], flow.dt * 3, flow.dt * 3 + 1e-12]
    xq[:4] = [0.0, flow.dx * 60, flow.dx * 5, -0.5]
    yq[:4] = [0.0, flow.dy * 40, flow.dy * 7, -0.25]
    out["interp_t"] = tq; out["interp_xy"] = np.stack([xq, yq], axis=1)
    out["interp_res"] = np.array([flow.interp(tq[i], (xq[i], yq[i])) for i in range(nq)])
    out["interp_field_t"] = np.array([0.0137, flow.dt * 10.25])
    out["interp_field"] = np.array([flow.interpField(t) for t in out["interp_field_t"]])

    # --- headingError, legacy/resources.py:26-46
    pairs = rng.uniform(-8, 8, (64, 2)); pairs[0] = [1., 1.]; pairs[1] = [0., np.pi]
    out["heading_pairs"] = pairs
    out["heading_err"] = np.array([ref_lres.headingError(a, b) for a, b in pairs])

    # --- AuvEnv episodes, legacy/verySimpleAuv.py:216-410
    ref_flow.ReconstructedFlow = SyntheticFlow  # AuvEnv.__init__ constructs flowGenerator.ReconstructedFlow("./turbulenceData")
    n_ep, n_steps = 6, 60
    ep = {k: [] for k in ("mults", "pos0", "heading0", "heading_target", "t_offset", "actions", "obs0", "obs", "reward", "done", "history")}
    for e in range(n_ep):
        np.random.seed(100 + e)  # the reference draws from the global numpy RNG
        env = ref_auv.AuvEnv(noiseMagCoeffs=0.1, noiseMagActuation=0.1, currentVelScale=1.0, currentTurbScale=2.0,
                             stopOnBoundsExceeded=(e != 1))
        env._max_episode_steps = 50 if e == 2 else 250
        if e == 3:
            obs0 = env.reset(applyNoise=False, fixedInitialValues=[np.array([0.3, -0.2]), 1.0, 4.0])
        elif e == 4:  # starts next to the boundary: bounds termination (legacy/verySimpleAuv.py:335-342)
            obs0 = env.reset(fixedInitialValues=[np.array([0.97, -0.98]), 0.3, 2.0])
        else:
            obs0 = env.reset()
        ep["mults"].append([env.mMult, env.IMult, env.XuuMult, env.YvvMult, env.NrrMult, env.XuMult, env.YvMult, env.NrMult,
                            env.XactMult, env.YactMult, env.NactMult])
        ep["pos0"].append(np.array(env.position, dtype=float)); ep["heading0"].append(env.heading)
        ep["heading_target"].append(env.headingTarget); ep["t_offset"].append(env.flowDataTimeOffset)
        ep["obs0"].append(obs0)
        arng = np.random.default_rng(500 + e)
        acts = arng.uniform(-1, 1, (n_steps, 3))
        if e == 4:
            acts[:, 0] = 1.0; acts[:, 1] = -1.0
        obs = np.full((n_steps, 11), np.nan); rew = np.full(n_steps, np.nan); done = np.zeros(n_steps, dtype=bool)
        hist = np.full((n_steps, 40), np.nan)
        for k in range(n_steps):
            o, r, d, _ = env.step(acts[k])
            obs[k], rew[k], done[k] = o, r, d
            if d:
                h = env.timeHistory.values  # DataFrame on done, legacy/verySimpleAuv.py:402-403
                hist[:k + 1] = h
                break
        else:
            hist[:] = np.array([list(row.values()) for row in env.timeHistory])
        ep["actions"].append(acts); ep["obs"].append(obs); ep["reward"].append(rew); ep["done"].append(done); ep["history"].append(hist)
        print("episode %d: %d steps, done at %s, return %.3f" % (e, k + 1, np.nonzero(done)[0], np.nansum(rew)))
    for k, v in ep.items():
        out["ep_" + k] = np.array(v)
    out["flow_time_quarter"] = np.array(env.flow.time[env.flow.time.shape[0] // 4])
    path = os.path.join(HERE, "golden_legacy.npz")
    np.savez_compressed(path, **out)
    print("%s  %.1f KiB" % (path, os.path.getsize(path) / 1024.))


if __name__ == "__main__":
    main()
