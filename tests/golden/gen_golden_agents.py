"""Golden vectors for the action producers that sit in front of the env step, produced by executing the
unmodified reference in this container:

    python tests/golden/gen_golden_agents.py

* lineOfSight, LOSNavigation.predict      dynamicsModel_BlueROV2_Heavy_3DoF.py:517-607
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_shims import import_current  # noqa: E402

ref_res, ref6, ref3 = import_current()


def main():
    rng = np.random.default_rng(2024)
    n = 3000
    p0 = rng.uniform(-1, 1, (n, 2))
    p1 = rng.uniform(-1, 1, (n, 2))
    # edge cases: second way-point inside the LOS circle, horizontal / vertical segments, segment far away,
    # vehicle beyond either end of the segment, tangent-ish configurations
    p1[:50] *= 0.3
    p0[50:80, 1] = p1[50:80, 1]                      # horizontal segments (sign(pathVec[1]) == 0 branch)
    p0[80:110, 0] = p1[80:110, 0]                    # vertical segments
    p0[110:160] = p0[110:160] * 0.1 + np.array([0.9, 0.9]); p1[110:160] = p1[110:160] * 0.1 + np.array([0.9, -0.9])
    p0[160:200] += 3.0; p1[160:200] += 4.0           # both way-points far away (delta < 0)
    p0[200:230] = np.array([0.8, 0.0]) + 0.01 * rng.standard_normal((30, 2)); p1[200:230] = np.array([0.8, 0.9])
    rnav = np.full(n, 0.5)
    rnav[230:400] = rng.uniform(0.1, 1.2, 170)
    tgt = np.array([ref3.lineOfSight(p0[i].copy(), p1[i].copy(), rnav[i]) for i in range(n)])
    obs = np.concatenate([p0[:500], p1[:500], rng.uniform(-1, 1, (500, 1))], axis=1)
    nav = ref3.LOSNavigation()
    act = np.array([nav.predict(obs[i])[0] for i in range(500)])
    path = os.path.join(HERE, "golden_agents.npz")
    np.savez_compressed(path, los_p0=p0, los_p1=p1, los_rnav=rnav, los_target=tgt, nav_obs=obs, nav_action=act)
    print(path, "%.1f KiB" % (os.path.getsize(path) / 1024.))


if __name__ == "__main__":
    main()
