"""Drop-in for tag_00_Dec2023_simpleControlTurbulence/verySimpleAuv_cyl.py: ``AuvEnvCyl`` (one vehicle following
way-points around a cylinder; numpy in / numpy out, old-Gym 4-tuple) plus ``AuvCylVecEnv`` for batches.  The env
step runs in the ``auv_step`` CUDA kernel (variant MVRL_AUV_CYL)."""
import numpy as np
import torch

from ..auv import AuvCylVecEnv, cyl_waypoints  # noqa: F401
from . import flowGenerator
from .resources import headingError  # noqa: F401
from .verySimpleAuv import HISTORY_COLUMNS, AuvEnv, PDController  # noqa: F401


class AuvEnvCyl(AuvEnv):
    """verySimpleAuv_cyl.py:22-345.  Same constructor signature as upstream (+ ``flow=`` / ``device=``)."""
    _VEC = AuvCylVecEnv

    def __init__(self, seed=None, dt=0.02, noiseMagCoeffs=0.0, noiseMagActuation=0.0,
                 currentVelScale=1.0, currentTurbScale=2.0, stopOnBoundsExceeded=True, flow=None, device="cuda"):
        super().__init__(seed=seed, dt=dt, noiseMagCoeffs=noiseMagCoeffs, noiseMagActuation=noiseMagActuation,
                         currentVelScale=currentVelScale, currentTurbScale=currentTurbScale,
                         stopOnBoundsExceeded=stopOnBoundsExceeded, flow=flow, device=device)
        self.Rcyl = 1.33
        self.xCyl = np.array([2.5, 0.])
        self.waypoints, self.wpThreshold = cyl_waypoints(self.Rcyl, self.xCyl)
        self.iWp = 0
        self._max_episode_steps = 1200
        self.xMinMax = [-2, 2]
        self.yMinMax = [-2, 2]

    def _sync(self):
        super()._sync()
        self.iWp = int(self._vec.iWp[0])
        self.positionTarget = self.waypoints[self.iWp, :2]
        self.headingTarget = float(self.waypoints[self.iWp, 2])

    def dataToState(self, pos, heading, velocities):
        """verySimpleAuv_cyl.py:84-115 (V0 scaling) for the given pose, against the stored previous errors."""
        perr = self.positionTarget - np.asarray(pos, dtype=float)
        herr = headingError(self.headingTarget, heading)
        herr_o, perr_o = (herr, perr) if self.herr_o is None else (self.herr_o, self.perr_o)
        c = lambda x: min(1., max(-1., x))
        return np.concatenate([[c(perr[0] / 0.2), c(perr[1] / 0.2), c(herr / (45. / 180. * np.pi)), c((herr - herr_o) / (2. / 180 * np.pi)),
                                c((perr[0] - perr_o[0]) / 0.025), c((perr[1] - perr_o[1]) / 0.025)],
                               np.clip(np.asarray(velocities) / [0.2, 0.2, 30. / 180. * np.pi], -1., 1.), np.zeros(2)])
