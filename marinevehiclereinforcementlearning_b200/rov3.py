"""Batched BlueROV2 Heavy 3DoF engine (kernel K3): host-side constants with the
reference's attribute names (dynamicsModel_BlueROV2_Heavy_3DoF.py:39-112) and
the vectorised env (…_3DoF.py:375-514).  All numerics run in libmvrl."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import ACT_RPM, ACT_SETPOINT
from .rov6 import ACTION_MODES, _Handle, _RovVecEnv, _device_index


class Rov3Constants:
    def __init__(self):
        self.rho_f = 1000.
        self.m = 11.4
        self.dispVol = self.m / self.rho_f
        self.Length = 0.457
        self.Width = 0.338
        self.CB = np.zeros(3)
        self.CG = np.array([0., 0., 0.02])
        self.I = np.array([[0.16, 0., 0.], [0., 0.16, 0.], [0., 0., 0.16]])
        self.Xudot, self.Yvdot, self.Nrdot = -5.5, -12.7, -0.12
        self.Yrdot = self.Nvdot = 0.
        self.Xuu, self.Yvv, self.Nrr = -18.18, -21.66, -1.55
        self.Yrr = self.Ypp = self.Nvv = self.Npp = 0.
        self.Xu, self.Yv, self.Nr = -4.03, -6.22, -0.07
        self.Yr = self.Yp = self.Nv = self.Np = 0.
        self.D_thruster = 0.1
        self.alphaThruster = 45. / 180. * np.pi
        self.l_x, self.l_y = 0.156, 0.111
        self.Kt_thruster = 40. / (1000. * (3500. / 60.) ** 2. * self.D_thruster ** 4.)
        # allocation with Length/2 arms, 3DoF.py:104-112
        A = np.array([[1., 1., -1., -1.], [1., -1., 1., -1.], [1., 1., 1., 1.]])
        A[0, :] = A[0, :] * np.cos(self.alphaThruster)
        A[1, :] = A[1, :] * np.sin(self.alphaThruster)
        A[2, :] = A[2, :] * np.sin(self.alphaThruster) * self.Length / 2.
        self.A = A
        self.Ainv = np.linalg.pinv(A)
        self.rpmMax, self.rpmDeadband = 3500., 300.
        self.pid = {"windup": np.array([2., 2., 90. / 180. * np.pi]), "Kp": np.array([20., 20., 20.]),
                    "Ki": np.array([0.1, 0.1, 0.1]), "Kd": np.array([5., 5., 0.5]), "max": np.array([150., 150., 100.])}

    def massMatrix(self):
        """3DoF.py:198-206."""
        m, xg, yg = self.m, self.CG[0], self.CG[1]
        Mrb = np.array([[m, 0., -m * yg], [0., m, m * xg], [-m * yg, m * xg, self.I[2, 2]]])
        return Mrb + -1. * np.diag([self.Xudot, self.Yvdot, self.Nrdot])

    def to_struct(self):
        p = _lib.MvrlRov3Params()
        for name in ("rho_f", "m", "Length", "dispVol", "Xudot", "Yvdot", "Nrdot", "Xu", "Yv", "Yr", "Nv", "Nr",
                     "Xuu", "Yvv", "Yrr", "Nvv", "Nrr", "D_thruster", "alphaThruster", "l_x", "l_y"):
            setattr(p, name, float(getattr(self, name)))
        p.xg, p.yg, p.Izz = float(self.CG[0]), float(self.CG[1]), float(self.I[2, 2])
        p.thrust_coef = self.rho_f * self.D_thruster ** 4. * self.Kt_thruster
        p.rpm_max, p.rpm_deadband = self.rpmMax, self.rpmDeadband
        M = self.massMatrix()
        p.M[:] = list(M.reshape(-1))
        p.Minv[:] = list(np.linalg.inv(M).reshape(-1))
        p.Ainv[:] = list(np.asarray(self.Ainv, dtype=float).reshape(-1))
        p.pid_Kp[:] = list(self.pid["Kp"]); p.pid_Ki[:] = list(self.pid["Ki"]); p.pid_Kd[:] = list(self.pid["Kd"])
        p.pid_windup[:] = list(self.pid["windup"]); p.pid_max[:] = list(self.pid["max"])
        return p

    def __setattr__(self, name, value):
        object.__setattr__(self, name, value)
        object.__setattr__(self, "_version", getattr(self, "_version", 0) + 1)

    def touch(self):
        object.__setattr__(self, "_version", self._version + 1)

    def fingerprint(self):
        return (id(self), self._version)


class Rov3Handle(_Handle):
    PREFIX = "mvrl_rov3"


class BlueROV2Heavy3DoFVecEnv(_RovVecEnv):
    """Batched ``BlueROV2Heavy3DoFEnv`` (3DoF.py:375-514): 6 states, 5
    observations; "setpoint" actions (3 in [-1, 1], reference semantics) or
    "rpm" (4 thruster rpm FP AP FS AS, stateless)."""
    HANDLE = Rov3Handle
    CONSTANTS = Rov3Constants
    STATE_DIM, OBS_DIM, SP_DIM, PATH_DIM, CTRL_DIM, AUX_DIM = 6, 5, 3, 4, 7, 7
    ACTION_DIM = {ACT_RPM: 4, ACT_SETPOINT: 3}


class Rov3Derivs:
    """Batched ``BlueROV2Heavy3DoF.derivs`` (3DoF.py:128-296) through
    ``mvrl_rov3_derivs``.  state [6, N]; rpm mode: act [4, N]; set-point mode:
    t [N], setpoint [3, N], ctrl [7, N] (updated in place)."""

    def __init__(self, consts=None, dtype=torch.float64, action_mode="setpoint", device="cuda"):
        self.consts = consts if consts is not None else Rov3Constants()
        self.dtype, self.action_mode = dtype, ACTION_MODES[action_mode]
        self.device = torch.device("cuda", _device_index(device))
        self._handle, self._key = None, None

    def _get_handle(self):
        key = self.consts.fingerprint()
        if self._handle is None or key != self._key:
            self._handle = Rov3Handle(self.consts, self.dtype, self.action_mode, device=self.device.index)
            self._key = key
        return self._handle

    @staticmethod
    def new_ctrl(n, dtype=torch.float64, device="cuda"):
        c = torch.zeros((7, n), dtype=dtype, device=device)
        c[0] = float("nan")
        return c

    def __call__(self, state, act=None, t=None, setpoint=None, ctrl=None, want_aux=False):
        h = self._get_handle()
        prep = lambda x: None if x is None else x.to(device=self.device, dtype=self.dtype).contiguous()
        state, act, t, setpoint = prep(state), prep(act), prep(t), prep(setpoint)
        n = state.shape[1]
        dstate = torch.empty_like(state)
        aux = torch.empty((7, n), dtype=self.dtype, device=self.device) if want_aux else None
        _lib.check(h.fn("derivs")(h._h, n, n, _lib.ptr(state), _lib.ptr(act), _lib.ptr(t), _lib.ptr(setpoint),
                                  _lib.ptr(ctrl), _lib.ptr(dstate), _lib.ptr(aux), _lib.current_stream(self.device)))
        return (dstate, aux) if want_aux else dstate

    def thrusterModel(self, u, rpm):
        """(Fthruster, Xthruster), 3DoF.py:114-126; u, rpm: tensors [N]."""
        h = self._get_handle()
        u = u.to(device=self.device, dtype=self.dtype).contiguous()
        rpm = rpm.to(device=self.device, dtype=self.dtype).contiguous()
        F, X = torch.empty_like(u), torch.empty_like(u)
        _lib.check(h.fn("thruster_model")(h._h, u.numel(), _lib.ptr(u), _lib.ptr(rpm), _lib.ptr(F), _lib.ptr(X),
                                          _lib.current_stream(self.device)))
        return F, X
