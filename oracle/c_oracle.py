"""ctypes front-end of the C oracle (oracle/mvrl_oracle.c).  TEST
INFRASTRUCTURE ONLY - see the header of mvrl_oracle.c."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import oracle_np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmvrl_oracle.so")
_d = C.c_double


class OrcRov6Params(C.Structure):
    _fields_ = [(n, _d) for n in ("rho_f", "m", "Length", "dispVol")] + [("CG", _d * 3), ("CB", _d * 3), ("I", _d * 9)] + \
        [(n, _d) for n in ("Xudot", "Yvdot", "Zwdot", "Kpdot", "Mqdot", "Nrdot", "Zvdot",
                           "Xu", "Yv", "Yp", "Yr", "Zw", "Zq", "Kv", "Kp", "Kr", "Mw", "Mq", "Nv", "Np", "Nr",
                           "Xuu", "Yvv", "Ypp", "Yrr", "Zww", "Zqq", "Kvv", "Kpp", "Krr", "Mww", "Mqq", "Nvv", "Npp", "Nrr",
                           "D_thruster", "Kt_thruster")] + [("A", _d * 48), ("Ainv", _d * 48)]


class OrcPid6(C.Structure):
    _fields_ = [("eOld", _d * 6), ("eInt", _d * 6), ("tOld", _d), ("margin", _d), ("has_old", C.c_int)]


# numpy view of an (OrcPid6 * n) array: lets the tests read / overwrite the controller state of all environments at once
PID6_DTYPE = np.dtype([("eOld", "f8", 6), ("eInt", "f8", 6), ("tOld", "f8"), ("margin", "f8"), ("has_old", "i4")], align=True)
assert PID6_DTYPE.itemsize == C.sizeof(OrcPid6)


class OrcRov6Env(C.Structure):
    _fields_ = [("mode", C.c_int), ("n_sub", C.c_int), ("max_steps", C.c_int), ("auto_reset", C.c_int),
                ("fixed_sp", C.c_int), ("threads", C.c_int), ("dt", _d), ("seed", C.c_uint64), ("env_id0", C.c_uint64)]


_lib = None


def load():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "mvrl_oracle.c")
        if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
            subprocess.run(["make", "-C", HERE], check=True, capture_output=True)
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")  # idle OpenMP workers must not spin against torch threads
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_angle_error.restype = _d
        _lib.orc_angle_error.argtypes = [_d, _d]
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def params_struct(p=None):
    p = p or oracle_np.Rov6Params()
    s = OrcRov6Params()
    for name, _ in OrcRov6Params._fields_:
        v = getattr(p, name)
        if name in ("CG", "CB"):
            getattr(s, name)[:] = [float(x) for x in v]
        elif name in ("I", "A", "Ainv"):
            getattr(s, name)[:] = [float(x) for x in np.asarray(v).reshape(-1)]
        else:
            setattr(s, name, float(v))
    return s


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def derivs6(state, act=None, mode=0, t=None, sp=None, ctrl=None, params=None, want_aux=False):
    lib = load()
    state = np.ascontiguousarray(state, dtype=np.float64).reshape(-1, 12)
    n = state.shape[0]
    out = np.empty((n, 12))
    gcf, cv = np.zeros((n, 6)), np.zeros((n, 8))
    act = None if act is None else np.ascontiguousarray(act, dtype=np.float64)
    t = None if t is None else np.ascontiguousarray(np.broadcast_to(t, (n,)), dtype=np.float64)
    sp = None if sp is None else np.ascontiguousarray(np.broadcast_to(sp, (n, 6)), dtype=np.float64)
    ps = params_struct(params)
    lib.orc_rov6_derivs(C.byref(ps), C.c_int(mode), C.c_long(n), _p(state), _p(act), _p(t), _p(sp),
                        None if ctrl is None else C.byref(ctrl), _p(out), _p(gcf), _p(cv))
    return (out, gcf, cv) if want_aux else out


class Rov6EnvC:
    """Same contract as oracle_np.Rov6EnvOracle, backed by orc_rov6_step."""

    def __init__(self, n, params=None, dt=0.2, max_steps=250, n_sub=8, mode=oracle_np.MODE_PID, seed=0,
                 auto_reset=False, env_id0=0, threads=0):
        self.lib = load()
        self.n, self.p = n, (params or oracle_np.Rov6Params())
        self.ps = params_struct(self.p)
        self.cfg = OrcRov6Env(mode=mode, n_sub=n_sub, max_steps=max_steps, auto_reset=int(auto_reset), fixed_sp=0,
                              threads=threads, dt=dt, seed=seed, env_id0=env_id0)
        self.mode = mode
        self._np = oracle_np.Rov6EnvOracle(n, params=self.p, dt=dt, max_steps=max_steps, n_sub=n_sub, mode=mode, seed=seed,
                                           auto_reset=auto_reset, env_id0=env_id0)

    def reset(self, initial_setpoint=None):
        obs = self._np.reset(initial_setpoint)  # reset draws/state set-up shared with the numpy oracle
        n = self.n
        self.cfg.fixed_sp = int(self._np.fixed_sp)
        self.state = np.zeros((n, 12))
        self.set_point = np.ascontiguousarray(self._np.set_point)
        self.path = np.ascontiguousarray(self._np.path)
        self.ctrl = (OrcPid6 * n)()
        self.ctrl_np = np.frombuffer(self.ctrl, dtype=PID6_DTYPE)   # shares memory with self.ctrl
        self.ctrl_np["margin"] = np.inf
        self.i_step = np.zeros(n, dtype=np.int32)
        self.time = np.zeros(n)
        self.episode = np.zeros(n, dtype=np.uint32)
        self.obs = np.ascontiguousarray(obs)
        self.done = np.zeros(n, dtype=np.uint8)
        self.term_obs = np.zeros((n, 9))
        self.aux = np.zeros((n, 14))
        self.mincos = np.ones(n)  # running min |cos(theta)| over all RK4 stages (conditioning diagnostic)
        self.dbmargin = np.full(n, np.inf)  # running min relative distance of a thruster demand from the dead-band edge
        return self.obs.copy()

    def step(self, action):
        action = np.ascontiguousarray(action, dtype=np.float64)
        self.lib.orc_rov6_step(C.byref(self.ps), C.byref(self.cfg), C.c_long(self.n), _p(self.state), _p(action),
                               _p(self.set_point), _p(self.path), self.ctrl, _p(self.i_step), _p(self.time), _p(self.episode),
                               _p(self.obs), _p(self.done), _p(self.term_obs), _p(self.aux), _p(self.mincos), _p(self.dbmargin))
        return self.obs, np.zeros(self.n), self.done.astype(bool), {"terminal_observation": self.term_obs}
