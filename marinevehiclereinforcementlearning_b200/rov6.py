"""Batched BlueROV2 Heavy 6DoF engine: host-side owner of the device buffers
and of the libmvrl handle.  All numerics run in the CUDA kernels behind the C
ABI (include/mvrl.h); this module only allocates torch tensors, marshals
pointers and mirrors the reference's attribute names.

Reference surface mirrored: dynamicsModel_BlueROV2_Heavy_6DoF.py:75-218
(vehicle constants), :445-594 (env reset/step semantics).
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import ACT_FORCE, ACT_RPM, ACT_SETPOINT

ACTION_MODES = {"rpm": ACT_RPM, "force": ACT_FORCE, "setpoint": ACT_SETPOINT, "pid": ACT_SETPOINT,
                ACT_RPM: ACT_RPM, ACT_FORCE: ACT_FORCE, ACT_SETPOINT: ACT_SETPOINT}
ACTION_DIM = {ACT_RPM: 8, ACT_FORCE: 6, ACT_SETPOINT: 6}

# 6DoF.py:46-54
PID6_DEFAULTS = dict(
    windup=[2., 2., 2., 90. / 180. * np.pi, 90. / 180. * np.pi, 90. / 180. * np.pi],
    max=[50., 50., 50., 1., 1., 2.],
    Kp=[25., 25., 25., 10., 10., 1.],
    Ki=[2., 2., 2., 0.1, 0.1, 0.2],
    Kd=[20., 20., 20., 5., 5., 0.65])


class Rov6Constants:
    """Vehicle constants with the reference's attribute names
    (6DoF.py:83-218).  Pure host-side set-up, done once per vehicle exactly as
    the reference does in ``__init__`` (numpy ``pinv`` for the allocation so the
    matrices are bit-identical to the reference's)."""

    def __init__(self):
        self.rho_f = 1000.
        self.m = 11.4
        self.dispVol = self.m / self.rho_f
        self.Length = 0.457
        self.Width = 0.338
        self.CB = np.array([0., 0., 0.])
        self.CG = np.array([0., 0., 0.05])
        self.I = np.array([[0.16, 0., 0.], [0., 0.16, 0.], [0., 0., 0.16]])
        self.x_origin = np.array([0., 0., 0.0375])
        self.Xudot, self.Yvdot, self.Zwdot = -5.5, -12.7, -14.57
        self.Kpdot = self.Mqdot = self.Nrdot = -0.12
        self.Yrdot = self.Zvdot = self.Nvdot = 0.
        self.Xuu, self.Yvv, self.Zww = -18.18, -21.66, -36.99
        self.Kpp = self.Mqq = self.Nrr = -1.55
        self.Yrr = self.Ypp = self.Zqq = self.Kvv = self.Krr = 0.
        self.Mww = -1.55
        self.Nvv = self.Npp = 0.
        self.Xu, self.Yv, self.Zw = -4.03, -6.22, -5.18
        self.Kp = self.Mq = self.Nr = -0.07
        self.Yr = self.Yp = self.Zq = self.Kv = self.Kr = self.Mw = self.Nv = self.Np = 0.
        self.D_thruster = 0.1
        self.alphaThruster = 33. / 180. * np.pi
        self.l_x, self.l_y, self.l_z = 0.1475, 0.101, 0.068
        self.l_x_v, self.l_y_v, self.l_z_v = 0.120, 0.22, 0.0
        self.Kt_thruster = 40. / (1000. * (3500. / 60.) ** 2. * self.D_thruster ** 4.)
        a = self.alphaThruster
        self.thrusterPositions = np.array([
            [self.l_x, self.l_y, self.l_z], [self.l_x, -self.l_y, self.l_z],
            [-self.l_x, self.l_y, self.l_z], [-self.l_x, -self.l_y, self.l_z],
            [self.l_x_v, self.l_y_v, self.l_z_v], [self.l_x_v, -self.l_y_v, self.l_z_v],
            [-self.l_x_v, self.l_y_v, self.l_z_v], [-self.l_x_v, -self.l_y_v, self.l_z_v]])
        self.thrusterNormals = np.array([
            [np.cos(a), -np.sin(a), 0.], [np.cos(a), np.sin(a), 0.],
            [-np.cos(a), -np.sin(a), 0.], [-np.cos(a), np.sin(a), 0.],
            [0., 0., -1.], [0., 0., 1.], [0., 0., 1.], [0., 0., -1.]])
        from .resources import computeThrustAllocation
        self.A, self.Ainv = computeThrustAllocation(self.thrusterPositions, self.thrusterNormals)
        self.disableThrusters = False
        self.rpmMax = 3500.
        self.rpmDeadband = 300.
        self.pid = {k: np.array(v, dtype=float) for k, v in PID6_DEFAULTS.items()}

    def massMatrix(self):
        """Mrb + Ma as assembled at 6DoF.py:286-299 (Ma[2,2] = -Zvdot)."""
        m, (xg, yg, zg) = self.m, self.CG
        M = np.array([
            [m, 0., 0., 0., m * zg, -m * yg],
            [0., m, 0., -m * zg, 0., m * xg],
            [0., 0., m, m * yg, -m * xg, 0.],
            [0., -m * zg, m * yg, 0., 0., 0.],
            [m * zg, 0., -m * xg, 0., 0., 0.],
            [-m * yg, m * xg, 0., 0., 0., 0.]])
        M[3:, 3:] = self.I
        return M + -1. * np.diag([self.Xudot, self.Yvdot, self.Zvdot, self.Kpdot, self.Mqdot, self.Nrdot])

    def to_struct(self):
        p = _lib.MvrlRov6Params()
        p.rho_f, p.m, p.Length = self.rho_f, self.m, self.Length
        p.CG[:] = list(map(float, self.CG))
        p.CB[:] = list(map(float, self.CB))
        p.I[:] = list(map(float, np.asarray(self.I).reshape(-1)))
        for name in ("Xudot", "Yvdot", "Zwdot", "Kpdot", "Mqdot", "Nrdot",
                     "Xu", "Yv", "Yp", "Yr", "Zw", "Zq", "Kv", "Kp", "Kr", "Mw", "Mq", "Nv", "Np", "Nr",
                     "Xuu", "Yvv", "Ypp", "Yrr", "Zww", "Zqq", "Kvv", "Kpp", "Krr", "Mww", "Mqq", "Nvv", "Npp", "Nrr"):
            setattr(p, name, float(getattr(self, name)))
        p.W = self.m * 9.81                                   # 6DoF.py:374
        p.B = self.dispVol * self.rho_f * 9.81                # 6DoF.py:375
        p.thrust_coef = self.rho_f * self.D_thruster ** 4. * self.Kt_thruster
        p.rpm_max, p.rpm_deadband = self.rpmMax, self.rpmDeadband
        M = self.massMatrix()
        p.M[:] = list(M.reshape(-1))
        p.Minv[:] = list(np.linalg.inv(M).reshape(-1))
        p.A[:] = list(np.asarray(self.A, dtype=float).reshape(-1))
        p.Ainv[:] = list(np.asarray(self.Ainv, dtype=float).reshape(-1))
        p.pid_Kp[:] = list(self.pid["Kp"]); p.pid_Ki[:] = list(self.pid["Ki"]); p.pid_Kd[:] = list(self.pid["Kd"])
        p.pid_windup[:] = list(self.pid["windup"]); p.pid_max[:] = list(self.pid["max"])
        p.disable_thrusters = 1 if self.disableThrusters else 0
        return p

    def __setattr__(self, name, value):
        object.__setattr__(self, name, value)
        object.__setattr__(self, "_version", getattr(self, "_version", 0) + 1)

    def touch(self):
        """Call after modifying an array attribute in place (e.g. ``rov.CG[2] = 0.1``)
        so that handles built from these constants are refreshed."""
        object.__setattr__(self, "_version", self._version + 1)

    def fingerprint(self):
        return (id(self), self._version)


class _Handle:
    """RAII wrapper of an opaque libmvrl handle (MvrlRov6* / MvrlRov3*)."""
    PREFIX = "mvrl_rov6"

    def __init__(self, consts, dtype, action_mode, n_sub=8, dt=0.2, max_steps=250, seed=0, env_id0=0,
                 auto_reset=False, fixed_sp=False, device=0, fast_math=False):
        _lib.require_cuda()
        self.lib = _lib.load()
        cfg = _lib.MvrlRov6Config(
            dtype=_lib.torch_dtype_code(dtype), action_mode=ACTION_MODES[action_mode], n_sub=int(n_sub),
            max_steps=int(max_steps), dt=float(dt), seed=int(seed) & (2 ** 64 - 1), env_id0=int(env_id0),
            auto_reset=int(bool(auto_reset)), fixed_sp=int(bool(fixed_sp)), device=int(device), fast_math=int(bool(fast_math)))
        self.cfg = cfg
        self.params = consts.to_struct()
        self._h = C.c_void_p()
        _lib.check(getattr(self.lib, self.PREFIX + "_create")(C.byref(self._h), C.byref(self.params), C.byref(cfg)))

    def fn(self, name):
        return getattr(self.lib, "%s_%s" % (self.PREFIX, name))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                getattr(self.lib, self.PREFIX + "_destroy")(h)
            except Exception:
                pass
            self._h = None


class Rov6Handle(_Handle):
    PREFIX = "mvrl_rov6"

    @property
    def specialised(self):
        return bool(self.lib.mvrl_rov6_is_specialised(self._h))

    @property
    def specialisation(self):
        """0 = generic kernels, 1 = default-sparsity kernels, 2 = default vehicle: fp32 kernels with compile-time constants."""
        return int(self.lib.mvrl_rov6_is_specialised(self._h))


def _device_index(device):
    _lib.require_cuda()
    d = torch.device(device)
    if d.type != "cuda":
        raise RuntimeError("the batched simulator runs on CUDA devices only (got %s); there is no CPU fallback" % d)
    return d.index if d.index is not None else torch.cuda.current_device()


class _RovVecEnv:
    """N vehicle environments stepped by one fused CUDA kernel (shared host logic
    of the 6DoF and 3DoF batched envs).

    Semantics per environment are those of the reference's Gym env with the
    integrator fixed to classic RK4 x ``n_sub``; the batch follows the SB3
    ``VecEnv`` convention (``step`` -> obs, rewards, dones, infos; auto-reset
    with ``terminal_observation``).  Tensors stay on the device.  Internally
    every array is structure-of-arrays ``[field, ld]``; ``obs`` / ``actions``
    are exposed as ``[N, k]`` views of them (``*_fm`` gives the feature-major
    tensors for transposition-free policies).
    """
    HANDLE = Rov6Handle
    CONSTANTS = None
    STATE_DIM, OBS_DIM, SP_DIM, PATH_DIM, CTRL_DIM, AUX_DIM = 12, 9, 6, 6, 13, 14
    ACTION_DIM = {ACT_RPM: 8, ACT_FORCE: 6, ACT_SETPOINT: 6}

    def __init__(self, num_envs, seed=0, dt=0.2, maxSteps=250, n_sub=8, action_mode="setpoint",
                 dtype=torch.float32, device="cuda", auto_reset=True, env_id0=0, fast_math=False,
                 record_aux=False, record_terminal_obs=True, collect_stats=True, vehicle=None):
        self.num_envs = int(num_envs)
        self.dt, self._max_episode_steps, self.n_sub = float(dt), int(maxSteps), int(n_sub)
        self.action_mode = ACTION_MODES[action_mode]
        if self.action_mode not in self.ACTION_DIM:
            raise ValueError("action_mode %r is not supported by %s" % (action_mode, type(self).__name__))
        self.dtype, self.seed, self.env_id0 = dtype, int(seed), int(env_id0)
        self.auto_reset, self.fast_math = bool(auto_reset), bool(fast_math)
        self.device = torch.device("cuda", _device_index(device))
        self.vehicle = vehicle if vehicle is not None else self.CONSTANTS()
        self.lenAction = self.ACTION_DIM[self.action_mode]
        self.lenObs = self.OBS_DIM
        self.fixedSp = False
        n = self.num_envs
        self.ld = max(32, ((n + 31) // 32) * 32)    # an empty batch still owns (non-null) buffers
        ld, dev = self.ld, self.device
        z = lambda k: torch.zeros((k, ld), dtype=dtype, device=dev)
        self._state, self._action, self._obs = z(self.STATE_DIM), z(self.lenAction), z(self.OBS_DIM)
        self._reward = torch.zeros(ld, dtype=dtype, device=dev)
        self._done = torch.zeros(ld, dtype=torch.uint8, device=dev)
        self._istep = torch.zeros(ld, dtype=torch.int32, device=dev)
        self._setpoint, self._path = z(self.SP_DIM), z(self.PATH_DIM)
        self._ctrl = z(self.CTRL_DIM)
        self._episode = torch.zeros(ld, dtype=torch.int32, device=dev)  # reinterpreted as uint32
        self._terminal_obs = z(self.OBS_DIM) if (record_terminal_obs and auto_reset) else None
        self._aux = z(self.AUX_DIM) if record_aux else None
        self._stats = torch.zeros(8, dtype=torch.float64, device=dev) if collect_stats else None
        if self._stats is not None:
            self._reset_stats()
        self._handle = None
        self._handle_key = None
        self._bufs = _lib.MvrlRov6Buffers(
            state=self._state.data_ptr(), action=self._action.data_ptr(), obs=self._obs.data_ptr(),
            reward=self._reward.data_ptr(), done=self._done.data_ptr(), istep=self._istep.data_ptr(),
            setpoint=self._setpoint.data_ptr(), path=self._path.data_ptr(), ctrl=self._ctrl.data_ptr(),
            episode=self._episode.data_ptr(),
            terminal_obs=None if self._terminal_obs is None else self._terminal_obs.data_ptr(),
            aux=None if self._aux is None else self._aux.data_ptr(),
            ep_stats=None if self._stats is None else self._stats.data_ptr())
        self._needs_episode_bump = False

    # -- handle management (re-created when a constant or a flag changes) ----
    def _get_handle(self):
        key = (self.vehicle.fingerprint(), self.fixedSp, self.auto_reset, self.n_sub, self.dt,
               self._max_episode_steps, self.seed, self.env_id0, self.fast_math, self.action_mode)
        if self._handle is None or key != self._handle_key:
            self._handle = self.HANDLE(self.vehicle, self.dtype, self.action_mode, n_sub=self.n_sub, dt=self.dt,
                                       max_steps=self._max_episode_steps, seed=self.seed, env_id0=self.env_id0,
                                       auto_reset=self.auto_reset, fixed_sp=self.fixedSp, device=self.device.index,
                                       fast_math=self.fast_math)
            self._handle_key = key
        return self._handle

    def _reset_stats(self):
        self._stats.zero_()
        self._stats[3] = float("inf")
        self._stats[4] = float("-inf")

    # -- views -----------------------------------------------------------------
    @property
    def obs_fm(self):
        return self._obs[:, :self.num_envs]

    @property
    def actions_fm(self):
        return self._action[:, :self.num_envs]

    @property
    def systemState(self):
        return self._state[:, :self.num_envs].T

    @property
    def state(self):
        return self.obs_fm.T

    @property
    def setPoint(self):
        return self._setpoint[:, :self.num_envs].T

    @property
    def path(self):
        """[N, 2, dims] way-points (6DoF.py:497, 509 / 3DoF.py:423, 435)."""
        return self._path[:, :self.num_envs].T.reshape(self.num_envs, 2, self.PATH_DIM // 2)

    @property
    def iStep(self):
        return self._istep[:self.num_envs]

    @property
    def time(self):
        return self._istep[:self.num_envs].to(torch.float64) * self.dt

    # -- reset / step ------------------------------------------------------------
    def reset(self, initialSetpoint=None, mask=None):
        """6DoF.py:485-529 / 3DoF.py:411-453.  ``initialSetpoint=None`` takes the
        random branch (for 6DoF defined here; the reference's own line raises -
        see DESIGN.md)."""
        self.fixedSp = initialSetpoint is not None
        h = self._get_handle()
        if self._needs_episode_bump:
            if mask is None:
                self._episode += 1
            else:
                self._episode[:self.num_envs] += torch.as_tensor(mask).to(device=self.device, dtype=torch.int32)
        self._needs_episode_bump = True
        sp = None
        if initialSetpoint is not None:
            vals = [float(v) for v in np.asarray(initialSetpoint, dtype=float).reshape(self.SP_DIM)]
            sp = (C.c_double * self.SP_DIM)(*vals)
        m = None
        if mask is not None:
            m = torch.as_tensor(mask).to(device=self.device, dtype=torch.uint8).contiguous()
        _lib.check(h.fn("reset")(h._h, self.num_envs, self.ld, C.byref(self._bufs), _lib.ptr(m), sp,
                                 _lib.current_stream(self.device)))
        return self.state

    def set_actions(self, actions):
        """Copy ``actions`` ([N, A] or feature-major [A, N]) into the SoA action buffer."""
        if actions.data_ptr() == self._action.data_ptr():
            return
        a = actions.to(device=self.device, dtype=self.dtype, non_blocking=True)
        if a.dim() == 2 and a.shape == (self.num_envs, self.lenAction):
            self.actions_fm.copy_(a.T)
        elif a.dim() == 2 and a.shape == (self.lenAction, self.num_envs):
            self.actions_fm.copy_(a)
        else:
            raise ValueError("actions must be [%d, %d] or [%d, %d], got %s" %
                             (self.num_envs, self.lenAction, self.lenAction, self.num_envs, tuple(a.shape)))

    def step_async(self, actions=None):
        """Launch the fused step kernel on the current stream (no sync)."""
        if actions is not None:
            self.set_actions(actions)
        h = self._get_handle()
        _lib.check(h.fn("step")(h._h, self.num_envs, self.ld, C.byref(self._bufs), _lib.current_stream(self.device)))

    def step(self, actions=None):
        self.step_async(actions)
        n = self.num_envs
        infos = {}
        if self._terminal_obs is not None:
            infos["terminal_observation"] = self._terminal_obs[:, :n].T
        return self.state, self._reward[:n], self._done[:n].bool(), infos

    def step_host(self, actions, obs_out=None, reward_out=None, done_out=None, chunks=0):
        """One env step for a caller that lives on the HOST: ``actions`` is a CPU tensor
        ``[N, A]`` (pinned memory lets the copies overlap), results land in CPU tensors
        ``obs [N, obs]``, ``reward [N]``, ``done [N]`` (uint8), allocated pinned on first use.
        Runs ``mvrl_rov6_step_host``: upload, SoA transpose, fused step, transpose back and
        download are pipelined over ``chunks`` equal pieces of the batch (0: the library's
        default).  Returns when the host tensors are complete."""
        if self.HANDLE.PREFIX != "mvrl_rov6":
            raise NotImplementedError("step_host / step_range are implemented for the 6DoF env")
        n = self.num_envs
        if actions.device.type != "cpu" or actions.dtype != self.dtype or tuple(actions.shape) != (n, self.lenAction) or not actions.is_contiguous():
            raise ValueError("actions must be a contiguous CPU tensor [%d, %d] of %s" % (n, self.lenAction, self.dtype))
        if obs_out is None:
            if getattr(self, "_h_obs", None) is None:
                self._h_obs = torch.empty((n, self.lenObs), dtype=self.dtype).pin_memory()
                self._h_reward = torch.zeros(n, dtype=self.dtype).pin_memory()   # reward == 0 (6DoF.py:575): never shipped over PCIe
                self._h_done = torch.empty(n, dtype=torch.uint8).pin_memory()
            obs_out, reward_out, done_out = self._h_obs, self._h_reward, self._h_done
        elif reward_out is not None:
            reward_out.zero_()   # caller-owned array: the library zero-fills it only when it captures the pipeline
        h = self._get_handle()
        self._bufs.action = self._action.data_ptr()
        _lib.check(h.lib.mvrl_rov6_step_host(h._h, n, self.ld, C.byref(self._bufs), C.c_void_p(actions.data_ptr()),
                                             C.c_void_p(obs_out.data_ptr()),
                                             None if reward_out is None else C.c_void_p(reward_out.data_ptr()),
                                             None if done_out is None else C.c_void_p(done_out.data_ptr()),
                                             int(chunks), _lib.current_stream(self.device)))
        return obs_out, reward_out, done_out

    def step_range_async(self, first, count):
        """Launch the fused step for environments ``[first, first + count)`` only (``mvrl_rov6_step_range``)."""
        if self.HANDLE.PREFIX != "mvrl_rov6":
            raise NotImplementedError("step_host / step_range are implemented for the 6DoF env")
        h = self._get_handle()
        _lib.check(h.lib.mvrl_rov6_step_range(h._h, int(first), int(count), self.ld, C.byref(self._bufs), _lib.current_stream(self.device)))

    def observe(self):
        """dataToState for the current device state (6DoF.py:467-483 / 3DoF.py:397-409),
        [N, obs]: resources.angleError / clip run on the device through torch views of the
        same SoA buffers the kernels use."""
        from . import resources
        n, L3 = self.num_envs, self.vehicle.Length * 3.
        half = self.PATH_DIM // 2
        pos = self._state[:half, :n]
        ang_sp = self._setpoint[half:, :n]
        ang = self._state[half:self.SP_DIM, :n]
        err = torch.stack([resources.angleError(ang_sp[k].contiguous(), ang[k].contiguous()) for k in range(ang.shape[0])])
        obs = torch.cat([(self._path[:half, :n] - pos) / L3, (self._path[half:, :n] - pos) / L3, err / (45. / 180. * np.pi)])
        return obs.clamp(-1., 1.).T

    # -- episode statistics (K5): device-side accumulators + optional all-reduce
    def episode_stats(self, reduce_group=None, reset=True):
        """{episodes, mean_length, mean_return, min_return, max_return,
        nonfinite}.  When torch.distributed is initialised the 8 accumulators
        are all-reduced over the process group (NCCL on GPUs) - the only
        collective of the framework, off the step path.  Episodes are counted
        where they END INSIDE THE KERNEL, i.e. with ``auto_reset=True``; an env
        created with ``auto_reset=False`` keeps integrating past ``done`` like
        the reference (6DoF.py:589-592) and reports 0 episodes here."""
        if self._stats is None:
            raise RuntimeError("collect_stats=False")
        from .distributed import reduce_episode_stats
        out = reduce_episode_stats(self._stats, group=reduce_group)
        if reset:
            self._reset_stats()
        return out

    # -- checkpoint / resume -------------------------------------------------------
    _STATE_KEYS = ("_state", "_istep", "_setpoint", "_path", "_ctrl", "_episode", "_obs")

    def state_dict(self):
        d = {k: getattr(self, k).clone() for k in self._STATE_KEYS}
        d["fixedSp"] = self.fixedSp
        return d

    def load_state_dict(self, d):
        for k, v in d.items():
            if k == "fixedSp":
                self.fixedSp = bool(v)
            else:
                getattr(self, k).copy_(v)
        self._needs_episode_bump = True


class BlueROV2Heavy6DoFVecEnv(_RovVecEnv):
    """Batched ``BlueROV2Heavy6DoFEnv`` (6DoF.py:445-594): 12 states, 9
    observations; actions per ``action_mode``: "setpoint" (reference Gym
    semantics, 6 in [-1, 1]), "force" (6 earth-frame generalised forces) or
    "rpm" (8 thruster rpm)."""
    HANDLE = Rov6Handle
    CONSTANTS = Rov6Constants


class Rov6Derivs:
    """Batched ``BlueROV2Heavy6DoF.derivs`` (6DoF.py:406-442): one derivative
    evaluation per environment through ``mvrl_rov6_derivs`` (kernel K2).

    Feature-major tensors: ``state`` [12, N]; ``act`` [8, N] rpm or [6, N]
    earth-frame forces; set-point mode takes ``t`` [N], ``setpoint`` [6, N] and
    the controller state ``ctrl`` [13, N] (updated in place like the reference
    mutates its controller)."""

    def __init__(self, consts=None, dtype=torch.float64, action_mode="rpm", device="cuda"):
        self.consts = consts if consts is not None else Rov6Constants()
        self.dtype, self.action_mode = dtype, ACTION_MODES[action_mode]
        self.device = torch.device("cuda", _device_index(device))
        self._handle, self._key = None, None

    def _get_handle(self):
        key = self.consts.fingerprint()
        if self._handle is None or key != self._key:
            self._handle = Rov6Handle(self.consts, self.dtype, self.action_mode, device=self.device.index)
            self._key = key
        return self._handle

    @staticmethod
    def new_ctrl(n, dtype=torch.float64, device="cuda"):
        """Fresh PID state (6DoF.py:37-41): eOld=None (NaN marker), eInt=0, tOld=0."""
        c = torch.zeros((13, n), dtype=dtype, device=device)
        c[0] = float("nan")
        return c

    def __call__(self, state, act=None, t=None, setpoint=None, ctrl=None, want_aux=False):
        h = self._get_handle()
        prep = lambda x: None if x is None else x.to(device=self.device, dtype=self.dtype).contiguous()
        state, act, t, setpoint = prep(state), prep(act), prep(t), prep(setpoint)
        if ctrl is not None and (ctrl.dtype != self.dtype or not ctrl.is_contiguous() or ctrl.device != self.device):
            raise ValueError("ctrl must be a contiguous [13, N] tensor of the model dtype on the model device")
        n = state.shape[1]
        dstate = torch.empty_like(state)
        aux = torch.empty((50, n), dtype=self.dtype, device=self.device) if want_aux else None
        _lib.check(h.lib.mvrl_rov6_derivs(h._h, n, n, _lib.ptr(state), _lib.ptr(act), _lib.ptr(t), _lib.ptr(setpoint),
                                          _lib.ptr(ctrl), _lib.ptr(dstate), _lib.ptr(aux), _lib.current_stream(self.device)))
        return (dstate, aux) if want_aux else dstate
