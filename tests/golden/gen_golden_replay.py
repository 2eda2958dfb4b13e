"""Golden vectors for the symmetry-augmenting replay buffer: the UNMODIFIED reference class ``CustomReplayBuffer``
(tag_00.../main_02_sbl_contrib_customBuffer.py:57-160) executed in this container.

    python tests/golden/gen_golden_replay.py

The module imports stable_baselines3 / sb3_contrib / gymnasium at the top (:20-44); none of them is installed and the
reference pins no versions (no requirements file).  Everything ``CustomReplayBuffer.add`` itself does is the reference's
own code; what it inherits is the storage set up by ``ReplayBuffer.__init__`` of stable_baselines3 (2.x: the reference
imports ``gymnasium.spaces``; ``common/buffers.py``), restated in ``_Sb3ReplayBuffer`` below:
``buffer_size = max(buffer_size // n_envs, 1)`` slots; ``observations`` / ``next_observations``
``[buffer_size, n_envs, *obs_shape]`` in the observation space's dtype; ``actions [buffer_size, n_envs, action_dim]``;
``rewards`` / ``dones`` / ``timeouts [buffer_size, n_envs]`` float32; ``pos = 0``, ``full = False``.  The rest of the
third-party surface is inert stubs.  The module's training script sits under ``if __name__ == "__main__"`` and does not run.
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_shims  # noqa: E402
from _ref_shims import LEGACY_ROOT, _Anything, _module  # noqa: E402


class _Box:
    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)


class _Discrete:
    pass


class _Sb3BaseBuffer:
    pass


class _Sb3ReplayBuffer(_Sb3BaseBuffer):
    """stable_baselines3.common.buffers.ReplayBuffer.__init__, optimize_memory_usage=False (see module docstring)."""

    def __init__(self, buffer_size, observation_space, action_space, device="auto", n_envs=1, optimize_memory_usage=False,
                 handle_timeout_termination=True):
        self.buffer_size = max(buffer_size // n_envs, 1)
        self.observation_space, self.action_space = observation_space, action_space
        self.obs_shape, self.action_dim = observation_space.shape, int(np.prod(action_space.shape))
        self.pos, self.full, self.device, self.n_envs = 0, False, device, n_envs
        self.optimize_memory_usage = optimize_memory_usage
        self.observations = np.zeros((self.buffer_size, n_envs, *self.obs_shape), dtype=observation_space.dtype)
        self.next_observations = np.zeros((self.buffer_size, n_envs, *self.obs_shape), dtype=observation_space.dtype)
        self.actions = np.zeros((self.buffer_size, n_envs, self.action_dim), dtype=action_space.dtype)
        self.rewards = np.zeros((self.buffer_size, n_envs), dtype=np.float32)
        self.dones = np.zeros((self.buffer_size, n_envs), dtype=np.float32)
        self.handle_timeout_termination = handle_timeout_termination
        self.timeouts = np.zeros((self.buffer_size, n_envs), dtype=np.float32)


def import_main_02():
    import pandas, torch, yaml  # noqa: F401  real packages first: their own imports must not see the stub modules
    _ref_shims.install_stubs()
    sb3 = _module("stable_baselines3")
    common = _module("stable_baselines3.common")
    sb3.common = common
    for sub, attrs in (("vec_env", {}), ("noise", {}), ("preprocessing", {}), ("type_aliases", {}), ("utils", {}),
                       ("buffers", {"BaseBuffer": _Sb3BaseBuffer, "ReplayBuffer": _Sb3ReplayBuffer})):
        setattr(common, sub, _module("stable_baselines3.common." + sub, **attrs))
    _module("sb3_contrib")
    gymn = _module("gymnasium")
    gymn.spaces = _module("gymnasium.spaces", Box=_Box, Discrete=_Discrete, Space=object)
    sys.path.insert(0, LEGACY_ROOT)
    cwd = os.getcwd()
    os.chdir(LEGACY_ROOT)   # the module's own imports (verySimpleAuv, resources) are relative to its directory
    try:
        spec = importlib.util.spec_from_file_location("ref_main_02", os.path.join(LEGACY_ROOT, "main_02_sbl_contrib_customBuffer.py"))
        mod = importlib.util.module_from_spec(spec)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
    return mod


def main():
    mod = import_main_02()
    n_envs, slots, n_add = 5, 23, 34
    buf = mod.CustomReplayBuffer(slots * n_envs + 3, _Box(-1, 1, (11,)), _Box(-1, 1, (3,)), device="cpu", n_envs=n_envs)
    assert buf.buffer_size == slots
    rng = np.random.default_rng(0)
    out = {"n_envs": np.array(n_envs), "buffer_size_arg": np.array(slots * n_envs + 3), "slots": np.array(slots)}
    keys = ("obs", "next_obs", "act", "rew", "done", "timeout")
    ins = {k: [] for k in keys}
    trace = []
    for k in range(n_add):
        obs = rng.uniform(-1, 1, (n_envs, 11)).astype(np.float32)
        nxt = rng.uniform(-1, 1, (n_envs, 11)).astype(np.float32)
        act = rng.uniform(-1, 1, (n_envs, 3)).astype(np.float32)
        rew = rng.uniform(-3, 3, n_envs).astype(np.float32)
        done = rng.uniform(size=n_envs) < 0.2
        tout = done & (rng.uniform(size=n_envs) < 0.5)
        infos = [({"TimeLimit.truncated": True} if t else {}) for t in tout]
        buf.add(obs, nxt, act, rew, done, infos)          # the reference's own add(), unmodified
        for key, v in zip(keys, (obs, nxt, act, rew, done, tout)):
            ins[key].append(v)
        trace.append((buf.pos, int(buf.full), buf.nRollovers))
    for key in keys:
        out["in_" + key] = np.array(ins[key])
    out["trace"] = np.array(trace)
    for name in ("observations", "next_observations", "actions", "rewards", "dones", "timeouts"):
        out["buf_" + name] = getattr(buf, name)
    assert buf.nRollovers > 2   # both regimes: with the four mirror images and without
    np.savez_compressed(os.path.join(HERE, "golden_replay.npz"), **out)
    print("wrote golden_replay.npz; final (pos, full, nRollovers) =", trace[-1])


if __name__ == "__main__":
    main()
