"""Rollout actor next to the env: the SB3 ``MlpPolicy`` of the reference's training scripts
(tag_00_Dec2023_simpleControlTurbulence/main_00_sbl.py:100-105: ``net_arch=[128, 128, 128]``, GELU) with a Gaussian action
head, evaluated by ONE tensor-core kernel (``mvrl_policy_act``, csrc/mvrl_policy.cu) directly on the batched env's
structure-of-arrays observation / action buffers - so that collecting a rollout step is two launches: actor, env step.

The parameters live in ordinary fp32 torch tensors (a learner may update them in place); ``sync_weights()`` repacks them
into the kernel's bf16 operand layout (canonical K-major core matrices for tcgen05.mma).  The learner side (losses, optimiser) is out of scope, like in SURVEY.md 8(e)."""
import ctypes as C
import math

import torch

from . import _lib
from .rov6 import _device_index

HIDDEN = 128


class MlpGaussianPolicy:
    """obs_dim -> 128 -> 128 -> 128 -> act_dim, GELU (tanh form), tanh-squashed mean, state-independent ``log_std``.

    ``act(env)`` / ``act_into(obs_fm, act_fm, ...)`` sample ``a = clip(mean + std * eps, -1, 1)`` with ``eps`` from Philox
    keyed on (seed, global env id, step): the draw of an environment does not depend on the batch layout or the number
    of GPUs.  ``predict(obs, deterministic)`` is the SB3-shaped entry point (numpy in / out) the reference's
    ``evaluate_agent`` loop calls (resources.py:145-198)."""

    def __init__(self, obs_dim, act_dim, device="cuda", seed=0, log_std_init=-0.5):
        self.device = torch.device("cuda", _device_index(device))
        self.obs_dim, self.act_dim, self.seed = int(obs_dim), int(act_dim), int(seed)
        g = torch.Generator().manual_seed(self.seed)
        dims = [self.obs_dim, HIDDEN, HIDDEN, HIDDEN, self.act_dim]
        self.weights = [torch.randn(dims[i + 1], dims[i], generator=g) / math.sqrt(dims[i]) for i in range(4)]   # nn.Linear layout [out, in]
        self.biases = [torch.zeros(dims[i + 1]) for i in range(4)]
        self.log_std = torch.full((self.act_dim,), float(log_std_init))
        self.step = 0
        self.lib = _lib.load()
        _lib.require_cuda()
        self._h = C.c_void_p()
        _lib.check(self.lib.mvrl_policy_create(C.byref(self._h), self.device.index, self.obs_dim, self.act_dim))
        self.sync_weights()

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self.lib.mvrl_policy_destroy(h)
            except Exception:
                pass
            self._h = None

    def sync_weights(self):
        """Repack the fp32 parameters (any device) into the kernel's bf16 operand layout.  Call after a learner update."""
        host = [t.detach().to("cpu", torch.float32).contiguous() for t in
                (self.weights[0], self.biases[0], self.weights[1], self.biases[1], self.weights[2], self.biases[2],
                 self.weights[3], self.biases[3], self.log_std)]
        _lib.check(self.lib.mvrl_policy_set_weights(self._h, *[C.c_void_p(t.data_ptr()) for t in host]))

    @property
    def logp_const(self):
        return float(-self.log_std.sum())

    def act_into(self, obs_fm, act_fm, n, logp=None, mean=None, eps=None, env_id0=0, step=None, deterministic=False):
        """obs_fm float32 [obs_dim, ld] -> act_fm float32 [act_dim, ld] (feature-major, as the envs hold them); no sync."""
        if step is None:
            step, self.step = self.step, self.step + 1
        ld = obs_fm.shape[1]
        for t, rows in ((obs_fm, self.obs_dim), (act_fm, self.act_dim), (mean, self.act_dim), (eps, self.act_dim)):
            if t is not None and (t.dtype != torch.float32 or t.device != self.device or t.shape[0] < rows or t.shape[1] != ld or not t.is_contiguous()):
                raise ValueError("feature-major float32 tensors [k, %d] on %s expected" % (ld, self.device))
        _lib.check(self.lib.mvrl_policy_act(self._h, int(n), int(ld), _lib.ptr(obs_fm), _lib.ptr(act_fm), _lib.ptr(logp), _lib.ptr(mean),
                                            _lib.ptr(eps), self.seed & (2 ** 64 - 1), int(env_id0), int(step) & 0xffffffff, int(bool(deterministic)),
                                            _lib.current_stream(self.device)))

    def act(self, env, logp=None, deterministic=False, step=None):
        """Sample actions for a batched env straight into its action buffer (the next ``env.step_async()`` reads them)."""
        self.act_into(env._obs, env._action, env.num_envs, logp=logp, env_id0=env.env_id0, step=step, deterministic=deterministic)
        return env.actions_fm

    def predict(self, obs, state=None, episode_start=None, deterministic=False):
        """SB3 ``predict``: numpy / tensor observations [obs_dim] or [N, obs_dim] -> (actions, None)."""
        x = torch.as_tensor(obs, dtype=torch.float32)
        single = x.dim() == 1
        x = x.reshape(-1, self.obs_dim).to(self.device)
        n = x.shape[0]
        ld = max(32, (n + 31) // 32 * 32)
        o = torch.zeros((self.obs_dim, ld), dtype=torch.float32, device=self.device)
        o[:, :n] = x.T
        a = torch.zeros((self.act_dim, ld), dtype=torch.float32, device=self.device)
        self.act_into(o, a, n, deterministic=deterministic)
        out = a[:, :n].T.cpu().numpy()
        return (out[0] if single else out), None
