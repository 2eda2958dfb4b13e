"""Drop-in for the numerical helpers of the reference's ``resources.py``
(``computeThrustAllocation`` :19-35, ``angleError`` :75-95,
``coordinateTransform`` :98-143) and its evaluation loop (``evaluate_agent`` :145-198).  The plotting /
SB3 training glue of that file is out of scope.

``angleError`` and ``coordinateTransform`` run as CUDA kernels
(``mvrl_angle_error`` / ``mvrl_coordinate_transform``): scalars in -> numpy
out like the reference; torch tensors ``[N]`` in -> tensors ``[N]`` /
``[N, dof, dof]`` out, on the device.  ``computeThrustAllocation`` is one-time
host set-up (numpy ``pinv``), exactly as in the reference.
"""
import numpy as np
import torch

from . import _lib


def computeThrustAllocation(thrusterPositions, thrusterNormals, x0=None):
    """A[:, i] = [n_i ; (r_i - x0) x n_i] and its pseudo-inverse."""
    pos = np.asarray(thrusterPositions, dtype=float)
    nrm = np.asarray(thrusterNormals, dtype=float)
    if x0 is None:
        x0 = np.zeros(3)
    A = np.vstack([nrm.T, np.cross(pos - np.asarray(x0, dtype=float), nrm).T])
    return A, np.linalg.pinv(A)


def _to_device(values):
    """-> (list of contiguous 1-D CUDA tensors of one dtype, was_scalar, dtype)."""
    _lib.require_cuda()
    tensors = [v for v in values if isinstance(v, torch.Tensor)]
    if tensors:
        dev = next((t.device for t in tensors if t.is_cuda), torch.device("cuda"))
        dtype = torch.float32 if all(t.dtype == torch.float32 for t in tensors) else torch.float64
        n = max(t.numel() for t in tensors)
        out = [torch.as_tensor(v, dtype=dtype, device=dev).reshape(-1).expand(n).contiguous() if not isinstance(v, torch.Tensor)
               else v.to(device=dev, dtype=dtype).reshape(-1).expand(n).contiguous() for v in values]
        return out, False, dtype
    arrs = [np.asarray(v, dtype=np.float64) for v in values]
    scalar = all(a.ndim == 0 for a in arrs)
    n = max(a.size for a in arrs)
    out = [torch.as_tensor(np.broadcast_to(a.reshape(-1), (n,)).copy(), device="cuda") for a in arrs]
    return out, scalar, torch.float64


def angleError(psi_d, psi):
    """Signed heading error with Python-modulo wrap (resources.py:75-95)."""
    (a, b), scalar, dtype = _to_device([psi_d, psi])
    out = torch.empty_like(a)
    lib = _lib.load()
    _lib.check(lib.mvrl_angle_error(_lib.torch_dtype_code(dtype), a.numel(), _lib.ptr(a), _lib.ptr(b), _lib.ptr(out),
                                    _lib.current_stream(a.device)))
    if isinstance(psi_d, torch.Tensor) or isinstance(psi, torch.Tensor):
        return out
    res = out.cpu().numpy()
    return float(res[0]) if scalar else res


headingError = angleError  # legacy name, tag_00.../resources.py:26-46


def coordinateTransform(phi, theta, psi, dof=["x", "y", "psi"]):
    """J(phi, theta, psi) for the active degrees of freedom (resources.py:98-143)."""
    if isinstance(dof, int):
        ndof = dof
    elif set(dof) == {"x", "y", "psi"}:
        ndof = 3
    elif set(dof) == {"x", "y", "z", "phi", "theta", "psi"}:
        ndof = 6
    else:
        raise ValueError("unsupported dof set %r" % (dof,))
    if ndof not in (3, 6):
        raise ValueError("dof must be 3 or 6")
    (p, t, s), scalar, dtype = _to_device([phi, theta, psi])
    n = p.numel()
    out = torch.empty((ndof * ndof, n), dtype=dtype, device=p.device)
    lib = _lib.load()
    _lib.check(lib.mvrl_coordinate_transform(_lib.torch_dtype_code(dtype), ndof, n, n, _lib.ptr(p), _lib.ptr(t), _lib.ptr(s),
                                             _lib.ptr(out), _lib.current_stream(p.device)))
    J = out.T.reshape(n, ndof, ndof)
    if any(isinstance(v, torch.Tensor) for v in (phi, theta, psi)):
        return J
    J = J.cpu().numpy()
    return J[0] if scalar else J


def evaluate_agent(agent, env, num_episodes=1, num_steps=None, deterministic=True, num_last_for_reward=None,
                   render=False, init=None, saveDir=None):
    """resources.py:145-198 (see ``vec_tools.evaluate_agent``)."""
    from .vec_tools import evaluate_agent as _impl
    return _impl(agent, env, num_episodes=num_episodes, num_steps=num_steps, deterministic=deterministic,
                 num_last_for_reward=num_last_for_reward, render=render, init=init, saveDir=saveDir)
