"""Drop-in for tag_00_Dec2023_simpleControlTurbulence/flowGenerator.py:
``ReconstructedFlow`` with ``scale`` / ``interp`` / ``interpField`` running as
CUDA kernels on a device-resident field (see ``auv.FlowField``)."""
import os

import numpy as np
import torch

from .. import _lib
from ..auv import FlowField, _device_index


def synthetic_base_field(lt_mean, nt, seed=7, sigma=0.05, kind="noise"):
    """Stand-in for the SPOD reconstruction when ``coeffs.npy`` / ``modes_r.npy`` are
    unavailable (they are absent from the reference checkout): long-time mean + seeded
    fluctuations, ``[nt, Ny, Nx, 3]``.  ``kind="noise"``: white Gaussian noise of standard
    deviation ``sigma``.  ``kind="modes"``: a handful of long-wave
    travelling modes of amplitude ``sigma`` - like real turbulence data (and unlike white noise)
    it extrapolates gently for x, y < 0, where ``interp`` reads because it ignores ``translate``."""
    rng = np.random.default_rng(seed)
    lt_mean = np.asarray(lt_mean, dtype=float)
    if kind == "noise":
        return lt_mean[None] + sigma * rng.standard_normal((nt,) + lt_mean.shape)
    if kind != "modes":
        raise ValueError("kind must be 'noise' or 'modes'")
    ny, nx, nf = lt_mean.shape
    t = np.arange(nt, dtype=np.float32)[:, None, None]
    y = np.arange(ny, dtype=np.float32)[None, :, None]
    x = np.arange(nx, dtype=np.float32)[None, None, :]
    f = np.repeat(lt_mean[None].astype(np.float32), nt, axis=0)
    for c in range(nf):
        for _ in range(4):
            kt, ky, kx = rng.uniform(0.05, 0.3), rng.uniform(0.02, 0.12), rng.uniform(0.02, 0.12)
            f[..., c] += (0.4 * sigma) * np.sin(kt * t + ky * y + kx * x + rng.uniform(0, 2 * np.pi))
    return f


class ReconstructedFlow(FlowField):
    """flowGenerator.py:13-159.  ``ReconstructedFlow(dataDir)`` reads the same
    files as the reference (``coeffs.npy``, ``modes_r.npy``, ``ltm.npy``,
    ``params_coeffs.yaml``, ``turbulence_coords.npy``) and raises
    ``FileNotFoundError`` like the reference when one is missing;
    ``ReconstructedFlow.from_base_field`` / ``.synthetic`` build the object from
    an explicit base field instead."""

    def __init__(self, dataDir, dtype=torch.float32, device="cuda"):
        import yaml
        coeffs = np.load(os.path.join(dataDir, "coeffs.npy"))
        modes = np.load(os.path.join(dataDir, "modes_r.npy"))
        self.lt_mean = np.load(os.path.join(dataDir, "ltm.npy"))
        base = self.reconstruct(modes, coeffs, self.lt_mean, dtype=dtype, device=device)
        with open(os.path.join(dataDir, "params_coeffs.yaml"), "r") as infile:
            params = yaml.safe_load(infile)
        coords = np.load(os.path.join(dataDir, "turbulence_coords.npy"))
        dx, dy = self._check_spacing(coords)
        super().__init__(base, baseDx=dx, baseDy=dy, baseDt=params["time_step"], baseCoords=coords, dtype=dtype, device=device)

    @staticmethod
    def reconstruct(modes, coeffs, lt_mean, dtype=torch.float32, device="cuda"):
        """flowGenerator.py:20-23: ``baseFlowData[t] = Re(modes @ coeffs[:, t]) + lt_mean`` for every time level at once, by
        kernel ``mvrl_flow_reconstruct`` (fp64 accumulation, real part / mean / conversion to ``dtype`` fused into the
        epilogue, output already in the ``[Nt, Ny, Nx, 3]`` layout).  ``modes [Ny, Nx, 3, K]`` and ``coeffs [K, Nt]`` real or
        complex (the pySPOD blobs are complex128)."""
        lib = _lib.load()
        dev = torch.device("cuda", _device_index(device))
        modes, coeffs = np.asarray(modes), np.asarray(coeffs)
        if modes.ndim < 2 or coeffs.ndim != 2 or modes.shape[-1] != coeffs.shape[0]:
            raise ValueError("modes [..., K] and coeffs [K, Nt] do not match: %s, %s" % (modes.shape, coeffs.shape))
        if tuple(np.shape(lt_mean)) != tuple(modes.shape[:-1]):
            raise ValueError("lt_mean %s does not match the modes' plane %s" % (np.shape(lt_mean), modes.shape[:-1]))
        k, nt = coeffs.shape
        plane = int(np.prod(modes.shape[:-1]))

        def as_f64(a):   # complex128 -> interleaved (re, im) doubles, anything real -> float64
            cplx = np.iscomplexobj(a)
            a = np.ascontiguousarray(a, dtype=np.complex128 if cplx else np.float64)
            return torch.as_tensor(a.view(np.float64), device=dev), int(cplx)

        m, m_c = as_f64(modes.reshape(plane, k))
        c, c_c = as_f64(coeffs)
        mean = torch.as_tensor(np.ascontiguousarray(lt_mean, dtype=np.float64).reshape(-1), device=dev)
        out = torch.empty((nt,) + tuple(modes.shape[:-1]), dtype=dtype, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.mvrl_flow_reconstruct(_lib.torch_dtype_code(dtype), plane, k, nt, _lib.ptr(m), m_c, _lib.ptr(c), c_c,
                                                 _lib.ptr(mean), _lib.ptr(out), _lib.current_stream(dev)))
            torch.cuda.current_stream(dev).synchronize()   # m, c, mean die with this frame
        return out

    @staticmethod
    def _check_spacing(coords):
        """flowGenerator.py:32-42: the grid must be uniform in x and in y ((y, x) orientation)."""
        dx = coords[0, 1:, 0] - coords[0, :-1, 0]
        dy = coords[1:, 0, 1] - coords[:-1, 0, 1]
        if not np.all(np.abs(dx - dx[0]) < 1e-6):
            raise ValueError("Non-uniform input grid spacing in the x-direction")
        if not np.all(np.abs(dy - dy[0]) < 1e-6):
            raise ValueError("Non-uniform input grid spacing in the y-direction")
        return float(dx[0]), float(dy[0])

    @classmethod
    def from_base_field(cls, baseFlowData, baseDx=0.005, baseDy=0.005, baseDt=0.002, baseCoords=None, lt_mean=None,
                        dtype=torch.float32, device="cuda"):
        self = cls.__new__(cls)
        self.lt_mean = lt_mean
        FlowField.__init__(self, baseFlowData, baseDx=baseDx, baseDy=baseDy, baseDt=baseDt, baseCoords=baseCoords, dtype=dtype, device=device)
        return self

    @classmethod
    def synthetic(cls, dataDir=None, lt_mean=None, nt=2000, seed=7, sigma=0.05, kind="noise", dtype=torch.float32, device="cuda"):
        """Mean field (``ltm.npy`` of ``dataDir`` or ``lt_mean``) + seeded noise, ``nt`` time levels."""
        coords, dx, dy, dt = None, 0.005, 0.005, 0.002
        if lt_mean is None:
            lt_mean = np.load(os.path.join(dataDir, "ltm.npy"))
            cpath = os.path.join(dataDir, "turbulence_coords.npy")
            if os.path.exists(cpath):
                coords = np.load(cpath)
                dx, dy = cls._check_spacing(coords)
        return cls.from_base_field(synthetic_base_field(lt_mean, nt, seed, sigma, kind), baseDx=dx, baseDy=dy, baseDt=dt,
                                   baseCoords=coords, lt_mean=np.asarray(lt_mean), dtype=dtype, device=device)

    # turbulence intensity on the plane, flowGenerator.py:47-51 (computed lazily: the env never needs it)
    def _intensity(self):
        f = self.baseFlowData  # == flowData after scale(1, 1, 1), where the reference evaluates it
        self.uPrime = torch.sqrt(torch.sum((f[..., 0] - 1.) ** 2., dim=0) / f.shape[0]).cpu().numpy()
        self.vPrime = torch.sqrt(torch.sum((f[..., 1] - 0.) ** 2., dim=0) / f.shape[0]).cpu().numpy()
        self.TI = np.sqrt(0.5 * (self.uPrime + self.vPrime))
        self.baseTI = self.TI[self.TI.shape[0] // 2, self.TI.shape[1] // 2]

    def __getattr__(self, name):
        if name in ("uPrime", "vPrime", "TI", "baseTI"):
            self._intensity()
            return self.__dict__[name]
        raise AttributeError(name)
