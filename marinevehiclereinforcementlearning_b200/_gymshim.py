"""``gym.Env`` / ``gym.spaces.Box`` when gym or gymnasium is installed, tiny
stand-ins otherwise (the reference only subclasses Env and declares Box
spaces: dynamicsModel_BlueROV2_Heavy_6DoF.py:445-465)."""
import numpy as np

try:  # pragma: no cover - neither package is in the build image
    import gymnasium as _gym
except Exception:  # noqa: BLE001
    try:
        import gym as _gym
    except Exception:  # noqa: BLE001
        _gym = None

if _gym is not None:
    Env = _gym.Env
    Box = _gym.spaces.Box
else:
    class Env:
        metadata = {}

        def render(self, mode="human"):
            pass

        def close(self):
            pass

    class Box:
        def __init__(self, low=-1.0, high=1.0, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            self.shape = tuple(shape) if shape is not None else np.shape(low)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

        def sample(self):
            return np.random.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))
