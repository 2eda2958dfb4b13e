"""Import shims that let the UNMODIFIED reference run in this container.

Used only by the ``gen_golden_*.py`` scripts (which need ``/root/reference``
and are therefore run here, never on the GPU box).  The reference imports
``matplotlib``, ``mpl_toolkits`` and ``gym`` at module scope
(dynamicsModel_BlueROV2_Heavy_6DoF.py:7-16, resources.py:8-17); none of them
is installed, and none of them is used by the numerical path, so inert stub
modules are enough.
"""
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"
LEGACY_ROOT = REFERENCE_ROOT + "/tag_00_Dec2023_simpleControlTurbulence"


class _Anything:
    """Object that swallows any attribute access / call (plot stubs)."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter(())


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__getattr__ = lambda attr: _Anything()  # PEP 562
    sys.modules[name] = m
    return m


def install_stubs():
    mpl = _module("matplotlib", rc=lambda *a, **k: None, rcParams={})
    for sub in ("pyplot", "animation", "widgets", "cm", "patches"):
        setattr(mpl, sub, _module("matplotlib." + sub))
    tk = _module("mpl_toolkits")
    tk.mplot3d = _module("mpl_toolkits.mplot3d", Axes3D=_Anything)

    class Env:  # gym.Env stand-in: the reference only subclasses it
        def __init__(self, *a, **k):
            pass

    class Box:
        def __init__(self, low=None, high=None, shape=None, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    def np_random(seed=None):
        return np.random.RandomState(seed), seed

    seeding = _module("gym.utils.seeding", np_random=np_random)
    utils = _module("gym.utils", seeding=seeding)
    spaces = _module("gym.spaces", Box=Box)
    _module("gym", Env=Env, spaces=spaces, utils=utils)


def import_current():
    """Top-level (current) generation: 3DoF/6DoF models + resources.py."""
    install_stubs()
    sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import resources as ref_res
        import dynamicsModel_BlueROV2_Heavy_6DoF as ref6
        import dynamicsModel_BlueROV2_Heavy_3DoF as ref3
    return ref_res, ref6, ref3


def import_legacy():
    """Legacy generation; must run in its own interpreter (module-name clash
    between /resources.py and legacy/resources.py)."""
    install_stubs()
    sys.path.insert(0, LEGACY_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import flowGenerator as ref_flow
        import verySimpleAuv as ref_auv
        import resources as ref_lres
    return ref_flow, ref_auv, ref_lres
