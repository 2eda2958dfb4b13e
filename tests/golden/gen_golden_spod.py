"""Golden vectors for the SPOD reconstruction: the UNMODIFIED reference constructor
``ReconstructedFlow.__init__`` (tag_00.../flowGenerator.py:14-51) executed in this container on a directory of
synthetic blobs (``_spod_blobs.py``; the real ``coeffs.npy`` / ``modes_r.npy`` are absent from the checkout), with the
reference's own ``ltm.npy``, ``turbulence_coords.npy`` and ``time_step``.

    python tests/golden/gen_golden_spod.py

Recorded: a strided sample and per-time-level plane sums of ``baseFlowData``, ``uPrime / vPrime / TI / baseTI``,
spacings, and - after the env's ``scale(11, 1, 2, translate=(-1.65, -1.1))`` (verySimpleAuv.py:104) - ``interp`` at
200 points and one ``interpField`` plane.  Two cases: complex128 blobs (what pySPOD writes) and real ones.
"""
import os
import sys
import tempfile

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_shims import LEGACY_ROOT, import_legacy  # noqa: E402
from _spod_blobs import write_spod_dir  # noqa: E402

ref_flow, _, _ = import_legacy()

N_MODES, NT, SEED = 37, 29, 21   # deliberately not multiples of the kernel's 16 / 64 tiles


def main():
    d = os.path.join(LEGACY_ROOT, "turbulenceData")
    ltm = np.load(os.path.join(d, "ltm.npy"))
    coords = np.load(os.path.join(d, "turbulence_coords.npy"))
    with open(os.path.join(d, "params_coeffs.yaml")) as f:
        time_step = yaml.safe_load(f)["time_step"]
    out = {"ltm": ltm, "coords": coords, "time_step": np.array(time_step), "n_modes": np.array(N_MODES), "nt": np.array(NT), "seed": np.array(SEED)}
    rng = np.random.default_rng(5)
    for tag, cplx in (("c", True), ("r", False)):
        with tempfile.TemporaryDirectory() as tmp:
            write_spod_dir(tmp, ltm, coords, time_step, N_MODES, NT, SEED, cplx)
            flow = ref_flow.ReconstructedFlow(tmp)          # the reference's own constructor, unmodified
        base = flow.baseFlowData
        out[tag + "_base_sample"] = base[::2, ::3, ::4, :].copy()
        out[tag + "_base_plane_sum"] = base.sum(axis=(1, 2))              # [nt, 3]
        out[tag + "_base_plane_abs_sum"] = np.abs(base).sum(axis=(1, 2))
        out[tag + "_base_last"] = base[-1].copy()
        for k in ("baseDt", "baseDx", "baseDy", "baseTI"):
            out[tag + "_" + k] = np.array(getattr(flow, k))
        for k in ("uPrime", "vPrime", "TI"):
            out[tag + "_" + k] = getattr(flow, k)
        flow.scale(11., 1., 2., translate=(-1.65, -1.1))                   # verySimpleAuv.py:104
        t = rng.uniform(0., flow.time[-1], 200)
        xy = np.stack([rng.uniform(0., 61 * flow.dx, 200), rng.uniform(0., 41 * flow.dy, 200)], axis=1)
        out[tag + "_interp_t"], out[tag + "_interp_xy"] = t, xy
        out[tag + "_interp_res"] = np.array([flow.interp(ti, p) for ti, p in zip(t, xy)])
        out[tag + "_interp_field_t"] = np.array(0.37 * flow.time[-1])
        out[tag + "_interp_field"] = flow.interpField(0.37 * flow.time[-1])
    np.savez_compressed(os.path.join(HERE, "golden_spod.npz"), **out)
    print("wrote golden_spod.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
