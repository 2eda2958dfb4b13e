"""Synthetic stand-ins for the SPOD blobs the reference checkout lacks (``coeffs.npy`` / ``modes_r.npy`` are listed in
.MISSING_LARGE_BLOBS), in the file layout ``ReconstructedFlow.__init__`` reads (tag_00.../flowGenerator.py:14-30).

Shared by ``gen_golden_spod.py`` (runs the UNMODIFIED reference constructor on such a directory, in the build
container) and by the tests (which build the same directory from the committed golden file and hand it to the CUDA
path) - so both sides read byte-identical inputs.
"""
import os

import numpy as np


def spod_blobs(plane_shape, n_modes, nt, seed, complex_valued=True):
    """modes [Ny, Nx, 3, K], coeffs [K, Nt]: seeded, amplitudes decaying with the mode index."""
    rng = np.random.default_rng(seed)
    amp = 0.08 / (1.0 + np.arange(n_modes)) ** 0.7
    modes = rng.standard_normal(tuple(plane_shape) + (n_modes,))
    coeffs = rng.standard_normal((n_modes, nt)) * amp[:, None]
    if complex_valued:
        modes = modes + 1j * rng.standard_normal(modes.shape)
        coeffs = coeffs + 1j * rng.standard_normal(coeffs.shape) * amp[:, None]
    return modes, coeffs


def write_spod_dir(path, lt_mean, coords, time_step, n_modes, nt, seed, complex_valued=True):
    os.makedirs(path, exist_ok=True)
    modes, coeffs = spod_blobs(np.shape(lt_mean), n_modes, nt, seed, complex_valued)
    np.save(os.path.join(path, "modes_r.npy"), modes)
    np.save(os.path.join(path, "coeffs.npy"), coeffs)
    np.save(os.path.join(path, "ltm.npy"), np.asarray(lt_mean))
    np.save(os.path.join(path, "turbulence_coords.npy"), np.asarray(coords))
    with open(os.path.join(path, "params_coeffs.yaml"), "w") as f:
        f.write("time_step: %r\nn_modes_save: %d\n" % (float(time_step), n_modes))
    return modes, coeffs
