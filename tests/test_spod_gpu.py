"""GPU parity of the SPOD reconstruction (``ReconstructedFlow.__init__``, tag_00.../flowGenerator.py:14-51; kernel
``mvrl_flow_reconstruct``) through the public constructor: the same directory of synthetic blobs the UNMODIFIED reference
constructor was run on (``tests/golden/gen_golden_spod.py``) is rebuilt from the committed golden file and handed to
``ReconstructedFlow(dataDir)``.  fp64 field: 1e-12; fp32 field: 1e-6 (one rounding of values of order 1)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import oracle_np as o

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from _spod_blobs import spod_blobs, write_spod_dir  # noqa: E402

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from marinevehiclereinforcementlearning_b200.tag_00_Dec2023_simpleControlTurbulence import flowGenerator

DEV = "cuda"


def rel_err(a, ref):
    ref = np.asarray(ref)
    scale = np.abs(ref) + np.abs(ref).max(axis=-1, keepdims=True)
    return (np.abs(np.asarray(a) - ref) / np.maximum(scale, 1e-300)).max()


@pytest.mark.parametrize("tag,cplx", [("c", True), ("r", False)])
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 1e-6)])
def test_reconstructed_flow_constructor_vs_reference_constructor(tmp_path, tag, cplx, dtype, tol):
    g = load_golden("spod")
    write_spod_dir(str(tmp_path), g["ltm"], g["coords"], float(g["time_step"]), int(g["n_modes"]), int(g["nt"]), int(g["seed"]), cplx)
    flow = flowGenerator.ReconstructedFlow(str(tmp_path), dtype=dtype, device=DEV)
    base = flow.baseFlowData.cpu().numpy().astype(float)
    assert base.shape == (int(g["nt"]), 41, 61, 3)
    assert np.abs(base[::2, ::3, ::4, :] - g[tag + "_base_sample"]).max() < tol
    assert np.abs(base[-1] - g[tag + "_base_last"]).max() < tol
    assert np.abs(base.sum(axis=(1, 2)) - g[tag + "_base_plane_sum"]).max() < tol * 41 * 61
    assert np.abs(np.abs(base).sum(axis=(1, 2)) - g[tag + "_base_plane_abs_sum"]).max() < tol * 41 * 61
    assert (flow.baseDt, flow.baseDx, flow.baseDy) == (float(g[tag + "_baseDt"]), float(g[tag + "_baseDx"]), float(g[tag + "_baseDy"]))
    # turbulence intensity attributes of the reference object (flowGenerator.py:47-51)
    assert np.abs(flow.uPrime - g[tag + "_uPrime"]).max() < 10 * tol and np.abs(flow.vPrime - g[tag + "_vPrime"]).max() < 10 * tol
    assert np.abs(flow.TI - g[tag + "_TI"]).max() < 10 * tol and abs(flow.baseTI - float(g[tag + "_baseTI"])) < 10 * tol
    # the env's scaling, then the reference's interp / interpField on the reconstructed field
    flow.scale(11., 1., 2., translate=(-1.65, -1.1))
    res = flow.interp(torch.as_tensor(g[tag + "_interp_t"], device=DEV), torch.as_tensor(g[tag + "_interp_xy"], device=DEV)).cpu().numpy()
    itol = 1e-11 if dtype == torch.float64 else 2e-5
    assert rel_err(res, g[tag + "_interp_res"]) < itol
    assert rel_err(flow.interpField(float(g[tag + "_interp_field_t"])).cpu().numpy(), g[tag + "_interp_field"]) < itol


def test_reconstruct_awkward_shapes_vs_oracle():
    """Tile edges of the kernel (64 x 64 outputs, K in steps of 16): sizes around them, one mode, one time level."""
    rng = np.random.default_rng(2)
    for plane_shape, k, nt in (((3, 5, 3), 1, 1), ((7, 9, 3), 16, 64), ((7, 9, 3), 17, 65), ((2, 11, 3), 33, 63), ((41, 61, 3), 5, 130)):
        modes, coeffs = spod_blobs(plane_shape, k, nt, seed=int(rng.integers(1 << 30)), complex_valued=bool(k & 1))
        mean = rng.standard_normal(plane_shape)
        ref = o.FlowOracle.reconstruct(modes, coeffs, mean)
        got = flowGenerator.ReconstructedFlow.reconstruct(modes, coeffs, mean, dtype=torch.float64, device=DEV).cpu().numpy()
        assert got.shape == ref.shape and np.abs(got - ref).max() < 1e-12, (plane_shape, k, nt)
    # mixed: complex modes with real coefficients and the reverse (the imaginary product vanishes)
    modes, coeffs = spod_blobs((4, 6, 3), 9, 20, seed=4, complex_valued=True)
    for m, c in ((modes, coeffs.real.copy()), (modes.real.copy(), coeffs)):
        ref = o.FlowOracle.reconstruct(m, c, np.zeros((4, 6, 3)))
        got = flowGenerator.ReconstructedFlow.reconstruct(m, c, np.zeros((4, 6, 3)), dtype=torch.float64, device=DEV).cpu().numpy()
        assert np.abs(got - ref).max() < 1e-12
    with pytest.raises(ValueError):
        flowGenerator.ReconstructedFlow.reconstruct(modes, coeffs[:-1], np.zeros((4, 6, 3)), device=DEV)


def test_reconstructed_flow_missing_blob_raises_like_the_reference(tmp_path):
    g = load_golden("spod")
    write_spod_dir(str(tmp_path), g["ltm"], g["coords"], float(g["time_step"]), 4, 3, 1)
    os.remove(os.path.join(str(tmp_path), "coeffs.npy"))
    with pytest.raises(FileNotFoundError):
        flowGenerator.ReconstructedFlow(str(tmp_path), device=DEV)
