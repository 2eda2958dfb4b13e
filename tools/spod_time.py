#!/usr/bin/env python
"""Times the SPOD reconstruction kernel (mvrl_flow_reconstruct) at the reference's data size: modes [41, 61, 3, K]
complex128, coeffs [K, Nt] complex128 (params_coeffs.yaml: 17 frequencies x 32 modes -> K = 544; Nt = 2000), fp32 output
field, against the library route it replaced (torch complex matmul + real + transpose + mean + cast).  Prints one JSON line."""
import json
import sys
import os
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marinevehiclereinforcementlearning_b200.tag_00_Dec2023_simpleControlTurbulence import flowGenerator  # noqa: E402
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
from _spod_blobs import spod_blobs  # noqa: E402

K, NT = 544, 2000
modes, coeffs = spod_blobs((41, 61, 3), K, NT, seed=1)
mean = np.random.default_rng(2).standard_normal((41, 61, 3))
dev = torch.device("cuda")
from marinevehiclereinforcementlearning_b200 import _lib  # noqa: E402
lib = _lib.load()
m = torch.as_tensor(np.ascontiguousarray(modes.reshape(-1, K)).view(np.float64), device=dev)
c = torch.as_tensor(np.ascontiguousarray(coeffs).view(np.float64), device=dev)
mu = torch.as_tensor(mean.reshape(-1), device=dev)
out = torch.empty((NT, 41, 61, 3), dtype=torch.float32, device=dev)
P = 41 * 61 * 3


def ours():
    _lib.check(lib.mvrl_flow_reconstruct(_lib.F32, P, K, NT, _lib.ptr(m), 1, _lib.ptr(c), 1, _lib.ptr(mu), _lib.ptr(out), _lib.current_stream(dev)))


mc = torch.as_tensor(modes.reshape(-1, K), device=dev)
cc = torch.as_tensor(coeffs, device=dev)


def library():
    return (torch.matmul(mc, cc).real.T.reshape(NT, 41, 61, 3) + mu.reshape(41, 61, 3)).to(torch.float32)


def timed(f, reps=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t_ours, t_lib = timed(ours), timed(library)
err = float((out - library()).abs().max())
t0 = time.time(); ref = np.real(modes.reshape(-1, K)[:, :] @ coeffs[:, :8]).T.reshape(8, 41, 61, 3) + mean; t_np = (time.time() - t0) / 8 * NT
err_np = float(np.abs(out[:8].cpu().numpy() - ref).max())
flop = 2.0 * 2 * P * K * NT   # two real products per complex term, FMA = 2
print(json.dumps({"what": "SPOD reconstruction, modes [41,61,3,%d] x coeffs [%d,%d] complex128 -> fp32 field" % (K, K, NT),
                  "ms_kernel": t_ours, "fp64_tflops": flop / t_ours / 1e9, "ms_torch_complex_matmul_route": t_lib,
                  "s_numpy_route_extrapolated_from_8_levels": t_np, "max_abs_diff_vs_torch": err, "max_abs_diff_vs_numpy_8_levels": err_np}))
