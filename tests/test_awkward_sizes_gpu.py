"""GPU: every kernel family on small, awkwardly sized batches (n = 1, 31, 130, 257; odd leading dimensions
and offsets; masks; NaN states; both fp32 instantiations; the host-buffer pipeline) must run without a
CUDA fault.  The script is tools/sanitize_smoke.py (also usable under compute-sanitizer where that is open)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_all_kernels_on_awkward_sizes():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize_smoke.py")], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "sanitize smoke ok" in res.stdout
