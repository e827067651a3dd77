"""The fallback row kernel: the edge-lane kernel is the default; the scan kernel (KNP_ROWS=scan) serves meshes whose edge
rings or vertex degrees exceed the lane-group tables, so it has to pass the same assembly parity tests.  The kernel is
chosen when the context is created, hence a pytest subprocess.  (Named z... to run after the main suite.)"""
import os
import subprocess
import sys
import pytest
from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_assembly_parity_with_the_scan_row_kernel():
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-q", "-m", "gpu",
                        "-k", "assembled_matrix or membrane_models or bitwise or values_against"],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ, KNP_ROWS="scan"), cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout
