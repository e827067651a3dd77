"""Opt-in kernel variants (kept for the record): the same assembly parity tests with the
thread-per-dof ELL row kernel (KNP_ROWS=ell) and with the list-driven phase 2b (KNP_ROWS_LISTS=1).  The variants are
selected when the context is created, hence one pytest subprocess per variant.  (Named z... to run after the main suite.)"""
import os
import subprocess
import sys
import pytest
from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("env", [{"KNP_ROWS": "ell"}, {"KNP_ROWS_LISTS": "1"}], ids=["ell", "lists"])
def test_assembly_parity_with_opt_in_row_kernels(env):
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-q", "-m", "gpu",
                        "-k", "assembled_matrix or membrane_models or bitwise or values_against"],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ, **env), cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout
