"""tcgen05 GEMM (3xBF16 split and single BF16) against float64, run in a subprocess with a timeout so a pipeline
deadlock fails the test instead of hanging the session."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tcgen05_gemm_matches_float64():
    try:
        r = subprocess.run([sys.executable, "-m", "tests.tc_check", "full"], cwd=ROOT, capture_output=True, text=True,
                           timeout=180)
    except subprocess.TimeoutExpired as e:
        pytest.fail(f"tcgen05 GEMM check timed out (pipeline deadlock?)\n{e.stdout}\n{e.stderr}")
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
