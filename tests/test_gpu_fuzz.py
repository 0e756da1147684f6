"""GPU (-m gpu): randomised differential test of the MSM and of commit+open against the oracle -- random sizes
(2^1..2^13), window widths, table modes, batched-affine rounds and scalar shapes (random, small, sparse, all equal,
near r, ragged length).  The driver is tools/fuzz_msm.py so that it can also be run by hand with more cases."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_randomised_msm_and_open_vs_oracle():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_msm.py"), "60"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "fuzz: 60 cases, 0 mismatches" in out.stdout, out.stdout
