"""GPU (-m gpu): a short run of the randomised differential test (tests/fuzz_kernels.py): random shapes, dtypes, scale
layouts, ranges, round / clamp modes and kernel variants (plain, fused ReLU, tensor zero-point, integer export, fused
abs-max), every result bit-exact against the numpy oracle (scale-gradient sums within the reduction tolerance)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [11, 12])
def test_fuzz_against_oracle(seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz_kernels.py"), "--cases", "150", "--seed", str(seed)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "fuzz OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
