"""GPU (-m gpu): the reference's OWN test-suite (copied unmodified into oracle/_ref/tests by oracle/make_ref.py) run on
the B200 after ``brevitas_b200.install()``:

* op level (``fuse=False``): tests/brevitas/function -- the STE wrappers must dispatch to ``torch.ops.autograd_ste_ops.
  <name>`` (the mock tests patch exactly those dotted names, test_ops_ste.py:21-22, 68) -- and tests/brevitas/core's
  statistics / scaling tests on the reference's own classes over the 12 kernels;
* module level (``fuse=True``): the same core tests plus the nn / proxy tests, i.e. whole ``QuantLinear`` /
  ``QuantConv2d`` / ``QuantReLU`` layers built by the reference's injector out of the fused classes.

Deselected, each for a reason that is not numerics:
* test_ops_ste.py::test_jit_annotations ties ``fn_prefix == torch`` to ``BREVITAS_JIT=1`` (SURVEY probe D.3): with
  the native namespace bound and the JIT off it fails by construction.
"""
import os
import subprocess
import sys

import pytest

from ref_util import ROOT, reference_root, reference_src

pytestmark = pytest.mark.gpu

COMPAT = os.path.join(ROOT, "brevitas_b200", "_compat")

FUNCTION = ["tests/brevitas/function/test_ops_ste.py", "tests/brevitas/function/test_ops.py",
            "tests/brevitas/function/test_autograd_ste_ops.py", "tests/brevitas/function/test_shape.py"]
CORE = ["tests/brevitas/core/test_stats.py", "tests/brevitas/core/test_standalone_scaling.py",
        "tests/brevitas/core/test_stats_view_wrapper.py"]
LAYERS = ["tests/brevitas/nn/test_linear.py", "tests/brevitas/nn/test_conv2d.py", "tests/brevitas/nn/test_act.py",
          "tests/brevitas/proxy"]
DESELECT = ["tests/brevitas/function/test_ops_ste.py::test_jit_annotations"]


def run_suite(files, fuse):
    root = reference_root()
    if root is None or not os.path.isdir(os.path.join(root, "tests", "brevitas")):
        pytest.skip("reference tests not available (oracle/make_ref.py)")
    desel = "".join(f", '--deselect', {d!r}" for d in DESELECT)
    code = f"""
import sys, unittest.mock
sys.modules['mock'] = unittest.mock
sys.path[:0] = [{os.path.join(ROOT, 'tests')!r}, {ROOT!r}, {root!r}]
import torch, pytest
import brevitas_b200
from brevitas_b200 import _kernels
brevitas_b200.install({reference_src()!r}, fuse={fuse})
import ref_cuda_plugin
rc = pytest.main(['-p', 'no:cacheprovider', '--rootdir={root}', '-c', '/dev/null', '-q', '--no-header', '-x'{desel}]
                 + {files!r}, plugins=[ref_cuda_plugin])
print('KERNEL_LAUNCHES', _kernels.launch_count)
sys.exit(int(rc))
"""
    env = dict(os.environ, BREVITAS_JIT="0")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=1500, env=env)
    tail = r.stdout[-3000:] + r.stderr[-1500:]
    assert r.returncode == 0, tail
    launches = int(r.stdout.split("KERNEL_LAUNCHES")[-1].split()[0])
    return launches, tail


def test_reference_function_and_core_tests_on_b200_ops():
    launches, tail = run_suite(FUNCTION + CORE, fuse=False)      # (mock / property tests: few reach a kernel)
    print(tail[-400:], "kernel launches:", launches)


def test_reference_core_and_layer_tests_on_fused_classes():
    launches, tail = run_suite(CORE + LAYERS, fuse=True)
    assert launches > 0
    print(tail[-400:])
