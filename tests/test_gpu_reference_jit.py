"""GPU (-m gpu): ``BREVITAS_JIT=1`` -- the reference scripts its STE wrappers and core modules with TorchScript and binds
``torch.ops.autograd_ste_ops.*`` AT COMPILE TIME (src/brevitas/function/ops_ste.py:38-47, src/brevitas/jit.py:19-29),
which only works because ``brevitas_b200.ops`` defines real dispatcher ops with schemas and autograd attached
(SURVEY.md §8b).  ``install()`` answers the reference's own ``cpp_extension.load('autograd_ste_ops', ...)`` call
(src/brevitas/__init__.py:60-71) without compiling its C++ plugin.  Runs in a subprocess (the flag is read at import)."""
import os
import subprocess
import sys

import pytest

from ref_util import ROOT, reference_src

pytestmark = pytest.mark.gpu

CODE = r"""
import os, sys
sys.path[:0] = [{tests!r}, {root!r}]
import torch
import brevitas_b200
from brevitas_b200 import _kernels
brevitas_b200.install({src!r}, fuse=False)
import brevitas
from brevitas.function import ops_ste
assert brevitas.config.JIT_ENABLED and brevitas.NATIVE_STE_BACKEND_LOADED and ops_ste.fn_prefix is torch
assert isinstance(ops_ste.round_ste, torch.jit.ScriptFunction), type(ops_ste.round_ste)
assert "autograd_ste_ops::round_ste_impl" in str(ops_ste.round_ste.graph)
assert "autograd_ste_ops::tensor_clamp_ste_impl" in str(ops_ste.tensor_clamp_ste.graph)
from brevitas.core.quant import IntQuant, RescalingIntQuant
iq = IntQuant(narrow_range=True, signed=True)
assert isinstance(iq, torch.jit.ScriptModule)
from test_gpu_reference_binding import compare_group, run_generator
total = 0
for group in ("ste", "int_quant", "weight_stats", "binary", "kat", "runtime_token"):
    before = _kernels.launch_count
    out = run_generator(group)
    # forward values bit-exact; gradients by tolerance: a scripted module's backward is TorchScript's own symbolic
    # autodiff (and fuser), whose op order differs from eager autograd's -- a property of the reference under the JIT
    checked, exact = compare_group(group, out, ("f32",), grads_by_tolerance=True)
    n = _kernels.launch_count - before
    assert n > 0, group
    total += n
    print(group, checked, "arrays", exact, "bit-exact", n, "kernel launches from TorchScript")
print("JIT_OK", total)
"""


def test_scripted_reference_runs_on_the_b200_ops():
    src = reference_src()
    if src is None:
        pytest.skip("reference not available (oracle/make_ref.py)")
    env = dict(os.environ, BREVITAS_JIT="1", PYTORCH_JIT="1")
    code = CODE.format(tests=os.path.join(ROOT, "tests"), root=ROOT, src=src)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=1500, env=env, cwd=ROOT)
    assert r.returncode == 0 and "JIT_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    print(r.stdout[-800:])
