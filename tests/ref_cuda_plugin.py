"""pytest plugin (test infrastructure): run the REFERENCE's own CPU-written tests on the GPU, unchanged.  Every tensor
a test, fixture or hypothesis strategy creates from non-tensor arguments is moved to ``cuda`` (``ref_util.FactoryToCuda``),
so modules and inputs built inside the tests live on the device and flow through ``torch.ops.autograd_ste_ops.*`` /
the fused ``tensor_quant`` modules that ``brevitas_b200.install()`` bound."""
import pytest
import torch

from ref_util import FactoryToCuda

# the legacy constructor torch.Tensor([...]) (tests/brevitas/core/test_stats.py:15) is not a dispatched factory call;
# it follows the default tensor type
torch.set_default_tensor_type(torch.cuda.FloatTensor)


@pytest.hookimpl(hookwrapper=True)
def pytest_runtest_protocol(item, nextitem):
    with FactoryToCuda():
        yield
