"""brevitas_b200 -- B200-native (sm_100a) fake-quantization hot path, a drop-in for the one native plugin and
the ``tensor_quant`` modules of Brevitas (reference: Giuseppe5/brevitas, src/brevitas/csrc, core/quant,
core/scaling, core/stats, function/ops_ste).  See DESIGN.md and INTEGRATION.md.

There is no CPU or eager fallback: the package requires ``libbrevitas_b200.so`` (built by
``__graft_entry__.build()`` / ``make -C brevitas_b200/csrc``) and CUDA tensors.
"""
__version__ = "0.1.0"

from . import _lib

_lib.load()          # fail loudly at import when the native library is missing

from . import ops  # noqa: E402,F401  (registers torch.ops.autograd_ste_ops.* and torch.ops.brevitas_b200.*)
from . import function  # noqa: E402,F401
from . import core  # noqa: E402,F401
from .binding import install, uninstall  # noqa: E402,F401  (binds an unmodified Brevitas installation to the kernels)
from .fused_bn import bn_act_quant, fuse_batch_norm, unfuse_batch_norm  # noqa: E402,F401
