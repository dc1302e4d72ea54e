"""Mirror of ``brevitas.function.ops_ste`` (src/brevitas/function/ops_ste.py:47-370).

Each wrapper dispatches to ``torch.ops.autograd_ste_ops.<name>_impl`` -- the reference's own plugin namespace,
here implemented by hand-written sm_100a kernels (``brevitas_b200.ops``).  Like the reference, under tracing the
plain op is emitted so that exporters see Round/Clip instead of a custom op (ops_ste.py:65-66).
"""
import torch
from torch import Tensor

from .. import ops as _ops  # noqa: F401  (registers the dispatcher ops)
from . import ops as _f

__all__ = ['round_ste', 'ceil_ste', 'floor_ste', 'tensor_clamp_ste', 'tensor_clamp_ste_', 'scalar_clamp_ste',
           'scalar_clamp_min_ste', 'binary_sign_ste', 'ternary_sign_ste', 'round_to_zero_ste', 'dpu_round_ste',
           'abs_binary_sign_grad']

fn_prefix = torch     # the "native backend" is always the one in use here


def _tracing() -> bool:
    return bool(torch._C._get_tracing_state())


def round_ste(x: Tensor) -> Tensor:
    """``torch.round`` (half to even) with identity gradient (reference: ops_ste.py:47-67)."""
    if _tracing():
        return torch.round(x)
    return torch.ops.autograd_ste_ops.round_ste_impl(x)


def ceil_ste(x: Tensor) -> Tensor:
    """``torch.ceil`` with identity gradient (reference: ops_ste.py:71-91)."""
    if _tracing():
        return torch.ceil(x)
    return torch.ops.autograd_ste_ops.ceil_ste_impl(x)


def floor_ste(x: Tensor) -> Tensor:
    """``torch.floor`` with identity gradient (reference: ops_ste.py:95-115)."""
    if _tracing():
        return torch.floor(x)
    return torch.ops.autograd_ste_ops.floor_ste_impl(x)


def tensor_clamp_ste(x: Tensor, min_val: Tensor, max_val: Tensor) -> Tensor:
    """Tensor-bound clamp with pass-through gradient to x only (reference: ops_ste.py:119-145)."""
    if _tracing():
        return _f.tensor_clamp(x, min_val, max_val)
    return torch.ops.autograd_ste_ops.tensor_clamp_ste_impl(x, min_val, max_val)


def tensor_clamp_ste_(x: Tensor, min_val: Tensor, max_val: Tensor) -> Tensor:
    """In-place variant (reference: ops_ste.py:149-172); mutates and returns ``x`` (Python-backend semantics)."""
    if _tracing():
        return _f.tensor_clamp_(x, min_val, max_val)
    return torch.ops.autograd_ste_ops.tensor_clamp_ste_impl_(x, min_val, max_val)


def scalar_clamp_ste(x: Tensor, min_val: float, max_val: float) -> Tensor:
    """``torch.clamp`` with scalar bounds and identity gradient (reference: ops_ste.py:176-201)."""
    if _tracing():
        return torch.clamp(x, min_val, max_val)
    return torch.ops.autograd_ste_ops.scalar_clamp_ste_impl(x, min_val, max_val)


def scalar_clamp_min_ste(x: Tensor, min_val: float) -> Tensor:
    """``torch.clamp_min`` with identity gradient (reference: ops_ste.py:205-229)."""
    if _tracing():
        return torch.clamp_min(x, min_val)
    return torch.ops.autograd_ste_ops.scalar_clamp_min_ste_impl(x, min_val)


def binary_sign_ste(x: Tensor) -> Tensor:
    """2-valued sign with identity gradient (reference: ops_ste.py:245-267)."""
    if _tracing():
        return _f.binary_sign(x)
    return torch.ops.autograd_ste_ops.binary_sign_ste_impl(x)


def ternary_sign_ste(x: Tensor) -> Tensor:
    """``torch.sign`` with identity gradient (reference: ops_ste.py:271-292)."""
    if _tracing():
        return torch.sign(x)
    return torch.ops.autograd_ste_ops.ternary_sign_ste_impl(x)


def round_to_zero_ste(x: Tensor) -> Tensor:
    """Round toward zero with identity gradient (reference: ops_ste.py:296-317)."""
    if _tracing():
        return _f.round_to_zero(x)
    return torch.ops.autograd_ste_ops.round_to_zero_ste_impl(x)


def dpu_round_ste(x: Tensor) -> Tensor:
    """DPU rounding with identity gradient (reference: ops_ste.py:321-342)."""
    if _tracing():
        return _f.dpu_round(x)
    return torch.ops.autograd_ste_ops.dpu_round_ste_impl(x)


def abs_binary_sign_grad(x: Tensor) -> Tensor:
    """``torch.abs`` whose gradient is ``binary_sign(x)`` (1 at 0) (reference: ops_ste.py:346-370)."""
    if _tracing():
        return torch.abs(x)
    return torch.ops.autograd_ste_ops.abs_binary_sign_grad_impl(x)
