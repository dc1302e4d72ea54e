"""Mirror of ``brevitas.function.ops`` (src/brevitas/function/ops.py) -- the non-STE helper functions.

The integer-range helpers work on the tiny 0-dim ``bit_width`` tensors exactly like the reference (they stay
on ATen); the element-wise functions over full tensors run on the sm_100a kernels.
"""
import torch
from torch import Tensor

__all__ = ['binary_sign', 'round_to_zero', 'dpu_round', 'tensor_clamp', 'tensor_clamp_', 'identity', 'max_int',
           'min_int']


def _k(name, x):
    from .. import _kernels
    return _kernels.unary(name, x)


def binary_sign(x: Tensor) -> Tensor:
    """2-valued sign ``(x >= 0) - (x < 0)`` (reference: function/ops.py:17-34).

    Built from comparisons in the reference, so the result carries no gradient; same here (detached)."""
    return _k("bvb_binary_sign_ste_impl", x.detach())


def round_to_zero(x: Tensor) -> Tensor:
    """``sign(x) * floor(abs(x))`` (reference: function/ops.py:38-53); forward values only (detached)."""
    return _k("bvb_round_to_zero_ste_impl", x.detach())


def dpu_round(x: Tensor) -> Tensor:
    """DPU rounding (reference: function/ops.py:57-72); forward values only (detached)."""
    return _k("bvb_dpu_round_ste_impl", x.detach())


def tensor_clamp(x: Tensor, min_val: Tensor, max_val: Tensor) -> Tensor:
    """Clamp with tensor bounds, differentiable w.r.t. x, min_val and max_val (reference: function/ops.py:76-100).

    This is the *standalone* differentiable form; it keeps the reference's two ``torch.where`` calls so that the
    autograd semantics (masked gradient, gradients to the bounds) are inherited verbatim.  Inside ``IntQuant`` the
    same arithmetic is fused into the quant kernels (``brevitas_b200::int_quant`` with ``CLAMP_MASKED``).
    """
    if x.is_cuda and _clamp_kernel_ok(x, min_val, max_val):
        return _TensorClampFn.apply(x, min_val, max_val)
    out = torch.where(x > max_val, max_val.type_as(x), x)
    out = torch.where(out < min_val, min_val.type_as(out), out)
    return out


def _clamp_kernel_ok(x: Tensor, min_val: Tensor, max_val: Tensor) -> bool:
    from .._kernels import _DTYPES, broadcast_pattern
    if x.dtype not in _DTYPES or not (min_val.is_cuda and max_val.is_cuda) or x.numel() == 0:
        return False
    try:                                    # bounds must follow the scale[(i / inner) % count] broadcast of the kernels
        broadcast_pattern(x.shape, min_val.shape)
        broadcast_pattern(x.shape, max_val.shape)
    except Exception:
        return False
    return True


class _TensorClampFn(torch.autograd.Function):
    """the where-based clamp (forward: bvb_tensor_clamp_ste_impl) with the gradients plain autograd gives the reference's
    two ``torch.where`` calls: masked for x, summed over the replaced elements for the bounds (bvb_tensor_clamp_bwd)"""

    @staticmethod
    def forward(ctx, x, min_val, max_val):
        from .. import _kernels
        ctx.save_for_backward(x, min_val, max_val)
        return _kernels.tensor_clamp(x, min_val, max_val)

    @staticmethod
    def backward(ctx, g):
        from .. import _kernels
        x, min_val, max_val = ctx.saved_tensors
        gx, gmin, gmax = _kernels.tensor_clamp_bwd(g.to(x.dtype), x, min_val, max_val, ctx.needs_input_grad[1],
                                                   ctx.needs_input_grad[2])
        if gmin is not None:
            gmin = gmin.to(min_val.dtype).view(min_val.shape)
        if gmax is not None:
            gmax = gmax.to(max_val.dtype).view(max_val.shape)
        return (gx.view(x.shape) if ctx.needs_input_grad[0] else None), gmin, gmax


def tensor_clamp_(x: Tensor, min_val: Tensor, max_val: Tensor) -> Tensor:
    """In-place clamp, not differentiable (reference: function/ops.py:104-111). Returns ``x``."""
    from .. import _kernels
    _kernels.tensor_clamp(x.detach(), min_val, max_val, inplace=True)
    return x


def identity(x: Tensor) -> Tensor:
    return x


def max_int(signed: bool, narrow_range: bool, bit_width: Tensor) -> Tensor:
    """Largest representable integer (reference: function/ops.py:133-160)."""
    if not signed and not narrow_range:
        return (2 ** bit_width) - 1
    if not signed and narrow_range:
        return (2 ** bit_width) - 2
    return (2 ** (bit_width - 1)) - 1


def min_int(signed: bool, narrow_range: bool, bit_width: Tensor) -> Tensor:
    """Smallest representable integer (reference: function/ops.py:164-191)."""
    if signed and narrow_range:
        return - (2 ** (bit_width - 1)) + 1
    if signed and not narrow_range:
        return - (2 ** (bit_width - 1))
    return 0 * bit_width
