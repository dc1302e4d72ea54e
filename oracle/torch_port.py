"""Plain-PyTorch (CPU, ATen) restatement of the reference's module graphs (TEST INFRASTRUCTURE, see __init__).

Purpose: (1) autograd-DERIVED gradients for the parity tests -- the numpy oracle has the closed form, this file
has the graph, and both are pinned against the real reference by tests/golden; (2) the ``cpu_baseline`` /
``--impl reference`` leg of bench.py: the same ATen op sequence the reference runs on the host cores
(``kind: "port"`` -- /root/reference does not exist on the GPU box).

Each function cites the reference lines it restates (paths relative to /root/reference/src/brevitas).
"""
import math

import torch
from torch import Tensor


# ---- STE primitives: ops/autograd_ste_ops.py ---------------------------------------------------------------------
class _Ste(torch.autograd.Function):
    """forward = fn(x); backward = identity (ops/autograd_ste_ops.py:62-64 and siblings)."""

    @staticmethod
    def forward(ctx, x, fn):
        return fn(x)

    @staticmethod
    def backward(ctx, g):
        return g, None


class _ClampSte(torch.autograd.Function):
    """ops/autograd_ste_ops.py:100-120: where-clamp forward, (g, None, None) backward."""

    @staticmethod
    def forward(ctx, x, lo, hi):
        return tensor_clamp(x, lo, hi)

    @staticmethod
    def backward(ctx, g):
        return g, None, None


class _AbsBinarySignGrad(torch.autograd.Function):
    """ops/autograd_ste_ops.py:355-377"""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(binary_sign(x).type(torch.int8))
        return torch.abs(x)

    @staticmethod
    def backward(ctx, g):
        (s,) = ctx.saved_tensors
        return s.float() * g


def binary_sign(x):                       # function/ops.py:17-34
    return torch.ge(x, 0.0).to(x.dtype) - torch.lt(x, 0.0).to(x.dtype)


def round_to_zero(x):                     # function/ops.py:38-53
    return torch.sign(x) * torch.floor(torch.abs(x))


def dpu_round(x):                         # function/ops.py:57-72
    return torch.where((x < 0.) & (x - torch.floor(x) == 0.5), torch.ceil(x), torch.round(x))


def tensor_clamp(x, min_val, max_val):    # function/ops.py:76-100
    out = torch.where(x > max_val, max_val.type_as(x), x)
    return torch.where(out < min_val, min_val.type_as(out), out)


def max_int(signed, narrow_range, bit_width):   # function/ops.py:133-160
    if not signed and not narrow_range:
        return (2 ** bit_width) - 1
    if not signed and narrow_range:
        return (2 ** bit_width) - 2
    return (2 ** (bit_width - 1)) - 1


def min_int(signed, narrow_range, bit_width):   # function/ops.py:164-191
    if signed and narrow_range:
        return - (2 ** (bit_width - 1)) + 1
    if signed and not narrow_range:
        return - (2 ** (bit_width - 1))
    return 0 * bit_width


ROUND = {"round": torch.round, "floor": torch.floor, "ceil": torch.ceil, "round_to_zero": round_to_zero,
         "dpu_round": dpu_round}


def round_ste(x, mode="round"):
    return _Ste.apply(x, ROUND[mode])


def binary_sign_ste(x):
    return _Ste.apply(x, binary_sign)


def abs_binary_sign_grad(x):
    return _AbsBinarySignGrad.apply(x)


def clamp_min_ste(x, min_val):
    return _Ste.apply(x, lambda v: torch.clamp_min(v, min_val))


# ---- IntQuant: core/quant/int_base.py:64-97 ---------------------------------------------------------------------
def int_quant(x, scale, zero_point, bit_width, signed, narrow_range, round_mode="round", clamp_ste=False,
              return_codes=False):
    y = x / scale
    y = y + zero_point
    lo = min_int(signed, narrow_range, bit_width)
    hi = max_int(signed, narrow_range, bit_width)
    y = round_ste(y, round_mode)
    y = _ClampSte.apply(y, lo, hi) if clamp_ste else tensor_clamp(y, lo, hi)
    codes = y
    y = y - zero_point
    y = y * scale
    return (y, codes) if return_codes else y


def int_scaling(signed, narrow_range, bit_width):    # core/scaling/int_scaling.py:20-24
    return -min_int(signed, narrow_range, bit_width) if signed else max_int(signed, narrow_range, bit_width)


# ---- RescalingIntQuant with abs-max statistics: core/quant/int.py:156-163, core/scaling/runtime.py:19-102,
#      core/stats/stats_op.py:129-141 -----------------------------------------------------------------------------
def absmax_threshold(x, view, scaling_min_val):
    """view: 'tensor' | 'rows' (dim 0 x rest) | 'batch_channel' (dims 0,1 x rest)."""
    if view == "tensor":
        stat = torch.max(torch.abs(x.reshape(-1)))
        shape = ()
    elif view == "rows":
        stat = torch.max(torch.abs(x.reshape(x.shape[0], -1)), dim=1)[0]
        shape = (x.shape[0],) + (1,) * (x.dim() - 1)
    else:
        stat = torch.max(torch.abs(x.reshape(x.shape[0], x.shape[1], -1)), dim=2)[0]
        shape = (x.shape[0], x.shape[1]) + (1,) * (x.dim() - 2)
    stat = stat.view(shape)
    if scaling_min_val is not None and scaling_min_val != 0:
        stat = clamp_min_ste(stat, scaling_min_val)
    return stat


def rescaling_int_quant_absmax(x, view, bit_width, signed, narrow_range, scaling_min_val=1e-10, round_mode="round",
                               clamp_ste=True):
    """e.g. Int8WeightPerChannelFloat: view='rows', signed, narrow, clamp_ste."""
    bw = torch.tensor(float(bit_width))
    threshold = absmax_threshold(x, view, scaling_min_val)
    scale = threshold / int_scaling(signed, narrow_range, bw)
    zero_point = torch.tensor(0.0)
    y = int_quant(x, scale, zero_point, bw, signed, narrow_range, round_mode, clamp_ste)
    return y, scale, zero_point, bw


# ---- BinaryQuant / ClampedBinaryQuant: core/quant/binary.py:60-64, 120-125 --------------------------------------
def binary_quant(x, scale):
    return binary_sign_ste(x) * scale


def clamped_binary_quant(x, scale):
    return binary_sign_ste(tensor_clamp(x, - scale, scale)) * scale


# ---- AbsPercentile: core/stats/stats_op.py:41-66 ----------------------------------------------------------------
def abs_percentile(x, q, reduce_dim=None):
    if reduce_dim is None:
        k = int(math.floor(.01 * q * x.numel() + 0.5))
        return x.abs().view(-1).kthvalue(k).values
    other = abs(reduce_dim - 1)
    k = int(math.floor(.01 * q * torch.narrow(x, dim=other, start=0, length=1).numel() + 0.5))
    return x.abs().kthvalue(k, dim=reduce_dim).values
