"""Batch-norm + ReLU + activation quantizer of a conv-net block as ONE fused operator.

``bn_act_quant(bn, act, x)`` computes what ``act(bn(x))`` computes for ``bn: torch.nn.BatchNorm2d`` and ``act`` a
``QuantReLU`` / ``QuantIdentity`` layer (this repository's mirror or the reference's ``brevitas.nn`` layer bound by
``brevitas_b200.install()``) whose quantizer is a ``RescalingIntQuant`` with a scale that does not depend on the
activation (learned ``ParameterScaling`` / ``ParameterFromRuntimeStatsScaling`` past its collection phase, constant,
eval-mode runtime statistics) -- in 8 passes over the activation per training step instead of 13: the normalised tensor
and the ReLU output are never written, the backward recomputes them from the conv output (csrc/bn_act_quant.cu).
Whenever a precondition does not hold (collection phase, NCHW tensor, odd channel count, unknown quantizer) it calls
the two modules one after the other, so it is always safe to use.

Numerics: the quantizer arithmetic is the reference chain, bit-identical to the unfused kernels on the same normalised
value; the batch statistics are summed in a different (fixed, fp64-combined) order than cuDNN's batch-norm, so the
normalised value can differ from the unfused pair by ~1 ulp and a quantized output by one step where that ulp crosses a
rounding boundary (tests/test_gpu_fused_bn.py states the bound).
"""
from typing import Optional

import torch
from torch import nn

from . import _kernels as K
from . import _lib


class _BnActQuantFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, scale, residual, running_mean, running_var, momentum, eps, training, zero_point, qmin,
                qmax, clamp_mode, relu):
        y, save_mean, save_invstd = K.bn_act_quant_fwd(x, gamma, beta, running_mean, running_var, momentum, eps, training,
                                                       scale, zero_point, qmin, qmax, relu, residual)
        ctx.save_for_backward(x, gamma, beta, scale, save_mean, save_invstd, residual)
        ctx.cfg = (zero_point, qmin, qmax, clamp_mode, relu, training)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, gamma, beta, scale, save_mean, save_invstd, residual = ctx.saved_tensors
        zero_point, qmin, qmax, clamp_mode, relu, training = ctx.cfg
        if not training:
            raise RuntimeError("bn_act_quant: backward through eval-mode batch-norm is not fused; call act(bn(x))")
        want_gs = ctx.needs_input_grad[3]
        gx, ggamma, gbeta, gscale, gres = K.bn_act_quant_bwd(
            gy, x, gamma, beta, save_mean, save_invstd, scale, zero_point, qmin, qmax, clamp_mode, relu, want_gs, residual,
            residual is not None and ctx.needs_input_grad[4])
        if gscale is not None:
            gscale = gscale.to(scale.dtype).view(scale.shape)
        return (gx, ggamma if gamma is not None else None, gbeta if beta is not None else None, gscale, gres,
                None, None, None, None, None, None, None, None, None, None)


def _tensor_quant_of(act: nn.Module):
    proxy = getattr(act, "act_quant", None)
    fq = getattr(proxy, "fused_activation_quant_proxy", None)
    if fq is None:
        return None, None, None
    return proxy, fq, getattr(fq, "tensor_quant", None)


def bn_act_quant(bn: nn.Module, act: nn.Module, x: torch.Tensor, residual: Optional[torch.Tensor] = None):
    """``act(bn(x))`` -- or ``act(bn(x) + residual)``, the closing block of a ResNet BasicBlock -- fused when possible
    (see the module docstring), the unfused modules otherwise"""
    from .core.quant import RescalingIntQuant, _NoDelay
    proxy, fq, tq = _tensor_quant_of(act)
    ok = (isinstance(x, torch.Tensor) and type(bn) is nn.BatchNorm2d and tq is not None and type(tq) is RescalingIntQuant
          and K.bn_act_quant_supported(x) and x.dtype == torch.float32 and bn.track_running_stats
          and (bn.training or not torch.is_grad_enabled() or not x.requires_grad)
          and type(fq.activation_impl) in (nn.ReLU, nn.Identity)
          and getattr(act, "input_quant", None) is not None and not act.input_quant.is_quant_enabled
          and type(tq.int_quant.delay_wrapper.delay_impl) is _NoDelay)
    if ok:
        independent = getattr(tq.scaling_impl, "input_independent", None)
        ok = independent is not None and independent()
    if ok:
        bit_width = tq.msb_clamp_bit_width_impl()
        cfg = tq._host_config(bit_width.dtype)
        ok = cfg is not None and cfg[0] is not None and cfg[3] == _lib.ROUND
    if ok:
        zp, qmin, qmax, rm, cm, _ = cfg
        scale = tq.scaling_impl(x) / tq.int_scaling_impl(bit_width)
        ok = scale.dtype == x.dtype and scale.numel() in (1, x.shape[1])
    if ok and residual is not None:
        ok = (isinstance(residual, torch.Tensor) and residual.shape == x.shape and residual.dtype == x.dtype
              and residual.stride() == x.stride())
    if not ok or (bn.training and bn.momentum is None):    # (cumulative moving average: left to torch)
        return act(bn(x)) if residual is None else act(bn(x) + residual)
    if bn.training:
        bn.num_batches_tracked.add_(1)
    relu = type(fq.activation_impl) is nn.ReLU
    y = _BnActQuantFn.apply(x, bn.weight, bn.bias, scale, residual, bn.running_mean, bn.running_var,
                            bn.momentum if bn.momentum is not None else 0.0, bn.eps, bn.training, zp, qmin, qmax, cm, relu)
    if not getattr(act, "return_quant_tensor", False):
        return y
    zero_point = tq.zero_point_impl(x, scale, bit_width)
    signed = proxy.is_signed if hasattr(proxy, "is_signed") else getattr(proxy, "signed", None)
    return _make_quant_tensor(act, y, scale, zero_point, bit_width, signed)


def _make_quant_tensor(act, value, scale, zero_point, bit_width, signed):
    """the QuantTensor type of whichever front-end ``act`` belongs to"""
    if type(act).__module__.startswith("brevitas.") or type(act).__module__.startswith("brevitas_examples"):
        from brevitas.quant_tensor import QuantTensor
        return QuantTensor(value, scale, zero_point, bit_width, signed, act.training)
    from .nn import QuantTensor
    return QuantTensor(value, scale, zero_point, bit_width, signed, act.training)
