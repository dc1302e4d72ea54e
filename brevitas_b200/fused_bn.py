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
        scale = tq.scaling_impl(x) / tq.int_threshold(bit_width)
        ok = scale.dtype == x.dtype and scale.numel() in (1, x.shape[1])
    if ok and residual is not None:
        ok = (isinstance(residual, torch.Tensor) and residual.shape == x.shape and residual.dtype == x.dtype
              and residual.stride() == x.stride())
    if not ok or (bn.training and bn.momentum is None):    # (cumulative moving average: left to torch)
        bn_f = getattr(bn, "_b200_unfused_forward", bn)   # modules prepared by fuse_batch_norm(): their own forwards
        act_f = getattr(act, "_b200_unfused_forward", act)
        return act_f(bn_f(x)) if residual is None else act_f(bn_f(x) + residual)
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


# ---- the same fusion for model code that calls the two modules itself (the reference's brevitas_examples models) -------
class _PendingBatchNorm:
    """What ``bn(x)`` -- or ``bn(x) + residual`` -- WILL be, handed from a fusing batch-norm to whatever consumes it.

    A quantized activation layer prepared by :func:`fuse_batch_norm` takes it apart and runs the fused operator.  Anything
    else gets the real tensor: torch functions through ``__torch_function__``, attribute access and arithmetic through the
    methods below (the batch-norm then runs once, unfused, exactly as the model code asked)."""
    __slots__ = ("bn", "x", "residual", "_value")

    def __init__(self, bn, x, residual=None):
        self.bn, self.x, self.residual, self._value = bn, x, residual, None

    def materialize(self) -> torch.Tensor:
        if self._value is None:
            v = self.bn._b200_unfused_forward(self.x)
            self._value = v if self.residual is None else v + self.residual
        return self._value

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        from torch.utils._pytree import tree_map
        real = lambda a: a.materialize() if isinstance(a, _PendingBatchNorm) else a      # noqa: E731
        return func(*tree_map(real, args), **tree_map(real, kwargs or {}))

    def _plus(self, other):
        if isinstance(other, _PendingBatchNorm):         # e.g. the batch-norm that closes a ResNet downsample path
            other = other.materialize()
        if (self._value is None and self.residual is None and isinstance(other, torch.Tensor)
                and other.shape == self.x.shape):
            return _PendingBatchNorm(self.bn, self.x, other)         # relu(bn(x) + identity): the residual flavour
        return self.materialize() + other

    __add__ = __radd__ = __iadd__ = _plus

    def __sub__(self, other): return self.materialize() - other
    def __rsub__(self, other): return other - self.materialize()
    def __mul__(self, other): return self.materialize() * other
    __rmul__ = __mul__
    def __truediv__(self, other): return self.materialize() / other
    def __neg__(self): return -self.materialize()
    def __getitem__(self, item): return self.materialize()[item]
    def __len__(self): return self.x.shape[0]

    def __getattr__(self, name):            # .shape, .view(...), .mean(...): the tensor's
        return getattr(self.materialize(), name)


def fuse_batch_norm(model: nn.Module) -> int:
    """Prepare ``model`` -- any model code on ``brevitas.nn`` / ``brevitas_b200.nn`` layers -- so that every
    ``BatchNorm2d`` whose output goes straight (or through one ``+ identity``) into a quantized ReLU / identity layer runs as
    the fused operator above, without touching the model's ``forward``.  Returns the number of batch-norms prepared; undo
    with :func:`unfuse_batch_norm`.  State dict, parameters, hooks on other modules and the results of every unfused path
    are unchanged."""
    import types
    count = 0
    for m in model.modules():
        if hasattr(m, "_b200_unfused_forward"):
            continue
        if type(m) is nn.BatchNorm2d:
            m._b200_unfused_forward = m.forward

            def bn_forward(self, x):
                if isinstance(x, torch.Tensor) and K.bn_act_quant_supported(x):
                    return _PendingBatchNorm(self, x)
                return self._b200_unfused_forward(x)
            m.forward = types.MethodType(bn_forward, m)
            count += 1
        elif _tensor_quant_of(m)[2] is not None:
            m._b200_unfused_forward = m.forward

            def act_forward(self, inp):
                if isinstance(inp, _PendingBatchNorm):
                    if inp._value is None:
                        return bn_act_quant(inp.bn, self, inp.x, inp.residual)
                    inp = inp.materialize()
                return self._b200_unfused_forward(inp)
            m.forward = types.MethodType(act_forward, m)
    return count


def unfuse_batch_norm(model: nn.Module) -> None:
    for m in model.modules():
        if "_b200_unfused_forward" in m.__dict__:
            del m.__dict__["forward"]
            del m.__dict__["_b200_unfused_forward"]
