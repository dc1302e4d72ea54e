"""Quantized layers for the BASELINE workloads: ``QuantLinear``, ``QuantConv2d``, ``QuantReLU``, ``QuantIdentity``,
``QuantHardTanh`` -- the callers either side of the hot path, mirroring ``brevitas.nn``.

Same constructor conventions as the reference (src/brevitas/nn/quant_linear.py:22-66, quant_conv.py:116-174,
quant_activation.py:14-101): quantizers are passed as classes (``weight_quant=Int8WeightPerChannelFloat``) and
refined by prefixed keyword arguments (``weight_bit_width=4``; ``bit_width=4`` for activation layers,
nn/mixin/base.py:64-68).  Same sub-module names as the reference so checkpoints line up:
``<layer>.weight_quant.tensor_quant....`` and ``<layer>.act_quant.fused_activation_quant_proxy.tensor_quant....``
(proxy/parameter_quant.py:83-89, proxy/runtime_quant.py:73-84).  The matmul / convolution itself is the stock
cuDNN / cuBLAS call exactly as in the reference (`F.linear`, `F.conv2d`): it is not part of the fake-quant path.

Not covered here (SURVEY.md §8f "next"): bias / accumulator quantizers that need the input scale, QuantTensor
arithmetic, export handlers.  ``return_quant_tensor`` returns a light ``QuantTensor`` tuple without the reference's
operator overloading.
"""
from typing import NamedTuple, Optional, Type, Union

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from .quant import (ActQuantizer, Int8ActPerTensorFloat, Int8WeightPerTensorFloat, Uint8ActPerTensorFloat,
                    WeightQuantizer)


class QuantTensor(NamedTuple):
    """value + quantization metadata (src/brevitas/quant_tensor/__init__.py:18-24), container only"""
    value: Tensor
    scale: Optional[Tensor] = None
    zero_point: Optional[Tensor] = None
    bit_width: Optional[Tensor] = None
    signed: Optional[bool] = None
    training: Optional[bool] = None


def _filter(prefix: str, kwargs: dict) -> dict:
    """``weight_bit_width=4`` -> ``{'bit_width': 4}`` (src/brevitas/nn/mixin/base.py:64-68)"""
    return {k[len(prefix):]: v for k, v in kwargs.items() if k.startswith(prefix)}


def _unpack(x: Union[Tensor, QuantTensor]) -> Tensor:
    return x.value if isinstance(x, QuantTensor) else x


class WeightQuantProxy(nn.Module):
    """proxy/parameter_quant.py:65-89: owns ``tensor_quant`` for one layer's weight"""

    def __init__(self, quantizer: Optional[Type[WeightQuantizer]], weight: nn.Parameter):
        super().__init__()
        self.tensor_quant = quantizer.tensor_quant(weight) if quantizer is not None else None
        self.signed = quantizer.signed if quantizer is not None else None

    @property
    def is_quant_enabled(self):
        return self.tensor_quant is not None

    def forward(self, w: Tensor) -> QuantTensor:
        if self.tensor_quant is None:
            return QuantTensor(w)
        out, scale, zero_point, bit_width = self.tensor_quant(w)
        return QuantTensor(out, scale, zero_point, bit_width, self.signed, self.training)


class FusedActivationQuantProxy(nn.Module):
    """activation then quantizer (proxy/runtime_quant.py:73-84)"""

    def __init__(self, activation_impl: Optional[nn.Module], tensor_quant: Optional[nn.Module]):
        super().__init__()
        self.activation_impl = activation_impl if activation_impl is not None else nn.Identity()
        self.tensor_quant = tensor_quant

    def forward(self, x):
        x = self.activation_impl(x)
        if self.tensor_quant is None:
            return x, None, None, None
        return self.tensor_quant(x)


class ActQuantProxy(nn.Module):
    """proxy/runtime_quant.py:87-164"""

    def __init__(self, quantizer: Optional[Type[ActQuantizer]], act_impl: Optional[nn.Module]):
        super().__init__()
        tq = quantizer.tensor_quant() if quantizer is not None else None
        self.signed = quantizer.signed if quantizer is not None else None
        self.fused_activation_quant_proxy = FusedActivationQuantProxy(act_impl, tq)

    @property
    def is_quant_enabled(self):
        return self.fused_activation_quant_proxy.tensor_quant is not None

    def forward(self, x: Tensor) -> QuantTensor:
        out, scale, zero_point, bit_width = self.fused_activation_quant_proxy(x)
        return QuantTensor(out, scale, zero_point, bit_width, self.signed, self.training)


class _QuantActLayer(nn.Module):
    """QuantNonLinearActLayer (src/brevitas/nn/quant_layer.py:100-150)"""

    def __init__(self, act_impl, act_quant, input_quant=None, return_quant_tensor=False, **kwargs):
        super().__init__()
        self.return_quant_tensor = return_quant_tensor
        aq = act_quant.let(**{k: v for k, v in kwargs.items() if not k.startswith("input_")}) if act_quant else None
        iq = input_quant.let(**_filter("input_", kwargs)) if input_quant else None
        self.input_quant = ActQuantProxy(iq, None)
        self.act_quant = ActQuantProxy(aq, act_impl)

    def forward(self, x):
        x = _unpack(x)
        if self.input_quant.is_quant_enabled:
            x = self.input_quant(x).value
        out = self.act_quant(x)
        return out if self.return_quant_tensor else out.value


class QuantReLU(_QuantActLayer):
    """nn/quant_activation.py:14-31 (default quantizer Uint8ActPerTensorFloat)"""

    def __init__(self, act_quant=Uint8ActPerTensorFloat, input_quant=None, return_quant_tensor=False, **kwargs):
        super().__init__(nn.ReLU(), act_quant, input_quant, return_quant_tensor, **kwargs)


class QuantIdentity(_QuantActLayer):
    """nn/quant_activation.py:82-101 (default quantizer Int8ActPerTensorFloat)"""

    def __init__(self, act_quant=Int8ActPerTensorFloat, return_quant_tensor=False, **kwargs):
        super().__init__(None, act_quant, None, return_quant_tensor, **kwargs)


class QuantHardTanh(_QuantActLayer):
    """nn/quant_activation.py:58-79: requires min_val / max_val for the scale init (as in the reference)"""

    def __init__(self, act_quant=Int8ActPerTensorFloat, input_quant=None, return_quant_tensor=False, **kwargs):
        act = nn.Hardtanh(kwargs.get("min_val", -1.0), kwargs.get("max_val", 1.0))
        super().__init__(act, act_quant, input_quant, return_quant_tensor, **kwargs)


class _QuantWBIOL:
    """the weight / input / output quantizer plumbing shared by Linear and Conv (nn/quant_layer.py:250-365)"""

    def _init_quant(self, weight_quant, input_quant, output_quant, return_quant_tensor, kwargs):
        self.return_quant_tensor = return_quant_tensor
        wq = weight_quant.let(**_filter("weight_", kwargs)) if weight_quant else None
        iq = input_quant.let(**_filter("input_", kwargs)) if input_quant else None
        oq = output_quant.let(**_filter("output_", kwargs)) if output_quant else None
        self.weight_quant = WeightQuantProxy(wq, self.weight)
        self.input_quant = ActQuantProxy(iq, None)
        self.output_quant = ActQuantProxy(oq, None)

    def quant_weight(self) -> QuantTensor:
        return self.weight_quant(self.weight)

    def _forward(self, x, inner):
        x = _unpack(x)
        if self.input_quant.is_quant_enabled:
            x = self.input_quant(x).value
        w = self.quant_weight().value
        out = inner(x, w, self.bias)
        if self.output_quant.is_quant_enabled:
            q = self.output_quant(out)
            return q if self.return_quant_tensor else q.value
        return QuantTensor(out) if self.return_quant_tensor else out


class QuantLinear(_QuantWBIOL, nn.Linear):
    """nn/quant_linear.py:22-66 (default weight quantizer Int8WeightPerTensorFloat)"""

    def __init__(self, in_features, out_features, bias=True, weight_quant=Int8WeightPerTensorFloat, bias_quant=None,
                 input_quant=None, output_quant=None, return_quant_tensor=False, **kwargs):
        nn.Linear.__init__(self, in_features, out_features, bias)
        if bias_quant is not None:
            raise NotImplementedError("bias quantizers need the input scale (SURVEY.md §8f rank 2)")
        self._init_quant(weight_quant, input_quant, output_quant, return_quant_tensor, kwargs)

    def forward(self, x):
        return self._forward(x, F.linear)


class QuantConv2d(_QuantWBIOL, nn.Conv2d):
    """nn/quant_conv.py:116-174 (default weight quantizer Int8WeightPerTensorFloat)"""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 weight_quant=Int8WeightPerTensorFloat, bias_quant=None, input_quant=None, output_quant=None,
                 return_quant_tensor=False, **kwargs):
        nn.Conv2d.__init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        if bias_quant is not None:
            raise NotImplementedError("bias quantizers need the input scale (SURVEY.md §8f rank 2)")
        self._init_quant(weight_quant, input_quant, output_quant, return_quant_tensor, kwargs)

    def forward(self, x):
        return self._forward(
            x, lambda a, w, b: F.conv2d(a, w, b, self.stride, self.padding, self.dilation, self.groups))
