"""Quantized layers for the BASELINE workloads: ``QuantLinear``, ``QuantConv2d``, ``QuantReLU``, ``QuantIdentity``,
``QuantHardTanh`` -- the callers either side of the hot path, mirroring ``brevitas.nn``.

Same constructor conventions as the reference (src/brevitas/nn/quant_linear.py:22-66, quant_conv.py:116-174,
quant_activation.py:14-101): quantizers are passed as classes (``weight_quant=Int8WeightPerChannelFloat``) and
refined by prefixed keyword arguments (``weight_bit_width=4``; ``bit_width=4`` for activation layers,
nn/mixin/base.py:64-68).  Same sub-module names as the reference so checkpoints line up:
``<layer>.weight_quant.tensor_quant....`` and ``<layer>.act_quant.fused_activation_quant_proxy.tensor_quant....``
(proxy/parameter_quant.py:83-89, proxy/runtime_quant.py:73-84).  The matmul / convolution itself is the stock
cuDNN / cuBLAS call exactly as in the reference (`F.linear`, `F.conv2d`): it is not part of the fake-quant path.

Bias quantizers that take the accumulator scale / bit-width (``IntBias`` ...; nn/quant_layer.py:302-365) and the
truncating average pool (``QuantAvgPool2d``, nn/quant_avg_pool.py:21-73) are covered; QuantTensor arithmetic and export
handlers are not (SURVEY.md §8 out of scope).  ``QuantTensor`` is a light tuple with ``set`` / ``view`` / ``reshape`` /
``flatten`` but without the reference's operator overloading.
"""
from typing import NamedTuple, Optional, Type, Union

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from .function.ops import max_int
from .function.ops_ste import ceil_ste
from .quant import (ActQuantizer, BiasQuantizer, Int8ActPerTensorFloat, Int8WeightPerTensorFloat, TruncQuantizer,
                    TruncTo8bit, Uint8ActPerTensorFloat, WeightQuantizer)


class QuantTensor(NamedTuple):
    """value + quantization metadata (src/brevitas/quant_tensor/__init__.py:18-24), container only"""
    value: Tensor
    scale: Optional[Tensor] = None
    zero_point: Optional[Tensor] = None
    bit_width: Optional[Tensor] = None
    signed: Optional[bool] = None
    training: Optional[bool] = None

    @property
    def is_not_none(self):
        return (self.value is not None and self.scale is not None and self.zero_point is not None
                and self.bit_width is not None and self.signed is not None)

    def set(self, **kwargs):
        return self._replace(**kwargs)

    def view(self, *args, **kwargs):
        return self.set(value=self.value.view(*args, **kwargs))

    def reshape(self, *args, **kwargs):
        return self.set(value=self.value.reshape(*args, **kwargs))

    def flatten(self, *args, **kwargs):
        return self.set(value=self.value.flatten(*args, **kwargs))

    def size(self, *args, **kwargs):
        return self.value.size(*args, **kwargs)

    def int(self, float_datatype: bool = False) -> Tensor:
        """integer representation ``round(value / scale + zero_point)`` (quant_tensor/__init__.py:174-187): int8 for
        signed <= 8 bits, uint8 for unsigned <= 8 bits, int32 otherwise -- one export kernel (1 read + 1 byte write)
        instead of div, add, round, cast.  A tensor-valued non-zero zero-point takes the literal sequence."""
        if not self.is_not_none:
            raise RuntimeError("QuantTensor not valid.")
        narrow = float(self.bit_width) <= 8.0
        dtype = (torch.int8 if self.signed else torch.uint8) if narrow else torch.int32
        scalar_zp = self.zero_point.numel() == 1
        if float_datatype or not self.value.is_cuda or not scalar_zp:
            from .core.quant import scalar_div
            from .function.ops_ste import round_ste
            int_value = round_ste(scalar_div(self.value, self.scale) + self.zero_point)
            return int_value if float_datatype else int_value.to(dtype)
        return torch.ops.brevitas_b200.int_quant_to_int(self.value.detach(), self.scale.detach(), float(self.zero_point),
                                                        None, None, 0, dtype)

    @property
    def shape(self):
        return self.value.shape


def _filter(prefix: str, kwargs: dict) -> dict:
    """``weight_bit_width=4`` -> ``{'bit_width': 4}`` (src/brevitas/nn/mixin/base.py:64-68)"""
    return {k[len(prefix):]: v for k, v in kwargs.items() if k.startswith(prefix)}


def _unpack(x: Union[Tensor, QuantTensor]) -> Tensor:
    return x.value if isinstance(x, QuantTensor) else x


def _as_quant_tensor(x: Union[Tensor, QuantTensor], training: bool) -> QuantTensor:
    return x if isinstance(x, QuantTensor) else QuantTensor(x, training=training)


class WeightQuantProxy(nn.Module):
    """proxy/parameter_quant.py:65-89: owns ``tensor_quant`` for one layer's weight"""

    def __init__(self, quantizer: Optional[Type[WeightQuantizer]], weight: nn.Parameter, output_channel_dim: int = 0):
        super().__init__()
        self.tensor_quant = quantizer.tensor_quant(weight, output_channel_dim) if quantizer is not None else None
        self.signed = quantizer.signed if quantizer is not None else None
        self.is_narrow_range = bool(quantizer.narrow_range) if quantizer is not None else False

    @property
    def is_quant_enabled(self):
        return self.tensor_quant is not None

    def max_uint_value(self, bit_width):
        """proxy/parameter_quant.py:61-62"""
        return max_int(False, self.is_narrow_range, bit_width)

    def forward(self, w: Tensor) -> QuantTensor:
        if self.tensor_quant is None:
            return QuantTensor(w)
        out, scale, zero_point, bit_width = self.tensor_quant(w)
        return QuantTensor(out, scale, zero_point, bit_width, self.signed, self.training)


class BiasQuantProxy(nn.Module):
    """proxy/parameter_quant.py:110-176: the bias quantizer may need the accumulator's scale and bit-width"""

    def __init__(self, quantizer: Optional[Type[BiasQuantizer]]):
        super().__init__()
        self.tensor_quant = quantizer.tensor_quant() if quantizer is not None else None
        self.signed = quantizer.signed if quantizer is not None else None
        self.requires_input_scale = bool(quantizer.requires_input_scale) if quantizer is not None else False
        self.requires_input_bit_width = bool(quantizer.requires_input_bit_width) if quantizer is not None else False

    @property
    def is_quant_enabled(self):
        return self.tensor_quant is not None

    def forward(self, x: Tensor, input_scale: Optional[Tensor] = None, input_bit_width: Optional[Tensor] = None):
        if self.tensor_quant is None:
            return QuantTensor(x, training=self.training)
        if self.requires_input_scale and input_scale is None:
            raise RuntimeError("Input scale required")
        if self.requires_input_bit_width and input_bit_width is None:
            raise RuntimeError("Input bit-width required")
        if self.requires_input_scale and self.requires_input_bit_width:
            out = self.tensor_quant(x, input_scale.view(-1), input_bit_width)
        elif self.requires_input_scale:
            out = self.tensor_quant(x, input_scale.view(-1))
        elif not self.requires_input_bit_width:
            out = self.tensor_quant(x)
        else:
            raise RuntimeError("Internally defined bit-width required")
        return QuantTensor(out[0], out[1], out[2], out[3], self.signed, self.training)


class TruncQuantProxy(nn.Module):
    """proxy/runtime_quant.py:178-198"""

    def __init__(self, quantizer: Optional[Type[TruncQuantizer]]):
        super().__init__()
        self.tensor_quant = quantizer.tensor_quant() if quantizer is not None else None

    @property
    def is_quant_enabled(self):
        return self.tensor_quant is not None

    def forward(self, x: QuantTensor) -> QuantTensor:
        if self.tensor_quant is None:
            return x
        v, s, zp, bw = self.tensor_quant(x.value, x.scale, x.zero_point, x.bit_width)
        return QuantTensor(v, s, zp, bw, x.signed, self.training)


class FusedActivationQuantProxy(nn.Module):
    """activation then quantizer (proxy/runtime_quant.py:73-84)"""

    def __init__(self, activation_impl: Optional[nn.Module], tensor_quant: Optional[nn.Module]):
        super().__init__()
        self.activation_impl = activation_impl if activation_impl is not None else nn.Identity()
        self.tensor_quant = tensor_quant

    def forward(self, x):
        if type(self.activation_impl) is nn.ReLU and self.tensor_quant is not None:
            fused = getattr(self.tensor_quant, 'forward_pre_relu', None)
            out = fused(x) if fused is not None else None
            if out is not None:                 # ReLU folded into the quantizer kernel (forward and backward)
                return out
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


class QuantSigmoid(_QuantActLayer):
    """nn/quant_activation.py:32-47 (default quantizer Uint8ActPerTensorFloat)"""

    def __init__(self, act_quant=Uint8ActPerTensorFloat, input_quant=None, return_quant_tensor=False, **kwargs):
        super().__init__(nn.Sigmoid(), act_quant, input_quant, return_quant_tensor, **kwargs)


class QuantTanh(_QuantActLayer):
    """nn/quant_activation.py:50-65 (default quantizer Int8ActPerTensorFloat)"""

    def __init__(self, act_quant=Int8ActPerTensorFloat, input_quant=None, return_quant_tensor=False, **kwargs):
        super().__init__(nn.Tanh(), act_quant, input_quant, return_quant_tensor, **kwargs)


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
    """the weight / bias / input / output quantizer plumbing shared by Linear and Conv (nn/quant_layer.py:250-365)"""

    def _init_quant(self, weight_quant, bias_quant, input_quant, output_quant, return_quant_tensor, kwargs):
        self.return_quant_tensor = return_quant_tensor
        wq = weight_quant.let(**_filter("weight_", kwargs)) if weight_quant else None
        bq = bias_quant.let(**_filter("bias_", kwargs)) if bias_quant else None
        iq = input_quant.let(**_filter("input_", kwargs)) if input_quant else None
        oq = output_quant.let(**_filter("output_", kwargs)) if output_quant else None
        self.weight_quant = WeightQuantProxy(wq, self.weight, getattr(self, "output_channel_dim", 0))
        self.bias_quant = BiasQuantProxy(bq)
        self.input_quant = ActQuantProxy(iq, None)
        self.output_quant = ActQuantProxy(oq, None)

    def quant_weight(self) -> QuantTensor:
        return self.weight_quant(self.weight)

    def _forward(self, x, inner):
        """forward_impl of the reference (nn/quant_layer.py:302-365): accumulator scale = weight scale x input scale,
        accumulator bit-width from max_acc_bit_width, optional bias quantization against them"""
        inp = _as_quant_tensor(x, self.training)
        quant_input = self.input_quant(inp.value) if self.input_quant.is_quant_enabled else inp
        quant_weight = self.quant_weight()
        output_scale = output_bit_width = output_zero_point = output_signed = None
        if quant_input.bit_width is not None and quant_weight.bit_width is not None:
            output_bit_width = self.max_acc_bit_width(quant_input.bit_width, quant_weight.bit_width)
        if quant_input.scale is not None and quant_weight.scale is not None:
            shape = (1, -1) + (1,) * (quant_input.value.dim() - 2)            # compute_channel_view_shape(inp, 1)
            output_scale = quant_weight.scale.view(shape) * quant_input.scale.view(shape)
        if quant_input.signed is not None:
            output_signed = bool(quant_input.signed or quant_weight.signed)
        if self.bias is not None:
            quant_bias = self.bias_quant(self.bias, output_scale, output_bit_width)
            out = inner(quant_input.value, quant_weight.value, quant_bias.value)
            # a bias that is NOT quantized at the accumulator scale (a float bias, or a bias quantizer with its own
            # scale) shifts the accumulator: it is carried as the output zero-point (quant_layer.py:337-341)
            if output_scale is not None and (quant_bias.scale is None
                                             or quant_bias.scale.data_ptr() != output_scale.data_ptr()):
                output_zero_point = - quant_bias.value.view(shape) / output_scale
            if quant_bias.bit_width is not None and output_bit_width is not None:
                output_bit_width = torch.where(quant_bias.bit_width > output_bit_width, quant_bias.bit_width,
                                               output_bit_width) + 1
        else:
            out = inner(quant_input.value, quant_weight.value, None)
        if self.return_quant_tensor and not self.output_quant.is_quant_enabled:
            # quant_layer.py:349-355 (the reference evaluates the two .any() on the host here, under the same condition)
            checkable = not (quant_input.value.is_cuda and torch.cuda.is_current_stream_capturing())   # a host read
            if checkable and quant_input.zero_point is not None and ((quant_input.zero_point != 0.0).any()
                                                                     or (quant_weight.zero_point != 0.0).any()):
                raise RuntimeError("Computing zero point of output accumulator not supported yet.")
            elif quant_input.zero_point is not None and output_zero_point is None:
                output_zero_point = quant_input.zero_point
        if self.output_quant.is_quant_enabled:
            q = self.output_quant(out)
            return q if self.return_quant_tensor else q.value
        if self.return_quant_tensor:
            return QuantTensor(out, output_scale, output_zero_point, output_bit_width, output_signed, self.training)
        return out


class QuantLinear(_QuantWBIOL, nn.Linear):
    """nn/quant_linear.py:22-66 (default weight quantizer Int8WeightPerTensorFloat)"""

    def __init__(self, in_features, out_features, bias=True, weight_quant=Int8WeightPerTensorFloat, bias_quant=None,
                 input_quant=None, output_quant=None, return_quant_tensor=False, **kwargs):
        nn.Linear.__init__(self, in_features, out_features, bias)
        self._init_quant(weight_quant, bias_quant, input_quant, output_quant, return_quant_tensor, kwargs)

    def max_acc_bit_width(self, input_bit_width, weight_bit_width):
        """nn/quant_linear.py:68-73"""
        max_input_val = max_int(bit_width=input_bit_width, signed=False, narrow_range=False)
        max_fc_val = self.weight_quant.max_uint_value(weight_bit_width)
        return ceil_ste(torch.log2(max_input_val * max_fc_val * self.in_features))

    def forward(self, x):
        return self._forward(x, F.linear)


class QuantConv2d(_QuantWBIOL, nn.Conv2d):
    """nn/quant_conv.py:116-174 (default weight quantizer Int8WeightPerTensorFloat)"""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 weight_quant=Int8WeightPerTensorFloat, bias_quant=None, input_quant=None, output_quant=None,
                 return_quant_tensor=False, **kwargs):
        nn.Conv2d.__init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        self._init_quant(weight_quant, bias_quant, input_quant, output_quant, return_quant_tensor, kwargs)

    def max_acc_bit_width(self, input_bit_width, weight_bit_width):
        """nn/quant_conv.py:199-206"""
        max_uint_input = max_int(bit_width=input_bit_width, signed=False, narrow_range=False)
        max_kernel_val = self.weight_quant.max_uint_value(weight_bit_width)
        group_size = self.out_channels // self.groups
        kernel_size = self.kernel_size[0] * self.kernel_size[1]
        return ceil_ste(torch.log2(max_uint_input * max_kernel_val * kernel_size * group_size))

    def forward(self, x):
        return self._forward(
            x, lambda a, w, b: F.conv2d(a, w, b, self.stride, self.padding, self.dilation, self.groups))


class QuantAvgPool2d(nn.AvgPool2d):
    """nn/quant_avg_pool.py:21-73: average pool on a QuantTensor, the sum rescaled back to integers and its extra
    bits truncated (``TruncTo8bit`` by default; ``bit_width=...`` refines it)"""

    def __init__(self, kernel_size, stride=None, trunc_quant=TruncTo8bit, return_quant_tensor=True, **kwargs):
        nn.AvgPool2d.__init__(self, kernel_size=kernel_size, stride=stride)
        self.return_quant_tensor = return_quant_tensor
        tq = trunc_quant.let(**{k[len("trunc_"):] if k.startswith("trunc_") else k: v for k, v in kwargs.items()}) \
            if trunc_quant else None
        self.trunc_quant = TruncQuantProxy(tq)

    @property
    def _avg_scaling(self):
        if isinstance(self.kernel_size, tuple):
            return self.kernel_size[0] * self.kernel_size[1]
        return self.kernel_size * self.kernel_size

    def max_acc_bit_width(self, input_bit_width):
        max_uint_input = max_int(bit_width=input_bit_width, signed=False, narrow_range=False)
        return ceil_ste(torch.log2(max_uint_input * self._avg_scaling))

    def forward(self, x):
        x = _as_quant_tensor(x, self.training)
        x = x.set(value=nn.AvgPool2d.forward(self, x.value))
        if self.trunc_quant.is_quant_enabled:
            assert x.is_not_none, "QuantAvgPool2d needs a QuantTensor input (value, scale, zero-point, bit-width, sign)"
            x = x.set(value=x.value * self._avg_scaling)           # remove the averaging: back to a sum of integers
            x = x.set(bit_width=self.max_acc_bit_width(x.bit_width))
            x = self.trunc_quant(x)
        return x if self.return_quant_tensor else x.value


class QuantConv1d(_QuantWBIOL, nn.Conv1d):
    """nn/quant_conv.py:22-113"""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 weight_quant=Int8WeightPerTensorFloat, bias_quant=None, input_quant=None, output_quant=None,
                 return_quant_tensor=False, **kwargs):
        nn.Conv1d.__init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        self._init_quant(weight_quant, bias_quant, input_quant, output_quant, return_quant_tensor, kwargs)

    def max_acc_bit_width(self, input_bit_width, weight_bit_width):
        max_uint_input = max_int(bit_width=input_bit_width, signed=False, narrow_range=False)
        max_kernel_val = self.weight_quant.max_uint_value(weight_bit_width)
        group_size = self.out_channels // self.groups
        return ceil_ste(torch.log2(max_uint_input * max_kernel_val * self.kernel_size[0] * group_size))

    def forward(self, x):
        return self._forward(
            x, lambda a, w, b: F.conv1d(a, w, b, self.stride, self.padding, self.dilation, self.groups))


class QuantConvTranspose2d(_QuantWBIOL, nn.ConvTranspose2d):
    """nn/quant_convtranspose.py: output channels live in dim 1 of the weight ``[in, out / groups, kh, kw]``: per-channel
    statistics see a permuted view, the scale has shape ``[1, out, 1, 1]``"""
    output_channel_dim = 1

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, groups=1, bias=True,
                 dilation=1, weight_quant=Int8WeightPerTensorFloat, bias_quant=None, input_quant=None, output_quant=None,
                 return_quant_tensor=False, **kwargs):
        nn.ConvTranspose2d.__init__(self, in_channels, out_channels, kernel_size, stride, padding, output_padding, groups,
                                    bias, dilation)
        self._init_quant(weight_quant, bias_quant, input_quant, output_quant, return_quant_tensor, kwargs)

    def max_acc_bit_width(self, input_bit_width, weight_bit_width):
        """nn/quant_convtranspose.py:193-201: overlapping kernel patches per output position"""
        max_uint_input = max_int(bit_width=input_bit_width, signed=False, narrow_range=False)
        max_kernel_val = self.weight_quant.max_uint_value(weight_bit_width)
        group_size = self.out_channels // self.groups
        overlap = 1
        for k, st in zip(self.kernel_size, self.stride):
            overlap *= max(round(k / st), 1)
        return ceil_ste(torch.log2(max_uint_input * max_kernel_val * overlap * group_size))

    def forward(self, x):
        return self._forward(
            x, lambda a, w, b: F.conv_transpose2d(a, w, b, self.stride, self.padding, self.output_padding, self.groups,
                                                  self.dilation))


class QuantConvTranspose1d(_QuantWBIOL, nn.ConvTranspose1d):
    """nn/quant_convtranspose.py:22-111: weight ``[in, out / groups, k]``, output channels in dim 1"""
    output_channel_dim = 1

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, groups=1, bias=True,
                 dilation=1, weight_quant=Int8WeightPerTensorFloat, bias_quant=None, input_quant=None, output_quant=None,
                 return_quant_tensor=False, **kwargs):
        nn.ConvTranspose1d.__init__(self, in_channels, out_channels, kernel_size, stride, padding, output_padding, groups,
                                    bias, dilation)
        self._init_quant(weight_quant, bias_quant, input_quant, output_quant, return_quant_tensor, kwargs)

    def max_acc_bit_width(self, input_bit_width, weight_bit_width):
        """nn/quant_convtranspose.py:104-111"""
        max_uint_input = max_int(bit_width=input_bit_width, signed=False, narrow_range=False)
        max_kernel_val = self.weight_quant.max_uint_value(weight_bit_width)
        group_size = self.out_channels // self.groups
        overlap = max(round(self.kernel_size[0] / self.stride[0]), 1)
        return ceil_ste(torch.log2(max_uint_input * max_kernel_val * overlap * group_size))

    def forward(self, x):
        return self._forward(
            x, lambda a, w, b: F.conv_transpose1d(a, w, b, self.stride, self.padding, self.output_padding, self.groups,
                                                  self.dilation))
