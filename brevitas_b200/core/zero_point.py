"""Mirror of src/brevitas/core/zero_point.py: the symmetric zero zero-point (:27-35) and the asymmetric families
(:38-54 ``_ScaleShiftZeroPoint``, :57-82 ``StatsFromParameterZeroPoint``, :85-186 ``ParameterFromRuntimeZeroPoint``,
:189-226 ``ParameterZeroPoint``).  Zero-points are tiny tensors (0-dim or one value per channel): they are computed
with the literal reference op sequence on the STE ops; the big-tensor work stays in the fused ``int_quant`` kernels,
which take the resulting zero-point."""
from typing import List, Optional, Tuple, Union

import torch
from torch import Tensor, nn
from torch.nn import Parameter

from ..function.ops_ste import abs_binary_sign_grad
from .stats import DEFAULT_MOMENTUM, IGNORE_MISSING_KEYS, SCALAR_SHAPE, _ParameterListStats
from .utils import StatelessBuffer, inplace_momentum_update, inplace_tensor_add

__all__ = ['ZeroZeroPoint', 'StatsFromParameterZeroPoint', 'ParameterFromRuntimeZeroPoint', 'ParameterZeroPoint']


class ZeroZeroPoint(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.zero_point = StatelessBuffer(torch.tensor(0.0))

    def forward(self, x: Tensor, scale: Tensor, bit_width: Tensor) -> Tensor:
        return self.zero_point()


class _ScaleShiftZeroPoint(nn.Module):
    """``zero_point / scale + min_int``, optionally rounded and clamped to the integer range (zero_point.py:38-54)."""

    def __init__(self, int_quant: nn.Module, quantize_zero_point: bool) -> None:
        super().__init__()
        self.int_quant = int_quant
        self.quantize_zero_point = quantize_zero_point

    def forward(self, zero_point: Tensor, scale: Tensor, bit_width: Tensor) -> Tensor:
        min_int = self.int_quant.min_int(bit_width)
        if self.quantize_zero_point:
            out = self.int_quant.to_int(scale, min_int, bit_width, zero_point)
        else:
            out = zero_point / scale + min_int
        return out


class StatsFromParameterZeroPoint(nn.Module):
    """Zero-point from a statistic of the tracked parameters, e.g. ``-min(w)`` (zero_point.py:57-82)."""

    def __init__(self, int_quant: nn.Module, quantize_zero_point: bool, zero_point_stats_input_view_shape_impl: nn.Module,
                 zero_point_stats_input_concat_dim: int, zero_point_stats_impl: nn.Module,
                 zero_point_shape: Tuple[int, ...], tracked_parameter_list: List[Parameter]) -> None:
        super().__init__()
        self.parameter_list_stats = _ParameterListStats(
            zero_point_stats_impl, zero_point_shape, zero_point_stats_input_view_shape_impl,
            zero_point_stats_input_concat_dim, tracked_parameter_list)
        self.scale_shift_zero_point = _ScaleShiftZeroPoint(int_quant, quantize_zero_point)

    def forward(self, x: Tensor, scale: Tensor, bit_width: Tensor) -> Tensor:
        stats = self.parameter_list_stats()
        return self.scale_shift_zero_point(-stats, scale, bit_width)


class ParameterFromRuntimeZeroPoint(nn.Module):
    """Collect a runtime statistic for ``collect_stats_steps`` steps, then learn it (zero_point.py:85-186)."""

    def __init__(self, collect_stats_steps: int, int_quant: nn.Module, quantize_zero_point: bool,
                 zero_point_stats_impl: nn.Module, zero_point_shape: Tuple[int, ...],
                 zero_point_stats_input_view_shape_impl: nn.Module,
                 zero_point_stats_momentum: Optional[float] = DEFAULT_MOMENTUM) -> None:
        super().__init__()
        assert collect_stats_steps > 0, 'Steps should be more than 0'
        self.collect_stats_steps = collect_stats_steps
        self.counter: int = 0
        self.zero_point_shape = zero_point_shape
        self.stats_input_view_shape_impl = zero_point_stats_input_view_shape_impl
        self.momentum = zero_point_stats_momentum
        self.value = Parameter(torch.full(zero_point_shape, 0.0))
        self.register_buffer('buffer', torch.full(zero_point_shape, 0.0))
        self.zero_point_stats_impl = zero_point_stats_impl
        self.scale_shift_zero_point = _ScaleShiftZeroPoint(int_quant, quantize_zero_point)

    def training_forward(self, x: Tensor) -> Tensor:
        if self.counter < self.collect_stats_steps:
            stats_input = self.stats_input_view_shape_impl(x)
            stats = self.zero_point_stats_impl(stats_input)
            stats = stats.view(self.zero_point_shape)
            new_counter = self.counter + 1
            if self.counter == 0:
                inplace_tensor_add(self.buffer, stats.detach())
            else:
                inplace_momentum_update(self.buffer, stats.detach(), self.momentum, self.counter, new_counter)
            self.counter = new_counter
            out = stats + 0. * self.value
        elif self.counter == self.collect_stats_steps:
            inplace_tensor_add(self.value.detach(), self.buffer)
            self.counter = self.counter + 1
            out = self.value
        else:
            out = self.value
        return out

    def forward(self, x: Tensor, scale: Tensor, bit_width: Tensor) -> Tensor:
        if self.training:
            out = self.training_forward(x)
        else:
            out = self.buffer if self.counter <= self.collect_stats_steps else self.value
        out = abs_binary_sign_grad(out)
        return self.scale_shift_zero_point(out, scale, bit_width)

    def state_dict(self, *args, destination=None, prefix='', keep_vars=False):
        out = super().state_dict(*args, destination=destination, prefix=prefix, keep_vars=keep_vars)
        del out[prefix + 'buffer']
        if self.counter == 0:
            del out[prefix + 'value']
        elif self.counter <= self.collect_stats_steps:
            out[prefix + 'value'] = self.buffer
        return out

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)
        value_key, buffer_key, training_key = prefix + 'value', prefix + 'buffer', prefix + 'training'
        if buffer_key in missing_keys:
            missing_keys.remove(buffer_key)
        if training_key in missing_keys:
            missing_keys.remove(training_key)
        if value_key not in missing_keys:
            self.counter = self.collect_stats_steps + 1
        if IGNORE_MISSING_KEYS and value_key in missing_keys:
            missing_keys.remove(value_key)


class ParameterZeroPoint(nn.Module):
    """Learned zero-point (zero_point.py:189-226)."""

    def __init__(self, zero_point_init: Union[float, Tensor], int_quant: nn.Module, quantize_zero_point: bool,
                 zero_point_shape: Optional[Tuple[int, ...]] = None) -> None:
        super().__init__()
        if (isinstance(zero_point_init, Tensor) and zero_point_shape is not None
                and zero_point_init.shape != SCALAR_SHAPE and zero_point_init.shape != zero_point_shape):
            raise RuntimeError("zero_point_init.shape is non-scalar and != from zero_point_shape.")
        if isinstance(zero_point_init, Tensor):
            zero_point_init = zero_point_init.detach()
        else:
            zero_point_init = torch.tensor(zero_point_init)
        if zero_point_init.shape == SCALAR_SHAPE and zero_point_shape is not None:
            zero_point_init = torch.full(zero_point_shape, zero_point_init)
        self.value = Parameter(zero_point_init)
        self.scale_shift_zero_point = _ScaleShiftZeroPoint(int_quant, quantize_zero_point)

    def forward(self, x: Tensor, scale: Tensor, bit_width: Tensor) -> Tensor:
        out = abs_binary_sign_grad(self.value)
        return self.scale_shift_zero_point(out, scale, bit_width)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)
        value_key = prefix + 'value'
        if IGNORE_MISSING_KEYS and value_key in missing_keys:
            missing_keys.remove(value_key)
