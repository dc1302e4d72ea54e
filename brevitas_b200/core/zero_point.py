"""Mirror of src/brevitas/core/zero_point.py: the symmetric zero zero-point (:27-35) and the asymmetric families
(:38-54 ``_ScaleShiftZeroPoint``, :57-82 ``StatsFromParameterZeroPoint``, :85-186 ``ParameterFromRuntimeZeroPoint``,
:189-226 ``ParameterZeroPoint``).  Zero-points are tiny tensors (0-dim or one value per channel): they are computed
with the literal reference op sequence on the STE ops; the big-tensor work stays in the fused ``int_quant`` kernels,
which take the resulting zero-point."""
from typing import List, Optional, Tuple, Union

import torch

from .. import config
from torch import Tensor, nn
from torch.nn import Parameter

from ..function.ops_ste import abs_binary_sign_grad
from .stats import DEFAULT_MOMENTUM, SCALAR_SHAPE, _ParameterListStats
from .utils import StatelessBuffer, inplace_momentum_update, inplace_tensor_add

__all__ = ['ZeroZeroPoint', 'StatsFromParameterZeroPoint', 'ParameterFromRuntimeZeroPoint', 'ParameterZeroPoint']


class ZeroZeroPoint(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.zero_point = StatelessBuffer(torch.tensor(0.0))

    def forward(self, x: Tensor, scale: Tensor, bit_width: Tensor) -> Tensor:
        return self.zero_point()


class _ScaleShiftZeroPoint(nn.Module):
    """Turns a real-valued offset into the quantizer's zero-point: ``offset / scale + min_int``, or -- with
    ``quantize_zero_point`` -- that value rounded and clamped by the quantizer's own ``to_int`` so it is a valid
    integer code (zero_point.py:38-54)."""

    def __init__(self, int_quant: nn.Module, quantize_zero_point: bool) -> None:
        super().__init__()
        self.int_quant = int_quant
        self.quantize_zero_point = quantize_zero_point

    def forward(self, zero_point: Tensor, scale: Tensor, bit_width: Tensor) -> Tensor:
        lowest_code = self.int_quant.min_int(bit_width)
        if not self.quantize_zero_point:
            return zero_point / scale + lowest_code
        return self.int_quant.to_int(scale, lowest_code, bit_width, zero_point)


class _OffsetZeroPoint(nn.Module):
    """Shared tail of the asymmetric zero-points: a real-valued offset (statistic, collected buffer or learned
    parameter) goes through ``_ScaleShiftZeroPoint``; missing-key policy for the optional ``value`` parameter."""

    def __init__(self, int_quant: nn.Module, quantize_zero_point: bool) -> None:
        super().__init__()
        self.scale_shift_zero_point = _ScaleShiftZeroPoint(int_quant, quantize_zero_point)

    def _to_codes(self, offset: Tensor, scale: Tensor, bit_width: Tensor) -> Tensor:
        return self.scale_shift_zero_point(offset, scale, bit_width)

    @staticmethod
    def _forgive_missing_value(prefix: str, missing_keys) -> None:
        key = prefix + 'value'
        if config.IGNORE_MISSING_KEYS and key in missing_keys:
            missing_keys.remove(key)


class StatsFromParameterZeroPoint(_OffsetZeroPoint):
    """Offset = minus a statistic of the tracked parameters, e.g. ``-min(w)`` (zero_point.py:57-82)."""

    def __init__(self, int_quant: nn.Module, quantize_zero_point: bool, zero_point_stats_input_view_shape_impl: nn.Module,
                 zero_point_stats_input_concat_dim: int, zero_point_stats_impl: nn.Module,
                 zero_point_shape: Tuple[int, ...], tracked_parameter_list: List[Parameter]) -> None:
        super().__init__(int_quant, quantize_zero_point)
        self.parameter_list_stats = _ParameterListStats(
            zero_point_stats_impl, zero_point_shape, zero_point_stats_input_view_shape_impl,
            zero_point_stats_input_concat_dim, tracked_parameter_list)

    def forward(self, x: Tensor, scale: Tensor, bit_width: Tensor) -> Tensor:
        return self._to_codes(-self.parameter_list_stats(), scale, bit_width)


class ParameterFromRuntimeZeroPoint(_OffsetZeroPoint):
    """A runtime statistic for ``collect_stats_steps`` training steps (kept as a running average in ``buffer``), then
    a learned parameter initialised from it (zero_point.py:85-186).  Three phases, driven by ``counter``:
    collecting (< steps), hand-over (== steps: the buffer is added into ``value`` once), learned (> steps)."""

    def __init__(self, collect_stats_steps: int, int_quant: nn.Module, quantize_zero_point: bool,
                 zero_point_stats_impl: nn.Module, zero_point_shape: Tuple[int, ...],
                 zero_point_stats_input_view_shape_impl: nn.Module,
                 zero_point_stats_momentum: Optional[float] = DEFAULT_MOMENTUM) -> None:
        super().__init__(int_quant, quantize_zero_point)
        assert collect_stats_steps > 0, 'Steps should be more than 0'
        self.collect_stats_steps = collect_stats_steps
        self.counter: int = 0
        self.zero_point_shape = zero_point_shape
        self.momentum = zero_point_stats_momentum
        self.stats_input_view_shape_impl = zero_point_stats_input_view_shape_impl
        self.zero_point_stats_impl = zero_point_stats_impl
        self.value = Parameter(torch.full(zero_point_shape, 0.0))
        self.register_buffer('buffer', torch.full(zero_point_shape, 0.0))

    def _observe(self, x: Tensor) -> Tensor:
        """one collection step: statistic of this batch, folded into the running buffer"""
        stats = self.zero_point_stats_impl(self.stats_input_view_shape_impl(x)).view(self.zero_point_shape)
        seen = self.counter
        if seen == 0:
            inplace_tensor_add(self.buffer, stats.detach())
        else:
            inplace_momentum_update(self.buffer, stats.detach(), self.momentum, seen, seen + 1)
        self.counter = seen + 1
        return stats + 0. * self.value            # keeps `value` in the autograd graph (DDP sees it used)

    def training_forward(self, x) -> Tensor:
        if self.counter < self.collect_stats_steps:
            return self._observe(x)
        if self.counter == self.collect_stats_steps:        # hand-over, exactly once
            inplace_tensor_add(self.value.detach(), self.buffer)
            self.counter += 1
        return self.value

    def forward(self, x: Tensor, scale: Tensor, bit_width: Tensor) -> Tensor:
        if self.training:
            offset = self.training_forward(x)
        elif self.counter <= self.collect_stats_steps:
            offset = self.buffer
        else:
            offset = self.value
        return self._to_codes(abs_binary_sign_grad(offset), scale, bit_width)

    def state_dict(self, *args, destination=None, prefix='', keep_vars=False):
        """the running buffer never appears; before the hand-over it is what gets saved as ``value``"""
        out = super().state_dict(*args, destination=destination, prefix=prefix, keep_vars=keep_vars)
        out.pop(prefix + 'buffer')
        if self.counter == 0:
            out.pop(prefix + 'value')
        elif self.counter <= self.collect_stats_steps:
            out[prefix + 'value'] = self.buffer
        return out

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)
        for never_saved in (prefix + 'buffer', prefix + 'training'):
            if never_saved in missing_keys:
                missing_keys.remove(never_saved)
        if prefix + 'value' not in missing_keys:
            self.counter = self.collect_stats_steps + 1          # a loaded value ends the collection phase
        self._forgive_missing_value(prefix, missing_keys)


class ParameterZeroPoint(_OffsetZeroPoint):
    """Learned offset (zero_point.py:189-226)."""

    def __init__(self, zero_point_init: Union[float, Tensor], int_quant: nn.Module, quantize_zero_point: bool,
                 zero_point_shape: Optional[Tuple[int, ...]] = None) -> None:
        super().__init__(int_quant, quantize_zero_point)
        init = zero_point_init.detach() if isinstance(zero_point_init, Tensor) else torch.tensor(zero_point_init)
        if zero_point_shape is not None:
            if init.shape == SCALAR_SHAPE:
                init = torch.full(zero_point_shape, init)
            elif init.shape != zero_point_shape:
                raise RuntimeError("zero_point_init.shape is non-scalar and != from zero_point_shape.")
        self.value = Parameter(init)

    def forward(self, x: Tensor, scale: Tensor, bit_width: Tensor) -> Tensor:
        return self._to_codes(abs_binary_sign_grad(self.value), scale, bit_width)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)
        self._forgive_missing_value(prefix, missing_keys)
