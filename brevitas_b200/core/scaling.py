"""Mirror of ``brevitas.core.scaling``: the producers of the quantization threshold / scale.

Reference: src/brevitas/core/scaling/standalone.py (ConstScaling :22, ParameterScaling :75,
ParameterFromRuntimeStatsScaling :155), runtime.py (StatsFromParameterScaling :19, _StatsScaling :50,
RuntimeStatsScaling :75, _AffineRescaling :105), int_scaling.py (IntScaling :11, PowerOfTwoIntScaling :27).
Scale tensors are tiny (0-dim .. [channels]); their arithmetic stays on ATen exactly as in the reference,
while the statistics over full tensors run on the sm_100a kernels or are fused into the quant kernel.
"""
from typing import List, Optional, Tuple, Union

import torch

from .. import config
from torch import Tensor, nn
from torch.nn import Parameter

from ..function.ops import max_int, min_int
from ..function.ops_ste import abs_binary_sign_grad
from . import stats as _stats
from .function_wrapper import Identity, OverBatchOverTensorView, ScalarClampMinSte
from .restrict_val import FloatRestrictValue, _ClampValue, _RestrictClampValue, _RestrictValue
from .stats import DEFAULT_MOMENTUM, SCALAR_SHAPE, AbsMaxPlan, _ParameterListStats, _RuntimeStats, _Stats
from .utils import StatelessBuffer, inplace_momentum_update, inplace_tensor_mul


class FusedStatsPlan:
    """What ``RescalingIntQuant`` needs to run statistic + quantization as one kernel."""

    __slots__ = ('geom', 'scaling_min_val', 'out_shape', 'on_absmax')

    def __init__(self, geom: AbsMaxPlan, scaling_min_val: float, out_shape, on_absmax=None):
        self.geom = geom
        self.scaling_min_val = scaling_min_val
        self.out_shape = out_shape
        self.on_absmax = on_absmax


class IntScaling(nn.Module):
    """Integer threshold of the range: ``-min_int`` if signed else ``max_int`` (int_scaling.py:11-24)."""

    def __init__(self, signed: bool, narrow_range: bool):
        super().__init__()
        self.signed = signed
        self.narrow_range = narrow_range

    def forward(self, bit_width: Tensor) -> Tensor:
        if self.signed:
            return - min_int(self.signed, self.narrow_range, bit_width)
        return max_int(self.signed, self.narrow_range, bit_width)


class PowerOfTwoIntScaling(nn.Module):
    """``max_int(signed, False, bw) + 1`` (int_scaling.py:27-36)."""

    def __init__(self, signed: bool):
        super().__init__()
        self.signed = signed

    def forward(self, bit_width: Tensor) -> Tensor:
        return max_int(self.signed, False, bit_width) + 1


class ConstScaling(nn.Module):
    """Constant threshold (standalone.py:22-72)."""

    def __init__(self, scaling_init: Union[float, Tensor], restrict_scaling_impl: Optional[nn.Module] = None,
                 scaling_min_val: Optional[float] = None) -> None:
        super().__init__()
        self.restrict_clamp_scaling = _RestrictClampValue(scaling_min_val, restrict_scaling_impl)
        if isinstance(scaling_init, Tensor):
            if restrict_scaling_impl is not None:
                scaling_init = restrict_scaling_impl.restrict_init_tensor(scaling_init)
            self.value = StatelessBuffer(scaling_init.detach())
        else:
            if restrict_scaling_impl is not None:
                scaling_init = restrict_scaling_impl.restrict_init_float(scaling_init)
            self.value = StatelessBuffer(torch.tensor(scaling_init))

    def forward(self, placeholder: Tensor) -> Tensor:
        return self.restrict_clamp_scaling(self.value())

    def input_independent(self) -> bool:
        """the threshold does not look at the tensor being quantized (lets QuantReLU fuse the ReLU into the kernel)"""
        return True


class ParameterScaling(nn.Module):
    """Learned threshold (standalone.py:75-152)."""

    def __init__(self, scaling_init: Union[float, Tensor], scaling_shape: Optional[Tuple[int, ...]] = None,
                 restrict_scaling_impl: Optional[nn.Module] = None, scaling_min_val: Optional[float] = None) -> None:
        super().__init__()
        if (isinstance(scaling_init, Tensor) and scaling_shape is not None
                and scaling_init.shape != SCALAR_SHAPE and scaling_init.shape != scaling_shape):
            raise RuntimeError("scaling_init.shape is non-scalar and != from scaling_shape.")
        scaling_init = scaling_init.detach() if isinstance(scaling_init, Tensor) else torch.tensor(scaling_init)
        if restrict_scaling_impl is not None:
            scaling_init = restrict_scaling_impl.restrict_init_tensor(scaling_init)
        if scaling_init.shape == SCALAR_SHAPE and scaling_shape is not None:
            scaling_init = torch.full(scaling_shape, scaling_init)
        self.value = Parameter(scaling_init)
        self.restrict_clamp_scaling = _RestrictClampValue(scaling_min_val, restrict_scaling_impl)

    def forward(self, placeholder: Tensor) -> Tensor:
        return abs_binary_sign_grad(self.restrict_clamp_scaling(self.value))

    def input_independent(self) -> bool:
        return True

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        value_key = prefix + 'value'
        retro = prefix + 'learned_value'
        if retro in state_dict:
            state_dict[value_key] = state_dict.pop(retro)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)
        if config.IGNORE_MISSING_KEYS and value_key in missing_keys:
            missing_keys.remove(value_key)


class ParameterFromRuntimeStatsScaling(nn.Module):
    """Statistics for ``collect_stats_steps`` training steps, then a learned parameter (standalone.py:155-298)."""

    def __init__(self, collect_stats_steps: int, scaling_stats_impl: nn.Module,
                 scaling_stats_input_view_shape_impl: Optional[nn.Module] = None,
                 scaling_shape: Tuple[int, ...] = SCALAR_SHAPE, restrict_scaling_impl: Optional[nn.Module] = None,
                 scaling_stats_momentum: Optional[float] = DEFAULT_MOMENTUM,
                 scaling_min_val: Optional[float] = None) -> None:
        super().__init__()
        assert collect_stats_steps > 0, 'Steps should be more than 0'
        self.collect_stats_steps = collect_stats_steps
        self.counter: int = 0
        self.stats_input_view_shape_impl = scaling_stats_input_view_shape_impl or OverBatchOverTensorView()
        self.stats = _Stats(scaling_stats_impl, scaling_shape)
        self.momentum = scaling_stats_momentum
        self.register_buffer('buffer', torch.full(scaling_shape, 1.0))
        self.value = Parameter(torch.full(scaling_shape, 1.0))
        self.restrict_scaling = _RestrictValue(restrict_scaling_impl)
        self.clamp_scaling = _ClampValue(scaling_min_val)
        if restrict_scaling_impl is not None:
            self.restrict_inplace_preprocess = restrict_scaling_impl.restrict_init_inplace_module()
            self.restrict_preprocess = restrict_scaling_impl.restrict_init_module()
        else:
            self.restrict_inplace_preprocess = Identity()
            self.restrict_preprocess = Identity()

    def pre_relu_collecting(self, x: Tensor) -> bool:
        """still collecting, and the statistic of relu(x) can be taken by the ReLU-folded select (whole-tensor
        AbsPercentile over a dense tensor, scalar threshold): the caller may then skip the ReLU pass"""
        from .function_wrapper import OverTensorView
        from .stats import AbsPercentile
        impl = self.stats.stats_impl
        return (self.training and self.counter < self.collect_stats_steps and type(impl) is AbsPercentile
                and type(self.stats_input_view_shape_impl) is OverTensorView and tuple(self.buffer.shape) == ()
                and impl.relu_tensor_supported(x))

    def training_forward(self, stats_input: Tensor, pre_relu: Optional[dict] = None) -> Tensor:
        if self.counter < self.collect_stats_steps:
            if pre_relu is not None:       # statistic of relu(stats_input), ReLU folded into the select kernel
                stats = self.stats.stats_impl.forward_relu_tensor(stats_input, pre_relu).view(self.buffer.shape)
            else:
                stats_input = self.stats_input_view_shape_impl(stats_input)
                stats = self.stats(stats_input)
            stats = stats + 0. * self.value      # keeps `value` in the graph (DDP, standalone.py:234-235)
            clamped_stats = self.clamp_scaling(stats)
            new_counter = self.counter + 1
            if self.counter == 0:
                inplace_tensor_mul(self.buffer, clamped_stats.detach())
            else:
                inplace_momentum_update(self.buffer, clamped_stats.detach(), self.momentum, self.counter, new_counter)
            self.counter = new_counter
            return abs_binary_sign_grad(clamped_stats)
        if self.counter == self.collect_stats_steps:
            self.restrict_inplace_preprocess(self.buffer)
            inplace_tensor_mul(self.value.detach(), self.buffer)
            self.counter = self.counter + 1
        return abs_binary_sign_grad(self.clamp_scaling(self.restrict_scaling(self.value)))

    def input_independent(self) -> bool:
        """true once the collection phase is over (or in eval mode): the statistics input is ignored"""
        return (not self.training) or self.counter >= self.collect_stats_steps

    def forward(self, stats_input: Tensor, pre_relu: Optional[dict] = None) -> Tensor:
        if self.training:
            return self.training_forward(stats_input, pre_relu)
        if self.counter <= self.collect_stats_steps:
            out = self.restrict_preprocess(self.buffer)
        else:
            out = self.value
        return abs_binary_sign_grad(self.clamp_scaling(self.restrict_scaling(out)))

    def state_dict(self, *args, destination=None, prefix='', keep_vars=False):
        out = super().state_dict(*args, destination=destination, prefix=prefix, keep_vars=keep_vars)
        out.pop(prefix + 'buffer', None)
        if self.counter == 0:
            out.pop(prefix + 'value', None)
        elif self.counter <= self.collect_stats_steps:
            out[prefix + 'value'] = self.restrict_preprocess(self.buffer)
        return out

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        value_key = prefix + 'value'
        retro = prefix + 'learned_value'
        if retro in state_dict:
            state_dict[value_key] = state_dict.pop(retro)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)
        if prefix + 'buffer' in missing_keys:
            missing_keys.remove(prefix + 'buffer')
        training_key = prefix + 'training'
        if training_key in missing_keys:
            missing_keys.remove(training_key)
        if value_key not in missing_keys:
            self.counter = self.collect_stats_steps + 1     # a loaded value ends the collection phase
        if config.IGNORE_MISSING_KEYS and value_key in missing_keys:
            missing_keys.remove(value_key)


class _AffineRescaling(nn.Module):
    """Learned affine transform of a statistic (runtime.py:105-134)."""

    def __init__(self, scaling_shape):
        super().__init__()
        self.affine_weight = Parameter(torch.ones(scaling_shape))
        self.affine_bias = Parameter(torch.zeros(scaling_shape))

    def forward(self, x):
        return abs_binary_sign_grad(x * self.affine_weight + self.affine_bias)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)
        for key in (prefix + 'affine_weight', prefix + 'affine_bias'):
            if config.IGNORE_MISSING_KEYS and key in missing_keys:
                missing_keys.remove(key)


class _StatsScaling(nn.Module):
    """restrict-pre -> optional affine -> restrict + clamp_min (runtime.py:50-72)."""

    def __init__(self, restrict_scaling_impl: nn.Module, scaling_shape: Tuple[int, ...],
                 scaling_min_val: Optional[float] = None, affine_rescaling: bool = False) -> None:
        super().__init__()
        self.affine_rescaling = _AffineRescaling(scaling_shape) if affine_rescaling else Identity()
        self.restrict_clamp_scaling = _RestrictClampValue(scaling_min_val, restrict_scaling_impl)
        self.restrict_scaling_pre = restrict_scaling_impl.restrict_init_module()

    def fused_min_val(self) -> Optional[float]:
        """scaling_min_val if this post-processing is exactly ``clamp_min_ste`` (or nothing), else None."""
        if type(self.affine_rescaling) is not Identity or type(self.restrict_scaling_pre) is not Identity:
            return None
        rc = self.restrict_clamp_scaling
        if type(rc.restrict_value_impl) not in (Identity, FloatRestrictValue):
            return None
        if type(rc.clamp_min_ste) is ScalarClampMinSte:
            return float(rc.clamp_min_ste.min_val)
        if type(rc.clamp_min_ste) is Identity:
            return 0.0
        return None

    def forward(self, stats: Tensor) -> Tensor:
        stats = self.restrict_scaling_pre(stats)
        stats = self.affine_rescaling(stats)
        return self.restrict_clamp_scaling(stats)


class StatsFromParameterScaling(nn.Module):
    """Threshold from a statistic of the tracked parameters, e.g. per-channel abs-max (runtime.py:19-47)."""

    def __init__(self, scaling_stats_impl: nn.Module, scaling_stats_input_view_shape_impl: nn.Module,
                 scaling_stats_input_concat_dim: int, tracked_parameter_list: List[Parameter],
                 restrict_scaling_impl: nn.Module, scaling_shape: Tuple[int, ...], affine_rescaling: bool = False,
                 scaling_min_val: Optional[float] = None) -> None:
        super().__init__()
        self.parameter_list_stats = _ParameterListStats(
            scaling_stats_impl, scaling_shape, scaling_stats_input_view_shape_impl, scaling_stats_input_concat_dim,
            tracked_parameter_list)
        self.stats_scaling_impl = _StatsScaling(restrict_scaling_impl, scaling_shape, scaling_min_val, affine_rescaling)

    def fused_stats_plan(self, x: Tensor) -> Optional[FusedStatsPlan]:
        mv = self.stats_scaling_impl.fused_min_val()
        if mv is None:
            return None
        geom = self.parameter_list_stats.fused_plan(x)
        if geom is None:
            return None
        return FusedStatsPlan(geom, mv, self.parameter_list_stats.stats.stats_output_shape)

    def forward(self, ignored: Tensor) -> Tensor:
        return self.stats_scaling_impl(self.parameter_list_stats())


class RuntimeStatsScaling(nn.Module):
    """Threshold from a statistic of the runtime input (batch stats in training, EMA in eval) (runtime.py:75-102)."""

    def __init__(self, scaling_stats_impl: nn.Module, scaling_stats_input_view_shape_impl: nn.Module,
                 restrict_scaling_impl: nn.Module, scaling_shape: Tuple[int, ...], affine_rescaling: bool,
                 scaling_stats_momentum: float = DEFAULT_MOMENTUM, scaling_min_val: Optional[float] = None) -> None:
        super().__init__()
        self.runtime_stats = _RuntimeStats(scaling_stats_impl, scaling_shape, scaling_stats_input_view_shape_impl,
                                           scaling_stats_momentum)
        self.stats_scaling_impl = _StatsScaling(restrict_scaling_impl, scaling_shape, scaling_min_val, affine_rescaling)

    def fused_stats_plan(self, x: Tensor) -> Optional[FusedStatsPlan]:
        mv = self.stats_scaling_impl.fused_min_val()
        if mv is None:
            return None
        geom = self.runtime_stats.fused_plan(x)
        if geom is None:
            return None
        rs = self.runtime_stats
        return FusedStatsPlan(geom, mv, rs.stats.stats_output_shape, on_absmax=rs.update_running)

    def forward(self, x: Tensor):
        return self.stats_scaling_impl(self.runtime_stats(x))
