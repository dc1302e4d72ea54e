"""Mirror of ``brevitas.core.stats`` for the hot path: AbsMax, AbsPercentile and the stats wrappers.

Reference: src/brevitas/core/stats/stats_op.py:41-66 (AbsPercentile), :129-141 (AbsMax);
stats_wrapper.py:19-114 (_Stats, _RuntimeStats, _ParameterListStats); view_wrapper.py:13-64.
The reductions run on the sm_100a kernels (``brevitas_b200::absmax_rows`` / ``absmax_tensor`` /
``abs_kth_value_rows``); when a quantizer can fuse the statistic with the quant-dequant pass it asks the
wrapper for an :class:`AbsMaxPlan` instead of calling it (see ``RescalingIntQuant``).
"""
import math
from typing import List, NamedTuple, Optional, Tuple

import torch
from torch import Tensor, nn

from .. import ops as _ops  # noqa: F401
from . import function_wrapper as fw

DEFAULT_MOMENTUM = 0.1
SCALAR_SHAPE = ()
IGNORE_MISSING_KEYS = False   # mirrors brevitas.config.IGNORE_MISSING_KEYS


class AbsMaxPlan(NamedTuple):
    """Geometry of an abs-max statistic that can be fused into the quant kernel."""
    kind: str            # 'rows' or 'tensor'
    rows: int
    cols: int


def _rows_cols_last(x: Tensor) -> Tuple[int, int]:
    cols = x.shape[-1]
    rows = x.numel() // cols if cols else 0
    return rows, cols


class AbsMax(nn.Module):
    """``max(abs(x))`` over the whole tensor or along one dim (stats_op.py:129-141); NaN-propagating."""

    def __init__(self, stats_reduce_dim: Optional[int] = None) -> None:
        super().__init__()
        self.stats_reduce_dim = stats_reduce_dim

    def forward(self, x: Tensor):
        if self.stats_reduce_dim is None:
            return torch.ops.brevitas_b200.absmax_tensor(x)
        dim = self.stats_reduce_dim % x.dim()
        if dim != x.dim() - 1:
            x = x.movedim(dim, -1)
        out_shape = x.shape[:-1]
        rows, cols = _rows_cols_last(x)
        return torch.ops.brevitas_b200.absmax_rows(x.contiguous(), rows, cols).view(out_shape)


class AbsPercentile(nn.Module):
    """k-th smallest of ``abs(x)`` with ``k = floor(.01 * q * n + .5)`` (stats_op.py:41-66); exact radix select."""

    def __init__(self, high_percentile_q: float, stats_reduce_dim: Optional[int], percentile_q=None):
        super().__init__()
        if percentile_q is not None:
            raise RuntimeError("percentile_q is deprecated, please pass high_percentile_q.")
        assert high_percentile_q <= 100, "q has to be a percentage"
        self.q = high_percentile_q
        self.stats_reduce_dim = stats_reduce_dim

    def forward(self, x: Tensor):
        if self.stats_reduce_dim is None:
            n = x.numel()
            k = int(math.floor(.01 * self.q * n + 0.5))
            val, _ = torch.ops.brevitas_b200.abs_kth_value_rows(x.reshape(-1), 1, n, k)
            return val.view(())
        assert len(x.size()) == 2, "Only 2-dim input is supported."
        if self.stats_reduce_dim % 2 == 0:
            x = x.t()
        rows, cols = x.shape
        k = int(math.floor(.01 * self.q * cols + 0.5))
        val, _ = torch.ops.brevitas_b200.abs_kth_value_rows(x.contiguous(), rows, cols, k)
        return val


def absmax_plan(stats_impl: nn.Module, view_impl: nn.Module, x: Tensor) -> Optional[AbsMaxPlan]:
    """Return the fused geometry when (view, AbsMax) reduces contiguous trailing elements of ``x``."""
    if type(stats_impl) is not AbsMax or not x.is_contiguous() or x.numel() == 0:
        return None
    dim = stats_impl.stats_reduce_dim
    if type(view_impl) is fw.OverTensorView and dim is None:
        return AbsMaxPlan('tensor', 1, x.numel())
    if type(view_impl) is fw.OverOutputChannelView and type(view_impl.permute_impl) is fw.Identity and dim in (1, -1):
        if x.dim() >= 1:
            return AbsMaxPlan('rows', x.shape[0], x.numel() // x.shape[0])
    if type(view_impl) is fw.OverBatchOverTensorView and dim in (1, -1) and x.dim() >= 1:
        return AbsMaxPlan('rows', x.shape[0], x.numel() // x.shape[0])
    if type(view_impl) is fw.OverBatchOverOutputChannelView and dim in (2, -1) and x.dim() >= 2:
        rows = x.shape[0] * x.shape[1]
        return AbsMaxPlan('rows', rows, x.numel() // rows)
    return None


class _Stats(nn.Module):
    """stats op followed by a reshape to the scaling shape (stats_wrapper.py:19-34)."""

    def __init__(self, stats_impl: nn.Module, stats_output_shape: Tuple[int, ...]) -> None:
        super().__init__()
        self.stats_output_shape = stats_output_shape
        self.stats_impl = stats_impl

    def forward(self, input: Tensor) -> Tensor:
        return self.stats_impl(input).view(self.stats_output_shape)


class _RuntimeStats(nn.Module):
    """Batch statistic in training with an EMA buffer used in eval (stats_wrapper.py:37-81)."""

    def __init__(self, stats_impl: nn.Module, stats_output_shape: Tuple[int, ...],
                 stats_input_view_shape_impl: nn.Module, stats_buffer_momentum: float = DEFAULT_MOMENTUM) -> None:
        super().__init__()
        self.first_batch = True
        self.stats_input_view_shape_impl = stats_input_view_shape_impl
        self.stats = _Stats(stats_impl, stats_output_shape)
        self.momentum = stats_buffer_momentum
        self.register_buffer('running_stats', torch.full(stats_output_shape, 1.0))

    def update_running(self, out: Tensor) -> None:
        """``running *= stat`` on the first batch, EMA afterwards (stats_wrapper.py:60-65)."""
        out = out.detach()
        rs = self.running_stats
        if rs.dtype == torch.float32 and rs.is_cuda and rs.is_contiguous() and out.numel() == rs.numel():
            torch.ops.brevitas_b200.running_stats_update_(rs, out.reshape(rs.shape), self.momentum, self.first_batch)
        elif self.first_batch:
            rs *= out
        else:
            rs *= (1 - self.momentum)
            rs += self.momentum * out
        self.first_batch = False

    def fused_plan(self, x: Tensor) -> Optional[AbsMaxPlan]:
        if not self.training:
            return None
        return absmax_plan(self.stats.stats_impl, self.stats_input_view_shape_impl, x)

    def forward(self, stats_input) -> Tensor:
        if self.training:
            stats_input = self.stats_input_view_shape_impl(stats_input)
            out = self.stats(stats_input)
            self.update_running(out)
        else:
            out = self.running_stats
        return out

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)
        key = prefix + 'running_stats'
        if IGNORE_MISSING_KEYS and key in missing_keys:
            missing_keys.remove(key)
        training_key = prefix + 'training'
        if training_key in missing_keys:
            missing_keys.remove(training_key)


class _ViewParameterWrapper(nn.Module):
    """Holds a tracked parameter without owning it in state dicts (view_wrapper.py:13-37)."""

    def __init__(self, parameter: nn.Parameter, view_shape_impl: nn.Module) -> None:
        super().__init__()
        self.parameter = parameter
        self.view_shape_impl = view_shape_impl

    def forward(self) -> Tensor:
        return self.view_shape_impl(self.parameter)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)
        key = prefix + 'parameter'
        if key in missing_keys:
            missing_keys.remove(key)

    def state_dict(self, *args, destination=None, prefix='', keep_vars=False):
        out = super().state_dict(*args, destination=destination, prefix=prefix, keep_vars=keep_vars)
        out.pop(prefix + 'parameter', None)
        return out


class _ViewCatParameterWrapper(_ViewParameterWrapper):
    """Concatenates the view of one more tracked parameter (view_wrapper.py:39-64)."""

    def __init__(self, parameter: nn.Parameter, view_shape_impl: nn.Module, cat_dim: int) -> None:
        super().__init__(parameter, view_shape_impl)
        self.cat_dim = cat_dim

    def forward(self, x: Tensor) -> Tensor:
        return torch.cat([self.view_shape_impl(self.parameter), x], dim=self.cat_dim)


class _ParameterListStats(nn.Module):
    """Statistic over the (concatenated) views of a list of parameters (stats_wrapper.py:83-114)."""

    def __init__(self, stats_impl: nn.Module, stats_output_shape: Tuple[int, ...],
                 stats_input_view_shape_impl: nn.Module, stats_input_concat_dim: int,
                 tracked_parameter_list: List[nn.Parameter]) -> None:
        super().__init__()
        self.stats_input_concat_dim = stats_input_concat_dim
        self.first_tracked_param = _ViewParameterWrapper(tracked_parameter_list[0], stats_input_view_shape_impl)
        if len(tracked_parameter_list) > 1:
            self.extra_tracked_params_list = nn.ModuleList([
                _ViewCatParameterWrapper(p, stats_input_view_shape_impl, stats_input_concat_dim)
                for p in tracked_parameter_list[1:]])
        else:
            self.extra_tracked_params_list = None
        self.stats = _Stats(stats_impl, stats_output_shape)

    def fused_plan(self, x: Tensor) -> Optional[AbsMaxPlan]:
        """Fusable only when the statistic is taken over exactly the tensor being quantized."""
        if self.extra_tracked_params_list is not None:
            return None
        p = self.first_tracked_param.parameter
        if x is not p:      # otherwise the two gradient paths (quant, statistic) end on different tensors
            return None
        return absmax_plan(self.stats.stats_impl, self.first_tracked_param.view_shape_impl, x)

    def forward(self) -> Tensor:
        stats_input = self.first_tracked_param()
        if self.extra_tracked_params_list is not None:
            for extra in self.extra_tracked_params_list:
                stats_input = extra(stats_input)
        return self.stats(stats_input)
