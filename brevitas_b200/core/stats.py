"""Mirror of ``brevitas.core.stats`` for the hot path: AbsMax, AbsPercentile and the stats wrappers.

Reference: src/brevitas/core/stats/stats_op.py:41-66 (AbsPercentile), :129-141 (AbsMax);
stats_wrapper.py:19-114 (_Stats, _RuntimeStats, _ParameterListStats); view_wrapper.py:13-64.
The reductions run on the sm_100a kernels (``brevitas_b200::absmax_rows`` / ``absmax_tensor`` /
``abs_kth_value_rows``); when a quantizer can fuse the statistic with the quant-dequant pass it asks the
wrapper for an :class:`AbsMaxPlan` instead of calling it (see ``RescalingIntQuant``).
"""
import contextlib
import math
import threading
from typing import List, NamedTuple, Optional, Tuple

import torch

from .. import config
from torch import Tensor, nn

from .. import ops as _ops  # noqa: F401
from . import function_wrapper as fw

DEFAULT_MOMENTUM = 0.1
SCALAR_SHAPE = ()


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

    def relu_tensor_supported(self, x: Tensor) -> bool:
        dense = x.is_contiguous() or (x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last))
        return self.stats_reduce_dim is None and x.is_cuda and dense and x.numel() > 0

    def forward_relu_tensor(self, x: Tensor, holder: dict) -> Tensor:
        """``self(relu(x).reshape(-1))`` without materialising relu(x); ``holder`` pairs this call with the quantizer call
        of the same forward so that the statistic's one-element gradient is folded into the quantizer's (ops.py)"""
        from ..ops import CollectingStat
        k = int(math.floor(.01 * self.q * x.numel() + 0.5))
        return CollectingStat.apply(x, k, holder)


# ---- remaining statistics of stats_op.py.  Those built on the per-row abs-max reuse the
# sm_100a reduction, the signed percentiles the exact radix select on order-preserving keys (``kth_value_rows``: the
# reference's ``kthvalue`` is a sort-class ATen op, 1.7 s on a ResNet-18 activation); min / mean / variance are the
# ATen reductions the reference issues on statistics-sized outputs -----------------
def _kth_signed(x: Tensor, k: int, dim: Optional[int]) -> Tensor:
    """``x.view(-1).kthvalue(k).values`` (dim None) or ``x.kthvalue(k, dim).values`` of a 2-D x"""
    if x.dtype not in (torch.float32, torch.bfloat16, torch.float16):
        # integer / fp64 statistics inputs (the reference's own tests feed integer tensors, tests/brevitas/core/
        # test_stats.py:44-72): ATen's kthvalue, as the reference
        if dim is None:
            return x.view(-1).kthvalue(k).values
        return x.kthvalue(k, dim=dim).values
    if dim is None:
        val, _ = torch.ops.brevitas_b200.kth_value_rows(x.reshape(-1), 1, x.numel(), k)
        return val.view(())
    if dim % 2 == 0:
        x = x.t()
    rows, cols = x.shape
    val, _ = torch.ops.brevitas_b200.kth_value_rows(x.contiguous(), rows, cols, k)
    return val


DEFAULT_STD_DEV_EPSILON = 1e-8


# ---- minimum and maximum from ONE read, shared between the statistics of one quantizer call ------------------------------
# The asymmetric weight quantizers take torch.max + torch.min (AbsMinMax -> scale) and torch.min again (NegativeMinOrZero ->
# zero-point) over the same view of the same weight.  `minmax_rows` returns both extrema from one pass; inside
# `shared_minmax()` (entered by RescalingIntQuant around its scale and zero-point computation) the second statistic reuses
# the result of the first instead of reading the tensor again.
class _MinMaxMemo(threading.local):
    table = None


_MEMO = _MinMaxMemo()


@contextlib.contextmanager
def shared_minmax():
    prev, _MEMO.table = _MEMO.table, {}
    try:
        yield
    finally:
        _MEMO.table = prev


def _minmax(x: Tensor, dim: Optional[int]):
    """(min, max) of the whole tensor (dim None) or along one dim of a 2-D tensor; None if the kernel does not apply"""
    if not (x.is_cuda and x.dtype in (torch.float32, torch.bfloat16, torch.float16) and x.numel() > 0
            and x.numel() < 2 ** 31 and (dim is None or x.dim() == 2)):
        return None
    table = _MEMO.table
    key = (x.data_ptr(), tuple(x.shape), tuple(x.stride()), x.dtype, x._version, dim)
    if table is not None and key in table:
        return table[key]
    if dim is None:
        mn, mx, _, _ = torch.ops.brevitas_b200.minmax_rows(x.reshape(-1), 1, x.numel(), True)
        res = (mn.view(()), mx.view(()))
    else:
        xx = x.t() if dim % 2 == 0 else x
        rows, cols = xx.shape
        mn, mx, _, _ = torch.ops.brevitas_b200.minmax_rows(xx.contiguous(), rows, cols, False)
        res = (mn, mx)
    if table is not None:
        table[key] = res
    return res


class NegativeMinOrZero(nn.Module):
    """``min(x)`` (whole tensor or along a dim) if it is <= 0 else 0 (stats_op.py:22-39)."""

    def __init__(self, stats_reduce_dim: Optional[int] = None) -> None:
        super().__init__()
        from .utils import StatelessBuffer
        self.stats_reduce_dim = stats_reduce_dim
        self.zero = StatelessBuffer(torch.tensor(0.0))

    def forward(self, x: Tensor) -> Tensor:
        both = _minmax(x, self.stats_reduce_dim)
        if both is not None:
            min_val = both[0]
        elif self.stats_reduce_dim is None:
            min_val = torch.min(x)
        else:
            min_val = torch.min(x, dim=self.stats_reduce_dim)[0]
        zero = self.zero().to(min_val.dtype)
        return torch.where(min_val <= zero, min_val, zero)


class NegativePercentileOrZero(nn.Module):
    """k-th smallest (signed) value with ``k = ceil(.01 * q * n)``, or 0 if positive (stats_op.py:69-97)."""

    def __init__(self, low_percentile_q, stats_reduce_dim: Optional[int] = None) -> None:
        super().__init__()
        from .utils import StatelessBuffer
        self.stats_reduce_dim = stats_reduce_dim
        self.q = low_percentile_q
        self.zero = StatelessBuffer(torch.tensor(0.0))

    def forward(self, x: Tensor) -> Tensor:
        if self.stats_reduce_dim is None:
            k = int(math.ceil(.01 * self.q * x.numel()))
        else:
            assert len(x.size()) == 2, "Only 2-dim input is supported."
            k = int(math.ceil(.01 * self.q * x.shape[self.stats_reduce_dim]))
        result = _kth_signed(x, k, self.stats_reduce_dim)
        zero = self.zero().to(result.dtype)
        return torch.where(result <= zero, result, zero)


class PercentileInterval(nn.Module):
    """``|high percentile - low percentile|`` (stats_op.py:100-126)."""

    def __init__(self, low_percentile_q, high_percentile_q, stats_reduce_dim: Optional[int] = None) -> None:
        super().__init__()
        self.stats_reduce_dim = stats_reduce_dim
        self.low_q = low_percentile_q
        self.high_q = high_percentile_q

    def forward(self, x: Tensor) -> Tensor:
        if self.stats_reduce_dim is None:
            n = x.numel()
            low_k = int(math.ceil(.01 * self.low_q * n))
            high_k = int(math.floor(.01 * self.high_q * n + 0.5))
            low_result = _kth_signed(x, low_k, None)
            high_result = _kth_signed(x, high_k, None)
        else:
            assert len(x.size()) == 2, "Only 2-dim input is supported."
            n = x.shape[self.stats_reduce_dim]
            low_k = int(math.ceil(.01 * self.low_q * n))
            high_k = int(math.floor(.01 * self.high_q * n + 0.5))
            low_result = _kth_signed(x, low_k, self.stats_reduce_dim)
            high_result = _kth_signed(x, high_k, self.stats_reduce_dim)
        return torch.abs(high_result - low_result)


class AbsMinMax(nn.Module):
    """``|max(x) - min(x)|`` (stats_op.py:144-158)."""

    def __init__(self, stats_reduce_dim: Optional[int] = None) -> None:
        super().__init__()
        self.stats_reduce_dim = stats_reduce_dim

    def forward(self, x: Tensor):
        both = _minmax(x, self.stats_reduce_dim)
        if both is not None:
            return torch.abs(both[1] - both[0])
        if self.stats_reduce_dim is None:
            return torch.abs(torch.max(x) - torch.min(x))
        max_val = torch.max(x, dim=self.stats_reduce_dim)[0]
        min_val = torch.min(x, dim=self.stats_reduce_dim)[0]
        return torch.abs(max_val - min_val)


class AbsMaxAve(nn.Module):
    """Mean of the per-row abs-max (stats_op.py:161-170); the abs-max is the sm_100a row reduction."""

    def __init__(self, stats_reduce_dim: int) -> None:
        super().__init__()
        self.absmax = AbsMax(stats_reduce_dim)

    def forward(self, x: Tensor):
        return torch.mean(self.absmax(x))


class AbsMaxL2(nn.Module):
    """L2 norm of the per-row abs-max over sqrt(#rows) (stats_op.py:173-185)."""

    def __init__(self, stats_reduce_dim: int) -> None:
        super().__init__()
        self.absmax = AbsMax(stats_reduce_dim)

    def forward(self, x: Tensor):
        per_channel_max = self.absmax(x)
        out = torch.norm(per_channel_max, p=2)
        return out / math.sqrt(per_channel_max.view(-1).shape[0])


class AbsAve(nn.Module):
    """``mean(abs(x))`` (stats_op.py:188-200)."""

    def __init__(self, stats_reduce_dim: Optional[int] = None) -> None:
        super().__init__()
        self.stats_reduce_dim = stats_reduce_dim

    def forward(self, x: Tensor):
        if self.stats_reduce_dim is None:
            return torch.mean(torch.abs(x))
        return torch.mean(torch.abs(x), dim=self.stats_reduce_dim)


class _MeanSigmaStdImpl(nn.Module):
    """``mean(|x|) + sigma * sqrt(var(|x|) + eps)`` (stats_op.py:221-246)."""

    def __init__(self, stats_reduce_dim: Optional[int] = None, std_dev_epsilon: float = DEFAULT_STD_DEV_EPSILON) -> None:
        super().__init__()
        self.stats_reduce_dim = stats_reduce_dim
        self.epsilon = std_dev_epsilon

    def forward(self, x: Tensor, sigma: Tensor):
        abs_val = torch.abs(x)
        if self.stats_reduce_dim is None:
            mean_val = torch.mean(abs_val)
            std_val = torch.sqrt(torch.var(abs_val) + self.epsilon)
        else:
            mean_val = torch.mean(torch.abs(x), dim=self.stats_reduce_dim)
            std_val = torch.sqrt(torch.var(abs_val, dim=self.stats_reduce_dim) + self.epsilon)
            mean_val = mean_val.view(-1)
            std_val = std_val.view(-1)
        return mean_val + sigma * std_val


class MeanSigmaStd(nn.Module):
    """stats_op.py:203-218 (constant sigma)."""

    def __init__(self, sigma: float, stats_reduce_dim: Optional[int] = None,
                 std_dev_epsilon: float = DEFAULT_STD_DEV_EPSILON) -> None:
        super().__init__()
        from .utils import StatelessBuffer
        self.impl = _MeanSigmaStdImpl(stats_reduce_dim, std_dev_epsilon)
        self.sigma = StatelessBuffer(torch.tensor(sigma))

    def forward(self, x: Tensor):
        return self.impl(x, self.sigma())


class MeanLearnedSigmaStd(nn.Module):
    """stats_op.py:249-285 (learned sigma).  The parameter is registered as ``value`` like in the reference (whose
    forward reads a non-existent ``self.sigma``, stats_op.py:269); older key names are mapped on load."""

    def __init__(self, sigma: float, stats_output_shape: Tuple[int, ...], stats_reduce_dim: Optional[int] = None,
                 std_dev_epsilon: float = DEFAULT_STD_DEV_EPSILON) -> None:
        super().__init__()
        self.impl = _MeanSigmaStdImpl(stats_reduce_dim, std_dev_epsilon)
        if stats_output_shape == SCALAR_SHAPE:
            self.value = nn.Parameter(torch.tensor(sigma))
        else:
            self.value = nn.Parameter(torch.full(stats_output_shape, sigma))

    def forward(self, x: Tensor):
        return self.impl(x, self.value.view(self.value.shape))

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        value_key = prefix + 'value'
        for retro_key in (prefix + 'sigma', prefix + 'learned_sigma'):
            if retro_key in state_dict:
                state_dict[value_key] = state_dict.pop(retro_key)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)
        if config.IGNORE_MISSING_KEYS and value_key in missing_keys:
            missing_keys.remove(value_key)


def absmax_plan(stats_impl: nn.Module, view_impl: nn.Module, x: Tensor) -> Optional[AbsMaxPlan]:
    """Return the fused geometry when (view, AbsMax) reduces contiguous trailing elements of ``x``."""
    if type(stats_impl) is not AbsMax or x.numel() == 0:
        return None
    # channels-last tensors keep dim 0 as the slowest dimension in memory: whole-tensor and per-dim-0-slice statistics
    # (and the element-wise quantization) do not care about the order inside a slice
    row_major = x.is_contiguous()
    dense = row_major or (x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)) \
        or (x.dim() == 5 and x.is_contiguous(memory_format=torch.channels_last_3d))
    if not dense:
        return None
    dim = stats_impl.stats_reduce_dim
    if type(view_impl) is fw.OverTensorView and dim is None:
        return AbsMaxPlan('tensor', 1, x.numel())
    if type(view_impl) is fw.OverOutputChannelView and type(view_impl.permute_impl) is fw.Identity and dim in (1, -1):
        if x.dim() >= 1:
            return AbsMaxPlan('rows', x.shape[0], x.numel() // x.shape[0])
    if type(view_impl) is fw.OverBatchOverTensorView and dim in (1, -1) and x.dim() >= 1:
        return AbsMaxPlan('rows', x.shape[0], x.numel() // x.shape[0])
    if not row_major:
        return None
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
        if config.IGNORE_MISSING_KEYS and key in missing_keys:
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
