"""Mirror of ``brevitas.core.function_wrapper`` (ops_ste.py, clamp.py, shape.py, misc.py).

The class identities matter: the fused quantizers read ``type(float_to_int_impl)`` / ``type(tensor_clamp_impl)``
to select the rounding and the clamp-gradient mode of the kernel.
"""
from typing import Optional, Tuple

import torch
from torch import Tensor, nn

from ..function import ops as F_ops
from ..function import ops_ste as F_ste
from ..function import shape as F_shape


class Identity(nn.Module):
    def forward(self, x: Tensor) -> Tensor:
        return x


# ---- float -> int with straight-through gradient (core/function_wrapper/ops_ste.py:14-92) -------------------
class RoundSte(nn.Module):
    def forward(self, x: Tensor) -> Tensor:
        return F_ste.round_ste(x)


class FloorSte(nn.Module):
    def forward(self, x: Tensor) -> Tensor:
        return F_ste.floor_ste(x)


class CeilSte(nn.Module):
    def forward(self, x: Tensor) -> Tensor:
        return F_ste.ceil_ste(x)


class RoundToZeroSte(nn.Module):
    def forward(self, x: Tensor) -> Tensor:
        return F_ste.round_to_zero_ste(x)


class DPURoundSte(nn.Module):
    def forward(self, x: Tensor) -> Tensor:
        return F_ste.dpu_round_ste(x)


# ---- clamps (core/function_wrapper/ops_ste.py:95-118, clamp.py:16-80) ---------------------------------------
class TensorClampSte(nn.Module):
    """Pass-through gradient (ops/autograd_ste_ops.py:118-120)."""

    def forward(self, x: Tensor, min_val: Tensor, max_val: Tensor) -> Tensor:
        return F_ste.tensor_clamp_ste(x, min_val, max_val)


class InplaceTensorClampSte(nn.Module):
    def forward(self, x: Tensor, min_val: Tensor, max_val: Tensor) -> Tensor:
        return F_ste.tensor_clamp_ste_(x, min_val, max_val)


class TensorClamp(nn.Module):
    """Differentiable clamp: masked gradient and gradients to the bounds (function/ops.py:98-99)."""

    def forward(self, x: Tensor, min_val: Tensor, max_val: Tensor) -> Tensor:
        return F_ops.tensor_clamp(x, min_val=min_val, max_val=max_val)


class ScalarClampSte(nn.Module):
    def __init__(self, min_val: float, max_val: float) -> None:
        super().__init__()
        self.min_val = min_val
        self.max_val = max_val

    def forward(self, x: Tensor) -> Tensor:
        return F_ste.scalar_clamp_ste(x, self.min_val, self.max_val)


class ScalarClampMinSte(nn.Module):
    def __init__(self, min_val: float) -> None:
        super().__init__()
        self.min_val = min_val

    def forward(self, x: Tensor) -> Tensor:
        return F_ste.scalar_clamp_min_ste(x, self.min_val)


class ClampMin(nn.Module):
    def __init__(self, min_val: float) -> None:
        super().__init__()
        self.min_val = min_val

    def forward(self, x: Tensor) -> Tensor:
        return x.clamp_min(self.min_val)


# ---- misc (core/function_wrapper/misc.py) -------------------------------------------------------------------
class PowerOfTwo(nn.Module):
    def forward(self, x: Tensor) -> Tensor:
        return 2.0 ** x


class LogTwo(nn.Module):
    def forward(self, x: Tensor) -> Tensor:
        return torch.log2(x)


class InplaceLogTwo(nn.Module):
    def forward(self, x: Tensor) -> Tensor:
        x.log2_()
        return x


# ---- views (core/function_wrapper/shape.py:30-115) ----------------------------------------------------------
class PermuteDims(nn.Module):
    def __init__(self, permute_dims: Tuple[int, ...]) -> None:
        super().__init__()
        self.permute_dims = permute_dims

    def forward(self, x: Tensor) -> Tensor:
        return x.permute(*self.permute_dims).contiguous()


class OverTensorView(nn.Module):
    """Flat view for whole-tensor statistics (function_wrapper/shape.py:30-45).  Every whole-tensor statistic (max, min,
    mean, k-th value) is invariant under a permutation of the elements, so a channels-last tensor is flattened in its
    MEMORY order -- a view -- instead of being copied into logical NCHW order first."""

    def forward(self, x: Tensor) -> Tensor:
        if x.dim() == 4 and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last):
            return x.permute(0, 2, 3, 1).reshape(-1)
        return x.reshape(F_shape.over_tensor(x))


class OverOutputChannelView(nn.Module):
    def __init__(self, permute_dims: Optional[Tuple[int, ...]] = None) -> None:
        super().__init__()
        self.permute_impl = PermuteDims(permute_dims) if permute_dims is not None else Identity()

    def forward(self, x: Tensor) -> Tensor:
        y = self.permute_impl(x)
        return y.reshape(F_shape.over_output_channels(y))


class OverBatchOverTensorView(nn.Module):
    def forward(self, x: Tensor) -> Tensor:
        return x.reshape(F_shape.over_batch_over_tensor(x))


class OverBatchOverOutputChannelView(nn.Module):
    def forward(self, x: Tensor) -> Tensor:
        return x.reshape(F_shape.over_batch_over_output_channels(x))


class StatsInputViewShapeImpl(object):
    OVER_TENSOR = OverTensorView
    OVER_OUTPUT_CHANNELS = OverOutputChannelView
    OVER_BATCH_OVER_TENSOR = OverBatchOverTensorView
    OVER_BATCH_OVER_OUTPUT_CHANNELS = OverBatchOverOutputChannelView


ROUND_MODE_OF = {}   # filled below: float_to_int_impl class -> round-mode tag of the C-ABI
CLAMP_MODE_OF = {}   # tensor_clamp_impl class -> clamp-gradient tag


def _fill_tables():
    from .. import _lib
    ROUND_MODE_OF.update({RoundSte: _lib.ROUND, FloorSte: _lib.FLOOR, CeilSte: _lib.CEIL,
                          RoundToZeroSte: _lib.ROUND_TO_ZERO, DPURoundSte: _lib.DPU_ROUND})
    CLAMP_MODE_OF.update({TensorClampSte: _lib.CLAMP_STE, TensorClamp: _lib.CLAMP_MASKED})


_fill_tables()
