"""Mirror of src/brevitas/core/restrict_val.py: value restrictions applied to scale factors (tiny tensors)."""
import math
from typing import Optional

import torch
from torch import Tensor, nn

from .function_wrapper import Identity, InplaceLogTwo, LogTwo, PowerOfTwo, RoundSte, ScalarClampMinSte


class _RestrictClampValue(nn.Module):
    """restrict, then clamp_min with STE (restrict_val.py:22-42)."""

    def __init__(self, scaling_min_val: Optional[float], restrict_value_impl: Optional[nn.Module]):
        super().__init__()
        if scaling_min_val is not None and scaling_min_val != 0:
            self.clamp_min_ste = ScalarClampMinSte(scaling_min_val)
        else:
            self.clamp_min_ste = Identity()
        self.restrict_value_impl = restrict_value_impl if restrict_value_impl is not None else Identity()

    def forward(self, x: Tensor) -> Tensor:
        return self.clamp_min_ste(self.restrict_value_impl(x))


class _RestrictValue(nn.Module):
    def __init__(self, restrict_value_impl: Optional[nn.Module]):
        super().__init__()
        self.restrict_value_impl = restrict_value_impl if restrict_value_impl is not None else Identity()

    def forward(self, x: Tensor) -> Tensor:
        return self.restrict_value_impl(x)


class _ClampValue(nn.Module):
    def __init__(self, scaling_min_val: Optional[float]):
        super().__init__()
        if scaling_min_val is not None and scaling_min_val != 0:
            self.clamp_min_ste = ScalarClampMinSte(scaling_min_val)
        else:
            self.clamp_min_ste = Identity()
        self.min_val = scaling_min_val

    def forward(self, x: Tensor) -> Tensor:
        return self.clamp_min_ste(x)


class FloatRestrictValue(nn.Module):
    """No restriction (restrict_val.py:80-99)."""

    def restrict_init_float(self, x: float) -> float:
        return x

    def restrict_init_tensor(self, x: Tensor) -> Tensor:
        return x

    def restrict_init_module(self):
        return Identity()

    def restrict_init_inplace_module(self):
        return Identity()

    def forward(self, x: Tensor) -> Tensor:
        return x


class LogFloatRestrictValue(nn.Module):
    """Value kept in the log2 domain, ``2 ** v`` on use (restrict_val.py:102-124)."""

    def __init__(self):
        super().__init__()
        self.power_of_two = PowerOfTwo()

    def restrict_init_float(self, x: float):
        return math.log2(x)

    def restrict_init_tensor(self, x: Tensor):
        return torch.log2(x)

    def restrict_init_module(self):
        return LogTwo()

    def restrict_init_inplace_module(self):
        return InplaceLogTwo()

    def forward(self, x: Tensor):
        return self.power_of_two(x)


class IntRestrictValue(nn.Module):
    def __init__(self, restrict_value_float_to_int_impl: Optional[nn.Module] = None):
        super().__init__()
        self.float_to_int_impl = restrict_value_float_to_int_impl or RoundSte()

    def restrict_init_float(self, x: float):
        return x

    def restrict_init_tensor(self, x: Tensor):
        return x

    def restrict_init_module(self):
        return Identity()

    def restrict_init_inplace_module(self):
        return Identity()

    def forward(self, x: Tensor):
        return self.float_to_int_impl(x)


class PowerOfTwoRestrictValue(nn.Module):
    """``2 ** round(v)`` with the rounding under STE (restrict_val.py:150-173)."""

    def __init__(self, restrict_value_float_to_int_impl: Optional[nn.Module] = None):
        super().__init__()
        self.float_to_int_impl = restrict_value_float_to_int_impl or RoundSte()
        self.power_of_two = PowerOfTwo()

    def restrict_init_float(self, x: float):
        return math.log2(x)

    def restrict_init_tensor(self, x: Tensor):
        return torch.log2(x)

    def restrict_init_module(self):
        return LogTwo()

    def restrict_init_inplace_module(self):
        return InplaceLogTwo()

    def forward(self, x: Tensor):
        return self.power_of_two(self.float_to_int_impl(x))
