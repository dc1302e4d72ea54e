"""Mirror of src/brevitas/core/bit_width: constant (const.py:14-40), learned (parameter.py:23-98), the "bits to remove"
parameter (:101-141) and the clamp on an incoming accumulator bit-width (const.py:43-78 ``MsbClampBitWidth``).
Bit-widths are 0-dim tensors; everything here is a handful of scalar ops on the STE kernels."""
import torch

from .. import config
from torch import Tensor, nn
from torch.nn import Parameter

from ..function.ops_ste import abs_binary_sign_grad, tensor_clamp_ste
from .utils import StatelessBuffer

MIN_INT_BIT_WIDTH = 2
NON_ZERO_EPSILON = 1e-6
REMOVE_ZERO_BIT_WIDTH = 0.1


class BitWidthConst(nn.Module):
    """Constant bit-width in a 0-dim float tensor that is not part of the checkpoint."""

    def __init__(self, bit_width: int) -> None:
        super().__init__()
        assert isinstance(bit_width, int)
        self.bit_width = StatelessBuffer(torch.tensor(float(bit_width)))
        self.bit_width_value = int(bit_width)      # host copy: lets the fused kernels avoid a device sync

    def forward(self) -> Tensor:
        return self.bit_width()


class BitWidthParameter(nn.Module):
    """Learned bit-width ``restrict(|offset| + base)`` (parameter.py:23-98)."""

    def __init__(self, bit_width: int, min_bit_width: int = MIN_INT_BIT_WIDTH, restrict_bit_width_impl: nn.Module = None,
                 override_pretrained_bit_width: bool = False) -> None:
        super().__init__()
        if restrict_bit_width_impl is None:
            from .function_wrapper import RoundSte
            from .restrict_val import IntRestrictValue
            restrict_bit_width_impl = IntRestrictValue(RoundSte())
        if bit_width < MIN_INT_BIT_WIDTH:
            raise RuntimeError("Int bit width has to be at least {}, instead is {}.".format(MIN_INT_BIT_WIDTH, bit_width))
        if min_bit_width < MIN_INT_BIT_WIDTH:
            raise RuntimeError("Min int bit width has to be at least {}, instead is {}.".format(MIN_INT_BIT_WIDTH,
                                                                                                 min_bit_width))
        if bit_width < min_bit_width:
            raise RuntimeError("Int bit width has to be at least {}, instead is {}.".format(min_bit_width, bit_width))
        bit_width = float(int(bit_width))
        min_bit_width = float(int(min_bit_width))
        bit_width_base = restrict_bit_width_impl.restrict_init_float(min_bit_width)
        bit_width = restrict_bit_width_impl.restrict_init_float(bit_width)
        self.bit_width_offset = Parameter(torch.tensor(bit_width - bit_width_base))
        self.bit_width_base = bit_width_base
        self.restrict_bit_width_impl = restrict_bit_width_impl
        self.override_pretrained = override_pretrained_bit_width

    def forward(self) -> Tensor:
        bit_width = abs_binary_sign_grad(self.bit_width_offset) + self.bit_width_base
        return self.restrict_bit_width_impl(bit_width)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        key = prefix + 'bit_width_offset'
        if self.override_pretrained and key in state_dict:
            del state_dict[key]
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)
        if config.IGNORE_MISSING_KEYS and key in missing_keys:
            missing_keys.remove(key)


class RemoveBitwidthParameter(nn.Module):
    """Learned number of bits to remove, ``1 / (eps + |coeff|)`` (parameter.py:101-141)."""

    def __init__(self, bit_width_to_remove: int, override_pretrained_bit_width: bool = False,
                 non_zero_epsilon: float = NON_ZERO_EPSILON, remove_zero_bit_width=REMOVE_ZERO_BIT_WIDTH):
        super().__init__()
        if bit_width_to_remove < 0:
            raise RuntimeError("Bit width to clamp has to be >= 0.")
        elif bit_width_to_remove == 0:
            bit_width_coeff_init = 1 / remove_zero_bit_width
        else:
            bit_width_coeff_init = 1 / bit_width_to_remove
        self.bit_width_coeff = Parameter(torch.tensor(bit_width_coeff_init))
        self.non_zero_epsilon = non_zero_epsilon
        self.override_pretrained = override_pretrained_bit_width

    def forward(self) -> Tensor:
        return 1.0 / (self.non_zero_epsilon + torch.abs(self.bit_width_coeff))

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        key = prefix + 'bit_width_coeff'
        if self.override_pretrained and key in state_dict:
            del state_dict[key]
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)
        if config.IGNORE_MISSING_KEYS and key in missing_keys:
            missing_keys.remove(key)


class MsbClampBitWidth(nn.Module):
    """``clamp_ste(|input_bit_width - bits_to_remove|, min, max)`` (const.py:43-78)."""

    def __init__(self, bit_width_to_remove_impl: nn.Module, min_overall_bit_width: int, max_overall_bit_width: int) -> None:
        super().__init__()
        self.min_overall_bit_width = BitWidthConst(min_overall_bit_width)
        self.max_overall_bit_width = BitWidthConst(max_overall_bit_width)
        self.bit_width_to_remove_impl = bit_width_to_remove_impl

    def forward(self, input_bit_width: Tensor) -> Tensor:
        bit_width_to_remove = self.bit_width_to_remove_impl()
        output_bit_width = torch.abs(input_bit_width - bit_width_to_remove)
        return tensor_clamp_ste(output_bit_width, self.min_overall_bit_width(), self.max_overall_bit_width())
