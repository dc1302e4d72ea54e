"""Mirror of src/brevitas/core/bit_width/const.py:14-40 (constant bit-width)."""
import torch
from torch import Tensor, nn

from .utils import StatelessBuffer


class BitWidthConst(nn.Module):
    """Constant bit-width in a 0-dim float tensor that is not part of the checkpoint."""

    def __init__(self, bit_width: int) -> None:
        super().__init__()
        assert isinstance(bit_width, int)
        self.bit_width = StatelessBuffer(torch.tensor(float(bit_width)))
        self.bit_width_value = int(bit_width)      # host copy: lets the fused kernels avoid a device sync

    def forward(self) -> Tensor:
        return self.bit_width()
