"""Mirror of src/brevitas/core/zero_point.py:27-35 (the symmetric, zero zero-point)."""
import torch
from torch import Tensor, nn

from .utils import StatelessBuffer


class ZeroZeroPoint(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.zero_point = StatelessBuffer(torch.tensor(0.0))

    def forward(self, x: Tensor, scale: Tensor, bit_width: Tensor) -> Tensor:
        return self.zero_point()
