"""View-shape helpers deciding the reduction geometry of the scale statistics.

Mirror of src/brevitas/function/shape.py:26-96 (same names, same return values).
"""
from typing import Tuple

from torch import Tensor

__all__ = ['over_tensor', 'over_output_channels', 'over_batch_over_tensor', 'over_batch_over_output_channels']


def over_tensor(x: Tensor) -> int:
    """Shape that flattens ``x`` completely (reference: function/shape.py:26-41)."""
    return -1


def over_output_channels(x: Tensor) -> Tuple[int, int]:
    """2-D shape with output channels (dim 0) as rows (reference: function/shape.py:44-60)."""
    return x.shape[0], -1


def over_batch_over_tensor(x: Tensor) -> Tuple[int, int]:
    """2-D shape with the batch (dim 0) as rows (reference: function/shape.py:63-79)."""
    return x.shape[0], -1


def over_batch_over_output_channels(x: Tensor) -> Tuple[int, int, int]:
    """3-D shape (batch, channel, rest) (reference: function/shape.py:82-96)."""
    return x.shape[0], x.shape[1], -1
