"""Mirror of src/brevitas/core/utils.py: ``StatelessBuffer`` and the in-place buffer helpers."""
from typing import Optional

import torch

VALUE_ATTR_NAME = 'value'


def inplace_tensor_add(tensor: torch.Tensor, value: torch.Tensor) -> torch.Tensor:
    tensor.add_(value)
    return tensor


def inplace_tensor_mul(tensor: torch.Tensor, value: torch.Tensor) -> torch.Tensor:
    tensor.mul_(value)
    return tensor


def inplace_momentum_update(tensor, update, momentum: Optional[float], counter: int, new_counter: int):
    """EMA (or cumulative average when momentum is None) of tiny scale buffers (core/utils.py:26-38)."""
    if momentum is None:
        tensor.mul_(counter / new_counter)
        tensor.add_(update / new_counter)
    else:
        tensor.mul_(1 - momentum)
        tensor.add_(momentum * update)
    return tensor


class StatelessBuffer(torch.nn.Module):
    """A buffer that follows ``.to()`` but never appears in state dicts (core/utils.py:41-64)."""

    def __init__(self, value: torch.Tensor):
        super().__init__()
        self.register_buffer(VALUE_ATTR_NAME, value)

    def forward(self):
        return self.value.detach()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)
        key = prefix + VALUE_ATTR_NAME
        if key in missing_keys:
            missing_keys.remove(key)

    def state_dict(self, *args, destination=None, prefix='', keep_vars=False):
        out = super().state_dict(*args, destination=destination, prefix=prefix, keep_vars=keep_vars)
        out.pop(prefix + VALUE_ATTR_NAME, None)
        return out
