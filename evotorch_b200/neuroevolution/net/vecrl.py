"""`evotorch.neuroevolution.net.vecrl` import path for `Policy` (the implementation lives in neuroevolution/policy.py) and
`reset_tensors` (vecrl.py:912-940)."""

from collections.abc import Iterable, Mapping
from typing import Any

import torch

from ..policy import Policy


def reset_tensors(x: Any, indices):
    """Zero the rows selected by `indices` (an index list or a boolean mask) in a tensor, or in every tensor of a nested
    container -- how the recurrent state of finished sub-environments is cleared."""
    if isinstance(x, torch.Tensor):
        x[indices] = 0
    elif isinstance(x, (str, bytes, bytearray)):
        return
    elif isinstance(x, Mapping):
        for value in x.values():
            reset_tensors(value, indices)
    elif isinstance(x, Iterable):
        for value in x:
            reset_tensors(value, indices)


__all__ = ["Policy", "reset_tensors"]
