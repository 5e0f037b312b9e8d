"""`RunningStat`: the numpy-facing wrapper around `RunningNorm` that the reference's non-vectorised rollouts use
(net/runningstat.py:25-152): observations are single numpy vectors, everything lives on the cpu."""

from __future__ import annotations

from copy import deepcopy
from typing import Optional, Union

import numpy as np
import torch
from torch import nn

from ..runningnorm import RunningNorm


class RunningStat:
    def __init__(self):
        self._rn: Optional[RunningNorm] = None

    def reset(self):
        self._rn = None

    @property
    def count(self) -> int:
        return 0 if self._rn is None else self._rn.count

    @property
    def sum(self) -> np.ndarray:
        return self._rn.sum.numpy()

    @property
    def sum_of_squares(self) -> np.ndarray:
        return self._rn.sum_of_squares.numpy()

    @property
    def mean(self) -> np.ndarray:
        return self._rn.mean.numpy()

    @property
    def stdev(self) -> np.ndarray:
        return self._rn.stdev.numpy()

    def update(self, x: Union[np.ndarray, "RunningStat"]):
        """Add one observation vector, or merge the contents of another RunningStat."""
        if isinstance(x, RunningStat):
            if x.count > 0:
                if self._rn is None:
                    self._rn = deepcopy(x._rn)
                else:
                    self._rn.update(x._rn)
            return
        if self._rn is None:
            x = np.array(x, dtype="float32")
            self._rn = RunningNorm(shape=x.shape, dtype="float32", device="cpu")
        self._rn.update(x)

    def normalize(self, x: Union[np.ndarray, list]) -> np.ndarray:
        return x if self._rn is None else self._rn.normalize(np.array(x, dtype="float32"))

    def to(self, device) -> "RunningStat":
        if torch.device(device) != torch.device("cpu"):
            raise ValueError(f"The received target device is {device!r}. However, RunningStat can only work on a cpu.")
        return self

    def to_layer(self) -> nn.Module:
        return self._rn.to_layer()

    def __copy__(self) -> "RunningStat":
        return deepcopy(self)

    def __repr__(self) -> str:
        return f"<{type(self).__name__}, count: {self.count}>"
