"""Running observation statistics and the normalisation layer made from them (mirrors net/runningnorm.py:47-613).

`RunningNorm` accumulates sum / sum of squares / count of (optionally masked) observation batches; `normalize(x)` returns
`clamp((x - mean) / stdev, low, high)` with `stdev = sqrt(max(E[x^2] - E[x]^2, min_variance))`.

On CUDA float32 the update of a batch is the K4 kernel in its raw-moments form (weights = the mask: sum_i m_i x_i and
sum_i m_i x_i^2 in one pass over the batch) and the count stays ON THE DEVICE: the reference's `int(torch.sum(mask))`
(runningnorm.py:318) is a host synchronisation per time step of every rollout; here nothing synchronises until somebody
asks for `.count`.  The K8 policy kernel reads the sums directly (`ops.mlp_forward(..., obs_sum=, obs_sumsq=, obs_count=)`),
so inside a rollout no normalised copy of the observations is ever materialised.
"""

from __future__ import annotations

from copy import deepcopy
from typing import Iterable, NamedTuple, Optional, Union

import torch
from torch import nn

from .. import ops


class CollectedStats(NamedTuple):
    mean: torch.Tensor
    stdev: torch.Tensor


def _clamp(x: torch.Tensor, lo: Optional[float], hi: Optional[float]) -> torch.Tensor:
    return x if (lo is None and hi is None) else torch.clamp(x, lo, hi)


class RunningNorm:
    def __init__(self, *, shape: Union[tuple, int], dtype=torch.float32, device=None, min_variance: float = 1e-2, clip: Optional[tuple] = None):
        self._shape = torch.Size(shape) if isinstance(shape, Iterable) else torch.Size([int(shape)])
        self._ndim = len(self._shape)
        self._dtype = dtype if isinstance(dtype, torch.dtype) else getattr(torch, str(dtype).replace("torch.", ""))
        self._device = torch.device("cpu" if device is None else device)
        self._min_variance = float(min_variance)
        self._lb, self._ub = (None, None) if clip is None else (float(clip[0]), float(clip[1]))
        self._sum: Optional[torch.Tensor] = None
        self._sum_of_squares: Optional[torch.Tensor] = None
        self._count: Union[int, torch.Tensor] = 0  # a 1-element int64 device tensor once CUDA batches have been seen

    # ------------------------------------------------------------------ properties
    @property
    def device(self) -> torch.device:
        return self._device

    @property
    def dtype(self) -> torch.dtype:
        return self._dtype

    @property
    def shape(self) -> tuple:
        return self._shape

    @property
    def min_variance(self) -> float:
        return self._min_variance

    @property
    def low(self) -> Optional[float]:
        return self._lb

    @property
    def high(self) -> Optional[float]:
        return self._ub

    @property
    def sum(self) -> Optional[torch.Tensor]:
        return self._sum

    @property
    def sum_of_squares(self) -> Optional[torch.Tensor]:
        return self._sum_of_squares

    @property
    def count(self) -> int:
        """Number of observations seen (reads the device counter: synchronises when the statistics live on a GPU)."""
        return int(self._count)

    @property
    def count_tensor(self) -> torch.Tensor:
        """The count as a 1-element int64 tensor on `device` (no synchronisation)."""
        if not isinstance(self._count, torch.Tensor):
            self._count = torch.tensor([self._count], dtype=torch.int64, device=self._device)
        return self._count

    def _has_data(self) -> bool:
        return self._sum is not None

    def reset(self):
        self._sum = self._sum_of_squares = None
        self._count = 0

    def to(self, device) -> "RunningNorm":
        device = torch.device(device)
        if device == self._device:
            return self
        new = RunningNorm(shape=self._shape, dtype=self._dtype, device=device, min_variance=self._min_variance,
                          clip=None if self._lb is None else (self._lb, self._ub))
        if self._has_data():
            new._sum, new._sum_of_squares = self._sum.to(device), self._sum_of_squares.to(device)
            new._count = self._count.to(device) if isinstance(self._count, torch.Tensor) else self._count
        return new

    def __copy__(self) -> "RunningNorm":
        return deepcopy(self)

    def __repr__(self) -> str:
        return f"<{type(self).__name__}, count: {self.count}>"

    # ------------------------------------------------------------------ update
    def _verify(self, x) -> torch.Tensor:
        x = torch.as_tensor(x, dtype=self._dtype, device=self._device)
        if x.ndim == self._ndim:
            if x.shape != self._shape:
                raise ValueError(f"This RunningNorm instance was initialized with shape: {self._shape}. However, the provided tensor has an"
                                 f" incompatible shape: {x.shape}.")
        elif x.ndim == self._ndim + 1:
            if x.shape[1:] != self._shape:
                raise ValueError(f"This RunningNorm instance was initialized with shape: {self._shape}. The provided tensor is shaped"
                                 f" {x.shape}. Accepting the tensor's leftmost dimension as the batch size, the remaining shape is"
                                 f" incompatible: {x.shape[1:]}")
        else:
            raise ValueError(f"This RunningNorm instance was initialized with shape: {self._shape}. The provided tensor is shaped {x.shape}."
                             f" The number of dimensions of the given tensor is incompatible.")
        return x

    def _accumulate(self, s1: torch.Tensor, s2: torch.Tensor, n):
        if self._has_data():
            self._sum += s1
            self._sum_of_squares += s2
            self._count = self._count + n
        else:
            self._sum, self._sum_of_squares = s1.clone(), s2.clone()
            self._count = n.clone() if isinstance(n, torch.Tensor) else n

    @torch.no_grad()
    def update(self, x, mask=None, *, verify: bool = True):
        """Add one observation, a batch of observations (optionally only the rows where `mask` is True), or the contents of
        another RunningNorm (runningnorm.py:229-410)."""
        if isinstance(x, RunningNorm):
            if mask is not None:
                raise ValueError("The `mask` argument is expected as None if the first argument is a RunningNorm.")
            if x._shape != self._shape:
                raise ValueError(f"The RunningNorm to be updated has the shape {self._shape}, the other one {x._shape}: incompatible.")
            if x._has_data():
                n = x._count.to(self._device) if isinstance(x._count, torch.Tensor) else x._count
                self._accumulate(x._sum.to(self._device, self._dtype), x._sum_of_squares.to(self._device, self._dtype), n)
            return
        if verify:
            x = self._verify(x)
        if x.ndim == self._ndim:
            if mask is not None:
                raise ValueError("The `mask` argument is expected as None if the first argument is a single observation.")
            self._accumulate(x, x.square(), 1)
            return
        if x.ndim != self._ndim + 1:
            raise ValueError(f"Invalid shape: {x.shape}")
        if mask is not None:
            mask = torch.as_tensor(mask, dtype=torch.bool, device=self._device)
            if mask.ndim != 1:
                raise ValueError(f"The `mask` tensor was expected as a 1-dimensional tensor. However, its shape is {mask.shape}.")
            if len(mask) != x.shape[0]:
                raise ValueError(f"The batch size of the observations is {x.shape[0]}, the `mask` has an incompatible length: {len(mask)}.")
        if self._ndim == 1 and ops.uses_kernels(x) and x.stride(-1) == 1 and x.shape[0] > 0:
            # K4, raw-moments form: s1 = sum_i w_i x_i, s2 = sum_i w_i x_i^2 with w = the mask (or ones), one pass, no host sync
            d = self._shape[0]
            w = torch.ones(x.shape[0], dtype=torch.float32, device=x.device) if mask is None else mask.to(torch.float32)
            zero, one = self._constants(d, x.device)
            s1, s2 = ops.grad(ops.GRAD_MOMENTS, x, w, zero, one, 1.0, 1.0)
            n = (torch.full((1,), x.shape[0], dtype=torch.int64, device=x.device) if mask is None
                 else mask.sum(dtype=torch.int64).reshape(1))
            if not isinstance(self._count, torch.Tensor):
                self._count = torch.tensor([self._count], dtype=torch.int64, device=x.device)
            self._accumulate(s1, s2, n)
            return
        if mask is not None:
            n = int(mask.sum())
            x = x * mask.to(self._dtype).reshape((x.shape[0],) + (1,) * (x.ndim - 1))
        else:
            n = x.shape[0]
        self._accumulate(x.sum(dim=0), x.square().sum(dim=0), n)

    def _constants(self, d: int, device) -> tuple:
        c = self.__dict__.get("_zero_one")
        if c is None or c[0].numel() != d or c[0].device != device:
            c = self.__dict__["_zero_one"] = (torch.zeros(d, dtype=torch.float32, device=device), torch.ones(d, dtype=torch.float32, device=device))
        return c

    # ------------------------------------------------------------------ statistics / normalisation
    @property
    @torch.no_grad()
    def stats(self) -> CollectedStats:
        n = self._count.to(self._dtype) if isinstance(self._count, torch.Tensor) else self._count
        mean = self._sum / n
        variance = torch.clamp(self._sum_of_squares / n - mean.square(), min=self._min_variance)
        return CollectedStats(mean=mean, stdev=torch.sqrt(variance))

    @property
    def mean(self) -> torch.Tensor:
        return self.stats.mean

    @property
    def stdev(self) -> torch.Tensor:
        return self.stats.stdev

    @torch.no_grad()
    def normalize(self, x, *, result_as_numpy: Optional[bool] = None, verify: bool = True):
        if not self._has_data():
            raise ValueError("Cannot do normalization because no data is collected yet.")
        if result_as_numpy is None:
            result_as_numpy = not isinstance(x, torch.Tensor)
        if verify:
            x = self._verify(x)
        mean, stdev = self.stats
        result = _clamp((x - mean) / stdev, self._lb, self._ub)
        return result.cpu().numpy() if result_as_numpy else result

    @torch.no_grad()
    def update_and_normalize(self, x, mask=None):
        as_numpy = not isinstance(x, torch.Tensor)
        x = self._verify(x)
        self.update(x, mask, verify=False)
        result = self.normalize(x, verify=False, result_as_numpy=False)
        return result.cpu().numpy() if as_numpy else result

    def to_layer(self) -> "ObsNormLayer":
        mean, stdev = self.stats
        return ObsNormLayer(mean=mean, stdev=stdev, low=self._lb, high=self._ub)


class ObsNormLayer(nn.Module):
    """clamp((x - mean) / stdev, low, high) as a module, for exporting a trained policy (runningnorm.py:583-613)."""

    def __init__(self, mean: torch.Tensor, stdev: torch.Tensor, low: Optional[float] = None, high: Optional[float] = None):
        super().__init__()
        self.register_buffer("_mean", mean)
        self.register_buffer("_stdev", stdev)
        self._lb = None if low is None else float(low)
        self._ub = None if high is None else float(high)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _clamp((x - self._mean) / self._stdev, self._lb, self._ub)
