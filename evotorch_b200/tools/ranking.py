"""Fitness -> utility transforms (mirrors evotorch.tools.ranking of the reference, tools/ranking.py:24-216).

CUDA float32 inputs are ranked by the hand-written radix-sort kernel (K3, csrc/evok_rank.cu); every other
tensor (CPU problems such as BASELINE config 1, float64, ...) takes the generic torch implementation
below.  Both paths share one contract that is *stricter* than the reference's: the sort is stable, i.e. equal
fitnesses keep ascending index order (the reference calls argsort without stable=True, so its tie order is
unspecified; on tie-free inputs all three agree bit for bit).
"""

from __future__ import annotations

from typing import Iterable

import torch

from .. import ops
from .readonlytensor import as_plain_tensor


def _sorted_positions(x: torch.Tensor, higher_is_better: bool) -> torch.Tensor:
    return torch.argsort(x, descending=(not higher_is_better), stable=True)


def centered(fitnesses: torch.Tensor, *, higher_is_better: bool = True) -> torch.Tensor:
    """Linearly spaced utilities in [-0.5, 0.5]; the best solution gets +0.5 (tools/ranking.py:24-53)."""
    with torch.no_grad():
        fitnesses = as_plain_tensor(fitnesses)
        x = fitnesses.reshape(-1)
        if ops.uses_kernels(x):
            return ops.rank(x.contiguous(), "centered", higher_is_better).reshape(fitnesses.shape)
        n = len(x)
        order = _sorted_positions(x, higher_is_better)
        table = (torch.arange(n, dtype=x.dtype, device=x.device) / (n - 1)) - 0.5
        out = torch.empty_like(x)
        out[order] = table
        return out.reshape(fitnesses.shape)


def linear(fitnesses: torch.Tensor, *, higher_is_better: bool = True) -> torch.Tensor:
    """Linearly spaced utilities in [0, 1] (tools/ranking.py:56-81)."""
    with torch.no_grad():
        fitnesses = as_plain_tensor(fitnesses)
        x = fitnesses.reshape(-1)
        if ops.uses_kernels(x):
            return ops.rank(x.contiguous(), "linear", higher_is_better).reshape(fitnesses.shape)
        n = len(x)
        order = _sorted_positions(x, higher_is_better)
        table = torch.arange(n, dtype=x.dtype, device=x.device) / (n - 1)
        out = torch.empty_like(x)
        out[order] = table
        return out.reshape(fitnesses.shape)


def nes(fitnesses: torch.Tensor, *, higher_is_better: bool = True) -> torch.Tensor:
    """NES utilities max(0, ln(n/2+1) - ln(n-p)), normalised to sum 1, minus 1/n (tools/ranking.py:84-124)."""
    with torch.no_grad():
        fitnesses = as_plain_tensor(fitnesses)
        x = fitnesses.reshape(-1)
        if ops.uses_kernels(x):
            return ops.rank(x.contiguous(), "nes", higher_is_better).reshape(fitnesses.shape)
        n = len(x)
        nf = torch.tensor(n, dtype=x.dtype, device=x.device)
        steps = torch.arange(n, dtype=x.dtype, device=x.device)
        table = torch.clamp_min(torch.log(nf / 2.0 + 1.0) - torch.log(nf - steps), 0.0)
        order = _sorted_positions(x, higher_is_better)
        position = torch.empty(n, dtype=order.dtype, device=x.device)
        position[order] = torch.arange(n, dtype=order.dtype, device=x.device)
        utils = table[position]
        utils = utils / torch.sum(utils)
        utils = utils - 1 / nf
        return utils.reshape(fitnesses.shape)


def normalized(fitnesses: torch.Tensor, *, higher_is_better: bool = True) -> torch.Tensor:
    """Standardised (zero mean, unit unbiased std) fitnesses, negated for minimisation (tools/ranking.py:127-160)."""
    with torch.no_grad():
        fitnesses = as_plain_tensor(fitnesses)
        if ops.uses_kernels(fitnesses) and fitnesses.ndim == 1:
            return ops.rank(fitnesses.contiguous(), "normalized", higher_is_better)
        g = fitnesses if higher_is_better else -fitnesses
        return (g - torch.mean(g)) / torch.std(g)


def raw(fitnesses: torch.Tensor, *, higher_is_better: bool = True) -> torch.Tensor:
    """The fitnesses themselves (negated for minimisation) (tools/ranking.py:163-183)."""
    return fitnesses if higher_is_better else -fitnesses


rankers = {"nes": nes, "centered": centered, "linear": linear, "normalized": normalized, "raw": raw}


def rank(fitnesses: Iterable[float], ranking_method: str, *, higher_is_better: bool) -> torch.Tensor:
    """Dispatch by name; KeyError on an unknown method exactly like the reference (tools/ranking.py:189-216)."""
    fitnesses = as_plain_tensor(torch.as_tensor(fitnesses))
    return rankers[ranking_method](fitnesses, higher_is_better=higher_is_better)
