"""Built-in vectorised objective functions with a fused evaluation kernel.

Each object is an ordinary vectorised fitness function (call it with an N x D tensor, get N fitnesses; it is marked
`__evotorch_vectorized__` like functions decorated with the reference's `@vectorized`, decorators.py:549) and
additionally carries `evok_objective_id`.  When a Problem is built around one of them, the Gaussian searchers
evaluate the population *inside* the sampling kernel (K1+K2 fused, csrc/evok_sample_eval.cu) so the N x D matrix is
written once and never re-read for evaluation.  Called directly on a CUDA fp32 population they run the stand-alone
row-reduction kernel (K2); on any other tensor the plain torch expression (same formula as the reference's README
example, README.md:86-89).
"""

from __future__ import annotations

import math
from typing import Callable

import torch

from . import ops


class BuiltinObjective:
    __evotorch_vectorized__ = True

    def __init__(self, name: str, objective_id: int, torch_fn: Callable):
        self.__name__ = name
        self.name = name
        self.evok_objective_id = objective_id
        self._torch_fn = torch_fn

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if x.ndim == 1:
            return self._torch_fn(x.unsqueeze(0))[0]
        if ops.uses_kernels(x) and x.ndim == 2 and x.stride(1) == 1:
            return ops.evaluate(self.evok_objective_id, x)
        return self._torch_fn(x)

    def __repr__(self) -> str:
        return f"<evotorch_b200.objectives.{self.name}>"


def _sphere(x: torch.Tensor) -> torch.Tensor:
    return torch.sum(x**2, dim=-1)


def _rastrigin(x: torch.Tensor) -> torch.Tensor:
    n = x.shape[-1]
    return 10 * n + torch.sum((x**2) - 10 * torch.cos(2 * math.pi * x), dim=-1)


def _ackley(x: torch.Tensor) -> torch.Tensor:
    n = x.shape[-1]
    return (-20.0 * torch.exp(-0.2 * torch.sqrt(torch.sum(x**2, dim=-1) / n)) - torch.exp(torch.sum(torch.cos(2 * math.pi * x), dim=-1) / n)
            + 20.0 + math.e)


sphere = BuiltinObjective("sphere", ops.OBJ_SPHERE, _sphere)
rastrigin = BuiltinObjective("rastrigin", ops.OBJ_RASTRIGIN, _rastrigin)
ackley = BuiltinObjective("ackley", ops.OBJ_ACKLEY, _ackley)

__all__ = ["sphere", "rastrigin", "ackley", "BuiltinObjective"]
