"""Shared helpers of the functional (ask/tell) API (reference: algorithms/functional/misc.py:26-163).

Batching convention (reference: `expects_ndim`, decorators.py:613): every tensor argument may carry extra LEFTMOST
dimensions; they are broadcast against each other and each batch item is an independent search.  On CUDA float32 every
batch item runs through the same K1-K5 kernels as the object-oriented searchers (one launch chain per item, all on the
current stream); anywhere else the generic torch path of those classes is used.

Scalar hyper-parameters (learning rates, momentum, ...) are stored as HOST tensors of the centre's dtype: the kernels take
them as launch arguments, so keeping them on the host means a `tell` never synchronises with the device.  (0-dim host
tensors still combine freely with device tensors in user code.)
"""

from __future__ import annotations

import importlib
from typing import Callable, Iterable, NamedTuple, Optional, Union

import torch

from ... import ops


def host_scalar(x, dtype: torch.dtype) -> torch.Tensor:
    """A (possibly batched) scalar hyper-parameter as a host tensor."""
    if isinstance(x, torch.Tensor):
        return x.detach().to(device="cpu", dtype=dtype)
    return torch.as_tensor(x, dtype=dtype, device="cpu")


def batch_shape_of(*pairs) -> torch.Size:
    """Broadcast batch shape of (tensor, core_ndim) pairs."""
    shapes = [tuple(t.shape[: t.ndim - nd]) for t, nd in pairs if t is not None]
    return torch.broadcast_shapes(*shapes) if shapes else torch.Size()


def expand_to(t: torch.Tensor, batch: torch.Size, core_ndim: int) -> torch.Tensor:
    """`t` with its batch dimensions broadcast to `batch` (a view)."""
    core = tuple(t.shape[t.ndim - core_ndim:]) if core_ndim else ()
    return t.expand(tuple(batch) + core)


def flat_items(t: torch.Tensor, batch: torch.Size, core_ndim: int) -> torch.Tensor:
    """`t` broadcast to `batch` and flattened to (B, *core) -- B = prod(batch), 1 when not batched."""
    core = tuple(t.shape[t.ndim - core_ndim:]) if core_ndim else ()
    return expand_to(t, batch, core_ndim).reshape((-1,) + core)


def scalar_items(t: torch.Tensor, batch: torch.Size) -> list:
    """Python floats of a (batched) host scalar, one per batch item."""
    return flat_items(t, batch, 0).tolist()


def on_kernels(*tensors) -> bool:
    return all(ops.uses_kernels(t) for t in tensors)


def get_stdev_init(*, center_init: torch.Tensor, stdev_init=None, radius_init=None) -> torch.Tensor:
    """Initial standard deviation from `stdev_init` (scalar / vector / batch of vectors) or from `radius_init`, the
    Euclidean norm of a constant stdev vector (misc.py:78-163)."""
    if not isinstance(center_init, torch.Tensor):
        raise TypeError("`center_init` is expected as a tensor")
    dtype, device, length = center_init.dtype, center_init.device, center_init.shape[-1]
    if stdev_init is None and radius_init is None:
        raise ValueError("Both `stdev_init` and `radius_init` are None. Please provide one of them.")
    if stdev_init is not None and radius_init is not None:
        raise ValueError("Both `stdev_init` and `radius_init` are given. Please specify only one of them.")
    if stdev_init is not None:
        stdev = torch.as_tensor(stdev_init, dtype=dtype, device=device)
        if stdev.ndim == 0:
            return stdev.repeat(length)
        if stdev.shape[-1] != length:
            raise ValueError(f"The shape of `stdev_init` {tuple(stdev.shape)} is incompatible with the solution length {length}"
                             f" implied by `center_init` {tuple(center_init.shape)}.")
        return stdev
    radius = torch.as_tensor(radius_init, dtype=dtype, device=device)
    element = torch.sqrt((radius**2) / length)
    return element[..., None] * torch.ones(length, dtype=dtype, device=device)


def vector_like_center(x: Union[float, Iterable], name: str, center: torch.Tensor) -> torch.Tensor:
    length = center.shape[-1]
    x = torch.as_tensor(x, dtype=center.dtype, device=center.device)
    if x.ndim == 0:
        return x.repeat(length)
    if x.shape[-1] != length:
        raise ValueError(f"`{name}` has an incompatible length. The length of `{name}`: {x.shape[-1]},"
                         f" but the solution length implied by the provided `center_init` is {length}.")
    return x


class OptimizerFunctions(NamedTuple):
    initialize: Callable
    ask: Callable
    tell: Callable


_OPTIMIZER_ALIASES = {"clipup": "clipup", "adam": "adam", "sgd": "sgd", "sga": "sgd", "momentum": "sgd"}


def get_functional_optimizer(optimizer: Union[str, tuple]) -> OptimizerFunctions:
    """Resolve "clipup" | "adam" | "sgd" (aliases "sga", "momentum") or a user triple (init, ask, tell) into the three
    functions of a functional optimizer (misc.py:26-75)."""
    if isinstance(optimizer, str):
        name = _OPTIMIZER_ALIASES.get(optimizer)
        if name is None:
            raise ValueError(f"Unrecognized functional optimizer name: {optimizer}")
        module = importlib.import_module(f"{__package__}.func{name}")
        return OptimizerFunctions(getattr(module, name), getattr(module, f"{name}_ask"), getattr(module, f"{name}_tell"))
    if isinstance(optimizer, Iterable):
        return OptimizerFunctions(*optimizer)
    raise TypeError(f"`get_functional_optimizer(...)` received an unrecognized argument: {optimizer!r} (of type {type(optimizer)})")


def draw_philox_seed() -> int:
    """A fresh 62-bit Philox key from torch's default (host) generator: `torch.manual_seed` makes the asks reproducible and
    drawing it never touches the device."""
    return int(torch.randint(0, 2**62, (1,), dtype=torch.int64).item())
