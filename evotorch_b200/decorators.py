"""Decorators that describe how a fitness function wants to be called (reference: evotorch/decorators.py:31-960).

    @vectorized        the function takes the whole N x D population and returns N fitnesses           (decorators.py:549)
    @rowwise           the function is written for ONE solution; it is vmapped over any leading dims  (:877)
    @expects_ndim(..)  general form of the same idea for several arguments                             (:613)
    @on_device(dev) / @on_cuda / @on_cuda(i) / @on_aux_device      where the function wants its inputs (:211-547)
    @pass_info         the function accepts extra keyword information from the problem               (:170)

`Problem` honours the markers these decorators leave on the function (`__evotorch_vectorized__`, `__evotorch_on_device__` +
`.device`, `__evotorch_on_aux_device__`, `__evotorch_pass_info__`).
"""

from __future__ import annotations

from numbers import Number
from typing import Callable, Optional

import numpy as np
import torch
from torch.func import vmap


def _marker(attribute: str, args: tuple, name: str) -> Callable:
    """Supports both `@name` and `@name()`: sets `fn.<attribute> = True`."""

    def mark(fn: Callable) -> Callable:
        setattr(fn, attribute, True)
        return fn

    if len(args) == 0:
        return mark
    if len(args) == 1 and callable(args[0]):
        return mark(args[0])
    raise TypeError(f"`{name}` received invalid arguments: {args!r}")


def vectorized(*args) -> Callable:
    """The decorated fitness function receives all solutions at once (a 2-D tensor) and returns one fitness per row."""
    return _marker("__evotorch_vectorized__", args, "vectorized")


def pass_info(*args) -> Callable:
    """The decorated callable accepts problem information (e.g. observation / action lengths) as extra keyword arguments."""
    return _marker("__evotorch_pass_info__", args, "pass_info")


def on_device(device) -> Callable:
    """The decorated fitness function wants its input on `device`; `fn.device` can be read and reassigned later."""
    device = torch.device(device)

    def mark(fn: Callable) -> Callable:
        fn.__evotorch_on_device__ = True
        fn.device = device
        return fn

    return mark


def on_cuda(*args) -> Callable:
    """`@on_cuda`, `@on_cuda()` -> device "cuda"; `@on_cuda(2)` -> device "cuda:2"."""
    if len(args) == 1 and callable(args[0]) and not isinstance(args[0], (int, np.integer)):
        return on_device("cuda")(args[0])
    if len(args) == 0:
        return on_device("cuda")
    if len(args) == 1:
        return on_device(torch.device("cuda", int(args[0])))
    raise TypeError(f"`on_cuda` received invalid arguments: {args!r}")


def on_aux_device(*args) -> Callable:
    """The decorated fitness function wants its input on the problem's auxiliary device (the first visible GPU if there is one,
    else the cpu): populations can then live on the host while evaluation happens on the accelerator."""
    return _marker("__evotorch_on_aux_device__", args, "on_aux_device")


def expects_ndim(*expected_ndims, allow_smaller_ndim: bool = False, randomness: str = "error") -> Callable:
    """Declare how many dimensions each positional argument of a function is written for.  Extra leftmost dimensions of the
    actual arguments are batch dimensions: they are broadcast against each other (aligned on the right) and the function is
    vmapped over them.  `None` marks an argument that is passed through untouched.  Numbers and numpy arrays are converted to
    tensors (dtype / device of the tensor arguments).

        @expects_ndim(2, 1)
        def f(a, b): ...                  # or: g = expects_ndim(f, (2, 1))
    """
    if len(expected_ndims) == 2 and callable(expected_ndims[0]) and isinstance(expected_ndims[1], (tuple, list)):
        fn, dims = expected_ndims
        return expects_ndim(*dims, allow_smaller_ndim=allow_smaller_ndim, randomness=randomness)(fn)
    for n in expected_ndims:
        if n is not None and (not isinstance(n, (int, np.integer)) or n < 0):
            raise TypeError(f"`expects_ndim` expects non-negative integers or None, got {n!r}")
    wanted = tuple(None if n is None else int(n) for n in expected_ndims)

    def decorate(fn: Callable) -> Callable:
        def call(*args):
            if len(args) != len(wanted):
                raise TypeError(f"The function decorated with `expects_ndim` was expecting {len(wanted)} positional arguments, got {len(args)}")
            tensors = [a for a, n in zip(args, wanted) if n is not None and isinstance(a, torch.Tensor)]
            ready = []
            for i, (a, n) in enumerate(zip(args, wanted)):
                if n is not None and not isinstance(a, torch.Tensor):
                    if isinstance(a, (bool, np.bool_)):
                        a = torch.as_tensor(a, dtype=torch.bool, device=tensors[0].device if tensors else None)
                    elif isinstance(a, Number):
                        if not tensors:
                            raise TypeError(f"Cannot decide the dtype / device for the scalar argument {a!r}: no tensor argument to follow")
                        a = torch.as_tensor(a, dtype=tensors[0].dtype, device=tensors[0].device)
                    elif isinstance(a, np.ndarray):
                        a = torch.as_tensor(a)
                    else:
                        raise TypeError(f"Received an argument of unexpected type: {a} (of type {type(a)})")
                if n is not None and a.ndim < n and not allow_smaller_ndim:
                    raise ValueError(f"The argument with index {i} has the shape {a.shape}, having {a.ndim} dimensions."
                                     f" However, it was expected as a tensor with {n} dimensions.")
                ready.append(a)
            extras = [0 if n is None else max(a.ndim - n, 0) for a, n in zip(ready, wanted)]
            wrapped = fn
            # one vmap per batch dimension, innermost first; an argument takes part while it still has batch dimensions left,
            # so batch dimensions align on the right like broadcasting
            for level in range(max(extras, default=0)):
                in_dims = tuple(0 if e > level else None for e in extras)
                wrapped = vmap(wrapped, in_dims=in_dims, randomness=randomness)
            return wrapped(*ready)

        call.__wrapped__ = fn
        call.__name__ = getattr(fn, "__name__", "expects_ndim_decorated")
        call.__doc__ = getattr(fn, "__doc__", None)
        return call

    return decorate


def rowwise(*args, randomness: str = "error") -> Callable:
    """Write the fitness function for ONE solution (a 1-D tensor); the decorated function accepts any number of leading batch
    dimensions and is marked `@vectorized`, so a `Problem` evaluates whole populations with it in one call."""

    def decorate(fn: Callable) -> Callable:
        decorated = expects_ndim(fn, (1,), randomness=randomness)
        decorated.__evotorch_vectorized__ = True
        return decorated

    if len(args) == 0:
        return decorate
    if len(args) == 1 and callable(args[0]):
        return decorate(args[0])
    raise TypeError("`rowwise` received invalid number of positional arguments")
