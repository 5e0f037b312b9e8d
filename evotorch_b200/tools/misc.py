"""Small tensor helpers on the hot path (mirrors the named functions of evotorch.tools.misc)."""

from __future__ import annotations

import math
from typing import Any, Optional, Union

import torch

RealOrVector = Union[float, int, torch.Tensor, list, tuple]

_DTYPE_NAMES = {"float16": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32, "float": torch.float32,
                "float64": torch.float64, "double": torch.float64, "int64": torch.int64, "int32": torch.int32, "bool": torch.bool}


def to_torch_dtype(dtype) -> torch.dtype:
    """Accept torch dtypes, their string names, numpy-style names and python types (tools/misc.py `to_torch_dtype`)."""
    if isinstance(dtype, torch.dtype):
        return dtype
    if dtype is float:
        return torch.float32
    if dtype is int:
        return torch.int64
    if dtype is bool:
        return torch.bool
    name = str(getattr(dtype, "__name__", dtype)).replace("torch.", "")
    if name in _DTYPE_NAMES:
        return _DTYPE_NAMES[name]
    raise TypeError(f"cannot interpret {dtype!r} as a torch dtype")


def extract_generator(generator: Any) -> Optional[torch.Generator]:
    """A torch.Generator, or any object with a `.generator` attribute such as a Problem (tools/misc.py:1523-1533)."""
    if generator is None or isinstance(generator, torch.Generator):
        return generator
    return generator.generator


def make_gaussian(*size, center=None, stdev=None, symmetric: bool = False, out: Optional[torch.Tensor] = None, dtype=None, device=None,
                  generator: Any = None) -> torch.Tensor:
    """Gaussian noise through torch's generator -- the `rng="torch"` sampler that reproduces the reference's population
    bit for bit on the same device (tools/misc.py:1663-1755): symmetric rows 2k / 2k+1 are (z_k*stdev)+center and
    ((-z_k)*stdev)+center.  The Philox sampler of the kernels is `ops.sample_eval`."""
    if out is None:
        out = torch.empty(*size, dtype=to_torch_dtype(dtype) if dtype is not None else torch.float32, device=device or "cpu")
    gen = extract_generator(generator)
    kw = {} if gen is None else {"generator": gen}
    if symmetric:
        if out.shape[0] % 2 != 0:
            raise ValueError(f"Symmetric sampling cannot be done if the leftmost dimension of the target tensor is odd: {tuple(out.shape)}")
        out[0::2, ...].normal_(**kw)
        out[1::2, ...] = out[0::2, ...]
        out[1::2, ...] *= -1
    else:
        out.normal_(**kw)
    if (center is None) != (stdev is None):
        raise ValueError("Please either specify none of `stdev` and `center`, or both of them.")
    if center is not None:
        out *= torch.as_tensor(stdev, dtype=out.dtype, device=out.device)
        out += torch.as_tensor(center, dtype=out.dtype, device=out.device)
    return out


def make_uniform(*size, lb=None, ub=None, out: Optional[torch.Tensor] = None, dtype=None, device=None, generator: Any = None) -> torch.Tensor:
    """Uniform samples in [lb, ub) (tools/misc.py:1540): out = lb + (ub-lb) * U[0,1)."""
    if out is None:
        out = torch.empty(*size, dtype=to_torch_dtype(dtype) if dtype is not None else torch.float32, device=device or "cpu")
    gen = extract_generator(generator)
    kw = {} if gen is None else {"generator": gen}
    out.uniform_(**kw)
    if (lb is None) != (ub is None):
        raise ValueError("Please either specify none of `lb` and `ub`, or both of them.")
    if lb is not None:
        lb = torch.as_tensor(lb, dtype=out.dtype, device=out.device)
        ub = torch.as_tensor(ub, dtype=out.dtype, device=out.device)
        out *= ub - lb
        out += lb
    return out


def modify_tensor(original: torch.Tensor, target: torch.Tensor, lb=None, ub=None, max_change=None, in_place: bool = False) -> torch.Tensor:
    """Move `original` towards `target` subject to bounds and a relative max-change limit (tools/misc.py:711-816):
    result = min(max(target, max(lb, o - |o|*c)), min(ub, o + |o|*c))."""
    if lb is None and ub is None and max_change is None:
        result = target
    else:
        def conv(x, name):
            t = x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=original.dtype, device=original.device)
            if t.ndim != 0 and t.shape != original.shape:
                raise IndexError(f"Argument {name}: shape mismatch. Shape of the original tensor: {original.shape}. Shape of {name}: {t.shape}.")
            return t

        lo = conv(float("-inf") if lb is None else lb, "lb")
        hi = conv(float("inf") if ub is None else ub, "ub")
        if max_change is not None:
            allowed = torch.abs(original) * conv(max_change, "max_change")
            lo = torch.max(lo, original - allowed)
            hi = torch.min(hi, original + allowed)
        result = torch.min(torch.max(target, lo), hi)
    if in_place:
        original[:] = result
        return original
    return result


def split_workload(workload: int, num_actors: int) -> list:
    """Near-equal integer split, the first `workload % num_actors` shares one larger (tools/misc.py:1113)."""
    base, extra = divmod(int(workload), int(num_actors))
    return [base + (1 if i < extra else 0) for i in range(num_actors)]


def stdev_from_radius(radius: float, solution_length: int) -> float:
    """sqrt(radius^2 / n) (tools/misc.py:1879)."""
    return math.sqrt((float(radius) ** 2) / int(solution_length))


def to_stdev_init(*, solution_length: int, stdev_init=None, radius_init=None):
    """Exactly one of stdev_init / radius_init (tools/misc.py:1925)."""
    if stdev_init is not None and radius_init is None:
        return stdev_init
    if stdev_init is None and radius_init is not None:
        return stdev_from_radius(radius_init, solution_length)
    if stdev_init is None:
        raise ValueError("Received both `stdev_init` and `radius_init` as None. Please provide a value either for `stdev_init` or for `radius_init`.")
    raise ValueError("Found both `stdev_init` and `radius_init` with values other than None. Please provide only one of them.")


def ensure_tensor_length_and_dtype(t: Any, length: int, dtype, about: Optional[str] = None, *, allow_scalar: bool = False,
                                   device=None) -> torch.Tensor:
    """Return `t` as a 1-D tensor of the given length/dtype/device; scalars are broadcast unless `allow_scalar`, in which
    case they are returned as 0-dim tensors (tools/misc.py:610)."""
    dtype = to_torch_dtype(dtype)
    t = torch.as_tensor(t, dtype=dtype, device=device)
    where = "" if about is None else f"{about}: "
    if t.ndim == 0:
        return t if allow_scalar else t.repeat(length)
    if t.ndim != 1 or len(t) != length:
        raise ValueError(f"{where}expected a vector of length {length}, got a tensor of shape {tuple(t.shape)}")
    return t
