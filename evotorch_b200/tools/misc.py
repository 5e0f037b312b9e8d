"""Small tensor helpers on the hot path (mirrors the named functions of evotorch.tools.misc)."""

from __future__ import annotations

import math
from collections.abc import Iterable
from numbers import Integral, Real
from typing import Any, Optional, Union

import numpy as np
import torch

RealOrVector = Union[float, int, torch.Tensor, list, tuple]

_DTYPE_NAMES = {"float16": torch.float16, "half": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32, "float": torch.float32,
                "float64": torch.float64, "double": torch.float64, "int8": torch.int8, "uint8": torch.uint8, "int16": torch.int16,
                "short": torch.int16, "int32": torch.int32, "int": torch.int32, "int64": torch.int64, "long": torch.int64, "bool": torch.bool}


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
    name = str(getattr(dtype, "__name__", dtype)).replace("torch.", "")  # "float32", np.float32, np.dtype("float32"), "torch.float32"
    if name in _DTYPE_NAMES:
        return _DTYPE_NAMES[name]
    raise TypeError(f"cannot interpret {dtype!r} as a torch dtype")


def extract_generator(generator: Any) -> Optional[torch.Generator]:
    """A torch.Generator, or any object with a `.generator` attribute such as a Problem (tools/misc.py:1523-1533)."""
    if generator is None or isinstance(generator, torch.Generator):
        return generator
    return generator.generator


def _shape_of(size: tuple) -> tuple:
    """`f(3, 4)`, `f((3, 4))` and `f(torch.Size([3, 4]))` all mean the shape (3, 4); no argument means a scalar."""
    if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
        return tuple(int(n) for n in size[0])
    return tuple(int(n) for n in size)


def _new_or_out(size: tuple, out: Optional[torch.Tensor], dtype, device, default_dtype=torch.float32) -> torch.Tensor:
    """The tensor a maker works on: a new one of the requested shape, or `out` (then shape / dtype / device must be omitted)."""
    if out is not None:
        if len(size) >= 1:
            raise ValueError(f"When `out` is provided (i.e. not None), the positional `size` arguments were not expected."
                             f" However, `size` arguments were received as {size!r}.")
        if dtype is not None or device is not None:
            raise ValueError("When `out` is provided (i.e. not None), the arguments `dtype` and `device` are expected as None.")
        return out
    return torch.empty(_shape_of(size), dtype=default_dtype if dtype is None else to_torch_dtype(dtype), device="cpu" if device is None else device)


def make_tensor(data: Any, *, dtype=None, device=None, read_only: bool = False) -> torch.Tensor:
    """A NEW tensor holding `data` (tools/misc.py:1138-1206).  `read_only` is accepted for signature compatibility; this package
    has no read-only tensor subclass, the returned tensor is an ordinary one."""
    dt = None if dtype is None else to_torch_dtype(dtype)
    if isinstance(data, torch.Tensor):
        return data.detach().clone().to(dtype=dt if dt is not None else data.dtype, device=data.device if device is None else device)
    return torch.tensor(data, dtype=dt, device="cpu" if device is None else device)


def make_empty(*size, dtype=None, device=None) -> torch.Tensor:
    return _new_or_out(size, None, dtype, device)


def make_zeros(*size, out: Optional[torch.Tensor] = None, dtype=None, device=None) -> torch.Tensor:
    return _new_or_out(size, out, dtype, device).zero_()


def make_ones(*size, out: Optional[torch.Tensor] = None, dtype=None, device=None) -> torch.Tensor:
    return _new_or_out(size, out, dtype, device).fill_(1)


def make_nan(*size, out: Optional[torch.Tensor] = None, dtype=None, device=None) -> torch.Tensor:
    return _new_or_out(size, out, dtype, device).fill_(float("nan"))


def make_I(size: Optional[int] = None, *, out: Optional[torch.Tensor] = None, dtype=None, device=None) -> torch.Tensor:
    """An n x n identity matrix (n = `size`), or `out` (a square matrix) turned into one (tools/misc.py:1456-1537)."""
    if isinstance(size, (tuple, list)):
        if len(size) != 1:
            raise ValueError(f"When the size argument is given as a tuple, `make_I(...)` expects the tuple to have only one element. The given tuple is {size}.")
        size = size[0]
    if size is None:
        if out is None:
            raise ValueError("`make_I(...)` needs either `size` or `out`")
        target = _new_or_out((), out, dtype, device)
    else:
        n = int(size)
        target = _new_or_out((n, n), None, dtype, device) if out is None else _new_or_out((n, n), out, dtype, device)
    if target.ndim != 2 or target.shape[0] != target.shape[1]:
        raise ValueError(f"An identity matrix needs a square target, got the shape {tuple(target.shape)}")
    target.zero_()
    target.fill_diagonal_(1)
    return target


def make_randint(*size, n, out: Optional[torch.Tensor] = None, dtype=None, device=None, generator: Any = None) -> torch.Tensor:
    """Uniform random integers in [0, n - 1]; int64 by default, float dtypes receive the integers as floats (tools/misc.py:1758-1832)."""
    target = _new_or_out(size, out, dtype, device, default_dtype=torch.int64)
    gen = extract_generator(generator)
    kw = {} if gen is None else {"generator": gen}
    if target.dtype.is_floating_point:
        target.copy_(torch.randint(0, int(n), target.shape, device=target.device, dtype=torch.int64, **kw))
    else:
        target.random_(0, int(n), **kw)
    return target


def make_gaussian(*size, center=None, stdev=None, symmetric: bool = False, out: Optional[torch.Tensor] = None, dtype=None, device=None,
                  generator: Any = None) -> torch.Tensor:
    """Gaussian noise through torch's generator -- the `rng="torch"` sampler that reproduces the reference's population
    bit for bit on the same device (tools/misc.py:1663-1755): symmetric rows 2k / 2k+1 are (z_k*stdev)+center and
    ((-z_k)*stdev)+center.  The Philox sampler of the kernels is `ops.sample_eval`."""
    out = _new_or_out(size, out, dtype, device)
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
    """Uniform samples (tools/misc.py:1540-1660).  Floating point: lb + (ub - lb) * U[0, 1).  Integer dtypes: uniformly among the
    integers lb .. ub inclusive (0 / 1 without bounds).  bool: fair coin, forced where lb == ub."""
    out = _new_or_out(size, out, dtype, device)
    if (lb is None) != (ub is None):
        raise ValueError(f"Expected both `lb` and `ub` as None, or both `lb` and `ub` as not None. lb: {lb!r}. ub: {ub!r}.")
    gen = extract_generator(generator)
    kw = {} if gen is None else {"generator": gen}
    if lb is not None:
        lb = torch.as_tensor(lb, dtype=out.dtype, device=out.device)
        ub = torch.as_tensor(ub, dtype=out.dtype, device=out.device)
    if out.dtype == torch.bool:
        out.random_(**kw)
        if lb is not None:
            out[torch.broadcast_to(~lb & ~ub, out.shape)] = False
            out[torch.broadcast_to(lb & ub, out.shape)] = True
    elif not out.dtype.is_floating_point:
        out.random_(**kw)
        if lb is None:
            out %= 2
        else:
            out -= lb
            out %= (ub - lb) + 1
            out += lb
    else:
        out.uniform_(**kw)
        if lb is not None:
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


# ------------------------------------------------------------------------------------------------
# Small type / shape predicates and helpers of the reference's tools/misc.py that host code and user scripts lean on
# (tools/misc.py:100-700).  No object-dtype / ObjectArray support: numeric problems only.
# ------------------------------------------------------------------------------------------------
def is_dtype_object(dtype) -> bool:
    return dtype in ("object", "Any", "O") if isinstance(dtype, str) else (dtype is object or dtype is Any)


def to_numpy_dtype(dtype) -> np.dtype:
    if isinstance(dtype, torch.dtype):
        return torch.empty(0, dtype=dtype).numpy().dtype
    if is_dtype_object(dtype):
        return np.dtype(object)
    return dtype if isinstance(dtype, np.dtype) else np.dtype(dtype)


def is_dtype_bool(t) -> bool:
    return to_numpy_dtype(t).kind == "b"


def is_dtype_integer(t) -> bool:
    return to_numpy_dtype(t).kind in ("u", "i")


def is_dtype_float(t) -> bool:
    return to_numpy_dtype(t).kind == "f"


def is_dtype_real(t) -> bool:
    return to_numpy_dtype(t).kind in ("u", "i", "f")


def is_sequence(x: Any) -> bool:
    if isinstance(x, (str, bytes)):
        return False
    if isinstance(x, (np.ndarray, torch.Tensor)):
        return x.ndim > 0
    return isinstance(x, Iterable)


def _is_scalar_of(x: Any, python_kind, dtype_test) -> bool:
    if isinstance(x, (torch.Tensor, np.ndarray)):
        return x.ndim == 0 and dtype_test(x.dtype)
    return isinstance(x, python_kind)


def is_bool(x: Any) -> bool:
    return _is_scalar_of(x, (bool, np.bool_), is_dtype_bool)


def is_integer(x: Any) -> bool:
    return (not is_bool(x)) and _is_scalar_of(x, Integral, is_dtype_integer)


def is_real(x: Any) -> bool:
    return (not is_bool(x)) and _is_scalar_of(x, Real, is_dtype_real)


def _is_vector_of(x: Any, scalar_test, dtype_test) -> bool:
    if isinstance(x, (torch.Tensor, np.ndarray)):
        return x.ndim == 1 and dtype_test(x.dtype)
    return isinstance(x, Iterable) and all(scalar_test(item) for item in x)


def is_bool_vector(x: Any) -> bool:
    return _is_vector_of(x, is_bool, is_dtype_bool)


def is_integer_vector(x: Any) -> bool:
    return _is_vector_of(x, is_integer, is_dtype_integer)


def is_real_vector(x: Any) -> bool:
    return _is_vector_of(x, is_real, is_dtype_real)


def clip_tensor(x: torch.Tensor, lb=None, ub=None, ensure_copy: bool = True) -> torch.Tensor:
    """max(min(x, ub), lb) with scalar or tensor bounds; never returns `x` itself unless `ensure_copy=False`."""
    result = x
    if lb is not None:
        result = torch.max(result, torch.as_tensor(lb, dtype=x.dtype, device=x.device))
    if ub is not None:
        result = torch.min(result, torch.as_tensor(ub, dtype=x.dtype, device=x.device))
    return x.clone() if (ensure_copy and result is x) else result


def numpy_copy(x, dtype=None) -> np.ndarray:
    """An independent numpy copy of a tensor / array / sequence, optionally cast."""
    if isinstance(x, torch.Tensor):
        result = x.detach().cpu().clone().numpy()
    elif isinstance(x, np.ndarray):
        result = x.copy()
    else:
        return np.array(x, dtype=dtype)
    return result if dtype is None else result.astype(to_numpy_dtype(dtype))


def expect_none(msg_prefix: str, **kwargs):
    for name, value in kwargs.items():
        if value is not None:
            raise ValueError(f"{msg_prefix}: expected `{name}` as None, however, it was found to be {value!r}")


def empty_tensor_like(source, *, shape=None, length: Optional[int] = None, dtype=None, device=None) -> torch.Tensor:
    """An uninitialised tensor taking shape / dtype / device from `source` unless overridden; `length` overrides the leftmost
    dimension only (tools/misc.py `empty_tensor_like`)."""
    if not isinstance(source, torch.Tensor):
        raise TypeError(f"`source` is expected as a torch.Tensor (object arrays are not supported here), got {type(source)}")
    if length is not None and shape is not None:
        raise ValueError("`length` and `shape` cannot be used together")
    if length is not None:
        if source.ndim == 0:
            raise ValueError("`length` can only be used with a source tensor of at least 1 dimension")
        shape = (int(length),) + tuple(source.shape[1:])
    elif shape is None:
        shape = tuple(source.shape)
    elif not isinstance(shape, Iterable):
        shape = (int(shape),)
    return torch.empty(tuple(shape), dtype=source.dtype if dtype is None else to_torch_dtype(dtype),
                       device=source.device if device is None else device)


def clone(x: Any, *, memo: Optional[dict] = None) -> Any:
    """An independent copy of `x` (tools/misc.py:588): `deep_clone` with deep-copy as the fallback for unknown objects."""
    from .cloning import deep_clone

    return deep_clone(x, otherwise_deepcopy=True, memo={} if memo is None else memo)


def device_of(x: Any) -> torch.device:
    """Device of a tensor, of a module (its first parameter) or of any object with a `device` attribute (tools/misc.py:2014-2037)."""
    if isinstance(x, torch.nn.Module):
        for param in x.parameters():
            return param.device
        raise ValueError(f"Cannot determine the device of the module {x}")
    return x.device


def dtype_of(x: Any):
    """dtype of a tensor / array, of a module (its first parameter) or of any object with a `dtype` attribute (tools/misc.py:2040-2063)."""
    if isinstance(x, torch.nn.Module):
        for param in x.parameters():
            return param.dtype
        raise ValueError(f"Cannot determine the dtype of the module {x}")
    return x.dtype
