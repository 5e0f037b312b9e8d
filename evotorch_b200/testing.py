"""Assertion helpers for tests written against this package (same names and behaviour as the reference's `evotorch.testing`,
testing.py:26-290, so that test code ports unchanged): tensors, arrays and plain sequences are all accepted."""

from __future__ import annotations

from numbers import Real
from typing import Any, Iterable, Optional, Union

import numpy as np
import torch


class TestingError(Exception):
    """Wrong use of a testing helper (as opposed to a failed assertion)."""

    __test__ = False  # not a pytest test class


def _as_numpy(x, dtype=None) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    if dtype is not None and isinstance(dtype, torch.dtype):
        dtype = torch.empty(0, dtype=dtype).numpy().dtype
    return np.asarray(x) if dtype is None else np.asarray(x, dtype=dtype)


def assert_allclose(actual, desired, *, rtol: Optional[float] = None, atol: Optional[float] = None, equal_nan: bool = True):
    """`actual` (cast to the dtype of `desired`) is close to `desired`; at least one tolerance has to be given."""
    if rtol is None and atol is None:
        raise TestingError("Both `rtol` and `atol` were found to be None. Please either specify `rtol`, `atol`, or both.")
    want = _as_numpy(desired)
    got = _as_numpy(actual, dtype=want.dtype)
    np.testing.assert_allclose(got, want, rtol=0.0 if rtol is None else rtol, atol=0.0 if atol is None else atol, equal_nan=bool(equal_nan))


def assert_almost_between(x, lb, ub, *, atol: Optional[float] = None):
    """Every element of `x` lies in [lb - atol, ub + atol] (bounds broadcast to the shape of `x`)."""
    x = _as_numpy(x)
    slack = 0.0 if atol is None else float(atol)
    lo = np.broadcast_to(_as_numpy(lb), x.shape).astype(x.dtype) - slack
    hi = np.broadcast_to(_as_numpy(ub), x.shape).astype(x.dtype) + slack
    assert np.all((x >= lo) & (x <= hi)), (f"The provided array is not within the desired boundaries. Provided array: {x}."
                                           f" Lower bound: {lb}. Upper bound: {ub}. Absolute tolerance: {atol}.")


def _numpy_dtype(dtype) -> np.dtype:
    if dtype == "Any" or dtype is Any:
        return np.dtype(object)
    if isinstance(dtype, torch.dtype):
        return torch.empty(0, dtype=dtype).numpy().dtype
    return np.dtype(dtype)


def assert_dtype_matches(x, dtype: Union[str, type, np.dtype, torch.dtype]):
    actual, expected = _numpy_dtype(x.dtype), _numpy_dtype(dtype)
    assert actual == expected, f"dtype mismatch. Encountered dtype: {actual}, expected dtype: {expected}"


def assert_shape_matches(x, shape: Union[tuple, int]):
    actual = tuple(x.shape) if hasattr(x, "shape") else tuple(np.asarray(x).shape)
    expected = tuple(shape) if isinstance(shape, Iterable) else (int(shape),)
    assert actual == expected, f"Encountered a shape mismatch. Shape of the tensor: {actual}. Expected shape: {expected}"


def assert_eachclose(x, value: Any, *, rtol: Optional[float] = None, atol: Optional[float] = None):
    """Every element of `x` is close to the scalar `value`."""
    x = _as_numpy(x)
    assert_allclose(x, np.full_like(x, value if isinstance(value, Real) else float(value)), rtol=rtol, atol=atol)
