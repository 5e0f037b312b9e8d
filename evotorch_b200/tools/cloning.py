"""Structure-preserving deep cloning (reference: tools/cloning.py:25-330).

`deep_clone(x)` walks containers and objects like `copy.deepcopy`, with three differences that matter for populations and
searchers: tensors are cloned with `.clone()` -- a slice of a 40 GB population becomes a small independent tensor instead of
dragging a copy of the whole storage along; shared references and cycles are preserved through the `memo` dictionary; and
read-only data stays read-only (a ReadOnlyTensor clones to a ReadOnlyTensor, a non-writeable numpy array stays non-writeable).
`Clonable` gives a class `clone()`, `copy.copy` and `copy.deepcopy` in those terms; `Serializable` pickles the same way.
"""

from __future__ import annotations

import copy
from collections import OrderedDict
from collections.abc import Mapping
from typing import Any, Optional

import numpy as np
import torch

from .readonlytensor import ReadOnlyTensor


def deep_clone(x: Any, *, otherwise_deepcopy: bool = False, otherwise_return: bool = False, otherwise_fail: bool = False,
               memo: Optional[dict] = None) -> Any:
    """Clone `x` recursively.  Exactly one of the `otherwise_*` flags says what to do with objects of unknown kinds:
    deep-copy them, return them as they are, or raise."""
    if int(bool(otherwise_deepcopy)) + int(bool(otherwise_return)) + int(bool(otherwise_fail)) != 1:
        raise ValueError("Expected exactly one of these arguments as True: `otherwise_deepcopy`, `otherwise_return`, `otherwise_fail`")
    memo = {} if memo is None else memo
    key = id(x)
    if key in memo:
        return memo[key]

    def again(item: Any) -> Any:
        return deep_clone(item, otherwise_deepcopy=otherwise_deepcopy, otherwise_return=otherwise_return, otherwise_fail=otherwise_fail, memo=memo)

    if x is None or isinstance(x, (bool, int, float, complex, str, bytes, type, range, slice)) or x is Ellipsis or x is NotImplemented:
        result = x
    elif isinstance(x, ReadOnlyTensor):
        result = x.clone(preserve_read_only=True)
    elif isinstance(x, torch.Tensor):
        result = x.detach().clone() if not x.requires_grad else x.clone()
    elif isinstance(x, np.ndarray):
        result = x.copy()
        result.flags["WRITEABLE"] = x.flags["WRITEABLE"]
    elif isinstance(x, Clonable):
        result = x.clone(memo=memo)
    elif isinstance(x, (dict, OrderedDict)):
        result = type(x)()
        memo[key] = result  # registered before its items are visited: cycles through this container resolve to the clone
        for k, v in x.items():
            result[again(k)] = again(v)
    elif isinstance(x, list):
        result = type(x)()
        memo[key] = result
        result.extend(again(item) for item in x)
    elif isinstance(x, set):
        result = type(x)()
        memo[key] = result
        result.update(again(item) for item in x)
    elif isinstance(x, frozenset):
        result = type(x)(again(item) for item in x)
    elif isinstance(x, tuple):
        items = [again(item) for item in x]
        result = type(x)(*items) if hasattr(x, "_fields") else type(x)(items)
    elif otherwise_deepcopy:
        result = copy.deepcopy(x, memo)
    elif otherwise_return:
        result = x
    else:
        raise TypeError(f"Do not know how to clone {x!r} (of type {type(x)}).")
    memo[key] = result
    return result


class Clonable:
    """Mixin: `clone()`, `copy.copy` and `copy.deepcopy` all mean "an independent object whose attributes are deep clones"."""

    def _get_cloned_state(self, *, memo: dict) -> dict:
        state = self.__getstate__() if type(self).__getstate__ is not object.__getstate__ and not isinstance(self, Serializable) else self.__dict__
        if not isinstance(state, Mapping):
            state = self.__dict__
        return {k: deep_clone(v, otherwise_deepcopy=True, memo=memo) for k, v in state.items()}

    def clone(self, *, memo: Optional[dict] = None) -> "Clonable":
        memo = {} if memo is None else memo
        if id(self) in memo:
            return memo[id(self)]
        new = object.__new__(type(self))
        memo[id(self)] = new
        new.__dict__.update(self._get_cloned_state(memo=memo))
        return new

    def __copy__(self) -> "Clonable":
        return self.clone()

    def __deepcopy__(self, memo: Optional[dict]) -> "Clonable":
        return self.clone(memo={} if memo is None else memo)


class Serializable(Clonable):
    """A Clonable that pickles as its cloned state (tensors detached from oversized storages)."""

    def __getstate__(self) -> dict:
        return self._get_cloned_state(memo={id(self): self})
