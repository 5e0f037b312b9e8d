"""`ReadOnlyTensor`: what `SolutionBatch.values` / `.evals` and `Solution.values` / `.evals` hand out (reference:
tools/readonlytensor.py:25-226).  It IS a torch tensor sharing storage with the population the kernels wrote -- every torch
function accepts it -- but in-place modification is refused: methods whose names end in "_" are hidden, augmented assignment
and item assignment raise, and it cannot be the `out=` target of a torch function.  `clone()` gives an ordinary tensor;
indexing / reshaping give a read-only view when they share storage and an ordinary tensor when they copy.
"""

from __future__ import annotations

from typing import Any, Optional

import numpy as np
import torch


def storage_ptr(x: torch.Tensor) -> int:
    """Address of the underlying storage (tools/misc.py `storage_ptr`): equal for tensors that share memory."""
    return x.untyped_storage().data_ptr()


def _refuse(self, *args, **kwargs):
    raise TypeError("The contents of a ReadOnlyTensor cannot be modified")


class ReadOnlyTensor(torch.Tensor):
    def __getattribute__(self, name: str) -> Any:
        if isinstance(name, str) and name.endswith("_") and not (name.startswith("__") and name.endswith("__")):
            raise AttributeError(f"A ReadOnlyTensor explicitly disables all members whose names end with '_'. Cannot access member {name!r}.")
        return super().__getattribute__(name)

    __setitem__ = __iadd__ = __isub__ = __imul__ = __itruediv__ = __ifloordiv__ = __imod__ = __ipow__ = __imatmul__ = _refuse
    __iand__ = __ior__ = __ixor__ = __ilshift__ = __irshift__ = __idiv__ = _refuse

    def _view_or_copy(self, other: torch.Tensor) -> torch.Tensor:
        """Results that do not share this tensor's storage are ordinary (mutable) tensors."""
        if isinstance(other, torch.Tensor) and storage_ptr(other) != storage_ptr(self):
            return other.as_subclass(torch.Tensor)
        return other

    def clone(self, *, preserve_read_only: bool = False, **kwargs) -> torch.Tensor:
        result = super().clone(**kwargs)
        return result if preserve_read_only else result.as_subclass(torch.Tensor)

    def __getitem__(self, index) -> torch.Tensor:
        return self._view_or_copy(super().__getitem__(index))

    def reshape(self, *args, **kwargs) -> torch.Tensor:
        return self._view_or_copy(super().reshape(*args, **kwargs))

    def numpy(self, *args, **kwargs) -> np.ndarray:
        array = torch.Tensor.numpy(self, *args, **kwargs)
        array.flags["WRITEABLE"] = False
        return array

    def __array__(self, *args, **kwargs) -> np.ndarray:
        array = super().__array__(*args, **kwargs)
        array.flags["WRITEABLE"] = False
        return array

    def __copy__(self):
        return self.clone(preserve_read_only=True)

    def __deepcopy__(self, memo):
        return self.clone(preserve_read_only=True)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        if kwargs is not None and isinstance(kwargs.get("out"), ReadOnlyTensor):
            raise TypeError(f"The `out` keyword argument passed to {func} is a ReadOnlyTensor: it cannot be the target of an in-place result.")
        return super().__torch_function__(func, types, args, kwargs)


def as_plain_tensor(x: Any) -> Any:
    """`x` itself, or -- for a ReadOnlyTensor -- an ordinary tensor view of the same memory.  Library code that only READS its
    input calls this first, so that the tensors it allocates "like" the input (and returns) are ordinary, writable tensors."""
    return x.as_subclass(torch.Tensor) if isinstance(x, ReadOnlyTensor) else x


def read_only_tensor(x: Any, *, dtype: Optional[torch.dtype] = None, device=None) -> ReadOnlyTensor:
    """A NEW read-only tensor holding a copy of `x`."""
    kw = {k: v for k, v in (("dtype", dtype), ("device", device)) if v is not None}
    return torch.tensor(x, **kw).as_subclass(ReadOnlyTensor) if not isinstance(x, torch.Tensor) else x.detach().clone().to(**kw).as_subclass(ReadOnlyTensor)


def as_read_only_tensor(x: Any, *, dtype: Optional[torch.dtype] = None, device=None) -> ReadOnlyTensor:
    """The read-only view of `x`: shares memory with it whenever `torch.as_tensor` can avoid a copy."""
    kw = {k: v for k, v in (("dtype", dtype), ("device", device)) if v is not None}
    return torch.as_tensor(x, **kw).as_subclass(ReadOnlyTensor)
