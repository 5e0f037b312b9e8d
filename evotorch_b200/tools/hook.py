"""`Hook`: an ordered collection of callables fired together (reference: tools/hook.py:25-190).

Calling the hook calls every registered function with the hook's stored positional / keyword arguments followed by the
call's own.  Functions may return None, a mapping or a sequence; mappings are merged into one dict, sequences are
concatenated into one list, mixing the two kinds is an error.  Searchers use hooks for their before/after-step callbacks
and status extensions, problems for before/after-eval and before/after-grad callbacks.
"""

from __future__ import annotations

from collections.abc import Iterable, Mapping, MutableSequence
from typing import Any, Callable, Optional, Union


class Hook(MutableSequence):
    def __init__(self, callables: Optional[Iterable[Callable]] = None, *, args: Optional[Iterable] = None, kwargs: Optional[Mapping] = None):
        self._funcs = list(callables) if callables is not None else []
        self._args = list(args) if args is not None else []
        self._kwargs = dict(kwargs) if kwargs is not None else {}

    # ------------------------------------------------------------------ firing
    def __call__(self, *args: Any, **kwargs: Any) -> Optional[Union[dict, list]]:
        call_args = [*self._args, *args]
        call_kwargs = {**self._kwargs, **kwargs}
        merged: Optional[Union[dict, list]] = None
        for func in self._funcs:
            returned = func(*call_args, **call_kwargs)
            if returned is None:
                continue
            if isinstance(returned, Mapping):
                kind, piece = dict, dict(returned)
            elif isinstance(returned, Iterable):
                kind, piece = list, list(returned)
            else:
                raise TypeError(f"Expected the function {func} to return None, or a dict-like object, or a list-like object."
                                f" However, the function returned an object of type {type(returned)!r}.")
            if merged is None:
                merged = piece
            elif not isinstance(merged, kind):
                got, had = ("dict-like", "list-like") if kind is dict else ("list-like", "dict-like")
                raise TypeError(f"The function {func} returned a {got} object. However, previous function(s) in this hook had returned"
                                f" {had} object(s). Such incompatible results cannot be accumulated.")
            elif kind is dict:
                merged.update(piece)
            else:
                merged.extend(piece)
        return merged

    def accumulate_dict(self, *args: Any, **kwargs: Any) -> dict:
        """Fire the hook; the functions are expected to return mappings (or None).  Always returns a dict."""
        result = self(*args, **kwargs)
        if result is None:
            return {}
        if isinstance(result, Mapping):
            return result
        raise TypeError(f"Expected the functions in this hook to accumulate dictionary-like objects. Instead, accumulated an object"
                        f" of type {type(result)}. Hint: are the functions registered in this hook returning non-dictionary iterables?")

    def accumulate_sequence(self, *args: Any, **kwargs: Any) -> list:
        """Fire the hook; the functions are expected to return sequences (or None).  Always returns a list."""
        result = self(*args, **kwargs)
        if result is None:
            return []
        if isinstance(result, Mapping):
            raise TypeError(f"Expected the functions in this hook to accumulate sequences (that are NOT dictionaries). Instead,"
                            f" accumulated a dict-like object of type {type(result)}.")
        return result

    # ------------------------------------------------------------------ stored arguments
    @property
    def args(self) -> list:
        return self._args

    @property
    def kwargs(self) -> dict:
        return self._kwargs

    # ------------------------------------------------------------------ MutableSequence
    def __getitem__(self, i):
        if isinstance(i, slice):
            return Hook(self._funcs[i], args=self._args, kwargs=self._kwargs)
        return self._funcs[i]

    def __setitem__(self, i, x):
        self._funcs[i] = x

    def __delitem__(self, i):
        del self._funcs[i]

    def insert(self, i: int, x: Callable):
        self._funcs.insert(i, x)

    def __len__(self) -> int:
        return len(self._funcs)

    def __repr__(self) -> str:
        parts = [repr(self._funcs)]
        if self._args:
            parts.append(f"args={self._args}")
        if self._kwargs:
            parts.append(f"kwargs={self._kwargs}")
        return f"{type(self).__name__}({', '.join(parts)})"

    __str__ = __repr__
