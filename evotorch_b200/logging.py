"""Loggers that hook into `searcher.log_hook` (mirrors the host-side part of evotorch.logging, logging.py:67-523)."""

from __future__ import annotations

import os
import pickle
import weakref
from datetime import datetime
from typing import Any, Iterable, Optional, Union

import torch


class Logger:
    """Calls `_log(status)` every `interval` generations and once more at the end of a run (logging.py:67-108)."""

    def __init__(self, searcher, *, interval: int = 1, after_first_step: bool = False):
        searcher.log_hook.append(self)
        searcher.end_of_run_hook.append(self._final)
        self._interval = int(interval)
        self._after_first_step = bool(after_first_step)
        self._steps = 0
        self._last_logged = None

    def __call__(self, status: dict):
        fire = (self._steps % self._interval == 0) if self._after_first_step else ((self._steps + 1) % self._interval == 0)
        self._steps += 1
        if fire:
            self._log(self._filter(status))
            self._last_logged = self._steps

    def _final(self, status: dict):
        if self._last_logged != self._steps:
            self._log(self._filter(status))
            self._last_logged = self._steps

    def _filter(self, status: dict) -> dict:
        return status

    def _log(self, status: dict):
        raise NotImplementedError


class ScalarLogger(Logger):
    """Keeps only scalar status entries (logging.py:394-425)."""

    def _filter(self, status: dict) -> dict:
        out = {}
        for k, v in status.items():
            if isinstance(v, (int, float, bool)):
                out[k] = v
            elif isinstance(v, torch.Tensor) and v.numel() == 1:
                out[k] = v.item()
        return out


class StdOutLogger(ScalarLogger):
    """Prints the scalar status entries each generation (logging.py:428-476)."""

    def __init__(self, searcher, *, interval: int = 1, after_first_step: bool = False, leading_keys: Iterable[str] = ("iter",)):
        super().__init__(searcher, interval=interval, after_first_step=after_first_step)
        self._leading_keys = list(leading_keys)

    def _log(self, status: dict):
        width = max((len(str(k)) for k in status), default=0)
        keys = [k for k in self._leading_keys if k in status] + [k for k in status if k not in self._leading_keys]
        for k in keys:
            print(str(k).rjust(width), ":", status[k])
        print()


class PandasLogger(ScalarLogger):
    """Collects the scalar status entries; `to_dataframe()` returns them as a pandas DataFrame (logging.py:479-523)."""

    def __init__(self, searcher, *, interval: int = 1, after_first_step: bool = False):
        super().__init__(searcher, interval=interval, after_first_step=after_first_step)
        self._data = []

    def _log(self, status: dict):
        self._data.append(dict(status))

    def to_dataframe(self, *, index: Optional[str] = "iter"):
        import pandas

        df = pandas.DataFrame(self._data)
        if index is not None and index in df.columns:
            df.set_index(index, inplace=True)
        return df


class PicklingLogger(Logger):
    """Every `interval` generations (and at the end of a run) pickle the chosen status items -- as host tensors, so the file loads
    on a machine without a GPU -- into `<directory>/<prefix>_generation<NNNNNN>.pickle` (logging.py:111-392).

    `items_to_save` are status keys ("center", "best", ...); `Solution`s are stored as their decision values.  The file
    also carries "beginning_time" / "now" / "elapsed".  With `checkpoint=True` the whole searcher (distribution, optimizer
    state, Philox key and generation counter) is stored under the key "searcher": `resume(file)` returns a searcher that
    continues the very same trajectory (the sampler is counter based, so the resumed run is bit-identical to an
    uninterrupted one)."""

    def __init__(self, searcher, *, interval: int, directory: Optional[str] = None, prefix: Optional[str] = None, zfill: int = 6,
                 items_to_save: Union[str, Iterable[str]] = ("center", "best"), make_policy_from: Optional[str] = None,
                 after_first_step: bool = False, verbose: bool = True, checkpoint: bool = False):
        super().__init__(searcher, interval=interval, after_first_step=after_first_step)
        self._searcher_ref = weakref.ref(searcher)
        self._items_to_save = (items_to_save,) if isinstance(items_to_save, str) else tuple(items_to_save)
        if prefix is None:
            prefix = f"{type(searcher.problem).__name__}_{datetime.now().strftime('%Y-%m-%d-%H.%M.%S')}_{os.getpid()}"
        self._prefix = str(prefix)
        self._directory = None if directory is None else str(directory)
        if self._directory is not None:
            os.makedirs(self._directory, exist_ok=True)
        self._verbose, self._zfill, self._checkpoint = bool(verbose), int(zfill), bool(checkpoint)
        self._make_policy_from = None if make_policy_from is None else str(make_policy_from)
        self._last_generation: Optional[int] = None
        self._last_file_name: Optional[str] = None

    def __getstate__(self) -> dict:
        state = dict(self.__dict__)
        state["_searcher_ref"] = None  # weak references do not pickle; `resume` re-binds
        return state

    @staticmethod
    def _as_cpu(x: Any) -> Any:
        from .core import Solution

        if isinstance(x, Solution):
            x = x.values
        if isinstance(x, torch.Tensor):
            x = x.detach().to("cpu").clone()
        return x

    def _log(self, status: dict):
        self.save()

    def _final(self, status: dict):
        searcher = self._searcher_ref() if self._searcher_ref is not None else None
        if searcher is not None and (self._last_generation is None or searcher.step_count > self._last_generation):
            self.save()

    def save(self, fname: Optional[str] = None) -> Optional[str]:
        """Write the pickle now; returns the file name (None if the searcher is gone)."""
        searcher = self._searcher_ref() if self._searcher_ref is not None else None
        if searcher is None:
            return None
        status = searcher.status
        data = {k: self._as_cpu(status[k]) for k in self._items_to_save if k in status}
        # neuro-evolution problems: the observation statistics and a ready-to-use policy go into the file too (logging.py:297-351)
        problem = searcher.problem
        if hasattr(problem, "observation_normalization") and hasattr(problem, "get_observation_stats") and problem.observation_normalization:
            stats = problem.get_observation_stats()
            data["obs_stats"] = stats.to("cpu") if hasattr(stats, "to") else stats
        if hasattr(problem, "to_policy"):
            if self.__dict__.get("_make_policy_from") is None:
                if "center" in status:
                    policy_key = "center"
                elif "pop_best" in status:
                    policy_key = "pop_best"
                else:
                    raise ValueError("PicklingLogger did not receive an explicit value for its `make_policy_from` argument."
                                     " The status dictionary of the search algorithm has neither 'center' nor 'pop_best'."
                                     " Therefore, it is not clear which status item is to be used for making a policy."
                                     " Please try instantiating a PicklingLogger with an explicit `make_policy_from` value.")
            else:
                policy_key = self._make_policy_from
            data["policy"] = problem.to_policy(status[policy_key]).to("cpu")
        begun = searcher.first_step_datetime
        if begun is not None:
            now = datetime.now()
            data.update(beginning_time=begun, now=now, elapsed=now - begun)
        if self._checkpoint:
            data["searcher"] = searcher
        if fname is None:
            fname = f"{self._prefix}_generation{str(searcher.step_count).zfill(self._zfill)}.pickle"
        if self._directory is not None:
            fname = os.path.join(self._directory, str(fname))
        with open(fname, "wb") as f:
            pickle.dump(data, f)
        self._last_generation, self._last_file_name = searcher.step_count, str(fname)
        if self._verbose:
            print("Saved to", fname)
        return str(fname)

    @property
    def last_generation(self) -> Optional[int]:
        return self._last_generation

    @property
    def last_file_name(self) -> Optional[str]:
        return self._last_file_name

    def unpickle_last_file(self) -> dict:
        with open(self._last_file_name, "rb") as f:
            return pickle.load(f)

    @staticmethod
    def resume(fname: str):
        """The searcher stored by a `checkpoint=True` logger, ready to `step()` / `run()` on."""
        with open(fname, "rb") as f:
            data = pickle.load(f)
        if "searcher" not in data:
            raise KeyError(f"{fname} holds no searcher: create the PicklingLogger with checkpoint=True")
        searcher = data["searcher"]
        for hook in list(searcher.log_hook):
            if isinstance(hook, PicklingLogger):
                hook._searcher_ref = weakref.ref(searcher)
        return searcher
