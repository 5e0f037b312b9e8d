"""Loggers that hook into `searcher.log_hook` (mirrors the host-side part of evotorch.logging, logging.py:67-523)."""

from __future__ import annotations

from typing import Iterable, Optional

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
