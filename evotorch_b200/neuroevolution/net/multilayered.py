"""`MultiLayered`: the container `str_to_net` builds for "A >> B >> C" (reference: net/multilayered.py:21-90).  It behaves like
`nn.Sequential` for feed-forward stacks; layers that return `(output, hidden_state)` tuples are supported too -- the hidden
states travel in a dict keyed by layer index -- so recurrent layers written in that convention compose."""

from __future__ import annotations

from typing import Optional

import torch
from torch import nn


class MultiLayered(nn.Module):
    def __init__(self, *layers: nn.Module):
        super().__init__()
        self._submodules = nn.ModuleList(layers)

    def forward(self, x: torch.Tensor, h: Optional[dict] = None):
        states_in = {} if h is None else h
        states_out = {}
        for index, layer in enumerate(self._submodules):
            state = states_in.get(index)
            produced = layer(x) if state is None else layer(x, state)
            if isinstance(produced, torch.Tensor):
                x = produced
            elif isinstance(produced, tuple):
                if len(produced) != 2:
                    raise ValueError(f"The layer number {index} returned a tuple of length {len(produced)}."
                                     f" A tensor or a tuple of two elements was expected.")
                x, new_state = produced
                if new_state is not None:
                    states_out[index] = new_state
            else:
                raise TypeError(f"The layer number {index} returned an object of type {type(produced)}."
                                f" A tensor or a tuple of two elements was expected.")
        return x if not states_out else (x, states_out)

    def __iter__(self):
        return iter(self._submodules)

    def __getitem__(self, i):
        return self._submodules[i]

    def __len__(self) -> int:
        return len(self._submodules)

    def append(self, module: nn.Module):
        self._submodules.append(module)
