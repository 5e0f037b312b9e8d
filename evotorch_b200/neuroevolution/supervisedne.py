"""`SupervisedNE`: minimise a loss over minibatches of a dataset by neuro-evolution (mirrors
evotorch.neuroevolution.supervisedne.SupervisedNE, supervisedne.py:31-348: same constructor arguments, `make_dataloader`,
`get_minibatch`, `loss`, `_evaluate_using_minibatch`, `_evaluate_network`, `_evaluate_batch`).

With `common_minibatch=True` (the default) every solution of a population is scored on the SAME minibatches
(supervisedne.py:337-347).  The reference still runs the networks one after the other; here the whole population goes through
`NEProblem.batched_forward` -- for a feed-forward net the first layer of all N networks is one dense product of the stacked
weight rows with the shared minibatch (tensor cores on CUDA float32) -- and the loss is applied per solution with `vmap`.
With `common_minibatch=False` each solution draws its own minibatches, exactly like the reference's loop.
"""

from __future__ import annotations

from typing import Any, Callable, Optional, Union

import torch
from torch import nn
from torch.utils.data import DataLoader, Dataset

from ..core import SolutionBatch
from .neproblem import NEProblem


class SupervisedNE(NEProblem):
    def __init__(self, dataset: Dataset, network: Union[str, nn.Module, Callable[[], nn.Module]], loss_func: Optional[Callable] = None, *,
                 network_args: Optional[dict] = None, initial_bounds=(-0.00001, 0.00001), minibatch_size: Optional[int] = None,
                 num_minibatches: Optional[int] = None, num_actors=None, common_minibatch: bool = True, num_gpus_per_actor=None,
                 actor_config: Optional[dict] = None, num_subbatches: Optional[int] = None, subbatch_size: Optional[int] = None, device=None):
        super().__init__(objective_sense="min", network=network, network_args=network_args, initial_bounds=initial_bounds,
                         num_actors=num_actors, num_gpus_per_actor=num_gpus_per_actor, actor_config=actor_config,
                         num_subbatches=num_subbatches, subbatch_size=subbatch_size, device=device)
        self.dataset = dataset
        self.dataloader: Optional[DataLoader] = None
        self.dataloader_iterator = None
        self._loss_func = loss_func
        self._minibatch_size = None if minibatch_size is None else int(minibatch_size)
        self._num_minibatches = 1 if num_minibatches is None else int(num_minibatches)
        self._common_minibatch = bool(common_minibatch)
        self._current_minibatches: Optional[list] = None
        self._batched_ok = True  # cleared if the loss or the network turns out not to be batchable over solutions

    # ------------------------------------------------------------------ data (supervisedne.py:226-300)
    def _make_dataloader(self) -> DataLoader:
        """Override point when no `minibatch_size` was given."""
        raise NotImplementedError

    def make_dataloader(self) -> DataLoader:
        if self._minibatch_size is None:
            return self._make_dataloader()
        return DataLoader(self.dataset, shuffle=True, batch_size=self._minibatch_size)

    def _prepare(self) -> None:
        self.dataloader = self.make_dataloader()

    def get_minibatch(self) -> Any:
        """The next minibatch of the DataLoader (restarting it when exhausted), moved to the network's device."""
        if self.dataloader is None:
            self._prepare()
        if self.dataloader_iterator is None:
            self.dataloader_iterator = iter(self.dataloader)
        batch = None
        try:
            batch = next(self.dataloader_iterator)
        except StopIteration:
            pass
        if batch is None:
            self.dataloader_iterator = iter(self.dataloader)
            batch = next(self.dataloader_iterator)
        return [var.to(self.network_device) for var in batch]

    # ------------------------------------------------------------------ loss (supervisedne.py:262-297)
    def _loss(self, y_hat: Any, y: Any) -> Union[float, torch.Tensor]:
        """Override point when no `loss_func` was given."""
        raise NotImplementedError

    def loss(self, y_hat: Any, y: Any) -> Union[float, torch.Tensor]:
        return self._loss(y_hat, y) if self._loss_func is None else self._loss_func(y_hat, y)

    def _evaluate_using_minibatch(self, network: nn.Module, batch: Any) -> Union[float, torch.Tensor]:
        with torch.no_grad():
            x, y = batch
            return self.loss(network(x), y)

    # ------------------------------------------------------------------ evaluation
    def _evaluate_network(self, network: nn.Module) -> torch.Tensor:
        """One network over `num_minibatches` minibatches, mean loss (supervisedne.py:327-335)."""
        loss = 0.0
        for batch_idx in range(self._num_minibatches):
            if not self._common_minibatch:
                self._current_minibatch = self.get_minibatch()
            else:
                self._current_minibatch = self._current_minibatches[batch_idx]
            loss += self._evaluate_using_minibatch(network, self._current_minibatch) / self._num_minibatches
        return loss

    def _evaluate_population(self, parameters: torch.Tensor) -> Optional[torch.Tensor]:
        """All solutions on the common minibatches at once: losses[i] = mean over the minibatches of loss(net_i(x), y)."""
        if not (self._common_minibatch and self._batched_ok and self._network_eval_func is None):
            return None
        if (type(self)._evaluate_network is not SupervisedNE._evaluate_network
                or type(self)._evaluate_using_minibatch is not SupervisedNE._evaluate_using_minibatch):
            return None  # a subclass customised the per-network evaluation: keep the reference's loop
        try:
            total = None
            for x, y in self._current_minibatches:
                y_hat = self.batched_forward(parameters, x)  # N x B x out
                losses = torch.vmap(lambda yh: torch.as_tensor(self.loss(yh, y)), in_dims=0)(y_hat)
                if losses.ndim != 1:
                    raise ValueError("the loss must return a scalar")
                total = losses if total is None else total + losses
            return total / self._num_minibatches
        except Exception:
            self._batched_ok = False  # e.g. a loss that vmap cannot trace, a network with buffers updated in forward: loop instead
            return None

    def _evaluate_batch(self, batch: SolutionBatch):
        if self._common_minibatch:
            self._current_minibatches = [self.get_minibatch() for _ in range(self._num_minibatches)]
        return super()._evaluate_batch(batch)
