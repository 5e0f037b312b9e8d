"""Functional gradient ascent with optional Polyak momentum (reference: algorithms/functional/funcsgd.py:23-130):

    velocity <- momentum * velocity + lr * g;   center <- center + velocity
"""

from __future__ import annotations

from typing import NamedTuple, Optional

import torch

from ... import ops
from .misc import batch_shape_of, expand_to, flat_items, host_scalar, on_kernels, scalar_items


class SGDState(NamedTuple):
    center: torch.Tensor
    velocity: torch.Tensor
    center_learning_rate: torch.Tensor
    momentum: torch.Tensor


def sgd(*, center_init, center_learning_rate, momentum=None) -> SGDState:
    center_init = torch.as_tensor(center_init)
    dtype = center_init.dtype
    return SGDState(center=center_init, velocity=torch.zeros_like(center_init), center_learning_rate=host_scalar(center_learning_rate, dtype),
                    momentum=host_scalar(0.0 if momentum is None else momentum, dtype))


def sgd_ask(state: SGDState) -> torch.Tensor:
    return state.center


def sgd_tell(state: SGDState, *, follow_grad) -> SGDState:
    center = state.center
    g = torch.as_tensor(follow_grad, dtype=center.dtype, device=center.device)
    lr, mom = state.center_learning_rate, state.momentum
    batch = batch_shape_of((center, 1), (state.velocity, 1), (g, 1), (lr, 0), (mom, 0))
    if on_kernels(center, g):
        d = center.shape[-1]
        new_center = expand_to(center, batch, 1).contiguous().clone()
        new_velocity = expand_to(state.velocity, batch, 1).contiguous().clone()
        gs = flat_items(g, batch, 1).contiguous()
        cs, vs = new_center.view(-1, d), new_velocity.view(-1, d)
        for b, (lr_b, mom_b) in enumerate(zip(scalar_items(lr, batch), scalar_items(mom, batch))):
            # velocity <- mom * velocity + lr * g, center += velocity: two axpy launches of the K5 family
            vs[b].mul_(mom_b)
            ops.axpy_(vs[b], gs[b], lr_b)
            ops.axpy_(cs[b], vs[b], 1.0)
    else:
        dev = center.device
        new_velocity = mom.to(dev)[..., None] * state.velocity + lr.to(dev)[..., None] * g
        new_center = center + new_velocity
    return SGDState(center=new_center, velocity=new_velocity, center_learning_rate=lr, momentum=mom)
