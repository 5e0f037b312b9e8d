"""Functional ClipUp (reference: algorithms/functional/funcclipup.py:23-151; Toklu et al., PPSN 2020).

    velocity <- momentum * velocity + lr * g / ||g||;   if ||velocity|| > max_speed: velocity <- max_speed * velocity / ||velocity||
    center   <- center + velocity
On CUDA float32 all batch items are ONE launch of the K5 `evok_clipup_batched` kernel (a CTA per item) writing the NEW state tensors.
"""

from __future__ import annotations

from typing import NamedTuple, Optional

import torch

from ... import ops
from .misc import batch_shape_of, expand_to, flat_items, host_scalar, on_kernels, scalar_items


class ClipUpState(NamedTuple):
    center: torch.Tensor
    velocity: torch.Tensor
    center_learning_rate: torch.Tensor
    momentum: torch.Tensor
    max_speed: torch.Tensor


def clipup(*, center_init, momentum=0.9, center_learning_rate=None, max_speed=None) -> ClipUpState:
    """Initial state.  One of `center_learning_rate` / `max_speed` may be omitted: max_speed = 2 * lr (funcclipup.py:31-92)."""
    center_init = torch.as_tensor(center_init)
    dtype = center_init.dtype
    if center_learning_rate is None and max_speed is None:
        raise ValueError("Both `center_learning_rate` and `max_speed` is missing. At least one of them is needed.")
    if center_learning_rate is not None:
        center_learning_rate = host_scalar(center_learning_rate, dtype)
    if max_speed is not None:
        max_speed = host_scalar(max_speed, dtype)
    if max_speed is None:
        max_speed = center_learning_rate * 2.0
    if center_learning_rate is None:
        center_learning_rate = max_speed / 2.0
    return ClipUpState(center=center_init, velocity=torch.zeros_like(center_init), center_learning_rate=center_learning_rate,
                       momentum=host_scalar(momentum, dtype), max_speed=max_speed)


def clipup_ask(state: ClipUpState) -> torch.Tensor:
    """The current search point."""
    return state.center


def clipup_tell(state: ClipUpState, *, follow_grad) -> ClipUpState:
    """The state after following `follow_grad` (batchable)."""
    center = state.center
    g = torch.as_tensor(follow_grad, dtype=center.dtype, device=center.device)
    lr, mom, cap = state.center_learning_rate, state.momentum, state.max_speed
    batch = batch_shape_of((center, 1), (state.velocity, 1), (g, 1), (lr, 0), (mom, 0), (cap, 0))
    if on_kernels(center, g):
        new_center = expand_to(center, batch, 1).contiguous().clone()
        new_velocity = expand_to(state.velocity, batch, 1).contiguous().clone()
        gs = flat_items(g, batch, 1).contiguous()
        cs, vs = new_center.view(-1, center.shape[-1]), new_velocity.view(-1, center.shape[-1])
        # ONE launch for all batch items (one CTA per item; the per-item hyper-parameters travel in the launch parameters)
        ops.clipup_batched_(gs, vs, cs, scalar_items(lr, batch), scalar_items(mom, batch), scalar_items(cap, batch))
    else:
        dev = center.device
        lr_, mom_, cap_ = (x.to(dev)[..., None] for x in (lr, mom, cap))
        velocity = mom_ * state.velocity + lr_ * (g / torch.linalg.vector_norm(g, dim=-1, keepdim=True))
        speed = torch.linalg.vector_norm(velocity, dim=-1, keepdim=True)
        new_velocity = torch.where(speed > cap_, cap_ * (velocity / speed), velocity)
        new_center = center + new_velocity
    return ClipUpState(center=new_center, velocity=new_velocity, center_learning_rate=lr, momentum=mom, max_speed=cap)
