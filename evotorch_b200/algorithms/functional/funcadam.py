"""Functional Adam (reference: algorithms/functional/funcadam.py:23-172; Kingma & Ba 2015), ascent form:

    t += 1;  m <- b1 m + (1 - b1) g;  v <- b2 v + (1 - b2) g^2;  center <- center + lr * (m / (1 - b1^t)) / (sqrt(v / (1 - b2^t)) + eps)
The step counter `t` lives on the host (the K5 kernel takes it as a launch argument: no device round trip per tell).
"""

from __future__ import annotations

from typing import NamedTuple

import torch

from ... import ops
from .misc import batch_shape_of, expand_to, flat_items, host_scalar, on_kernels, scalar_items


class AdamState(NamedTuple):
    center: torch.Tensor
    center_learning_rate: torch.Tensor
    beta1: torch.Tensor
    beta2: torch.Tensor
    epsilon: torch.Tensor
    m: torch.Tensor
    v: torch.Tensor
    t: torch.Tensor


def adam(*, center_init, center_learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8) -> AdamState:
    center_init = torch.as_tensor(center_init)
    dtype = center_init.dtype
    return AdamState(center=center_init, center_learning_rate=host_scalar(center_learning_rate, dtype), beta1=host_scalar(beta1, dtype),
                     beta2=host_scalar(beta2, dtype), epsilon=host_scalar(epsilon, dtype), m=torch.zeros_like(center_init),
                     v=torch.zeros_like(center_init), t=torch.zeros(center_init.shape[:-1], dtype=dtype, device="cpu"))


def adam_ask(state: AdamState) -> torch.Tensor:
    return state.center


def adam_tell(state: AdamState, *, follow_grad) -> AdamState:
    center = state.center
    g = torch.as_tensor(follow_grad, dtype=center.dtype, device=center.device)
    lr, b1, b2, eps = state.center_learning_rate, state.beta1, state.beta2, state.epsilon
    batch = batch_shape_of((center, 1), (state.m, 1), (g, 1), (lr, 0), (b1, 0), (b2, 0), (eps, 0), (state.t, 0))
    t = expand_to(state.t, batch, 0) + 1
    if on_kernels(center, g):
        d = center.shape[-1]
        new_center = expand_to(center, batch, 1).contiguous().clone()
        new_m = expand_to(state.m, batch, 1).contiguous().clone()
        new_v = expand_to(state.v, batch, 1).contiguous().clone()
        gs = flat_items(g, batch, 1).contiguous()
        cs, ms, vs = new_center.view(-1, d), new_m.view(-1, d), new_v.view(-1, d)
        items = zip(scalar_items(t, batch), scalar_items(lr, batch), scalar_items(b1, batch), scalar_items(b2, batch), scalar_items(eps, batch))
        for b, (t_b, lr_b, b1_b, b2_b, eps_b) in enumerate(items):
            ops.adam_step(gs[b], ms[b], vs[b], int(t_b), lr_b, b1_b, b2_b, eps_b, mu=cs[b])
    else:
        dev = center.device
        lr_, b1_, b2_, eps_, t_ = (x.to(dev)[..., None] for x in (lr, b1, b2, eps, t))
        new_m = b1_ * state.m + (1 - b1_) * g
        new_v = b2_ * state.v + (1 - b2_) * (g**2.0)
        mhat = new_m / (1 - b1_**t_)
        vhat = new_v / (1 - b2_**t_)
        new_center = center + lr_ * mhat / (torch.sqrt(vhat) + eps_)
    return AdamState(center=new_center, center_learning_rate=lr, beta1=b1, beta2=b2, epsilon=eps, m=new_m, v=new_v, t=t)
