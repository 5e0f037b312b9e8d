"""Center optimizers of the distribution-based searchers (mirrors evotorch.optimizers, optimizers.py:31-457).

Every optimizer exposes `ascent(g) -> step` (what the reference's `_follow_gradient` calls, distributions.py:387).
On CUDA float32 state the step is ONE kernel (csrc/evok_update.cu) that reduces the norms on the device, so there
is no host synchronisation (the reference's ClipUp does `if normx > limit` on the host, optimizers.py:313).
Optimizer state (velocity / moments) is held in ordinary torch tensors so pickling keeps working.
"""

from __future__ import annotations

from collections.abc import Mapping
from typing import Callable, Optional

import torch

from . import ops
from .tools.misc import ensure_tensor_length_and_dtype, to_torch_dtype


class _VectorOptimizer:
    def _prep(self, g, n: int, dtype, device) -> torch.Tensor:
        return ensure_tensor_length_and_dtype(g, n, dtype, about=f"{type(self).__name__}.ascent", device=device).contiguous()


class ClipUp(_VectorOptimizer):
    """ClipUp (Toklu et al. 2020): v <- clip_norm(momentum*v + stepsize*g/||g||, max_speed) (optimizers.py:231-357)."""

    _param_group_items = {"lr": "_stepsize", "max_speed": "_max_speed", "momentum": "_momentum"}
    _param_group_item_lb = {"lr": 0.0, "max_speed": 0.0, "momentum": 0.0}
    _param_group_item_ub = {"momentum": 1.0}

    def __init__(self, *, solution_length: int, dtype, stepsize: float, momentum: float = 0.9, max_speed: Optional[float] = None,
                 device="cpu"):
        stepsize, momentum = float(stepsize), float(momentum)
        max_speed = stepsize * 2.0 if max_speed is None else float(max_speed)  # optimizers.py:274-275
        if stepsize < 0.0:
            raise ValueError(f"Invalid stepsize: {stepsize}")
        if momentum < 0.0 or momentum > 1.0:
            raise ValueError(f"Invalid momentum: {momentum}")
        if max_speed < 0.0:
            raise ValueError(f"Invalid max_speed: {max_speed}")
        self._stepsize, self._momentum, self._max_speed = stepsize, momentum, max_speed
        self._dtype = to_torch_dtype(dtype)
        self._device = torch.device(device)
        self._velocity = torch.zeros(int(solution_length), dtype=self._dtype, device=self._device)
        self._param_groups = (ClipUpParameterGroup(self),)

    @torch.no_grad()
    def ascent(self, globalg, *, cloned_result: bool = True) -> torch.Tensor:
        g = self._prep(globalg, len(self._velocity), self._dtype, self._device)
        if ops.uses_kernels(g):
            step = torch.empty_like(g)
            ops.clipup_step(g, self._velocity, self._stepsize, self._momentum, self._max_speed, step_out=step)
            return step
        v = self._momentum * self._velocity + (g / torch.norm(g)) * self._stepsize
        vnorm = torch.norm(v)
        if vnorm > self._max_speed:
            v = v * (self._max_speed / vnorm)
        self._velocity = v
        return v.clone() if cloned_result else v

    @torch.no_grad()
    def ascent_into_(self, globalg: torch.Tensor, mu: torch.Tensor) -> None:
        """Fused `mu += ascent(g)` (one kernel on CUDA)."""
        if ops.uses_kernels(globalg) and ops.uses_kernels(mu):
            ops.clipup_step(globalg.contiguous(), self._velocity, self._stepsize, self._momentum, self._max_speed, mu=mu)
        else:
            mu += self.ascent(globalg, cloned_result=False)

    @property
    def contained_optimizer(self) -> "ClipUp":
        return self

    @property
    def param_groups(self) -> tuple:
        return self._param_groups


class ClipUpParameterGroup(Mapping):
    """Dictionary-like view of ClipUp's hyper-parameters (`optimizer.param_groups[0]["lr"] = ...`, optimizers.py:367-418)."""

    def __init__(self, clipup: ClipUp):
        self.clipup = clipup

    def __getitem__(self, key: str) -> float:
        return getattr(self.clipup, ClipUp._param_group_items[key])

    def __setitem__(self, key: str, value: float):
        attr = ClipUp._param_group_items[key]
        value = float(value)
        if key in ClipUp._param_group_item_lb and value < ClipUp._param_group_item_lb[key]:
            raise ValueError(f"Invalid value for {key!r}: {value}")
        if key in ClipUp._param_group_item_ub and value > ClipUp._param_group_item_ub[key]:
            raise ValueError(f"Invalid value for {key!r}: {value}")
        setattr(self.clipup, attr, value)

    def __iter__(self):
        return iter(ClipUp._param_group_items)

    def __len__(self) -> int:
        return len(ClipUp._param_group_items)

    def __repr__(self) -> str:
        return f"<{type(self).__name__}: {dict(self)}>"


class Adam(_VectorOptimizer):
    """Adam as an ascent-step producer (optimizers.py:101-165 wraps torch.optim.Adam on a dummy parameter; the
    resulting step is lr * m_hat / (sqrt(v_hat) + eps))."""

    def __init__(self, *, solution_length: int, dtype, device="cpu", stepsize: Optional[float] = None, beta1: Optional[float] = None,
                 beta2: Optional[float] = None, epsilon: Optional[float] = None, amsgrad: Optional[bool] = None):
        if (beta1 is None) != (beta2 is None):
            raise ValueError("The arguments beta1 and beta2 were expected as both None, or as both real numbers.")
        if amsgrad:
            raise NotImplementedError("amsgrad is not supported by the kernel-backed Adam")
        self._lr = 1e-3 if stepsize is None else float(stepsize)
        self._b1 = 0.9 if beta1 is None else float(beta1)
        self._b2 = 0.999 if beta2 is None else float(beta2)
        self._eps = 1e-8 if epsilon is None else float(epsilon)
        self._dtype, self._device = to_torch_dtype(dtype), torch.device(device)
        self._m = torch.zeros(int(solution_length), dtype=self._dtype, device=self._device)
        self._v = torch.zeros_like(self._m)
        self._t = 0
        self.param_groups = [{"lr": self._lr, "betas": (self._b1, self._b2), "eps": self._eps}]

    def _hyper(self):
        pg = self.param_groups[0]
        return float(pg["lr"]), float(pg["betas"][0]), float(pg["betas"][1]), float(pg["eps"])

    @torch.no_grad()
    def ascent(self, globalg, *, cloned_result: bool = True) -> torch.Tensor:
        g = self._prep(globalg, len(self._m), self._dtype, self._device)
        lr, b1, b2, eps = self._hyper()
        self._t += 1
        if ops.uses_kernels(g):
            step = torch.empty_like(g)
            ops.adam_step(g, self._m, self._v, self._t, lr, b1, b2, eps, step_out=step)
            return step
        self._m = b1 * self._m + (1 - b1) * g
        self._v = b2 * self._v + (1 - b2) * g * g
        bc1, bc2 = 1 - b1**self._t, 1 - b2**self._t
        denom = self._v.sqrt() / (bc2**0.5) + eps
        return (lr / bc1) * (self._m / denom)

    @torch.no_grad()
    def ascent_into_(self, globalg: torch.Tensor, mu: torch.Tensor) -> None:
        if ops.uses_kernels(globalg) and ops.uses_kernels(mu):
            lr, b1, b2, eps = self._hyper()
            self._t += 1
            ops.adam_step(globalg.contiguous(), self._m, self._v, self._t, lr, b1, b2, eps, mu=mu)
        else:
            mu += self.ascent(globalg)

    @property
    def contained_optimizer(self) -> "Adam":
        return self


class SGD(_VectorOptimizer):
    """SGD with optional momentum as an ascent-step producer (optimizers.py:168-228)."""

    def __init__(self, *, solution_length: int, dtype, stepsize: float, device="cpu", momentum: Optional[float] = None,
                 dampening: Optional[bool] = None, nesterov: Optional[bool] = None):
        if dampening or nesterov:
            raise NotImplementedError("dampening / nesterov are not supported by the kernel-backed SGD")
        self._lr = float(stepsize)
        self._momentum = 0.0 if momentum is None else float(momentum)
        self._dtype, self._device = to_torch_dtype(dtype), torch.device(device)
        self._n = int(solution_length)
        self._buf = torch.zeros(self._n, dtype=self._dtype, device=self._device) if self._momentum != 0.0 else None
        self._first = True
        self.param_groups = [{"lr": self._lr, "momentum": self._momentum}]

    @torch.no_grad()
    def ascent(self, globalg, *, cloned_result: bool = True) -> torch.Tensor:
        g = self._prep(globalg, self._n, self._dtype, self._device)
        lr, mom = float(self.param_groups[0]["lr"]), float(self.param_groups[0]["momentum"])
        first, self._first = self._first, False
        if ops.uses_kernels(g):
            step = torch.empty_like(g)
            ops.sgd_step(g, self._buf, first, lr, mom, step_out=step)
            return step
        if mom != 0.0:
            self._buf = g.clone() if first else mom * self._buf + g
            d = self._buf
        else:
            d = g
        return lr * d

    @torch.no_grad()
    def ascent_into_(self, globalg: torch.Tensor, mu: torch.Tensor) -> None:
        mu += self.ascent(globalg)

    @property
    def contained_optimizer(self) -> "SGD":
        return self


def get_optimizer_class(s: str, optimizer_config: Optional[dict] = None) -> Callable:
    """Name -> optimizer class, optionally pre-configured (optimizers.py:421-457)."""
    if s in ("clipsgd", "clipsga", "clipup"):
        cls = ClipUp
    elif s == "adam":
        cls = Adam
    elif s in ("sgd", "sga"):
        cls = SGD
    else:
        raise ValueError(f"Unknown optimizer: {s!r}")
    if optimizer_config is None:
        return cls

    def configured(*args, **kwargs):
        conf = dict(optimizer_config)
        conf.update(kwargs)
        return cls(*args, **conf)

    return configured
