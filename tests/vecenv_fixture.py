"""A deterministic vectorised toy environment (test fixture; reference convention of vecgymne.py: auto-reset, tensors in,
tensors out).  State x in R^n_obs per sub-environment:  x <- 2 tanh(x A^T + a B^T + c) + 0.05 (c: a constant push the policy has to learn to cancel), reward = 1 - mean(x^2) + 0.1 a_0,
episode i ends after `lengths[i]` steps and restarts from its initial state.  `as_numpy()` gives the same dynamics on numpy
arrays for the CPU oracle.  Sub-environment i is the same whatever `num_envs` is."""

import numpy as np
import torch


class ToyVecEnv:
    def __init__(self, num_envs: int, n_obs: int, n_act: int, *, seed: int = 0, device="cpu", max_len: int = 9):
        g = np.random.default_rng(seed)
        self.act_length = n_act
        self.A = torch.as_tensor(g.standard_normal((n_obs, n_obs)) * 0.4 / np.sqrt(n_obs), dtype=torch.float32, device=device)
        self.B = torch.as_tensor(g.standard_normal((n_obs, n_act)) * 0.5, dtype=torch.float32, device=device)
        self.c = torch.as_tensor(g.standard_normal(n_obs) * 0.8, dtype=torch.float32, device=device)
        self.x0 = torch.as_tensor(np.random.default_rng(seed + 1).standard_normal((num_envs, n_obs)) * 1.5 + 0.5, dtype=torch.float32, device=device)
        self.lengths = torch.as_tensor(3 + (np.arange(num_envs) * 7 + seed) % (max_len - 2), dtype=torch.int64, device=device)
        self.x, self.t = None, None

    def reset(self):
        self.x = self.x0.clone()
        self.t = torch.zeros_like(self.lengths)
        return self.x.clone()

    def step(self, actions):
        a = torch.as_tensor(actions, dtype=torch.float32, device=self.x.device)
        self.x = 2.0 * torch.tanh(self.x @ self.A.T + a @ self.B.T + self.c) + 0.05
        reward = 1.0 - self.x.square().mean(dim=1) + 0.1 * a[:, 0]
        self.t += 1
        done = self.t >= self.lengths
        self.x = torch.where(done[:, None], self.x0, self.x)
        self.t = torch.where(done, torch.zeros_like(self.t), self.t)
        return self.x.clone(), reward, done, {}

    def as_numpy(self):
        return _NumpyView(self)


class _NumpyView:
    def __init__(self, env: ToyVecEnv):
        self.env = ToyVecEnv.__new__(ToyVecEnv)
        for k in ("A", "B", "c", "x0", "lengths"):
            setattr(self.env, k, getattr(env, k).detach().cpu().clone())
        self.env.x = self.env.t = None

    def reset(self):
        return self.env.reset().numpy()

    def step(self, actions):
        o, r, d, i = self.env.step(torch.as_tensor(np.asarray(actions, dtype=np.float32)))
        return o.numpy(), r.numpy(), d.numpy(), i
