"""The vectorised rollout loop around the batched policy forward (mirrors `VecGymNE._evaluate_subbatch`, vecgymne.py:744-916).

`env` is any vectorised environment in the reference's convention: `reset() -> obs[num_envs, n_in]`,
`step(actions[num_envs, n_out]) -> (obs, reward[num_envs], done[num_envs], info)`, auto-resetting finished sub-environments,
all tensors on one device.  Solution i drives sub-environment i for `num_episodes` episodes; its score is the (mean) episode
return.  Per time step the reference gathers the active rows (`obs[mask]`: a host synchronisation), normalises a copy,
scatters it back and calls a vmapped module; here the statistics update is one K4 launch on the masked batch and the
normalisation, clipping and masking happen inside the K8 policy kernel while it loads the observation -- inactive policies
are not even read.  The only synchronisation left per step is the loop's own termination test.
"""

from __future__ import annotations

from typing import NamedTuple, Optional

import torch

from .policy import Policy
from .runningnorm import RunningNorm


class RolloutResult(NamedTuple):
    scores: torch.Tensor        # (num_solutions,) mean episode return
    interactions: int           # environment steps taken by active sub-environments
    episodes: int               # num_solutions * num_episodes


@torch.no_grad()
def rollout(policy: Policy, parameters: torch.Tensor, env, *, num_episodes: int = 1, obs_norm: Optional[RunningNorm] = None,
            update_stats: bool = True, collected_stats: Optional[RunningNorm] = None, decrease_rewards_by: Optional[float] = None,
            alive_bonus_schedule: Optional[tuple] = None, action_noise_stdev: Optional[float] = None,
            max_steps: Optional[int] = None) -> RolloutResult:
    num_solutions = parameters.shape[0]
    obs = env.reset()
    device = obs.device
    num_envs = obs.shape[0]
    params = parameters.to(device)
    if num_solutions > num_envs:
        raise ValueError(f"Received incompatible number of environments: {num_envs} for {num_solutions} solutions")
    if num_solutions < num_envs:  # surplus sub-environments replay solution 0 and never count (vecgymne.py:766-779)
        padded = torch.empty(num_envs, params.shape[1], dtype=params.dtype, device=device)
        padded[:num_solutions] = params
        padded[num_solutions:] = params[0]
        params = padded
    active = torch.zeros(num_envs, dtype=torch.bool, device=device)
    active[:num_solutions] = True
    policy.set_parameters(params)

    episodes_done = torch.zeros(num_envs, dtype=torch.int64, device=device)
    scores = torch.zeros(num_envs, dtype=torch.float32, device=device)
    if alive_bonus_schedule is not None:
        bonus_t0, bonus_t1, alive_bonus = alive_bonus_schedule
        steps_alive = torch.zeros(num_envs, dtype=torch.int64, device=device)
    interactions = torch.zeros((), dtype=torch.int64, device=device)

    def observe(o: torch.Tensor):
        if obs_norm is not None and update_stats:
            if collected_stats is not None:
                collected_stats.update(o, active)
            obs_norm.update(o, active)

    observe(obs)
    steps = 0
    while True:
        actions = policy(torch.as_tensor(obs, dtype=params.dtype), obs_norm=obs_norm, active=active)
        if action_noise_stdev is not None:  # uniform noise, exactly like vecgymne.py:844
            actions = actions + torch.rand_like(actions) * action_noise_stdev
        obs, reward, done, _ = env.step(actions)
        done = torch.as_tensor(done, dtype=torch.bool, device=device)
        reward = torch.as_tensor(reward, dtype=torch.float32, device=device)
        if decrease_rewards_by is not None:
            reward = reward - decrease_rewards_by
        if alive_bonus_schedule is not None:
            steps_alive += active
            full = active & (steps_alive >= bonus_t1)
            scores += full * alive_bonus
            if bonus_t1 > bonus_t0:
                partial = active & (steps_alive >= bonus_t0) & (steps_alive < bonus_t1)
                scores += partial * (alive_bonus * (steps_alive - bonus_t0).to(torch.float32) / float(bonus_t1 - bonus_t0))
            steps_alive.masked_fill_(active & done, 0)
        scores += torch.where(active, reward, torch.zeros_like(reward))
        interactions += active.sum()
        episodes_done += done
        active[:num_solutions] &= episodes_done[:num_solutions] < num_episodes
        steps += 1
        if (max_steps is not None and steps >= max_steps) or not bool(active.any()):  # the loop's one host synchronisation per step
            break
        observe(obs)

    fitnesses = scores[:num_solutions]
    if num_episodes > 1:
        fitnesses = fitnesses / num_episodes
    return RolloutResult(scores=fitnesses, interactions=int(interactions), episodes=num_solutions * int(num_episodes))
