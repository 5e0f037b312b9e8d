"""`VecNE`: a neuroevolution `Problem` whose solutions are flat policy parameters, evaluated by vectorised rollouts
(mirrors the evaluation side of `VecGymNE`, neuroevolution/vecgymne.py:84-1060, without its gym dependency: the environment
is any object following the reference's vectorised-env convention, or a factory `env(num_envs, **env_config)`).

One generation of `PGPE(VecNE(...))` on a GPU = K1 sampling of the N x L parameter matrix, then per time step one K4 launch
(masked observation statistics) + one K8 launch (normalise + clip + policy forward, finished policies skipped), then K3-K5.
"""

from __future__ import annotations

from typing import Callable, Iterable, Mapping, Optional, Union

import torch
from torch import nn

from ..core import Problem, Solution, SolutionBatch
from .policy import Policy, fill_parameters
from .rollout import rollout
from .runningnorm import RunningNorm


class VecNE(Problem):
    def __init__(self, env: Union[Callable, object], network: Union[Callable, nn.Module], *, env_config: Optional[Mapping] = None,
                 max_num_envs: Optional[int] = None, network_args: Optional[Mapping] = None, observation_normalization: bool = False,
                 decrease_rewards_by: Optional[float] = None, alive_bonus_schedule: Optional[tuple] = None,
                 action_noise_stdev: Optional[float] = None, num_episodes: int = 1, device=None, seed: Optional[int] = None):
        self._env_source, self._env_config = env, dict(env_config or {})
        if isinstance(network, nn.Module):
            if network_args:
                raise ValueError("`network_args` is expected as None when the network is given as a torch.nn.Module instance")
            net = network
        elif isinstance(network, str):
            # a structure string (net/parser.py): obs_length / act_length / obs_shape come from a one-environment probe, like the
            # reference's `_env_constants_for_str_net` (vecgymne.py:68-80)
            from .net.parser import str_to_net

            net = str_to_net(network, **dict(self._probe_env_constants(), **dict(network_args or {})))
        else:
            net = network(**dict(network_args or {}))
        self._env, self._env_size = None, None
        self._max_num_envs = None if max_num_envs is None else int(max_num_envs)
        self._policy = Policy(net)
        self._observation_normalization = bool(observation_normalization)
        self._decrease_rewards_by = None if decrease_rewards_by is None else float(decrease_rewards_by)
        self._alive_bonus_schedule = None
        if alive_bonus_schedule is not None:  # (t, bonus) or (t0, t1, bonus) (vecgymne.py:389-401)
            sched = tuple(alive_bonus_schedule)
            if len(sched) == 2:
                sched = (int(sched[0]), int(sched[0]), float(sched[1]))
            elif len(sched) == 3:
                sched = (int(sched[0]), int(sched[1]), float(sched[2]))
            else:
                raise ValueError(f"alive_bonus_schedule was expected to have 2 or 3 items, but it has {len(sched)}: {alive_bonus_schedule!r}")
            self._alive_bonus_schedule = sched
        self._action_noise_stdev = None if action_noise_stdev is None else float(action_noise_stdev)
        self._num_episodes = int(num_episodes)
        self._obs_stats: Optional[RunningNorm] = None
        self._collected_stats: Optional[RunningNorm] = None
        self._interaction_count = 0
        self._episode_count = 0
        super().__init__("max", initial_bounds=(-0.00001, 0.00001), solution_length=self._policy.parameter_length, device=device,
                         dtype=torch.float32, seed=seed)

    def _probe_env_constants(self) -> dict:
        env = self._env_source(1, **self._env_config) if callable(self._env_source) else self._env_source
        obs = env.reset()
        obs_shape = tuple(obs.shape[1:])
        act_length = None
        space = getattr(env, "single_action_space", None)
        if space is not None and getattr(space, "shape", None):
            act_length = int(space.shape[0])
        for name in ("act_length", "action_length", "n_act"):
            if act_length is None and hasattr(env, name):
                act_length = int(getattr(env, name))
        constants = {"obs_length": int(obs_shape[0]), "obs_shape": obs_shape}
        if act_length is not None:
            constants.update(act_length=act_length, act_shape=(act_length,))
        return constants

    # ------------------------------------------------------------------ bookkeeping (vecgymne.py:457-494)
    @property
    def observation_normalization(self) -> bool:
        return self._observation_normalization

    @property
    def interaction_count(self) -> int:
        return self._interaction_count

    @property
    def episode_count(self) -> int:
        return self._episode_count

    @property
    def max_num_envs(self) -> Optional[int]:
        return self._max_num_envs

    def _extra_status(self, batch: SolutionBatch) -> dict:
        return dict(total_interaction_count=self._interaction_count, total_episode_count=self._episode_count)

    def __getstate__(self) -> dict:
        state = super().__getstate__()
        state["_env"], state["_env_size"] = None, None  # environments are rebuilt on demand
        return state

    # ------------------------------------------------------------------ observation statistics (vecgymne.py:649-715)
    def _ensure_obsnorm(self):
        if not self._observation_normalization:
            raise ValueError("This feature can only be used when observation_normalization=True.")

    def get_observation_stats(self) -> Optional[RunningNorm]:
        self._ensure_obsnorm()
        return self._obs_stats

    def set_observation_stats(self, rn: RunningNorm):
        self._ensure_obsnorm()
        self._obs_stats = rn

    def pop_observation_stats(self) -> Optional[RunningNorm]:
        """The statistics collected since the last pop (what a worker would send to the main process)."""
        self._ensure_obsnorm()
        result, self._collected_stats = self._collected_stats, None
        return result

    def update_observation_stats(self, rn: RunningNorm):
        self._ensure_obsnorm()
        if self._obs_stats is None:
            self._obs_stats = rn
        else:
            self._obs_stats.update(rn)

    # ------------------------------------------------------------------ evaluation
    def _get_env(self, num_envs: int):
        if not callable(self._env_source):
            return self._env_source  # a ready-made vectorised environment: its size is what it is
        if self._env is None or num_envs > self._env_size:  # a larger environment is reused: surplus sub-environments are padding (vecgymne.py:497)
            self._env, self._env_size = self._env_source(num_envs, **self._env_config), num_envs
        return self._env

    def _evaluate_batch(self, batch: SolutionBatch):
        if self._max_num_envs is None or len(batch) <= self._max_num_envs:
            self._evaluate_subbatch(batch)
        else:
            for piece in batch.split(max_size=self._max_num_envs):
                self._evaluate_subbatch(piece)

    def _evaluate_subbatch(self, batch: SolutionBatch):
        n = len(batch)
        env = self._get_env(n if self._max_num_envs is None else min(n, self._max_num_envs) if callable(self._env_source) else n)
        values = batch.access_values(keep_evals=True)
        if self._observation_normalization and self._obs_stats is None:
            probe = env.reset()
            self._obs_stats = RunningNorm(shape=probe.shape[1:], dtype=torch.float32, device=probe.device)
        if self._observation_normalization and self._collected_stats is None:
            self._collected_stats = RunningNorm(shape=self._obs_stats.shape, dtype=torch.float32, device=self._obs_stats.device)
        result = rollout(self._policy, values, env, num_episodes=self._num_episodes,
                         obs_norm=self._obs_stats if self._observation_normalization else None, collected_stats=self._collected_stats,
                         decrease_rewards_by=self._decrease_rewards_by, alive_bonus_schedule=self._alive_bonus_schedule,
                         action_noise_stdev=self._action_noise_stdev)
        self._interaction_count += result.interactions
        self._episode_count += result.episodes
        batch.set_evals(result.scores.to(batch.device))

    # ------------------------------------------------------------------ exporting a solution (vecgymne.py:927-1062)
    def make_net(self, solution: Iterable) -> nn.Module:
        """A copy of the network carrying the parameters of `solution` (no wrappers)."""
        from copy import deepcopy

        if isinstance(solution, Solution):
            solution = solution.values
        vector = torch.as_tensor(solution, dtype=torch.float32).detach().to("cpu")
        net = deepcopy(self._policy._net).to("cpu")
        fill_parameters(net, vector)
        return net

    def to_policy(self, solution: Iterable, *, with_wrapper_modules: bool = True) -> nn.Module:
        """The network with `solution`'s parameters, preceded by the observation-normalisation layer when there is one."""
        net = self.make_net(solution)
        if with_wrapper_modules and self._observation_normalization and self._obs_stats is not None and self._obs_stats.sum is not None:
            return nn.Sequential(self._obs_stats.to("cpu").to_layer(), net)
        return net
