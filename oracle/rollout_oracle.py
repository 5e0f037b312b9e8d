"""TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's vectorised rollout loop
(`VecGymNE._evaluate_subbatch`, neuroevolution/vecgymne.py:744-916, with `_normalize_observation` :604-647 and
`RunningNorm` net/runningnorm.py:229-533), written with the reference's own operation order: boolean-mask gather of the
active observations, `update_and_normalize` on the gathered rows, scatter back into a clone, policy on the full batch.

PINNED: `tests/golden/rollout_golden.npz` holds scores, interaction / episode counters and observation statistics produced by the
REAL `VecGymNE` loop (tests/golden/gen_rollout_golden.py runs it unmodified on the toy environment; gymnasium itself is absent,
the functional stand-ins under tests/golden/_refstubs/gymnasium are enough for TorchWrapper / VecGymNE), and
tests/test_rollout.py::test_rollout_oracle_matches_the_real_reference_rollouts checks this restatement against it.  RunningNorm is
pinned separately (runningnorm_golden.npz), the policy is the numpy MLP of es_oracle (pinned against the reference's `Policy`).
"""

from typing import Optional

import numpy as np

from . import es_oracle as O


class RunningNormOracle:
    """net/runningnorm.py:229-533 in float32 numpy."""

    def __init__(self, n: int, min_variance: float = 1e-2, clip: Optional[tuple] = None):
        self.sum = np.zeros(n, dtype=np.float32)
        self.sumsq = np.zeros(n, dtype=np.float32)
        self.count = 0
        self.min_variance, self.clip = np.float32(min_variance), clip

    def update(self, x: np.ndarray):
        x = x.astype(np.float32)
        self.sum = self.sum + x.sum(axis=0, dtype=np.float32)
        self.sumsq = self.sumsq + np.square(x).sum(axis=0, dtype=np.float32)
        self.count += x.shape[0]

    def normalize(self, x: np.ndarray) -> np.ndarray:
        mean = self.sum / np.float32(self.count)
        var = np.maximum(self.sumsq / np.float32(self.count) - np.square(mean), self.min_variance)
        y = (x.astype(np.float32) - mean) / np.sqrt(var)
        if self.clip is not None:
            y = np.clip(y, np.float32(self.clip[0]), np.float32(self.clip[1]))
        return y.astype(np.float32)


def rollout(params: np.ndarray, n_in: int, n_hidden: int, n_out: int, activation: str, env, *, num_episodes: int = 1, obs_norm: Optional[RunningNormOracle] = None,
            decrease_rewards_by: Optional[float] = None, alive_bonus_schedule: Optional[tuple] = None) -> tuple:
    """Returns (scores[num_solutions], interactions).  `env`: reset() / step(actions) on numpy arrays (vecgymne.py:744-916)."""
    num_solutions = params.shape[0]
    obs = np.asarray(env.reset(), dtype=np.float32)
    num_envs = obs.shape[0]
    p = params
    if num_solutions < num_envs:
        p = np.concatenate([params, np.repeat(params[:1], num_envs - num_solutions, axis=0)], axis=0)
    active = np.zeros(num_envs, dtype=bool)
    active[:num_solutions] = True
    num_eps = np.zeros(num_envs, dtype=np.int64)
    score = np.zeros(num_envs, dtype=np.float32)
    t_per_env = np.zeros(num_envs, dtype=np.int64)
    total = 0

    def normalize(o, mask):
        if obs_norm is None or not mask.any():
            return o
        sel = o[mask]
        obs_norm.update(sel)
        out = o.copy()
        out[mask] = obs_norm.normalize(sel)
        return out

    obs = normalize(obs, active)
    while True:
        act = O.mlp_policy_forward(p, obs, n_in, n_hidden, n_out, activation).astype(np.float32)
        obs, reward, done, _ = env.step(act)
        obs, reward, done = np.asarray(obs, dtype=np.float32), np.asarray(reward, dtype=np.float32), np.asarray(done, dtype=bool)
        if decrease_rewards_by is not None:
            reward = reward - np.float32(decrease_rewards_by)
        if alive_bonus_schedule is not None:
            t0, t1, bonus = alive_bonus_schedule
            t_per_env[active] += 1
            score[active & (t_per_env >= t1)] += np.float32(bonus)
            if t1 > t0:
                part = active & (t_per_env >= t0) & (t_per_env < t1)
                score[part] += np.float32(bonus) * ((t_per_env[part] - t0).astype(np.float32) / np.float32(t1 - t0))
            t_per_env[active & done] = 0
        score[active] += reward[active]
        total += int(active.sum())
        num_eps[done] += 1
        active[:num_solutions] = active[:num_solutions] & (num_eps[:num_solutions] < num_episodes)
        if not active[:num_solutions].any():
            break
        obs = normalize(obs, active)
    fit = score[:num_solutions]
    if num_episodes > 1:
        fit = fit / np.float32(num_episodes)
    return fit, total
