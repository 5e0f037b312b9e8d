"""Golden rollouts from the REAL reference loop (`evotorch.neuroevolution.VecGymNE._evaluate_subbatch`, vecgymne.py:744-916),
driven on the toy vectorised environment of tests/vecenv_fixture.py.  gymnasium is not installed here; the functional stand-ins
under _refstubs/gymnasium (spaces.Box, vector.VectorEnv) are enough for the reference's TorchWrapper / VecGymNE to run unmodified.

    PYTHONPATH=tests/golden/_refstubs:/root/reference/src EVOTORCH_VERBOSE_LEVEL=0 python tests/golden/gen_rollout_golden.py
"""

import os
import sys

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import evotorch  # noqa: E402
from evotorch.neuroevolution import VecGymNE  # noqa: E402
from evotorch.neuroevolution.net.vecrl import BaseVectorEnv  # noqa: E402
from gymnasium.spaces import Box  # noqa: E402  (the stand-in)

from vecenv_fixture import ToyVecEnv  # noqa: E402

assert "/root/reference" in evotorch.__file__
N_OBS, N_HID, N_ACT = 12, 16, 4


class GymToy(BaseVectorEnv):
    """ToyVecEnv behind the gymnasium vector-env interface the reference expects (new 5-tuple step API, numpy arrays)."""

    def __init__(self, num_envs, seed=3):
        space = lambda n: Box(-np.inf, np.inf, shape=(n,), dtype=np.float32)  # noqa: E731
        super().__init__(num_envs, space(N_OBS), space(N_ACT))
        self.toy = ToyVecEnv(num_envs, N_OBS, N_ACT, seed=seed)

    def reset(self, **kwargs):
        return self.toy.reset().numpy(), {}

    def step(self, actions):
        obs, reward, done, info = self.toy.step(torch.as_tensor(np.asarray(actions)))
        return obs.numpy(), reward.numpy(), done.numpy(), np.zeros_like(done.numpy()), info

    def seed(self, value):
        pass


def make_net():
    return nn.Sequential(nn.Linear(N_OBS, N_HID), nn.Tanh(), nn.Linear(N_HID, N_ACT))


rng = np.random.default_rng(8)
out = {}
VARIANTS = {
    "plain": dict(),
    "normalised": dict(observation_normalization=True),
    "episodes_bonus": dict(observation_normalization=True, num_episodes=3, decrease_rewards_by=0.25, alive_bonus_schedule=(2, 5, 0.5)),
    "bonus_single_step": dict(num_episodes=2, alive_bonus_schedule=(3, 0.3)),
}
for tag, kw in VARIANTS.items():
    prob = VecGymNE(lambda num_envs, **k: GymToy(num_envs, **k), make_net(), env_config=dict(seed=3), **kw)
    length = prob.solution_length
    for call, n in enumerate((50, 41)):  # the second, smaller batch reuses the 50 sub-environments: 9 of them are padding
        params = (rng.standard_normal((n, length)) * 0.4).astype(np.float32)
        batch = prob.generate_batch(n)
        batch.access_values()[:] = torch.as_tensor(params)
        prob.evaluate(batch)
        out[f"{tag}/{call}/params"] = params
        out[f"{tag}/{call}/scores"] = batch.evals[:, 0].numpy().copy()
        out[f"{tag}/{call}/interactions"] = np.array(prob.interaction_count)
        out[f"{tag}/{call}/episodes"] = np.array(prob.episode_count)
        if kw.get("observation_normalization"):
            st = prob.get_observation_stats()
            out[f"{tag}/{call}/stats_sum"], out[f"{tag}/{call}/stats_sumsq"] = st.sum.numpy().copy(), st.sum_of_squares.numpy().copy()
            out[f"{tag}/{call}/stats_count"] = np.array(st.count)
np.savez_compressed(os.path.join(HERE, "rollout_golden.npz"), **out)
print("wrote", len(out), "arrays", file=sys.stderr)
