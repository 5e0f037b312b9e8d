"""RunningNorm, the policy kernel's fused observation normalisation / masking, and the rollout loop."""

import os

import numpy as np
import pytest
import torch
from torch import nn

from evotorch_b200.neuroevolution import ObsNormLayer, Policy, RunningNorm, rollout
from oracle import rollout_oracle as RO
from vecenv_fixture import ToyVecEnv

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "runningnorm_golden.npz"))
DEVICES = ["cpu", pytest.param("cuda", marks=pytest.mark.gpu)]


def T(x, device, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype, device=device)


@pytest.mark.parametrize("device", DEVICES)
@pytest.mark.parametrize("tag,clip", [("noclip", None), ("clip", (-1.5, 2.0))])
def test_running_norm_matches_reference(tag, clip, device):
    rn = RunningNorm(shape=9, dtype="float32", device=device, min_variance=1e-2, clip=clip)
    with pytest.raises(ValueError):
        rn.normalize(T(GOLD["probe"], device))
    probe = T(GOLD["probe"], device)
    for i in range(4):
        mask = GOLD[f"mask{i}"]
        rn.update(T(GOLD[f"batch{i}"], device), None if mask.all() and i == 0 else T(mask, device, torch.bool))
        assert rn.count == int(GOLD[f"{tag}/count{i}"])
        np.testing.assert_allclose(rn.sum.cpu().numpy(), GOLD[f"{tag}/sum{i}"], rtol=2e-6, atol=1e-5)
        np.testing.assert_allclose(rn.sum_of_squares.cpu().numpy(), GOLD[f"{tag}/sumsq{i}"], rtol=2e-6, atol=1e-4)
        np.testing.assert_allclose(rn.mean.cpu().numpy(), GOLD[f"{tag}/mean{i}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(rn.stdev.cpu().numpy(), GOLD[f"{tag}/stdev{i}"], rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(rn.normalize(probe).cpu().numpy(), GOLD[f"{tag}/norm{i}"], rtol=2e-4, atol=2e-5)
    rn.update(T(GOLD["single"], device))
    assert rn.count == int(GOLD[f"{tag}/count_single"])
    np.testing.assert_allclose(rn.normalize(probe).cpu().numpy(), GOLD[f"{tag}/norm_single"], rtol=2e-4, atol=2e-5)
    other = RunningNorm(shape=9, dtype="float32", device=device)
    other.update(T(GOLD["batch2"], device))
    rn.update(other)
    assert rn.count == int(GOLD[f"{tag}/count_merged"])
    np.testing.assert_allclose(rn.mean.cpu().numpy(), GOLD[f"{tag}/mean_merged"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rn.stdev.cpu().numpy(), GOLD[f"{tag}/stdev_merged"], rtol=2e-5, atol=1e-6)
    layer = rn.to_layer()
    assert isinstance(layer, ObsNormLayer)
    np.testing.assert_allclose(layer(probe).cpu().numpy(), GOLD[f"{tag}/layer"], rtol=2e-4, atol=2e-5)
    got = rn.update_and_normalize(T(GOLD["batch1"], device), T(GOLD["mask1"], device, torch.bool))
    np.testing.assert_allclose(got.cpu().numpy(), GOLD[f"{tag}/update_and_normalize"], rtol=2e-4, atol=2e-5)
    # numpy in -> numpy out; shape and mask validation (runningnorm.py:190-212, :287-307)
    assert isinstance(rn.normalize(GOLD["probe"]), np.ndarray)
    with pytest.raises(ValueError):
        rn.update(torch.zeros(4, 8, device=device))
    with pytest.raises(ValueError):
        rn.update(torch.zeros(4, 9, device=device), torch.ones(3, dtype=torch.bool, device=device))
    with pytest.raises(ValueError):
        rn.update(torch.zeros(9, device=device), torch.ones(1, dtype=torch.bool, device=device))
    moved = rn.to("cpu")
    assert moved.device.type == "cpu" and moved.count == rn.count
    rn.reset()
    assert rn.count == 0 and rn.sum is None


def _net(n_in, n_hidden, n_out):
    return nn.Sequential(nn.Linear(n_in, n_hidden), nn.Tanh(), nn.Linear(n_hidden, n_out))


@pytest.mark.parametrize("device", DEVICES)
def test_policy_call_with_normalisation_and_mask(device):
    g = np.random.default_rng(2)
    n, n_in, n_hidden, n_out = 37, 11, 8, 3
    policy = Policy(_net(n_in, n_hidden, n_out))
    params = T(g.standard_normal((n, policy.parameter_length)) * 0.3, device)
    obs = T(g.standard_normal((n, n_in)) * 3 + 1, device)
    active = T(g.random(n) < 0.7, device, torch.bool)
    rn = RunningNorm(shape=n_in, dtype="float32", device=device, clip=(-2.0, 2.0))
    rn.update(obs, active)
    policy.set_parameters(params)
    got = policy(obs, obs_norm=rn, active=active)
    from oracle import es_oracle as O

    ref_norm = RO.RunningNormOracle(n_in, clip=(-2.0, 2.0))
    ref_norm.update(obs.cpu().numpy()[active.cpu().numpy()])
    want = O.mlp_policy_forward(params.cpu().numpy(), ref_norm.normalize(obs.cpu().numpy()), n_in, n_hidden, n_out, "tanh")
    want[~active.cpu().numpy()] = 0.0
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-4, atol=1e-5)
    # mask alone / normalisation alone
    np.testing.assert_allclose(policy(obs, active=active).cpu().numpy()[active.cpu().numpy()],
                               policy(obs).cpu().numpy()[active.cpu().numpy()], rtol=0, atol=0)
    assert torch.count_nonzero(policy(obs, active=active)[~active]) == 0
    np.testing.assert_allclose(policy(obs, obs_norm=rn).cpu().numpy(), policy(rn.normalize(obs)).cpu().numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("device", DEVICES)
@pytest.mark.parametrize("variant", ["plain", "normalised", "episodes_bonus_padding"])
def test_rollout_matches_the_reference_loop(variant, device):
    g = np.random.default_rng(8)
    n_in, n_hidden, n_out = 12, 16, 4
    policy = Policy(_net(n_in, n_hidden, n_out))
    num_solutions = 50 if variant != "episodes_bonus_padding" else 41
    num_envs = 50
    params_np = (g.standard_normal((num_solutions, policy.parameter_length)) * 0.4).astype(np.float32)
    env = ToyVecEnv(num_envs, n_in, n_out, seed=3, device=device)
    kw, okw = {}, {}
    if variant != "plain":
        kw["obs_norm"] = RunningNorm(shape=n_in, dtype="float32", device=device, clip=(-5.0, 5.0))
        okw["obs_norm"] = RO.RunningNormOracle(n_in, clip=(-5.0, 5.0))
    if variant == "episodes_bonus_padding":
        kw.update(num_episodes=3, decrease_rewards_by=0.25, alive_bonus_schedule=(2, 5, 0.5))
        okw.update(num_episodes=3, decrease_rewards_by=0.25, alive_bonus_schedule=(2, 5, 0.5))
    result = rollout(policy, T(params_np, device), env, **kw)
    want_scores, want_steps = RO.rollout(params_np, n_in, n_hidden, n_out, "tanh", env.as_numpy(), **okw)
    assert result.interactions == want_steps and result.episodes == num_solutions * kw.get("num_episodes", 1)
    assert result.scores.shape == (num_solutions,)
    np.testing.assert_allclose(result.scores.cpu().numpy(), want_scores, rtol=2e-4, atol=2e-4)
    if "obs_norm" in kw:
        assert kw["obs_norm"].count == okw["obs_norm"].count
        np.testing.assert_allclose(kw["obs_norm"].sum.cpu().numpy(), okw["obs_norm"].sum, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("device", DEVICES)
def test_vecne_problem_drives_pgpe(device):
    """The neuroevolution problem class on top of the rollout loop (VecGymNE's evaluation side): PGPE improves the return,
    the counters and the observation statistics advance, sub-batching by `max_num_envs` gives the same scores, and the
    exported policy reproduces the batched kernel's actions."""
    from evotorch_b200.algorithms import PGPE
    from evotorch_b200.neuroevolution import VecNE
    from evotorch_b200 import SolutionBatch

    n_in, n_out = 10, 3
    make_env = lambda num_envs, **kw: ToyVecEnv(num_envs, n_in, n_out, device=device, **kw)  # noqa: E731
    prob = VecNE(make_env, _net(n_in, 12, n_out), env_config=dict(seed=5), observation_normalization=True, num_episodes=2,
                 decrease_rewards_by=0.1, device=device, seed=7)
    assert prob.solution_length == n_in * 12 + 12 + 12 * n_out + n_out and prob.senses == ["max"]
    searcher = PGPE(prob, popsize=64, center_learning_rate=0.05, stdev_learning_rate=0.1, stdev_init=0.1)
    searcher.step()
    first = searcher.status["mean_eval"]
    assert searcher.status["total_episode_count"] == 128 and searcher.status["total_interaction_count"] > 128
    searcher.run(25)
    assert searcher.status["mean_eval"] > first + 1.0
    stats = prob.get_observation_stats()
    assert stats.count == searcher.status["total_interaction_count"] and prob.pop_observation_stats().count == stats.count
    assert prob.pop_observation_stats() is None

    # same solutions, evaluated in sub-batches of at most 20 environments (fresh problems: the statistics start equal)
    def scores(max_envs):
        p = VecNE(make_env, _net(n_in, 12, n_out), env_config=dict(seed=5), max_num_envs=max_envs, device=device, seed=7)
        batch = SolutionBatch(p, popsize=50, empty=True)
        batch.set_values(searcher.population.values[:50].clone())
        p.evaluate(batch)
        return batch.evals.clone().view(-1)

    whole, pieces = scores(None), scores(20)
    # 50 solutions in pieces of <= 20 -> 17 + 17 + 16; sub-environment i has its own start state and episode length, so only the
    # first piece meets the same environments as the unsplit evaluation
    torch.testing.assert_close(pieces[:17], whole[:17], rtol=1e-5, atol=1e-5)
    assert not torch.allclose(pieces[17:34], whole[17:34])

    center = searcher.status["center"]
    module = prob.to_policy(center)
    obs = torch.randn(4, n_in)
    pol = Policy(_net(n_in, 12, n_out))
    pol.set_parameters(center.cpu().expand(4, -1).contiguous())
    want = pol(obs, obs_norm=stats.to("cpu"))
    torch.testing.assert_close(module(obs), want, rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        VecNE(make_env, _net(n_in, 12, n_out), device=device).get_observation_stats()


ROLLOUT_GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "rollout_golden.npz"))
ROLLOUT_VARIANTS = {
    "plain": dict(),
    "normalised": dict(observation_normalization=True),
    "episodes_bonus": dict(observation_normalization=True, num_episodes=3, decrease_rewards_by=0.25, alive_bonus_schedule=(2, 5, 0.5)),
    "bonus_single_step": dict(num_episodes=2, alive_bonus_schedule=(3, 0.3)),
}


@pytest.mark.parametrize("device", DEVICES)
@pytest.mark.parametrize("tag", list(ROLLOUT_VARIANTS))
def test_vecne_matches_the_real_reference_rollouts(tag, device):
    """`tests/golden/rollout_golden.npz` was produced by the reference's own `VecGymNE` loop (vecgymne.py:744-916) on the toy
    environment: scores, interaction / episode counters and the running observation statistics must agree, including the second
    call, where 41 solutions run in the 50 sub-environments left over from the first (9 padding environments)."""
    from evotorch_b200 import SolutionBatch
    from evotorch_b200.neuroevolution import VecNE

    n_obs, n_hid, n_act = 12, 16, 4
    prob = VecNE(lambda num_envs, **kw: ToyVecEnv(num_envs, n_obs, n_act, device=device, **kw), _net(n_obs, n_hid, n_act), env_config=dict(seed=3),
                 device=device, **ROLLOUT_VARIANTS[tag])
    for call, n in enumerate((50, 41)):
        g = lambda key: ROLLOUT_GOLD[f"{tag}/{call}/{key}"]  # noqa: E731
        batch = SolutionBatch(prob, popsize=n, empty=True)
        batch.set_values(T(g("params"), device))
        prob.evaluate(batch)
        np.testing.assert_allclose(batch.evals[:, 0].cpu().numpy(), g("scores"), rtol=3e-4, atol=3e-4)
        assert prob.interaction_count == int(g("interactions")) and prob.episode_count == int(g("episodes"))
        assert prob.status["total_interaction_count"] == prob.interaction_count
        if ROLLOUT_VARIANTS[tag].get("observation_normalization"):
            stats = prob.get_observation_stats()
            assert stats.count == int(g("stats_count"))
            np.testing.assert_allclose(stats.sum.cpu().numpy(), g("stats_sum"), rtol=1e-4, atol=2e-3)
            np.testing.assert_allclose(stats.sum_of_squares.cpu().numpy(), g("stats_sumsq"), rtol=1e-4, atol=2e-3)
    assert prob._env_size == 50  # the larger environment was reused


def test_rollout_oracle_matches_the_real_reference_rollouts():
    """Pins oracle/rollout_oracle.py to the same golden file."""
    n_obs, n_hid, n_act = 12, 16, 4
    for tag, kw in ROLLOUT_VARIANTS.items():
        okw = {k: v for k, v in kw.items() if k in ("num_episodes", "decrease_rewards_by")}
        if "alive_bonus_schedule" in kw:
            sched = kw["alive_bonus_schedule"]
            okw["alive_bonus_schedule"] = sched if len(sched) == 3 else (sched[0], sched[0], sched[1])
        norm = RO.RunningNormOracle(n_obs) if kw.get("observation_normalization") else None
        total = 0
        for call, n in enumerate((50, 41)):
            env = ToyVecEnv(50, n_obs, n_act, seed=3).as_numpy()
            scores, steps = RO.rollout(ROLLOUT_GOLD[f"{tag}/{call}/params"], n_obs, n_hid, n_act, "tanh", env, obs_norm=norm, **okw)
            total += steps
            np.testing.assert_allclose(scores, ROLLOUT_GOLD[f"{tag}/{call}/scores"], rtol=3e-4, atol=3e-4)
            assert total == int(ROLLOUT_GOLD[f"{tag}/{call}/interactions"])
            if norm is not None:
                assert norm.count == int(ROLLOUT_GOLD[f"{tag}/{call}/stats_count"])


def test_network_structure_strings():
    """net/parser.py: `str_to_net` (BASELINE config 4 writes its policy as "Linear(376, 256) >> Tanh() >> Linear(256, 17)"), the
    `MultiLayered` it builds, and a VecNE problem whose network is such a string with obs_length / act_length filled in from the
    environment.  (The reference's own tests/test_neuroevolution_net_parser.py passes against this parser.)"""
    from evotorch_b200.neuroevolution import VecNE
    from evotorch_b200.neuroevolution.net import MultiLayered, NetParsingError, str_to_net

    net = str_to_net("Linear(376, 256) >> Tanh() >> Linear(256, 17)")
    assert isinstance(net, MultiLayered) and len(net) == 3 and net[0].in_features == 376 and isinstance(net[1], nn.Tanh)
    policy = Policy(net)
    assert policy.parameter_length == 100_881 and policy._spec == ([376, 256, 17], ["tanh", "none"])  # the kernel path recognises it
    x = torch.randn(5, 376)
    ref = nn.Sequential(*list(net))
    torch.testing.assert_close(net(x), ref(x))
    with pytest.raises(NetParsingError, match="Unrecognized module class"):
        str_to_net("Nope(3)")
    with pytest.raises(NetParsingError, match=r"at line\(1\) at column\(8\): Unknown constant: n"):
        str_to_net("Linear(n, 2)")
    prob = VecNE(lambda num_envs, **kw: ToyVecEnv(num_envs, 10, 3, **kw), "Linear(obs_length, hidden) >> Tanh() >> Linear(hidden, act_length)",
                 network_args=dict(hidden=7), env_config=dict(seed=1))
    assert prob.solution_length == 10 * 7 + 7 + 7 * 3 + 3
