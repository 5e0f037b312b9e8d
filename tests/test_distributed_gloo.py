"""World-size-2 (gloo, CPU) test of the sharded generation: all-gather of fitnesses -> GLOBAL ranking -> all-reduce of the
partial gradients -> replicated update must reproduce the single-process run on the same population."""

import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from evotorch_b200 import Problem
from evotorch_b200.algorithms import CEM, PGPE, SNES
from evotorch_b200.distributed import shard_rows


def rastrigin(x):
    n = x.shape[1]
    return 10 * n + torch.sum((x**2) - 10 * torch.cos(2 * np.pi * x), 1)


class GlobalPopulationProblem(Problem):
    """A CPU problem whose populations are slices of a globally defined matrix (a function of the generation index, the
    distribution and the GLOBAL row index) -- the property the Philox sampler gives the CUDA path."""

    def __init__(self, symmetric, popsize, **kw):
        super().__init__("min", rastrigin, initial_bounds=(-5.12, 5.12), solution_length=6, vectorized=True, seed=3, **kw)
        self._sym, self._popsize, self._gen_index = symmetric, popsize, 0

    def sample_and_evaluate(self, distribution, batch):
        g = torch.Generator().manual_seed(1000 + self._gen_index)
        self._gen_index += 1
        n, d = self._popsize, self.solution_length
        if self._sym:
            z = torch.randn(n // 2, d, generator=g)
            x = torch.empty(n, d)
            x[0::2] = distribution.mu + distribution.sigma * z
            x[1::2] = distribution.mu - distribution.sigma * z
        else:
            x = distribution.mu + distribution.sigma * torch.randn(n, d, generator=g)
        r0 = self.philox_row0
        batch.access_values()[:] = x[r0:r0 + len(batch)]
        self.evaluate(batch)


def make_searcher(kind, popsize, distributed):
    sym = kind == "pgpe"
    prob = GlobalPopulationProblem(sym, popsize)
    mu0 = torch.linspace(-2, 2, 6)
    if kind == "pgpe":
        return PGPE(prob, popsize=popsize, center_learning_rate=0.4, stdev_learning_rate=0.1, stdev_init=1.0, center_init=mu0,
                    distributed=distributed)
    if kind == "pgpe_nes_nonsym":
        return PGPE(prob, popsize=popsize, center_learning_rate=0.4, stdev_learning_rate=0.1, stdev_init=1.0, center_init=mu0,
                    symmetric=False, ranking_method="nes", optimizer="adam", distributed=distributed)
    if kind == "snes":
        return SNES(prob, popsize=popsize, stdev_init=1.0, center_init=mu0, ranking_method="centered", distributed=distributed)
    return CEM(prob, popsize=popsize, parenthood_ratio=0.3, stdev_init=1.0, center_init=mu0, distributed=distributed)


CASES = [("pgpe", 26), ("pgpe_nes_nonsym", 25), ("snes", 24), ("cem", 31)]


def _worker(rank, world, init_file, out_dir):
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        for kind, popsize in CASES:
            s = make_searcher(kind, popsize, distributed=True)
            assert s._distributed
            traj = []
            for _ in range(4):
                s.step()
                traj.append(torch.cat([s.status["center"], s.status["stdev"], torch.tensor([s.status["mean_eval"]])]))
            torch.save(torch.stack(traj), os.path.join(out_dir, f"{kind}_{rank}.pt"))
        # unseeded problem, no centre given, torch RNG: every rank would start from its own centre and (with equal seeds) draw the
        # same noise -- the searcher must adopt rank 0's state and the ranks must sample different shards
        for seed in (None, 5):
            prob = Problem("min", rastrigin, initial_bounds=(-5.12, 5.12), solution_length=6, vectorized=True, seed=seed)
            s = PGPE(prob, popsize=20, center_learning_rate=0.4, stdev_learning_rate=0.1, stdev_init=1.0, distributed=True)
            for _ in range(3):
                s.step()
            shard = next(iter(prob._grad_batches.values())).values.clone()
            torch.save({"state": torch.cat([s.status["center"], s.status["stdev"], s._optimizer._velocity]), "shard": shard},
                       os.path.join(out_dir, f"unseeded_{seed}_{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_shard_rows_never_splits_pairs():
    for popsize, ws, sym in ((26, 2, True), (1_000_000, 8, True), (25, 2, False), (10, 3, True)):
        total, prev_end = 0, 0
        for r in range(ws):
            row0, n, counts = shard_rows(popsize, ws, r, sym)
            assert row0 == prev_end and (not sym or (row0 % 2 == 0 and n % 2 == 0)) and counts[r] == n
            prev_end, total = row0 + n, total + n
        assert total == popsize
    with pytest.raises(ValueError):
        shard_rows(25, 2, 0, True)


@pytest.mark.timeout(300)
def test_two_rank_generation_equals_single_process():
    world = 2
    with tempfile.TemporaryDirectory() as tmp:
        init_file = os.path.join(tmp, "rendezvous")
        mp.spawn(_worker, args=(world, init_file, tmp), nprocs=world, join=True)
        for kind, popsize in CASES:
            # single-process run on the same global populations: the distributed protocol with world size 1
            s = make_searcher(kind, popsize, distributed=False)
            s._step = s._step_distributed
            s._distributed = True
            ref = []
            for _ in range(4):
                s.step()
                ref.append(torch.cat([s.status["center"], s.status["stdev"], torch.tensor([float(s._mean_eval)])]))
            ref = torch.stack(ref).numpy()
            r0 = torch.load(os.path.join(tmp, f"{kind}_0.pt")).numpy()
            r1 = torch.load(os.path.join(tmp, f"{kind}_1.pt")).numpy()
            np.testing.assert_array_equal(r0, r1)  # replicated update: every rank holds identical mu / sigma
            np.testing.assert_allclose(r0, ref, rtol=2e-5, atol=2e-6, err_msg=kind)
        for seed in (None, 5):
            a = torch.load(os.path.join(tmp, f"unseeded_{seed}_0.pt"))
            b = torch.load(os.path.join(tmp, f"unseeded_{seed}_1.pt"))
            assert torch.equal(a["state"], b["state"])  # one replicated distribution / optimizer state
            assert not torch.equal(a["shard"], b["shard"])  # but different samples in the two shards
