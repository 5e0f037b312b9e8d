"""More seeded trajectories of the reference's Gaussian searchers, one per option the first golden file does not exercise:

    PYTHONPATH=tests/golden/_refstubs:/root/reference/src EVOTORCH_VERBOSE_LEVEL=0 python tests/golden/gen_searcher_variants_golden.py

(CPU float32, seed 11; the package's CPU path draws from the same torch generator stream, so the runs compare step by step.)
"""

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import evotorch  # noqa: E402
from evotorch import Problem  # noqa: E402
from evotorch.algorithms import CEM, PGPE, SNES, XNES  # noqa: E402

from searcher_variants import VARIANTS, objective  # noqa: E402

assert "/root/reference" in evotorch.__file__
ALGOS = {"PGPE": PGPE, "SNES": SNES, "CEM": CEM, "XNES": XNES}
out = {}
for tag, (algo, D, sense, fn, kw, gens) in VARIANTS.items():
    prob = Problem(sense, objective(fn), initial_bounds=(-5.12, 5.12), solution_length=D, vectorized=True, seed=11, dtype=torch.float32)
    s = ALGOS[algo](prob, **kw)
    mus, sigs, fs, xs = [], [], [], []
    for _ in range(gens):
        s.step()
        mus.append(s.status["center"].numpy().copy()); sigs.append(s.status["stdev"].numpy().copy())
        fs.append(s.population.evals[:, 0].numpy().copy()); xs.append(s.population.values.numpy().copy())
    out[f"{tag}/mu"], out[f"{tag}/sigma"], out[f"{tag}/f"], out[f"{tag}/X"] = np.stack(mus), np.stack(sigs), np.stack(fs), np.stack(xs)
    out[f"{tag}/popsize"] = np.array(len(s.population))
np.savez_compressed(os.path.join(HERE, "searcher_variants_golden.npz"), **out)
print("wrote", len(out), "arrays", file=sys.stderr)
