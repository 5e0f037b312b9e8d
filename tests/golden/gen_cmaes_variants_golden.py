"""CMA-ES option variants from the REAL reference (same seeds -> same torch-generator stream on CPU, so the package's CPU
path can be compared step by step):

    PYTHONPATH=tests/golden/_refstubs:/root/reference/src EVOTORCH_VERBOSE_LEVEL=0 python tests/golden/gen_cmaes_variants_golden.py
"""

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))

import evotorch  # noqa: E402
from evotorch import Problem  # noqa: E402
from evotorch.algorithms import CMAES  # noqa: E402

assert "/root/reference" in evotorch.__file__


def sphere(x):
    return torch.sum(x**2, dim=-1)


def shifted(x):  # to be MAXIMISED
    return -torch.sum((x - 1.5) ** 2 * torch.arange(1, x.shape[-1] + 1, dtype=x.dtype), dim=-1)


VARIANTS = {
    "separable": dict(sense="min", fn="sphere", D=8, kw=dict(stdev_init=1.0, popsize=14, separable=True)),
    "no_active": dict(sense="min", fn="sphere", D=6, kw=dict(stdev_init=0.7, popsize=10, active=False)),
    "csa_squared_bounds": dict(sense="min", fn="sphere", D=6, kw=dict(stdev_init=1.0, popsize=12, csa_squared=True, stdev_min=0.6, stdev_max=1.1)),
    "maximise_default_popsize": dict(sense="max", fn="shifted", D=7, kw=dict(stdev_init=2.0)),
    "ratios_no_limit": dict(sense="min", fn="sphere", D=5, kw=dict(stdev_init=1.0, popsize=16, c_1_ratio=0.5, c_mu_ratio=2.0, c_sigma_ratio=1.5,
                                                                    damp_sigma_ratio=0.8, c_c_ratio=1.2, c_m=0.9, limit_C_decomposition=False)),
}
out = {}
for tag, v in VARIANTS.items():
    fn = sphere if v["fn"] == "sphere" else shifted
    prob = Problem(v["sense"], fn, initial_bounds=(-3, 3), solution_length=v["D"], vectorized=True, seed=11, dtype=torch.float32)
    cma = CMAES(prob, **v["kw"])
    rec = {k: [] for k in ("m", "sigma", "C", "p_sigma", "p_c", "f")}
    out[f"{tag}/popsize"] = np.array(cma.popsize)
    for _ in range(7):
        cma.step()
        rec["m"].append(cma.m.numpy().copy()); rec["sigma"].append(np.array(float(cma.sigma))); rec["C"].append(cma.C.numpy().copy())
        rec["p_sigma"].append(cma.p_sigma.numpy().copy()); rec["p_c"].append(cma.p_c.numpy().copy())
        rec["f"].append(cma.population.evals[:, 0].numpy().copy())
    for k, val in rec.items():
        out[f"{tag}/{k}"] = np.stack(val)
np.savez_compressed(os.path.join(HERE, "cmaes_variants_golden.npz"), **out)
print("wrote", len(out), "arrays", file=sys.stderr)
