"""Searcher configurations shared by the golden generator (run against the reference) and the tests (run against the
package): tag -> (algorithm, dim, sense, objective name, keyword arguments, generations)."""

import math

import torch

VARIANTS = {
    "pgpe_sgd_momentum_radius_bounds": ("PGPE", 8, "min", "rastrigin", dict(popsize=32, center_learning_rate=0.1, stdev_learning_rate=0.1, radius_init=4.5,
                                                                            optimizer="sgd", optimizer_config={"momentum": 0.5}, stdev_min=0.5, stdev_max=1.7), 6),
    "pgpe_linear_no_max_change": ("PGPE", 8, "min", "rastrigin", dict(popsize=32, center_learning_rate=0.4, stdev_learning_rate=0.3, stdev_init=1.0,
                                                                      ranking_method="linear", stdev_max_change=None), 6),
    "pgpe_clipup_config_normalized": ("PGPE", 8, "max", "rastrigin", dict(popsize=32, center_learning_rate=0.3, stdev_learning_rate=0.1, stdev_init=0.8,
                                                                          optimizer="clipup", optimizer_config={"max_speed": 0.45, "momentum": 0.8},
                                                                          ranking_method="normalized"), 6),
    "pgpe_raw_nonsym_plain": ("PGPE", 6, "min", "sphere", dict(popsize=25, center_learning_rate=0.02, stdev_learning_rate=0.01, stdev_init=0.5, symmetric=False,
                                                               optimizer=None, ranking_method="raw"), 5),
    "snes_noscale_lrs_radius": ("SNES", 8, "min", "rastrigin", dict(popsize=20, radius_init=5.0, center_learning_rate=0.7, stdev_learning_rate=0.2,
                                                                   scale_learning_rate=False, stdev_min=0.3), 6),
    "snes_default_popsize_adam": ("SNES", 10, "min", "rastrigin", dict(stdev_init=1.5, optimizer="adam", center_learning_rate=0.05, ranking_method="centered"), 6),
    "cem_bounds": ("CEM", 8, "min", "rastrigin", dict(popsize=40, parenthood_ratio=0.5, radius_init=6.0, stdev_min=0.8, stdev_max=2.5, stdev_max_change=0.15), 6),
    "cem_max": ("CEM", 6, "max", "rastrigin", dict(popsize=30, parenthood_ratio=0.2, stdev_init=1.0), 5),
    "xnes_radius_lrs": ("XNES", 5, "min", "sphere", dict(popsize=14, radius_init=3.0, center_learning_rate=0.8, stdev_learning_rate=0.1, scale_learning_rate=False), 5),
}


def objective(name: str):
    if name == "sphere":
        return lambda x: torch.sum(x**2, dim=-1)
    return lambda x: 10 * x.shape[-1] + torch.sum(x**2 - 10 * torch.cos(2 * math.pi * x), dim=-1)
