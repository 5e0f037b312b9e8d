"""Time PGPE generations with a lazy (never materialised) population: python scripts/lazy_bench.py POPSIZE DIM [STEPS].
Prints one JSON line per configuration (CUDA-event timed, after 3 warm-up generations)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from evotorch_b200 import Problem, ops  # noqa: E402
from evotorch_b200.algorithms import PGPE  # noqa: E402
from evotorch_b200.objectives import rastrigin  # noqa: E402


def run(popsize: int, dim: int, steps: int, lazy: bool) -> dict:
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    prob = Problem("min", rastrigin, initial_bounds=(-5.12, 5.12), solution_length=dim, device="cuda", seed=1, lazy_population=lazy)
    s = PGPE(prob, popsize=popsize, center_learning_rate=0.5, stdev_learning_rate=0.1, stdev_init=1.0)
    for _ in range(3):
        s.step()
    ops.enable_timers()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(steps):
        s.step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    timers = {k: round(v[1], 4) for k, v in ops.timer_results().items()}
    ops.disable_timers()
    return {"popsize": popsize, "dim": dim, "lazy": lazy, "ms_per_generation": round(ms, 3), "generations_per_s": round(1e3 / ms, 3),
            "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 3), "population_gb_if_materialised": round(popsize * dim * 4 / 2 ** 30, 1),
            "kernel_ms": timers, "mean_eval": float(s.status["mean_eval"])}


if __name__ == "__main__":
    popsize, dim = int(sys.argv[1]), int(sys.argv[2])
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    modes = [True] if popsize * dim * 4 > 150 * 2 ** 30 else [False, True]
    for lazy in modes:
        print(json.dumps(run(popsize, dim, steps, lazy)), flush=True)
