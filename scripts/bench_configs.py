"""Secondary configurations of BASELINE.json (configs[2], configs[3]) on one GPU: CMA-ES generations/s at D=1024, N=4096 with a
per-stage breakdown, and the batched policy forward at 65 536 x MLP(376-256-17).  Writes gpurun_out/configs.json."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from evotorch_b200 import Problem, ops  # noqa: E402
from evotorch_b200.algorithms import CMAES  # noqa: E402
from evotorch_b200.neuroevolution import Policy  # noqa: E402
from evotorch_b200.objectives import sphere  # noqa: E402

dev = torch.device("cuda", 0)
out = {}


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


# ---- cfg3: CMA-ES sphere D=1024 popsize=4096
prob = Problem("min", sphere, initial_bounds=(-3, 3), solution_length=1024, device=dev, seed=0)
c = CMAES(prob, stdev_init=1.0, popsize=4096)
for _ in range(5):
    c.step()
torch.cuda.synchronize()
K = 30
a = ev()
for _ in range(K):
    c.step()
b = ev()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / K
out["cfg3_cmaes"] = {"ms_per_generation": ms, "generations_per_s": 1e3 / ms, "popsize": 4096, "dim": 1024,
                     "flop_model": 2 * 2 * 4096 * 1024 * 1024 + 1024**3 / 3}
# stage breakdown (each stage timed in isolation, synchronised)
stages = {}


def timed(name, fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    x = ev()
    for _ in range(reps):
        r = fn()
    y = ev()
    torch.cuda.synchronize()
    stages[name] = x.elapsed_time(y) / reps
    return r


zs, ys, xs = timed("sample_distribution (Philox z, Y = Z A^T, X = m + sigma Y)", c.sample_distribution)
aw = timed("evaluate + argsort + weight gather", lambda: c.get_population_weights(xs))
timed("update_m (2 weighted row sums, K4)", lambda: c.update_m(zs, ys, aw))
h = c._h_sig()
timed("update_C (weighted SYRK Y^T diag(w) Y + rank-1)", lambda: c.update_C(zs, ys, aw, h))
timed("decompose_C (Cholesky, cuSOLVER)", c.decompose_C)
timed("matmul Z A^T only", lambda: zs @ c.A.T)
timed("syrk (Y^T*w) @ Y only", lambda: (ys.T * aw) @ ys)
out["cfg3_cmaes"]["stage_ms"] = stages
print(json.dumps(out["cfg3_cmaes"], indent=1), flush=True)

# ---- cfg4: batched policy forward, B = 1 observation per policy
net = torch.nn.Sequential(torch.nn.Linear(376, 256), torch.nn.Tanh(), torch.nn.Linear(256, 17))
pol = Policy(net)
NP = 65536
P = torch.empty(NP, pol.parameter_length, device=dev).normal_(0, 0.1)
obs = torch.randn(NP, 376, device=dev)
pol.set_parameters(P)
for _ in range(3):
    pol(obs)
torch.cuda.synchronize()
a = ev()
for _ in range(10):
    act = pol(obs)
b = ev()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
gb = 4.0 * NP * pol.parameter_length / 1e9
out["cfg4_policy_forward"] = {"ms": ms, "forwards_per_s": 1e3 / ms, "gbs": gb / ms * 1e3, "policies": NP, "params": pol.parameter_length,
                              "observations_per_policy": 1, "activation": "tanh"}
print(json.dumps(out["cfg4_policy_forward"], indent=1), flush=True)
# ---- small / medium problems: eager stepping vs CUDA-graph replay (launch-bound regime)
from evotorch_b200.algorithms import PGPE, SNES  # noqa: E402
from evotorch_b200.objectives import rastrigin  # noqa: E402


def gens_per_s(make, graph, K=300):
    s = make()
    if graph:
        s.enable_cuda_graph()
    s.run(10)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.run(K)
    torch.cuda.synchronize()
    return K / (time.perf_counter() - t0)


small = {}
for tag, make in (("snes_cfg1_shape_N1000_D100", lambda: SNES(Problem("min", rastrigin, initial_bounds=(-5.12, 5.12), solution_length=100, device=dev, seed=1),
                                                              popsize=1000, stdev_init=10.0)),
                  ("pgpe_N10000_D1000", lambda: PGPE(Problem("min", rastrigin, initial_bounds=(-5.12, 5.12), solution_length=1000, device=dev, seed=1),
                                                     popsize=10000, center_learning_rate=0.5, stdev_learning_rate=0.1, stdev_init=1.0)),
                  ("pgpe_N100000_D10000_cfg2", lambda: PGPE(Problem("min", rastrigin, initial_bounds=(-5.12, 5.12), solution_length=10000, device=dev, seed=1),
                                                            popsize=100000, center_learning_rate=0.5, stdev_learning_rate=0.1, stdev_init=1.0))):
    small[tag] = {"eager_gen_per_s": gens_per_s(make, False), "cuda_graph_gen_per_s": gens_per_s(make, True)}
    print(tag, small[tag], flush=True)
out["eager_vs_cuda_graph"] = small

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
