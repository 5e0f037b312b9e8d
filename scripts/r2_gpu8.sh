#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest cholesky / cmaes"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -q --maxfail=10 -k "cholesky or cmaes" 2>&1 | tail -30
echo "== timing"; timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/r2_cholesky.txt
import json, os, torch
from evotorch_b200 import Problem, ops
from evotorch_b200.algorithms import CMAES
from evotorch_b200.objectives import sphere
dev = torch.device("cuda", 0)
def t(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for n in (256, 512, 1024, 2048, 4096):
    B = torch.randn(n, n, device=dev); A = B @ B.T / n + torch.eye(n, device=dev)
    L = torch.empty_like(A); info = torch.zeros((), dtype=torch.int32, device=dev)
    ours = t(lambda: ops.cholesky(A, out=L))
    lib = t(lambda: torch.linalg.cholesky_ex(A, check_errors=False, out=(L, info)))
    print(json.dumps({"cholesky_n": n, "evok_ms": ours, "cusolver_ms": lib, "speedup": lib / ours}), flush=True)
def run(graph, K=50):
    prob = Problem("min", sphere, initial_bounds=(-3, 3), solution_length=1024, device=dev, seed=0)
    c = CMAES(prob, stdev_init=1.0, popsize=4096)
    if graph: c.enable_cuda_graph()
    for _ in range(6): c.step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K): c.step()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / K, c._graph is not None, float(c.status["mean_eval"])
for lib in ("0", "1"):
    os.environ["EVOTORCH_B200_EVOK_CHOLESKY"] = "0" if lib == "1" else "1"
    for graph in (False, True):
        ms, g, me = run(graph)
        print(json.dumps({"cfg3": "cuSOLVER potrf" if lib == "1" else "evok_cholesky", "cuda_graph": g, "ms_per_generation": ms, "generations_per_s": 1e3 / ms, "mean_eval": me}), flush=True)
PY
