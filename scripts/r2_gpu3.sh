#!/bin/bash
# Round-2 third GPU pass (1 GPU): GEMM with in-kernel lo derivation, fused CMA-ES generation, graph replay; cfg3 timing.
set -u
mkdir -p gpurun_out
echo "== pytest gemm / cmaes / xnes"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -q --maxfail=10 -k "gemm or cmaes or xnes or rank_table" 2>&1 | tail -40 | tee gpurun_out/r2_pytest_cmaes.txt
echo "== cfg3 timing"; timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/r2_cfg3.txt
import json, torch, time
from evotorch_b200 import Problem, ops
from evotorch_b200.algorithms import CMAES
from evotorch_b200.objectives import sphere
dev = torch.device("cuda", 0)
def run(graph, fused=True, K=50):
    prob = Problem("min", sphere, initial_bounds=(-3, 3), solution_length=1024, device=dev, seed=0)
    c = CMAES(prob, stdev_init=1.0, popsize=4096)
    if not fused: c._fused_ok = lambda: False
    if graph: c.enable_cuda_graph()
    for _ in range(6): c.step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K): c.step()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / K, c._graph is not None, float(c.status["mean_eval"])
for name, kw in (("op-by-op (round 1 path)", dict(graph=False, fused=False)), ("fused eager", dict(graph=False)), ("fused + CUDA graph", dict(graph=True))):
    ms, g, me = run(**kw)
    print(json.dumps({"cfg3": name, "ms_per_generation": ms, "generations_per_s": 1e3 / ms, "graph_captured": g, "mean_eval": me}), flush=True)
# GEMM alone, both shapes, convert on/off
import os
for (M, N, K) in ((4096, 1024, 1024), (1024, 1024, 4096), (8192, 8192, 8192)):
    A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev); C = torch.empty(M, N, device=dev)
    for _ in range(3): ops.gemm_nt(A, B, C)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): ops.gemm_nt(A, B, C)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(json.dumps({"gemm": [M, N, K], "ms": ms, "fp32_equiv_tflops": 2.0 * M * N * K / ms / 1e9, "tf32_tensor_tflops": 6.0 * M * N * K / ms / 1e9}), flush=True)
# Cholesky alone
Cm = torch.randn(1024, 1024, device=dev); Cm = Cm @ Cm.T + 1024 * torch.eye(1024, device=dev)
L = torch.empty_like(Cm); info = torch.zeros((), dtype=torch.int32, device=dev)
for _ in range(3): torch.linalg.cholesky_ex(Cm, check_errors=False, out=(L, info))
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): torch.linalg.cholesky_ex(Cm, check_errors=False, out=(L, info))
b.record(); torch.cuda.synchronize()
print(json.dumps({"cholesky_ex_1024_ms": a.elapsed_time(b) / 20}))
PY
