#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -25 | tee gpurun_out/r2_pytest_gpu_b.txt
echo "== pytest cuda-parametrised NE"; timeout 600 python -m pytest tests/test_neproblem.py -q 2>&1 | tail -5
echo "== rank latency"; timeout 300 python scripts/rank_bench.py 2>&1 | tail -15 | tee gpurun_out/r2_rank_latency.txt
