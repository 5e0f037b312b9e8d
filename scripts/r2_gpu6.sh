#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest functional + batched"; timeout 900 python -m pytest tests/test_functional_api.py tests/test_gpu_parity.py -q --maxfail=10 -k "functional or batched or cuda" 2>&1 | tail -25
echo "== functional bench"; timeout 300 python scripts/functional_bench.py 64 1000 100 2>&1 | tail -3 | tee gpurun_out/r2_functional_bench.txt
timeout 300 python scripts/functional_bench.py 16 10000 1000 2>&1 | tail -1 | tee -a gpurun_out/r2_functional_bench.txt
