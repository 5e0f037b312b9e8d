#!/bin/bash
# full GPU suite + smoke + default bench line with the final library
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r2_pytest_gpu.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench default"; timeout 900 python bench.py > gpurun_out/r2_bench_1gpu_final.json 2> gpurun_out/r2_bench_1gpu_final.err; tail -c 300 gpurun_out/r2_bench_1gpu_final.json; tail -3 gpurun_out/r2_bench_1gpu_final.err
