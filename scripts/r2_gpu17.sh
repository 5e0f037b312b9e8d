#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "shared_minibatch" > gpurun_out/pytest_sf.log 2>&1; grep -n "Error\|assert\|^E " gpurun_out/pytest_sf.log | head -20; tail -3 gpurun_out/pytest_sf.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_gather_persistent -s 1 -c 1 -f -o gpurun_out/prof_gather_gemm python scripts/sf_probe.py > gpurun_out/ncu_gg.log 2>&1; tail -1 gpurun_out/ncu_gg.log
