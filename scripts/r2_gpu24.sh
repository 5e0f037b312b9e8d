#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitize.py > gpurun_out/r2_sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/r2_sanitize_memcheck.log
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_pytest_gpu.txt 2>&1; tail -2 gpurun_out/r2_pytest_gpu.txt
