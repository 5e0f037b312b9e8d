#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "checkpoint_resume" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -3
