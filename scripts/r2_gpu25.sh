#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -x -q -k "shared_minibatch or gemm or cmaes or xnes or syrk" 2>&1 | tail -2
timeout 300 python scripts/sf_bench.py 65536 | cut -c1-90
timeout 300 python scripts/gemm_bench.py | cut -c1-120
