#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -x -q -k "shared_minibatch or gemm or cmaes or xnes or syrk" 2>&1 | tail -3
timeout 300 python scripts/sf_bench.py 65536 | tee gpurun_out/r2_sf_bench.json
timeout 300 python scripts/gemm_bench.py | tee gpurun_out/r2_gemm_bench.json
EVOK_GEMM_B_LO_TMA=0 timeout 300 python scripts/gemm_bench.py | tee -a gpurun_out/r2_gemm_bench.json
