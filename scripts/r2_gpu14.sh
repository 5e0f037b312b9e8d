#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "shared_minibatch" 2>&1 | tail -5
timeout 300 python scripts/sf_bench.py 65536 | tee gpurun_out/r2_sf_bench.json
timeout 300 python scripts/sf_bench.py 16384 | tee -a gpurun_out/r2_sf_bench.json
EVOK_GATHER_PERSISTENT=0 timeout 300 python scripts/sf_bench.py 16384 | tee -a gpurun_out/r2_sf_bench.json
