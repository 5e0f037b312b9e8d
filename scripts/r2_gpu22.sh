#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/gather_trace.py | awk 'NR<=2 || (NR>=10 && NR<=26)'
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "shared_minibatch" 2>&1 | tail -2
timeout 300 python scripts/sf_bench.py 65536
