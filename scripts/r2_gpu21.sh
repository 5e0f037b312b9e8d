#!/bin/bash
set -u
mkdir -p gpurun_out
for d in 0 3; do EVOK_GATHER_DEBUG=$d timeout 300 python scripts/sf_bench.py 16384; done
