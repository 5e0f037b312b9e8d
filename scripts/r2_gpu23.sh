#!/bin/bash
set -u
mkdir -p gpurun_out
for i in 1 2; do
SF_LIB=evotorch_b200/lib/libevok_prev.so timeout 300 python scripts/sf_bench.py 65536 | cut -c1-80
timeout 300 python scripts/sf_bench.py 65536 | cut -c1-80
done
