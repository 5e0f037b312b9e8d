#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "shared_minibatch" 2>&1 | tail -3
timeout 300 python scripts/sf_bench.py 65536
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches_sf.csv python scripts/sf_probe.py > /dev/null 2>&1
grep -o 'evok::[a-z_0-9]*[^"]*"[^n]*ns","[0-9]*' gpurun_out/launches_sf.csv | sed 's/(.*)//' | tail -3
