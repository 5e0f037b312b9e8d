#!/bin/bash
# full GPU suite + shared-minibatch bench + launch list of the forward
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r2_pytest_gpu.txt
timeout 300 python scripts/sf_bench.py 65536 | tee gpurun_out/r2_sf_bench.json
timeout 300 python scripts/sf_bench.py 16384 | tee -a gpurun_out/r2_sf_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches_sf.csv python scripts/sf_probe.py > /dev/null 2>&1
grep -o 'evok::[a-z_0-9]*[^"]*"[^n]*ns","[0-9]*' gpurun_out/launches_sf.csv | sed 's/(.*)//' | tail -4
