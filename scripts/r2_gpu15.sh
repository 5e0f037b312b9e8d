#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches_sf.csv python scripts/sf_probe.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_gather_persistent -s 1 -c 1 -f -o gpurun_out/prof_gather_gemm python scripts/sf_probe.py > gpurun_out/ncu_gg.log 2>&1; tail -2 gpurun_out/ncu_gg.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_tail -s 1 -c 1 -f -o gpurun_out/prof_mlp_tail python scripts/sf_probe.py > gpurun_out/ncu_tail.log 2>&1; tail -2 gpurun_out/ncu_tail.log
