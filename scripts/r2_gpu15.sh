#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches_sf.csv python scripts/sf_probe.py > /dev/null 2>&1; grep -o 'evok::[a-z_0-9]*[^"]*"[^n]*ns","[0-9]*' gpurun_out/launches_sf.csv | sed 's/(CUtensor.*GemmParams)//' | tail -6
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_tail -s 1 -c 1 -f -o gpurun_out/prof_mlp_tail python scripts/sf_probe.py > gpurun_out/ncu_tail.log 2>&1; tail -2 gpurun_out/ncu_tail.log
