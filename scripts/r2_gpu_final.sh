#!/bin/bash
# Round-2 final single-GPU pass: smoke, default bench line (with other_configs + cpu_baseline), launch list of the bench command.
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench default"; timeout 900 python bench.py > gpurun_out/r2_bench_1gpu_final.json 2> gpurun_out/r2_bench_1gpu_final.err; tail -c 600 gpurun_out/r2_bench_1gpu_final.json; tail -3 gpurun_out/r2_bench_1gpu_final.err
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --no-e2e --no-cpu-baseline --no-other-configs --steps 3 --warmup 3 > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log | cut -c1-200
