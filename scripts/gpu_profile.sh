#!/bin/bash
# ncu evidence for the round: launch list of the bench command + one full capture of each HBM-bound kernel (1 GPU).
set -u
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B --steps 3 --warmup 3 > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sample_eval_kernel -s 3 -c 1 -f -o gpurun_out/prof_sample_eval $B --steps 2 --warmup 3 --popsize 200000 > gpurun_out/ncu_se.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:grad_partial -s 2 -c 1 -f -o gpurun_out/prof_grad $B --steps 2 --warmup 3 --popsize 200000 > gpurun_out/ncu_grad.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:radix_scatter -s 8 -c 1 -f -o gpurun_out/prof_scatter $B --steps 2 --warmup 3 > gpurun_out/ncu_sc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_forward -s 1 -c 1 -f -o gpurun_out/prof_mlp python scripts/kbench.py 20000 1000 > gpurun_out/ncu_mlp.log 2>&1
ls -la gpurun_out | head -30
