#!/bin/bash
# Round-2 first GPU pass: parity tests (incl. the new ones), default bench, ncu captures of the SHIPPED kernels at the metric size.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 2>&1 | tail -60 | tee gpurun_out/r2_pytest_gpu.txt
echo "== bench metric (1M x 10k)"; timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; tail -c 2500 gpurun_out/r2_bench_1gpu.json; tail -5 gpurun_out/r2_bench_1gpu.err
B="python bench.py --no-e2e --no-cpu-baseline"
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B --steps 3 --warmup 3 > gpurun_out/ncu_list.log 2>&1; tail -2 gpurun_out/ncu_list.log
echo "== ncu full: sample_eval @ metric size"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:sample_eval_kernel -s 3 -c 1 -f -o gpurun_out/prof_sample_eval $B --steps 2 --warmup 3 > gpurun_out/ncu_se.log 2>&1; tail -2 gpurun_out/ncu_se.log
echo "== ncu full: grad_partial_tma @ metric size"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:grad_partial_tma -s 2 -c 1 -f -o gpurun_out/prof_grad $B --steps 2 --warmup 3 > gpurun_out/ncu_grad.log 2>&1; tail -2 gpurun_out/ncu_grad.log
ls -la gpurun_out | head -40
