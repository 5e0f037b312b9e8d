#!/bin/bash
# Run on the B200 box through gpurun: smoke, GPU parity tests, a bench line, the ncu launch list and one full capture
# of the two HBM-bound kernels.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q --maxfail=40 -x 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.txt
echo "== bench cfg2 (100k x 10k)"; timeout 900 python bench.py --steps 20 --warmup 3 --popsize 100000 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -c 3000 gpurun_out/bench_cfg2.json; tail -5 gpurun_out/bench_cfg2.err
echo "== bench metric (1M x 10k)"; timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 3000 gpurun_out/bench_1gpu.json; tail -5 gpurun_out/bench_1gpu.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; tail -3 gpurun_out/ncu_list.log
echo "== ncu full: sample_eval"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:sample_eval_kernel -s 3 -c 1 -f -o gpurun_out/prof_sample_eval python bench.py --steps 2 --warmup 3 --popsize 100000 --no-e2e --no-cpu-baseline > gpurun_out/ncu_se.log 2>&1; tail -3 gpurun_out/ncu_se.log
echo "== ncu full: grad"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:grad_partial_kernel -s 2 -c 1 -f -o gpurun_out/prof_grad python bench.py --steps 2 --warmup 3 --popsize 100000 --no-e2e --no-cpu-baseline > gpurun_out/ncu_grad.log 2>&1; tail -3 gpurun_out/ncu_grad.log
ls -la gpurun_out
