#!/bin/bash
# Round-2 second GPU pass (1 GPU): new parity tests, K1 occupancy variants, default bench with other_configs, reference arm.
set -u
mkdir -p gpurun_out
echo "== pytest new parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -q --maxfail=10 2>&1 | tail -30 | tee gpurun_out/r2_pytest_parity.txt
echo "== kbench variants (400k x 10k)"; timeout 600 python scripts/kbench.py 400000 10000 2>&1 | tee gpurun_out/r2_kbench_variants.txt | tail -12
echo "== bench default"; timeout 900 python bench.py > gpurun_out/r2_bench_1gpu_b.json 2> gpurun_out/r2_bench_1gpu_b.err; tail -c 3500 gpurun_out/r2_bench_1gpu_b.json; tail -5 gpurun_out/r2_bench_1gpu_b.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; tail -c 2500 gpurun_out/r2_bench_ref.json; tail -3 gpurun_out/r2_bench_ref.err
