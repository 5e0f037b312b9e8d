#!/bin/bash
# Round-2 multi-GPU pass: metric workload with the sharded ranking (and without, for comparison), sharded_parity, BASELINE config 5.
# usage: bash scripts/r2_gpu_multi.sh <N>
set -u
N=${1:-2}
mkdir -p gpurun_out
run() { timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $N "$@"; }
echo "== metric, sharded rank"; EVOTORCH_B200_SHARDED_RANK=1 run --steps 40 --warmup 5 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; tail -c 1800 gpurun_out/r2_bench_${N}gpu.json; tail -3 gpurun_out/r2_bench_${N}gpu.err
echo "== metric, replicated rank (default)"; EVOTORCH_B200_SHARDED_RANK=0 run --steps 40 --warmup 5 --no-e2e --no-sharded-parity > gpurun_out/r2_bench_${N}gpu_replicated_rank.json 2> gpurun_out/r2_bench_${N}gpu_replicated_rank.err; tail -c 600 gpurun_out/r2_bench_${N}gpu_replicated_rank.json | head -c 400; echo; tail -2 gpurun_out/r2_bench_${N}gpu_replicated_rank.err
if [ "${SKIP_CFG5:-0}" != "1" ]; then echo "== cfg5 (1M x 100k)"; run --config cfg5 --steps 10 --warmup 3 > gpurun_out/r2_bench_cfg5_${N}gpu.json 2> gpurun_out/r2_bench_cfg5_${N}gpu.err; tail -c 2500 gpurun_out/r2_bench_cfg5_${N}gpu.json; tail -3 gpurun_out/r2_bench_cfg5_${N}gpu.err; fi
