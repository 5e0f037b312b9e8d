#!/bin/bash
# per-kernel-group timings of the sharded generation in EAGER mode (CUDA events around each group), both ranking protocols
set -u
N=${1:-8}
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $N "$@"; }
for v in 1 0; do
  EVOTORCH_B200_SHARDED_RANK=$v run --steps 40 --warmup 5 --cuda-graph 0 --no-e2e --no-sharded-parity > gpurun_out/r2_diag_${N}gpu_sharded${v}.json 2> gpurun_out/r2_diag_${N}gpu_sharded${v}.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_diag_${N}gpu_sharded${v}.json')); print('sharded_rank=$v', d['ms_per_step'], {k: round(x['ms'],4) for k,x in d['kernels'].items()})"
done
