#!/bin/bash
# The driver's scaling command at N GPUs: the DEFAULT bench line (replicated ranking + push kernel, e2e and sharded_parity legs included).
# usage: bash scripts/r2_gpu_scale.sh <N>
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/r2_scale_${N}gpu.json 2> gpurun_out/r2_scale_${N}gpu.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_scale_${N}gpu.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("n_gpus", "value", "ms_per_step", "scaling", "gpu_launches", "clocks")}, "e2e", (d.get("e2e") or {}).get("value"), "parity", (d.get("sharded_parity") or {}).get("ok"),
      "roofline", d["roofline"]["frac"], d["roofline"]["ms_per_launch"])
PY
tail -2 gpurun_out/r2_scale_${N}gpu.err
