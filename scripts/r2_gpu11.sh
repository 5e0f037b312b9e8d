#!/bin/bash
set -u
mkdir -p gpurun_out
cat > gpurun_out/sf.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from evotorch_b200.neuroevolution import Policy
dev = torch.device("cuda", 0)
pol = Policy(torch.nn.Sequential(torch.nn.Linear(376, 256), torch.nn.Tanh(), torch.nn.Linear(256, 17)))
P = torch.empty(4096, pol.parameter_length, device=dev).normal_(0, 0.1)
x = torch.randn(256, 376, device=dev)
for _ in range(3):
    y = pol.forward_shared(P, x)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 2 -c 1 -f -o gpurun_out/prof_gemm_gather python gpurun_out/sf.py > gpurun_out/ncu_gg.log 2>&1; tail -3 gpurun_out/ncu_gg.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_sf.csv python gpurun_out/sf.py > /dev/null 2>&1; tail -8 gpurun_out/launches_sf.csv
