#!/bin/bash
set -u
mkdir -p gpurun_out
cat > gpurun_out/sf2.py <<'PY'
import sys, os, json, torch
sys.path.insert(0, '.')
from evotorch_b200 import ops
from evotorch_b200.neuroevolution import Policy
dev = torch.device("cuda", 0)
pol = Policy(torch.nn.Sequential(torch.nn.Linear(376, 256), torch.nn.Tanh(), torch.nn.Linear(256, 17)))
P = torch.empty(4096, pol.parameter_length, device=dev).normal_(0, 0.1)
x = torch.randn(256, 376, device=dev)
ops.enable_timers()
for _ in range(6):
    y = pol.forward_shared(P, x)
torch.cuda.synchronize()
print(os.environ.get("EVOK_GATHER_DEBUG", "0"), ops.timer_results())
PY
for d in 0 1 2; do EVOK_GATHER_DEBUG=$d python gpurun_out/sf2.py; done
