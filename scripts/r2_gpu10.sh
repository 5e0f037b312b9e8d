#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest sharded rank / shared forward / NE"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_neproblem.py -q --maxfail=10 -k "sharded or shared or neproblem or supervised" 2>&1 | tail -25
echo "== shared forward timing"; timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/r2_shared_forward.txt
import json, torch
from evotorch_b200.neuroevolution import Policy
dev = torch.device("cuda", 0)
net = torch.nn.Sequential(torch.nn.Linear(376, 256), torch.nn.Tanh(), torch.nn.Linear(256, 17))
pol = Policy(net)
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for NP, B in ((65536, 256), (65536, 64), (65536, 16), (8192, 256)):
    P = torch.empty(NP, pol.parameter_length, device=dev).normal_(0, 0.1)
    x = torch.randn(B, 376, device=dev)
    ms = t(lambda: pol.forward_shared(P, x))
    useful = 2.0 * NP * B * (376 * 256 + 256 * 17)
    rec = {"policies": NP, "B": B, "ms": ms, "fp32_equiv_tflops": useful / ms / 1e9, "tensor_tflops_3xtf32": 3 * 2.0 * NP * B * 376 * 256 / ms / 1e9,
           "param_gbs": 4.0 * NP * pol.parameter_length / ms / 1e6}
    if NP <= 8192:
        ref = t(lambda: torch.vmap(pol._call_one, in_dims=(0, None))(P, x), 3)
        rec["torch_vmap_ms"] = ref
    print(json.dumps(rec), flush=True)
    del P
    torch.cuda.empty_cache()
PY
