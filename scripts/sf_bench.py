"""Shared-minibatch policy forward (cfg4, B = 256): time of the whole forward and of its two kernels, and torch vmap beside it."""
import json
import sys

import torch

sys.path.insert(0, ".")
import os

import evotorch_b200._native as _nat

if os.environ.get("SF_LIB"):  # A/B against a variant build (evotorch_b200/lib/libevok_<tag>.so)
    _nat.LIB_PATH = os.environ["SF_LIB"]
from evotorch_b200 import ops
from evotorch_b200.neuroevolution import Policy

dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
pol = Policy(torch.nn.Sequential(torch.nn.Linear(376, 256), torch.nn.Tanh(), torch.nn.Linear(256, 17)).to(dev))
P = torch.empty(n, pol.parameter_length, device=dev).normal_(0, 0.1)
x = torch.randn(256, 376, device=dev)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for _ in range(2):
    pol.forward_shared(P, x)
ms = timed(lambda: pol.forward_shared(P, x), 5)
ops.enable_timers()
for _ in range(3):
    pol.forward_shared(P, x)
torch.cuda.synchronize()
parts = ops.timer_results()
out = {"n": n, "B": 256, "ms_per_forward": ms, "parts_ms": parts, "useful_tflops": 2.0 * n * 256 * (376 * 256 + 256 * 17) / ms / 1e9}
if n <= 16384:
    f = lambda: torch.vmap(pol._call_one, in_dims=(0, None), chunk_size=4096)(P, x)
    out["torch_vmap_ms"] = timed(f, 3)
print(json.dumps(out))
