"""Time evok_rank (centered) over population sizes around the counting / radix switch; CUDA events, eager launches and graph replay."""
import sys

import torch

sys.path.insert(0, ".")
from evotorch_b200 import ops  # noqa: E402

dev = "cuda"
for n in (32, 1000, 4096, 8192, 8193, 20000, 100000, 125000, 250000, 500000, 1000000):
    f = torch.randn(n, device=dev)
    w = torch.empty(n, device=dev)
    for _ in range(5):
        ops.rank(f, "centered", False, out=w)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ops.rank(f, "centered", False, out=w)
        s.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(20):
                ops.rank(f, "centered", False, out=w)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    t_graph = a.elapsed_time(b) / 200 * 1e3
    a.record()
    for _ in range(200):
        ops.rank(f, "centered", False, out=w)
    b.record()
    torch.cuda.synchronize()
    t_eager = a.elapsed_time(b) / 200 * 1e3
    print(f"N = {n:8d}: rank {t_graph:8.2f} us (graph replay)  {t_eager:8.2f} us (eager launches)", flush=True)
