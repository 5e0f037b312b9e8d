"""Per-generation timing of the sharded graph-mode PGPE at the metric shape (run under torchrun): prints, for rank 0, the
per-step device times with and without the nvidia-smi clock sampler of bench.py running."""
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

os.environ.setdefault("NCCL_DEBUG", "WARN")
sys.path.insert(0, ".")
from evotorch_b200 import Problem  # noqa: E402
from evotorch_b200.algorithms import PGPE  # noqa: E402
from evotorch_b200.objectives import rastrigin  # noqa: E402
from evotorch_b200.peer import enable_peer_exchange  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N, D = 1_000_000, 10_000
prob = Problem("min", rastrigin, initial_bounds=(-5.12, 5.12), solution_length=D, device=dev, seed=0)
px = enable_peer_exchange(prob, N)
s = PGPE(prob, popsize=N, center_learning_rate=0.5, stdev_learning_rate=0.1, stdev_init=1.0, distributed=True).enable_cuda_graph()
for _ in range(5):
    s.step()


def timed(steps, label, sampler):
    proc = None
    dist.barrier()
    torch.cuda.synchronize()
    if sampler and rank == 0:
        proc = subprocess.Popen(["nvidia-smi", "--query-gpu=index,clocks.sm", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(local)],
                                stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    t0 = time.perf_counter()
    evs[0].record()
    host = []
    for i in range(steps):
        h0 = time.perf_counter()
        s.step()
        host.append((time.perf_counter() - h0) * 1e3)
        evs[i + 1].record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    if proc is not None:
        proc.terminate()
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    tot = torch.tensor([evs[0].elapsed_time(evs[-1])], device=dev)
    dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"label": label, "ms_per_step_max_over_ranks": round(float(tot) / steps, 3), "wall_ms": round(wall, 1),
                          "device_ms": [round(x, 2) for x in per], "host_ms": [round(x, 2) for x in host]}), flush=True)


timed(30, "no sampler", False)
timed(30, "with nvidia-smi sampler", True)
timed(30, "no sampler again", False)
assert not px.timed_out()
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
