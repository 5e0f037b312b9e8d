"""Run under torchrun (one rank per GPU).  The peer-exchange generation (fitness gather and gradient reduction done by the
producing kernels over NVLink, no NCCL in the loop) must reproduce the NCCL-sharded and the single-GPU trajectories, be
bit-identical under CUDA-graph replay, and is timed against the NCCL path.

    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/check_peer_exchange.py
"""
import gc
import json
import os
import sys

import torch
import torch.distributed as dist

os.environ.setdefault("EVOTORCH_B200_PEER_TIMEOUT_S", "5")
os.environ.setdefault("NCCL_DEBUG", "WARN")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evotorch_b200 import Problem  # noqa: E402
from evotorch_b200.algorithms import CEM, PGPE, SNES  # noqa: E402
from evotorch_b200.objectives import rastrigin  # noqa: E402
from evotorch_b200.peer import enable_peer_exchange  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
exchanges = []


def say(*a):
    if rank == 0:
        print(*a, flush=True)


def make(kind, *, distributed=True, peer=False, lazy=False, graph=False, popsize=20000, dim=1000):
    prob = Problem("min", rastrigin, initial_bounds=(-5.12, 5.12), solution_length=dim, device=dev, seed=17, lazy_population=lazy)
    if peer:
        exchanges.append(enable_peer_exchange(prob, popsize))
    if kind == "pgpe":
        s = PGPE(prob, popsize=popsize, center_learning_rate=0.5, stdev_learning_rate=0.1, stdev_init=1.0, distributed=distributed)
    elif kind == "snes":
        s = SNES(prob, popsize=popsize, stdev_init=2.0, distributed=distributed)
    else:
        s = CEM(prob, popsize=popsize, parenthood_ratio=0.2, stdev_init=2.0, distributed=distributed)
    return s.enable_cuda_graph() if graph else s


def trajectory(s, gens):
    out = []
    for _ in range(gens):
        s.step()
        out.append(torch.cat([s.status["center"], s.status["stdev"]]).clone())
    torch.cuda.synchronize()
    return torch.stack(out)


def reldiff(a, b):
    return float((a - b).abs().max() / b.abs().max())


ok = True
TIMING_ONLY = os.environ.get("PEER_TIMING_ONLY", "0") == "1"
for kind in (() if TIMING_ONLY else ("pgpe", "snes", "cem")):
    # NCCL first, alone (its collectives must not interleave with anything else), then the peer variants
    t_nccl = trajectory(make(kind), 6)
    dist.barrier()
    t_peer = trajectory(make(kind, peer=True), 6)
    t_peer_graph_s = make(kind, peer=True, graph=True)
    t_peer_graph = trajectory(t_peer_graph_s, 6)
    t_peer_lazy = trajectory(make(kind, peer=True, lazy=True), 6)
    single = make(kind, distributed=False)
    single.step()
    t_single = trajectory(single, 6)
    d1, d2, d3 = reldiff(t_peer, t_nccl), reldiff(t_peer, t_single), reldiff(t_peer_lazy, t_peer)
    same = torch.equal(t_peer, t_peer_graph) and t_peer_graph_s._graph is not None
    # every rank must hold bit-identical parameters (fixed reduction order)
    ref = t_peer[-1].clone()
    dist.broadcast(ref, src=0)
    replicated = torch.equal(ref, t_peer[-1])
    good = d1 < 1e-5 and d2 < 1e-5 and d3 < 2e-4 and same and replicated
    ok = ok and good
    say(f"{kind}: peer vs nccl {d1:.2e}  peer vs 1-GPU {d2:.2e}  lazy-peer vs peer {d3:.2e}  graph==eager {same}  ranks identical {replicated}"
        f"  {'OK' if good else 'MISMATCH'}")

timeouts = [px.timed_out() for px in exchanges]
ok = ok and not any(timeouts)
say("wait timeouts:", sum(timeouts))


# ---- timing: one generation, CUDA-graph replay, NCCL collectives vs peer exchange
def time_generation(popsize, dim, peer, steps=50):
    s = make("pgpe", peer=peer, graph=True, popsize=popsize, dim=dim)
    for _ in range(6):
        s.step()
    assert s._graph is not None
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        s.step()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / steps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    s._graph = None
    del s
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    return float(t)


rows = []
if TIMING_ONLY:  # the metric shape only, three interleaved rounds, every measurement printed
    for rnd in range(3):
        say(json.dumps({"round": rnd, "nccl_ms": round(time_generation(1000000, 10000, False), 4),
                        "peer_ms": round(time_generation(1000000, 10000, True), 4)}))
for popsize, dim in (() if TIMING_ONLY else ((10000, 1000), (100000, 1000), (100000, 10000), (1000000, 10000))):
    # A/B/A/B, best of two each (the big shapes run at the power cap: single measurements wander by a few per cent)
    nccl_ms = time_generation(popsize, dim, False)
    peer_ms = time_generation(popsize, dim, True)
    nccl_ms = min(nccl_ms, time_generation(popsize, dim, False))
    peer_ms = min(peer_ms, time_generation(popsize, dim, True))
    rows.append({"popsize": popsize, "dim": dim, "world": world, "nccl_ms": round(nccl_ms, 4), "peer_ms": round(peer_ms, 4),
                 "speedup": round(nccl_ms / peer_ms, 3)})
    say(json.dumps(rows[-1]))
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/peer_exchange_{world}gpu.json", "w") as fh:
        json.dump({"parity": "PASS" if ok else "FAIL", "timing": rows}, fh, indent=1)
say("PEER_EXCHANGE", "PASS" if ok else "FAIL", "world", world)
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0 if ok else 1)
