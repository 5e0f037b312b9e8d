"""Run under torchrun (one rank per GPU): the sharded PGPE/SNES/CEM generations must reproduce the single-GPU run with the
same seed (the Philox population is shard-invariant; ranking is global), up to fp32 summation order.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_multi_gpu.py
"""
import gc
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evotorch_b200 import Problem  # noqa: E402
from evotorch_b200.algorithms import CEM, PGPE, SNES  # noqa: E402
from evotorch_b200.objectives import rastrigin  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def make(kind, distributed):
    prob = Problem("min", rastrigin, initial_bounds=(-5.12, 5.12), solution_length=1000, device=dev, seed=17)
    if kind == "pgpe":
        return PGPE(prob, popsize=20000, center_learning_rate=0.5, stdev_learning_rate=0.1, stdev_init=1.0, distributed=distributed)
    if kind == "snes":
        return SNES(prob, popsize=20000, stdev_init=2.0, distributed=distributed)
    return CEM(prob, popsize=20000, parenthood_ratio=0.2, stdev_init=2.0, distributed=distributed)


ok = True
for kind in ("pgpe", "snes", "cem"):
    sharded = make(kind, True)
    assert sharded._distributed
    single = make(kind, False)
    # the sharded searcher updates from generation 1 on; the single-process searcher's first step only samples
    single.step()
    for gen in range(6):
        sharded.step()
        single.step()
        dmu = float((sharded.status["center"] - single.status["center"]).abs().max() / single.status["center"].abs().max())
        dsg = float((sharded.status["stdev"] - single.status["stdev"]).abs().max() / single.status["stdev"].abs().max())
        good = dmu < 1e-5 and dsg < 1e-5
        ok = ok and good
        if rank == 0:
            print(f"{kind} gen {gen}: rel diff mu {dmu:.2e} sigma {dsg:.2e} mean_eval {sharded.status['mean_eval']:.3f} {'OK' if good else 'MISMATCH'}", flush=True)
    # every rank holds the same replicated distribution
    ref = sharded.status["center"].clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(ref, sharded.status["center"]), "ranks diverged"
# CUDA-graph replay of the sharded generation (collectives captured) must equal the eager sharded run bit for bit.
# The two searchers run one after the other (never interleaved: replayed collectives use the replaying stream).
for kind in ("pgpe", "snes"):
    traj = {}
    for mode in ("eager", "graph"):
        s_ = make(kind, True)
        if mode == "graph":
            s_.enable_cuda_graph()
        steps = []
        for gen in range(8):
            s_.step()
            steps.append(torch.cat([s_.status["center"], s_.status["stdev"]]).clone())
        traj[mode] = (torch.stack(steps), s_._graph is not None, s_.status["mean_eval"])
        torch.cuda.synchronize()
        s_._graph = None  # release the captured graph (and the NCCL work it references) before anything else uses the communicator
        del s_
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
    good = torch.equal(traj["eager"][0], traj["graph"][0]) and traj["graph"][1] and not traj["eager"][1]
    ok = ok and good
    if rank == 0:
        print(f"{kind}: graph-replayed sharded run == eager sharded run: {'OK' if good else 'MISMATCH'} (graph captured: {traj['graph'][1]})", flush=True)
dist.barrier()
if rank == 0:
    print("MULTI_GPU_PARITY", "PASS" if ok else "FAIL", "world", world, flush=True)
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0 if ok else 1)  # skip the (slow, occasionally hanging) communicator teardown: the process is done
