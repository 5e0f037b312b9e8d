"""pgpe_tell / cem_tell on a batch of independent searches -- one launch per stage for all items vs one launch chain
per item (EVOTORCH_B200_FUNCTIONAL_LOOP=1).  python scripts/functional_bench.py [items] [popsize] [dim]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evotorch_b200.algorithms.functional import cem, cem_ask, cem_tell, pgpe, pgpe_ask, pgpe_tell  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
D = int(sys.argv[3]) if len(sys.argv) > 3 else 100
dev = "cuda"
torch.manual_seed(0)
center = torch.randn(B, D, device=dev)


def f(x):
    return torch.sum(x * x, dim=-1)


def run(kind, loop, iters=30):
    os.environ["EVOTORCH_B200_FUNCTIONAL_LOOP"] = "1" if loop else "0"
    if kind == "pgpe":
        st = pgpe(center_init=center, center_learning_rate=0.1, stdev_learning_rate=0.1, objective_sense="min", stdev_init=1.0)
        ask, tell = (lambda s: pgpe_ask(s, popsize=N)), pgpe_tell
    else:
        st = cem(center_init=center, parenthood_ratio=0.25, objective_sense="min", stdev_init=1.0)
        ask, tell = (lambda s: cem_ask(s, popsize=N)), cem_tell
    t_ask = t_tell = 0.0
    for i in range(iters + 3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        x = ask(st)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        ev = f(x)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        st = tell(st, x, ev)
        torch.cuda.synchronize(); t3 = time.perf_counter()
        if i >= 3:
            t_ask += t1 - t0
            t_tell += t3 - t2
    return 1e3 * t_ask / iters, 1e3 * t_tell / iters, float(f(st.optimizer_state.center if kind == "pgpe" else st.center).mean())


out = {"items": B, "popsize": N, "dim": D}
for kind in ("pgpe", "cem"):
    a1, t1, m1 = run(kind, loop=True)
    a2, t2, m2 = run(kind, loop=False)
    out[kind] = {"loop_ask_ms": a1, "loop_tell_ms": t1, "batched_ask_ms": a2, "batched_tell_ms": t2, "tell_speedup": t1 / t2, "ask_speedup": a1 / a2,
                 "final_mean_f_loop": m1, "final_mean_f_batched": m2}
print(json.dumps(out))
