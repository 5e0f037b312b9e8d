"""Timeline of one CTA of gemm_gather_persistent_kernel: clock64() stamps of the converter / MMA / TMA / epilogue roles per K-block.
Needs the trace build:  python -c "from evotorch_b200.build import build; build(defines=('EVOK_GEMM_TRACE',), tag='trace')"
"""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
import evotorch_b200._native as nat

nat.LIB_PATH = os.path.join(os.path.dirname(nat.LIB_PATH), "libevok_trace.so")
from evotorch_b200.neuroevolution import Policy

dev = torch.device("cuda", 0)
pol = Policy(torch.nn.Sequential(torch.nn.Linear(376, 256), torch.nn.Tanh(), torch.nn.Linear(256, 17)).to(dev))
P = torch.empty(4096, pol.parameter_length, device=dev).normal_(0, 0.1)
x = torch.randn(256, 376, device=dev)
for _ in range(2):
    pol.forward_shared(P, x)
torch.cuda.synchronize()
trace = torch.zeros(512, 16, dtype=torch.int64, device=dev)
os.environ["EVOK_GATHER_TRACE_PTR"] = hex(trace.data_ptr())
pol.forward_shared(P, x)
torch.cuda.synchronize()
del os.environ["EVOK_GATHER_TRACE_PTR"]
t = trace.cpu()
names = {0: "conv_top", 14: "conv_data_landed", 1: "conv_after_barsync", 2: "conv_after_empty_lo", 3: "conv_arrived", 4: "gather(g)_stage_free", 5: "conv_iter_end",
         6: "mma_conv_a_ok", 7: "mma_full_b_ok", 8: "mma_committed", 9: "tma_b_issue", 10: "epi_wait_chunk", 11: "epi_chunk_ready", 12: "epi_chunk_folded"}
base = int(t[120, 0])
rows = []
for g in range(120, 160):
    rows.append({"g": g, **{names[s]: int(t[g, s]) - base for s in sorted(names) if s < 10 or s == 14}})
out = {"clock": "SM cycles relative to conv_top of K-block 120 (CTA 0); 12 K-blocks per tile", "blocks": rows,
       "chunks": [{"chunk": c, **{names[s]: int(t[c, s]) - base for s in (10, 11, 12)}} for c in range(30, 42)]}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gather_trace.json", "w"))
# compact text view: per K-block deltas
print("g  conv: top->landed->barsync->empty_lo->arrived->iter_end | stage_free(g) | mma: conv_a_ok->full_b_ok->committed | per-block period")
prev = None
for r in rows:
    per = r["conv_top"] - prev if prev is not None else 0
    prev = r["conv_top"]
    print(f'{r["g"]:3d} {r["conv_top"]:7d} +{r["conv_data_landed"] - r["conv_top"]:5d} +{r["conv_after_barsync"] - r["conv_data_landed"]:5d} '
          f'+{r["conv_after_empty_lo"] - r["conv_after_barsync"]:5d} +{r["conv_arrived"] - r["conv_after_empty_lo"]:5d} +{r["conv_iter_end"] - r["conv_arrived"]:5d} | '
          f'{r["gather(g)_stage_free"]:7d} | {r["mma_conv_a_ok"]:7d} +{r["mma_full_b_ok"] - r["mma_conv_a_ok"]:5d} +{r["mma_committed"] - r["mma_full_b_ok"]:5d} | {per:6d}')
