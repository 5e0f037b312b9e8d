"""Timeline of CTA (0, 0, 0) of gemm_tf32x3_kernel (trace build, see scripts/gather_trace.py): per K-block stamps of the TMA / converter /
MMA roles, per chunk stamps of the epilogue, kernel start / store phase."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
import evotorch_b200._native as nat

nat.LIB_PATH = os.path.join(os.path.dirname(nat.LIB_PATH), "libevok_trace.so")
from evotorch_b200 import ops

dev = "cuda"
out = {}
for (M, N, K) in [(4096, 1024, 1024), (8192, 8192, 8192)]:
    A, B = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev)
    for _ in range(2):
        ops.gemm_nt(A, B)
    torch.cuda.synchronize()
    trace = torch.zeros(512, 16, dtype=torch.int64, device=dev)
    os.environ["EVOK_GATHER_TRACE_PTR"] = hex(trace.data_ptr())
    ops.gemm_nt(A, B)
    torch.cuda.synchronize()
    del os.environ["EVOK_GATHER_TRACE_PTR"]
    t = trace.cpu()
    t0 = int(t[0, 13])
    nkb = min((K + 31) // 32, 512)
    print(f"== {M} x {N} x {K}: kernel start 0, store phase starts {int(t[1, 13]) - t0}, stores issued {int(t[2, 13]) - t0} (cycles)")
    print(" i  tma_issue  conv_top +landed +arrived | mma_ready +committed | period")
    prev = None
    rows = []
    for i in list(range(0, min(nkb, 14))) + list(range(max(14, nkb - 6), nkb)):
        r = {k: int(t[i, s]) - t0 for k, s in (("tma", 9), ("ctop", 0), ("cland", 1), ("carr", 3), ("mrdy", 6), ("mcom", 8))}
        rows.append({"i": i, **r})
        per = r["mrdy"] - prev if prev is not None else 0
        prev = r["mrdy"]
        print(f'{i:3d} {r["tma"]:8d} {r["ctop"]:8d} +{r["cland"] - r["ctop"]:6d} +{r["carr"] - r["cland"]:6d} | {r["mrdy"]:8d} +{r["mcom"] - r["mrdy"]:6d} | {per:6d}')
    nch = (nkb + 3) // 4
    ch = [{"chunk": c, "wait": int(t[c, 10]) - t0, "ready": int(t[c, 11]) - t0} for c in range(min(nch, 128))]
    print(" chunks (wait_start, ready):", [(c["chunk"], c["wait"], c["ready"]) for c in ch[:6]], "...", [(c["chunk"], c["wait"], c["ready"]) for c in ch[-3:]])
    out[f"{M}x{N}x{K}"] = {"blocks": rows, "chunks": ch, "store_phase_start": int(t[1, 13]) - t0, "stores_issued": int(t[2, 13]) - t0}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gemm_trace.json", "w"))
