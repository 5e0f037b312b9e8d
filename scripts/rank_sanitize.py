import sys
import torch
sys.path.insert(0, ".")
from evotorch_b200 import ops
for n in (1, 2, 5, 63, 64, 65, 1000, 1024, 1025, 2047, 4096, 4097, 8191, 8192):
    f = torch.randn(n, device="cuda").round(decimals=1)
    perm = torch.empty(n, dtype=torch.int64, device="cuda")
    for m in ("centered", "linear", "nes", "normalized", "raw"):
        ops.rank(f, m, True, perm=perm)
        ops.rank(f, m, False)
    ops.argsort(f, True)
    ops.elite_mask(f, n // 2)
torch.cuda.synchronize()
print("RANK_SANITIZE_DONE")
