import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evotorch_b200 import ops
torch.manual_seed(0)
dev = "cuda"
for (M, N, K) in [(4096, 1024, 1024), (1024, 1024, 4096), (128, 256, 1024), (512, 512, 1024)]:
    # exact integer data: any mismatch is a data-path bug, not rounding
    A = torch.randint(-8, 9, (M, K), device=dev).float()
    B = torch.randint(-8, 9, (N, K), device=dev).float()
    ref = A.double() @ B.double().T
    for rep in range(3):
        C = ops.gemm_nt(A, B)
        torch.cuda.synchronize()
        bad = (C.double() != ref)
        print(f"int data M={M} N={N} K={K} rep {rep}: mismatches {int(bad.sum())} of {M*N}", end="")
        if bad.any():
            idx = bad.nonzero()
            print(" first bad", idx[:5].tolist(), "rows uniq", idx[:, 0].unique()[:10].tolist(), "cols uniq", idx[:, 1].unique()[:10].tolist(), end="")
        print(flush=True)
    A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev)
    ref = A.double() @ B.double().T
    C = ops.gemm_nt(A, B)
    err = (C.double() - ref)
    rel = err.abs() / ref.abs().max()
    print(f"randn M={M} N={N} K={K}: max {float(rel.max()):.2e} mean {float(rel.mean()):.2e} p99.9 {float(rel.flatten().kthvalue(int(0.999*rel.numel())).values):.2e}"
          f" | signed mean err/|C|: {float((err*ref.sign()).mean()/ref.abs().mean()):.2e}")
    # per-tile max error map
    tm = rel.reshape(M // 128 if M >= 128 else 1, -1, N // 256 if N >= 256 else 1, 256 if N >= 256 else N).amax(dim=(1, 3)) if M % 128 == 0 and N % 256 == 0 else None
    if tm is not None:
        print("   per-tile max err: min %.2e max %.2e" % (float(tm.min()), float(tm.max())))
    # hi-only (single-pass TF32) reference for comparison
    Ah = (A.view(torch.int32) & -8192).view(torch.float32); Bh = (B.view(torch.int32) & -8192).view(torch.float32)
    r1 = Ah.double() @ Bh.double().T
    print("   single-pass TF32 ideal err %.2e" % float(((r1 - ref).abs() / ref.abs().max()).max()))
