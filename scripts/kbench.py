"""Time the two HBM-bound kernels (fused sample+eval, grad) of every libevok build variant found in evotorch_b200/lib/.

    python scripts/kbench.py [popsize] [dim]      (on the GPU box)
Variants are built on the CPU box with scripts/build_variants.py.
"""
import ctypes
import glob
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from evotorch_b200 import _native as nat  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
REPS = 10
dev = torch.device("cuda", 0)
mu = torch.empty(D, device=dev).uniform_(-5.12, 5.12)
sg = torch.ones(D, device=dev)
X = torch.empty(N, D, device=dev)
f = torch.empty(N, device=dev)
w = torch.randn(N, device=dev) / N
gm, gs = torch.empty(D, device=dev), torch.empty(D, device=dev)
ws = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream().cuda_stream


def load(path):
    h = ctypes.CDLL(path)
    for name, (res, args) in nat._SIGNATURES.items():
        fn = getattr(h, name)
        fn.restype, fn.argtypes = res, args
    return h


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(REPS):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / REPS


results = {}
ref_g = None
libs = sorted(glob.glob(os.path.join(ROOT, "evotorch_b200", "lib", "libevok*.so")))
for path in libs:
    tag = os.path.basename(path)[len("libevok"):-3].lstrip("_") or "default"
    h = load(path)

    def sample(obj=2, store=True):
        rc = h.evok_sample_eval(obj, X.data_ptr() if store else None, D, mu.data_ptr(), sg.data_ptr(), 0, N, D, 1, 7, 1, None,
                                f.data_ptr() if obj else None, stream)
        assert rc == 0, rc

    def grad():
        rc = h.evok_grad(1, X.data_ptr(), D, w.data_ptr(), mu.data_ptr(), sg.data_ptr(), N, D, 1.0, 1.0, gm.data_ptr(), gs.data_ptr(),
                         ws.data_ptr(), ws.numel(), stream)
        assert rc == 0, rc

    t_se = timeit(sample)
    t_s = timeit(lambda: sample(0, True))
    t_lazy = timeit(lambda: sample(2, False))
    t_g = timeit(grad)
    t_rng = timeit(lambda: sample(1, False))       # sphere, no store: Philox + Box-Muller + one FFMA per element = the RNG floor
    t_sphere = timeit(lambda: sample(1, True))     # sphere with the store
    sample(2, True)
    grad()
    torch.cuda.synchronize()
    cur = torch.cat([gm, gs]).double()
    if ref_g is None:
        ref_g = cur.clone()
    gdiff = float((cur - ref_g).abs().max() / ref_g.abs().max())
    gb = 4.0 * N * D / 1e9
    results[tag] = {"sample_eval_ms": t_se, "sample_eval_gbs": gb / t_se * 1e3, "sample_only_ms": t_s, "sample_only_gbs": gb / t_s * 1e3,
                    "lazy_eval_ms": t_lazy, "rng_only_ms": t_rng, "sphere_store_ms": t_sphere, "grad_ms": t_g, "grad_gbs": 0.5 * gb / t_g * 1e3}
    print(f"{tag:28s} sample_eval {t_se:7.3f} ms {gb / t_se * 1e3:7.0f} GB/s | sample {t_s:7.3f} ms {gb / t_s * 1e3:7.0f} GB/s | "
          f"lazy {t_lazy:7.3f} ms | rng-only {t_rng:7.3f} ms | sphere+store {t_sphere:7.3f} ms | grad {t_g:7.3f} ms {0.5 * gb / t_g * 1e3:7.0f} GB/s (rel diff vs default {gdiff:.1e})", flush=True)
# K8 at BASELINE config 4: 65 536 policies x 100 881 parameters (26.4 GB), one observation each
del X, f, w
torch.cuda.empty_cache()
NP, dims, acts = 65536, [376, 256, 17], [1, 0]
L = 376 * 256 + 256 + 256 * 17 + 17
P = torch.empty(NP, L, device=dev).normal_(0, 0.1)
obs = torch.randn(NP, 376, device=dev)
out = torch.empty(NP, 17, device=dev)
h = load(os.path.join(ROOT, "evotorch_b200", "lib", "libevok.so"))
d_arr, a_arr = (ctypes.c_int32 * 3)(*dims), (ctypes.c_int32 * 2)(*acts)
t_mlp = timeit(lambda: h.evok_mlp_forward(P.data_ptr(), L, obs.data_ptr(), 376, out.data_ptr(), 17, NP, 2, d_arr, a_arr, stream))
print(f"mlp_forward cfg4 (65536 x 100881, B=1): {t_mlp:7.3f} ms  {4.0 * NP * L / 1e9 / t_mlp * 1e3:7.0f} GB/s", flush=True)
results["_mlp_cfg4"] = {"ms": t_mlp, "gbs": 4.0 * NP * L / 1e9 / t_mlp * 1e3}
ref_t = timeit(lambda: torch.bmm(P[:, :376 * 256].view(NP, 256, 376), obs.unsqueeze(-1)))
print(f"torch.bmm layer-1 only: {ref_t:7.3f} ms  {4.0 * NP * 376 * 256 / 1e9 / ref_t * 1e3:7.0f} GB/s")
del P, obs, out
torch.cuda.empty_cache()
X = torch.empty(N, D, device=dev)
# reference points: torch copy (read+write) and torch fill (write only)
Y = torch.empty_like(X[: N // 2])
t_copy = timeit(lambda: Y.copy_(X[: N // 2]))
t_fill = timeit(lambda: X.fill_(1.0))
print(f"torch copy {2 * 0.5 * 4.0 * N * D / 1e9 / t_copy * 1e3:7.0f} GB/s (read+write) | torch fill {4.0 * N * D / 1e9 / t_fill * 1e3:7.0f} GB/s (write only)")
results["_torch"] = {"copy_gbs": 4.0 * N * D / 1e9 / t_copy * 1e3, "fill_gbs": 4.0 * N * D / 1e9 / t_fill * 1e3}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"N": N, "D": D, "results": results}, open(os.path.join(ROOT, "gpurun_out", "kbench.json"), "w"), indent=1)
