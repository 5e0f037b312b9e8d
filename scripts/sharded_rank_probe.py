"""Time evok_rank_sharded with 8 simulated ranks (125 k keys each) on ONE GPU: per-rank kernel cost without NVLink."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from evotorch_b200 import _native as nat, ops
lib = nat.lib()
DEV = "cuda"
R, nl = 8, 125_000
n = R * nl
f = torch.randn(n, device=DEV)
offs = [r * nl for r in range(R + 1)]
keys = [torch.zeros(n, dtype=torch.int32, device=DEV) for _ in range(R)]
fsum = [torch.zeros(R, dtype=torch.float64, device=DEV) for _ in range(R)]
flags = [torch.zeros(R, dtype=torch.int64, device=DEV) for _ in range(R)]
epoch = [torch.zeros(1, dtype=torch.int64, device=DEV) for _ in range(R)]
done = [torch.zeros(4, dtype=torch.int32, device=DEV) for _ in range(R)]
err = torch.zeros(1, dtype=torch.int32, device=DEV)
w = [torch.empty(nl, device=DEV) for _ in range(R)]
mean = [torch.zeros(1, device=DEV) for _ in range(R)]
ws = [torch.empty(lib.evok_rank_workspace_bytes(nl), dtype=torch.uint8, device=DEV) for _ in range(R)]
tab = lambda ts: (ctypes.c_void_p * R)(*[t.data_ptr() for t in ts])
c_offs = (ctypes.c_int64 * (R + 1))(*offs)
st = torch.cuda.current_stream().cuda_stream
def one_round():
    # all ranks on ONE stream, in rank order: rank r's merge would wait for the later ranks' pushes, so first run every rank's
    # sort + push with a private copy of the call where merge waits are satisfied: do two passes (push of all, then merges see flags)
    for r in range(R):
        rc = lib.evok_rank_sharded(0, f[offs[r]:offs[r + 1]].data_ptr(), n, 0, R, r, c_offs, tab(keys), tab(fsum), tab(flags), epoch[r].data_ptr(),
                                   done[r].data_ptr(), err.data_ptr(), int(2e6), w[r].data_ptr(), mean[r].data_ptr(), ws[r].data_ptr(), ws[r].numel(), st)
        assert rc == 0
# single stream: rank 0's merge times out (2 ms) waiting for ranks 1..7 -> instead time the pieces separately
import time
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
wl = torch.empty(nl, device=DEV)
print("local rank of 125k keys (sort + scatter)  us:", t(lambda: ops.rank(f[:nl], "centered", False, out=wl)))
# world = 1 sharded call on 125 k keys: sort + push (local only) + merge (no searches)
k1 = [torch.zeros(nl, dtype=torch.int32, device=DEV)]; fs1 = [torch.zeros(1, dtype=torch.float64, device=DEV)]; fl1 = [torch.zeros(1, dtype=torch.int64, device=DEV)]
e1 = torch.zeros(1, dtype=torch.int64, device=DEV); d1 = torch.zeros(4, dtype=torch.int32, device=DEV)
t1 = lambda ts: (ctypes.c_void_p * 1)(*[x.data_ptr() for x in ts])
o1 = (ctypes.c_int64 * 2)(0, nl)
print("sharded call, world 1 (sort+push+merge, no searches) us:", t(lambda: lib.evok_rank_sharded(0, f.data_ptr(), nl, 0, 1, 0, o1, t1(k1), t1(fs1), t1(fl1), e1.data_ptr(), d1.data_ptr(), err.data_ptr(), int(2e6), wl.data_ptr(), mean[0].data_ptr(), ws[0].data_ptr(), ws[0].numel(), st)))
# 8 simulated ranks on 8 streams, total wall time / 8
streams = [torch.cuda.Stream() for _ in range(R)]
def all_ranks():
    for r in range(R):
        with torch.cuda.stream(streams[r]):
            lib.evok_rank_sharded(0, f[offs[r]:offs[r + 1]].data_ptr(), n, 0, R, r, c_offs, tab(keys), tab(fsum), tab(flags), epoch[r].data_ptr(),
                                  done[r].data_ptr(), err.data_ptr(), int(5e8), w[r].data_ptr(), mean[r].data_ptr(), ws[r].data_ptr(), ws[r].numel(), streams[r].cuda_stream)
for _ in range(3): all_ranks()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): all_ranks()
torch.cuda.synchronize()
print("8 simulated ranks on one GPU, per round us:", (time.perf_counter() - t0) / 20 * 1e6, "err", int(err.item()))
