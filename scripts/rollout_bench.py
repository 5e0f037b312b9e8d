"""cfg4-sized policy step of a rollout (N = 65 536 policies, Linear(376,256)-Tanh-Linear(256,17), one observation each):
the K8 kernel with the observation statistics update + normalisation + clipping + active-mask fused, against the reference's
per-step op sequence in torch eager (mask gather -> RunningNorm.update_and_normalize -> scatter -> vmap(functional_call))."""
import json
import sys

import torch
from torch import nn
from torch.func import functional_call, vmap

sys.path.insert(0, ".")
from evotorch_b200.neuroevolution import Policy, RunningNorm  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda")
net = nn.Sequential(nn.Linear(376, 256), nn.Tanh(), nn.Linear(256, 17))
policy = Policy(net)
L = policy.parameter_length
params = torch.randn(N, L, device=dev) * 0.1
obs = torch.randn(N, 376, device=dev) * 2 + 0.5
policy.set_parameters(params)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


rows = []
for frac in (1.0, 0.5, 0.1):
    active = torch.rand(N, device=dev) < frac
    rn = RunningNorm(shape=376, dtype="float32", device=dev, clip=(-10.0, 10.0))
    rn.update(obs, active)

    def fused():
        rn.update(obs, active)
        return policy(obs, obs_norm=rn, active=active)

    # the reference's sequence (vecgymne.py:822-839) on the same GPU, eager
    ref_sum, ref_sumsq, ref_count = torch.zeros(376, device=dev), torch.zeros(376, device=dev), 0
    names = [n for n, _ in net.named_parameters()]
    shapes = [p.shape for _, p in net.named_parameters()]
    net_dev = net.to(dev)

    def unflat(flat):
        out, o = {}, 0
        for n_, s_ in zip(names, shapes):
            k = s_.numel()
            out[n_] = flat[o:o + k].reshape(s_)
            o += k
        return out

    def reference():
        global ref_count
        sel = obs[active]                                   # boolean gather: host sync
        s1, s2 = sel.sum(0), sel.square().sum(0)
        n = int(active.sum())                               # host sync (runningnorm.py:318)
        mean = s1 / max(n, 1)
        std = torch.sqrt(torch.clamp(s2 / max(n, 1) - mean.square(), min=1e-2))
        normed = torch.clamp((sel - mean) / std, -10.0, 10.0)
        full = obs.clone()
        full[active] = normed
        return vmap(lambda f, x: functional_call(net_dev, unflat(f), (x,)))(params, full)

    t_fused, t_ref = timeit(fused), timeit(reference, reps=5)
    t_plain = timeit(lambda: policy(obs))
    rows.append({"N": N, "active_fraction": frac, "fused_step_ms": round(t_fused, 4), "plain_forward_ms": round(t_plain, 4),
                 "torch_reference_step_ms": round(t_ref, 4), "speedup_vs_torch": round(t_ref / t_fused, 2),
                 "param_bytes_read_gb": round(4.0 * L * float(active.sum()) / 1e9, 2),
                 "achieved_tbs": round(4.0 * L * float(active.sum()) / 1e12 / (t_fused * 1e-3), 3)})
    print(json.dumps(rows[-1]), flush=True)
