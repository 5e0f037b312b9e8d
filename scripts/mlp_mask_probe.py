"""A few masked, normalising policy forwards (N = 16 384 policies of cfg4's network, half of them active) -- the target of an
ncu capture: DRAM bytes read should be about the ACTIVE parameter rows only."""
import sys

import torch
from torch import nn

sys.path.insert(0, ".")
from evotorch_b200.neuroevolution import Policy, RunningNorm  # noqa: E402

N, dev = 16384, torch.device("cuda")
policy = Policy(nn.Sequential(nn.Linear(376, 256), nn.Tanh(), nn.Linear(256, 17)))
params = torch.randn(N, policy.parameter_length, device=dev) * 0.1
obs = torch.randn(N, 376, device=dev)
torch.manual_seed(0)
active = torch.rand(N, device=dev) < 0.5
rn = RunningNorm(shape=376, dtype="float32", device=dev, clip=(-10.0, 10.0))
rn.update(obs, active)
policy.set_parameters(params)
for _ in range(4):
    out = policy(obs, obs_norm=rn, active=active)
torch.cuda.synchronize()
print("active", int(active.sum()), "of", N, "-> parameter bytes of active rows", 4 * policy.parameter_length * int(active.sum()))
