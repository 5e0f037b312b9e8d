"""Golden fitnesses from the REAL reference's neuro-evolution problems (`evotorch.neuroevolution.NEProblem`, neproblem.py:33-429,
and `SupervisedNE`, supervisedne.py:31-348): a fixed population evaluated one solution at a time by the reference's own
`_evaluate` loop.  The package must reproduce them through its batched route.

    PYTHONPATH=tests/golden/_refstubs:/root/reference/src EVOTORCH_VERBOSE_LEVEL=0 python tests/golden/gen_ne_golden.py
"""

import os

import numpy as np
import torch
from torch import nn
from torch.utils.data import TensorDataset

import evotorch
from evotorch import SolutionBatch
from evotorch.neuroevolution import NEProblem, SupervisedNE

assert "/root/reference" in evotorch.__file__
HERE = os.path.dirname(os.path.abspath(__file__))
out = {}
g = torch.Generator().manual_seed(11)

# ---- SupervisedNE, common minibatch = the whole dataset (so that the DataLoader's shuffling cannot change the loss)
for tag, (n_in, n_hid, n_out, act, B, N) in {"mlp_tanh": (6, 16, 3, nn.Tanh, 40, 24), "mlp_relu_wide": (20, 128, 5, nn.ReLU, 64, 10)}.items():
    X = torch.randn(B, n_in, generator=g)
    Y = torch.randn(B, n_out, generator=g)

    def make_net(n_in=n_in, n_hid=n_hid, n_out=n_out, act=act):
        return nn.Sequential(nn.Linear(n_in, n_hid), act(), nn.Linear(n_hid, n_out))

    prob = SupervisedNE(TensorDataset(X, Y), make_net, nn.MSELoss(), minibatch_size=B, num_minibatches=2, common_minibatch=True)
    P = torch.randn(N, prob.solution_length, generator=g) * 0.5
    batch = SolutionBatch(prob, N)
    batch.set_values(P)
    prob.evaluate(batch)
    out[f"supervised/{tag}/X"], out[f"supervised/{tag}/Y"], out[f"supervised/{tag}/P"] = X.numpy(), Y.numpy(), P.numpy()
    out[f"supervised/{tag}/loss"] = batch.evals[:, 0].numpy().copy()
    out[f"supervised/{tag}/dims"] = np.array([n_in, n_hid, n_out])
    # the same with an L1 loss handed in as a plain function
    prob = SupervisedNE(TensorDataset(X, Y), make_net, lambda yh, y: torch.mean(torch.abs(yh - y)), minibatch_size=B, common_minibatch=True)
    batch = SolutionBatch(prob, N)
    batch.set_values(P)
    prob.evaluate(batch)
    out[f"supervised/{tag}/l1"] = batch.evals[:, 0].numpy().copy()

# ---- NEProblem with a deterministic network evaluator (+ evaluation data)
probe = torch.randn(9, 5, generator=g)


def evaluator(net):
    y = net(probe)
    return torch.sum(y * y), torch.stack([y.mean(), y.max()])


prob = NEProblem("max", nn.Sequential(nn.Linear(5, 7), nn.Tanh(), nn.Linear(7, 2)), evaluator, eval_data_length=2)
P = torch.randn(12, prob.solution_length, generator=g)
batch = SolutionBatch(prob, 12)
batch.set_values(P)
prob.evaluate(batch)
out["neproblem/probe"], out["neproblem/P"], out["neproblem/evals"] = probe.numpy(), P.numpy(), batch.evals.numpy().copy()
net = prob.make_net(P[3])
out["neproblem/make_net_out"] = net(probe).detach().numpy()
out["neproblem/solution_length"] = np.array([prob.solution_length])

np.savez_compressed(os.path.join(HERE, "ne_golden.npz"), **out)
print("wrote", len(out), "arrays")
