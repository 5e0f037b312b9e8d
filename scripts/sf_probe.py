import sys, torch
sys.path.insert(0, '.')
from evotorch_b200.neuroevolution import Policy
dev = torch.device("cuda", 0)
pol = Policy(torch.nn.Sequential(torch.nn.Linear(376, 256), torch.nn.Tanh(), torch.nn.Linear(256, 17)))
P = torch.empty(4096, pol.parameter_length, device=dev).normal_(0, 0.1)
x = torch.randn(256, 376, device=dev)
for _ in range(3):
    y = pol.forward_shared(P, x)
torch.cuda.synchronize()
