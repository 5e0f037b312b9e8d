"""Generate the golden fixtures in this directory by running the REAL reference.

Run (in the build container, where /root/reference exists):

    PYTHONPATH=tests/golden/_refstubs:/root/reference/src EVOTORCH_VERBOSE_LEVEL=0 \
        python tests/golden/gen_golden.py

The reference (nnaisense/evotorch @ cebcac4f) is pure Python on torch; `ray` and `gymnasium` are not
installed here, so the two import stubs under `_refstubs/` stand in for them (import-time only; no
actor / gym code path is exercised).  Everything is computed on CPU in float32.  The outputs are small
`.npz` files that travel to the GPU box, where the reference does not exist.
"""

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))

import evotorch  # noqa: E402  (the reference)
from evotorch import Problem  # noqa: E402
from evotorch.algorithms import CEM, CMAES, PGPE, SNES, XNES  # noqa: E402
from evotorch.distributions import (  # noqa: E402
    ExpGaussian,
    ExpSeparableGaussian,
    SeparableGaussian,
    SymmetricSeparableGaussian,
)
from evotorch.optimizers import SGD, Adam, ClipUp  # noqa: E402
from evotorch.tools import modify_tensor  # noqa: E402
from evotorch.tools.misc import make_gaussian  # noqa: E402
from evotorch.tools.ranking import rank  # noqa: E402

assert "/root/reference" in evotorch.__file__, evotorch.__file__


def T(x):
    return torch.as_tensor(np.asarray(x), dtype=torch.float32)


def npy(t):
    return t.detach().cpu().numpy().copy()


def rastrigin(x: torch.Tensor) -> torch.Tensor:
    A = 10
    (_, n) = x.shape
    return A * n + torch.sum((x**2) - A * torch.cos(2 * np.pi * x), 1)


def sphere(x: torch.Tensor) -> torch.Tensor:
    return torch.sum(x**2, dim=-1)


out = {}

# ---------------------------------------------------------------- ranking
rng = np.random.default_rng(1234)
rank_inputs = {
    "appxB": np.array([3, 1, 7, 2, 5, 4], dtype=np.float32),
    "reftest0": np.array([0.0, 4.0, 8.0], dtype=np.float32),  # shapes of tests/test_ranking.py vectors
    "reftest1": np.array([-3.0, 10.0, 2.0, 7.5], dtype=np.float32),
    "rand257": rng.standard_normal(257).astype(np.float32) * 100,
    "rand1000": (rng.standard_normal(1000) * 717 + 1.1e3).astype(np.float32),
    "n2": np.array([2.0, -1.0], dtype=np.float32),
}
for name, f in rank_inputs.items():
    assert len(np.unique(f)) == len(f), name  # tie-free: the unstable reference sort is well defined
    out[f"rank/{name}/f"] = f
    for method in ("centered", "linear", "nes", "normalized", "raw"):
        for hib in (True, False):
            out[f"rank/{name}/{method}/{int(hib)}"] = npy(rank(T(f), method, higher_is_better=hib))

# Tied fitnesses (fp32 Rastrigin at D=10k collides massively): the reference's
# default argsort is unstable, so for THIS vector only the reference is run with torch's sort forced to
# stable=True -- the tie-break contract of the new engine (equal keys keep ascending index order).
_orig_tensor_argsort, _orig_argsort = torch.Tensor.argsort, torch.argsort
torch.Tensor.argsort = lambda self, *a, **k: _orig_tensor_argsort(self, *a, **{**k, "stable": True})
torch.argsort = lambda x, *a, **k: _orig_argsort(x, *a, **{**k, "stable": True})
tied = np.round(rng.standard_normal(600) * 3).astype(np.float32)
tied[::50] = 0.0
tied[1::50] = -0.0
assert len(np.unique(tied)) < 40
out["rank/tied600/f"] = tied
for method in ("centered", "linear", "nes", "normalized", "raw"):
    for hib in (True, False):
        out[f"rank/tied600/{method}/{int(hib)}"] = npy(rank(T(tied), method, higher_is_better=hib))
torch.Tensor.argsort, torch.argsort = _orig_tensor_argsort, _orig_argsort

# ---------------------------------------------------------------- sampling layout
g = torch.Generator().manual_seed(77)
raw = make_gaussian(10, 7, symmetric=True, generator=g, dtype=torch.float32)
mu = T(rng.standard_normal(7))
sg = T(np.abs(rng.standard_normal(7)) + 0.1)
g = torch.Generator().manual_seed(77)
shaped = make_gaussian(10, 7, center=mu, stdev=sg, symmetric=True, generator=g, dtype=torch.float32)
out["sample/sym/raw"] = npy(raw)
out["sample/sym/mu"] = npy(mu)
out["sample/sym/sigma"] = npy(sg)
out["sample/sym/out"] = npy(shaped)
g = torch.Generator().manual_seed(78)
raw = make_gaussian(9, 7, generator=g, dtype=torch.float32)
g = torch.Generator().manual_seed(78)
shaped = make_gaussian(9, 7, center=mu, stdev=sg, generator=g, dtype=torch.float32)
out["sample/nonsym/raw"] = npy(raw)
out["sample/nonsym/out"] = npy(shaped)

# ---------------------------------------------------------------- gradients
N, D = 64, 16
mu = T(rng.standard_normal(D))
sg = T(np.abs(rng.standard_normal(D)) * 0.5 + 0.2)
Z = rng.standard_normal((N // 2, D)).astype(np.float32)
Xsym = np.empty((N, D), dtype=np.float32)
Xsym[0::2] = npy(mu) + npy(sg) * Z
Xsym[1::2] = npy(mu) - npy(sg) * Z
Xns = (npy(mu) + npy(sg) * rng.standard_normal((N, D))).astype(np.float32)
fsym = npy(rastrigin(T(Xsym)))
fns = npy(rastrigin(T(Xns)))
assert len(np.unique(fsym)) == N and len(np.unique(fns)) == N
out["grad/mu"], out["grad/sigma"] = npy(mu), npy(sg)
out["grad/Xsym"], out["grad/fsym"] = Xsym, fsym
out["grad/Xns"], out["grad/fns"] = Xns, fns
for method in ("centered", "linear", "nes", "normalized", "raw"):
    for sense in ("min", "max"):
        for div in ("num_directions", "num_solutions", "total_weight", "weight_stdev", None):
            p = {"mu": mu, "sigma": sg}
            if div is not None:
                p.update({"divide_mu_grad_by": div, "divide_sigma_grad_by": div})
            d = SymmetricSeparableGaussian(p)
            gr = d.compute_gradients(T(Xsym), T(fsym), objective_sense=sense, ranking_method=method)
            out[f"grad/sym/{method}/{sense}/{div}/mu"] = npy(gr["mu"])
            out[f"grad/sym/{method}/{sense}/{div}/sigma"] = npy(gr["sigma"])
            d = SeparableGaussian(dict(p))
            gr = d.compute_gradients(T(Xns), T(fns), objective_sense=sense, ranking_method=method)
            out[f"grad/sep/{method}/{sense}/{div}/mu"] = npy(gr["mu"])
            out[f"grad/sep/{method}/{sense}/{div}/sigma"] = npy(gr["sigma"])
        d = ExpSeparableGaussian({"mu": mu, "sigma": sg})
        gr = d.compute_gradients(T(Xns), T(fns), objective_sense=sense, ranking_method=method)
        out[f"grad/exp/{method}/{sense}/mu"] = npy(gr["mu"])
        out[f"grad/exp/{method}/{sense}/sigma"] = npy(gr["sigma"])
for ratio in (0.5, 0.25, 0.1):
    for sense in ("min", "max"):
        d = SeparableGaussian({"mu": mu, "sigma": sg, "parenthood_ratio": ratio})
        gr = d.compute_gradients(T(Xns), T(fns), objective_sense=sense, ranking_method=None)
        out[f"grad/cem/{ratio}/{sense}/mu"] = npy(gr["mu"])
        out[f"grad/cem/{ratio}/{sense}/sigma"] = npy(gr["sigma"])

# XNES gradient
Dx = 5
A = T(np.eye(Dx) * 0.7 + 0.1 * rng.standard_normal((Dx, Dx)))
mux = T(rng.standard_normal(Dx))
dist = ExpGaussian({"mu": mux, "sigma": A.clone()})
Xx = npy(dist.sample(20, generator=torch.Generator().manual_seed(5)))
fx = npy(sphere(T(Xx)))
out["xnes/mu"], out["xnes/A"], out["xnes/A_inv"] = npy(mux), npy(dist.A), npy(dist.A_inv)
out["xnes/X"], out["xnes/f"] = Xx, fx
for method in ("nes", "centered"):
    gr = dist.compute_gradients(T(Xx), T(fx), objective_sense="min", ranking_method=method)
    out[f"xnes/{method}/d"], out[f"xnes/{method}/M"] = npy(gr["d"]), npy(gr["M"])
    upd = dist.update_parameters(gr, learning_rates={"mu": 1.0, "sigma": 0.3})
    out[f"xnes/{method}/new_mu"] = npy(upd.mu)
    out[f"xnes/{method}/new_A"] = npy(upd.A)
    out[f"xnes/{method}/new_A_inv"] = npy(upd.A_inv)

# ---------------------------------------------------------------- optimizers
grads7 = rng.standard_normal((7, 12)).astype(np.float32)
out["opt/grads"] = grads7
for i, (ss, mom, ms) in enumerate([(0.1, 0.9, None), (0.1, 0.95, 0.3), (0.2, 0.5, 0.3), (0.15, 0.9, 0.3)]):
    kw = dict(solution_length=12, dtype="float32", stepsize=ss, momentum=mom)
    if ms is not None:
        kw["max_speed"] = ms
    opt = ClipUp(**kw)
    out[f"opt/clipup/{i}/cfg"] = np.array([ss, mom, -1.0 if ms is None else ms], dtype=np.float64)
    out[f"opt/clipup/{i}/steps"] = np.stack([npy(opt.ascent(T(gv))) for gv in grads7])
opt = Adam(solution_length=12, dtype="float32", stepsize=0.05)
out["opt/adam/0/steps"] = np.stack([npy(opt.ascent(T(gv))) for gv in grads7])
opt = Adam(solution_length=12, dtype="float32", stepsize=0.01, beta1=0.8, beta2=0.95, epsilon=1e-6)
out["opt/adam/1/steps"] = np.stack([npy(opt.ascent(T(gv))) for gv in grads7])
opt = SGD(solution_length=12, dtype="float32", stepsize=0.1)
out["opt/sgd/0/steps"] = np.stack([npy(opt.ascent(T(gv))) for gv in grads7])
opt = SGD(solution_length=12, dtype="float32", stepsize=0.1, momentum=0.8)
out["opt/sgd/1/steps"] = np.stack([npy(opt.ascent(T(gv))) for gv in grads7])

# ---------------------------------------------------------------- modify_tensor
x = T([10, 11, 12]); tgt = T([0, 21, 22])
out["modify/orig"], out["modify/target"] = npy(x), npy(tgt)
out["modify/lb5"] = npy(modify_tensor(x, tgt, lb=5))
out["modify/lb5ub20"] = npy(modify_tensor(x, tgt, lb=5, ub=20))
out["modify/mc05"] = npy(modify_tensor(x, tgt, max_change=0.5))
out["modify/lb7ub17mc05"] = npy(modify_tensor(x, tgt, lb=7, ub=17, max_change=0.5))
xo = T(rng.standard_normal(16)); xt = T(rng.standard_normal(16) * 3)
out["modify/r/orig"], out["modify/r/target"] = npy(xo), npy(xt)
out["modify/r/mc02"] = npy(modify_tensor(xo, xt, max_change=0.2))
out["modify/r/lbm1ub1mc05"] = npy(modify_tensor(xo, xt, lb=-1.0, ub=1.0, max_change=0.5))

# ---------------------------------------------------------------- seeded trajectories
def run_traj(tag, make_searcher, n_gen, D, sense="min", fn=rastrigin):
    prob = Problem(sense, fn, initial_bounds=(-5.12, 5.12), solution_length=D, vectorized=True, seed=11, dtype=torch.float32)
    s = make_searcher(prob)
    mus, sigs, Xs, fs = [], [], [], []
    for _ in range(n_gen):
        s.step()
        mus.append(npy(s.status["center"])); sigs.append(npy(s.status["stdev"]))
        Xs.append(npy(s.population.values)); fs.append(npy(s.population.evals[:, 0]))
    out[f"traj/{tag}/mu"], out[f"traj/{tag}/sigma"] = np.stack(mus), np.stack(sigs)
    out[f"traj/{tag}/X"], out[f"traj/{tag}/f"] = np.stack(Xs), np.stack(fs)
    for t in range(n_gen):
        assert len(np.unique(fs[t])) == len(fs[t]), (tag, t)
    return s


run_traj("pgpe", lambda p: PGPE(p, popsize=32, center_learning_rate=0.5, stdev_learning_rate=0.1, stdev_init=1.0), 6, 8)
run_traj("pgpe_max", lambda p: PGPE(p, popsize=32, center_learning_rate=0.5, stdev_learning_rate=0.1, stdev_init=1.0),
         5, 8, sense="max")
run_traj("pgpe_nonsym_adam", lambda p: PGPE(p, popsize=30, center_learning_rate=0.05, stdev_learning_rate=0.1, stdev_init=1.0,
                                             symmetric=False, optimizer="adam"), 5, 8)
run_traj("pgpe_nes_rank", lambda p: PGPE(p, popsize=32, center_learning_rate=0.3, stdev_learning_rate=0.1, radius_init=4.0,
                                          ranking_method="nes", optimizer=None, stdev_min=0.01, stdev_max=2.0), 5, 8)
run_traj("snes", lambda p: SNES(p, popsize=24, stdev_init=2.0), 6, 8)
run_traj("snes_clipup", lambda p: SNES(p, popsize=24, stdev_init=2.0, optimizer="clipup", center_learning_rate=0.2,
                                        stdev_max_change=0.3), 5, 8)
run_traj("cem", lambda p: CEM(p, popsize=40, parenthood_ratio=0.25, stdev_init=2.0, stdev_max_change=0.5), 5, 8)
s = run_traj("xnes", lambda p: XNES(p, popsize=16, stdev_init=1.5), 5, 5, fn=sphere)

# CMA-ES: record the internally sampled zs too
prob = Problem("min", sphere, initial_bounds=(-3, 3), solution_length=6, vectorized=True, seed=3, dtype=torch.float32)
cma = CMAES(prob, stdev_init=1.0, popsize=12)
rec = {k: [] for k in ("Z", "f", "m", "sigma", "C", "A", "p_sigma", "p_c")}
orig_sample = cma.sample_distribution


def recording_sample(num_samples=None):
    zs, ys, xs = orig_sample(num_samples)
    rec["Z"].append(npy(zs))
    return zs, ys, xs


cma.sample_distribution = recording_sample
out["cmaes/m0"] = npy(cma.m)
out["cmaes/weights"] = npy(cma.weights)
out["cmaes/consts"] = np.array([float(cma.mu_eff), float(cma.c_sigma), float(cma.damp_sigma), float(cma.c_c), float(cma.c_1),
                                float(cma.c_mu), float(cma.decompose_C_freq)], dtype=np.float64)
for _ in range(6):
    cma.step()
    rec["f"].append(npy(cma.population.evals[:, 0])); rec["m"].append(npy(cma.m)); rec["sigma"].append(npy(cma.sigma).reshape(1))
    rec["C"].append(npy(cma.C)); rec["A"].append(npy(cma.A)); rec["p_sigma"].append(npy(cma.p_sigma)); rec["p_c"].append(npy(cma.p_c))
for k, v in rec.items():
    out[f"cmaes/{k}"] = np.stack(v)

# appendix-C constants at cfg3 (D=1024, popsize=4096) - construction only
prob = Problem("min", sphere, initial_bounds=(-3, 3), solution_length=1024, vectorized=True, seed=3, dtype=torch.float32)
cma = CMAES(prob, stdev_init=1.0, popsize=4096)
out["cmaes/cfg3_consts"] = np.array([float(cma.mu_eff), float(cma.c_sigma), float(cma.damp_sigma), float(cma.c_c), float(cma.c_1),
                                     float(cma.c_mu), float(cma.decompose_C_freq), float(torch.sum(cma.weights))], dtype=np.float64)

# ---------------------------------------------------------------- batched policy forward
from evotorch.neuroevolution.net.vecrl import Policy  # noqa: E402

net = torch.nn.Sequential(torch.nn.Linear(11, 8), torch.nn.Tanh(), torch.nn.Linear(8, 3))
pol = Policy(net)
L = pol.parameter_length
params = T(rng.standard_normal((6, L)) * 0.3)
obs = T(rng.standard_normal((6, 11)))
pol.set_parameters(params)
out["policy/params"], out["policy/obs"] = npy(params), npy(obs)
out["policy/act"] = npy(pol(obs))
out["policy/dims"] = np.array([11, 8, 3, L])

np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
print("wrote", len(out), "arrays ->", os.path.join(HERE, "reference_golden.npz"), file=sys.stderr)
