"""Golden vectors for RunningNorm / ObsNormLayer from the REAL reference (evotorch.neuroevolution.net.runningnorm):

    PYTHONPATH=tests/golden/_refstubs:/root/reference/src EVOTORCH_VERBOSE_LEVEL=0 python tests/golden/gen_runningnorm_golden.py
"""

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))

import evotorch  # noqa: E402
from evotorch.neuroevolution.net.runningnorm import RunningNorm  # noqa: E402

assert "/root/reference" in evotorch.__file__
rng = np.random.default_rng(5)
out = {}
D = 9
batches = [(rng.standard_normal((n, D)) * rng.uniform(0.01, 4.0, size=D) + rng.uniform(-3, 3, size=D)).astype(np.float32) for n in (1, 7, 32, 5)]
masks = [None, rng.random(7) < 0.6, rng.random(32) < 0.3, np.zeros(5, dtype=bool)]
single = rng.standard_normal(D).astype(np.float32)
probe = (rng.standard_normal((6, D)) * 5).astype(np.float32)
out["probe"], out["single"] = probe, single
for tag, clip in (("noclip", None), ("clip", (-1.5, 2.0))):
    rn = RunningNorm(shape=D, dtype="float32", min_variance=1e-2, clip=clip)
    for i, (b, m) in enumerate(zip(batches, masks)):
        out[f"batch{i}"] = b
        out[f"mask{i}"] = np.ones(len(b), dtype=bool) if m is None else m
        rn.update(torch.as_tensor(b), None if m is None else torch.as_tensor(m))
        out[f"{tag}/sum{i}"], out[f"{tag}/sumsq{i}"] = rn.sum.numpy().copy(), rn.sum_of_squares.numpy().copy()
        out[f"{tag}/count{i}"] = np.array(rn.count)
        out[f"{tag}/mean{i}"], out[f"{tag}/stdev{i}"] = rn.mean.numpy().copy(), rn.stdev.numpy().copy()
        out[f"{tag}/norm{i}"] = rn.normalize(torch.as_tensor(probe)).numpy().copy()
    rn.update(torch.as_tensor(single))
    out[f"{tag}/count_single"] = np.array(rn.count)
    out[f"{tag}/norm_single"] = rn.normalize(torch.as_tensor(probe)).numpy().copy()
    other = RunningNorm(shape=D, dtype="float32")
    other.update(torch.as_tensor(batches[2]))
    rn.update(other)
    out[f"{tag}/count_merged"] = np.array(rn.count)
    out[f"{tag}/mean_merged"], out[f"{tag}/stdev_merged"] = rn.mean.numpy().copy(), rn.stdev.numpy().copy()
    out[f"{tag}/layer"] = rn.to_layer()(torch.as_tensor(probe)).numpy().copy()
    out[f"{tag}/update_and_normalize"] = rn.update_and_normalize(torch.as_tensor(batches[1]), torch.as_tensor(masks[1])).numpy().copy()
np.savez_compressed(os.path.join(HERE, "runningnorm_golden.npz"), **out)
print("wrote", len(out), "arrays", file=sys.stderr)
