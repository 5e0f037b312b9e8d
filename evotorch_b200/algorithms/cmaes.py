"""CMA-ES with a full covariance matrix and active weights (mirrors evotorch.algorithms.cmaes.CMAES, cmaes.py:90-606:
same constructor arguments, hyper-parameter formulas and status keys).

Per generation (cmaes.py:567-606):
    z ~ N(0, I) (K1 Philox sampler on CUDA)  ->  Y = Z A^T, X = m + sigma*Y (one GEMM with the affine epilogue)
    evaluate (K2 for built-in objectives)  ->  stable argsort (K3)  ->  rank -> weight gather
    weighted recombinations sum_i w_i z_i, sum_i w_i y_i (K4 weighted column sums)
    evolution paths, sigma, rank-1 + rank-mu update of C, Cholesky.
The rank-mu term is computed as Y^T diag(w) Y (a weighted SYRK): the reference materialises an N x D x D broadcast
temporary (cmaes.py:548; 16 GiB at D = 1024, N = 4096).  On CUDA fp32 both contractions run on the hand-written tcgen05
kernel (csrc/evok_gemm.cu: TMA -> 128B-swizzled smem -> tcgen05.mma.kind::tf32 with 3xTF32 operand splitting -> TMEM ->
register accumulation); the Cholesky factorisation stays on cuSOLVER (torch.linalg.cholesky).
"""

from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch

from .. import ops
from ..core import Problem, Solution, SolutionBatch
from .searchalgorithm import SearchAlgorithm, SinglePopulationAlgorithmMixin


def _safe_divide(a, b):
    tolerance = 1e-8
    if abs(b) < tolerance:
        b = (-tolerance) if b < 0 else tolerance
    return a / b


class CMAES(SearchAlgorithm, SinglePopulationAlgorithmMixin):
    def __init__(self, problem: Problem, *, stdev_init, popsize: Optional[int] = None, center_init=None, c_m: float = 1.0,
                 c_sigma: Optional[float] = None, c_sigma_ratio: float = 1.0, damp_sigma: Optional[float] = None,
                 damp_sigma_ratio: float = 1.0, c_c: Optional[float] = None, c_c_ratio: float = 1.0, c_1: Optional[float] = None,
                 c_1_ratio: float = 1.0, c_mu: Optional[float] = None, c_mu_ratio: float = 1.0, active: bool = True,
                 csa_squared: bool = False, stdev_min: Optional[float] = None, stdev_max: Optional[float] = None,
                 separable: bool = False, limit_C_decomposition: bool = True, obj_index: Optional[int] = None):
        SearchAlgorithm.__init__(self, problem, center=self._get_center, stepsize=self._get_sigma)
        problem.ensure_numeric()
        problem.ensure_unbounded()
        self._obj_index = problem.normalize_obj_index(obj_index)
        d = problem.solution_length
        if not popsize:
            popsize = 4 + int(np.floor(3 * np.log(d)))  # cmaes.py:270-272
        self.popsize = int(popsize)
        self.mu = int(np.floor(popsize / 2))
        self._population = problem.generate_batch(popsize=popsize)
        self.separable = bool(separable)

        if center_init is None:
            center_init = problem.generate_values(1)
        elif isinstance(center_init, Solution):
            center_init = center_init.values.clone()
        self.m = problem.make_tensor(center_init).squeeze().clone()
        if not (self.m.ndim == 1 and len(self.m) == d):
            raise ValueError(f"The initial center point was expected as a vector of length {d}."
                             " However, the provided `center_init` has (or implies) a different shape.")
        self.sigma = problem.make_tensor(stdev_init)
        if separable:
            self.C = problem.make_ones(d)
            self.A = problem.make_ones(d)
        else:
            self.C = problem.make_I(d)
            self.A = self.C.clone()

        # weights and learning rates (cmaes.py:300-385); host scalars in float64, the weight vector in the problem dtype
        raw_weights = problem.make_tensor(np.log((popsize + 1) / 2) - torch.log(torch.arange(popsize) + 1))
        positive, negative = raw_weights[: self.mu], raw_weights[self.mu:]
        self.mu_eff = float(torch.sum(positive).pow(2.0) / torch.sum(positive.pow(2.0)))
        self.c_m, self.active, self.csa_squared = c_m, bool(active), bool(csa_squared)
        self.stdev_min, self.stdev_max = stdev_min, stdev_max
        mu_eff = self.mu_eff
        if c_sigma is None:
            c_sigma = (mu_eff + 2.0) / (d + mu_eff + 3)
        self.c_sigma = c_sigma_ratio * c_sigma
        if damp_sigma is None:
            damp_sigma = 1 + 2 * max(0.0, math.sqrt((mu_eff - 1) / (d + 1)) - 1) + self.c_sigma
        self.damp_sigma = damp_sigma_ratio * damp_sigma
        if c_c is None:
            if separable:
                c_c = (1 + (1 / d) + (mu_eff / d)) / (d**0.5 + (1 / d) + 2 * (mu_eff / d))
            else:
                c_c = (4 + mu_eff / d) / (d + (4 + 2 * mu_eff / d))
        self.c_c = c_c_ratio * c_c
        if c_1 is None:
            if separable:
                c_1 = 1.0 / (d + 2.0 * np.sqrt(d) + mu_eff / d)
            else:
                c_1 = min(1, popsize / 6) * 2 / ((d + 1.3) ** 2.0 + mu_eff)
        self.c_1 = float(c_1_ratio * c_1)
        if c_mu is None:
            if separable:
                c_mu = (0.25 + mu_eff + (1.0 / mu_eff) - 2) / (d + 4 * np.sqrt(d) + (mu_eff / 2.0))
            else:
                c_mu = min(1 - self.c_1, 2 * ((0.25 + mu_eff - 2 + (1 / mu_eff)) / ((d + 2) ** 2.0 + mu_eff)))
        self.c_mu = float(c_mu_ratio * c_mu)
        self.variance_discount_sigma = math.sqrt(self.c_sigma * (2 - self.c_sigma) * mu_eff)
        self.variance_discount_c = math.sqrt(self.c_c * (2 - self.c_c) * mu_eff)

        positive = positive / torch.sum(positive)
        if self.active:
            mu_eff_neg = float(torch.sum(negative).pow(2.0) / torch.sum(negative.pow(2.0)))
            alpha = min(1 + self.c_1 / self.c_mu, 1 + 2 * mu_eff_neg / (mu_eff + 2), (1 - self.c_mu - self.c_1) / (d * self.c_mu))
            negative = alpha * negative / torch.sum(torch.abs(negative))
        else:
            negative = torch.zeros_like(negative)
        self.weights = torch.cat([positive, negative], dim=-1)
        self._weights_sum = float(torch.sum(self.weights))

        self.p_sigma = problem.make_zeros(d)
        self.p_c = problem.make_zeros(d)
        self.unbiased_expectation = np.sqrt(d) * (1 - (1 / (4 * d)) + 1 / (21 * d**2))
        if limit_C_decomposition:
            self.decompose_C_freq = max(1, int(np.floor(_safe_divide(1, 10 * d * (self.c_1 + self.c_mu)))))
        else:
            self.decompose_C_freq = 1
        SinglePopulationAlgorithmMixin.__init__(self)

    # ------------------------------------------------------------------ accessors
    @property
    def population(self) -> SolutionBatch:
        return self._population

    @property
    def obj_index(self) -> int:
        return self._obj_index

    def _get_center(self) -> torch.Tensor:
        return self.m

    def _get_sigma(self) -> float:
        return float(self.sigma)

    # ------------------------------------------------------------------ one generation
    def sample_distribution(self, num_samples: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """zs ~ N(0, I); ys = zs A^T; xs = m + sigma ys (cmaes.py:408-430)."""
        n = self.popsize if num_samples is None else int(num_samples)
        problem = self._problem
        d = problem.solution_length
        zs = problem.make_empty(num_solutions=n)
        if ops.uses_kernels(zs) and problem.rng == "philox":
            seed, stream_id = problem.next_philox_stream()
            zero, one = problem.make_zeros(d), problem.make_ones(d)
            ops.sample_eval(ops.OBJ_NONE, zs, zero, one, n_rows=n, symmetric=False, seed=seed, stream_id=stream_id)
        else:
            problem.make_gaussian(out=zs)
        if self.separable:
            ys = self.A.unsqueeze(0) * zs
        elif ops.uses_kernels(zs) and ops.uses_kernels(self.A):
            # K6: one tcgen05 GEMM (3xTF32, fp32-accurate) with the affine epilogue xs = m + sigma * ys fused in
            ys = torch.empty_like(zs)
            xs = torch.empty_like(zs)
            ops.gemm_nt(zs, self.A.contiguous(), ys, out2=xs, alpha=self.sigma.reshape(1), bias=self.m.contiguous())
            return zs, ys, xs
        else:
            ys = zs @ self.A.T
        xs = self.m.unsqueeze(0) + self.sigma * ys
        return zs, ys, xs

    def get_population_weights(self, xs: torch.Tensor) -> torch.Tensor:
        """Evaluate, sort best-first, weight of each solution = weights[rank] (cmaes.py:432-452)."""
        self._population.set_values(xs)
        self._problem.evaluate(self._population)
        indices = self._population.argsort(obj_index=self.obj_index)
        ranks = torch.empty_like(indices)
        ranks[indices] = torch.arange(self.popsize, dtype=indices.dtype, device=indices.device)
        return self.weights[ranks]

    def _weighted_rowsum(self, w: torch.Tensor, rows: torch.Tensor) -> torch.Tensor:
        """sum_i w_i rows_i -- K4 (moments form with a zero centre) on CUDA fp32."""
        if ops.uses_kernels(rows) and ops.uses_kernels(w):
            d = rows.shape[1]
            zero = torch.zeros(d, dtype=rows.dtype, device=rows.device)
            one = torch.ones(d, dtype=rows.dtype, device=rows.device)
            s1, _ = ops.grad(ops.GRAD_MOMENTS, rows.contiguous(), w.contiguous(), zero, one, 1.0, 1.0)
            return s1
        return torch.mv(rows.T, w)

    def update_m(self, zs, ys, assigned_weights) -> Tuple[torch.Tensor, torch.Tensor]:
        """Weighted recombination of the mu best (exactly the solutions with positive weight) (cmaes.py:454-481)."""
        positive = torch.clamp_min(assigned_weights, 0.0)
        local_m_displacement = self._weighted_rowsum(positive, zs)
        shaped_m_displacement = self._weighted_rowsum(positive, ys)
        self.m = self.m + self.c_m * self.sigma * shaped_m_displacement
        return local_m_displacement, shaped_m_displacement

    def update_p_sigma(self, local_m_displacement: torch.Tensor) -> None:
        self.p_sigma = (1 - self.c_sigma) * self.p_sigma + self.variance_discount_sigma * local_m_displacement

    def update_sigma(self) -> None:
        d = self._problem.solution_length
        if self.csa_squared:
            exponential_update = (torch.norm(self.p_sigma).pow(2.0) / d - 1) / 2
        else:
            exponential_update = torch.norm(self.p_sigma) / self.unbiased_expectation - 1
        self.sigma = self.sigma * torch.exp((self.c_sigma / self.damp_sigma) * exponential_update)

    def _h_sig(self) -> torch.Tensor:
        """cmaes.py:31-46 (uses the generation counter before it is incremented); stays a device scalar: no host sync."""
        d = self.p_sigma.shape[-1]
        squared_sum = torch.norm(self.p_sigma).pow(2.0) / (1 - (1 - self.c_sigma) ** (2 * self._steps_count + 1))
        return ((squared_sum / d) - 1 < 1 + 4.0 / (d + 1)).to(self.p_sigma.dtype)

    def update_p_c(self, shaped_m_displacement: torch.Tensor, h_sig: torch.Tensor) -> None:
        self.p_c = (1 - self.c_c) * self.p_c + h_sig * self.variance_discount_c * shaped_m_displacement

    def update_C(self, zs, ys, assigned_weights, h_sig) -> None:
        """Rank-1 + rank-mu update with active (negative) weights (cmaes.py:519-553)."""
        d = self._problem.solution_length
        if self.active:
            assigned_weights = torch.where(assigned_weights > 0, assigned_weights,
                                           d * assigned_weights / torch.sum(zs * zs, dim=-1))
        c1a = self.c_1 * (1 - (1 - h_sig**2) * self.c_c * (2 - self.c_c))
        weighted_pc = (self.c_1 / (c1a + 1e-23)) ** 0.5
        if self.separable:
            r1_update = c1a * (self.p_c.pow(2.0) - self.C)
            rmu_update = self.c_mu * (self._weighted_rowsum(assigned_weights, ys.pow(2.0)) - torch.sum(assigned_weights) * self.C)
        else:
            pc = weighted_pc * self.p_c
            r1_update = c1a * (torch.outer(pc, pc) - self.C)
            if ops.uses_kernels(ys) and ops.uses_kernels(assigned_weights):
                # K7: weighted SYRK Y^T diag(w) Y as one tcgen05 GEMM over K-major (w*Y)^T and Y^T (split-K over the population)
                syrk = ops.gemm_nt(ops.transpose_scale(ys.contiguous(), assigned_weights.contiguous()), ops.transpose_scale(ys.contiguous()))
            else:
                syrk = (ys.T * assigned_weights) @ ys  # no N x D x D temporary either
            rmu_update = self.c_mu * (syrk - self._weights_sum * self.C)
        self.C = self.C + r1_update + rmu_update

    def _limit_stdev(self) -> None:
        """cmaes.py:49-79."""
        diag = self.C if self.separable else torch.diag(self.C)
        stdevs = torch.clamp(self.sigma * torch.sqrt(diag), min=self.stdev_min, max=self.stdev_max)
        unscaled = (stdevs / self.sigma).pow(2.0)
        if self.separable:
            self.C = unscaled
        else:
            self.C = self.C.clone()
            torch.diagonal(self.C)[:] = unscaled

    def decompose_C(self) -> None:
        self.A = self.C.pow(0.5) if self.separable else torch.linalg.cholesky(self.C)

    def _step(self):
        zs, ys, xs = self.sample_distribution()
        assigned_weights = self.get_population_weights(xs)
        local_m_displacement, shaped_m_displacement = self.update_m(zs, ys, assigned_weights)
        self.update_p_sigma(local_m_displacement)
        self.update_sigma()
        h_sig = self._h_sig()
        self.update_p_c(shaped_m_displacement, h_sig)
        self.update_C(zs, ys, assigned_weights, h_sig)
        if self.stdev_min is not None or self.stdev_max is not None:
            self._limit_stdev()
        if (self._steps_count + 1) % self.decompose_C_freq == 0:
            self.decompose_C()
