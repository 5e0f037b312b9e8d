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
register accumulation; the lo halves of the 3xTF32 operands are derived inside the kernel, so operands are read from HBM once); the
glue between the contractions is fused into four small kernels (csrc/evok_cmaes.cu, evok_rank_table) and the covariance update is
applied by the SYRK's epilogue, so a generation is ~14 launches with no host reads and replays from a CUDA graph
(`enable_cuda_graph()`).  The Cholesky factorisation stays on cuSOLVER (torch.linalg.cholesky_ex): the repo's own persistent
tile-dataflow kernel (csrc/evok_chol.cu, EVOTORCH_B200_EVOK_CHOLESKY=1) is correct but measured 2.2x slower at D = 1024.
"""

from __future__ import annotations

import math
import os
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
        self._use_graph, self._graph = os.environ.get("EVOTORCH_B200_CUDA_GRAPH", "0") == "1", None
        SinglePopulationAlgorithmMixin.__init__(self)

    # ------------------------------------------------------------------ accessors
    @property
    def population(self) -> SolutionBatch:
        return self._population

    @property
    def obj_index(self) -> int:
        return self._obj_index

    def _get_center(self) -> torch.Tensor:
        return self.m.clone() if self.__dict__.get("_fused") is not None else self.m  # the fused step updates `m` in place

    def _get_sigma(self) -> float:
        return float(self.sigma)

    # ------------------------------------------------------------------ one generation
    def sample_distribution(self, num_samples: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """zs ~ N(0, I); ys = zs A^T; xs = m + sigma ys (cmaes.py:408-430)."""
        n = self.popsize if num_samples is None else int(num_samples)
        problem = self._problem
        d = problem.solution_length
        fs = self.__dict__.get("_fused") if n == self.popsize else None  # persistent buffers of the fused generation
        zs = problem.make_empty(num_solutions=n) if fs is None else fs["zs"]
        if ops.uses_kernels(zs) and problem.rng == "philox":
            seed, stream_id = problem.next_philox_stream()
            zero, one = (problem.make_zeros(d), problem.make_ones(d)) if fs is None else (fs["zero"], fs["one"])
            ops.sample_eval(ops.OBJ_NONE, zs, zero, one, n_rows=n, symmetric=False, seed=seed, stream_id=stream_id,
                            stream_offset=problem.philox_stream_offset)
        else:
            problem.make_gaussian(out=zs)
        if self.separable:
            ys = self.A.unsqueeze(0) * zs
        elif ops.uses_kernels(zs) and ops.uses_kernels(self.A):
            # K6: one tcgen05 GEMM (3xTF32, fp32-accurate) with the affine epilogue xs = m + sigma * ys fused in; in the fused
            # generation xs IS the population's value buffer
            ys = torch.empty_like(zs) if fs is None else fs["ys"]
            xs = torch.empty_like(zs) if fs is None else self._population._data
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
        if self.separable:
            self.A = self.C.pow(0.5)
        elif ops.uses_kernels(self.C) and self.C.is_contiguous() and os.environ.get("EVOTORCH_B200_EVOK_CHOLESKY", "0") == "1":
            self.A = ops.cholesky(self.C)
        else:
            self.A = torch.linalg.cholesky(self.C)

    # ------------------------------------------------------------------ fused generation (CUDA float32, full covariance)
    def _fused_ok(self) -> bool:
        return (not self.separable and self.stdev_min is None and self.stdev_max is None and ops.uses_kernels(self.m) and ops.uses_kernels(self.C)
                and self._problem.rng == "philox" and self._population._data.is_contiguous()
                and self._population._evdata.shape[1] == 1 and self._population._evdata.dtype == torch.float32)

    def _fused_state(self) -> dict:
        fs = self.__dict__.get("_fused")
        if fs is None:
            p, n, d = self._problem, self.popsize, self._problem.solution_length
            dev = self.m.device
            new = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)  # noqa: E731
            fs = self._fused = dict(zs=new(n, d), ys=new(n, d), aw=new(n), w_pos=new(n), w_act=new(n), local=new(d), shaped=new(d), scratch=new(d),
                                    k=new(3), zero=p.make_zeros(d), one=p.make_ones(d), info=torch.zeros((), dtype=torch.int32, device=dev),
                                    steps_dev=None)
            # state tensors become persistent buffers that the kernels update in place (pointer-stable: CUDA-graph replay)
            self.m, self.p_sigma, self.p_c = self.m.contiguous().clone(), self.p_sigma.contiguous().clone(), self.p_c.contiguous().clone()
            self.sigma = self.sigma.reshape(()).clone()
            self.C, self.A = self.C.contiguous().clone(), self.A.contiguous().clone()
            self._consts = (self.c_m, self.c_sigma, self.damp_sigma, self.c_c, self.c_1, self.c_mu, self.variance_discount_sigma,
                            self.variance_discount_c, float(self.unbiased_expectation), self._weights_sum)
        return fs

    def _step_fused(self):
        """One generation as a short chain of kernels with no host reads (cmaes.py:567-606):
        K1 z-sampling -> GEMM (Y = Z A^T, X = m + sigma Y straight into the population) -> evaluate -> rank-to-weights (K3, one launch)
        -> row weights (positive part / active reweighting, one pass over Z) -> two weighted row sums (K4) -> fused vector update
        (m, p_sigma, sigma, h_sig, p_c + the covariance coefficients) -> weighted SYRK with the covariance update in its epilogue ->
        Cholesky.  Every state tensor is updated in place."""
        fs = self._fused_state()
        zs, ys, xs = self.sample_distribution()
        pop = self._population
        if xs.data_ptr() == pop._data.data_ptr():
            pop._evdata.fill_(float("nan"))
        else:  # an overriding `sample_distribution` (e.g. recorded draws in the tests) returns its own tensors
            pop.set_values(xs)
        self._problem.evaluate(pop)
        f = pop._evdata.view(-1)
        ops.rank_table(f, self._problem.senses[self._obj_index] == "max", self.weights, out=fs["aw"])
        ops.cmaes_row_weights(fs["aw"], zs, self.active, fs["w_pos"], fs["w_act"])
        ops.grad(ops.GRAD_MOMENTS, zs, fs["w_pos"], fs["zero"], fs["one"], 1.0, 1.0, out_mu=fs["local"], out_sigma=fs["scratch"])
        ops.grad(ops.GRAD_MOMENTS, ys, fs["w_pos"], fs["zero"], fs["one"], 1.0, 1.0, out_mu=fs["shaped"], out_sigma=fs["scratch"])
        ops.cmaes_vector_update(fs["local"], fs["shaped"], self.m, self.p_sigma, self.p_c, self.sigma, self._consts, self.csa_squared, fs["k"],
                                steps=self._steps_count, steps_dev=fs["steps_dev"])
        ops.weighted_syrk_update(ys, fs["w_act"], fs["k"], self.C, u=self.p_c, out=self.C)
        if fs["steps_dev"] is not None or (self._steps_count + 1) % self.decompose_C_freq == 0:
            if os.environ.get("EVOTORCH_B200_EVOK_CHOLESKY", "0") == "1":
                ops.cholesky(self.C, out=self.A)  # the repo's own tile-dataflow kernel (csrc/evok_chol.cu): correct, but measured
                # 2.2x SLOWER than cuSOLVER's potrf at D = 1024 (0.75 vs 0.34 ms, profiles/r02_cholesky.txt), so it is not the default
            else:
                torch.linalg.cholesky_ex(self.C, check_errors=False, out=(self.A, fs["info"]))

    # ------------------------------------------------------------------ CUDA-graph replay of the fused generation
    def enable_cuda_graph(self, enabled: bool = True):
        """Capture the fused generation into a CUDA graph and replay it from `step()` (one graph launch per generation; the
        z-sampler reads a device-side generation counter, `_h_sig` a device-side step counter, so the replayed trajectory equals
        eager stepping).  Used when the configuration is capturable: fused path, built-in objective, no evaluation hooks, Cholesky
        every generation (`decompose_C_freq == 1`); otherwise stepping stays eager."""
        self._use_graph = bool(enabled)
        self._graph = None
        return self

    def _graph_capturable(self) -> bool:
        prob = self._problem
        return (self._fused_ok() and self.decompose_C_freq == 1 and prob.evok_objective_id is not None and len(prob.before_eval_hook) == 0
                and len(prob.after_eval_hook) == 0 and not prob.stores_solution_stats and "sample_distribution" not in self.__dict__)

    def _step_graph(self):
        from .. import _native as nat

        prob = self._problem
        if self._graph is None:
            self._step_fused()  # warm every kernel / workspace / cuSOLVER handle eagerly, right before the capture
            fs = self._fused_state()
            prob.philox_stream_offset = torch.zeros(1, dtype=torch.int32, device=self.m.device)
            fs["steps_dev"] = torch.full((1,), self._steps_count + 1, dtype=torch.int64, device=self.m.device)
            base = prob._philox_stream
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            before = ops.launch_count()
            try:
                with nat.private_workspaces() as store, torch.cuda.graph(graph):
                    self._step_fused()
                    prob.philox_stream_offset.add_(1)
            except Exception:  # e.g. a library call inside the step that cannot be captured on this build: stay eager
                prob.philox_stream_offset, fs["steps_dev"] = None, None
                prob._philox_stream = base
                self._use_graph = False
                torch.cuda.synchronize()
                return
            self._graph_kernels = ops.launch_count() - before
            ops.count_replayed_launches(-self._graph_kernels)
            prob._philox_stream = base  # the capture consumed a host-side stream id without running anything
            prob.philox_stream_offset.zero_()
            self._graph, self._graph_workspaces = graph, store
            return
        self._graph.replay()
        ops.count_replayed_launches(self._graph_kernels)
        prob._philox_stream += 1

    def __getstate__(self) -> dict:
        state = dict(self.__dict__)
        state["_graph"] = None
        state.pop("_graph_workspaces", None)
        if state.get("_fused") is not None:
            state["_fused"] = None  # scratch buffers are rebuilt on the first step after loading
        return state

    def _step(self):
        if self._fused_ok():
            if self.__dict__.get("_use_graph") and self._graph_capturable():
                self._step_graph()
            else:
                self._graph = None
                if self.__dict__.get("_fused") is not None and self._fused.get("steps_dev") is not None:
                    self._fused["steps_dev"], self._problem.philox_stream_offset = None, None
                self._step_fused()
            return
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
