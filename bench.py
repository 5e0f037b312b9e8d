#!/usr/bin/env python
"""bench.py -- PGPE generations/s on synthetic Rastrigin (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--popsize P] [--dim D]

Workload (config.workload): PGPE (symmetric sampling, ClipUp, centered ranking, stdev_max_change 0.2; the reference's
defaults), Rastrigin, popsize 1,000,000 x dim 10,000 fp32 -- the configuration BASELINE.json's metric is quoted on; the
40 GB population fits one B200.  With N > 1 (torchrun, one rank per GPU) the SAME population is row-sharded over the ranks
(strong scaling): per generation one all-gather of the fitness vector and one all-reduce of the stacked gradients.

One "step" = one generation through the public API (`searcher.step()`): rank -> weighted gradient reduction -> ClipUp /
sigma update -> fused Philox sampling + evaluation of a fresh population.

JSON line (rank 0): value = generations/s, device-timed (CUDA events, max over ranks) with the population resident in HBM;
e2e = the same generation driven through `Problem.sample_and_compute_gradients` with a HOST-resident distribution (mu, sigma in
pinned host memory are copied to the device every step, gradients and mean fitness are copied back; the reference's
`dist_on_cpu` actor protocol, core.py:2958); roofline = the dominant kernel (fused sample+evaluate) timed live with CUDA
events; cpu_baseline = the reference's torch-CPU op sequence (oracle/ref_cpu_path.py) on this box's host cores.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "PGPE generations/sec at popsize=1Mxdim=10k (Rastrigin, fp32)"
UNIT = "generations/s"
LR_MU, LR_SIGMA, STDEV_INIT, SEED = 0.5, 0.1, 1.0, 0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--popsize", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=10_000)
    ap.add_argument("--cpu-sample-popsize", type=int, default=2_000, help="population rows of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cuda-graph", type=int, default=-1, help="1/0: replay each generation from a CUDA graph. Default: 0 at N = 1 (kernels are timed live inside the timed region), 1 at N > 1 (the captured graph includes the NCCL collectives: +5 %% at 2-8 GPUs; the fused kernel is then timed stand-alone right after the timed region)")
    ap.add_argument("--peer", type=int, default=-1, help="1/0: at N > 1 move fitnesses and gradients between the GPUs from inside the producing kernels (NVLink peer memory, evotorch_b200/peer.py) instead of NCCL all_gather/all_reduce. Default: 1 at N > 1")
    return ap.parse_args()


def workload_config(args, n_gpus, collectives="nccl"):
    how = {"nccl": "NCCL all_gather(fitness) + all_reduce(grad)",
           "peer": "fitness gather + gradient reduction fused into the producing kernels over NVLink peer memory (no NCCL in the loop)"}[collectives]
    return {
        "workload": f"PGPE(symmetric, ClipUp, centered ranking, stdev_max_change=0.2) on Rastrigin, popsize={args.popsize}, dim={args.dim}, fp32",
        "popsize": args.popsize,
        "dim": args.dim,
        "center_learning_rate": LR_MU,
        "stdev_learning_rate": LR_SIGMA,
        "stdev_init": STDEV_INIT,
        "parallelism": f"population row-sharded over {n_gpus} GPU(s); {how}" if n_gpus > 1 else "single GPU",
        "l2": "inputs larger than L2 (population %.1f GB >> 126 MB): no flush needed" % (4.0 * args.popsize * args.dim / 1e9 / n_gpus),
    }


# ----------------------------------------------------------------------------------------------------- CPU baseline
def cpu_reference_run(args, steps: int, warmup: int) -> dict:
    """Time the reference's torch-CPU op sequence on a bounded sample (popsize `cpu_sample_popsize`, full dim) with all host
    threads, and scale linearly in popsize to the full workload (every op on the path is linear in N apart from the
    O(N log N) argsort of N floats, which is < 1 % of a generation)."""
    import torch

    from oracle.ref_cpu_path import PGPEReferencePath

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_sample = min(args.cpu_sample_popsize, args.popsize)
    n_sample -= n_sample % 2
    path = PGPEReferencePath(args.dim, n_sample, center_learning_rate=LR_MU, stdev_learning_rate=LR_SIGMA, stdev_init=STDEV_INIT, seed=SEED)
    for _ in range(max(warmup, 1)):
        path.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        path.step()
    dt = time.perf_counter() - t0
    sample_gps = steps / dt
    full_gps = sample_gps * (n_sample / args.popsize)
    torch_eager_gpu = None
    if torch.cuda.is_available() and args.impl != "reference":
        # the same torch op sequence, eager, on this GPU ("PyTorch path" comparator); bounded sample, scaled like the CPU leg
        try:
            n_gpu = min(100_000, args.popsize)
            n_gpu -= n_gpu % 2
            gpath = PGPEReferencePath(args.dim, n_gpu, center_learning_rate=LR_MU, stdev_learning_rate=LR_SIGMA, stdev_init=STDEV_INIT, seed=SEED,
                                      device="cuda")
            for _ in range(3):
                gpath.step()
            torch.cuda.synchronize()
            g0 = time.perf_counter()
            for _ in range(10):
                gpath.step()
            torch.cuda.synchronize()
            gdt = (time.perf_counter() - g0) / 10
            torch_eager_gpu = {"value": (1.0 / gdt) * (n_gpu / args.popsize), "unit": UNIT,
                               "sample": f"10 generations at popsize={n_gpu} x dim={args.dim} ({1e3 * gdt:.2f} ms each), scaled linearly in popsize "
                                         f"to {args.popsize}; the reference's torch op sequence, eager, on cuda:0"}
            del gpath
            torch.cuda.empty_cache()
        except Exception as exc:  # e.g. out of memory for the temporaries: report, do not fail the bench
            torch_eager_gpu = {"unavailable": repr(exc)[:200]}
    return {
        "value": full_gps,
        "unit": UNIT,
        "cores": cores,
        "kind": "port",
        "torch_eager_gpu": torch_eager_gpu,
        "sample": f"{steps} generations at popsize={n_sample} x dim={args.dim} ({dt:.2f} s, {sample_gps:.4f} gen/s); "
                  f"scaled linearly in popsize to {args.popsize}; torch {torch.__version__} CPU, {torch.get_num_threads()} threads",
        "sample_ms_per_step": 1e3 * dt / steps,
    }


# ----------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.lines, self.proc, self.thread = [], None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            # nvidia-smi takes a moment to start (NVML initialisation over all GPUs of the box, during which driver calls of this process
            # can stall): wait for its first sample so that none of that falls into the timed region
            deadline = time.perf_counter() + 5.0
            while not self.lines and time.perf_counter() < deadline and self.proc.poll() is None:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin: float, t_end: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for t, line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            inside = t_begin <= t <= t_end + 0.1
            try:
                if inside:
                    sm.append(float(parts[1]))
                smax = float(parts[2])
            except ValueError:
                continue
            if inside:
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------- our arm
def measured_peak_gbs() -> tuple:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def run_ours(args):
    import torch
    import torch.distributed as dist

    from evotorch_b200 import Problem, ops
    from evotorch_b200.algorithms import PGPE
    from evotorch_b200.distributions import SymmetricSeparableGaussian
    from evotorch_b200.objectives import rastrigin
    from evotorch_b200.optimizers import ClipUp
    from evotorch_b200.tools import modify_tensor

    # NCCL prints its version banner to STDOUT at NCCL_DEBUG=VERSION (set in some images): keep stdout = the one JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, D, K, W = args.popsize, args.dim, args.steps, args.warmup

    def barrier_sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    problem = Problem("min", rastrigin, initial_bounds=(-5.12, 5.12), solution_length=D, device=dev, seed=SEED)
    collectives, px = "nccl", None
    if world > 1 and (args.peer == 1 or args.peer < 0):
        try:
            from evotorch_b200.peer import enable_peer_exchange

            px = enable_peer_exchange(problem, N)
            collectives = "peer"
        except Exception as exc:  # e.g. CUDA IPC not permitted in this container: keep the NCCL collectives (still the GPU path)
            print(f"[bench] peer exchange unavailable ({exc!r}); using NCCL collectives", file=sys.stderr)
    searcher = PGPE(problem, popsize=N, center_learning_rate=LR_MU, stdev_learning_rate=LR_SIGMA, stdev_init=STDEV_INIT,
                    distributed=(world > 1))
    use_graph = (world > 1) if args.cuda_graph < 0 else args.cuda_graph == 1
    if use_graph:
        searcher.enable_cuda_graph()
    for _ in range(max(W, 3)):
        searcher.step()

    # ---- device-resident timing (value) + live per-kernel timing (roofline)
    clocks = ClockSampler(local_rank) if rank == 0 else None  # started (and warmed up) BEFORE the barrier: rank 0 must not enter late
    barrier_sync()
    launches0 = ops.launch_count()
    ops.enable_timers()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.perf_counter()
    ev0.record()
    for _ in range(K):
        searcher.step()
    ev1.record()
    barrier_sync()
    t_end = time.perf_counter()
    elapsed_ms = max_over_ranks(ev0.elapsed_time(ev1))
    timers = ops.timer_results()
    ops.disable_timers()
    launches = ops.launch_count() - launches0
    clock_info = clocks.stop(t_begin, t_end) if clocks is not None else None
    mean_eval = float(searcher.status["mean_eval"])
    value = K / (elapsed_ms / 1e3)

    # ---- roofline of the dominant kernel (fused sample + evaluate): algorithmic bytes = the population written once
    n_local = N // world
    peak, peak_src = measured_peak_gbs()
    kern = {}
    for name, (cnt, ms) in timers.items():
        kern[name] = {"launches_timed": cnt, "ms": ms}
    if "sample_eval" not in timers:  # CUDA-graph mode: kernels are not individually timed; time the fused kernel on its own
        pop = searcher._population if searcher._population is not None else next(iter(problem._grad_batches.values()))
        d0 = searcher._distribution
        ops.enable_timers()
        for _ in range(5):
            ops.sample_eval(problem.evok_objective_id, pop._data, d0.mu, d0.sigma, n_rows=len(pop), symmetric=True, seed=1, stream_id=12345,
                            f=pop._evdata.view(-1))
        torch.cuda.synchronize()
        timers = dict(timers, **ops.timer_results())
        ops.disable_timers()
        kern = {name: {"launches_timed": cnt, "ms": ms, "note": "timed stand-alone after the run (CUDA-graph mode)"} for name, (cnt, ms) in timers.items()}
    fused_ms = timers["sample_eval"][1]
    fused_bytes = 4.0 * n_local * D + 4.0 * n_local
    achieved = fused_bytes / (fused_ms * 1e-3) / 1e9
    # DRAM traffic of the kernel from the committed ncu --set full capture (profiles/traffic.json), scaled by rows x columns
    traffic, traffic_note = None, "no capture"
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            cap = json.load(fh)["sample_eval"]
        traffic = cap["ratio"] * fused_bytes
        traffic_note = (f"dram__bytes_read+write = {cap['ratio']:.4f} x algorithmic bytes in {cap['source']} "
                        f"(captured at popsize {cap['capture']['popsize']}, scaled linearly to this launch)")
    except Exception:
        pass
    roofline = {"kernel": "evok::sample_eval_kernel<RASTRIGIN, symmetric, store, vec4>", "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": fused_bytes, "ms_per_launch": fused_ms,
                "timing": ("CUDA events around every launch inside the timed region" if not use_graph else
                           "generations replayed from a CUDA graph: the kernel was timed stand-alone (5 launches, CUDA events) right after the timed region"),
                "share_of_step": fused_ms / (elapsed_ms / K)}
    if "grad" in timers:
        g_ms = timers["grad"][1]
        g_bytes = 4.0 * (n_local // 2) * D
        kern["grad"].update({"algorithmic_bytes": g_bytes, "achieved_gbs": g_bytes / (g_ms * 1e-3) / 1e9,
                             "frac": g_bytes / (g_ms * 1e-3) / 1e9 / peak})
    model_bytes = 10.0 * n_local * D  # write X + read X (evaluate) + read the + rows (gradient)
    traffic_bytes = 6.0 * n_local * D  # what this engine actually moves: evaluation is fused into the write

    # ---- end to end: host-resident distribution -> device generation -> gradients back to the host, every step
    e2e = None
    if not args.no_e2e:
        del searcher
        torch.cuda.empty_cache()
        mu_host = torch.empty(D, dtype=torch.float32).pin_memory()
        sigma_host = torch.empty(D, dtype=torch.float32).pin_memory()
        mu_host.copy_(torch.empty(D).uniform_(-5.12, 5.12, generator=torch.Generator().manual_seed(SEED)))
        sigma_host.fill_(STDEV_INIT)
        hdist = SymmetricSeparableGaussian({"mu": mu_host, "sigma": sigma_host, "divide_mu_grad_by": "num_directions",
                                            "divide_sigma_grad_by": "num_directions"})
        assert hdist.mu.data_ptr() == mu_host.data_ptr() and hdist.mu.is_pinned()  # the distribution lives in the pinned buffers
        hopt = ClipUp(solution_length=D, dtype=torch.float32, stepsize=LR_MU, device="cpu")

        def e2e_step():
            # H2D: mu, sigma (pinned) -> device inside sample_and_compute_gradients; D2H: gradients + mean fitness
            res = problem.sample_and_compute_gradients(hdist, N, ranking_method="centered")[0]
            _ = float(res["mean_eval"])
            upd = hdist.update_parameters(res["gradients"], learning_rates={"sigma": LR_SIGMA}, optimizers={"mu": hopt})
            new_sigma = modify_tensor(sigma_host, upd.sigma, max_change=0.2)
            mu_host.copy_(upd.mu)
            sigma_host.copy_(new_sigma)

        for _ in range(max(W, 3)):
            e2e_step()
        barrier_sync()
        t0 = time.perf_counter()
        for _ in range(K):
            e2e_step()
        barrier_sync()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        h2d = 2 * D * 4
        d2h = 2 * D * 4 + 4
        e2e = {"value": K / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e2e_ms / K,
               "api": "Problem.sample_and_compute_gradients(host-resident SymmetricSeparableGaussian) + update_parameters/modify_tensor on the host"}

    def finish():
        # leave without tearing the NCCL communicators down (teardown after graph-captured collectives can hang); every rank
        # has passed the final barrier and rank 0 has flushed its JSON line
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            os._exit(0)

    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    if rank != 0:
        finish()
        return

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(W, 3), "ms_per_step": elapsed_ms / K,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args, world, collectives), cuda_graph=bool(use_graph)), "impl": "ours",
        "gpu_launches": int(launches), "clocks": clock_info, "e2e": e2e, "roofline": roofline, "kernels": kern,
        "whole_generation": {"model_bytes_per_gen_per_gpu": model_bytes, "model_gbs": model_bytes * value / 1e9,
                             "model_frac_of_peak": model_bytes * value / 1e9 / peak,
                             "moved_bytes_per_gen_per_gpu": traffic_bytes, "moved_gbs": traffic_bytes * value / 1e9,
                             "note": "model = SURVEY 8(d) 10*N*D bytes (unfused write+read+half read); moved = 6*N*D (evaluation fused into the sampling write)"},
        "mean_eval_after": mean_eval,
    }
    if px is not None:
        if px.timed_out():
            raise RuntimeError("a peer-exchange wait timed out during the run: the numbers above are invalid")
        line["peer_exchange"] = {"wait_timeouts": 0, "buffer_bytes": px.nbytes}
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_reference_run(args, steps=3, warmup=1)
    emit(line)
    finish()


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the path (its torch-CPU op sequence, restated in
    oracle/ref_cpu_path.py and checked bit-identical against the real reference in the build container), all host threads,
    same metric / config.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    n_gpus = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    base = cpu_reference_run(args, steps=max(1, min(args.steps, 5)), warmup=min(max(args.warmup, 1), 2))
    line = {
        "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 / base["value"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, n_gpus), "impl": "reference",
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_RESULT_FD = None


def claim_stdout():
    """stdout must carry exactly ONE JSON line, but libraries write there too (NCCL prints its version banner to stdout when
    the image sets NCCL_DEBUG): keep a private duplicate of the real stdout for the result and point fd 1 at stderr."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


if __name__ == "__main__":
    a = parse_args()
    claim_stdout()
    try:
        if a.impl == "reference":
            run_reference(a)
        else:
            run_ours(a)
    except BaseException:
        import traceback

        traceback.print_exc()
        sys.stderr.flush()
        raise
