"""BASELINE.json configurations 3, 4 and 5 at their multi-GPU sizes, one rank per GPU (weak scaling: fixed work per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_configs_multi.py [3 4 5]

  cfg 3  Arrhenius RK4 x 500, DRAM, 262,144 chains per GPU (chains sharded by global index, no collective in the run)
  cfg 4  Ito-SDE, d = 10, N = 1e4, 131,072 chains per GPU (1,048,576 over 8) + the NCCL all-reduce of the chain moments
         for the pooled covariance and Gelman-Rubin R-hat
  cfg 5  Monte-Carlo propagation, 1.25e7 samples per GPU (1e8 over 8): evaluation, then statistics with the brackets /
         counts / candidate histograms reduced over NCCL
Timing: CUDA events on the engine's stream between barriers, max over ranks; rank 0 prints one JSON line per config."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybitup_b200 import synthetic                                          # noqa: E402
from pybitup_b200.distributed import Reducer                                # noqa: E402
from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig      # noqa: E402
from pybitup_b200.propagation import DevicePropagation                      # noqa: E402


def max_over_ranks(x, dev):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def timed(fn, dev, reps=3):
    """median over reps of (device ms between events, wall ms), each the max over ranks"""
    out = []
    for _ in range(reps):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        w = (time.perf_counter() - w0) * 1e3
        out.append((max_over_ranks(e0.elapsed_time(e1), dev), max_over_ranks(w, dev)))
    out.sort()
    return out[len(out) // 2]


def main():
    which = sys.argv[1:] or ["3", "4", "5"]
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    red = Reducer()
    say = (lambda d: print(json.dumps(d), flush=True)) if rank == 0 else (lambda d: None)
    if "3" in which:
        p = synthetic.arrhenius_problem()
        C, T = 262144, 20
        cfg = SamplerConfig(algo="DRAM", seed=1, updating_it=10, eps_v=1e-10, gamma_dr=0.2, chain_offset=rank * C)
        with ChainEngine(ProblemSpec.from_dict(p), cfg, C, device=local) as eng:
            eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=1e-3, seed=rank), p["prop_cov"])
            eng.run(T, thin=T, keep_samples=False)
            eng.synchronize()
            ms, _ = timed(lambda: eng.run(T, thin=T, keep_samples=False), dev)
            ev = float(red(np.array([eng.counters()["n_eval"].mean() / eng.iteration]), "sum")[0]) / world
            say({"cfg": 3, "n_gpus": world, "what": "Arrhenius RK4x500 DRAM, 262,144 chains per GPU x %d iterations" % T, "ms": ms,
                 "chain_steps_per_s": world * C * T / (ms * 1e-3), "evals_per_step": ev, "scaling": "weak"})
    if "4" in which:
        p = synthetic.regression_problem()
        C, T = 131072, 10
        cfg = SamplerConfig(algo="ISDE", seed=1, isde_h=0.05, isde_f0=4.0, isde_gradient="Analytical", chain_offset=rank * C)
        with ChainEngine(ProblemSpec.from_dict(p), cfg, C, device=local) as eng:
            eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=1e-2, seed=rank), np.eye(10) * 2.5e-5)
            eng.run(T, thin=T, keep_samples=False)
            eng.synchronize()
            ms, _ = timed(lambda: eng.run(T, thin=T, keep_samples=False), dev)
            st = eng.statistics(red)
            _, ms_stat = timed(lambda: eng.statistics(red), dev)
            say({"cfg": 4, "n_gpus": world, "what": "regression d=10 N=1e4 ISDE, 131,072 chains per GPU x %d iterations" % T, "ms": ms,
                 "chain_steps_per_s": world * C * T / (ms * 1e-3), "chains_total": int(st["n_chains"]),
                 "moment_allreduce_and_rhat_ms": ms_stat, "rhat_max": float(np.nanmax(st["rhat"])), "scaling": "weak"})
    if "5" in which:
        p = synthetic.arrhenius_problem()
        S = 12500000
        with ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo="RWMH"), 1, device=local) as eng:
            with DevicePropagation(eng, S) as dp:
                mean, std = p["x0"], 0.01 * np.abs(p["x0"])
                dp.evaluate_gaussian(mean, std, seed=0, sample_offset=rank * S)
                _, ms_eval = timed(lambda: dp.evaluate_gaussian(mean, std, seed=0, sample_offset=rank * S), dev)
                st = dp.statistics(reduce=red, n_total=world * S)
                how = dp.last_select
                _, ms_stat = timed(lambda: dp.statistics(reduce=red, n_total=world * S), dev)
                say({"cfg": 5, "n_gpus": world, "what": "Arrhenius propagation, 1.25e7 samples per GPU x 500 points", "eval_ms": ms_eval,
                     "statistics_ms": ms_stat, "statistics_select": how, "samples_total": world * S,
                     "samples_per_s_eval": world * S / (ms_eval * 1e-3), "samples_per_s_total": world * S / ((ms_eval + ms_stat) * 1e-3),
                     "mean_first_last": [float(st["mean"][0]), float(st["mean"][-1])], "ci_last": [float(st["ci_lb"][-1]), float(st["ci_ub"][-1])],
                     "scaling": "weak"})
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
