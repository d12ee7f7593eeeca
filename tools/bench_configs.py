"""The five BASELINE.json configurations at full per-GPU size on ONE B200 (the graded metric is bench.py's; this
records the others).  Prints one JSON line per configuration; algorithmic flop counts are SURVEY.md 8(d)'s."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybitup_b200 import synthetic, _capi                                  # noqa: E402
from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig      # noqa: E402
from pybitup_b200.propagation import DevicePropagation                      # noqa: E402


FP64_INST_RK4 = 77           # DFMA + DMUL + DADD in the RK4 loop of mh_step_kernel<ArrheniusModel,...> (cuobjdump -sass)
FP64_INST_RK4_EVAL = 77      # same loop in eval_model_kernel<ArrheniusModel, 4> (no residual / sum of squares)


def timed_runs(eng, steps, reps):
    eng.run(steps, thin=steps, keep_samples=False)
    eng.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.run(steps, thin=steps, keep_samples=False)
        e1.record()
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def main():
    which = sys.argv[1:] or ["1", "2", "3", "4", "5"]
    peak, _ = _capi.measure_fp64_peak(0, 4096, -100)
    print(json.dumps({"fp64_peak_tflops_sustained": peak}))
    if "1" in which:      # linear, d=3, N=50, RWMH, 1 chain x 1e5 iterations: latency-bound, ns per step
        p = synthetic.linear_problem()
        with ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo="RWMH", seed=1), 1) as eng:
            eng.init_chains(p["x0"][None, :], p["prop_cov"])
            ms = timed_runs(eng, 100000, 3)
            print(json.dumps({"cfg": 1, "what": "linear RWMH, 1 chain x 1e5 iterations", "geom": eng.launch_geometry(), "ms": ms,
                              "ns_per_step": ms * 1e6 / 1e5, "chain_steps_per_s": 1e5 / (ms * 1e-3)}))
        C = 1 << 20
        with ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo="RWMH", seed=1), C) as eng:
            eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=1e-3), p["prop_cov"])
            ms = timed_runs(eng, 500, 3)
            r = C * 500 / (ms * 1e-3)
            print(json.dumps({"cfg": "1b", "what": "linear RWMH, 1,048,576 chains x 500 iterations", "geom": eng.launch_geometry(), "ms": ms,
                              "chain_steps_per_s": r, "tflops_alg": r * 512 / 1e12, "frac_fp64": r * 512 / 1e12 / peak}))
    if "2" in which:
        p = synthetic.cubic_problem()
        C = 65536
        for algo in ("AMH", "RWMH", "DR", "DRAM"):
            with ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo=algo, seed=1, updating_it=10, eps_v=1e-10, gamma_dr=0.2), C) as eng:
                eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=1e-2), p["prop_cov"])
                ms = timed_runs(eng, 500, 3)
                ev = eng.counters()["n_eval"].mean() / eng.iteration
                r = C * 500 / (ms * 1e-3)
                print(json.dumps({"cfg": 2, "what": "cubic %s, 65,536 chains x 500 iterations" % algo, "geom": eng.launch_geometry(), "ms": ms,
                                  "chain_steps_per_s": r, "evals_per_step": ev, "tflops_alg": r * ev * 10163 / 1e12,
                                  "frac_fp64": r * ev * 10163 / 1e12 / peak}))
    if "3" in which:
        p = synthetic.arrhenius_problem()
        C = 262144
        with ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo="DRAM", seed=1, updating_it=10, eps_v=1e-10, gamma_dr=0.2), C) as eng:
            eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=1e-3), p["prop_cov"])
            ms = timed_runs(eng, 20, 3)
            ev = eng.counters()["n_eval"].mean() / eng.iteration
            r = C * 20 / (ms * 1e-3)
            print(json.dumps({"cfg": 3, "what": "Arrhenius RK4x500 DRAM, 262,144 chains x 20 iterations", "geom": eng.launch_geometry(), "ms": ms,
                              "chain_steps_per_s": r, "evals_per_step": ev, "tflops_alg": r * ev * 2.17e4 / 1e12,
                              "frac_fp64_alg": r * ev * 2.17e4 / 1e12 / peak,
                              # the kernel's RK4 step is 77 FP64 instructions (cuobjdump -sass; four stages 2^(y + n log2 u) by table-driven base-2
                              # exp / log): issue slots of the FP64 pipe in use, each counted as one DFMA (2 flops)
                              "fp64_inst_per_rk4_step": FP64_INST_RK4, "frac_fp64_issue": r * ev * 500 * FP64_INST_RK4 * 2 / 1e12 / peak}))
    if "4" in which:
        p = synthetic.regression_problem()
        C = 131072           # 1,048,576 chains over 8 GPUs
        cfg = SamplerConfig(algo="ISDE", seed=1, isde_h=0.05, isde_f0=4.0, isde_gradient="Analytical")
        with ChainEngine(ProblemSpec.from_dict(p), cfg, C) as eng:
            eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=1e-2), np.eye(10) * 2.5e-5)
            ms = timed_runs(eng, 10, 3)
            r = C * 10 / (ms * 1e-3)
            st = eng.statistics()
            print(json.dumps({"cfg": 4, "what": "regression d=10 N=1e4 ISDE (streamed records), 131,072 chains x 10 iterations",
                              "geom": eng.launch_geometry(), "ms": ms, "chain_steps_per_s": r, "tflops_alg": r * 4.4e5 / 1e12,
                              "frac_fp64": r * 4.4e5 / 1e12 / peak, "status_flags": int(np.count_nonzero(eng.counters()["status"])),
                              "rhat_max": float(np.nanmax(st["rhat"]))}))
    if "5" in which:
        p = synthetic.arrhenius_problem()
        S = 12500000         # 1e8 samples over 8 GPUs
        with ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo="RWMH"), 1) as eng:
            with DevicePropagation(eng, S) as dp:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                dp.evaluate_gaussian(p["x0"], 0.01 * np.abs(p["x0"]), seed=0, sample_offset=0)
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                st = dp.statistics()
                t2 = time.perf_counter()
                how = dp.last_select
                st = dp.statistics()          # again with the subsample / candidate buffers already allocated
                t3 = time.perf_counter()
                stf = dp.statistics(bracket=False)
                t4 = time.perf_counter()
                assert np.array_equal(st["ci_lb"], stf["ci_lb"]) and np.array_equal(st["ci_ub"], stf["ci_ub"])
                print(json.dumps({"cfg": 5, "what": "Arrhenius propagation, 1.25e7 samples x 500 points, evaluations in HBM (50 GB)",
                                  "eval_s": t1 - t0, "statistics_s": t2 - t1, "statistics_select": how, "statistics_s_second_call": t3 - t2,
                                  "statistics_s_full_radix_select": t4 - t3, "samples_per_s_eval": S / (t1 - t0),
                                  "samples_per_s_total": S / (t2 - t0), "tflops_alg_eval": S * 2.15e4 / (t1 - t0) / 1e12,
                                  "frac_fp64_issue_eval": S * 500 * FP64_INST_RK4_EVAL * 2 / (t1 - t0) / 1e12 / peak,
                                  "hbm_write_GBs": S * 4000 / (t1 - t0) / 1e9, "mean_first_last": [float(st["mean"][0]), float(st["mean"][-1])],
                                  "ci_last": [float(st["ci_lb"][-1]), float(st["ci_ub"][-1])]}))


if __name__ == "__main__":
    main()
