"""Exploratory timing of the step kernel on one GPU (not the graded bench; see bench.py)."""
import json
import sys
import os
import time

LAST_GEOM = None

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybitup_b200 import synthetic, _capi            # noqa: E402
from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig  # noqa: E402


def time_run(p, algo, C, steps, lanes=0, block=0, reps=3, q=0, **kw):
    cfg = SamplerConfig(algo=algo, lanes_per_chain=lanes, block_threads=block, chains_per_thread=q, seed=1, **kw)
    eng = ChainEngine(ProblemSpec.from_dict(p), cfg, C)
    eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=1e-3), p["prop_cov"])
    eng.run(steps, thin=steps, keep_samples=False)      # warm-up
    eng.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.run(steps, thin=steps, keep_samples=False)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    global LAST_GEOM
    LAST_GEOM = eng.launch_geometry()
    ctr = eng.counters()
    acc = 1.0 - ctr["n_rejected"].mean() / eng.iteration
    evals = ctr["n_eval"].mean() / eng.iteration
    eng.close()
    return best, acc, evals


def main():
    tf, ms = _capi.measure_fp64_peak(0, 4096, 5)
    print(json.dumps({"fp64_peak_tflops": tf, "ms": ms}))
    which = sys.argv[1:] or ["cubic", "linear", "arrhenius", "heat"]
    if "cubic" in which:
        p = synthetic.cubic_problem()
        flops = 1000 * 10 + 20 + 140
        for lanes, block, q in ((0, 0, 0), (1, 448, 1), (1, 224, 2), (2, 448, 2), (2, 224, 2), (2, 512, 2), (4, 448, 2), (4, 512, 2)):
            ms, acc, ev = time_run(p, "AMH", 65536, 100, lanes, block, q=q, updating_it=10, eps_v=1e-10)
            rate = 65536 * 100 / (ms * 1e-3)
            print(json.dumps({"cfg": "cubic_am_65536", "lanes": lanes, "block": block, "geom": LAST_GEOM, "ms": ms, "steps_per_s": rate,
                              "tflops_alg": rate * flops / 1e12, "frac": rate * flops / 1e12 / tf, "acc": acc}))
        for algo in ("RWMH", "DR", "DRAM"):
            ms, acc, ev = time_run(p, algo, 65536, 100, 0, 0, updating_it=10, eps_v=1e-10, gamma_dr=0.2)
            rate = 65536 * 100 / (ms * 1e-3)
            print(json.dumps({"cfg": "cubic_%s_65536" % algo, "ms": ms, "steps_per_s": rate, "evals_per_step": ev,
                              "tflops_alg": rate * ev * flops / 1e12, "acc": acc}))
    if "scan" in which:
        # time per launch vs N: slope = data loop, intercept = per-iteration scalar part
        for algo in ("RWMH", "AMH"):
            for q in (1, 2):
                for n in (250, 500, 1000, 2000, 4000):
                    p = synthetic.cubic_problem(n_points=n)
                    ms, acc, ev = time_run(p, algo, 65536, 100, 1, 0, q=q, updating_it=10, eps_v=1e-10)
                    print(json.dumps({"cfg": "scan", "algo": algo, "q": q, "N": n, "geom": LAST_GEOM, "ms": ms,
                                      "us_per_step_per_launch": ms * 10.0}))
    if "linear" in which:
        p = synthetic.linear_problem()
        for C in (1, 148 * 128, 1 << 20):
            for lanes, q in (((0, 0), (8, 1)) if C == 1 else ((0, 0), (1, 1), (1, 2))):
                steps = 20000 if C == 1 else 200
                ms, acc, ev = time_run(p, "RWMH", C, steps, lanes, 0, q=q)
                rate = C * steps / (ms * 1e-3)
                print(json.dumps({"cfg": "linear_rwmh", "C": C, "geom": LAST_GEOM, "ms": ms, "steps_per_s": rate,
                                  "ns_per_step": ms * 1e6 / steps, "tflops_alg": rate * 512 / 1e12, "acc": acc}))
    if "arrhenius" in which:
        p = synthetic.arrhenius_problem()
        for block in (0, 128):
            ms, acc, ev = time_run(p, "DRAM", 262144, 4, 1, block, updating_it=10, eps_v=1e-10, gamma_dr=0.2, reps=2)
            rate = 262144 * 4 / (ms * 1e-3)
            print(json.dumps({"cfg": "arrhenius_dram_262144", "block": block, "ms": ms, "steps_per_s": rate,
                              "evals_per_step": ev, "tflops_alg": rate * ev * 2.17e4 / 1e12, "acc": acc}))
    if "heat" in which:
        p = synthetic.heat_problem()
        ms, acc, ev = time_run(p, "DRAM", 65536, 200, 0, 0, updating_it=10, eps_v=1e-10, gamma_dr=0.2)
        print(json.dumps({"cfg": "heat_dram_65536", "ms": ms, "steps_per_s": 65536 * 200 / (ms * 1e-3), "acc": acc}))


if __name__ == "__main__":
    main()
