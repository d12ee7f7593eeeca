"""Arrhenius kernels: how much do shared-memory bank conflicts of the table lookups cost?  Identical samples / chains
(every lane reads the same table entry: broadcast) against dispersed ones."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pybitup_b200 import synthetic
from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig
from pybitup_b200.propagation import DevicePropagation

def timed(fn, reps=3):
    ms = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))

p = synthetic.arrhenius_problem()
S = 12500000 // 4
with ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo="RWMH"), 1) as eng:
    with DevicePropagation(eng, S) as dp:
        for rel in (0.0, 1e-4, 1e-3, 1e-2, 5e-2):
            std = rel * np.abs(p["x0"])
            dp.evaluate_gaussian(p["x0"], std, seed=0, sample_offset=0)
            ms = timed(lambda: dp.evaluate_gaussian(p["x0"], std, seed=0, sample_offset=0))
            print("eval S=%d rel_std=%g: %.2f ms  %.3e samples/s" % (S, rel, ms, S / ms * 1e3), flush=True)
C, T = 262144, 20
for jit in (0.0, 1e-3, 1e-2, 5e-2):
    cfg = SamplerConfig(algo="DRAM", seed=1, updating_it=10, eps_v=1e-10, gamma_dr=0.2)
    with ChainEngine(ProblemSpec.from_dict(p), cfg, C) as eng:
        x0 = synthetic.chain_starts(p["x0"], C, jitter=jit, seed=0)
        x0 = np.minimum(np.maximum(x0, p["lb"]), p["ub"])
        eng.init_chains(x0, p["prop_cov"])
        eng.run(T, thin=T, keep_samples=False)
        ms = timed(lambda: eng.run(T, thin=T, keep_samples=False))
        ev = float(eng.read_state()["n_eval"].mean()) / eng.iteration
        print("cfg3 jitter=%g: %.2f ms  %.3e chain-steps/s  evals/step %.2f  -> %.3e evals/s" % (jit, ms, C * T / ms * 1e3, ev, C * T * ev / ms * 1e3), flush=True)
