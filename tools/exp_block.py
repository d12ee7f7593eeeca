"""cfg 3 (Arrhenius DRAM, 262,144 chains): block size against the wave quantisation (1024 resident threads per SM)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pybitup_b200 import synthetic
from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig

def timed(fn, reps=3):
    ms = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))

p = synthetic.arrhenius_problem()
T = 20
for C in (262144, 32768):
    for B in (0, 64, 96, 128, 160, 192, 224, 256, 320, 448, 512):
        cfg = SamplerConfig(algo="DRAM", seed=1, updating_it=10, eps_v=1e-10, gamma_dr=0.2, block_threads=B)
        try:
            with ChainEngine(ProblemSpec.from_dict(p), cfg, C) as eng:
                eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=1e-3, seed=0), p["prop_cov"])
                eng.run(T, thin=T, keep_samples=False)
                ms = timed(lambda: eng.run(T, thin=T, keep_samples=False))
                print("C=%d block=%d %s: %.2f ms  %.4e chain-steps/s" % (C, B, eng.launch_geometry(), ms, C * T / ms * 1e3), flush=True)
        except Exception as e:
            print("C=%d block=%d failed: %s" % (C, B, e))
