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
for C, L in ((262144, 0), (262144, 1), (262144, 2), (131072, 0), (131072, 1), (131072, 2), (65536, 0), (65536, 1), (65536, 2), (32768, 0), (32768, 1), (32768, 2), (8192, 0), (8192, 1), (8192, 2)):
    for B in (0,) if L == 0 else (0, 128, 256, 512):
        cfg = SamplerConfig(algo="DRAM", seed=1, updating_it=10, eps_v=1e-10, gamma_dr=0.2, block_threads=B, lanes_per_chain=L)
        try:
            with ChainEngine(ProblemSpec.from_dict(p), cfg, C) as eng:
                eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=1e-3, seed=0), p["prop_cov"])
                eng.run(T, thin=T, keep_samples=False)
                ms = timed(lambda: eng.run(T, thin=T, keep_samples=False))
                print("C=%d lanes=%d block=%d %s: %.2f ms  %.4e chain-steps/s" % (C, L, B, eng.launch_geometry(), ms, C * T / ms * 1e3), flush=True)
        except Exception as e:
            print("C=%d block=%d failed: %s" % (C, B, e))
