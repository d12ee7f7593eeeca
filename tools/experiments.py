"""Measurements behind the design decisions recorded in DESIGN.md (run on a B200: python tools/experiments.py <name>).

  conflicts   Arrhenius kernels: what do the shared-memory bank conflicts of the table lookups cost?  Identical samples /
              chains (every lane reads the same table entry: a broadcast) against dispersed ones.
  lanes       cfg 3 (Arrhenius DRAM): one lane per chain against the speculative pair (both delayed-rejection stages
              evaluated at once), over chain counts and block sizes.
  rep         cfg 3: plain lookup tables against the conflict-free replicas (192 KB of shared memory, one block per SM), over
              chain counts and block sizes.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                                                           # noqa: E402
import torch                                                                 # noqa: E402
from pybitup_b200 import synthetic                                           # noqa: E402
from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig      # noqa: E402
from pybitup_b200.propagation import DevicePropagation                      # noqa: E402


def timed(fn, reps=3):
    ms = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def conflicts():
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
            x0 = np.minimum(np.maximum(synthetic.chain_starts(p["x0"], C, jitter=jit, seed=0), p["lb"]), p["ub"])
            eng.init_chains(x0, p["prop_cov"])
            eng.run(T, thin=T, keep_samples=False)
            ms = timed(lambda: eng.run(T, thin=T, keep_samples=False))
            print("cfg3 jitter=%g %s: %.2f ms  %.3e chain-steps/s" % (jit, eng.launch_geometry(), ms, C * T / ms * 1e3), flush=True)


def lanes():
    p = synthetic.arrhenius_problem()
    T = 20
    for C in (262144, 131072, 65536, 32768, 8192):
        for L in (0, 1, 2):
            for B in (0,) if L == 0 else (0, 128, 256, 512):
                cfg = SamplerConfig(algo="DRAM", seed=1, updating_it=10, eps_v=1e-10, gamma_dr=0.2, block_threads=B, lanes_per_chain=L)
                with ChainEngine(ProblemSpec.from_dict(p), cfg, C) as eng:
                    eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=1e-3, seed=0), p["prop_cov"])
                    eng.run(T, thin=T, keep_samples=False)
                    ms = timed(lambda: eng.run(T, thin=T, keep_samples=False))
                    print("C=%d lanes=%d block=%d %s: %.2f ms  %.4e chain-steps/s" % (C, L, B, eng.launch_geometry(), ms, C * T / ms * 1e3), flush=True)


def rep():
    p = synthetic.arrhenius_problem()
    T = 20
    for C in (262144, 131072, 65536, 32768, 8192):
        for force, B in (("2,1,0,0,0", 0), ("2,1,0,0,1", 0), ("2,1,0,0,1", 1024), ("2,1,0,0,1", 896), ("2,1,0,0,1", 768), ("2,1,0,0,1", 512), ("", 0)):
            if force:
                os.environ["PBU_FORCE_VARIANT"] = force
            else:
                os.environ.pop("PBU_FORCE_VARIANT", None)
            cfg = SamplerConfig(algo="DRAM", seed=1, updating_it=10, eps_v=1e-10, gamma_dr=0.2, block_threads=B)
            with ChainEngine(ProblemSpec.from_dict(p), cfg, C) as eng:
                eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=1e-3, seed=0), p["prop_cov"])
                eng.run(T, thin=T, keep_samples=False)
                ms = timed(lambda: eng.run(T, thin=T, keep_samples=False))
                print("C=%d force=[%s] block=%d %s: %.2f ms  %.4e chain-steps/s" % (C, force, B, eng.launch_geometry(), ms, C * T / ms * 1e3), flush=True)


if __name__ == "__main__":
    {"conflicts": conflicts, "lanes": lanes, "rep": rep}[sys.argv[1]]()
