import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybitup_b200 import synthetic
from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig
p = synthetic.cubic_problem()
C = 65536
os.environ["PBU_FORCE_VARIANT"] = "4,4,1,1"
for name, kw in (("RWMH", dict(algo="RWMH")), ("AM literal u=10", dict(algo="AMH", updating_it=10)), ("AM literal u=1e6", dict(algo="AMH", updating_it=1000000)),
                 ("AM welford u=10", dict(algo="AMH", updating_it=10, am_welford=True)), ("AM welford u=1e6", dict(algo="AMH", updating_it=1000000, am_welford=True))):
    cfg = SamplerConfig(seed=1, eps_v=1e-10, **kw)
    eng = ChainEngine(ProblemSpec.from_dict(p), cfg, C)
    eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=1e-3), p["prop_cov"])
    ms = []
    for i in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.run(200, thin=200, keep_samples=False); e1.record(); e1.synchronize()
        ms.append(e0.elapsed_time(e1) / 2)
    print("%-20s ms/100 %.3f" % (name, min(ms)))
    eng.close()
