"""Why is the bench workload slower than the probe?  Vary start jitter, iterations per launch, keep_samples."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybitup_b200 import synthetic
from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig

p = synthetic.cubic_problem()
C = 65536
for jitter, T, keep, seed in ((1e-3, 100, False, 0), (1e-2, 100, False, 0), (1e-2, 100, True, 0), (1e-3, 1000, False, 0), (1e-2, 1000, True, 1000), (0.0, 100, False, 0)):
    cfg = SamplerConfig(algo="AMH", seed=1, updating_it=10, eps_v=1e-10)
    eng = ChainEngine(ProblemSpec.from_dict(p), cfg, C)
    eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=jitter, seed=seed), p["prop_cov"])
    ms = []
    for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.run(T, thin=T, keep_samples=keep); e1.record(); e1.synchronize()
        ms.append(e0.elapsed_time(e1) * 100.0 / T)
    ctr = eng.counters()
    print(json.dumps({"jitter": jitter, "T": T, "keep": keep, "ms_per_100": [round(m, 3) for m in ms],
                      "acc": 1 - ctr["n_rejected"].mean() / eng.iteration, "geom": eng.launch_geometry()}))
    eng.close()
