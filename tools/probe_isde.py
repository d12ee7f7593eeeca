import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybitup_b200 import synthetic
from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig
p = synthetic.regression_problem()
C = 131072
for lanes, q, block in ((0, 0, 0), (1, 1, 512), (4, 1, 512), (8, 1, 512)):
    try:
        cfg = SamplerConfig(algo="ISDE", seed=1, isde_h=0.05, isde_f0=4.0, isde_gradient="Analytical", lanes_per_chain=lanes, chains_per_thread=q, block_threads=block)
        eng = ChainEngine(ProblemSpec.from_dict(p), cfg, C)
    except Exception as ex:
        print("skip", lanes, q, block, str(ex)[:80]); continue
    eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=1e-2), np.eye(10) * 2.5e-5)
    ms = []
    for i in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.run(5, thin=5, keep_samples=False); e1.record(); e1.synchronize()
        ms.append(e0.elapsed_time(e1) / 5)
    r = C / (min(ms) * 1e-3)
    print(json.dumps({"want": [lanes, q, block], "geom": eng.launch_geometry(), "ms_per_it": min(ms), "steps_per_s": r, "tflops_alg": r * 4.4e5 / 1e12}))
    eng.close()
