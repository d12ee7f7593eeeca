import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybitup_b200 import synthetic
from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig
p = synthetic.cubic_problem()
C = 65536
for force in ("", "2,2,1,0", "4,4,0,1", "8,8,0,1", "4,4,1,1", "8,8,1,1"):
    for algo in ("AMH", "RWMH", "DRAM"):
        if force:
            os.environ["PBU_FORCE_VARIANT"] = force
        else:
            os.environ.pop("PBU_FORCE_VARIANT", None)
        cfg = SamplerConfig(algo=algo, seed=1, updating_it=10, eps_v=1e-10, gamma_dr=0.2)
        try:
            eng = ChainEngine(ProblemSpec.from_dict(p), cfg, C)
        except Exception as ex:
            print("skip", force, algo, str(ex)[:100]); continue
        eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=1e-3), p["prop_cov"])
        ms = []
        for i in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.run(200, thin=200, keep_samples=False); e1.record(); e1.synchronize()
            ms.append(e0.elapsed_time(e1) / 2)
        ev = eng.counters()["n_eval"].mean() / eng.iteration
        print("variant", force or "auto", algo, eng.launch_geometry(), "ms/100 %.3f" % min(ms), "evals/step %.3f" % ev,
              "TF_alg %.2f" % (C * 100 * ev * 10163 / (min(ms) * 1e-3) / 1e12))
        eng.close()
