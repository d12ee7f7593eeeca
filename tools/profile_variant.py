"""One launch sequence of cubic cfg 2 with a forced kernel variant, for ncu: profile_variant.py <algo> <variant>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
algo, var = sys.argv[1], sys.argv[2]
if var != "auto":
    os.environ["PBU_FORCE_VARIANT"] = var
from pybitup_b200 import synthetic
from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig
p = synthetic.cubic_problem()
eng = ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo=algo, seed=1, updating_it=10, eps_v=1e-10, gamma_dr=0.2), 65536)
eng.init_chains(synthetic.chain_starts(p["x0"], 65536, jitter=1e-3), p["prop_cov"])
for _ in range(3):
    eng.run(100, thin=100, keep_samples=False)
eng.synchronize()
print("done", eng.launch_geometry())
