"""cfg 1b (linear RWMH, 1,048,576 chains) against the launch geometry of the step kernel."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybitup_b200 import synthetic                                          # noqa: E402
from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig      # noqa: E402
from tools.bench_configs import timed_runs                                  # noqa: E402

p = synthetic.linear_problem()
C = 1 << 20
for lanes in (0, 1, 2):
    for blk in (0, 32, 64, 128, 256, 512):
        try:
            cfg = SamplerConfig(algo="RWMH", seed=1, block_threads=blk, lanes_per_chain=lanes)
            with ChainEngine(ProblemSpec.from_dict(p), cfg, C) as eng:
                eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=1e-3), p["prop_cov"])
                ms = timed_runs(eng, 500, 3)
                print(json.dumps({"lanes": lanes, "block_threads": blk, "ms_per_500": ms, "geom": eng.launch_geometry()}), flush=True)
        except Exception as ex:
            print(json.dumps({"lanes": lanes, "block_threads": blk, "error": str(ex)[:100]}), flush=True)
