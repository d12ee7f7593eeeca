"""cfg 3 (Arrhenius DRAM, 262,144 chains) against the block size of the step kernel (wave quantisation: 4 x 256-thread
blocks per SM hold 592 of the 1024 blocks at a time)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybitup_b200 import synthetic                                          # noqa: E402
from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig      # noqa: E402
from tools.bench_configs import timed_runs                                  # noqa: E402

p = synthetic.arrhenius_problem()
C = 262144
for blk in (0, 64, 128, 192, 224, 256, 288, 320, 512):
    try:
        cfg = SamplerConfig(algo="DRAM", seed=1, updating_it=10, eps_v=1e-10, gamma_dr=0.2, block_threads=blk)
        with ChainEngine(ProblemSpec.from_dict(p), cfg, C) as eng:
            eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=1e-3), p["prop_cov"])
            ms = timed_runs(eng, 20, 3)
            print(json.dumps({"block_threads": blk, "ms_per_20": ms, "geom": eng.launch_geometry()}), flush=True)
    except Exception as ex:
        print(json.dumps({"block_threads": blk, "error": str(ex)[:120]}), flush=True)
