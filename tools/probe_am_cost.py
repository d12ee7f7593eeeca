"""Decomposes the cost of adaptive Metropolis over random-walk MH at the cfg 2 shape: literal vs Welford recursion,
Cholesky refresh every 10 vs (almost) never."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybitup_b200 import synthetic                                          # noqa: E402
from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig      # noqa: E402
from tools.bench_configs import timed_runs                                  # noqa: E402

p = synthetic.cubic_problem()
C = 65536
for name, kw in (("RWMH", dict(algo="RWMH")), ("AM literal, refresh 10", dict(algo="AMH", updating_it=10)),
                 ("AM literal, refresh 100000", dict(algo="AMH", updating_it=100000)),
                 ("AM welford, refresh 10", dict(algo="AMH", updating_it=10, am_welford=True)),
                 ("AM welford, refresh 100000", dict(algo="AMH", updating_it=100000, am_welford=True))):
    with ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(seed=1, eps_v=1e-10, **kw), C) as eng:
        eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=1e-2), p["prop_cov"])
        ms = timed_runs(eng, 500, 3)
        print(json.dumps({"what": name, "ms_per_500": ms, "geom": eng.launch_geometry()}))
