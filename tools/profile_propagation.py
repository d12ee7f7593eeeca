"""cfg 5 shape (evaluation + statistics of S samples on one GPU), for an ncu launch list:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_propagation.py [S]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybitup_b200 import synthetic                                          # noqa: E402
from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig      # noqa: E402
from pybitup_b200.propagation import DevicePropagation                      # noqa: E402

S = int(float(sys.argv[1])) if len(sys.argv) > 1 else 12500000
p = synthetic.arrhenius_problem()
with ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo="RWMH"), 1) as eng:
    with DevicePropagation(eng, S) as dp:
        dp.evaluate_gaussian(p["x0"], 0.01 * np.abs(p["x0"]), seed=0, sample_offset=0)
        st = dp.statistics()
        print(dp.last_select, float(st["ci_lb"][-1]), float(st["ci_ub"][-1]))
