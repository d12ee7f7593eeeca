"""One short launch sequence of a named config, for ncu (tools/profile_one.py <cfg> [steps] [launches])."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybitup_b200 import synthetic                     # noqa: E402
from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig  # noqa: E402

CFG = {
    "cubic": (lambda: synthetic.cubic_problem(), "AMH", 65536, dict(updating_it=10, eps_v=1e-10)),
    "linear": (lambda: synthetic.linear_problem(), "RWMH", 1 << 20, {}),
    "arrhenius": (lambda: synthetic.arrhenius_problem(), "DRAM", 262144, dict(updating_it=10, eps_v=1e-10, gamma_dr=0.2)),
    "heat": (lambda: synthetic.heat_problem(), "DRAM", 65536, dict(updating_it=10, eps_v=1e-10, gamma_dr=0.2)),
    "isde": (lambda: synthetic.regression_problem(), "ISDE", 131072, dict(isde_h=0.05, isde_f0=4.0, isde_gradient="Analytical")),
}


def main():
    name = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    launches = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    lanes = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    build, algo, C, kw = CFG[name]
    p = build()
    eng = ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo=algo, seed=1, lanes_per_chain=lanes, **kw), C)
    import numpy as np
    eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=1e-3), np.eye(len(p["x0"])) * 2.5e-5 if algo == "ISDE" else p["prop_cov"])
    for _ in range(launches):
        eng.run(steps, thin=steps, keep_samples=False)
    eng.synchronize()
    print("done", name, eng.iteration)


if __name__ == "__main__":
    main()
