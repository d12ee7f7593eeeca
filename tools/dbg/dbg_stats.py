import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pybitup_b200 import synthetic
from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig

a0 = synthetic.arrhenius_problem(n_steps=40, dT=15.0)
unc = np.array([0, 1])
base = dict(a0, unc_idx=unc, lb=a0["lb"][:2], ub=a0["ub"][:2], x0=a0["x0"][:2], prop_cov=a0["prop_cov"][:2, :2])
dg_b = dict(T0=350.0, dT=10.0, n_steps=50, beta=30.0 / 60.0)
fb = synthetic.arrhenius_reference(a0["theta_nom"], dg_b["T0"], dg_b["dT"], dg_b["n_steps"], dg_b["beta"])
two = dict(base, y=np.concatenate([a0["y"], fb[None, :]], axis=1), std_y=np.full(90, 5e-3),
           blocks=[dict(theta_nom=a0["theta_nom"], unc_idx=unc, design=a0["design"], n_points=40),
                   dict(theta_nom=a0["theta_nom"], unc_idx=unc, design=dg_b, n_points=50)])
for name, p in (("single", base), ("two", two)):
    for algo in ("RWMH", "DRAM"):
        C = 64
        with ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo=algo, seed=9, updating_it=20, eps_v=1e-12, gamma_dr=0.2), C) as eng:
            eng.init_chains(np.tile(p["x0"], (C, 1)), p["prop_cov"])
            for _ in range(2):
                res = eng.run(300, thin=1)
            eng.synchronize()
            xs = res.samples.cpu().numpy()
            st = eng.statistics()
            x, lp = eng.state()
            print(name, algo, "state mean", x.mean(axis=0), "samples mean", xs.mean(axis=(0, 2)), "stat mean", st["mean"], "n", st["n"], st["n_chains"],
                  "rej", eng.counters()["n_rejected"].mean())
