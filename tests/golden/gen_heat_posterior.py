"""Posterior moments of the reference's heat-conduction case by brute-force quadrature (401 x 401 grid over
phi in [-19.4, -17.4], h in [1.80e-3, 2.02e-3]) of the oracle's Gaussian likelihood with the uniform prior.
Printed values are hard-coded in tests/test_gpu_dropin.py as the statistical known answer:

    mean = (-18.418010706591474, 0.0019146479526173403)
    var  = (0.023064064360743076, 2.309329769136012e-10), cov = -2.2297224870054396e-06

(The reference's own saved chain, test/test_pce/mcmc_chain.csv, has mean (-18.365, 1.9091e-3): a single DRAM chain of
1e4 iterations with an effective sample size of ~146 by batch means, i.e. a Monte-Carlo standard error of 0.0165 on
phi; it sits 3 standard errors from the quadrature value.)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cpu_models          # noqa: E402
from pybitup_b200 import synthetic     # noqa: E402

p = synthetic.heat_problem()
x, y, s = p["design"]["x"], p["y"][0], p["std_y"]
phis, hs = np.linspace(-19.4, -17.4, 401), np.linspace(0.00180, 0.00202, 401)
L = np.zeros((401, 401))
for i, ph in enumerate(phis):
    for j, h in enumerate(hs):
        f = cpu_models.heat_conduction(np.array([0.95, 0.95, 2.37, 21.29, ph, h]), x)
        L[i, j] = -0.5 * np.sum(((y - f) / s) ** 2)
W = np.exp(L - L.max())
W /= W.sum()
mp, mh = (W.sum(1) * phis).sum(), (W.sum(0) * hs).sum()
print("mean", mp, mh)
print("var", (W.sum(1) * (phis - mp) ** 2).sum(), (W.sum(0) * (hs - mh) ** 2).sum(), "cov", (W * np.outer(phis - mp, hs - mh)).sum())
