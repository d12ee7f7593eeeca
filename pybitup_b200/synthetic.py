"""Synthetic inverse problems of the shapes BASELINE.json names.

Plain numpy, fixed seeds.  Each builder returns a dict of arrays that both the
CUDA engine (``pybitup_b200.problem.ProblemSpec.from_dict``) and the test
oracle understand:

    model, theta_nom[P], unc_idx[d], y[n_runs,N], std_y[N], lb[d], ub[d], gamma,
    design{...}, x0[d], prop_cov[d,d], theta_true[P]

The reference ships no forward models (``documentation/pybitup_doc.tex:280``);
these definitions are this repo's and are frozen here.
"""
import numpy as np

R_GAS = 8.314462618


def linear_problem(n_points=50, seed=0):
    """cfg 1: f = th0 + th1 x + th2 x^2 stored as a design matrix, d = 3, N = 50, sigma = 0.1."""
    x = np.linspace(0.0, 1.0, n_points)
    Phi = np.stack([np.ones_like(x), x, x * x], axis=1)
    theta = np.array([1.0, 2.0, 3.0])
    rng = np.random.default_rng(seed)
    y = Phi @ theta + 0.1 * rng.standard_normal(n_points)
    return dict(model="linear", theta_nom=theta.copy(), unc_idx=np.arange(3), y=y[None, :],
                std_y=np.full(n_points, 0.1), lb=np.full(3, -10.0), ub=np.full(3, 10.0), gamma=1.0,
                design=dict(Phi=Phi, x=x), x0=theta.copy(),
                prop_cov=np.diag([1e-3, 1e-2, 1e-2]), theta_true=theta)


def cubic_problem(n_points=1000, seed=0):
    """cfg 2: cubic by Horner, d = 4, N = 1000, sigma = 0.1, x in [-1, 1]."""
    x = np.linspace(-1.0, 1.0, n_points)
    theta = np.array([1.0, -2.0, 0.5, 3.0])
    rng = np.random.default_rng(seed)
    f = ((theta[3] * x + theta[2]) * x + theta[1]) * x + theta[0]
    y = f + 0.1 * rng.standard_normal(n_points)
    return dict(model="poly", theta_nom=theta.copy(), unc_idx=np.arange(4), y=y[None, :],
                std_y=np.full(n_points, 0.1), lb=np.full(4, -10.0), ub=np.full(4, 10.0), gamma=1.0,
                design=dict(x=x), x0=theta.copy(),
                prop_cov=np.diag([1e-4, 1e-3, 1e-3, 1e-3]) * 0.1, theta_true=theta)


def arrhenius_reference(theta, T0, dT, n_steps, beta):
    """Host-side restatement of the RK4 pyrolysis model, used ONLY to synthesise data."""
    log10A, E, n, F = (float(v) for v in theta)
    c0 = np.log(10.0) * log10A - np.log(beta)
    ER = E / R_GAS
    out = np.empty(n_steps)
    alpha = 0.0
    g0 = np.exp(c0 - ER / T0)
    for k in range(n_steps):
        gm = np.exp(c0 - ER / (T0 + (k + 0.5) * dT))
        g1 = np.exp(c0 - ER / (T0 + (k + 1.0) * dT))
        pw = lambda u: u ** n if u > 0.0 else 0.0
        k1 = dT * g0 * pw(1.0 - alpha)
        k2 = dT * gm * pw(1.0 - (alpha + 0.5 * k1))
        k3 = dT * gm * pw(1.0 - (alpha + 0.5 * k2))
        k4 = dT * g1 * pw(1.0 - (alpha + k3))
        alpha = alpha + (k1 + 2.0 * k2 + 2.0 * k3 + k4) / 6.0
        out[k] = 1.0 - F * alpha
        g0 = g1
    return out


def arrhenius_problem(n_steps=500, dT=2.0, seed=0):
    """cfg 3/5: one-reaction pyrolysis, theta = (log10 A, E, n, F), 500 RK4 steps of 2 K from 300 K."""
    T0, beta = 300.0, 10.0 / 60.0
    theta = np.array([6.0, 1.0e5, 1.5, 0.3])
    rng = np.random.default_rng(seed)
    f = arrhenius_reference(theta, T0, dT, n_steps, beta)
    sigma = 5e-3
    y = f + sigma * rng.standard_normal(n_steps)
    x = T0 + dT * np.arange(1, n_steps + 1)
    return dict(model="arrhenius", theta_nom=theta.copy(), unc_idx=np.arange(4), y=y[None, :],
                std_y=np.full(n_steps, sigma), lb=0.7 * theta, ub=1.3 * theta, gamma=1.0,
                design=dict(T0=T0, dT=dT, n_steps=n_steps, beta=beta, x=x), x0=theta.copy(),
                prop_cov=np.diag((2e-4 * theta) ** 2), theta_true=theta)


def regression_problem(n_points=10000, d=10, seed=0):
    """cfg 4: correlated Gaussian regression, Phi rows ~ N(0, Sigma), Sigma_ij = 0.9^|i-j|, noise 0.5."""
    rng = np.random.default_rng(seed)
    idx = np.arange(d)
    Sigma = 0.9 ** np.abs(idx[:, None] - idx[None, :])
    Lc = np.linalg.cholesky(Sigma)
    Phi = rng.standard_normal((n_points, d)) @ Lc.T
    theta = np.linspace(-1.0, 1.0, d)
    y = Phi @ theta + 0.5 * rng.standard_normal(n_points)
    return dict(model="linear", theta_nom=theta.copy(), unc_idx=np.arange(d), y=y[None, :],
                std_y=np.full(n_points, 0.5), lb=np.full(d, -10.0), ub=np.full(d, 10.0), gamma=1.0,
                design=dict(Phi=Phi, x=np.arange(n_points, dtype=np.float64)), x0=theta.copy(),
                prop_cov=np.eye(d) * 1e-4, theta_true=theta)


def heat_problem():
    """The reference's own heat-conduction case (pybitup/test/test_pce/heat_conduction_1.json,
    data from Smith 2013 Tab. 3.2 as listed in test_pce/generateDataFile.py:31-32)."""
    x = np.arange(10.0, 70.0, 4.0)
    y = np.array([96.14, 80.12, 67.66, 57.96, 50.90, 44.84, 39.75, 36.16, 33.31, 31.15,
                  29.28, 27.88, 27.18, 26.40, 25.86])
    theta = np.array([0.95, 0.95, 2.37, 21.29, -18.41, 0.00191])
    return dict(model="heat", theta_nom=theta.copy(), unc_idx=np.array([4, 5]), y=y[None, :],
                std_y=np.full(15, 0.2604), lb=np.array([-100.0, 0.00001]), ub=np.array([0.0, 1.0]),
                gamma=1.0, design=dict(x=x), x0=np.array([-18.41, 0.00191]),
                prop_cov=np.array([[2.1034e-2, -2.0286e-6], [-2.0286e-6, 2.0972e-10]]),
                theta_true=theta)


def chain_starts(x0, n_chains, jitter=0.01, seed=0):
    """Per-chain starting points x0 + jitter*N(0,1)*|x0| (scale 1 where x0 == 0)."""
    rng = np.random.default_rng(seed)
    x0 = np.asarray(x0, dtype=np.float64)
    scale = np.where(x0 == 0.0, 1.0, np.abs(x0))
    return x0[None, :] + jitter * scale[None, :] * rng.standard_normal((n_chains, x0.size))
