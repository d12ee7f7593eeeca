"""CPU (numpy) forward models used by the oracle.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this module; the product path (``pybitup_b200``) never does.

Parity status
-------------
* ``heat_conduction``  follows the reference's own test model
  (``pybitup/test/test_pce/heat_conduction.py:14-41``) and is pinned by the
  reference's fixtures (``tests/golden/heat_*``).
* ``linear_design``, ``poly_horner`` and ``arrhenius_rk4`` are NOT in the
  reference (``documentation/pybitup_doc.tex:280`` says no models are shipped;
  the examples live in the un-vendored jcoheur/pybitup-examples repo).  Their
  arithmetic is defined HERE and is therefore "parity unpinned" as model
  arithmetic; what IS pinned is the sampler/likelihood arithmetic wrapped
  around them, because the golden vectors were produced by running the
  unmodified reference samplers on ``bi.Model`` subclasses that call exactly
  these functions (``tests/golden/gen_golden.py``).

Every model maps a FULL parameter vector ``theta`` (length P) and the model's
design data to a float64 vector of N evaluations, as ``Model.fun_x`` does in
the reference (``pybitup/bayesian_inference.py:306-308, 371``).
"""
import math

import numpy as np

R_GAS = 8.314462618  # J/(mol K)

MODEL_IDS = {"linear": 0, "poly": 1, "arrhenius": 2, "heat": 3}


# --------------------------------------------------------------------------- linear
def linear_design(theta, Phi):
    """f_k = sum_j Phi[k, j] * theta[j], accumulated j = 0..P-1 in order."""
    theta = np.asarray(theta, dtype=np.float64)
    f = np.zeros(Phi.shape[0])
    for j in range(Phi.shape[1]):
        f = f + Phi[:, j] * theta[j]
    return f


def linear_design_grad(theta, Phi):
    """d f_k / d theta_j = Phi[k, j]  ->  dict j -> f64[N]."""
    return {j: Phi[:, j].copy() for j in range(Phi.shape[1])}


# --------------------------------------------------------------------------- polynomial
def poly_horner(theta, x):
    """f = theta[0] + theta[1] x + ... + theta[P-1] x^(P-1), by Horner from the top."""
    theta = np.asarray(theta, dtype=np.float64)
    f = np.full_like(x, theta[-1], dtype=np.float64)
    for j in range(len(theta) - 2, -1, -1):
        f = f * x + theta[j]
    return f


def poly_grad(theta, x):
    """d f / d theta_j = x^j, built by repeated multiplication."""
    out = {}
    p = np.ones_like(x, dtype=np.float64)
    for j in range(len(theta)):
        out[j] = p.copy()
        p = p * x
    return out


# --------------------------------------------------------------------------- Arrhenius / RK4
def arrhenius_tables(T0, dT, n_steps):
    """1/T at the mid point and at the end point of every RK4 step (shared by all chains)."""
    k = np.arange(n_steps, dtype=np.float64)
    inv_mid = 1.0 / (T0 + (k + 0.5) * dT)
    inv_end = 1.0 / (T0 + (k + 1.0) * dT)
    return inv_mid, inv_end


def arrhenius_rk4(theta, T0, dT, n_steps, beta):
    """Single-reaction pyrolysis, d(alpha)/dT = (A/beta) exp(-E/(R T)) (1-alpha)^n.

    theta = (log10 A [1/s], E [J/mol], n, F);  alpha(T0) = 0;  classic RK4 with
    step dT over n_steps steps;  output y_k = 1 - F*alpha_k at nodes k = 1..n_steps.
    (1-alpha) is clamped at 0 before the power.  The rate constant is evaluated as
    exp(c0 - (E/R) * (1/T)) with c0 = ln(10)*log10A - ln(beta) and 1/T tabulated.
    """
    log10A, E, n, F = (float(v) for v in theta)
    inv_mid, inv_end = arrhenius_tables(T0, dT, n_steps)
    c0 = math.log(10.0) * log10A - math.log(beta)
    ER = E / R_GAS
    h = float(dT)
    out = np.empty(n_steps)
    alpha = 0.0
    g0 = math.exp(c0 - ER * (1.0 / T0))

    def pw(u):
        return math.pow(u, n) if u > 0.0 else 0.0

    for k in range(n_steps):
        gm = math.exp(c0 - ER * inv_mid[k])
        g1 = math.exp(c0 - ER * inv_end[k])
        k1 = h * g0 * pw(1.0 - alpha)
        k2 = h * gm * pw(1.0 - (alpha + 0.5 * k1))
        k3 = h * gm * pw(1.0 - (alpha + 0.5 * k2))
        k4 = h * g1 * pw(1.0 - (alpha + k3))
        alpha = alpha + (k1 + 2.0 * k2 + 2.0 * k3 + k4) / 6.0
        out[k] = 1.0 - F * alpha
        g0 = g1
    return out


# --------------------------------------------------------------------------- heat conduction
def heat_conduction(theta, x, L=70.0):
    """Steady 1-D fin temperature; restates pybitup/test/test_pce/heat_conduction.py:27-41.

    theta = (a, b, k, T_amb, phi, h).
    """
    a, b, k, T_amb, phi, h = (float(v) for v in theta)
    gamma = math.sqrt(2.0 * (a + b) * h / (a * b * k))
    ep = math.exp(gamma * L)
    em = math.exp(-gamma * L)
    c1 = -phi / (k * gamma) * (ep * (h + k * gamma) / (em * (h - k * gamma) + ep * (h + k * gamma)))
    c2 = phi / (k * gamma) + c1
    return c1 * np.exp(-gamma * x) + c2 * np.exp(gamma * x) + T_amb
