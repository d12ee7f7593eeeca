"""CPU restatement of the reference's MCMC hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this module.  ``pybitup_b200`` (the product) never imports it
and has no CPU fallback.

It restates, in plain numpy / Python floats, the arithmetic of

* ``pybitup/metropolis_hastings_algorithms.py``  (RWMH :158-224, AM :397-487,
  DR :512-567, DRAM :570-595, GradientBasedMCMC :621-777, ito_SDE :911-1040,
  computeGradientFD :1043-1088),
* ``pybitup/bayesian_inference.py``  (BayesianPosterior.compute_log_value
  :44-64, Likelihood.arg_gauss_likelihood :495-542, compute_grad_log_value
  :573-623, Model.run_model :338-371),
* ``pybitup/distributions.py``  (Uniform.compute_value :282-289,
  Mixture.compute_log_value :456-479),

with the random draws INJECTED (a ``Streams`` object standing in for the
stdlib ``random`` module the reference calls at :213, :314, :523, :561).

Pinning: ``tests/test_oracle_golden.py`` checks every function here against
golden vectors recorded from the UNMODIFIED reference run in the build
container (``tests/golden/gen_golden.py``; vectors in ``tests/golden/*.npz``).
The reference's own unit tests assert nothing about chains, so
those recorded runs are the pin.
"""
from dataclasses import dataclass, field
from fractions import Fraction

import numpy as np
import scipy.linalg

from . import cpu_models


# --------------------------------------------------------------------------- problem
@dataclass
class Problem:
    """Flat description of one Bayesian inverse problem (one model, n_runs data sets)."""
    model: str                       # 'linear' | 'poly' | 'arrhenius' | 'heat'
    theta_nom: np.ndarray            # f64[P] nominal full parameter vector (Model.param_nom)
    unc_idx: np.ndarray              # int[d] position of each uncertain parameter in theta
    y: np.ndarray                    # f64[n_runs, N]
    std_y: np.ndarray                # f64[N]
    lb: np.ndarray                   # f64[d]
    ub: np.ndarray                   # f64[d]
    gamma: float = 1.0
    design: dict = field(default_factory=dict)   # model specific (Phi | x | T0,dT,n_steps,beta)
    # further models of the same likelihood (Likelihood.models is a dict, summed over at bayesian_inference.py:500-524):
    # Problems that share lb / ub / gamma with this one; each has its own data, nominal parameters and name map
    extra: list = field(default_factory=list)
    # element-wise re-parametrisation of the uncertain parameters (the Model.parametrization_* hooks a user of the
    # reference writes, bayesian_inference.py:314-336): {"kind": [...], "a", "b", "c", "det_jac": bool} with
    # kind 0: X = a + b Y, kind 1: X = a + b exp(c Y); det_jac False = the hook left at the base class's constant 1
    param: dict = None

    def backward(self, Y):
        """Model.parametrization_backward (:320-324)."""
        Y = np.asarray(Y, dtype=np.float64)
        if self.param is None:
            return Y
        k, a, b, c = (np.asarray(self.param[n], dtype=np.float64) for n in ("kind", "a", "b", "c"))
        return np.where(k == 0, a + b * Y, a + b * np.exp(c * Y))

    def dXdY(self, X):
        k, a, b, c = (np.asarray(self.param[n], dtype=np.float64) for n in ("kind", "a", "b", "c"))
        return np.where(k == 0, b, c * (np.asarray(X, dtype=np.float64) - a))

    def det_jac(self, X):
        """Model.parametrization_det_jac (:326-330)."""
        if self.param is None or not self.param.get("det_jac", True):
            return 1
        return float(np.prod(1.0 / self.dXdY(X)))

    @property
    def members(self):
        return [self] + list(self.extra)

    @property
    def d(self):
        return len(self.unc_idx)

    @property
    def n_points(self):
        return sum(m.y.shape[1] for m in self.members)

    def full_theta(self, X):
        # Model.run_model: uncertain values written into the nominal vector (:361-369)
        th = np.array(self.theta_nom, dtype=np.float64)
        for i, idx in enumerate(self.unc_idx):
            th[idx] = X[i]
        return th

    def model_eval(self, X):
        """Model outputs of every model of the likelihood, concatenated in order."""
        if not self.extra:
            return self.model_eval_one(X)
        return np.concatenate([m.model_eval_one(X) for m in self.members])

    def model_eval_one(self, X):
        th = self.full_theta(X)
        if self.model == "linear":
            return cpu_models.linear_design(th, self.design["Phi"])
        if self.model == "poly":
            return cpu_models.poly_horner(th, self.design["x"])
        if self.model == "arrhenius":
            dg = self.design
            return cpu_models.arrhenius_rk4(th, dg["T0"], dg["dT"], dg["n_steps"], dg["beta"])
        if self.model == "heat":
            return cpu_models.heat_conduction(th, self.design["x"])
        raise ValueError(self.model)

    def model_grad(self, X):
        """d f / d X_i for the uncertain parameters, dict i -> f64[N] (analytical)."""
        th = self.full_theta(X)
        if self.model == "linear":
            g = cpu_models.linear_design_grad(th, self.design["Phi"])
        elif self.model == "poly":
            g = cpu_models.poly_grad(th, self.design["x"])
        else:
            raise ValueError("no analytical gradient for model %r" % self.model)
        return {i: g[int(idx)] for i, idx in enumerate(self.unc_idx)}


def log_prior(prob, X):
    """Mixture of Uniform: sum_i log(p_i(X_i)); bounds inclusive (distributions.py:282-289, :456-479)."""
    Y = 0.0
    with np.errstate(divide="ignore"):
        for i in range(prob.d):
            if X[i] < prob.lb[i] or X[i] > prob.ub[i]:
                p = 0
            else:
                p = 1 / abs(prob.ub[i] - prob.lb[i])
            Y += np.log(p)
    return Y


def arg_gauss_likelihood(prob, X):
    """J = sum_models sum_runs sum_k ((y - f)/(std_y*gamma))^2   (bayesian_inference.py:495-542)."""
    J = 0
    for m in prob.members:
        f = m.model_eval_one(X)
        for c_run in range(m.y.shape[0]):
            arg_exp = (m.y[c_run] - f) / (m.std_y * prob.gamma)
            J = J + np.sum(arg_exp ** 2, axis=0)
    return J


def log_post(prob, Y):
    """BayesianPosterior.compute_log_value (:44-64): X = backward(Y); prior(X) - log(det_jac(X)) + loglike(X)."""
    X = prob.backward(Y)
    lp = log_prior(prob, X)
    if lp == -np.inf:
        return -np.inf
    log_like = -(1 / 2) * arg_gauss_likelihood(prob, X)
    return lp - np.log(prob.det_jac(X)) + log_like


def grad_log_post(prob, Y):
    """Analytical gradient, BayesianPosterior.compute_grad_log_like (:80-89): compute_grad_log_value (:573-623) with the
    model's inv_jac = diag(dX/dY), minus compute_grad_log_det_jac (:626-636).

    Only run 0 enters (:606-608) and the prior does not (:80-82); the models of the likelihood add up (:596).
    """
    X = prob.backward(Y)
    grad = np.zeros(prob.d)
    inv_jac = np.eye(prob.d) if prob.param is None else np.diag(prob.dXdY(X))
    for m in prob.members:
        f = m.model_eval_one(X)
        mg = m.model_grad(X)
        ss_x = (m.y[0] - f) / (m.std_y ** 2 * prob.gamma ** 2)
        for i in range(prob.d):
            my_mat = np.matmul(ss_x, mg[i])
            if prob.param is None:
                grad[i] = grad[i] + my_mat
            else:
                for k in range(prob.d):
                    grad[k] = grad[k] + my_mat * inv_jac[i, k]                     # :621
    if prob.param is not None:
        det = prob.det_jac(X)
        if prob.param.get("det_jac", True):
            k_, a = np.asarray(prob.param["kind"]), np.asarray(prob.param["a"], dtype=np.float64)
            grad_det_jac = np.where(k_ == 0, 0.0, -det / (X - a))
        else:
            grad_det_jac = np.zeros(prob.d)
        grad = grad - np.matmul(grad_det_jac, inv_jac) / det                        # :634, :89
    return grad


def grad_fd(fun, X, eps=1e-10):
    """computeGradientFD, forward scheme (:1069-1088)."""
    n = X.size
    g = np.zeros(n)
    f0 = fun(X)
    for i in range(n):
        Xp = np.array(X)
        delta = Xp[i] * eps
        if delta == 0.0:
            delta = eps
        Xp[i] = X[i] + delta
        g[i] = (fun(Xp) - f0) / delta
    return g


# --------------------------------------------------------------------------- streams
class Streams:
    """Stand-in for the stdlib ``random`` module: pre-generated normals and uniforms."""

    def __init__(self, normals, uniforms):
        self.normals = np.asarray(normals, dtype=np.float64)
        self.uniforms = np.asarray(uniforms, dtype=np.float64)
        self.i_n = 0
        self.i_u = 0

    def gauss(self, mu=0.0, sigma=1.0):
        v = self.normals[self.i_n]
        self.i_n += 1
        return mu + sigma * float(v)

    def random(self):
        v = self.uniforms[self.i_u]
        self.i_u += 1
        return float(v)


def py_min1(r):
    """Python's ``min(1, r)``: returns r only when r < 1, so NaN gives 1."""
    return r if r < 1 else 1


def chol_upper(V):
    """Upper Cholesky factor, V = R^T R: the reference's literal call, scipy.linalg.cholesky with
    its default lower=False (:143, :487, :592, :703, :775)."""
    return scipy.linalg.cholesky(np.asarray(V, dtype=np.float64))


def inv_upper(R):
    """scipy.linalg.inv, as the reference calls it on the Cholesky factor (:517, :594, :704)."""
    return scipy.linalg.inv(np.asarray(R, dtype=np.float64))


def fma(a, b, c):
    """Correctly rounded a*b + c (one rounding), by exact rational arithmetic: what an FMA instruction returns."""
    a, b, c = float(a), float(b), float(c)
    if not (np.isfinite(a) and np.isfinite(b) and np.isfinite(c)):
        return a * b + c
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def chol_upper_plain(V):
    """Upper Cholesky in the operation order of LAPACK's unblocked dpotf2('U') -- the routine behind
    scipy.linalg.cholesky at these sizes (mha.py:143, :487, :592) -- written out so that the CUDA device function
    (pbu_common.cuh::chol_upper_packed) can be held to it operation by operation:

        dot = sum_k R_kj^2          accumulated from zero with fused multiply-adds (ddot)
        ajj = sqrt(V_jj - dot)      one subtraction
        t_i = sum_k R_kj R_ki       from zero with fused multiply-adds (dgemv), i > j
        R_ji = (V_ji - t_i) * (1/ajj)   scaled by the RECIPROCAL (dscal)

    Up to d = 4 this is bit-identical to scipy with OpenBLAS (test_cholesky_potf2_order_is_scipys); from d = 5 on
    OpenBLAS's vectorised dot / gemv kernels sum in lane order (e.g. ((p0+p2)+(p1+p3)) for a length-4 dot on this
    image's SkylakeX build), which depends on the BLAS build; the two then agree to rounding."""
    V = np.asarray(V, dtype=np.float64)
    n = V.shape[0]
    R = np.zeros((n, n))
    for j in range(n):
        dot = 0.0
        for k in range(j):
            dot = fma(R[k, j], R[k, j], dot)
        s = V[j, j] - dot
        if not s > 0.0:
            raise np.linalg.LinAlgError("matrix not positive definite")
        ajj = float(np.sqrt(s))
        R[j, j] = ajj
        rinv = 1.0 / ajj
        for i in range(j + 1, n):
            t = 0.0
            for k in range(j):
                t = fma(R[k, j], R[k, i], t)
            R[j, i] = (V[j, i] - t) * rinv
    return R


def proposal_matvec(R, z):
    """R @ z of the proposal, the reference's literal call (np.matmul, mha.py:193, :543)."""
    return np.matmul(R, z)


def proposal_matvec_plain(R, z):
    """R @ z (R upper triangular) in the summation order the CUDA kernels use (pbu_common.cuh::upper_matvec), which up
    to d = 4 is the order numpy/OpenBLAS produces on the image the golden vectors were recorded on
    (test_proposal_matvec_order_is_numpys); from d = 5 on a fused multiply-add chain over j >= i."""
    R = np.asarray(R, dtype=np.float64)
    d = R.shape[0]
    p = [[float(R[i, j]) * float(z[j]) for j in range(d)] for i in range(d)]
    if d == 1:
        return np.array([p[0][0]])
    if d == 2:
        return np.array([fma(R[0, 0], z[0], p[0][1]), p[1][1]])
    if d == 3:
        return np.array([fma(R[0, 2], z[2], fma(R[0, 0], z[0], p[0][1])), fma(R[1, 2], z[2], p[1][1]), p[2][2]])
    if d == 4:
        return np.array([(p[0][0] + p[0][2]) + (p[0][1] + p[0][3]), p[1][2] + (p[1][1] + p[1][3]), p[2][2] + p[2][3], p[3][3]])
    out = np.zeros(d)
    for i in range(d):
        s = 0.0
        for j in range(i, d):
            s = fma(R[i, j], z[j], s)
        out[i] = s
    return out


class kernel_order:
    """Context manager: run the restatement with the WRITTEN-OUT operation orders the CUDA kernels implement (dpotf2
    Cholesky, back-substitution inverse, proposal mat-vec) instead of the scipy / numpy calls the reference makes.  Up to
    d = 4 the two are bit-identical on the image the golden vectors come from; the written-out forms make a device-vs-
    oracle comparison independent of the BLAS build of the machine the test runs on."""

    def __enter__(self):
        g = globals()
        self.saved = (g["chol_upper"], g["inv_upper"], g["proposal_matvec"])
        g["chol_upper"], g["inv_upper"], g["proposal_matvec"] = chol_upper_plain, inv_upper_plain, proposal_matvec_plain
        return self

    def __exit__(self, *a):
        g = globals()
        g["chol_upper"], g["inv_upper"], g["proposal_matvec"] = self.saved


def inv_upper_plain(R):
    """Inverse of an upper-triangular matrix by back substitution."""
    n = R.shape[0]
    X = np.zeros((n, n))
    for j in range(n):
        X[j, j] = 1.0 / R[j, j]
        for i in range(j - 1, -1, -1):
            s = 0.0
            for k in range(i + 1, j + 1):
                s += R[i, k] * X[k, j]
            X[i, j] = -s / R[i, i]
    return X


# --------------------------------------------------------------------------- MH family
def run_mh(prob, x0, V, n_it, streams, adaptive=False, delayed=False,
           updating_it=10, eps_v=0.0, gamma_dr=0.2, am_welford=False):
    """RWMH / AM / DR / DRAM for ONE chain, literal reference semantics.

    ``am_welford=True`` replaces the literal Haario recursion (:455, :463) by the engine's production update
    (north_star: "the adaptive-covariance update is a Welford accumulation"; pbu_mh_kernel.cuh / pbu_mh_coop.cuh):
    running mean and sum of outer products of deviations over x_0 .. x_i,
        delta = x_i - mean;  mean += delta / (i + 1);  M2 += delta (x_i - mean)^T;
        V_i = S_d (M2 / i + eps_v I)
    i.e. the same estimator S_d (cov(x_0..x_i) + eps I) without the recursion's cancellation (and without its i = 1
    quirks Q3 / Q4).  Written with the kernel's fused multiply-adds.  This is NOT reference arithmetic: it pins the
    engine's own opt-in mode.

    Returns a dict with chain[T+1,d], lp[T+1] (log-posterior of the current state after
    each iteration), acc[T] (0 rejected, 1 accepted at stage 1, 2 at stage 2),
    eval_x / eval_lp (every posterior evaluation in call order), R, V_i, X_av, n_rejected,
    mean_r and n_eval.
    """
    d = prob.d
    cur = np.array(x0, dtype=np.float64)
    evals_x, evals_lp = [], []

    def post(X):
        v = log_post(prob, X)
        evals_x.append(np.array(X))
        evals_lp.append(v)
        return v

    lp_cur = post(cur)
    if adaptive and delayed:
        lp_cur = post(cur)          # DRAM runs MetropolisHastings.__init__ twice (:588-589)
    R = chol_upper(V)
    S_d = 2.38 ** 2 / d
    eps_Id = eps_v * np.eye(d)
    V_i = np.array(V, dtype=np.float64)
    X_av = None                      # aliases the state until the first adaptation (Q3)
    inv_V = None
    if delayed:
        inv_R = inv_upper(R)
        inv_V = inv_R * np.transpose(inv_R)      # elementwise (:517-518) -> diag(1/R_ii^2)
    chain = np.zeros((n_it + 1, d))
    lps = np.zeros(n_it + 1)
    acc = np.zeros(n_it, dtype=np.int32)
    chain[0] = cur
    lps[0] = lp_cur
    n_rejected = 0
    mean_r = 0.0
    arg_MAP, MAP_val = np.array(cur), lp_cur             # estimate_max state (:138-139)
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        for it in range(1, n_it + 1):
            z = np.array([streams.gauss(0, 1) for _ in range(d)])
            new = cur + proposal_matvec(R, z)                # :193
            lp_new = post(new)
            if lp_new == -np.inf:
                r = 0
            else:
                r = np.exp(lp_new - lp_cur)                  # :205
            alpha = py_min1(r)
            u = streams.random()
            if u < alpha:
                cur = np.array(new)
                if lp_new > MAP_val:                         # estimate_max (:226-237): first-stage acceptances only (Q10)
                    arg_MAP, MAP_val = np.array(new), lp_new
                lp_cur = lp_new
                acc[it - 1] = 1
            elif delayed:
                z2 = np.array([streams.gauss(0, 1) for _ in range(d)])
                new2 = cur + gamma_dr * proposal_matvec(R, z2)   # :543
                lp2 = post(new2)
                r_12 = np.exp(lp_new - lp2)
                alpha_12 = py_min1(r_12)
                diff1 = new - new2
                diff2 = new - cur
                M2 = np.matmul(np.matmul(diff1, inv_V), diff1)
                M4 = np.matmul(np.matmul(diff2, inv_V), diff2)
                r_2 = np.exp(lp2 - lp_cur) * (np.exp(-1 / 2 * M2) / np.exp(-1 / 2 * M4)) \
                    * (1 - alpha_12) / (1 - alpha)           # :555-556
                alpha_2 = py_min1(r_2)
                u2 = streams.random()
                if u2 < alpha_2:
                    cur = np.array(new2)
                    lp_cur = lp2
                    acc[it - 1] = 2
                else:
                    n_rejected += 1
            else:
                n_rejected += 1

            if adaptive and am_welford:
                i = it
                if X_av is None:
                    X_av = np.array(x0, dtype=np.float64)      # mean of {x_0}
                    V_i = np.zeros((d, d))                      # M2
                inv_n = 1.0 / (i + 1.0)
                sc = S_d / i
                dl = cur - X_av
                mu = np.array([fma(dl[a], inv_n, X_av[a]) for a in range(d)])
                M2 = np.array(V_i)
                for a in range(d):
                    for b in range(a, d):
                        M2[a, b] = M2[b, a] = fma(dl[a], cur[b] - mu[b], V_i[a, b])
                V_i, X_av = M2, mu
                if i % updating_it == 0:
                    Vc = np.array([[fma(sc, M2[a, b], (S_d * eps_v) if a == b else 0.0) for b in range(d)] for a in range(d)])
                    R = chol_upper(Vc)
                    if delayed:
                        inv_R = inv_upper(R)
                        inv_V = np.transpose(inv_R) * inv_R
                mean_r += r
            elif adaptive:
                i = it
                X_i = cur
                X_av_prev = cur if X_av is None else X_av    # aliasing at i == 1 (Q3)
                X_av_ip = X_i + (i) / (i + 1) * (X_av_prev - X_i)          # :455
                V_ip = (i - 1) / i * V_i + S_d / i * (
                    i * np.tensordot(X_av_prev, X_av_prev, axes=0)
                    - (i + 1) * np.tensordot(X_av_ip, X_av_ip, axes=0)
                    + np.tensordot(X_i, X_i, axes=0) + eps_Id)             # :463
                V_i = V_ip
                X_av = X_av_ip
                if i % updating_it == 0:
                    R = chol_upper(V_i)                      # :487 / :592
                    if delayed:
                        inv_R = inv_upper(R)
                        inv_V = np.transpose(inv_R) * inv_R  # :595
                mean_r += r                                  # :419 (unclipped, Q9)
            else:
                mean_r += alpha                              # :170
            chain[it] = cur
            lps[it] = lp_cur
    if adaptive and am_welford and n_it > 0:
        V_i = S_d / n_it * V_i + S_d * eps_v * np.eye(d)       # what pbu_engine_get_proposal reports for this mode
    return dict(chain=chain, lp=lps, acc=acc, eval_x=np.array(evals_x), eval_lp=np.array(evals_lp),
                R=R, V_i=V_i, X_av=(cur if X_av is None else X_av), n_rejected=n_rejected,
                mean_r=mean_r, n_eval=len(evals_lp), arg_MAP=arg_MAP, MAP_val=MAP_val)


# --------------------------------------------------------------------------- Ito-SDE
def _csv_quantise(v):
    """What survives np.savetxt(fmt='%f') + pandas.read_csv (bayesian_inference.py:134; Q11/Q18)."""
    return np.array([float("%f" % t) for t in v])


def run_isde(prob, x0, n_it, streams, h, f0, gradient="Analytical", C0=None,
             starting_it=None, update_it=1, end_it=None, csv_quantise=False, track_map=False):
    """ito_SDE.run_algorithm (:932-1040) for ONE chain.

    track_map: estimate_max_distr -- one MORE posterior evaluation per iteration, at the new state (:1011-1019).

    C0: initial preconditioner (identity when None, the 'Identity' branch :661-664).
    Adaptation window as GradientBasedMCMC.adapt_covariance (:736-777).  The reference
    initialises the window from the chain CSV re-read from disk (6-decimal text, Q11);
    ``csv_quantise=True`` reproduces that, ``False`` uses the exact states (what the
    CUDA engine does, a documented deviation).
    Returns chain[T+1,d], P[T+1,d], final C, X_av, V_i.
    """
    d = prob.d
    xi = np.array(x0, dtype=np.float64)
    P = np.zeros(d)
    C = np.eye(d) if C0 is None else np.array(C0, dtype=np.float64)
    L_c = chol_upper(C)
    inv_L_c = inv_upper(L_c)
    if starting_it is None:
        starting_it = n_it + 10
    if end_it is None:
        end_it = n_it + 1
    S_d = 1.0
    eps_Id = 1e-10 * np.eye(d)
    X_av = np.array(x0, dtype=np.float64)
    V_i = None
    hfm = 1 - h * f0 / 4
    hfp = 1 + h * f0 / 4
    if gradient == "Analytical":
        def grad(X):
            return grad_log_post(prob, X)
    else:
        def grad(X):
            return grad_fd(lambda Z: log_post(prob, Z), X, eps=0.0000000001)
    chain = np.zeros((n_it + 1, d))
    Ps = np.zeros((n_it + 1, d))
    chain[0] = xi
    arg_MAP, MAP_val = np.array(xi), (log_post(prob, xi) if track_map else None)     # :138-139 (MetropolisHastings.__init__)
    rows = [_csv_quantise(xi) if csv_quantise else np.array(xi)]
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        for it in range(1, n_it + 1):
            cur = xi
            if starting_it <= it <= end_it + 1:
                if it == starting_it:
                    vals = np.array(rows)
                    X_av = np.mean(vals, axis=0)
                    V_i = S_d * (np.atleast_2d(np.cov(vals, rowvar=False)) + eps_Id)
                elif it <= end_it:
                    X_i = cur
                    X_av_ip = X_i + (it) / (it + 1) * (X_av - X_i)
                    V_ip = (it - 1) / it * V_i + S_d / it * (
                        it * np.tensordot(X_av, X_av, axes=0)
                        - (it + 1) * np.tensordot(X_av_ip, X_av_ip, axes=0)
                        + np.tensordot(X_i, X_i, axes=0) + eps_Id)
                    V_i = V_ip
                    X_av = X_av_ip
                if it % update_it == 0:
                    C = V_i
                    L_c = chol_upper(C)
                    inv_L_c = inv_upper(L_c)
            WP = np.sqrt(h) * np.array([streams.gauss(0, 1) for _ in range(d)])    # :990
            xi_n = xi + h / 2 * np.matmul(C, P)                                      # :991
            grad_phi = -grad(xi_n)                                                   # :994
            P_np = hfm / hfp * P + h / hfp * (-grad_phi) + np.sqrt(f0) / hfp * np.matmul(inv_L_c, WP)
            xi_np = xi_n + h / 2 * np.matmul(C, P_np)                                # :1001
            if track_map:
                lp_np = log_post(prob, xi_np)                                        # :1016
                if lp_np > MAP_val:
                    arg_MAP, MAP_val = np.array(xi_np), lp_np
            P = P_np
            xi = xi_np
            chain[it] = xi
            Ps[it] = P
            rows.append(_csv_quantise(xi) if csv_quantise else np.array(xi))
    return dict(chain=chain, P=Ps, C=C, X_av=X_av, V_i=V_i, arg_MAP=arg_MAP, MAP_val=MAP_val)


# --------------------------------------------------------------------------- Hamiltonian Monte Carlo
def run_hmc(prob, x0, n_it, streams, step_size, num_steps, gradient="Analytical", C0=None,
            starting_it=None, update_it=1, end_it=None):
    """HamiltonianMonteCarlo (:779-884) driven by MetropolisHastings.run_algorithm (:158-182) for ONE chain.

    Per iteration: adapt_covariance (:736-777, as for ISDE); p = z @ inv_L_c^T (d normals, :829-830); half momentum
    step with the gradient at the CURRENT state (:836); L position / momentum leapfrog steps with the preconditioner
    C (:842-850); final half step and sign flip (:852-855); r = exp(U_cur - U_new + K_cur - K_new) with
    K = sum((p @ C) * p)/2 (:862-879; the `is -np.inf` test never fires, exp(-inf) gives 0 anyway); one uniform (:213).
    On acceptance the stored gradient becomes the one at the end of the trajectory (:221).
    Returns chain[T+1,d], acc[T], lp[T+1], n_rejected, mean_r, final C.
    """
    d = prob.d
    cur = np.array(x0, dtype=np.float64)
    C = np.eye(d) if C0 is None else np.array(C0, dtype=np.float64)
    L_c = chol_upper(C)
    inv_L_c = inv_upper(L_c)
    inv_L_c_T = inv_upper(np.transpose(L_c)) if C0 is None else np.transpose(inv_L_c)
    if starting_it is None:
        starting_it = n_it + 10
    if end_it is None:
        end_it = n_it + 1
    S_d, eps_Id = 1.0, 1e-10 * np.eye(d)
    X_av, V_i = np.array(x0, dtype=np.float64), None
    if gradient == "Analytical":
        def grad(X):
            return grad_log_post(prob, X)
    else:
        def grad(X):
            return grad_fd(lambda Z: log_post(prob, Z), X, eps=0.0000000001)
    lp_cur = log_post(prob, cur)
    g_cur = grad(cur)
    chain = np.zeros((n_it + 1, d))
    lps = np.zeros(n_it + 1)
    acc = np.zeros(n_it, dtype=np.int32)
    chain[0], lps[0] = cur, lp_cur
    rows = [np.array(cur)]
    n_rejected, mean_r = 0, 0.0
    arg_MAP, MAP_val = np.array(cur), lp_cur
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        for it in range(1, n_it + 1):
            if starting_it <= it <= end_it + 1:
                if it == starting_it:
                    vals = np.array(rows)
                    X_av = np.mean(vals, axis=0)
                    V_i = S_d * (np.atleast_2d(np.cov(vals, rowvar=False)) + eps_Id)
                elif it <= end_it:
                    X_i = cur
                    X_av_ip = X_i + (it) / (it + 1) * (X_av - X_i)
                    V_i = (it - 1) / it * V_i + S_d / it * (
                        it * np.tensordot(X_av, X_av, axes=0) - (it + 1) * np.tensordot(X_av_ip, X_av_ip, axes=0)
                        + np.tensordot(X_i, X_i, axes=0) + eps_Id)
                    X_av = X_av_ip
                if it % update_it == 0:
                    C = V_i
                    L_c = chol_upper(C)
                    inv_L_c = inv_upper(L_c)
                    inv_L_c_T = np.transpose(inv_L_c)
            new = np.array(cur)
            z = np.array([streams.gauss(0, 1) for _ in range(d)])
            p = np.matmul(z, inv_L_c_T)
            cur_p = p
            p = p - step_size * (-g_cur) / 2
            g_new = g_cur
            for i in range(num_steps):
                new = new + step_size * (np.matmul(p, C))
                g_new = grad(new)
                if i != (num_steps - 1):
                    p = p - step_size * (-g_new)
            p = p - step_size * (-g_new) / 2
            p = -p
            cur_K = np.sum(np.matmul(cur_p, C) * cur_p) / 2
            lp_new = log_post(prob, new)
            prop_K = np.sum(np.matmul(p, C) * p) / 2
            r = np.exp((-lp_cur) - (-lp_new) + cur_K - prop_K)
            alpha = py_min1(r)
            u = streams.random()
            if u < alpha:
                cur = np.array(new)
                if lp_new > MAP_val:                         # accept_reject -> estimate_max (:218)
                    arg_MAP, MAP_val = np.array(new), lp_new
                lp_cur = lp_new
                g_cur = g_new
                acc[it - 1] = 1
            else:
                n_rejected += 1
            mean_r += alpha
            chain[it], lps[it] = cur, lp_cur
            rows.append(np.array(cur))
    return dict(chain=chain, lp=lps, acc=acc, n_rejected=n_rejected, mean_r=mean_r, C=C, arg_MAP=arg_MAP, MAP_val=MAP_val)


# --------------------------------------------------------------------------- propagation
def propagate(prob, samples):
    """Monte-Carlo propagation, emulator == 'None' (solve_problem.py:543-647).

    Returns fun_eval[S,N] and the statistics the reference writes: mean, min, max
    (computed as plain reductions, NOT with the reference's elif/±1e5 quirks, Q15) and the
    2.5 / 97.5 percentiles (np.percentile, :643-645).
    """
    S = samples.shape[0]
    fe = np.zeros((S, prob.n_points))
    for s in range(S):
        fe[s] = prob.model_eval(samples[s])
    return dict(fun_eval=fe, mean=fe.mean(axis=0), min=fe.min(axis=0), max=fe.max(axis=0),
                ci_lb=np.percentile(fe, 2.5, axis=0), ci_ub=np.percentile(fe, 97.5, axis=0))
