"""Prior densities of the Bayesian path (host-side mirror of ``pybitup/distributions.py``).

Only what the sampling path can use is provided: ``set_probability_dist`` building a ``Mixture`` of ``Uniform``
components (reference :10-37, :271-289, :424-479) -- the only prior that works in the reference.
The objects keep the reference's attribute names (``mixture_components[i].lb / .ub``) because the engine lowers the
prior to a box from exactly those; ``compute_value`` / ``compute_log_value`` are the scalar
definitions the device code implements, used here for single evaluations only.
"""
import numpy as np

IMPLEMENTED_PDFS = ["Gaussian", "Uniform", "Mixture", "SoizePDF"]


def set_probability_dist(name_list, hyperparam, n_rand_var):
    """Same call as the reference (distributions.py:10-37); the name of the LAST entry decides (as there)."""
    name = None
    for name in name_list:
        if name not in IMPLEMENTED_PDFS:
            raise ValueError("No {} density function implemented on the device path \n"
                             "Implemented prior functions are {}.".format(name, " ".join(IMPLEMENTED_PDFS)))
    if name == "Gaussian":
        return Gaussian(hyperparam, n_rand_var)
    if name == "Uniform":
        return Uniform(hyperparam, n_rand_var)
    if name == "SoizePDF":
        return SoizePDF(hyperparam, n_rand_var=2)          # no hyper-parameters, dimension fixed to two (:32-34)
    return Mixture(hyperparam, n_rand_var)


class ProbabilityDistribution:
    """Generic density (distributions.py:40-175): what the samplers need from a sampled distribution."""

    def __init__(self, hyperparam, n_rand_var):
        self.hyperparam = hyperparam
        self.n_rand_var = n_rand_var
        self.dim = n_rand_var
        self.name_random_var = ["X{}".format(i) for i in range(int(n_rand_var))]

    def save_sample(self, IO_fileID, value):
        np.savetxt(IO_fileID['MChains'], np.array([value]), fmt="%f", delimiter=",")          # :155-159

    def update_eval(self):
        pass

    def save_value(self, IO_util, current_it, model_eval=None):
        pass


class Gaussian(ProbabilityDistribution):
    """Multivariate normal density, hyperparam = [mean (1 x d), cov (d x d)] (distributions.py:178-252).
    On the device it is the linear "model" A x with A = L^-1 (cov = L L^T): (x-m)^T cov^-1 (x-m) = |A x - A m|^2."""

    def __init__(self, hyperparam, n_rand_var):
        ProbabilityDistribution.__init__(self, hyperparam, n_rand_var)
        self.mean = np.atleast_2d(np.array(hyperparam[0], dtype=np.float64))
        if self.mean.shape[0] != 1:
            self.mean = self.mean.T
        self.cov = np.atleast_2d(np.array(hyperparam[1], dtype=np.float64))
        self.dim = self.cov.shape[0]
        self.name_random_var = ["X{}".format(i) for i in range(self.dim)]
        self.distr_support = np.stack([self.mean[0] - 4 * np.sqrt(np.diag(self.cov)), self.mean[0] + 4 * np.sqrt(np.diag(self.cov))], axis=1)
        self.inv_cov = np.linalg.inv(self.cov)
        self.log_gauss_coeff = float(-0.5 * (self.dim * np.log(2 * np.pi) + np.linalg.slogdet(self.cov)[1]))
        self.gauss_coeff = float(np.exp(self.log_gauss_coeff))

    def compute_exp_arg(self, X):
        diff_x = np.asarray(X, dtype=np.float64) - self.mean[0]
        return float(diff_x @ self.inv_cov @ diff_x)

    def compute_value(self, X):
        return self.gauss_coeff * np.exp(-0.5 * self.compute_exp_arg(X))

    def compute_log_value(self, X):
        return self.log_gauss_coeff - 0.5 * self.compute_exp_arg(X)

    def lower(self):
        from .engine import ProblemSpec
        d = self.dim
        A = np.linalg.inv(np.linalg.cholesky(self.cov))
        return ProblemSpec(model="linear", theta_nom=self.mean[0].copy(), unc_idx=np.arange(d, dtype=np.int32),
                           y=(A @ self.mean[0])[None, :], std_y=np.ones(d), lb=np.full(d, -np.inf), ub=np.full(d, np.inf),
                           gamma=1.0, design={"Phi": A}, log_const=self.log_gauss_coeff)


class SoizePDF(ProbabilityDistribution):
    """2-D test density of Soize (2017), p. 65 (distributions.py:483-506); device model "soize"."""

    def __init__(self, hyperparam, n_rand_var=2):
        ProbabilityDistribution.__init__(self, hyperparam, 2)
        self.c0 = 1

    def compute_value(self, X):
        return self.c0 * np.exp(-15 * (X[0] ** 3 - X[1]) ** 2 - (X[1] - 0.3) ** 4)

    def compute_log_value(self, X):
        return np.log(self.c0) - 15 * (X[0] ** 3 - X[1]) ** 2 - (X[1] - 0.3) ** 4

    def lower(self):
        from .engine import ProblemSpec
        return ProblemSpec(model="soize", theta_nom=np.zeros(2), unc_idx=np.arange(2, dtype=np.int32), y=np.zeros((1, 2)),
                           std_y=np.ones(2), lb=np.full(2, -np.inf), ub=np.full(2, np.inf), gamma=1.0, design={},
                           log_const=float(np.log(self.c0)))


class Uniform(ProbabilityDistribution):
    """1-D uniform density on [lb, ub], bounds inclusive (distributions.py:271-289)."""

    def __init__(self, hyperparam, n_rand_var):
        ProbabilityDistribution.__init__(self, hyperparam, n_rand_var)
        self.lb = hyperparam[0]
        self.ub = hyperparam[1]
        self.prob = 1 / abs(hyperparam[1] - hyperparam[0])
        self.distr_support = np.array([self.lb, self.ub])

    def compute_value(self, X):
        return 0 if (X < self.lb or X > self.ub) else self.prob


class Mixture(ProbabilityDistribution):
    """Product of independent 1-D components, hyperparam = [names, params] (distributions.py:424-479)."""

    def __init__(self, hyperparam, n_rand_var):
        ProbabilityDistribution.__init__(self, hyperparam, n_rand_var)
        self.distr_names = hyperparam[0]
        self.distr_param = hyperparam[1]
        self.n_components = len(self.distr_names)
        self.dim = len(self.distr_names)
        self.distr_support = np.zeros([self.dim, 2])
        self.mixture_components = []
        for i in range(self.n_components):
            c = set_probability_dist([self.distr_names[i]], self.distr_param[i], n_rand_var)
            if not isinstance(c, Uniform):
                raise ValueError("only Uniform mixture components can be lowered to the device prior")
            self.mixture_components.append(c)
            self.distr_support[i] = c.distr_support

    def compute_value(self, X):
        Y = 1
        for i, c in enumerate(self.mixture_components):
            Y *= c.compute_value(X[i])
        return Y

    def compute_log_value(self, X):
        Y = 0.0
        with np.errstate(divide="ignore"):
            for i, c in enumerate(self.mixture_components):
                Y += np.log(c.compute_value(X[i]))
        return Y

    def box(self):
        """(lb, ub) arrays of the d components: what the device prior is built from."""
        return (np.array([float(c.lb) for c in self.mixture_components]),
                np.array([float(c.ub) for c in self.mixture_components]))
