"""Prior densities of the Bayesian path (host-side mirror of ``pybitup/distributions.py``).

Only what the sampling path can use is provided: ``set_probability_dist`` building a ``Mixture`` of ``Uniform``
components (reference :10-37, :271-289, :424-479) -- the only prior that works in the reference.
The objects keep the reference's attribute names (``mixture_components[i].lb / .ub``) because the engine lowers the
prior to a box from exactly those; ``compute_value`` / ``compute_log_value`` are the scalar
definitions the device code implements, used here for single evaluations only.
"""
import numpy as np

IMPLEMENTED_PDFS = ["Uniform", "Mixture"]


def set_probability_dist(name_list, hyperparam, n_rand_var):
    """Same call as the reference (distributions.py:10-37); the name of the LAST entry decides (as there)."""
    name = None
    for name in name_list:
        if name not in IMPLEMENTED_PDFS:
            raise ValueError("No {} density function implemented on the device path \n"
                             "Implemented prior functions are {}.".format(name, " ".join(IMPLEMENTED_PDFS)))
    if name == "Uniform":
        return Uniform(hyperparam, n_rand_var)
    return Mixture(hyperparam, n_rand_var)


class ProbabilityDistribution:
    def __init__(self, hyperparam, n_rand_var):
        self.hyperparam = hyperparam
        self.n_rand_var = n_rand_var
        self.dim = n_rand_var
        self.name_random_var = ["X{}".format(i) for i in range(int(n_rand_var))]


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
