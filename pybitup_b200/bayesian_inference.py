"""Model / Data / Likelihood / BayesianPosterior: host-side mirror of ``pybitup/bayesian_inference.py``.

The user-facing contract is the reference's (``from pybitup_b200 import bayesian_inference as bi``;
``class MyModel(bi.Model)``): same constructor signatures, attribute names and call order as the reference's
``Model`` (:254-389), ``Data`` (:141-210), ``Likelihood`` (:399-696) and ``BayesianPosterior`` (:15-139).  What
changes is where the arithmetic runs: a model declares which REGISTERED DEVICE MODEL it is

    class Cubic(bi.Model):
        device_model = "poly"          # 'linear' | 'poly' | 'arrhenius' | 'heat'

(and, for models whose design is not just ``x``, overrides ``device_design()``), and ``BayesianPosterior.lower()``
turns the whole posterior into the flat ``ProblemSpec`` the CUDA engine consumes.  Anything that cannot be lowered
raises ``UnsupportedProblem``: there is no CPU fallback.  ``fun_x`` is never called on the sampling path.
"""
import pathlib

import numpy as np

from . import _capi as capi
from .engine import ProblemSpec


class Data:
    """Experimental data of one model: x, y = {run: f64[N]}, std_y (bayesian_inference.py:154-182)."""

    def __init__(self, name="", x=np.array([1]), y=np.array([1]), std_y=np.array([1])):
        self.name = list([name])
        self.x = x
        self.num_points = np.array([len(self.x)])
        if not isinstance(y, dict):
            y = {0: np.asarray(y, dtype=np.float64)}
        self.y = y
        self.n_runs = len(y.keys())
        self.n_data_set = 1
        self.std_y = std_y
        self.index_data_set = np.array([[0, len(self.x) - 1]])
        self.mean_y = np.mean(np.stack([np.asarray(self.y[r], dtype=np.float64) for r in range(self.n_runs)]), axis=0)

    def size_x(self, i):
        return self.num_points[i]


class Model:
    """Forward model contract of the reference (bayesian_inference.py:254-389) plus the device declaration."""

    device_model = None      # name of the registered device model this class corresponds to

    def __init__(self, x=[], param=[], scaling_factors_parametrization=1, name=""):
        self.name = name
        self.P = scaling_factors_parametrization
        self._x = x
        self._param = param
        self.param_nom = []
        self.param_names = []
        self.input_file_name = []
        self.unpar_name = {}
        self.unpar_name_dict = {}
        self.model_eval = 0
        self.model_grad = 0

    def size_x(self):
        return len(self._x)

    param = property(lambda self: self._param, lambda self, v: setattr(self, "_param", v))
    x = property(lambda self: self._x, lambda self, v: setattr(self, "_x", v))

    # -- user hooks (reference names) --------------------------------------------------------
    def fun_x(self, *ext_parameters):
        return 1

    def d_fx_dparam(self, *ext_parameters):
        return 1

    def parametrization_forward(self, X=1, P=1):
        return X

    def parametrization_backward(self, Y=1, P=1):
        return Y

    def parametrization_det_jac(self, X=1):
        return 1

    def parametrization_inv_jac(self, X=1):
        return X

    # -- device declaration ----------------------------------------------------------------------
    def device_design(self):
        """Design data of the registered device model.  Default: the design points ``x`` (poly / heat); ``linear``
        models return {'Phi': [N, P]}; ``arrhenius`` models return {'T0', 'dT', 'n_steps', 'beta'}."""
        return {"x": np.asarray(self.x, dtype=np.float64)}

    def uncertain_index(self):
        """Positions of the uncertain parameters inside the full parameter vector, with the reference's
        substring matching (Model.run_model, :351-369): parameter names that occur inside the
        blank-joined uncertain names are uncertain, and the n-th uncertain VALUE goes to the n-th such name."""
        unpar = list(self.unpar_name)
        n_model = len(self.param_nom)
        if len(unpar) >= n_model:
            return np.arange(n_model, dtype=np.int32)          # param = var_param (:368-369)
        joined = " ".join(unpar)
        idx = [i for i, name in enumerate(self.param_names) if joined.find(name) >= 0]
        if len(idx) < len(unpar):
            raise ValueError("uncertain parameters {} not found among the model parameters {}".format(unpar, list(self.param_names)))
        return np.asarray(idx[: len(unpar)], dtype=np.int32)

    def run_model(self, var_param):
        """Kept for interface compatibility: evaluates the model through the user's numpy ``fun_x`` with the
        reference's parameter mapping.  NOT used by the device samplers or by device propagation."""
        if self.input_file_name:
            self.model_eval = self.fun_x(self.input_file_name, self.unpar_name, var_param)
            return
        idx = self.uncertain_index()
        if len(list(self.unpar_name)) < len(self.param_nom):
            vec = self.param_nom
            for n, i in enumerate(idx):
                vec[i] = var_param[n]
            self.param = vec
        else:
            self.param = var_param
        self.model_eval = self.fun_x()

    def get_gradient_model(self, var_param):
        self.model_grad = self.d_fx_dparam()


class Likelihood:
    """Gaussian likelihood over {model_id: Data} and {model_id: Model} (bayesian_inference.py:399-542)."""

    def __init__(self, exp_data, model_list, gamma=1.0):
        self.data = exp_data
        self.models = model_list
        self.SS_X = 0
        self.arg_LL = 0
        self.gamma = gamma
        self.model_eval = {}

    def lower(self, lb, ub):
        """ProblemSpec of the model/data pairs with the box prior [lb, ub].

        Several models (the dicts are walked in key order, as arg_gauss_likelihood does at :500) become model blocks of
        one device problem: same registered device model, same number of parameters and of runs; each keeps its own
        nominal parameters, name map and design."""
        ids = list(self.models.keys())
        if not ids:
            raise capi.UnsupportedProblem(capi.PBU_ERR_UNSUPPORTED, "the likelihood has no model")
        specs = []
        for mid in ids:
            model, data = self.models[mid], self.data[mid]
            if model.input_file_name:
                raise capi.UnsupportedProblem(capi.PBU_ERR_UNSUPPORTED, "external-solver models (input_file) cannot run on the device")
            if model.device_model not in capi.MODEL_IDS:
                raise capi.UnsupportedProblem(
                    capi.PBU_ERR_UNSUPPORTED,
                    "model class {} does not declare a registered device model (device_model = one of {}); "
                    "there is no CPU fallback".format(type(model).__name__, sorted(capi.MODEL_IDS)))
            y = np.stack([np.asarray(data.y[r], dtype=np.float64) for r in range(data.n_runs)])
            specs.append(dict(model=model.device_model, theta_nom=np.array(model.param_nom, dtype=np.float64),
                              unc_idx=model.uncertain_index(), y=y, std_y=np.asarray(data.std_y, dtype=np.float64),
                              design=model.device_design()))
        first = specs[0]
        blocks = None
        y, std_y = first["y"], first["std_y"]
        if len(specs) > 1:
            for sp in specs[1:]:
                if (sp["model"] != first["model"] or sp["theta_nom"].shape != first["theta_nom"].shape
                        or sp["y"].shape[0] != first["y"].shape[0]):
                    raise capi.UnsupportedProblem(
                        capi.PBU_ERR_UNSUPPORTED,
                        "the models {} of one posterior must be the same registered device model with the same number of "
                        "parameters and of data runs to be lowered together".format(ids))
            blocks = [dict(theta_nom=sp["theta_nom"], unc_idx=sp["unc_idx"], design=sp["design"], n_points=sp["y"].shape[1])
                      for sp in specs]
            y = np.concatenate([sp["y"] for sp in specs], axis=1)
            std_y = np.concatenate([sp["std_y"] for sp in specs])
        return ProblemSpec(model=first["model"], theta_nom=first["theta_nom"], unc_idx=first["unc_idx"], y=y, std_y=std_y,
                           lb=np.asarray(lb, dtype=np.float64), ub=np.asarray(ub, dtype=np.float64),
                           gamma=float(self.gamma), design=first["design"], blocks=blocks)

    def model_slices(self):
        """{model id: slice of the concatenated points} in the order lower() lays the models out."""
        out, at = {}, 0
        for mid in self.models.keys():
            n = len(np.asarray(self.data[mid].std_y))
            out[mid] = slice(at, at + n)
            at += n
        return out


_NICE_RATES = (1.0, -1.0, float(np.log(10.0)), -float(np.log(10.0)), float(np.log(2.0)), -float(np.log(2.0)), 0.5, -0.5, 2.0, -2.0)


def _snap(v, candidates, rtol=1e-9):
    for c in candidates:
        if abs(v - c) <= rtol * max(1.0, abs(c)):
            return float(c)
    return float(v)


def lower_parametrization(model, Y0, need_gradient=False):
    """The user's ``parametrization_*`` hooks (bayesian_inference.py:314-336) as a REGISTERED device change of variables,
    or ``None`` for the identity.  The hooks are arbitrary Python; the device knows element-wise maps of two kinds,
    X_i = a_i + b_i Y_i (scaling / shift) and X_i = a_i + b_i exp(c_i Y_i) (log, log10, shifted-log sampling), with
    -log(det_jac(X)) affine in Y (true for these maps, and for a hook left at the base class's constant 1).  The hooks
    are PROBED around the starting point: each component is identified from three equally spaced evaluations, then the
    identified map -- and, for analytical-gradient samplers, ``parametrization_inv_jac`` / ``grad_det_jac`` -- must
    reproduce the hooks at random points to 1e-11, else ``UnsupportedProblem`` (there is no CPU fallback)."""
    Y0 = np.asarray(Y0, dtype=np.float64).copy()
    d = Y0.size

    def back(Y):
        X = np.asarray(model.parametrization_backward(np.array(Y, dtype=np.float64)), dtype=np.float64)
        if X.shape != (d,):
            raise capi.UnsupportedProblem(capi.PBU_ERR_UNSUPPORTED, "parametrization_backward must map d values to d values")
        return X

    def neg_log_det(Y):
        with np.errstate(all="ignore"):
            return -float(np.log(model.parametrization_det_jac(back(Y))))

    X0 = back(Y0)
    rng = np.random.default_rng(12345)
    h = 0.05 * np.maximum(np.abs(Y0), 1.0)
    probes = [Y0 + h * rng.uniform(-2.0, 2.0, d) for _ in range(6)]
    if (np.array_equal(X0, Y0) and all(np.array_equal(back(P), P) for P in probes)
            and all(neg_log_det(P) == 0.0 for P in [Y0] + probes)):
        return None
    refuse = lambda why: capi.UnsupportedProblem(      # noqa: E731
        capi.PBU_ERR_UNSUPPORTED, "the model's re-parametrisation cannot be lowered to the device: " + why)
    kind, a, b, c = np.zeros(d, dtype=np.int32), np.zeros(d), np.ones(d), np.zeros(d)
    for i in range(d):
        e = np.zeros(d)
        e[i] = h[i]
        X1, X2 = back(Y0 + e), back(Y0 + 2 * e)
        others = np.arange(d) != i
        if not (np.array_equal(X1[others], X0[others]) and np.array_equal(X2[others], X0[others])):
            raise refuse("it is not element-wise (component %d of Y moves other components of X)" % i)
        d1, d2 = X1[i] - X0[i], X2[i] - X1[i]
        if d1 == 0.0:
            raise refuse("component %d of X does not depend on Y" % i)
        if abs(d2 - d1) <= 1e-10 * abs(d1):
            b[i] = _snap(0.5 * (d1 + d2) / h[i], (1.0, -1.0))
            a[i] = X0[i] - b[i] * Y0[i]
            if abs(a[i]) <= 1e-12 * max(1.0, abs(X0[i])):
                a[i] = 0.0
        else:
            rho = d2 / d1
            if not rho > 0.0:
                raise refuse("component %d is neither affine nor exponential in Y" % i)
            kind[i] = 1
            c[i] = _snap(np.log(rho) / h[i], _NICE_RATES)
            b[i] = d1 / (np.exp(c[i] * Y0[i]) * np.expm1(c[i] * h[i]))
            b[i] = _snap(b[i], (1.0, -1.0))
            a[i] = X0[i] - b[i] * np.exp(c[i] * Y0[i])
            if abs(a[i]) <= 1e-11 * max(1.0, abs(X0[i])):
                a[i] = 0.0
            else:
                a[i] = _snap(a[i], (round(a[i]),), rtol=1e-10)

    def dev_back(Y):
        return np.where(kind == 1, a + b * np.exp(c * Y), a + b * Y)

    for P in [Y0] + probes:
        Xu, Xd = back(P), dev_back(P)
        if np.max(np.abs(Xu - Xd) / np.maximum(np.abs(Xu), 1e-300)) > 1e-11:
            raise refuse("it is not of a registered kind (X = a + b Y or X = a + b exp(c Y) per component)")
    # -log(det_jac) = c0 + s . Y
    L0 = neg_log_det(Y0)
    if not np.isfinite(L0):
        raise refuse("parametrization_det_jac is not positive at the starting point")
    s_ = np.zeros(d)
    for i in range(d):
        e = np.zeros(d)
        e[i] = h[i]
        s_[i] = _snap((neg_log_det(Y0 + e) - L0) / h[i], (0.0, c[i]), rtol=1e-8)
    c0 = L0 - float(s_ @ Y0)
    true_c0 = float(np.sum(np.log(np.abs(np.where(kind == 1, b * c, b)))))
    for cand in (0.0, true_c0):
        if abs(c0 - cand) <= 1e-9 * max(1.0, abs(cand)):
            c0 = cand
    for P in probes:
        Lu, Ld = neg_log_det(P), c0 + float(s_ @ P)
        if abs(Lu - Ld) > 1e-10 * max(1.0, abs(Lu)):
            raise refuse("log(parametrization_det_jac) is not affine in the sampling variables")
    # analytical gradient: inv_jac must be diag(dX/dY) and grad_det_jac consistent with det_jac (:604-636)
    grad_ok, why = True, ""
    try:
        for P in probes[:3]:
            X = back(P)
            inv = np.asarray(model.parametrization_inv_jac(X), dtype=np.float64)
            dxdy = np.where(kind == 1, c * (X - a), b)
            if inv.shape != (d, d) or np.max(np.abs(inv - np.diag(dxdy))) > 1e-10 * np.max(np.abs(dxdy)):
                grad_ok, why = False, "parametrization_inv_jac is not diag(dX/dY)"
                break
            gld = np.matmul(np.asarray(model.grad_det_jac(X), dtype=np.float64), inv) / model.parametrization_det_jac(X)
            if np.max(np.abs(-gld - s_)) > 1e-9 * max(1.0, np.max(np.abs(s_))):
                grad_ok, why = False, "grad_det_jac is not the gradient of parametrization_det_jac"
                break
    except Exception as exc:                       # the base class has no grad_det_jac and its inv_jac returns X (:332-336)
        grad_ok, why = False, "parametrization_inv_jac / grad_det_jac: %s" % exc
    if need_gradient and not grad_ok:
        raise refuse(why)
    return dict(kind=kind, a=a, b=b, c=c, ldj_c0=float(c0), ldj_s=s_, grad_ok=grad_ok, grad_why=why)


class BayesianPosterior:
    """prior x likelihood in the sampling space (bayesian_inference.py:15-139)."""

    def __init__(self, prior, likelihood, class_model, param_init):
        self.param_init = param_init
        self.prior = prior
        self.likelihood = likelihood
        self.model = class_model
        self.dim = len(param_init)
        self.name_random_var = list(self.model.unpar_name)
        self._engine = None

    def get_dim(self):
        return len(self.param_init)

    def lower(self):
        """Flat problem descriptor for the CUDA engine; refuses what it cannot represent."""
        if not hasattr(self.prior, "box"):
            raise capi.UnsupportedProblem(capi.PBU_ERR_UNSUPPORTED, "only a Mixture of Uniform priors can be lowered (SURVEY Q12)")
        lb, ub = self.prior.box()
        spec = self.likelihood.lower(lb, ub)
        # the change of variables of the posterior is the one of `self.model` (:50, :61; solve_problem.py:336-339)
        spec.param = lower_parametrization(self.model, self.param_init)
        return spec

    # single evaluations through the device (diagnostics; the samplers never call these per iteration)
    def _eval_engine(self):
        if self._engine is None:
            from .engine import ChainEngine, SamplerConfig
            self._engine = ChainEngine(self.lower(), SamplerConfig(algo="RWMH"), 1)
        return self._engine

    def compute_log_value(self, Y):
        return float(self._eval_engine().log_posterior(np.asarray(Y, dtype=np.float64)[None, :])[0])

    def compute_value(self, Y):
        return float(np.exp(self.compute_log_value(Y)))

    def save_sample(self, IO_fileID, value):
        """Chain row in the sampling space and in the model space (bayesian_inference.py:131-138)."""
        np.savetxt(IO_fileID['MChains_reparam'], np.array([value]), fmt="%f", delimiter=",")
        X = self.model.parametrization_backward(value)
        np.savetxt(IO_fileID['MChains'], np.array([X]), fmt="%f", delimiter=",")

    def save_value(self, IO_util, current_it, model_eval):
        """output/model_eval/<id>_fun_eval-<it>.npy (Likelihood.write_fun_eval, :448-454)."""
        slices = self.likelihood.model_slices()
        for model_id in self.likelihood.models.keys():
            p = pathlib.Path(IO_util['path']['fun_eval_folder'], f"{model_id}_fun_eval-{current_it}.npy")
            np.save(p, np.asarray(model_eval)[slices[model_id]])
