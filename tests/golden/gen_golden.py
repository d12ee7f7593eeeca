"""Generate golden vectors by running the UNMODIFIED reference (jcoheur/pybitup).

Run in the build container only (needs /root/reference):

    python tests/golden/gen_golden.py

For each case it drives ``pybitup.solve_problem.Sampling(json).sample(models)``
(``solve_problem.py:125-348``) end to end -- JSON parsing, CSV data loading,
prior / likelihood / posterior construction, ``Sampler`` dispatch, the sampler
loop -- with the stdlib ``random`` module the samplers call replaced by
pre-generated streams (``mha.random = Streams``), and records
at FULL precision (the chain CSV is 6-decimal text):

* every ``BayesianPosterior.compute_log_value`` call (input, output),
* the state handed to ``save_sample`` after every iteration,
* final sampler attributes (R, V_i, X_av_i, n_rejected, mean_r, P for ISDE).

The forward models of the linear / polynomial / Arrhenius cases are this
repo's (``oracle/cpu_models.py``); the heat-conduction case is the reference's
own test case (``pybitup/test/test_pce/heat_conduction*.{py,json,csv}``).
Outputs: ``tests/golden/<case>.npz`` (a few tens of KB each).
"""
import json
import os
import sys
import tempfile

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim, cpu_models          # noqa: E402
from oracle.mcmc_oracle import Streams           # noqa: E402
from pybitup_b200 import synthetic               # noqa: E402

ref_shim.install()
import pybitup.solve_problem as sp               # noqa: E402  (the reference)
import pybitup.bayesian_inference as bi          # noqa: E402
import pybitup.metropolis_hastings_algorithms as mha  # noqa: E402


# ----------------------------------------------------------------------------- reference-side models
class LinearModel(bi.Model):
    Phi = None

    def fun_x(self):
        return cpu_models.linear_design(self.param, self.Phi)

    def d_fx_dparam(self):
        g = cpu_models.linear_design_grad(self.param, self.Phi)
        return {name: g[i] for i, name in enumerate(self.param_names)}

    def parametrization_inv_jac(self, X=1):
        return np.eye(len(X))

    def grad_det_jac(self, X=1):
        return np.zeros(len(X))


class PolyModel(bi.Model):
    def fun_x(self):
        return cpu_models.poly_horner(self.param, self.x)

    def d_fx_dparam(self):
        g = cpu_models.poly_grad(self.param, self.x)
        return {name: g[i] for i, name in enumerate(self.param_names)}

    def parametrization_inv_jac(self, X=1):
        return np.eye(len(X))

    def grad_det_jac(self, X=1):
        return np.zeros(len(X))


class ArrheniusModel(bi.Model):
    cfg = None

    def fun_x(self):
        c = self.cfg
        return cpu_models.arrhenius_rk4(self.param, c["T0"], c["dT"], c["n_steps"], c["beta"])


def heat_model():
    sys.path.insert(0, os.path.join(ref_shim.REFERENCE_ROOT, "pybitup", "test", "test_pce"))
    import heat_conduction                       # the reference's own model file
    return heat_conduction.HeatConduction()


# ----------------------------------------------------------------------------- recording hooks
class Recorder:
    def __init__(self):
        self.eval_x, self.eval_lp, self.saved = [], [], []
        self.sampler = None


def run_reference(workdir, json_name, models, normals, uniforms):
    rec = Recorder()
    streams = Streams(normals, uniforms)
    orig_clv = bi.BayesianPosterior.compute_log_value
    orig_save = bi.BayesianPosterior.save_sample
    orig_cov = mha.MetropolisHastings.compute_covariance
    orig_term = mha.MetropolisHastings.terminate_loop
    orig_random = mha.random

    def clv(self, Y):
        v = orig_clv(self, Y)
        rec.eval_x.append(np.array(Y, dtype=np.float64))
        rec.eval_lp.append(float(v))
        return v

    def save(self, IO_fileID, value):
        rec.saved.append(np.array(value, dtype=np.float64))
        return orig_save(self, IO_fileID, value)

    def cov(self, n_iterations):
        # harness workaround: the reference reads its own chain file while
        # the write handle is unflushed; flush so short runs do not crash.
        for f in self.IO_fileID.values():
            f.flush()
        return orig_cov(self, n_iterations)

    def term(self):
        rec.sampler = self
        return orig_term(self)

    bi.BayesianPosterior.compute_log_value = clv
    bi.BayesianPosterior.save_sample = save
    mha.MetropolisHastings.compute_covariance = cov
    mha.MetropolisHastings.terminate_loop = term
    mha.random = streams
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        with np.errstate(all="ignore"):
            s = sp.Sampling(json_name)
            s.sample(models)
            s.__del__()
    finally:
        os.chdir(cwd)
        bi.BayesianPosterior.compute_log_value = orig_clv
        bi.BayesianPosterior.save_sample = orig_save
        mha.MetropolisHastings.compute_covariance = orig_cov
        mha.MetropolisHastings.terminate_loop = orig_term
        mha.random = orig_random
    rec.n_normals_used = streams.i_n
    rec.n_uniforms_used = streams.i_u
    return rec


# ----------------------------------------------------------------------------- case builders
def write_case(workdir, prob, names, algo_block, x_col):
    """CSV + JSON exactly in the reference's input format (heat_conduction_1.json is the model)."""
    N = prob["y"].shape[1]
    df = pd.DataFrame({"x": x_col, "y": prob["y"][0], "std_y": prob["std_y"]})
    df.to_csv(os.path.join(workdir, "data.csv"))
    unc = [int(i) for i in prob["unc_idx"]]
    prior = {}
    for k, idx in enumerate(unc):
        prior[names[idx]] = {"initial_val": float(prob["x0"][k]), "prior_name": "Uniform",
                             "prior_param": [float(prob["lb"][k]), float(prob["ub"][k])]}
    js = {"Sampling": {
        "BayesianPosterior": {
            "Data": [{"Type": "ReadFromFile", "FileName": "data", "xField": ["x"],
                      "yField": ["y"], "sigmaField": ["std_y"]}],
            "Model": [{"param_names": names, "param_values": [float(v) for v in prob["theta_nom"]],
                       "parametrization": "no"}],
            "Prior": {"Distribution": "Mixture", "Param": prior},
            "Likelihood": {"function": "Gaussian", "distance": "L2"}},
        "Algorithm": algo_block}}
    with open(os.path.join(workdir, "case.json"), "w") as f:
        f.write("/* generated by tests/golden/gen_golden.py */\n")
        json.dump(js, f, indent=1)
    return N


def mh_algo(name, n_it, cov, **kw):
    blk = {"name": name, "n_iterations": n_it, "save_eval_freq": 10 ** 9,
           "proposal": {"name": "Gaussian", "covariance": {"type": "full", "value": np.asarray(cov).tolist()}}}
    if name == "AMH":
        blk["AMH"] = {"starting_it": 100, "updating_it": kw["updating_it"], "eps_v": kw["eps_v"]}
    if name == "DR":
        blk["DR"] = {"gamma": kw["gamma"]}
    if name == "DRAM":
        blk["DRAM"] = {"starting_it": 100, "updating_it": kw["updating_it"], "eps_v": kw["eps_v"],
                       "gamma": kw["gamma"]}
    return blk


def isde_algo(n_it, h, f0, gradient, starting_it, update_it, end_it=None):
    cov = {"value": "Identity", "starting_it": starting_it, "update_it": update_it}
    if end_it is not None:
        cov["end_it"] = end_it
    return {"name": "ISDE", "n_iterations": n_it, "save_eval_freq": 10 ** 9,
            "covariance": cov, "gradient": gradient, "ISDE": {"h": h, "f0": f0}}


def hmc_algo(n_it, step_size, num_steps, gradient, starting_it, update_it, end_it=None):
    cov = {"value": "Identity", "starting_it": starting_it, "update_it": update_it}
    if end_it is not None:
        cov["end_it"] = end_it
    return {"name": "HMC", "n_iterations": n_it, "save_eval_freq": 10 ** 9, "covariance": cov, "gradient": gradient,
            "HMC": {"step_size": step_size, "num_steps": num_steps}}


def build_model(prob, names):
    if prob["model"] == "linear":
        m = LinearModel()
        m.Phi = prob["design"]["Phi"]
    elif prob["model"] == "poly":
        m = PolyModel()
    elif prob["model"] == "arrhenius":
        m = ArrheniusModel()
        m.cfg = prob["design"]
    else:
        m = heat_model()
    return m


from tests.helpers import install_parametrization    # noqa: E402  (the hooks as a user would write them)


def pack_problem(prob):
    out = {"model": prob["model"], "theta_nom": prob["theta_nom"], "unc_idx": prob["unc_idx"],
           "y": prob["y"], "std_y": prob["std_y"], "lb": prob["lb"], "ub": prob["ub"],
           "gamma": prob["gamma"], "x0": prob["x0"], "prop_cov": prob["prop_cov"]}
    for k, v in prob["design"].items():
        out["design_" + k] = v
    return out


def run_case(case, prob, names, algo, seed, stream_len, param=None):
    """param: element-wise re-parametrisation of the uncertain parameters installed on the reference-side model
    (Model.parametrization_* hooks, bayesian_inference.py:314-336): {"kind": [...], "a": [...], "b": [...], "c": [...]} with
    kind 0: X = a + b Y, kind 1: X = a + b exp(c Y); "det_jac": whether parametrization_det_jac is the true |dY/dX|
    (False: the base class's constant 1)."""
    rng = np.random.default_rng(seed)
    normals = rng.standard_normal(stream_len)
    uniforms = rng.random(stream_len)
    with tempfile.TemporaryDirectory() as wd:
        write_case(wd, prob, names, algo, prob["design"].get("x", np.arange(prob["y"].shape[1], dtype=float)))
        model = build_model(prob, names)
        if param is not None:
            install_parametrization(model, param)
        rec = run_reference(wd, "case.json", {"case": model}, normals, uniforms)
    s = rec.sampler
    out = pack_problem(prob)
    out.update(algo_json=json.dumps(algo), normals=normals[:rec.n_normals_used],
               uniforms=uniforms[:rec.n_uniforms_used],
               eval_x=np.array(rec.eval_x), eval_lp=np.array(rec.eval_lp), chain=np.array(rec.saved),
               n_rejected=s.n_rejected, mean_r=float(s.mean_r))
    for attr in ("R", "V_i", "X_av_i", "inv_V", "C_approx", "P_nm", "xi_nm", "L_c", "inv_L_c", "arg_MAP", "MAP_val"):
        if hasattr(s, attr):
            out["final_" + attr] = np.array(getattr(s, attr), dtype=np.float64)
    if param is not None:
        out["param_json"] = json.dumps(param)
    path = os.path.join(HERE, case + ".npz")
    np.savez_compressed(path, **out)
    print("%-22s evals=%d chain=%s n_rej=%d -> %s (%d B)" % (
        case, len(rec.eval_lp), np.array(rec.saved).shape, s.n_rejected, os.path.basename(path),
        os.path.getsize(path)))


def main():
    lin = synthetic.linear_problem()
    lin_names = ["th0", "th1", "th2"]
    T = 300
    run_case("linear_rwmh", lin, lin_names, mh_algo("RWMH", T, lin["prop_cov"]), 11, 4 * T * 4)
    full_cov = np.array([[1e-3, 2e-3, -1e-3], [2e-3, 1e-2, 1e-3], [-1e-3, 1e-3, 1e-2]])
    lin2 = dict(lin, prop_cov=full_cov)
    run_case("linear_rwmh_fullcov", lin2, lin_names, mh_algo("RWMH", T, full_cov), 12, 4 * T * 4)
    run_case("linear_am", lin, lin_names, mh_algo("AMH", T, lin["prop_cov"], updating_it=10, eps_v=1e-10),
             13, 4 * T * 4)
    run_case("linear_dr", lin2, lin_names, mh_algo("DR", T, full_cov, gamma=0.2), 14, 8 * T * 4)
    run_case("linear_dram", lin, lin_names,
             mh_algo("DRAM", T, lin["prop_cov"], updating_it=10, eps_v=1e-10, gamma=0.2), 15, 8 * T * 4)

    # prior-box rejection path: tight box so that proposals leave it often
    lin3 = dict(lin, lb=np.array([0.95, 1.9, 2.9]), ub=np.array([1.05, 2.1, 3.1]), prop_cov=lin["prop_cov"] * 4)
    run_case("linear_dr_tightbox", lin3, lin_names, mh_algo("DR", T, lin3["prop_cov"], gamma=0.2), 16, 8 * T * 4)

    cub = synthetic.cubic_problem(n_points=100)
    cub = dict(cub, prop_cov=cub["prop_cov"] * 10)
    cub_names = ["c0", "c1", "c2", "c3"]
    run_case("cubic_am", cub, cub_names, mh_algo("AMH", T, cub["prop_cov"], updating_it=10, eps_v=1e-10),
             21, 5 * T * 4)

    # 60 RK4 steps of 10 K (300 -> 900 K) so that the reaction (~650 K) is inside the ramp
    arr = synthetic.arrhenius_problem(n_steps=60, dT=10.0)
    arr_names = ["log10A", "E", "n", "F"]
    run_case("arrhenius_rwmh", arr, arr_names, mh_algo("RWMH", 200, arr["prop_cov"]), 30, 10 * 200 * 4)
    run_case("arrhenius_dr", arr, arr_names, mh_algo("DR", 200, arr["prop_cov"] , gamma=0.2), 32, 10 * 200 * 4)
    run_case("arrhenius_dram", arr, arr_names,
             mh_algo("DRAM", 200, arr["prop_cov"], updating_it=10, eps_v=1e-10, gamma=0.2), 31, 10 * 200 * 4)

    # the reference's own case: heat conduction, DRAM, eps_v = 0 as in heat_conduction_1.json:33-51
    heat = synthetic.heat_problem()
    heat_names = ["a", "b", "k", "T_amb", "phi", "h"]
    run_case("heat_dram", heat, heat_names,
             mh_algo("DRAM", 500, heat["prop_cov"], updating_it=10, eps_v=0.0, gamma=0.2), 41, 6 * 500 * 2)
    run_case("heat_rwmh", heat, heat_names, mh_algo("RWMH", 300, heat["prop_cov"]), 42, 6 * 300 * 2)

    reg = synthetic.regression_problem(n_points=200, d=5)
    reg_names = ["b%d" % i for i in range(5)]
    run_case("regression_isde", reg, reg_names, isde_algo(200, 0.02, 4.0, "Analytical", 10 ** 6, 10), 51, 200 * 5)
    run_case("regression_isde_adapt", reg, reg_names, isde_algo(300, 0.02, 4.0, "Analytical", 100, 10, 250),
             52, 300 * 5)
    run_case("linear_isde_fd", lin, lin_names, isde_algo(100, 0.002, 4.0, "Numerical", 10 ** 6, 10), 53, 100 * 3)


def gen_hmc():
    reg = synthetic.regression_problem(n_points=200, d=5)
    reg_names = ["b%d" % i for i in range(5)]
    run_case("regression_hmc", reg, reg_names, hmc_algo(150, 0.01, 5, "Analytical", 10 ** 6, 10), 61, 150 * 6)
    run_case("regression_hmc_adapt", reg, reg_names, hmc_algo(400, 0.01, 4, "Analytical", 150, 25, 300), 62, 400 * 6)


def write_multi_case(workdir, probs, names, algo_block):
    """Two (or more) Data / Model entries in one posterior: solve_problem.py:186-271 pairs them by position with the
    keys of the `models` dict; the likelihood sums over them (bayesian_inference.py:500-524)."""
    p0 = probs[0]
    data, model = [], []
    for m, prob in enumerate(probs):
        x_col = prob["design"].get("x", np.arange(prob["y"].shape[1], dtype=float))
        pd.DataFrame({"x": x_col, "y": prob["y"][0], "std_y": prob["std_y"]}).to_csv(os.path.join(workdir, "data%d.csv" % m))
        data.append({"Type": "ReadFromFile", "FileName": "data%d" % m, "xField": ["x"], "yField": ["y"], "sigmaField": ["std_y"]})
        model.append({"param_names": names, "param_values": [float(v) for v in prob["theta_nom"]], "parametrization": "no"})
    prior = {}
    for k, idx in enumerate(int(i) for i in p0["unc_idx"]):
        prior[names[idx]] = {"initial_val": float(p0["x0"][k]), "prior_name": "Uniform",
                             "prior_param": [float(p0["lb"][k]), float(p0["ub"][k])]}
    js = {"Sampling": {"BayesianPosterior": {"Data": data, "Model": model, "Prior": {"Distribution": "Mixture", "Param": prior},
                                             "Likelihood": {"function": "Gaussian", "distance": "L2"}},
                       "Algorithm": algo_block}}
    with open(os.path.join(workdir, "case.json"), "w") as f:
        json.dump(js, f, indent=1)


def run_multi_case(case, probs, names, algo, seed, stream_len):
    rng = np.random.default_rng(seed)
    normals = rng.standard_normal(stream_len)
    uniforms = rng.random(stream_len)
    with tempfile.TemporaryDirectory() as wd:
        write_multi_case(wd, probs, names, algo)
        models = {"m%d" % m: build_model(prob, names) for m, prob in enumerate(probs)}
        rec = run_reference(wd, "case.json", models, normals, uniforms)
    s = rec.sampler
    out = pack_problem(probs[0])
    out["n_models"] = len(probs)
    for m, prob in enumerate(probs[1:], start=1):
        for k, v in pack_problem(prob).items():
            if k in ("theta_nom", "unc_idx", "y", "std_y") or k.startswith("design_"):
                out["m%d_%s" % (m, k)] = v
    out.update(algo_json=json.dumps(algo), normals=normals[:rec.n_normals_used], uniforms=uniforms[:rec.n_uniforms_used],
               eval_x=np.array(rec.eval_x), eval_lp=np.array(rec.eval_lp), chain=np.array(rec.saved),
               n_rejected=s.n_rejected, mean_r=float(s.mean_r))
    for attr in ("R", "V_i", "X_av_i", "inv_V", "C_approx", "P_nm", "xi_nm", "L_c", "inv_L_c"):
        if hasattr(s, attr):
            out["final_" + attr] = np.array(getattr(s, attr), dtype=np.float64)
    path = os.path.join(HERE, case + ".npz")
    np.savez_compressed(path, **out)
    print("%-22s evals=%d chain=%s n_rej=%d -> %s (%d B)" % (case, len(rec.eval_lp), np.array(rec.saved).shape, s.n_rejected,
                                                             os.path.basename(path), os.path.getsize(path)))


def gen_multi():
    """Several models in one likelihood: (1) one pyrolysis reaction observed at two heating rates (two Arrhenius models
    with their own ramp, beta and fixed F), log10A and E uncertain, DRAM; (2) two cubic data sets on different grids,
    all four coefficients uncertain, Ito-SDE with the analytical gradient summed over both models."""
    a0 = synthetic.arrhenius_problem(n_steps=40, dT=15.0)
    names = ["log10A", "E", "n", "F"]
    theta_b = np.array([6.0, 1.0e5, 1.5, 0.35])
    dg_b = dict(T0=350.0, dT=10.0, n_steps=50, beta=30.0 / 60.0)
    rng = np.random.default_rng(77)
    f_b = synthetic.arrhenius_reference(theta_b, dg_b["T0"], dg_b["dT"], dg_b["n_steps"], dg_b["beta"])
    y_b = f_b + 5e-3 * rng.standard_normal(50)
    unc = np.array([0, 1])
    common = dict(unc_idx=unc, lb=a0["lb"][:2], ub=a0["ub"][:2], x0=a0["x0"][:2], prop_cov=a0["prop_cov"][:2, :2] * 4.0)
    pa = dict(a0, **common)
    pb = dict(a0, theta_nom=theta_b, y=y_b[None, :], std_y=np.full(50, 5e-3),
              design=dict(dg_b, x=dg_b["T0"] + dg_b["dT"] * np.arange(1, 51)), **common)
    run_multi_case("multi_arrhenius_dram", [pa, pb], names,
                   mh_algo("DRAM", 200, common["prop_cov"], updating_it=10, eps_v=1e-10, gamma=0.2), 71, 10 * 200 * 2)

    c0 = synthetic.cubic_problem(n_points=60)
    c0 = dict(c0, prop_cov=c0["prop_cov"] * 10)
    xb = np.linspace(0.0, 2.0, 45)
    th = c0["theta_nom"]
    yb = ((th[3] * xb + th[2]) * xb + th[1]) * xb + th[0] + 0.1 * rng.standard_normal(45)
    c1 = dict(c0, y=yb[None, :], std_y=np.full(45, 0.1), design=dict(x=xb))
    cn = ["c0", "c1", "c2", "c3"]
    run_multi_case("multi_cubic_am", [c0, c1], cn, mh_algo("AMH", 200, c0["prop_cov"], updating_it=10, eps_v=1e-10), 72, 5 * 200 * 4)
    run_multi_case("multi_cubic_isde", [c0, c1], cn, isde_algo(150, 0.02, 4.0, "Analytical", 10 ** 6, 10), 73, 150 * 4)


def subset(prob, unc, scale_cov=1.0):
    """The same problem with only the parameters `unc` uncertain (the others stay at theta_nom)."""
    unc = np.asarray(unc)
    return dict(prob, unc_idx=unc, lb=prob["lb"][unc], ub=prob["ub"][unc], x0=prob["x0"][unc],
                prop_cov=prob["prop_cov"][np.ix_(unc, unc)] * scale_cov)


def gen_widened():
    """Round-2 widening: MAP tracking (estimate_max_distr, mha.py:226-237, :1011-1019), models with FIXED parameters
    (d < P, Model.run_model :351-366) and re-parametrised models (parametrization_* hooks, :314-336; posterior :44-64,
    gradient :573-636)."""
    lin = synthetic.linear_problem()
    lin_names = ["th0", "th1", "th2"]
    T = 300
    full_cov = np.array([[1e-3, 2e-3, -1e-3], [2e-3, 1e-2, 1e-3], [-1e-3, 1e-3, 1e-2]])
    lin2 = dict(lin, prop_cov=full_cov)
    run_case("linear_dr_map", lin2, lin_names, dict(mh_algo("DR", T, full_cov, gamma=0.2), estimate_max_distr="yes"), 81, 8 * T * 4)
    run_case("linear_am_map", lin, lin_names, dict(mh_algo("AMH", T, lin["prop_cov"], updating_it=10, eps_v=1e-10),
                                                   estimate_max_distr="yes"), 82, 4 * T * 4)
    reg = synthetic.regression_problem(n_points=200, d=5)
    reg_names = ["b%d" % i for i in range(5)]
    run_case("regression_isde_map", reg, reg_names, dict(isde_algo(150, 0.02, 4.0, "Analytical", 10 ** 6, 10), estimate_max_distr="yes"),
             83, 150 * 5)
    run_case("regression_hmc_map", reg, reg_names, dict(hmc_algo(120, 0.01, 5, "Analytical", 10 ** 6, 10), estimate_max_distr="yes"),
             84, 120 * 6)

    # ---- fixed parameters: d < P
    cub = synthetic.cubic_problem(n_points=100)
    cub = dict(cub, prop_cov=cub["prop_cov"] * 10)
    cub_names = ["c0", "c1", "c2", "c3"]
    run_case("cubic_fixed_am", subset(cub, [1, 3]), cub_names, mh_algo("AMH", T, subset(cub, [1, 3])["prop_cov"], updating_it=10, eps_v=1e-10),
             85, 5 * T * 2)
    run_case("linear_fixed_dram", subset(lin, [0, 2]), lin_names,
             mh_algo("DRAM", T, subset(lin, [0, 2])["prop_cov"], updating_it=10, eps_v=1e-10, gamma=0.2), 86, 8 * T * 2)
    run_case("cubic_fixed_isde", subset(cub, [0, 2, 3]), cub_names, isde_algo(150, 0.004, 4.0, "Analytical", 10 ** 6, 10), 87, 150 * 3)
    run_case("regression_fixed_hmc", subset(reg, [1, 4]), reg_names, hmc_algo(100, 0.01, 4, "Analytical", 10 ** 6, 10), 88, 100 * 3)

    # ---- re-parametrised models.  The prior box lives in the model space X; the chain runs in Y.
    # (1) every parameter sampled in log space, true Jacobian: Y = ln X
    lin_pos = dict(lin, lb=np.array([0.1, 0.1, 0.1]), ub=np.array([10.0, 10.0, 10.0]))
    logp = dict(kind=[1, 1, 1], a=[0.0, 0.0, 0.0], b=[1.0, 1.0, 1.0], c=[1.0, 1.0, 1.0], det_jac=True)
    cov_log = np.diag([1e-3, 2.5e-3, 1.2e-3])
    run_case("linear_logparam_am", dict(lin_pos, prop_cov=cov_log), lin_names,
             mh_algo("AMH", T, cov_log, updating_it=10, eps_v=1e-10), 91, 4 * T * 4, param=logp)
    # (2) affine scaling X = a + b Y (the `scaling_factors_parametrization` use case), DRAM
    aff = dict(kind=[0, 0, 0, 0], a=[1.0, -2.0, 0.0, 3.0], b=[0.01, 0.05, 0.05, 0.1], c=[0.0] * 4, det_jac=True)
    cov_aff = cub["prop_cov"] / np.outer(aff["b"], aff["b"])
    run_case("cubic_affine_dram", dict(cub, prop_cov=cov_aff), cub_names,
             mh_algo("DRAM", T, cov_aff, updating_it=10, eps_v=1e-10, gamma=0.2), 92, 8 * T * 4, param=aff)
    # (3) mixed kinds, base-10 exponent, and a user who left det_jac at the base class's 1
    mix = dict(kind=[0, 1, 0, 1], a=[0.0, -4.0, 0.5, 0.0], b=[1.0, 1.0, 0.2, 1.0], c=[0.0, 1.0, 0.0, float(np.log(10.0))], det_jac=False)
    cov_mix = np.diag([1e-4, 2.5e-4, 2.5e-2, 1e-4]) * 0.5
    run_case("cubic_mixed_dr", dict(cub, prop_cov=cov_mix), cub_names, mh_algo("DR", T, cov_mix, gamma=0.2), 93, 8 * T * 4, param=mix)
    # (4) analytical gradient through the change of variables (inv_jac, grad_det_jac): Ito-SDE and HMC in log space
    # (shifted exponential X = -3 + exp(Y): the regression coefficients straddle zero)
    reg_pos = dict(reg, lb=np.full(5, -2.9), ub=np.full(5, 10.0))
    logp5 = dict(kind=[1] * 5, a=[-3.0] * 5, b=[1.0] * 5, c=[1.0] * 5, det_jac=True)
    run_case("regression_logparam_isde", reg_pos, reg_names, isde_algo(150, 0.002, 4.0, "Analytical", 10 ** 6, 10), 94, 150 * 5, param=logp5)
    run_case("regression_logparam_hmc", reg_pos, reg_names, hmc_algo(100, 0.002, 4, "Analytical", 10 ** 6, 10), 95, 100 * 6, param=logp5)


def gen_hessian_fd():
    """computeHessianFD (mha.py:1115-1212) of a fixed smooth function, all four flag combinations; pins
    pybitup_b200.metropolis_hastings_algorithms.compute_hessian_fd including the in-place aliasing of :1186-1189."""
    A = np.array([[4., 1., .5, .2], [1., 3., .3, .1], [.5, .3, 2., .4], [.2, .1, .4, 1.5]])
    x = np.array([0.7, -1.3, 2.1, 0.4])
    f = lambda z: float(-0.5 * z @ A @ z + 0.1 * np.sin(z).sum() + 0.05 * z[0] * z[1] * z[2])   # noqa: E731
    out = {}
    for full in (True, False):
        for pos in (True, False):
            out["H_%d_%d" % (full, pos)] = mha.computeHessianFD(f, x, eps=0.001, compute_all_element=full, positive_diag=pos)
    np.savez(os.path.join(HERE, "_hessian_fd.npz"), A=A, x=x, **out)


def gen_near_pd():
    """nearPD (mha.py:1253-1284) on an indefinite symmetric matrix, a positive definite one and a badly scaled
    inverse-Hessian-like one; pins pybitup_b200.metropolis_hastings_algorithms.nearPD."""
    import warnings
    rng = np.random.default_rng(5)
    mats = {}
    B = rng.standard_normal((4, 4))
    mats["indefinite"] = (B + B.T) / 2.0
    mats["pd"] = B @ B.T + 0.5 * np.eye(4)
    S = np.diag([1e-3, 2.0, 40.0])
    C = np.array([[1.0, 0.95, -0.9], [0.95, 1.0, 0.2], [-0.9, 0.2, 1.0]])          # not a valid correlation matrix
    mats["scaled"] = S @ C @ S
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                     # np.matrix PendingDeprecationWarning
        for k, A in mats.items():
            out["A_" + k] = A
            out["Y_" + k] = np.asarray(mha.nearPD(A.copy()))
    np.savez(os.path.join(HERE, "_near_pd.npz"), **out)


if __name__ == "__main__":
    if "--nearpd" in sys.argv:
        gen_near_pd()
    elif "--multi" in sys.argv:
        gen_multi()
    elif "--hessian" in sys.argv:
        gen_hessian_fd()
    elif "--hmc" in sys.argv:
        gen_hmc()
    elif "--widened" in sys.argv:
        gen_widened()
    else:
        main()
        gen_hmc()
        gen_hessian_fd()
        gen_near_pd()
        gen_multi()
        gen_widened()
