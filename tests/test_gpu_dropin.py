"""End to end through the reference's Python surface: JSON input -> Sampling(...).sample(models) /
Propagation(...).propagate(models) with a bi.Model subclass, on the CUDA engine.  Checks the reference's output files
 and the statistical known answers of the reference's own heat-conduction case (§8c)."""
import json
import os

import numpy as np
import pandas as pd
import pytest

from oracle import cpu_models

pytestmark = pytest.mark.gpu

# posterior moments of the heat-conduction case by grid quadrature (tests/golden/gen_heat_posterior.py)
TRUE_MEAN = np.array([-18.418010706591474, 0.0019146479526173403])
TRUE_COV = np.array([[0.023064064360743076, -2.2297224870054396e-06], [-2.2297224870054396e-06, 2.309329769136012e-10]])
# moments of the reference's saved DRAM chain (pybitup/test/test_pce/mcmc_chain.csv): one chain of 1e4
# iterations, effective sample size ~146 (batch means), Monte-Carlo standard error 0.0165 on phi and 1.7e-6 on h
REF_MEAN = np.array([-18.365, 1.9091e-3])
REF_SE = np.array([0.0165, 1.7e-6])

HEAT_JSON = """
/* heat conduction, after pybitup/test/test_pce/heat_conduction_1.json */
{
  "Sampling": {
    "BayesianPosterior": {
      "Data": [{"Type": "ReadFromFile", "FileName": "heat_conduction_data", "xField": ["x"], "yField": ["T"], "sigmaField": ["std_T"]}],
      "Model": [{"param_names": ["a", "b", "k", "T_amb", "phi", "h"],
                 "param_values": [0.95, 0.95, 2.37, 21.29, -18.41, 0.00191], "parametrization": "no"}],
      "Prior": {"Distribution": "Mixture",
                "Param": {"phi": {"initial_val": -18.41, "prior_name": "Uniform", "prior_param": [-100, 0.0]},
                          "h": {"initial_val": 0.00191, "prior_name": "Uniform", "prior_param": [0.00001, 1.0]}}},
      "Likelihood": {"function": "Gaussian", "distance": "L2"}
    },
    "Algorithm": {
      "name": "DRAM", // the reference's own test uses DRAM
      "DRAM": {"starting_it": 1e2, "updating_it": 1e1, "eps_v": 1e-12, "gamma": 0.2},
      "n_iterations": 4e3,
      "n_chains": 64, "seed": 11, "save_eval_freq": 1000,
      "proposal": {"name": "Gaussian", "covariance": {"type": "full", "value": [[2.1034e-2, -2.0286e-6], [-2.0286e-6, 2.0972e-10]]}}
    }
  },
  "Propagation": {
    "Model": [{"model_id": "heat", "design_points": {"filename": "design_points.csv", "field": "x"},
               "param_names": ["a", "b", "k", "T_amb", "phi", "h"],
               "param_values": [0.95, 0.95, 2.37, 21.29, -18.41, 0.00191], "parametrization": "no", "emulator": "None"}],
    "Uncertain_param": {"phi": {"filename": "output/mcmc_chain.csv", "field": 0}, "h": {"filename": "output/mcmc_chain.csv", "field": 1}},
    "Model_evaluation": {}
  }
}
"""


def _heat_model():
    from pybitup_b200 import bayesian_inference as bi

    class HeatConduction(bi.Model):
        device_model = "heat"

        def fun_x(self):
            return cpu_models.heat_conduction(self.param, np.asarray(self.x, dtype=np.float64))

    return HeatConduction()


def test_sampling_and_propagation_from_json(tmp_path, monkeypatch):
    from pybitup_b200 import synthetic
    from pybitup_b200.solve_problem import Sampling, Propagation
    monkeypatch.chdir(tmp_path)
    p = synthetic.heat_problem()
    pd.DataFrame({"x": p["design"]["x"], "T": p["y"][0], "std_T": p["std_y"]}).to_csv("heat_conduction_data.csv", index=False)
    pd.DataFrame({"x": np.linspace(10.0, 66.0, 29)}).to_csv("design_points.csv", index=False)
    with open("heat.json", "w") as f:
        f.write(HEAT_JSON)
    model = _heat_model()
    job = Sampling("heat.json")
    run = job.sample({"heat": model})
    job.close()
    T = 4000
    chain = pd.read_csv("output/mcmc_chain.csv", header=None).values
    assert chain.shape == (T + 1, 2) and np.allclose(chain[0], [-18.41, 0.00191])
    assert pd.read_csv("output/mcmc_chain_reparam.csv", header=None).values.shape == (T + 1, 2)
    lines = open("output/output.txt").read().splitlines()
    assert lines[0] == "$RandomVarName$" and lines[1].split() == ["phi", "h"] and int(lines[3]) == T     # post_process.py:41-50
    assert "Rejection rate is" in "\n".join(lines) and "Covariance Matrix is" in "\n".join(lines)
    assert np.loadtxt("cov.csv", delimiter=",").shape == (2, 2)
    ev = np.load("output/model_eval/heat_fun_eval-1000.npy")
    assert ev.shape == (15,) and np.allclose(ev, cpu_models.heat_conduction(
        np.array([0.95, 0.95, 2.37, 21.29, run.chain[1000, 0], run.chain[1000, 1]]), p["design"]["x"]), rtol=1e-11)
    assert os.path.exists("output/data.bin")
    # the caller's initial vector was updated in place to the final state (mha.py:95-99, :215)
    assert np.array_equal(run.param_init, run.chain[-1]) and not np.array_equal(run.chain[-1], [-18.41, 0.00191])
    # statistical known answer: pooled posterior moments of 64 chains vs the reference's saved chain
    st = run.statistics
    assert st["n_chains"] == 64 and st["n"] == 64 * T
    assert np.all(np.abs(st["mean"] - TRUE_MEAN) < 5 * np.sqrt(np.diag(TRUE_COV)) / np.sqrt(64 * T / 30.0))
    assert np.all(np.abs(st["cov"] - TRUE_COV) < 0.1 * np.abs(TRUE_COV))
    assert np.all(np.abs(st["mean"] - REF_MEAN) < 4 * REF_SE)          # within the reference chain's own Monte-Carlo error
    assert np.all(st["rhat"] < 1.1)
    assert 0.3 < 1.0 - run.n_rejected / T < 0.99

    # ---- propagation of the chain through the model (solve_problem.py:467-647)
    res = Propagation("heat.json").propagate({"heat": _heat_model()})["heat"]
    xs = pd.read_csv("output/mcmc_chain.csv").values          # first row becomes the header, as in the reference (Q15)
    x_design = np.linspace(10.0, 66.0, 29)
    fe = np.array([cpu_models.heat_conduction(np.array([0.95, 0.95, 2.37, 21.29, s[0], s[1]]), x_design) for s in xs])
    assert res["n_samples"] == T
    assert np.allclose(res["mean"], fe.mean(axis=0), rtol=1e-11)
    assert np.allclose(res["ci_lb"], np.percentile(fe, 2.5, axis=0), rtol=1e-10)
    assert np.allclose(res["ci_ub"], np.percentile(fe, 97.5, axis=0), rtol=1e-10)
    ci = pd.read_csv("output/heat_CI.csv")
    iv = pd.read_csv("output/heat_interval.csv")
    assert list(ci.columns) == ["x", "CI_lb", "CI_ub"] and list(iv.columns) == ["x", "mean", "lower_bound", "upper_bound"]
    assert np.allclose(iv["upper_bound"], fe.max(axis=0), rtol=1e-11) and len(ci) == 29


@pytest.mark.parametrize("precond", ["Hessian_diag", "from_file", "Hessian", "nearPD"])
def test_isde_from_json_on_cubic(tmp_path, monkeypatch, precond):
    from pybitup_b200 import synthetic, bayesian_inference as bi
    from pybitup_b200.solve_problem import Sampling
    monkeypatch.chdir(tmp_path)
    p = synthetic.cubic_problem(n_points=400)
    pd.DataFrame({"x": p["design"]["x"], "y": p["y"][0], "s": p["std_y"]}).to_csv("cubic.csv", index=False)

    class Cubic(bi.Model):
        device_model = "poly"

    # "Hessian" / "nearPD": the reference's full finite-difference Hessian (aliased stencil) need not be
    # positive definite; GradientBasedMCMC then falls back to the identity (mha.py:713-722), for which only a short,
    # small-step run is meaningful here
    exact = precond in ("Hessian_diag", "from_file")
    T = 1500 if exact else 60
    cfg = {"Sampling": {"BayesianPosterior": {
        "Data": [{"Type": "ReadFromFile", "FileName": "cubic", "xField": ["x"], "yField": ["y"], "sigmaField": ["s"]}],
        "Model": [{"param_names": ["c0", "c1", "c2", "c3"], "param_values": [1.0, -2.0, 0.5, 3.0]}],
        "Prior": {"Distribution": "Mixture", "Param": {n: {"initial_val": v, "prior_name": "Uniform", "prior_param": [-10, 10]}
                                                      for n, v in zip(["c0", "c1", "c2", "c3"], [1.0, -2.0, 0.5, 3.0])}},
        "Likelihood": {"function": "Gaussian"}},
        "Algorithm": {"name": "ISDE", "n_iterations": T, "n_chains": 32, "seed": 5, "save_eval_freq": 1e9,
                      "covariance": {"value": precond, "starting_it": 200, "update_it": 50, "end_it": 1000},
                      "gradient": "Analytical", "ISDE": {"h": 0.3 if exact else 1e-4, "f0": 4.0}}}}
    with open("cubic.json", "w") as f:
        json.dump(cfg, f)
    x = p["design"]["x"]
    if precond == "from_file":          # ./cov.csv of an earlier run (mha.py:694-697): here the exact posterior covariance
        Phi0 = np.stack([x ** j for j in range(4)], axis=1)
        np.savetxt("cov.csv", np.linalg.inv(Phi0.T @ Phi0) * 0.1 ** 2, delimiter=",")
    job = Sampling("cubic.json")
    run = job.sample({"cubic": Cubic()})
    job.close()
    assert pd.read_csv("output/mcmc_chain.csv", header=None).values.shape == (T + 1, 4)
    aux = pd.read_csv("output/aux_variables.csv", header=None).values
    # T + 1 blocks of d rows: (0, xi_0) from ito_SDE.__init__ (mha.py:921), then (P_n, xi_n) after every update (:1032)
    assert aux.shape == ((T + 1) * 4, 2) and np.all(aux[:4, 0] == 0.0)
    assert np.allclose(aux[-4:, 1], run.chain[-1], atol=1e-6)
    if not exact:
        assert np.all(np.isfinite(run.chain)) and run.C_approx.shape == (4, 4)
        assert np.all(np.linalg.eigvalsh((run.C_approx + run.C_approx.T) / 2) > 0)      # a factorisable matrix, or the identity
        return
    # posterior of a linear-Gaussian problem: mean = least squares solution, cov = (Phi^T Phi)^-1 sigma^2
    Phi = np.stack([x ** j for j in range(4)], axis=1)
    cov = np.linalg.inv(Phi.T @ Phi) * 0.1 ** 2
    mean = cov @ Phi.T @ p["y"][0] / 0.1 ** 2
    st = run.statistics
    assert np.all(np.abs(st["mean"] - mean) < 4 * np.sqrt(np.diag(cov)) / np.sqrt(32 * 20))
    assert np.all(np.abs(np.diag(st["cov"]) / np.diag(cov) - 1.0) < 0.35)


@pytest.mark.parametrize("name", ["Gaussian", "SoizePDF"])
def test_sampling_a_known_distribution(tmp_path, monkeypatch, name):
    """Sampling.Distribution (solve_problem.py:134-164): the reference's sampler test bed, on the device."""
    from pybitup_b200.solve_problem import Sampling
    from pybitup_b200 import distributions
    monkeypatch.chdir(tmp_path)
    mean, cov = [1.0, -2.0, 0.5], [[2.0, 0.3, 0.1], [0.3, 1.0, 0.2], [0.1, 0.2, 0.5]]
    if name == "Gaussian":
        dist = {"name": "Gaussian", "hyperparameters": [[mean], cov], "n_rand_var": 3, "init_val": mean}
        prop = (np.array(cov) * 0.8).tolist()
    else:
        dist = {"name": "SoizePDF", "hyperparameters": [], "n_rand_var": 2, "init_val": [0.0, 0.3]}
        prop = [[0.2, 0.0], [0.0, 0.2]]
    T = 3000
    cfg = {"Sampling": {"Distribution": dist,
                        "Algorithm": {"name": "AMH", "AMH": {"starting_it": 100, "updating_it": 50, "eps_v": 1e-8}, "n_iterations": T,
                                      "n_chains": 256, "seed": 3, "proposal": {"covariance": {"type": "full", "value": prop}}}}}
    with open("dist.json", "w") as f:
        json.dump(cfg, f)
    job = Sampling("dist.json")
    run = job.sample()
    job.close()
    d = len(dist["init_val"])
    assert pd.read_csv("output/mcmc_chain.csv", header=None).values.shape == (T + 1, d)
    st = run.statistics
    if name == "Gaussian":
        se = np.sqrt(np.diag(cov) / (256 * T / 40.0))
        assert np.all(np.abs(st["mean"] - mean) < 5 * se)
        assert np.all(np.abs(st["cov"] - np.array(cov)) < 0.1 * np.sqrt(np.outer(np.diag(cov), np.diag(cov))))
        # log-density on the device vs scipy (what the reference's Gaussian.compute_log_value returns, :234-236)
        import scipy.stats
        from pybitup_b200.engine import ChainEngine, SamplerConfig
        g = distributions.set_probability_dist(["Gaussian"], [[mean], cov], 3)
        X = np.random.default_rng(0).standard_normal((20, 3))
        with ChainEngine(g.lower(), SamplerConfig(), 1) as eng:
            assert np.allclose(eng.log_posterior(X), scipy.stats.multivariate_normal(mean, cov).logpdf(X), rtol=1e-12)
    else:
        # moments of the Soize density by quadrature on a fine grid
        s = distributions.SoizePDF([], 2)
        g0, g1 = np.meshgrid(np.linspace(-2.5, 2.5, 1201), np.linspace(-2.5, 3.0, 1201), indexing="ij")
        w = np.exp(-15 * (g0 ** 3 - g1) ** 2 - (g1 - 0.3) ** 4)
        w /= w.sum()
        m = np.array([(w * g0).sum(), (w * g1).sum()])
        v = np.array([(w * (g0 - m[0]) ** 2).sum(), (w * (g1 - m[1]) ** 2).sum()])
        assert np.all(np.abs(st["mean"] - m) < 0.05) and np.all(np.abs(np.diag(st["cov"]) / v - 1) < 0.15)
        from pybitup_b200.engine import ChainEngine, SamplerConfig
        X = np.random.default_rng(1).uniform(-1.5, 1.5, (30, 2))
        with ChainEngine(s.lower(), SamplerConfig(), 1) as eng:
            assert np.allclose(eng.log_posterior(X), [s.compute_log_value(x) for x in X], rtol=1e-13, atol=1e-13)


def test_hmc_from_json(tmp_path, monkeypatch):
    from pybitup_b200 import synthetic, bayesian_inference as bi
    from pybitup_b200.solve_problem import Sampling
    monkeypatch.chdir(tmp_path)
    p = synthetic.cubic_problem(n_points=300)
    pd.DataFrame({"x": p["design"]["x"], "y": p["y"][0], "s": p["std_y"]}).to_csv("cubic.csv", index=False)

    class Cubic(bi.Model):
        device_model = "poly"

    T = 800
    names = ["c0", "c1", "c2", "c3"]
    cfg = {"Sampling": {"BayesianPosterior": {
        "Data": [{"Type": "ReadFromFile", "FileName": "cubic", "xField": ["x"], "yField": ["y"], "sigmaField": ["s"]}],
        "Model": [{"param_names": names, "param_values": [1.0, -2.0, 0.5, 3.0]}],
        "Prior": {"Distribution": "Mixture", "Param": {n: {"initial_val": v, "prior_name": "Uniform", "prior_param": [-10, 10]}
                                                      for n, v in zip(names, [1.0, -2.0, 0.5, 3.0])}},
        "Likelihood": {"function": "Gaussian"}},
        "Algorithm": {"name": "HMC", "n_iterations": T, "n_chains": 64, "seed": 8, "save_eval_freq": 1e9,
                      "covariance": {"value": "Hessian_diag", "starting_it": 100, "update_it": 50, "end_it": 500},
                      "gradient": "Analytical", "HMC": {"step_size": 0.4, "num_steps": 4}}}}
    with open("hmc.json", "w") as f:
        json.dump(cfg, f)
    job = Sampling("hmc.json")
    run = job.sample({"cubic": Cubic()})
    job.close()
    assert pd.read_csv("output/mcmc_chain.csv", header=None).values.shape == (T + 1, 4)
    x = p["design"]["x"]
    Phi = np.stack([x ** j for j in range(4)], axis=1)
    cov = np.linalg.inv(Phi.T @ Phi) * 0.1 ** 2
    mean = cov @ Phi.T @ p["y"][0] / 0.1 ** 2
    st = run.statistics
    assert np.all(np.abs(st["mean"] - mean) < 5 * np.sqrt(np.diag(cov)) / np.sqrt(64 * 20))
    assert np.all(np.abs(np.diag(st["cov"]) / np.diag(cov) - 1.0) < 0.3)
    assert 0.3 < 1.0 - run.n_rejected / T <= 1.0


def test_two_models_in_one_posterior_from_json(tmp_path, monkeypatch):
    """Two Data / Model entries (one reaction observed at two heating rates): the likelihood sums over both models
    (bayesian_inference.py:500-524); per-model .npy evaluations; posterior concentrates on the true (log10A, E)."""
    from pybitup_b200 import synthetic, bayesian_inference as bi
    from pybitup_b200.solve_problem import Sampling
    monkeypatch.chdir(tmp_path)
    names = ["log10A", "E", "n", "F"]
    truth = np.array([6.0, 1.0e5, 1.5, 0.3])
    ramps = [dict(T0=300.0, dT=15.0, n_steps=40, beta=10.0 / 60.0), dict(T0=350.0, dT=10.0, n_steps=50, beta=30.0 / 60.0)]
    rng = np.random.default_rng(3)

    class Pyro(bi.Model):
        device_model = "arrhenius"
        ramp = None

        def device_design(self):
            return dict(self.ramp)

    models, data_blk = {}, []
    for m, r in enumerate(ramps):
        f = synthetic.arrhenius_reference(truth, r["T0"], r["dT"], r["n_steps"], r["beta"])
        T = r["T0"] + r["dT"] * np.arange(1, r["n_steps"] + 1)
        pd.DataFrame({"T": T, "y": f + 5e-3 * rng.standard_normal(len(T)), "s": np.full(len(T), 5e-3)}).to_csv("rate%d.csv" % m, index=False)
        data_blk.append({"Type": "ReadFromFile", "FileName": "rate%d" % m, "xField": ["T"], "yField": ["y"], "sigmaField": ["s"]})
        mod = Pyro()
        mod.ramp = r
        models["rate%d" % m] = mod
    n_it = 1500
    cfg = {"Sampling": {"BayesianPosterior": {
        "Data": data_blk,
        "Model": [{"param_names": names, "param_values": [6.0, 1.0e5, 1.5, 0.3]}] * 2,
        "Prior": {"Distribution": "Mixture", "Param": {
            "log10A": {"initial_val": 6.0, "prior_name": "Uniform", "prior_param": [4.2, 7.8]},
            "E": {"initial_val": 1.0e5, "prior_name": "Uniform", "prior_param": [0.7e5, 1.3e5]}}},
        "Likelihood": {"function": "Gaussian"}},
        # (adaptive Metropolis rather than DRAM: on this ridge-shaped posterior the reference's delayed-rejection ratio,
        # with its element-wise inv_V -- SURVEY Q2 -- underflows to 0/0 = nan, which min(1, nan) accepts, Q6)
        "Algorithm": {"name": "AMH", "n_iterations": n_it, "n_chains": 64, "seed": 9, "save_eval_freq": 300,
                      "proposal": {"covariance": {"type": "diag", "value": [(1e-4 * 6.0) ** 2, (1e-4 * 1e5) ** 2]}},
                      "AMH": {"starting_it": 10, "updating_it": 50, "eps_v": 1e-12}}}}
    with open("pyro.json", "w") as f:
        json.dump(cfg, f)
    job = Sampling("pyro.json")
    run = job.sample(models)
    job.close()
    chain = pd.read_csv("output/mcmc_chain.csv", header=None).values
    assert chain.shape == (n_it + 1, 2)
    for m, r in enumerate(ramps):
        ev = np.load("output/model_eval/rate%d_fun_eval-300.npy" % m)
        assert ev.shape == (r["n_steps"],) and np.all(ev <= 1.0) and ev[-1] < 0.75
    st = run.statistics
    assert not np.any(run.status), "status flags (nan ratio / not PD) raised"
    assert abs(st["mean"][0] - 6.0) < 0.1 and abs(st["mean"][1] - 1.0e5) < 1500.0, st["mean"]
    # the two heating rates together pin the (log10A, E) ridge tighter than either alone would
    assert np.sqrt(st["cov"][0, 0]) < 0.05


def test_reparametrised_model_with_fixed_parameter_and_map_from_json(tmp_path, monkeypatch):
    """A user model with its own parametrization_* hooks (log sampling, bayesian_inference.py:314-336), one FIXED
    parameter (d = 3 < P = 4, Model.run_model :351-366), estimate_max_distr and estimate_sigma, through
    Sampling(json).sample(models): chain files in both spaces, MAP / sigma files (mha.py:332-355, solve_problem.py:90-111),
    and the posterior of the linear-Gaussian problem in the MODEL space."""
    from pybitup_b200 import synthetic, bayesian_inference as bi
    from pybitup_b200.solve_problem import Sampling
    monkeypatch.chdir(tmp_path)
    p = synthetic.cubic_problem(n_points=400)
    pd.DataFrame({"x": p["design"]["x"], "y": p["y"][0], "s": p["std_y"]}).to_csv("cubic.csv", index=False)

    class CubicLog(bi.Model):           # as a user of the reference writes it: sample log(c) for the positive coefficients
        device_model = "poly"

        def parametrization_forward(self, X=1, P=1):
            return np.log(X)

        def parametrization_backward(self, Y=1, P=1):
            return np.exp(Y)

        def parametrization_det_jac(self, X=1):
            return np.prod(1.0 / np.asarray(X))

    names, vals = ["c0", "c1", "c2", "c3"], [1.0, -2.0, 0.5, 3.0]
    unc = ["c0", "c2", "c3"]                                     # c1 stays at its nominal value
    T = 3000
    cfg = {"Sampling": {"BayesianPosterior": {
        "Data": [{"Type": "ReadFromFile", "FileName": "cubic", "xField": ["x"], "yField": ["y"], "sigmaField": ["s"]}],
        "Model": [{"param_names": names, "param_values": vals}],
        "Prior": {"Distribution": "Mixture", "Param": {n: {"initial_val": vals[names.index(n)], "prior_name": "Uniform",
                                                           "prior_param": [1e-3, 10]} for n in unc}},
        "Likelihood": {"function": "Gaussian"}},
        "Algorithm": {"name": "AMH", "AMH": {"starting_it": 100, "updating_it": 20, "eps_v": 1e-10}, "n_iterations": T,
                      "n_chains": 128, "seed": 9, "save_eval_freq": 1e9, "save_all_chains": "yes", "estimate_max_distr": "yes",
                      "estimate_sigma": "yes",
                      "proposal": {"covariance": {"type": "diag", "value": [2e-5, 2e-4, 2e-5]}}}}}
    with open("cubic.json", "w") as f:
        json.dump(cfg, f)
    job = Sampling("cubic.json")
    run = job.sample({"cubic": CubicLog()})
    job.close()
    assert run.problem.param is not None and list(run.problem.param["kind"]) == [1, 1, 1] and list(run.problem.unc_idx) == [0, 2, 3]
    X = pd.read_csv("output/mcmc_chain.csv", header=None).values
    Y = pd.read_csv("output/mcmc_chain_reparam.csv", header=None).values
    assert X.shape == (T + 1, 3) and np.allclose(X[0], [1.0, 0.5, 3.0]) and np.allclose(X, np.exp(Y), atol=2e-6)
    # MAP files: arg max in the MODEL space, value = exp(log MAP) (mha.py:352-355)
    arg_map = np.loadtxt("output/arg_MAP_estimation.csv", delimiter=",")
    assert np.allclose(arg_map, np.exp(run.arg_MAP), atol=1e-6)
    assert run.MAP_val >= run.MAP_val_all[0] - 1e-12 and run.arg_MAP_all.shape == (128, 3)
    assert np.isclose(np.loadtxt("output/MAP_estimation.csv"), np.exp(run.MAP_val), atol=1e-6)
    sig = pd.read_csv("output/estimated_sigma.csv")
    assert np.allclose(sig["model_id"].values, 0.1)             # the unchanged std_y
    # posterior in the model space: c1 fixed -> least squares of y + 2x on (1, x^2, x^3)
    x = p["design"]["x"]
    Phi = np.stack([np.ones_like(x), x ** 2, x ** 3], axis=1)
    cov = np.linalg.inv(Phi.T @ Phi) * 0.1 ** 2
    mean = cov @ Phi.T @ (p["y"][0] + 2.0 * x) / 0.1 ** 2
    allc = np.exp(run.chains_all[T // 2:])                      # [T/2, d, C] in the model space
    emp_mean = allc.mean(axis=(0, 2))
    sd = np.sqrt(np.diag(cov))
    assert np.all(np.abs(emp_mean - mean) < 0.15 * sd), (emp_mean, mean, sd)
    emp_sd = np.sqrt(np.mean((allc - emp_mean[None, :, None]) ** 2, axis=(0, 2)))
    assert np.all(np.abs(emp_sd / sd - 1.0) < 0.15), (emp_sd, sd)
    # the device MAP is where the posterior density in Y peaks: prior(X) |dX/dY| like(X) -> close to the least-squares point
    best = np.exp(run.arg_MAP_all[np.argmax(run.MAP_val_all)])
    assert np.all(np.abs(best - mean) < 1.0 * sd)
