"""Time the UNMODIFIED reference on a synthetic workload.  MEASUREMENT INFRASTRUCTURE ONLY (bench.py's CPU legs).

One call = one process = one chain, exactly as a user of jcoheur/pybitup would run it:
``pybitup.solve_problem.Sampling("case.json").sample({"case": model})`` (solve_problem.py:125-348) in its own working
directory, data read from the CSV the JSON names, unseeded stdlib ``random`` for the draws, chain rows written with
``np.savetxt`` every iteration (bayesian_inference.py:131-138 -- 50 % of the reference's time).
``save_eval_freq`` is set above ``n_iterations`` so that the ``.npy`` model-evaluation dumps are off.

Three variants, for the three figures bench.py reports:
  "as_shipped"          the reference as it is;
  "save_sample_stubbed" the same with BayesianPosterior.save_sample replaced by a no-op (labelled as such);
  "port"                the numpy restatement oracle/mcmc_oracle.py (no files, injected draws).

Nothing here imports the ``pybitup_b200`` package: ``synthetic.py`` (plain numpy problem builders) is loaded by file
path so that the CUDA library is never mapped into a reference-arm process.
"""
import importlib.util
import json
import os
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def load_synthetic():
    """pybitup_b200/synthetic.py as a stand-alone module (numpy only), without importing the package."""
    spec = importlib.util.spec_from_file_location("_pbu_synthetic", os.path.join(ROOT, "pybitup_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _names(prob):
    return {"linear": ["th%d" % i for i in range(len(prob["theta_nom"]))],
            "poly": ["c%d" % i for i in range(len(prob["theta_nom"]))],
            "arrhenius": ["log10A", "E", "n", "F"],
            "heat": ["a", "b", "k", "T_amb", "phi", "h"]}[prob["model"]]


def algo_block(name, n_it, prob, updating_it=10, eps_v=1e-10, gamma=0.2, isde=None):
    blk = {"name": name, "n_iterations": int(n_it), "save_eval_freq": 10 ** 9}
    if name in ("RWMH", "AMH", "DR", "DRAM"):
        blk["proposal"] = {"name": "Gaussian", "covariance": {"type": "full", "value": np.asarray(prob["prop_cov"]).tolist()}}
    if name == "AMH":
        blk["AMH"] = {"starting_it": 100, "updating_it": updating_it, "eps_v": eps_v}
    if name == "DR":
        blk["DR"] = {"gamma": gamma}
    if name == "DRAM":
        blk["DRAM"] = {"starting_it": 100, "updating_it": updating_it, "eps_v": eps_v, "gamma": gamma}
    if name == "ISDE":
        blk.update({"covariance": {"value": "Identity", "starting_it": 10 ** 9, "update_it": 10}, "gradient": "Analytical",
                    "ISDE": dict(isde or {"h": 0.05, "f0": 4.0})})
    return blk


def write_case(workdir, prob, algo):
    """data.csv + case.json in the reference's input format (test/test_pce/heat_conduction_1.json is the template)."""
    import pandas as pd
    names = _names(prob)
    N = prob["y"].shape[1]
    x_col = prob["design"].get("x", np.arange(N, dtype=float))
    pd.DataFrame({"x": x_col, "y": prob["y"][0], "std_y": prob["std_y"]}).to_csv(os.path.join(workdir, "data.csv"))
    prior = {}
    for k, idx in enumerate(int(i) for i in prob["unc_idx"]):
        prior[names[idx]] = {"initial_val": float(prob["x0"][k]), "prior_name": "Uniform",
                             "prior_param": [float(prob["lb"][k]), float(prob["ub"][k])]}
    js = {"Sampling": {
        "BayesianPosterior": {
            "Data": [{"Type": "ReadFromFile", "FileName": "data", "xField": ["x"], "yField": ["y"], "sigmaField": ["std_y"]}],
            "Model": [{"param_names": names, "param_values": [float(v) for v in prob["theta_nom"]], "parametrization": "no"}],
            "Prior": {"Distribution": "Mixture", "Param": prior},
            "Likelihood": {"function": "Gaussian", "distance": "L2"}},
        "Algorithm": algo}}
    with open(os.path.join(workdir, "case.json"), "w") as f:
        json.dump(js, f, indent=1)


def _reference_modules():
    from oracle import ref_shim
    ref_shim.install(prefer_copy=True)
    import pybitup.bayesian_inference as bi
    import pybitup.solve_problem as sp
    return bi, sp


def build_model(bi, prob):
    """The user-side ``bi.Model`` subclass of the problem (the reference ships none, pybitup_doc.tex:280): numpy
    forward models from oracle/cpu_models.py behind the reference's ``fun_x`` contract."""
    from oracle import cpu_models

    if prob["model"] == "poly":
        class M(bi.Model):
            def fun_x(self):
                return cpu_models.poly_horner(self.param, self.x)

            def d_fx_dparam(self):
                g = cpu_models.poly_grad(self.param, self.x)
                return {name: g[i] for i, name in enumerate(self.param_names)}
    elif prob["model"] == "linear":
        Phi = prob["design"]["Phi"]

        class M(bi.Model):
            def fun_x(self):
                return cpu_models.linear_design(self.param, Phi)

            def d_fx_dparam(self):
                g = cpu_models.linear_design_grad(self.param, Phi)
                return {name: g[i] for i, name in enumerate(self.param_names)}
    elif prob["model"] == "arrhenius":
        c = prob["design"]

        class M(bi.Model):
            def fun_x(self):
                return cpu_models.arrhenius_rk4(self.param, c["T0"], c["dT"], c["n_steps"], c["beta"])
    else:
        raise ValueError(prob["model"])
    M.parametrization_inv_jac = lambda self, X=1: np.eye(len(X))
    M.grad_det_jac = lambda self, X=1: np.zeros(len(X))
    return M()


def run_reference_chain(prob, algo, variant="as_shipped"):
    """One chain through the reference's own entry point; returns (iterations, seconds of Sampling.sample)."""
    bi, sp = _reference_modules()
    orig_save = bi.BayesianPosterior.save_sample
    if variant == "save_sample_stubbed":
        # no-op after the first three rows: compute_covariance re-reads the chain file at the end (mha.py:300) and needs
        # it to exist with more than one row; the handles are flushed once so that it does
        calls = [0]

        def stub(self, IO_fileID, value):
            calls[0] += 1
            if calls[0] <= 3:
                orig_save(self, IO_fileID, value)
                if calls[0] == 3:
                    for k in ("MChains", "MChains_reparam"):
                        IO_fileID[k].flush()
        bi.BayesianPosterior.save_sample = stub
    elif variant != "as_shipped":
        raise ValueError(variant)
    cwd = os.getcwd()
    try:
        with tempfile.TemporaryDirectory() as wd:
            write_case(wd, prob, algo)
            os.chdir(wd)
            model = build_model(bi, prob)
            import contextlib
            import io
            with np.errstate(all="ignore"), contextlib.redirect_stdout(io.StringIO()):
                s = sp.Sampling("case.json")
                t0 = time.perf_counter()
                s.sample({"case": model})
                dt = time.perf_counter() - t0
                s.__del__()
    finally:
        os.chdir(cwd)
        bi.BayesianPosterior.save_sample = orig_save
    return int(algo["n_iterations"]), dt


def run_port_chain(prob, algo, seed=0):
    """The same chain through the numpy restatement (oracle/mcmc_oracle.py); returns (iterations, seconds, chain)."""
    from oracle import mcmc_oracle as orc
    p = orc.Problem(model=prob["model"], theta_nom=prob["theta_nom"], unc_idx=prob["unc_idx"], y=prob["y"], std_y=prob["std_y"],
                    lb=prob["lb"], ub=prob["ub"], gamma=prob["gamma"], design=prob["design"])
    n_it, name = int(algo["n_iterations"]), algo["name"]
    d = len(prob["unc_idx"])
    rng = np.random.default_rng(seed)
    st = orc.Streams(rng.standard_normal(2 * n_it * d), rng.random(2 * n_it))
    kw = dict(adaptive=name in ("AMH", "DRAM"), delayed=name in ("DR", "DRAM"))
    if kw["adaptive"]:
        kw.update(updating_it=int(algo[name]["updating_it"]), eps_v=float(algo[name]["eps_v"]))
    if kw["delayed"]:
        kw.update(gamma_dr=float(algo[name]["gamma"]))
    t0 = time.perf_counter()
    with np.errstate(all="ignore"):
        out = orc.run_mh(p, prob["x0"], prob["prop_cov"], n_it, st, **kw)
    return n_it, time.perf_counter() - t0, out["chain"]


# ---- one workload, one worker per core -----------------------------------------------------------------------------
def worker(args):
    """(builder name, builder kwargs, algorithm name, iterations, variant, seed) -> (iterations, seconds)."""
    builder, bkw, algo_name, n_it, variant, seed = args
    syn = load_synthetic()
    prob = getattr(syn, builder)(**bkw)
    prob = dict(prob, x0=syn.chain_starts(prob["x0"], 1, jitter=0.01, seed=seed)[0])
    algo = algo_block(algo_name, n_it, prob)
    if variant == "port":
        n, dt, _ = run_port_chain(prob, algo, seed)
        return n, dt
    return run_reference_chain(prob, algo, variant)


def rate(pool, n_procs, builder, bkw, algo_name, n_it, variant):
    """chain-steps/s of n_procs independent single-chain processes running concurrently (rate = all iterations over the
    time of the slowest process), and the wall time of the whole call."""
    jobs = [(builder, bkw, algo_name, n_it, variant, s) for s in range(n_procs)]
    t0 = time.perf_counter()
    res = [worker(jobs[0])] if pool is None else pool.map(worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    return sum(n for n, _ in res) / max(dt for _, dt in res), wall
