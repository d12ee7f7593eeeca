"""K4: chain-moment sums on the device, pooled covariance / R-hat, pooled-covariance proposal."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


def _reference(ch):          # ch [C, T, d]
    C, T, d = ch.shape
    allx = ch.reshape(-1, d)
    W = ch.var(axis=1, ddof=1).mean(axis=0)
    B_over_n = ch.mean(axis=1).var(axis=0, ddof=1)
    return allx.mean(axis=0), np.cov(allx, rowvar=False), np.sqrt(((T - 1) / T * W + B_over_n) / W)


@pytest.mark.parametrize("algo", ["RWMH", "ISDE"])
def test_chain_sums_match_numpy(algo):
    from pybitup_b200 import synthetic, diagnostics
    from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig
    p = synthetic.cubic_problem(n_points=120)
    C, T, thin = 300, 120, 3
    cfg = SamplerConfig(algo=algo, seed=4, isde_h=0.01, isde_f0=4.0)
    with ChainEngine(ProblemSpec.from_dict(p), cfg, C) as eng:
        eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=0.02, seed=1), p["prop_cov"] if algo == "RWMH" else None)
        r1 = eng.run(T // 2, thin=thin)
        r2 = eng.run(T // 2, thin=thin)          # moments keep accumulating across launches
        eng.synchronize()
        ch = np.concatenate([r1.samples.cpu().numpy(), r2.samples.cpu().numpy()], axis=0).transpose(2, 0, 1)   # [C, T/thin, d]
        st = eng.statistics()
        mean = ch.mean(axis=1)
        dev = ch - mean[:, None, :]
        sums_np = diagnostics.chain_sums_numpy(np.full(C, ch.shape[1]), mean, np.einsum("cti,ctj->cij", dev, dev), p["x0"])
        assert H.rel_err(eng.chain_sums(p["x0"]), sums_np) < 1e-9
        m, cov, rhat = _reference(ch)
        assert H.rel_err(st["mean"], m) < 1e-12 and H.rel_err(st["cov"], cov) < 1e-9 and H.rel_err(st["rhat"], rhat) < 1e-9
        assert st["n_chains"] == C and st["n"] == C * ch.shape[1]
        eng.reset_moments()
        assert eng.chain_sums()[1] == 0.0


def test_pooled_covariance_proposal():
    """Pooled-covariance mode: after set_proposal every chain proposes with the Cholesky factor of the given matrix."""
    from pybitup_b200 import synthetic
    from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig
    p = synthetic.linear_problem()
    C = 64
    with ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo="RWMH", seed=2), C) as eng:
        eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=1e-3), p["prop_cov"])
        eng.run(200, thin=1, keep_samples=False)
        st = eng.statistics()
        V = 2.38 ** 2 / 3 * st["cov"]
        eng.set_proposal(V)
        R = eng.proposal()["R"]
        assert H.rel_err(R[0].T @ R[0], V) < 1e-12 and np.array_equal(R[0], R[-1])
        rej0 = eng.counters()["n_rejected"].sum()
        eng.run(200, thin=1, keep_samples=False)
        acc = 1.0 - (eng.counters()["n_rejected"].sum() - rej0) / (200.0 * C)
        # (the proposal is cur + R @ z with R upper, the reference's convention, SURVEY Q1: not tuned for correlation)
        assert 0.01 < acc < 0.9
        with pytest.raises(np.linalg.LinAlgError):
            eng.set_proposal(-np.eye(3))


@pytest.mark.parametrize("algo", ["RWMH", "ISDE"])
def test_pooled_statistics_on_device_and_apply(algo):
    """The device-resident pipeline (two passes over the per-chain moments + finalisation kernel): statistics equal to numpy
    over all kept states; apply=True turns the pooled covariance into every chain's proposal (MH: R = chol(S_d (cov + eps
    I)), mha.py:463, :486-487) or preconditioner (ISDE: C_approx, inv_L_c, mha.py:745-746, :773-777) without a host copy."""
    import scipy.linalg
    from pybitup_b200 import synthetic
    from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig
    p = synthetic.cubic_problem(n_points=120)
    C, T, d = 5000, 60, 4           # three chain slices in the reduction
    cfg = SamplerConfig(algo=algo, seed=4, isde_h=0.01, isde_f0=4.0, eps_v=1e-9)
    with ChainEngine(ProblemSpec.from_dict(p), cfg, C) as eng:
        eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=0.02, seed=1), p["prop_cov"] if algo == "RWMH" else None)
        res = eng.run(T, thin=1)
        stats = eng.pooled_statistics(apply=True)               # queued behind the launch; no synchronisation here
        st = eng.statistics_dict(stats)
        ch = res.samples.cpu().numpy().transpose(2, 0, 1)       # [C, T, d]
        m, cov, rhat = _reference(ch)
        assert st["n_chains"] == C and st["n"] == C * T and st["proposal_ok"]
        assert H.rel_err(st["mean"], m) < 1e-12 and H.rel_err(st["cov"], cov) < 1e-9 and H.rel_err(st["rhat"], rhat) < 1e-9
        scale, eps = (2.38 ** 2 / d, 1e-9) if algo == "RWMH" else (1.0, 1e-10)
        V = scale * (cov + eps * np.eye(d))
        assert H.rel_err(st["V"], V) < 1e-9
        assert H.rel_err(st["R"], scipy.linalg.cholesky(st["V"])) < 1e-12
        assert H.rel_err(st["inv_R"] @ st["R"], np.eye(d)) < 1e-9 or np.allclose(st["inv_R"] @ st["R"], np.eye(d), atol=1e-12)
        prop = eng.proposal()
        if algo == "RWMH":
            assert np.array_equal(prop["R"][0], st["R"]) and np.array_equal(prop["R"][-1], st["R"])
            # the next launch proposes with the pooled factor (it reaches the non-adaptive kernel as a parameter)
            x_before, _ = eng.state()
            eng.set_streams(np.ones((C, d)), np.zeros((C, 1)) + 0.0)       # z = 1, u = 0: accept whatever is inside the box
            eng.run(1, keep_samples=False)
            x_after, _ = eng.state()
            assert H.rel_err(x_after - x_before, np.tile(st["R"] @ np.ones(d), (C, 1))) < 1e-9
        else:
            assert np.array_equal(prop["R"][0], st["inv_R"]) and np.array_equal(prop["R"][-1], st["inv_R"])
            assert np.array_equal(eng.isde_state()["C"][C // 2], st["V"])
        # a second call reuses the persistent buffers and keeps accumulating
        eng.run(10, thin=1, keep_samples=False)
        assert eng.statistics()["n"] == C * (T + 10 + (1 if algo == "RWMH" else 0))


def test_pooled_statistics_refuses_a_singular_covariance():
    """All chains at the same point: the pooled covariance is zero, eps = 0 -> not positive definite; the proposal stays."""
    from pybitup_b200 import synthetic
    from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig
    p = synthetic.linear_problem()
    with ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo="RWMH", seed=1, eps_v=0.0), 32) as eng:
        eng.init_chains(p["x0"], p["prop_cov"])
        eng.set_streams(np.zeros((32, 3 * 4)), np.ones((32, 4)))       # z = 0: nobody moves
        eng.run(4, keep_samples=False)
        R0 = eng.proposal()["R"][0]
        st = eng.statistics_dict(eng.pooled_statistics(apply=True))
        assert not st["proposal_ok"] and np.array_equal(eng.proposal()["R"][0], R0)
        eng.set_streams(None, None)
        eng.run(4, keep_samples=False)                                   # still runs, with the old factor


def test_json_driven_pooled_covariance_modes(tmp_path, monkeypatch):
    """"pooled_covariance": "yes" through the JSON surface: AM with the covariance of all chains every updating_it
    iterations, Ito-SDE with the pooled preconditioner inside its adaptation window."""
    import json
    import pandas as pd
    from pybitup_b200 import synthetic, bayesian_inference as bi
    from pybitup_b200.solve_problem import Sampling
    monkeypatch.chdir(tmp_path)
    p = synthetic.cubic_problem(n_points=200)

    class Poly(bi.Model):
        device_model = "poly"

    pd.DataFrame({"x": p["design"]["x"], "y": p["y"][0], "std_y": p["std_y"]}).to_csv("data.csv", index=False)
    names = ["c0", "c1", "c2", "c3"]
    prior = {n: {"initial_val": float(p["x0"][i]), "prior_name": "Uniform", "prior_param": [-10.0, 10.0]} for i, n in enumerate(names)}
    base = {"Data": [{"Type": "ReadFromFile", "FileName": "data", "xField": ["x"], "yField": ["y"], "sigmaField": ["std_y"]}],
            "Model": [{"param_names": names, "param_values": [float(v) for v in p["theta_nom"]], "parametrization": "no"}],
            "Prior": {"Distribution": "Mixture", "Param": prior}, "Likelihood": {"function": "Gaussian", "distance": "L2"}}
    algos = {
        "am": {"name": "AMH", "n_iterations": 1500, "n_chains": 512, "seed": 3, "init_jitter": 0.01, "pooled_covariance": "yes",
               "save_eval_freq": 10 ** 9, "AMH": {"starting_it": 100, "updating_it": 50, "eps_v": 1e-12},
               "proposal": {"name": "Gaussian", "covariance": {"type": "diag", "value": [1e-5, 1e-4, 1e-4, 1e-4]}}},
        "isde": {"name": "ISDE", "n_iterations": 400, "n_chains": 512, "seed": 3, "init_jitter": 0.01, "pooled_covariance": "yes",
                 "save_eval_freq": 10 ** 9, "gradient": "Analytical", "ISDE": {"h": 0.2, "f0": 4.0},
                 "covariance": {"value": "Matrix", "matrix_value": np.diag([1e-4, 5e-4, 5e-4, 1e-3]).tolist(),
                                "starting_it": 100, "update_it": 50, "end_it": 300}}}
    # closed-form posterior of the cubic (a linear-Gaussian problem)
    x = p["design"]["x"]
    Phi = np.stack([x ** k for k in range(4)], axis=1)
    A = Phi.T @ Phi / 0.1 ** 2
    cov_true = np.linalg.inv(A)
    mean_true = cov_true @ (Phi.T @ p["y"][0]) / 0.1 ** 2
    for key, algo in algos.items():
        with open(key + ".json", "w") as f:
            json.dump({"Sampling": {"BayesianPosterior": base, "Algorithm": algo}}, f)
        job = Sampling(key + ".json")
        run = job.sample({"m": Poly()})
        job.close()
        st = run.statistics
        assert st["n_chains"] == 512 and run.pooled
        assert pd.read_csv("output/mcmc_chain.csv", header=None).values.shape == (algo["n_iterations"] + 1, 4)
        if key == "am":
            # the factor every chain ends with is the pooled one of the last refresh (the last iteration)
            assert np.allclose(run.R.T @ run.R, 2.38 ** 2 / 4 * (st["cov"] + 1e-12 * np.eye(4)), rtol=1e-9)
            assert 0.1 < 1.0 - run.n_rejected / 1500 < 0.6                # tuned by the pooled covariance
        else:
            # after end_it the preconditioner is frozen: the pooled covariance of the states up to iteration 300
            assert np.allclose(run.C_approx, run.C_approx.T) and np.all(np.linalg.eigvalsh(run.C_approx) > 0)
            assert np.all(np.abs(np.diag(run.C_approx) / np.diag(cov_true) - 1.0) < 0.5)
        sd = np.sqrt(np.diag(cov_true))
        assert np.all(np.abs(st["mean"] - mean_true) < 0.2 * sd)
        assert np.all(st["rhat"] < 1.3)
