"""GPU parity tests (run on the B200 box: ``pytest -m gpu``).

Every test drives the CUDA engine through the C ABI (``pybitup_b200._capi`` -> libpybitup_b200.so)
and checks it against (a) golden vectors recorded from the UNMODIFIED reference and (b) the CPU
oracle on the same seeded inputs.  Bar (BASELINE.json north_star): with identical proposal / uniform
streams, log-posteriors agree to 1e-10 relative and accept/reject sequences are identical.
"""
import numpy as np
import pytest

from oracle import mcmc_oracle as orc
from tests import helpers as H

pytestmark = pytest.mark.gpu

MH_CASES = [c for c in H.golden_cases() if "isde" not in c and "hmc" not in c]
TOL = 1e-10


def _engine(g_or_dict, algo, n_chains, **kw):
    from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig
    p = H.problem_dict(g_or_dict)
    name = algo["name"]
    cfg = SamplerConfig(algo=name, **kw)
    if name in ("AMH", "DRAM"):
        cfg.updating_it = int(algo[name]["updating_it"])
        cfg.eps_v = float(algo[name]["eps_v"])
    if name in ("DR", "DRAM"):
        cfg.gamma_dr = float(algo[name]["gamma"])
    return ChainEngine(ProblemSpec.from_dict(p), cfg, n_chains)


def _eval_sequence(res, lp0, n_init, delayed, chain=0):
    lp1 = res.trace_lp1[:, chain].cpu().numpy()
    lp2 = res.trace_lp2[:, chain].cpu().numpy()
    acc = res.trace_acc[:, chain].cpu().numpy()
    seq = [lp0] * n_init
    for s in range(len(lp1)):
        seq.append(lp1[s])
        if delayed and acc[s] != 1:
            seq.append(lp2[s])
    return np.array(seq), acc


@pytest.mark.parametrize("lanes", [1, 2, 8, -4])
@pytest.mark.parametrize("case", MH_CASES)
def test_mh_golden_parity(case, lanes):
    """CUDA engine vs the recorded reference run, fed the same streams.  lanes < 0: cooperative teams of -lanes lanes."""
    g = H.load_golden(case)
    algo = g["algo"]
    if g["model"] == "arrhenius" and not (lanes == 1 or (lanes == 2 and algo["name"] in ("DR", "DRAM"))):
        pytest.skip("sequential ODE model: one lane per chain (or the speculative pair of delayed rejection)")
    if "n_models" in g and lanes < 0:
        pytest.skip("several models per likelihood run in the replicated layout (the cooperative kernel is single-model)")
    if g["model"] in ("linear", "poly") and len(g["unc_idx"]) < len(g["theta_nom"]) and lanes != 1:
        pytest.skip("models with fixed parameters (d < P) are compiled for one lane per chain")
    T = int(algo["n_iterations"])
    tol = TOL          # every fixture: the kernel's Cholesky runs in LAPACK's dpotf2 order (pbu_common.cuh)
    n_chains = 3 if lanes > 0 else 11       # the same chain several times: every chain must reproduce the reference
    eng = _engine(g, algo, n_chains, lanes_per_chain=abs(lanes), cooperative=(1 if lanes < 0 else -1), track_map=H.wants_map(g))
    eng.init_chains(np.tile(H.start_point(g), (n_chains, 1)), g["prop_cov"])      # (sampling space: forward(initial_val))
    x_init, lp_init = eng.state()
    eng.set_streams(np.tile(g["normals"], (n_chains, 1)), np.tile(g["uniforms"], (n_chains, 1)))
    res = eng.run(T, thin=1, keep_lp=True, trace=True)
    eng.synchronize()
    delayed = algo["name"] in ("DR", "DRAM")
    n_init = 2 if algo["name"] == "DRAM" else 1
    ref_chain = H.golden_chain(g)
    moved_ref = np.any(ref_chain[1:] != ref_chain[:-1], axis=1)
    samples = res.samples.cpu().numpy()           # [T, d, C]
    ctr = eng.counters()
    for c in range(n_chains):
        seq, acc = _eval_sequence(res, lp_init[c], n_init, delayed, c)
        assert len(seq) == len(g["eval_lp"]), "number of posterior evaluations differs"
        assert H.rel_err(seq, g["eval_lp"]) < tol
        assert np.array_equal(acc > 0, moved_ref), "accept/reject sequence differs from the reference"
        chain = np.concatenate([x_init[c][None, :], samples[:, :, c]], axis=0)
        assert H.rel_err(chain, ref_chain) < tol
        assert ctr["n_rejected"][c] == int(g["n_rejected"])
        assert ctr["n_eval"][c] == len(g["eval_lp"])
        assert ctr["status"][c] & 3 == 0
        assert H.rel_err(ctr["mean_r"][c], float(g["mean_r"])) < max(tol, 1e-9)
    if H.wants_map(g):          # estimate_max on the device (mha.py:226-237): every copy of the chain finds the reference's MAP
        arg_map, map_val = eng.map_estimate()
        for c in range(n_chains):
            assert H.rel_err(arg_map[c], g["final_arg_MAP"]) < tol
            assert H.rel_err(map_val[c], float(g["final_MAP_val"])) < tol
    prop = eng.proposal()
    assert H.rel_err(prop["R"][0], g["final_R"]) < max(1e-7, 100 * tol)
    if algo["name"] in ("AMH", "DRAM"):
        assert H.rel_err(prop["V"][0], g["final_V_i"]) < max(1e-6, 100 * tol)
        assert H.rel_err(prop["X_av"][0], g["final_X_av_i"]) < max(tol, 1e-9)
    # all injected draws consumed in the reference's order: d normals + 1 uniform per stage
    x_fin, lp_fin = eng.state()
    assert H.rel_err(x_fin[0], ref_chain[-1]) < tol
    eng.close()


@pytest.mark.parametrize("algo_name", ["RWMH", "AMH", "DR", "DRAM"])
@pytest.mark.parametrize("builder,lanes", [("linear_problem", 1), ("linear_problem", 8), ("cubic_problem", 2),
                                           ("heat_problem", 1), ("arrhenius_problem", 1), ("cubic_problem", -4),
                                           ("linear_problem", -4), ("heat_problem", -4), ("arrhenius_problem", 2)])
def test_many_chains_vs_oracle(builder, lanes, algo_name):
    """64 different chains (own start, own streams) against the CPU oracle, chain by chain.  (arrhenius, 2): speculative
    delayed rejection -- a pair of lanes evaluates both proposal stages of a chain at once."""
    if builder == "arrhenius_problem" and lanes == 2 and algo_name not in ("DR", "DRAM"):
        pytest.skip("the two-lane layout of a sequential model is the speculative delayed rejection")
    _many_chains(builder, lanes, algo_name)


@pytest.mark.parametrize("algo_name", ["RWMH", "AMH", "DR", "DRAM"])
def test_replicated_tables_step_kernel_vs_oracle(algo_name, monkeypatch):
    """The Arrhenius step kernels with the conflict-free replicated exp2 / log2 tables (192 KB of shared memory, what the
    planner picks when the chains fill the GPU), forced here for 64 chains: same chains as the oracle's."""
    lanes = 2 if algo_name in ("DR", "DRAM") else 1
    monkeypatch.setenv("PBU_FORCE_VARIANT", "%d,1,0,0,1" % lanes)
    geo = _many_chains("arrhenius_problem", lanes, algo_name)
    assert geo["smem_bytes"] > 190000 and geo["lanes"] == lanes


@pytest.mark.parametrize("algo_name", ["RWMH", "AMH", "DRAM"])
def test_uniform_weight_kernel(algo_name, monkeypatch):
    """Polynomial model, one std_y for all points: the engine picks the uniform-weight cooperative kernel (PolyModelU: the
    weight folded into the coefficients, records [x | w*y]).  Same accept codes and states (rounding apart) as the general
    kernel and as the oracle; unequal std_y fall back to the general kernel."""
    from pybitup_b200 import synthetic
    p = synthetic.cubic_problem(n_points=333)
    prob = H.oracle_problem(p)
    n_chains, T, d = 64, 50, prob.d
    rng = np.random.default_rng(11)
    x0 = synthetic.chain_starts(p["x0"], n_chains, jitter=1e-3, seed=2)
    normals, uniforms = rng.standard_normal((n_chains, 2 * T * d)), rng.random((n_chains, 2 * T))
    algo = {"name": algo_name, "AMH": {"updating_it": 7, "eps_v": 1e-10}, "DRAM": {"updating_it": 7, "eps_v": 1e-10, "gamma": 0.3}}
    runs = {}
    for uni in ("1", "0"):
        monkeypatch.setenv("PBU_UNIFORM_W", uni)
        eng = _engine(p, algo, n_chains, lanes_per_chain=4, cooperative=1)
        assert eng.launch_geometry()["uniform_weight"] == int(uni) and eng.launch_geometry()["cooperative"] == 1
        eng.init_chains(x0, p["prop_cov"])
        eng.set_streams(normals, uniforms)
        res = eng.run(T, thin=1, trace=True)
        eng.synchronize()
        runs[uni] = (res.samples.cpu().numpy(), res.trace_acc.cpu().numpy(), eng.state()[1])
        eng.close()
    assert np.array_equal(runs["1"][1], runs["0"][1])
    assert H.rel_err(runs["1"][0], runs["0"][0]) < 1e-11 and H.rel_err(runs["1"][2], runs["0"][2]) < 1e-11
    import tests.helpers as helpers
    for c in range(0, n_chains, 7):
        with orc.kernel_order():
            out = orc.run_mh(prob, x0[c], p["prop_cov"], T, orc.Streams(normals[c], uniforms[c]), **helpers.mh_kwargs(algo))
        assert np.array_equal(runs["1"][1][:, c], out["acc"])
        assert H.rel_err(runs["1"][0][:, :, c], out["chain"][1:]) < 1e-9
    monkeypatch.delenv("PBU_UNIFORM_W")
    p2 = dict(p)
    p2["std_y"] = p["std_y"] * (1.0 + 0.5 * rng.random(p["std_y"].shape))
    eng = _engine(p2, algo, n_chains, lanes_per_chain=4, cooperative=1)
    assert eng.launch_geometry()["uniform_weight"] == 0
    eng.close()


def _many_chains(builder, lanes, algo_name):
    from pybitup_b200 import synthetic
    kwargs = {"linear_problem": {}, "cubic_problem": {"n_points": 257}, "heat_problem": {},
              "arrhenius_problem": {"n_steps": 40, "dT": 15.0}}[builder]
    p = getattr(synthetic, builder)(**kwargs)
    prob = H.oracle_problem(p)
    n_chains, T = 64, 60
    d = prob.d
    rng = np.random.default_rng(123)
    x0 = synthetic.chain_starts(p["x0"], n_chains, jitter=1e-3, seed=5)
    x0 = np.minimum(np.maximum(x0, p["lb"]), p["ub"])
    normals = rng.standard_normal((n_chains, 2 * T * d))
    uniforms = rng.random((n_chains, 2 * T))
    algo = {"name": algo_name, "AMH": {"updating_it": 7, "eps_v": 1e-10}, "DR": {"gamma": 0.3},
            "DRAM": {"updating_it": 7, "eps_v": 1e-10, "gamma": 0.3}}
    eng = _engine(p, algo, n_chains, lanes_per_chain=abs(lanes), cooperative=(1 if lanes < 0 else -1))
    geo = eng.launch_geometry()
    eng.init_chains(x0, p["prop_cov"])
    _, lp_init = eng.state()
    eng.set_streams(normals, uniforms)
    res = eng.run(T, thin=1, trace=True)
    eng.synchronize()
    x_fin, lp_fin = eng.state()
    samples = res.samples.cpu().numpy()
    acc_dev = res.trace_acc.cpu().numpy()
    tol = 1e-9
    import tests.helpers as helpers
    for c in range(0, n_chains, 3):
        st = orc.Streams(normals[c], uniforms[c])
        with orc.kernel_order():      # the written-out operation orders of the kernels (bit-identical to scipy / numpy up to d = 4)
            out = orc.run_mh(prob, x0[c], p["prop_cov"], T, st, **helpers.mh_kwargs(algo))
        assert np.array_equal(acc_dev[:, c], out["acc"]), "chain %d: accept codes differ" % c
        assert H.rel_err(samples[:, :, c], out["chain"][1:]) < tol
        assert H.rel_err(lp_fin[c], out["lp"][-1]) < tol
        assert H.rel_err(lp_init[c], out["lp"][0]) < 1e-12
    eng.close()
    return geo


@pytest.mark.parametrize("algo_name", ["AMH", "DRAM"])
@pytest.mark.parametrize("builder,lanes", [("linear_problem", 1), ("cubic_problem", 2), ("cubic_problem", -4), ("heat_problem", 8),
                                           ("arrhenius_problem", 1)])
def test_welford_adaptation_vs_oracle(builder, lanes, algo_name):
    """am_welford=1 (the production update north_star names): 48 chains with injected streams against the oracle's
    Welford restatement -- identical accept codes, states and final (V_i, X_av, R) -- in every kernel layout."""
    from pybitup_b200 import synthetic
    kwargs = {"linear_problem": {}, "cubic_problem": {"n_points": 257}, "heat_problem": {},
              "arrhenius_problem": {"n_steps": 40, "dT": 15.0}}[builder]
    p = getattr(synthetic, builder)(**kwargs)
    prob = H.oracle_problem(p)
    n_chains, T, d = 48, 60, prob.d
    rng = np.random.default_rng(77)
    x0 = np.minimum(np.maximum(synthetic.chain_starts(p["x0"], n_chains, jitter=1e-3, seed=9), p["lb"]), p["ub"])
    normals = rng.standard_normal((n_chains, 2 * T * d))
    uniforms = rng.random((n_chains, 2 * T))
    algo = {"name": algo_name, "AMH": {"updating_it": 7, "eps_v": 1e-10}, "DRAM": {"updating_it": 7, "eps_v": 1e-10, "gamma": 0.3}}
    eng = _engine(p, algo, n_chains, lanes_per_chain=abs(lanes), cooperative=(1 if lanes < 0 else -1), am_welford=True)
    eng.init_chains(x0, p["prop_cov"])
    eng.set_streams(normals, uniforms)
    res = eng.run(T, thin=1, trace=True)
    eng.synchronize()
    samples = res.samples.cpu().numpy()
    acc_dev = res.trace_acc.cpu().numpy()
    prop = eng.proposal()
    import tests.helpers as helpers
    for c in range(0, n_chains, 5):
        with orc.kernel_order():
            out = orc.run_mh(prob, x0[c], p["prop_cov"], T, orc.Streams(normals[c], uniforms[c]), am_welford=True,
                             **helpers.mh_kwargs(algo))
        assert np.array_equal(acc_dev[:, c], out["acc"]), "chain %d: accept codes differ" % c
        assert H.rel_err(samples[:, :, c], out["chain"][1:]) < 1e-9
        assert H.rel_err(prop["X_av"][c], out["X_av"]) < 1e-10
        assert H.rel_err(prop["V"][c], out["V_i"]) < 1e-8
        assert H.rel_err(prop["R"][c], out["R"]) < 1e-7
    assert not np.any(eng.counters()["status"] & 3)          # bit 2 (a NaN delayed-rejection ratio, Q6) is the reference's behaviour
    eng.close()


@pytest.mark.parametrize("builder", ["linear_problem", "cubic_problem", "heat_problem", "arrhenius_problem",
                                     "regression_problem"])
def test_logposterior_and_model_eval(builder):
    from pybitup_b200 import synthetic
    kwargs = {"regression_problem": {"n_points": 300, "d": 10}, "arrhenius_problem": {"n_steps": 80, "dT": 8.0},
              "cubic_problem": {"n_points": 1000}}.get(builder, {})
    p = getattr(synthetic, builder)(**kwargs)
    prob = H.oracle_problem(p)
    rng = np.random.default_rng(7)
    S = 200
    X = p["x0"][None, :] * (1.0 + 0.02 * rng.standard_normal((S, prob.d)))
    X[::17] = p["ub"] + 1.0            # outside the prior box -> -inf, model not evaluated
    X[5] = p["lb"]                     # bounds are inclusive (distributions.py:284)
    eng = _engine(p, {"name": "RWMH"}, 4)
    lp = eng.log_posterior(X)
    f = eng.model_eval(X[1:9])
    with np.errstate(all="ignore"):
        lp_ref = np.array([orc.log_post(prob, x) for x in X])
        f_ref = np.array([prob.model_eval(x) for x in X[1:9]])
    assert np.array_equal(np.isinf(lp), np.isinf(lp_ref))
    assert H.rel_err(lp, lp_ref) < 1e-11
    assert H.rel_err(f, f_ref) < 1e-11
    eng.close()


def test_multi_run_data():
    """n_runs > 1: the sum of squares runs over every experimental run (bayesian_inference.py:507-524)."""
    from pybitup_b200 import synthetic
    p = synthetic.cubic_problem(n_points=129)
    rng = np.random.default_rng(3)
    y0 = p["y"][0]
    p["y"] = np.stack([y0, y0 + 0.05 * rng.standard_normal(y0.size), y0 - 0.02, y0 + 0.01])
    prob = H.oracle_problem(p)
    X = p["x0"][None, :] * (1.0 + 0.01 * rng.standard_normal((50, 4)))
    for lanes in (1, 8):
        eng = _engine(p, {"name": "RWMH"}, 2, lanes_per_chain=lanes)
        lp = eng.log_posterior(X)
        assert H.rel_err(lp, [orc.log_post(prob, x) for x in X]) < 1e-11
        eng.close()


def test_heat_known_answer_on_device():
    """log-likelihood of the reference's heat-conduction case at the nominal point."""
    from pybitup_b200 import synthetic
    p = synthetic.heat_problem()
    eng = _engine(p, {"name": "RWMH"}, 1)
    lp = eng.log_posterior(np.array([[-18.41, 0.00191]]))[0]
    log_prior = np.log(1 / 100.0) + np.log(1 / (1.0 - 0.00001))
    assert abs((lp - log_prior) - (-5.9943119569)) < 1e-8
    eng.close()
