"""Streaming mode: point records cut into shared-memory tiles (TMA ring) instead of being resident.  Forced here
with small tiles on small problems; results must agree with the resident run and with the oracle."""
import numpy as np
import pytest

from oracle import mcmc_oracle as orc
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("algo", ["RWMH", "DRAM"])
@pytest.mark.parametrize("builder,kwargs,lanes", [("cubic_problem", {"n_points": 301}, 1), ("cubic_problem", {"n_points": 301}, 8),
                                                  ("regression_problem", {"n_points": 500, "d": 10}, 1),
                                                  ("arrhenius_problem", {"n_steps": 60, "dT": 10.0}, 1)])
def test_mh_streamed_equals_resident(builder, kwargs, lanes, algo):
    from pybitup_b200 import synthetic
    from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig
    p = getattr(synthetic, builder)(**kwargs)
    C, T, d = 70, 40, len(p["x0"])
    rng = np.random.default_rng(5)
    x0 = synthetic.chain_starts(p["x0"], C, jitter=1e-3, seed=2)
    x0 = np.minimum(np.maximum(x0, p["lb"]), p["ub"])
    normals, uniforms = rng.standard_normal((C, 2 * T * d)), rng.random((C, 2 * T))
    outs = []
    for tile in (0, 37):
        cfg = SamplerConfig(algo=algo, updating_it=5, eps_v=1e-10, gamma_dr=0.3, lanes_per_chain=lanes, tile_points=tile)
        eng = ChainEngine(ProblemSpec.from_dict(p), cfg, C)
        eng.init_chains(x0, p["prop_cov"])
        eng.set_streams(normals, uniforms)
        res = eng.run(T, thin=1, trace=True)
        eng.synchronize()
        outs.append((res.samples.cpu().numpy(), res.trace_acc.cpu().numpy(), res.trace_lp1.cpu().numpy()))
        eng.close()
    assert np.array_equal(outs[0][1], outs[1][1])
    tol = 1e-10
    assert H.rel_err(outs[1][0], outs[0][0]) < tol
    assert H.rel_err(outs[1][2], outs[0][2]) < tol


def test_streamed_logposterior_and_model_vs_oracle():
    from pybitup_b200 import synthetic
    from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig
    p = synthetic.regression_problem(n_points=700, d=10)
    prob = H.oracle_problem(p)
    rng = np.random.default_rng(1)
    X = p["x0"][None, :] * (1.0 + 0.05 * rng.standard_normal((130, 10)))
    eng = ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo="RWMH", tile_points=64), 4)
    lp = eng.log_posterior(X)
    f = eng.model_eval(X[:5])
    assert H.rel_err(lp, [orc.log_post(prob, x) for x in X]) < 1e-11
    assert H.rel_err(f, [prob.model_eval(x) for x in X[:5]]) < 1e-11
    eng.close()


def test_isde_streamed_vs_oracle():
    from pybitup_b200 import synthetic
    from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig
    p = synthetic.regression_problem(n_points=600, d=10)
    prob = H.oracle_problem(p)
    C, T, d = 33, 30, 10
    rng = np.random.default_rng(4)
    x0 = synthetic.chain_starts(p["x0"], C, jitter=1e-2, seed=9)
    normals = rng.standard_normal((C, T * d))
    cfg = SamplerConfig(algo="ISDE", isde_h=0.01, isde_f0=4.0, isde_gradient="Analytical", tile_points=100)
    eng = ChainEngine(ProblemSpec.from_dict(p), cfg, C)
    eng.init_chains(x0, None)
    eng.set_streams(normals, None)
    res = eng.run(T, thin=1)
    eng.synchronize()
    xs = res.samples.cpu().numpy()
    for c in (0, 7, 32):
        out = orc.run_isde(prob, x0[c], T, orc.Streams(normals[c], np.zeros(1)), 0.01, 4.0, gradient="Analytical")
        assert H.rel_err(xs[:, :, c], out["chain"][1:]) < 1e-8
    eng.close()
