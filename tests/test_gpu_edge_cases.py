"""Edge cases of the engine through the C ABI: degenerate sizes, thinning phase across launches, out-of-prior
proposals, NaN policy, stream exhaustion, error paths."""
import numpy as np
import pytest

from oracle import mcmc_oracle as orc
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _mk(p, algo="RWMH", C=1, **kw):
    from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig
    return ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo=algo, **kw), C)


def test_zero_steps_single_point_single_chain():
    from pybitup_b200 import synthetic
    p = synthetic.cubic_problem(n_points=1)
    prob = H.oracle_problem(p)
    with _mk(p, "DRAM", 1, updating_it=3, eps_v=1e-8, gamma_dr=0.5) as eng:
        eng.init_chains(p["x0"][None, :], p["prop_cov"])
        x0, lp0 = eng.state()
        assert H.rel_err(lp0[0], orc.log_post(prob, p["x0"])) < 1e-12
        res = eng.run(0)                                     # nothing to do
        assert res.samples.shape[0] == 0 and eng.iteration == 0
        rng = np.random.default_rng(0)
        nz, un = rng.standard_normal((1, 80)), rng.random((1, 20))
        eng.set_streams(nz, un)
        res = eng.run(10, trace=True)
        eng.synchronize()
        out = orc.run_mh(prob, p["x0"], p["prop_cov"], 10, orc.Streams(nz[0], un[0]), adaptive=True, delayed=True,
                         updating_it=3, eps_v=1e-8, gamma_dr=0.5)
        assert np.array_equal(res.trace_acc.cpu().numpy()[:, 0], out["acc"])


def test_thinning_phase_continues_across_launches():
    from pybitup_b200 import synthetic
    p = synthetic.linear_problem()
    C, d = 5, 3
    rng = np.random.default_rng(2)
    nz, un = rng.standard_normal((C, 200 * d)), rng.random((C, 200))
    x0 = synthetic.chain_starts(p["x0"], C, jitter=1e-3)
    with _mk(p, "RWMH", C) as a, _mk(p, "RWMH", C) as b:
        for e in (a, b):
            e.init_chains(x0, p["prop_cov"])
            e.set_streams(nz, un)
        whole = a.run(100, thin=7).samples.cpu().numpy()                       # rows at iterations 7, 14, ..., 98
        parts = [b.run(n, thin=7).samples.cpu().numpy() for n in (10, 3, 1, 50, 36)]
        assert whole.shape[0] == 14 and sum(x.shape[0] for x in parts) == 14
        assert np.array_equal(np.concatenate(parts, axis=0), whole)
        assert np.array_equal(a.state()[0], b.state()[0])


def test_proposals_outside_the_prior_box_are_rejected_without_model_evaluation():
    from pybitup_b200 import synthetic
    p = synthetic.linear_problem()
    p["lb"], p["ub"] = p["x0"] - 1e-4, p["x0"] + 1e-4          # a box much tighter than the proposal
    prob = H.oracle_problem(p)
    C, T = 8, 50
    rng = np.random.default_rng(1)
    nz, un = rng.standard_normal((C, T * 3)), rng.random((C, T))
    with _mk(p, "RWMH", C, lanes_per_chain=1, chains_per_thread=2) as eng:
        eng.init_chains(np.tile(p["x0"], (C, 1)), p["prop_cov"])
        eng.set_streams(nz, un)
        res = eng.run(T, trace=True)
        eng.synchronize()
        lp1 = res.trace_lp1.cpu().numpy()
        assert np.isinf(lp1).mean() > 0.9                    # almost every proposal leaves the box: log-posterior -inf
        for c in range(C):
            out = orc.run_mh(prob, p["x0"], p["prop_cov"], T, orc.Streams(nz[c], un[c]))
            assert np.array_equal(res.trace_acc.cpu().numpy()[:, c], out["acc"])
            assert H.rel_err(res.samples.cpu().numpy()[:, :, c], out["chain"][1:]) < 1e-12


def test_nan_ratio_policy_and_status_flags():
    """min(1, nan) == 1 accepts in the reference; nan_rejects=1 rejects instead; both set the status bit."""
    from pybitup_b200 import synthetic
    p = synthetic.cubic_problem(n_points=16)
    p["y"] = p["y"].copy()
    p["y"][0, 3] = np.nan                                     # a NaN datum makes every log-posterior NaN
    for nan_rejects, expect_moves in ((False, True), (True, False)):
        with _mk(p, "RWMH", 4, nan_rejects=nan_rejects, seed=3) as eng:
            eng.init_chains(np.tile(p["x0"], (4, 1)), p["prop_cov"])
            eng.run(20, keep_samples=False)
            ctr = eng.counters()
            assert np.all(ctr["status"] & 4)
            assert (np.all(ctr["n_rejected"] == 0)) == expect_moves
            assert (np.all(ctr["n_rejected"] == 20)) == (not expect_moves)


def test_injected_stream_exhaustion_is_flagged():
    from pybitup_b200 import synthetic
    p = synthetic.linear_problem()
    with _mk(p, "RWMH", 2) as eng:
        eng.init_chains(np.tile(p["x0"], (2, 1)), p["prop_cov"])
        eng.set_streams(np.zeros((2, 3 * 5)), np.full((2, 5), 0.5))      # enough for 5 iterations only
        eng.run(8, keep_samples=False)
        assert np.all(eng.counters()["status"] & 2)
        eng.set_streams(None, None)                                       # back to Philox
        eng.run(8, keep_samples=False)


def test_error_paths():
    from pybitup_b200 import synthetic, _capi
    from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig
    p = synthetic.linear_problem()
    with _mk(p, "RWMH", 2) as eng:
        with pytest.raises(_capi.EngineError):
            eng.run(5)                                        # init_chains not called
        with pytest.raises(np.linalg.LinAlgError):
            eng.init_chains(np.tile(p["x0"], (2, 1)), -np.eye(3))          # scipy.linalg.cholesky raises LinAlgError
        with pytest.raises(ValueError):
            eng.init_chains(np.zeros((3, 3)), None)
    with pytest.raises(ValueError):
        SamplerConfig(algo="NUTS")._c_struct()
    q = dict(p)
    q["theta_nom"] = np.arange(11.0)                          # no linear kernel with 11 parameters is compiled
    q["unc_idx"] = np.arange(11)
    q["design"] = dict(Phi=np.ones((50, 11)))
    q["lb"], q["ub"] = np.full(11, -1.0), np.full(11, 1.0)
    with pytest.raises(_capi.UnsupportedProblem):
        ChainEngine(ProblemSpec.from_dict(q), SamplerConfig(), 1)
    with pytest.raises(_capi.UnsupportedProblem):
        ChainEngine(ProblemSpec.from_dict(synthetic.heat_problem()), SamplerConfig(algo="ISDE", isde_h=0.1, isde_f0=1.0), 1)


def test_philox_chains_do_not_depend_on_sharding_or_launch_shape():
    """Philox is keyed by (seed, global chain index, iteration): the chains are the same whatever the lane count,
    chains-per-thread, block size, launch splitting or rank sharding."""
    from pybitup_b200 import synthetic
    p = synthetic.cubic_problem(n_points=100)
    C = 96
    x0 = synthetic.chain_starts(p["x0"], C, jitter=1e-3)
    ref = None
    for lanes, q, block, split, shard in ((1, 1, 0, (30,), None), (2, 2, 64, (7, 23), None), (8, 1, 128, (30,), None),
                                          (1, 2, 32, (30,), (40, 56))):
        kw = dict(lanes_per_chain=lanes, chains_per_thread=q, block_threads=block, seed=77, updating_it=5, eps_v=1e-10)
        if shard is None:
            with _mk(p, "AMH", C, **kw) as eng:
                eng.init_chains(x0, p["prop_cov"])
                for n in split:
                    eng.run(n, keep_samples=False)
                x = eng.state()[0]
        else:
            xs, first = [], 0
            for cnt in shard:
                with _mk(p, "AMH", cnt, chain_offset=first, **kw) as eng:
                    eng.init_chains(x0[first:first + cnt], p["prop_cov"])
                    eng.run(30, keep_samples=False)
                    xs.append(eng.state()[0])
                first += cnt
            x = np.concatenate(xs)
        if ref is None:
            ref = x
        else:
            assert H.rel_err(x, ref) < 1e-9, (lanes, q, block)
