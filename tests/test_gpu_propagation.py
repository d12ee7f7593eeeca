"""K3: Monte-Carlo propagation on the device against the oracle restatement of solve_problem.py:543-647."""
import numpy as np
import pytest

from oracle import mcmc_oracle as orc
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("builder,kwargs,S", [("arrhenius_problem", {"n_steps": 60, "dT": 10.0}, 1500),
                                              ("heat_problem", {}, 4001), ("cubic_problem", {"n_points": 33}, 777),
                                              ("linear_problem", {}, 40)])
def test_propagation_statistics_match_oracle(builder, kwargs, S):
    from pybitup_b200 import synthetic
    from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig
    from pybitup_b200.propagation import propagate
    p = getattr(synthetic, builder)(**kwargs)
    prob = H.oracle_problem(p)
    rng = np.random.default_rng(8)
    X = p["x0"][None, :] * (1.0 + 0.01 * rng.standard_normal((S, prob.d)))
    X[3] = X[2]                      # ties
    ref = orc.propagate(prob, X)
    with ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo="RWMH"), 1) as eng:
        out = propagate(eng, X, return_evals=True)
    assert H.rel_err(out["fun_eval"], ref["fun_eval"]) < 1e-11
    assert H.rel_err(out["mean"], ref["mean"]) < 1e-12
    # min / max / percentiles are exact order statistics of the device evaluations
    fe = out["fun_eval"]
    assert np.array_equal(out["min"], fe.min(axis=0)) and np.array_equal(out["max"], fe.max(axis=0))
    assert np.array_equal(out["ci_lb"], np.percentile(fe, 2.5, axis=0))
    assert np.array_equal(out["ci_ub"], np.percentile(fe, 97.5, axis=0))
    assert H.rel_err(out["ci_lb"], ref["ci_lb"]) < 1e-10 and H.rel_err(out["ci_ub"], ref["ci_ub"]) < 1e-10


def test_propagation_gaussian_samples_and_sharding():
    """Samples drawn on the device: sharding them over two engines (as two GPUs would) and summing the moment /
    histogram arrays gives the statistics of the unsharded run, bit for bit on the order statistics."""
    from pybitup_b200 import synthetic
    from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig
    from pybitup_b200.propagation import DevicePropagation
    p = synthetic.arrhenius_problem(n_steps=40, dT=15.0)
    mean, std = p["x0"], 0.01 * np.abs(p["x0"])
    S = 6000
    spec = ProblemSpec.from_dict(p)
    with ChainEngine(spec, SamplerConfig(algo="RWMH"), 1) as eng:
        with DevicePropagation(eng, S) as dp:
            dp.evaluate_gaussian(mean, std, seed=7, sample_offset=0)
            whole = dp.statistics()
            fe = dp.evals()
        halves = [DevicePropagation(eng, S // 2), DevicePropagation(eng, S // 2)]
        for i, h in enumerate(halves):
            h.evaluate_gaussian(mean, std, seed=7, sample_offset=i * (S // 2))
        # emulate the all-reduce over two ranks by running the protocol of both shards in lock step
        import ctypes as C
        from pybitup_b200 import _capi as capi
        from pybitup_b200.propagation import percentile_ranks, lerp_percentile
        N = p["y"].shape[1]
        sums, mns, mxs = [], [], []
        for h in halves:
            s, mn, mx = np.empty(N), np.empty(N), np.empty(N)
            capi.check(capi.lib.pbu_propagation_moments(h._h, capi.dptr(s), capi.dptr(mn), capi.dptr(mx)))
            sums.append(s), mns.append(mn), mxs.append(mx)
        assert H.rel_err((sums[0] + sums[1]) / S, whole["mean"]) < 1e-13
        assert np.array_equal(np.minimum(*mns), whole["min"]) and np.array_equal(np.maximum(*mxs), whole["max"])
        ranks = np.array(list(percentile_ranks(S, 2.5)) + list(percentile_ranks(S, 97.5)), dtype=np.int64)
        for h in halves:
            capi.check(capi.lib.pbu_propagation_select_begin(h._h, ranks.ctypes.data_as(C.POINTER(C.c_int64)), 4, None, None, None))
        for ps in range(8):
            hs = []
            for h in halves:
                hist = np.empty((N, 4, 256), dtype=np.uint64)
                capi.check(capi.lib.pbu_propagation_select_pass(h._h, ps, hist.ctypes.data_as(C.POINTER(C.c_uint64))))
                hs.append(hist)
            tot = np.ascontiguousarray(hs[0] + hs[1])
            for h in halves:
                capi.check(capi.lib.pbu_propagation_select_update(h._h, ps, tot.ctypes.data_as(C.POINTER(C.c_uint64))))
        vals = np.empty((N, 4))
        capi.check(capi.lib.pbu_propagation_select_result(halves[0]._h, capi.dptr(vals)))
        lb = np.array([lerp_percentile(S, 2.5, vals[k, 0], vals[k, 1]) for k in range(N)])
        ub = np.array([lerp_percentile(S, 97.5, vals[k, 2], vals[k, 3]) for k in range(N)])
        assert np.array_equal(lb, whole["ci_lb"]) and np.array_equal(ub, whole["ci_ub"])
        assert np.array_equal(whole["ci_lb"], np.percentile(fe, 2.5, axis=0))
        for h in halves:
            h.close()
