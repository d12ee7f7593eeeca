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


@pytest.mark.parametrize("d", [4, 2])
def test_replicated_table_kernel_is_bit_identical(d, monkeypatch):
    """Arrhenius evaluations with the conflict-free replicated lookup tables (1024-thread persistent blocks, what cfg 5
    runs) against the plain-table kernel: the same table values, the same arithmetic -> identical bits; and against
    the oracle.  S is not a multiple of the block, several grid strides per block."""
    from pybitup_b200 import synthetic
    from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig
    from pybitup_b200.propagation import DevicePropagation
    p = synthetic.arrhenius_problem(n_steps=50, dT=12.0)
    if d < 4:                        # E and F fixed at their nominal values: the d = 2 of 4 instantiation
        idx = np.array([0, 2])
        p.update(unc_idx=idx, lb=p["lb"][idx], ub=p["ub"][idx], x0=p["x0"][idx])
    S = 2 * 148 * 1024 + 1234
    mean, std = p["x0"], 0.03 * np.abs(p["x0"])
    out = {}
    with ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo="RWMH"), 1) as eng:
        for rep in ("0", "1"):
            monkeypatch.setenv("PBU_REP_TABLES", rep)
            with DevicePropagation(eng, S) as dp:
                dp.evaluate_gaussian(mean, std, seed=3, sample_offset=0)
                out[rep] = dp.evals()
                if rep == "1":
                    X = dp.samples()[:300]
    assert np.array_equal(out["0"], out["1"])
    ref = orc.propagate(H.oracle_problem(p), X)
    assert H.rel_err(out["1"][:300], ref["fun_eval"]) < 1e-11


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


def _bracket_problem():
    from pybitup_b200 import synthetic
    return synthetic.arrhenius_problem(n_steps=24, dT=25.0)


@pytest.mark.parametrize("percentiles", [(2.5, 97.5), (50.0,), (0.001, 99.999)])
def test_bracketed_select_is_exact(percentiles):
    """The one-pass bracketed select (subsample brackets -> one pass -> select among the candidates) returns the same
    exact order statistics as np.percentile / the full radix select, ties and open brackets included."""
    from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig
    from pybitup_b200.propagation import DevicePropagation
    p = _bracket_problem()
    S = 200_000 + 37
    with ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo="RWMH"), 1) as eng:
        with DevicePropagation(eng, S) as dp:
            dp.evaluate_gaussian(p["x0"], 0.01 * np.abs(p["x0"]), seed=3)
            fast = dp.statistics(percentiles, bracket=True)
            assert dp.last_select == "bracket"
            slow = dp.statistics(percentiles, bracket=False)
            assert dp.last_select == "radix"
            fe = dp.evals()
    for q in percentiles:
        assert np.array_equal(fast["p%g" % q], np.percentile(fe, q, axis=0))
        assert np.array_equal(fast["p%g" % q], slow["p%g" % q])
    assert np.array_equal(fast["min"], fe.min(axis=0)) and np.array_equal(fast["max"], fe.max(axis=0))
    assert H.rel_err(fast["mean"], fe.mean(axis=0)) < 1e-12
    assert np.array_equal(fast["mean"], slow["mean"])


def test_bracketed_select_falls_back_when_the_subsample_misleads():
    """Samples laid out so that exactly the chunks the subsample reads come from another distribution: the brackets miss
    the wanted ranks, the counts of the pass show it, and the full radix select gives the exact answer."""
    from pybitup_b200 import synthetic
    from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig
    from pybitup_b200.propagation import DevicePropagation
    p = synthetic.cubic_problem(n_points=9)
    S = 64 * 256 * 12
    rng = np.random.default_rng(5)
    X = p["x0"][None, :] * (1.0 + 0.01 * rng.standard_normal((S, len(p["x0"]))))
    chunk = (np.arange(S) // 256) % DevicePropagation.BRACKET_STRIDE == 0
    X[chunk, 0] += 5.0
    with ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo="RWMH"), 1) as eng:
        with DevicePropagation(eng, S) as dp:
            dp.evaluate(X)
            st = dp.statistics(bracket=True)
            assert dp.last_select == "radix (bracket missed)"
            fe = dp.evals()
    assert np.array_equal(st["ci_lb"], np.percentile(fe, 2.5, axis=0))
    assert np.array_equal(st["ci_ub"], np.percentile(fe, 97.5, axis=0))
    assert np.array_equal(st["min"], fe.min(axis=0)) and np.array_equal(st["max"], fe.max(axis=0))


def test_bracketed_select_sharded_equals_unsharded():
    """Two shards run the bracket protocol in lock step through a reducer that combines them as an all-reduce would:
    bit-identical order statistics to the single-engine run."""
    from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig
    from pybitup_b200.propagation import DevicePropagation
    import threading
    p = _bracket_problem()
    S = 300_000
    mean, std = p["x0"], 0.01 * np.abs(p["x0"])
    with ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo="RWMH"), 1) as eng:
        with DevicePropagation(eng, S) as dp:
            dp.evaluate_gaussian(mean, std, seed=11)
            whole = dp.statistics(bracket=True)
        halves = [DevicePropagation(eng, S // 2), DevicePropagation(eng, S // 2)]
        for i, h in enumerate(halves):
            h.evaluate_gaussian(mean, std, seed=11, sample_offset=i * (S // 2))
        # a two-party all-reduce between two threads (each drives one shard; the C ABI serialises on the engine's stream)
        barrier = threading.Barrier(2)
        slots = [None, None]
        lock = threading.Lock()

        def make_reduce(me):
            def red(a, op):
                a = np.ascontiguousarray(a)
                slots[me] = a
                barrier.wait()
                other = slots[1 - me]
                res = {"sum": a + other, "min": np.minimum(a, other), "max": np.maximum(a, other)}[op]
                barrier.wait()
                return res
            return red

        outs, errs = [None, None], []

        def run(me):
            try:
                outs[me] = halves[me].statistics(reduce=make_reduce(me), n_total=S, bracket=True)
                assert halves[me].last_select == "bracket"
            except Exception as ex:          # pragma: no cover
                errs.append(ex)
                barrier.abort()

        th = [threading.Thread(target=run, args=(i,)) for i in range(2)]
        [t.start() for t in th]
        [t.join() for t in th]
        for h in halves:
            h.close()
    assert not errs, errs
    for o in outs:
        assert np.array_equal(o["ci_lb"], whole["ci_lb"]) and np.array_equal(o["ci_ub"], whole["ci_ub"])
        assert np.array_equal(o["min"], whole["min"]) and np.array_equal(o["max"], whole["max"])
        assert H.rel_err(o["mean"], whole["mean"]) < 1e-13


def test_bracketed_select_with_constant_and_heavily_tied_design_points():
    """A design point where every sample gives the same value (and one where half of them do): the bracket collapses to
    a == b, the values are counted instead of collected, and the answer is still the exact order statistic."""
    from pybitup_b200 import synthetic
    from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig
    from pybitup_b200.propagation import DevicePropagation
    p = synthetic.cubic_problem(n_points=7)
    S = 300_000
    rng = np.random.default_rng(2)
    d = len(p["x0"])
    Xc = np.repeat(p["x0"][None, :], S, axis=0)                      # every sample identical: all points constant
    Xh = Xc * (1.0 + 0.01 * rng.standard_normal((S, d)) * (rng.random((S, 1)) < 0.5))      # half of the samples identical
    with ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo="RWMH"), 1) as eng:
        for X, pcts in ((Xc, (2.5, 97.5)), (Xh, (40.0, 60.0)), (Xh, (2.5, 97.5))):
            with DevicePropagation(eng, S) as dp:
                dp.evaluate(X)
                st = dp.statistics(pcts, bracket=True)
                assert dp.last_select == "bracket"
                fe = dp.evals()
            for q in pcts:
                assert np.array_equal(st["p%g" % q], np.percentile(fe, q, axis=0))
            assert np.array_equal(st["min"], fe.min(axis=0)) and np.array_equal(st["max"], fe.max(axis=0))
