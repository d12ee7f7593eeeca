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
