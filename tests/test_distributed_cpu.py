"""Host logic of the multi-GPU path on CPU: sharding, the gloo all-reduce of chain sums (world_size 2) and the
pooled covariance / R-hat formulas, against plain numpy over all chains."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from pybitup_b200 import diagnostics
from pybitup_b200.distributed import shard_range


def _chains(seed=0, C=12, T=50, d=3):
    rng = np.random.default_rng(seed)
    base = rng.standard_normal((C, 1, d)) * 0.3 + np.array([5.0, -2.0, 100.0])
    return base + rng.standard_normal((C, T, d)) * np.array([1.0, 0.1, 3.0])


def _moments(ch):
    mean = ch.mean(axis=1)
    dev = ch - mean[:, None, :]
    M2 = np.einsum("cti,ctj->cij", dev, dev)
    return np.full(ch.shape[0], ch.shape[1]), mean, M2


def _reference(ch):
    C, T, d = ch.shape
    allx = ch.reshape(-1, d)
    W = ch.var(axis=1, ddof=1).mean(axis=0)
    B_over_n = ch.mean(axis=1).var(axis=0, ddof=1)
    rhat = np.sqrt(((T - 1) / T * W + B_over_n) / W)
    return allx.mean(axis=0), np.cov(allx, rowvar=False), rhat


def test_shard_range_covers_everything():
    for n, w in ((10, 3), (262144, 8), (7, 8), (1, 1)):
        got = []
        for r in range(w):
            f, c = shard_range(n, w, r)
            got.extend(range(f, f + c))
        assert got == list(range(n))


def test_finalize_single_rank_matches_numpy():
    ch = _chains()
    cnt, mean, M2 = _moments(ch)
    from pybitup_b200.distributed import global_chain_statistics
    st = global_chain_statistics(lambda mu: diagnostics.chain_sums_numpy(cnt, mean, M2, mu), 3)
    m, cov, rhat = _reference(ch)
    assert np.allclose(st["mean"], m, rtol=1e-13) and np.allclose(st["cov"], cov, rtol=1e-11)
    assert np.allclose(st["rhat"], rhat, rtol=1e-11)
    assert st["n_chains"] == 12 and st["n"] == 12 * 50


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pybitup_b200.distributed import Reducer, global_chain_statistics, shard_range
    ch = _chains()
    first, count = shard_range(ch.shape[0], world, rank)
    cnt, mean, M2 = _moments(ch[first:first + count])
    red = Reducer()
    st = global_chain_statistics(lambda mu: diagnostics.chain_sums_numpy(cnt, mean, M2, mu), 3, red)
    mn = red(np.array([float(rank)]), "min")
    mx = red(np.array([float(rank)]), "max")
    hist = red(np.arange(6, dtype=np.int64).reshape(2, 3) * (rank + 1), "sum")
    # the packed reductions of the bracketed select (propagation.py): brackets as (a, -b) under min with open ends, counts
    # as float64 under sum (exact below 2^53); gloo has no in-place device reduction, so the host path is taken
    ab = red(np.array([[-np.inf, 0.25 + rank], [-(1.0 + rank), -np.inf]]), "min")
    big = red(np.array([float(2 ** 51 + rank), 12500000.0 * (rank + 1)]), "sum")
    q.put((rank, st["mean"], st["cov"], st["rhat"], st["n_chains"], mn[0], mx[0], hist, ab, big, red.device_sum is None))
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_matches_numpy():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m, cov, rhat = _reference(_chains())
    for rank, mean_r, cov_r, rhat_r, nch, mn, mx, hist, ab, big, host_path in res:
        assert np.array_equal(ab, np.array([[-np.inf, 0.25], [-2.0, -np.inf]])) and host_path
        assert big[0] == float(2 ** 52 + 1) and big[1] == 37500000.0
        assert np.allclose(mean_r, m, rtol=1e-13) and np.allclose(cov_r, cov, rtol=1e-11) and np.allclose(rhat_r, rhat, rtol=1e-11)
        assert nch == 12 and mn == 0.0 and mx == 1.0
        assert np.array_equal(hist, np.arange(6).reshape(2, 3) * 3)
    # both ranks hold the bit-identical result
    assert np.array_equal(res[0][2], res[1][2])


def test_ess_batch_means_iid_and_correlated():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((20000, 2))
    assert 0.7 * 20000 < diagnostics.ess_batch_means(x) < 1.4 * 20000
    y = np.zeros_like(x)
    for t in range(1, len(x)):
        y[t] = 0.9 * y[t - 1] + x[t]
    assert diagnostics.ess_batch_means(y) < 0.15 * 20000


# ---- the rank-aware sampling driver (host logic only: no engine is created before run_algorithm) -------------------------
def _sampler_worker(rank, world, port, workdir, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    os.chdir(workdir)
    from pybitup_b200 import bayesian_inference as bi, distributions, inference_problem, solve_problem, synthetic
    from pybitup_b200 import metropolis_hastings_algorithms as mha
    job = solve_problem.SolveProblem("case.json")          # initialises the process group (gloo here), rank-aware handles
    assert dist.is_initialized() and dist.get_world_size() == world
    job.IO_util['fileID']['MChains'] = job._open(os.path.join(workdir, "output", "mcmc_chain.csv"), "w+")
    p = synthetic.cubic_problem(n_points=20)

    class Poly(bi.Model):
        device_model = "poly"

    m = Poly()
    m.x, m.param, m.param_nom, m.param_names, m.unpar_name = p["design"]["x"], p["theta_nom"].copy(), p["theta_nom"].copy(), list("abcd"), list("abcd")
    data = {"m": bi.Data("y", m.x, {0: p["y"][0]}, p["std_y"])}
    prior = distributions.set_probability_dist(["Mixture"], [["Uniform"] * 4, [[-10.0, 10.0]] * 4], 4)
    post = bi.BayesianPosterior(prior, bi.Likelihood(data, {"m": m}, 1.0), m, p["x0"].copy())
    opt = {"n_chains": 11, "init_jitter": 0.05, "seed": 3, "pooled_covariance": "yes"}
    s = mha.AdaptiveMetropolisHastings(job.IO_util, 100, p["x0"].copy(), p["prop_cov"], post, 100, 10, 1e-10, opt)
    cfg = s._config()
    q.put((rank, s.chain_first, s.n_chains, s._start_points(), s.writer, job.IO_util['fileID']['MChains'].name, cfg.algo,
           [s._launch_len(done) for done in (0, 3, 10, 95)], s._pooled_due(20), s._pooled_due(25)))
    job.close()
    dist.destroy_process_group()


def test_two_rank_sampler_sharding(tmp_path):
    """Under torchrun-style ranks the sampler shards the GLOBAL chain index space, draws shard-invariant starting points,
    gives rank 0 alone the output files, and in pooled mode cuts its launches at the refresh interval."""
    with open(tmp_path / "case.json", "w") as f:
        f.write("{}")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sampler_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, f0, c0, x0, w0, name0, algo0, lens0, due20, due25), (r1, f1, c1, x1, w1, name1, algo1, _, _, _) = res
    assert (f0, c0, f1, c1) == (0, 6, 6, 5) and w0 and not w1
    assert name0.endswith("mcmc_chain.csv") and name1 == os.devnull
    assert algo0 == algo1 == "RWMH"                       # pooled mode: the non-adaptive kernel with a refreshed factor
    assert lens0 == [10, 7, 10, 5] and due20 and not due25
    # shard-invariant starting points: the two shards stacked are what a single process draws
    from pybitup_b200 import synthetic
    p = synthetic.cubic_problem(n_points=20)
    x_all = np.tile(p["x0"], (11, 1))
    rng = np.random.default_rng(3 + 7919)
    x_all[1:] += 0.05 * np.abs(p["x0"]) * rng.standard_normal((10, 4))
    assert np.array_equal(np.vstack([x0, x1]), x_all)
