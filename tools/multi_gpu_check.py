"""Run under torchrun (one rank per GPU): chains and propagation samples sharded over ranks give the same pooled
statistics as one GPU running everything (Philox streams are keyed by the GLOBAL chain / sample index)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybitup_b200 import synthetic                                   # noqa: E402
from pybitup_b200.distributed import Reducer, shard_range           # noqa: E402
from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig  # noqa: E402
from pybitup_b200.propagation import DevicePropagation              # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    red = Reducer()
    p = synthetic.cubic_problem(n_points=200)
    spec = ProblemSpec.from_dict(p)
    C_total, T = 1000, 150
    x0_all = synthetic.chain_starts(p["x0"], C_total, jitter=0.01, seed=3)
    first, count = shard_range(C_total, world, rank)
    out = {}
    for algo, kw in (("AMH", dict(updating_it=10, eps_v=1e-10)), ("ISDE", dict(isde_h=0.01, isde_f0=4.0))):
        cfg = SamplerConfig(algo=algo, seed=9, chain_offset=first, **kw)
        with ChainEngine(spec, cfg, count, device=local) as eng:
            eng.init_chains(x0_all[first:first + count], p["prop_cov"] if algo == "AMH" else None)
            eng.run(T, thin=1, keep_samples=False)
            st = eng.statistics(red)
        if rank == 0:
            cfg = SamplerConfig(algo=algo, seed=9, chain_offset=0, **kw)
            with ChainEngine(spec, cfg, C_total, device=local) as eng:
                eng.init_chains(x0_all, p["prop_cov"] if algo == "AMH" else None)
                eng.run(T, thin=1, keep_samples=False)
                ref = eng.statistics()
            err = max(np.max(np.abs(st["mean"] - ref["mean"]) / np.abs(ref["mean"])), np.max(np.abs(st["cov"] - ref["cov"]) / np.abs(ref["cov"])),
                      np.max(np.abs(st["rhat"] - ref["rhat"])))
            out[algo] = dict(n_chains=st["n_chains"], err=float(err), rhat=st["rhat"].tolist())
            assert st["n_chains"] == C_total and err < 1e-10, (algo, err)
    # propagation: samples drawn on the device, sharded
    pa = synthetic.arrhenius_problem(n_steps=50, dT=12.0)
    S = 8192
    s_first, s_count = shard_range(S, world, rank)
    with ChainEngine(ProblemSpec.from_dict(pa), SamplerConfig(algo="RWMH"), 1, device=local) as eng:
        with DevicePropagation(eng, s_count) as dp:
            dp.evaluate_gaussian(pa["x0"], 0.01 * np.abs(pa["x0"]), seed=5, sample_offset=s_first)
            st = dp.statistics(reduce=red, n_total=S)
        if rank == 0:
            with DevicePropagation(eng, S) as dp:
                dp.evaluate_gaussian(pa["x0"], 0.01 * np.abs(pa["x0"]), seed=5, sample_offset=0)
                ref = dp.statistics()
            assert np.array_equal(st["ci_lb"], ref["ci_lb"]) and np.array_equal(st["ci_ub"], ref["ci_ub"])
            assert np.array_equal(st["min"], ref["min"]) and np.array_equal(st["max"], ref["max"])
            assert np.allclose(st["mean"], ref["mean"], rtol=1e-13)
            out["propagation"] = "sharded == unsharded"
    # the one-pass bracketed select over NCCL (brackets min / max-reduced, counts and candidate histograms summed)
    S = 1 << 21
    s_first, s_count = shard_range(S, world, rank)
    with ChainEngine(ProblemSpec.from_dict(pa), SamplerConfig(algo="RWMH"), 1, device=local) as eng:
        with DevicePropagation(eng, s_count) as dp:
            dp.evaluate_gaussian(pa["x0"], 0.01 * np.abs(pa["x0"]), seed=6, sample_offset=s_first)
            st = dp.statistics(reduce=red, n_total=S, bracket=True)
            how = dp.last_select
        if rank == 0:
            with DevicePropagation(eng, S) as dp:
                dp.evaluate_gaussian(pa["x0"], 0.01 * np.abs(pa["x0"]), seed=6, sample_offset=0)
                ref = dp.statistics(bracket=False)
            assert how == "bracket", how
            assert np.array_equal(st["ci_lb"], ref["ci_lb"]) and np.array_equal(st["ci_ub"], ref["ci_ub"])
            assert np.array_equal(st["min"], ref["min"]) and np.array_equal(st["max"], ref["max"])
            assert np.allclose(st["mean"], ref["mean"], rtol=1e-13)
            out["propagation_bracketed"] = "sharded == unsharded (2^21 samples)"
    # ---- the product API: Sampling(json).sample(models) under torchrun shards n_chains by rank, rank 0 writes the files,
    # the pooled statistics come from the device all-reduce; pooled-covariance mode refreshes the Ito-SDE preconditioner
    # from the chains of ALL ranks.  Compared with one engine holding every chain on this rank's GPU.
    import tempfile
    import pandas as pd
    from pybitup_b200 import bayesian_inference as bi
    from pybitup_b200.solve_problem import Sampling

    class Poly(bi.Model):
        device_model = "poly"

    names = ["c0", "c1", "c2", "c3"]
    C_total, T = 4096, 300
    C0 = np.diag([1e-4, 5e-4, 5e-4, 1e-3])                # rough posterior scales: a stable start for h = 0.2
    wd = [tempfile.mkdtemp() if rank == 0 else None]
    dist.broadcast_object_list(wd, src=0)
    os.chdir(wd[0])
    if rank == 0:
        pd.DataFrame({"x": p["design"]["x"], "y": p["y"][0], "std_y": p["std_y"]}).to_csv("data.csv", index=False)
        prior = {n: {"initial_val": float(p["x0"][i]), "prior_name": "Uniform", "prior_param": [-10.0, 10.0]} for i, n in enumerate(names)}
        js = {"Sampling": {"BayesianPosterior": {
            "Data": [{"Type": "ReadFromFile", "FileName": "data", "xField": ["x"], "yField": ["y"], "sigmaField": ["std_y"]}],
            "Model": [{"param_names": names, "param_values": [float(v) for v in p["theta_nom"]], "parametrization": "no"}],
            "Prior": {"Distribution": "Mixture", "Param": prior}, "Likelihood": {"function": "Gaussian", "distance": "L2"}},
            "Algorithm": {"name": "ISDE", "n_iterations": T, "n_chains": C_total, "seed": 21, "init_jitter": 0.01,
                          "pooled_covariance": "yes", "save_eval_freq": 10 ** 9, "gradient": "Analytical", "ISDE": {"h": 0.2, "f0": 4.0},
                          "covariance": {"value": "Matrix", "matrix_value": C0.tolist(), "starting_it": 100, "update_it": 50,
                                         "end_it": 250}}}}
        with open("case.json", "w") as f:
            json.dump(js, f)
    dist.barrier()
    job = Sampling("case.json")
    run = job.sample({"m": Poly()})
    job.close()
    assert run.n_chains_total == C_total and run.world == world and run.n_chains == shard_range(C_total, world, rank)[1]
    st = run.statistics
    if rank == 0:
        assert pd.read_csv("output/mcmc_chain.csv", header=None).values.shape == (T + 1, 4)
        x0_all = np.tile(p["x0"], (C_total, 1))
        rng = np.random.default_rng(21 + 7919)
        x0_all[1:] += 0.01 * np.abs(p["x0"]) * rng.standard_normal((C_total - 1, 4))
        cfg = SamplerConfig(algo="ISDE", seed=21, isde_h=0.2, isde_f0=4.0, adapt_start=0, adapt_update=50)
        with ChainEngine(spec, cfg, C_total, device=local) as eng:
            eng.init_chains(x0_all, C0)
            for it in range(50, T + 1, 50):
                eng.run(50, thin=1, keep_samples=False)
                if 100 <= it <= 251:
                    eng.pooled_statistics(apply=True)
            ref = eng.statistics()
            Cref = eng.isde_state()["C"][0]
        err = max(np.max(np.abs(st["mean"] - ref["mean"]) / np.abs(ref["mean"])), np.max(np.abs(st["cov"] - ref["cov"]) / np.abs(ref["cov"])),
                  np.max(np.abs(st["rhat"] - ref["rhat"])), np.max(np.abs(run.C_approx - Cref) / np.abs(Cref)))
        out["sampling_json_isde_pooled"] = dict(n_chains=st["n_chains"], ranks=world, err=float(err), rhat=st["rhat"].tolist())
        assert st["n_chains"] == C_total and st["n"] == C_total * T and err < 1e-8, err
    # ---- cost of one pooled-statistics call (two passes + two all-reduces + finalisation), device timed
    C_big = 131072
    pr = synthetic.regression_problem()
    with ChainEngine(ProblemSpec.from_dict(pr), SamplerConfig(algo="ISDE", seed=1, isde_h=0.05, isde_f0=4.0, chain_offset=rank * C_big),
                     C_big, device=local) as eng:
        eng.init_chains(synthetic.chain_starts(pr["x0"], C_big, jitter=0.01, seed=rank), None)
        eng.run(2, thin=1, keep_samples=False)
        for _ in range(3):
            eng.pooled_statistics(red, apply=True)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            eng.pooled_statistics(red, apply=True)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 20], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out["pooled_statistics_ms_131072_chains_d10"] = float(ms.item())
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_OK " + json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
