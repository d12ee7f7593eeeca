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
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_OK " + json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
