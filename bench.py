#!/usr/bin/env python
"""Headline benchmark of the batched-chain MCMC hot path (BASELINE.json metric: chain-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1], the configuration the metric is quoted on for one GPU):
cubic polynomial model, d = 4 parameters, N = 1000 data points, adaptive Metropolis,
65,536 chains per B200 (weak scaling: every rank owns its own 65,536 chains, no data-path collective).
One bench "step" = ITERS_PER_STEP sampler iterations of all chains = one launch of the fused step kernel.

What is timed
  value    chain states resident in HBM, CUDA events around each kernel launch on the launching stream,
           L2 flushed between timed launches, max over ranks (barrier + synchronize on both sides).
  e2e      the same through the public host API with HOST buffers: every step uploads the chains' starting
           points from pinned host memory, initialises the chains, runs ITERS_PER_STEP iterations, reduces the
           pooled moments of ALL chains (pooled mean / covariance, Gelman-Rubin R-hat; at N > 1 this is the NCCL
           all-reduce of the chain sums, the only collective of the path) and reads the final states,
           log-posteriors, counters and statistics back into pinned host buffers.  At N > 1 rank 0 then re-runs
           every rank's chains on its own GPU and checks the pooled moments (shard == whole).
  roofline the step kernel is FP64-pipe bound (the data shared by all chains stays in shared
           memory, so HBM does not bind); achieved = algorithmic flops per launch / mean launch duration,
           peak = a DFMA micro-benchmark run live on the same GPU (MEASURED_PEAKS.json has no FP64 entry).
           The HBM view (algorithmic bytes / duration vs MEASURED_PEAKS.json hbm_gbs) is reported beside it.
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, a byte-for-byte copy of the package made by
           oracle/make_ref.py) through Sampling(json).sample(models), one chain on one core, on a bounded sample
           of the same workload (rank 0, N = 1 only); three labelled figures: as shipped, with the per-iteration
           chain-file writes stubbed, and the numpy port.
  configs  the other BASELINE.json configurations (1, 1b, 3 weak + strong, 4 with its moment all-reduce, 5) at their
           per-GPU sizes, each a short device-timed measurement with its own roofline, in the same run.

--impl reference times the unmodified reference's CPU implementation of the path on all host cores, one chain
per process through Sampling(json).sample(models); the processes never import the pybitup_b200 package.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "chain-steps/sec"
UNIT = "chain-steps/s"
CHAINS_PER_GPU = 65536
ITERS_PER_STEP = 1000
N_POINTS = 1000
D = 4
WORKLOAD = "cubic polynomial (Horner), d=4, N=1000, adaptive Metropolis (updating_it=10, eps_v=1e-10), " \
           "65536 chains per GPU, %d iterations per step" % ITERS_PER_STEP
AM = dict(updating_it=10, eps_v=1e-10)
# SURVEY.md 8(d), lower-bound convention (FMA = 2): N*(2p+4) + proposal d(d+1) + accept 3 + AM 8d^2+3d
FLOPS_PER_CHAIN_STEP = N_POINTS * (2 * 3 + 4) + D * (D + 1) + 3 + (8 * D * D + 3 * D)
# per launch and chain: state in + out (x, lp, R, xav, V, counters) ; shared data once per block
STATE_BYTES_PER_CHAIN = 2 * 8 * (D + 1 + 10 + D + 10 + 4)
DATA_BYTES = N_POINTS * 4 * 8


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for n, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        self.f.close()
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU legs
# The CPU legs time the UNMODIFIED reference (oracle/_ref, a byte-for-byte copy of the package made by oracle/make_ref.py)
# through its own entry point, Sampling(json).sample(models): one chain per process, one process per host core.  They never
# import the pybitup_b200 package (oracle/ref_bench.py loads the numpy problem builders by file path), so the CUDA
# library is not mapped into these processes.
BENCH_PROBLEM = ("cubic_problem", {"n_points": N_POINTS}, "AMH")


def bench_config(chains_per_gpu=CHAINS_PER_GPU, iters=ITERS_PER_STEP):
    """The `config` object of the JSON line: identical in both arms."""
    return {"workload": WORKLOAD, "chains_per_gpu": int(chains_per_gpu), "iterations_per_step": int(iters)}


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def run_reference_arm(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    from oracle import make_ref, ref_bench
    make_ref.make()                          # no-op on the GPU box (prebuilt copy), refresh in the build container
    cores = host_cores()
    n_it = args.ref_iters
    import multiprocessing as mp
    os.environ.setdefault("PYTHONWARNINGS", "ignore")     # the reference's own SyntaxWarnings (post_process.py) in the workers
    pool = mp.get_context("spawn").Pool(cores) if cores > 1 else None
    builder, bkw, algo = BENCH_PROBLEM
    try:
        for _ in range(min(args.warmup, 1)):
            ref_bench.rate(pool, cores, builder, bkw, algo, 500, "as_shipped")
        rates, t_all = [], 0.0
        for _ in range(args.steps):
            r, wall = ref_bench.rate(pool, cores, builder, bkw, algo, n_it, "as_shipped")
            rates.append(r)
            t_all += wall
        stub, _ = ref_bench.rate(pool, cores, builder, bkw, algo, n_it, "save_sample_stubbed")
        port, _ = ref_bench.rate(pool, cores, builder, bkw, algo, n_it, "port")
    finally:
        if pool is not None:
            pool.close()
            pool.join()
    value = float(np.mean(rates))
    sample = "%d independent chains (one process per core) x %d AM iterations of the same cubic problem per bench step, each " \
             "through the unmodified reference's Sampling(json).sample(models)" % (cores, n_it)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_all / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args.chains, args.iters),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample,
                         "variants": {"as_shipped": value, "save_sample_stubbed": float(stub), "port": float(port)},
                         "variants_note": "as_shipped = the unmodified reference (oracle/_ref), chain rows written with np.savetxt "
                                          "every iteration; save_sample_stubbed = the same with BayesianPosterior.save_sample a "
                                          "no-op after three rows; port = the numpy restatement oracle/mcmc_oracle.py"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def cpu_baseline_leg():
    """cpu_baseline of the GPU arm (rank 0, one GPU): a bounded sample in a child process, so that this process keeps the
    reference package and its import stubs out of its module table."""
    code = ("import json,sys; sys.path.insert(0, %r); from oracle import make_ref, ref_bench as rb; make_ref.make(); "
            "b=%r; n=30000; "
            "a=rb.worker((b[0], b[1], b[2], n, 'as_shipped', 0)); s=rb.worker((b[0], b[1], b[2], n, 'save_sample_stubbed', 0)); "
            "from oracle import ref_bench; syn=rb.load_synthetic(); p=getattr(syn,b[0])(**b[1]); "
            "p=dict(p, x0=syn.chain_starts(p['x0'],1,jitter=0.01,seed=0)[0]); "
            "m=rb.run_port_chain(p, rb.algo_block(b[2], 100000, p), 0); "
            "import numpy as np; np.save(sys.argv[1], m[2]); "
            "print(json.dumps({'as_shipped': a[0]/a[1], 'stubbed': s[0]/s[1], 'port': m[0]/m[1], 'n': n, 't': a[1]+s[1]+m[1]}))"
            % (ROOT, BENCH_PROBLEM))
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "chain.npy")
        out = subprocess.run([sys.executable, "-W", "ignore", "-c", code, path], capture_output=True, text=True, timeout=600)
        if out.returncode != 0:
            raise RuntimeError("cpu_baseline leg failed: " + out.stderr[-2000:])
        r = json.loads(out.stdout.strip().splitlines()[-1])
        chain = np.load(path)
    cpu = {"value": r["as_shipped"], "unit": UNIT, "cores": 1, "kind": "reference",
           "sample": "1 chain x %d AM iterations of the same cubic problem through the unmodified reference's "
                     "Sampling(json).sample(models) (oracle/_ref), %.1f s for the three variants" % (r["n"], r["t"]),
           "variants": {"as_shipped": r["as_shipped"], "save_sample_stubbed": r["stubbed"], "port": r["port"]}}
    return cpu, chain


# ------------------------------------------------------------------------------------------------ GPU arm
FP64_INST_RK4 = 68.5         # FP64 instructions of one RK4 step of the Arrhenius kernels (137 per two steps, cuobjdump -sass)
OTHER_INST_RK4 = 62.5        # the other instructions of that step: an FP64 instruction holds the issue port for two cycles, these for one
FP64_INST_RK4_EVAL = 66.5    # the evaluation kernel of the propagation (no residual, no sum of squares): 133 FP64 + 128 other per two steps
OTHER_INST_RK4_EVAL = 64.0


def _finite(o):
    """JSON has no inf / nan: replace them by null."""
    if isinstance(o, dict):
        return {k: _finite(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_finite(v) for v in o]
    if isinstance(o, float) and not np.isfinite(o):
        return None
    return o


def _timed(fn, reps, barrier, max_over_ranks, flush=None):
    """median over reps of the device time of fn() in ms (CUDA events on the current stream, max over ranks)"""
    import torch
    ms = []
    for _ in range(reps):
        if flush is not None:
            flush()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ms.append(max_over_ranks(e0.elapsed_time(e1)))
    return float(np.median(ms))


def measure_configs(rank, local_rank, world, red, fp64_peak, hbm_peak, barrier, max_over_ranks, flush, which):
    """The other BASELINE.json configurations at their per-GPU sizes, each a short measurement in the same run (device
    timed, L2 flushed before every timed launch, max over ranks).  Flop / byte counts per unit: SURVEY.md 8(d)."""
    import torch
    from pybitup_b200 import synthetic
    from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig
    from pybitup_b200.propagation import DevicePropagation
    out = {}

    def frac(tflops):
        return tflops / fp64_peak

    if "1" in which and rank == 0:
        # cfg 1: ONE chain x 1e5 iterations (the reference's own CPU-runnable case): latency of a serial chain, no roofline
        p = synthetic.linear_problem()
        with ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo="RWMH", seed=1), 1, device=local_rank) as eng:
            eng.init_chains(p["x0"][None, :], p["prop_cov"])
            eng.run(1000, thin=1000, keep_samples=False)
            ms = _timed(lambda: eng.run(100000, thin=100000, keep_samples=False), 3, lambda: torch.cuda.synchronize(), lambda v: v)
            out["cfg1"] = {"workload": "linear model d=3 N=50, RWMH, 1 chain x 1e5 iterations (one GPU thread group; rank 0 only)",
                           "ms": ms, "ns_per_chain_step": ms * 1e6 / 1e5, "chain_steps_per_s": 1e5 / (ms * 1e-3),
                           "roofline": None, "note": "a single serial chain is bound by dependent-instruction latency: no throughput roofline applies"}
    if "1b" in which:
        p = synthetic.linear_problem()
        C, T = 1 << 20, 200
        with ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo="RWMH", seed=1, chain_offset=rank * C), C, device=local_rank) as eng:
            eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=1e-3, seed=rank), p["prop_cov"])
            eng.run(T, thin=T, keep_samples=False)
            ms = _timed(lambda: eng.run(T, thin=T, keep_samples=False), 3, barrier, max_over_ranks, flush)
            r = world * C * T / (ms * 1e-3)
            tf = r / world * 512 / 1e12
            out["cfg1b"] = {"workload": "linear model d=3 N=50, RWMH, 1,048,576 chains per GPU x %d iterations" % T, "scaling": "weak",
                            "ms": ms, "chain_steps_per_s": r, "launch": eng.launch_geometry(),
                            "roofline": {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": frac(tf),
                                         "flops_per_chain_step": 512}}
    if "3" in which:
        p = synthetic.arrhenius_problem()
        T = 20
        for mode, C in (("weak", 262144), ("strong", 262144 // world)):
            if mode == "strong" and world == 1:
                continue
            first = rank * C
            cfg = SamplerConfig(algo="DRAM", seed=1, updating_it=10, eps_v=1e-10, gamma_dr=0.2, chain_offset=first)
            with ChainEngine(ProblemSpec.from_dict(p), cfg, C, device=local_rank) as eng:
                eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=1e-3, seed=rank), p["prop_cov"])
                eng.run(T, thin=T, keep_samples=False)
                ms = _timed(lambda: eng.run(T, thin=T, keep_samples=False), 3, barrier, max_over_ranks, flush)
                ev = float(eng.read_state()["n_eval"].mean()) / eng.iteration
                # speculative delayed rejection (two lanes per chain): both stages are evaluated every step, the second one
                # for nothing when the first proposal is accepted; n_eval counts what the reference would have evaluated
                ev_run = 2.0 if eng.launch_geometry()["lanes"] == 2 else ev
                r = world * C * T / (ms * 1e-3)
                tf_alg = r / world * ev * 2.17e4 / 1e12
                tf_issue = r / world * ev_run * 500 * FP64_INST_RK4 * 2 / 1e12
                out["cfg3" if mode == "weak" else "cfg3_strong"] = {
                    "workload": "Arrhenius pyrolysis ODE (RK4 x 500), d=4, DRAM, %d chains per GPU (%d in all) x %d iterations" % (C, world * C, T),
                    "scaling": mode, "ms": ms, "chain_steps_per_s": r, "evals_per_chain_step": ev, "evals_executed_per_chain_step": ev_run,
                    "launch": eng.launch_geometry(),
                    "roofline": {"bound": "fp64", "achieved": tf_alg, "peak": fp64_peak, "unit": "TFLOP/s", "frac": frac(tf_alg),
                                 "flops_per_evaluation": 2.17e4,
                                 "fp64_issue": {"achieved": tf_issue, "frac": frac(tf_issue), "fp64_inst_per_rk4_step": FP64_INST_RK4,
                                                "other_inst_per_rk4_step": OTHER_INST_RK4,
                                                "issue_slot_bound_frac": 2 * FP64_INST_RK4 / (2 * FP64_INST_RK4 + OTHER_INST_RK4),
                                                "note": "FP64 instructions the kernel issues, each counted as an FMA (2 flops): what the pipe "
                                                        "is busy with; the algorithmic count takes exp / pow as 1 flop each.  "
                                                        "issue_slot_bound_frac is the most this instruction mix can reach "
                                                        "(2 issue cycles per FP64 instruction, 1 per other)"}}}
    if "4" in which:
        p = synthetic.regression_problem()
        C, T = 131072, 4
        cfg = SamplerConfig(algo="ISDE", seed=1, isde_h=0.05, isde_f0=4.0, isde_gradient="Analytical", chain_offset=rank * C)
        with ChainEngine(ProblemSpec.from_dict(p), cfg, C, device=local_rank) as eng:
            eng.init_chains(synthetic.chain_starts(p["x0"], C, jitter=1e-2, seed=rank), np.eye(10) * 2.5e-5)
            eng.run(T, thin=1, keep_samples=False)
            eng.pooled_statistics(red, apply=True)
            ms = _timed(lambda: eng.run(T, thin=1, keep_samples=False), 3, barrier, max_over_ranks, flush)
            ms_stat = _timed(lambda: eng.pooled_statistics(red, apply=True), 5, barrier, max_over_ranks)
            st = eng.statistics_dict(eng.pooled_statistics(red))
            r = world * C * T / (ms * 1e-3)
            tf = r / world * 4.4e5 / 1e12
            out["cfg4"] = {"workload": "correlated Gaussian regression d=10 N=1e4 (records streamed), Ito-SDE, 131,072 chains per GPU "
                                       "(%d in all) x %d iterations, pooled preconditioner" % (world * C, T), "scaling": "weak", "ms": ms,
                           "chain_steps_per_s": r, "launch": eng.launch_geometry(),
                           "pooled_moment_allreduce_ms": ms_stat,
                           "pooled_moment_allreduce_what": "two passes over the per-chain moments, %s, finalisation (pooled covariance, R-hat, "
                                                           "Cholesky) and the refresh of every chain's preconditioner, all on the device"
                                                           % ("two NCCL all-reduces (12 + 97 doubles)" if world > 1 else "no collective on one GPU"),
                           "chains_total": st["n_chains"], "rhat_max_after_%d_iterations_from_dispersed_starts" % (5 * T): float(np.nanmax(st["rhat"])),
                           "roofline": {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": frac(tf),
                                        "flops_per_chain_step": 4.4e5}}
    if "5" in which:
        p = synthetic.arrhenius_problem()
        S = 12500000
        with ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo="RWMH"), 1, device=local_rank) as eng:
            with DevicePropagation(eng, S) as dp:
                mean, std = p["x0"], 0.01 * np.abs(p["x0"])
                dp.evaluate_gaussian(mean, std, seed=0, sample_offset=rank * S)
                ms_eval = _timed(lambda: dp.evaluate_gaussian(mean, std, seed=0, sample_offset=rank * S), 2, barrier, max_over_ranks)
                rd = red if (red is not None and red.active and world > 1) else None
                dp.statistics(reduce=rd, n_total=world * S)
                barrier()
                t0 = time.perf_counter()
                st = dp.statistics(reduce=rd, n_total=world * S)
                torch.cuda.synchronize()
                ms_stat = max_over_ranks((time.perf_counter() - t0) * 1e3)
                tf_issue = S * 500 * FP64_INST_RK4_EVAL * 2 / (ms_eval * 1e-3) / 1e12
                out["cfg5"] = {"workload": "Monte-Carlo propagation through the pyrolysis ODE, 1.25e7 samples per GPU (%.3g in all) x 500 "
                                           "points, evaluations materialised in HBM (50 GB per GPU)" % (world * S), "scaling": "weak",
                               "eval_ms": ms_eval, "statistics_ms": ms_stat, "statistics_select": dp.last_select,
                               "samples_per_s_eval": world * S / (ms_eval * 1e-3),
                               "samples_per_s_total": world * S / ((ms_eval + ms_stat) * 1e-3),
                               "ci_last_point": [float(st["ci_lb"][-1]), float(st["ci_ub"][-1])],
                               "roofline": {"bound": "fp64", "achieved": S * 2.15e4 / (ms_eval * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                                            "frac": frac(S * 2.15e4 / (ms_eval * 1e-3) / 1e12), "flops_per_sample": 2.15e4,
                                            "fp64_issue": {"achieved": tf_issue, "frac": frac(tf_issue), "fp64_inst_per_rk4_step": FP64_INST_RK4_EVAL,
                                                           "other_inst_per_rk4_step": OTHER_INST_RK4_EVAL,
                                                           "kernel": "eval_model_rep_kernel<ArrheniusModel>: conflict-free replicated exp2 / log2 "
                                                                     "tables (192 KB of shared memory), 1024-thread persistent blocks"},
                                            "hbm_write": {"achieved": S * 4000 / (ms_eval * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                                          "frac": S * 4000 / (ms_eval * 1e-3) / 1e9 / hbm_peak}}}
    return out


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from pybitup_b200 import _capi, synthetic
    from pybitup_b200.distributed import Reducer
    from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    red = Reducer()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    C, T = args.chains, args.iters
    K, W = args.steps, args.warmup
    p = synthetic.cubic_problem(n_points=N_POINTS)
    spec = ProblemSpec.from_dict(p)
    cfg = SamplerConfig(algo="AMH", seed=2026, chain_offset=rank * C, **AM)
    x0 = synthetic.chain_starts(p["x0"], C, jitter=0.01, seed=1000 + rank)

    fp64_burst, _ = _capi.measure_fp64_peak(local_rank, 4096, 5)
    fp64_peak, _ = _capi.measure_fp64_peak(local_rank, 4096, -150)        # ~0.7 s of back-to-back DFMA launches
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def flush():
        flush_buf.fill_(1)

    # ---- device-resident timing -----------------------------------------------------------------
    eng = ChainEngine(spec, cfg, C, device=local_rank)
    eng.init_chains(x0, p["prop_cov"])
    geom = eng.launch_geometry()
    coop = bool(geom["cooperative"])                                              # teams of G lanes own G chains
    # PolyModelU: the uniform-weight form of the polynomial model (one std_y for all points: the weight is folded into the
    # coefficients once per evaluation, records [x | w*y]; pbu_models.cuh) -- picked by the engine when the data allow it
    kernel_name = "%s<%s<4>,D=4,G=%d,ALGO=AM%s>" % ("mh_coop_kernel" if coop else "mh_step_kernel",
                                                     "PolyModelU" if geom["uniform_weight"] else "PolyModel", geom["lanes"],
                                                     ",MIXED" if geom["mixed"] else "")
    for _ in range(W):
        eng.run(T, thin=T, keep_samples=True)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    for k in range(K):
        flush_buf.fill_(k & 0xff)                                          # L2 flush between timed launches
        ev[k][0].record()
        eng.run(T, thin=T, keep_samples=True)
        ev[k][1].record()
    barrier()
    clk = clocks.stop()
    ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = max_over_ranks(sum(ms))
    ctr = eng.read_state()
    acc_rate = 1.0 - float(ctr["n_rejected"].mean()) / max(eng.iteration, 1)
    bad = int(np.count_nonzero(ctr["status"]))
    eng.close()
    value = world * C * T * K / (total_ms * 1e-3)

    # ---- end to end through the host API --------------------------------------------------------
    # every step: starting points from pinned host memory -> device, chain initialisation, T iterations, the pooled
    # moments of ALL chains (at N > 1: the NCCL all-reduce of the chain sums, pooled covariance and R-hat), and the final
    # states, log-posteriors, counters and statistics back into pinned host buffers
    x0_pin = torch.from_numpy(x0).pin_memory()
    x0_host = x0_pin.numpy()
    eng = ChainEngine(spec, cfg, C, device=local_rank)
    n_stats = int(_capi.lib.pbu_engine_pooled_stats_len(eng._h))
    stats_pin = torch.empty(n_stats, dtype=torch.float64).pin_memory()
    h2d = C * 8 * D                               # the chains' starting points (everything else is built on the device)
    d2h = C * 8 * (D + 1) + C * (8 + 8 + 8 + 4) + 8 * n_stats

    def e2e_step():
        eng.init_chains(x0_host, p["prop_cov"])
        eng.run(T, thin=max(T // 10, 1), keep_samples=False)     # ten kept states per chain feed the moments (R-hat needs > 1)
        stats = eng.pooled_statistics(red)
        stats_pin.copy_(stats, non_blocking=True)
        return eng.read_state()                    # one synchronisation: everything above has landed

    for _ in range(min(W, 3)):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record()
    for k in range(K):
        st_host = e2e_step()
    e1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    t_e2e = max_over_ranks(max(e0.elapsed_time(e1), 1e3 * t_wall))
    e2e_value = world * C * T * K / (t_e2e * 1e-3)
    pooled = eng.statistics_dict(stats_pin)
    ms_stat = _timed(lambda: eng.pooled_statistics(red), 5, barrier, max_over_ranks)
    eng.close()

    # ---- shard == whole (N > 1): rank 0 runs every rank's chains on its own GPU and must find the same pooled moments ----
    shard_check = None
    if world > 1:
        if rank == 0:
            x_all = np.concatenate([synthetic.chain_starts(p["x0"], C, jitter=0.01, seed=1000 + r) for r in range(world)], axis=0)
            with ChainEngine(spec, SamplerConfig(algo="AMH", seed=2026, chain_offset=0, **AM), world * C, device=local_rank) as whole:
                whole.init_chains(x_all, p["prop_cov"])
                whole.run(T, thin=max(T // 10, 1), keep_samples=False)
                ref = whole.statistics()
            err = max(float(np.max(np.abs(pooled["mean"] - ref["mean"]) / np.abs(ref["mean"]))),
                      float(np.max(np.abs(pooled["cov"] - ref["cov"]) / np.abs(ref["cov"]))))
            shard_check = {"chains": int(pooled["n_chains"]), "max_rel_err_pooled_mean_cov_vs_one_gpu": err, "ok": bool(err < 1e-9)}
            assert pooled["n_chains"] == world * C and err < 1e-9, shard_check
        barrier()

    # ---- effective sample size per chain-step (same batch-means estimator on both sides) ----------
    ess = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from pybitup_b200.diagnostics import ess_batch_means
        n_probe, T_ess = 16, 24000           # the CPU chain below is cut into segments of the same length
        with ChainEngine(spec, cfg, 4096, device=local_rank) as eng2:
            eng2.init_chains(synthetic.chain_starts(p["x0"], 4096, jitter=0.01, seed=5), p["prop_cov"])
            eng2.run(2000, thin=2000, keep_samples=False)                 # burn-in / adaptation
            res = eng2.run(T_ess, thin=1, sample_chains=n_probe)
            eng2.synchronize()
            smp = res.samples.cpu().numpy()                               # [T, d, n_probe]
        ess_step = float(np.nanmean([ess_batch_means(smp[:, :, c]) for c in range(n_probe)])) / T_ess
        ess = {"estimator": "batch means, min over parameters", "gpu_ess_per_chain_step": ess_step,
               "gpu_ess_per_s": ess_step * value}

    # ---- CPU baseline (rank 0, single GPU runs only): the unmodified reference, in a child process -----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu, chain = cpu_baseline_leg()
        if ess is not None:
            from pybitup_b200.diagnostics import ess_batch_means
            segs = [chain[2000 + i * 24000: 2000 + (i + 1) * 24000] for i in range((len(chain) - 2000) // 24000)]
            e_cpu = float(np.mean([ess_batch_means(sg) for sg in segs])) / 24000
            ess.update({"cpu_ess_per_chain_step": e_cpu, "cpu_ess_per_s": e_cpu * cpu["variants"]["port"],
                        "cpu_ess_note": "chain of the numpy port (same sampler, injected draws); rate of the port"})

    # ---- the other configurations, in the same run --------------------------------------------------
    configs = None
    if not args.no_configs:
        configs = measure_configs(rank, local_rank, world, red, fp64_peak, hbm_peak, barrier, max_over_ranks, flush,
                                  [c.strip() for c in args.configs.split(",")])

    if rank == 0:
        ms_launch = total_ms / K
        flops_launch = float(FLOPS_PER_CHAIN_STEP) * C * T
        achieved = flops_launch / (ms_launch * 1e-3) / 1e12
        bytes_launch = float(STATE_BYTES_PER_CHAIN) * C + DATA_BYTES * ((C + 127) // 128)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": bench_config(C, T),
            "run": {"rng": "Philox4x32-10", "l2": "flushed between timed launches (256 MiB fill)", "acceptance_rate": acc_rate,
                    "chains_with_status_flags": bad},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "includes": "upload of the starting points from pinned memory, chain initialisation, %d iterations, pooled "
                                "moments of all chains (%s), final states / counters / statistics into pinned host buffers"
                                % (T, "NCCL all-reduce of the chain sums over %d ranks" % world if world > 1 else "one GPU: no collective"),
                    "pooled_statistics_ms": ms_stat, "pooled_rhat_max": float(np.nanmax(pooled["rhat"])),
                    "pooled_chains": int(pooled["n_chains"]), "shard_equals_whole": shard_check},
            "gpu_launches": K,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak, "traffic": None,
                         "kernel": kernel_name, "launch": geom,
                         "flops_per_chain_step": FLOPS_PER_CHAIN_STEP,
                         "fp64_instructions_per_point": "5 = 3 Horner DFMA + 1 residual (DFMA -f*w + w*y; a DADD w*y - (w f) in the uniform-weight "
                                                        "form) + 1 square-accumulate DFMA: the same count in both forms",
                         "peak_source": "live DFMA micro-benchmark on this GPU, sustained over ~0.7 s of back-to-back launches "
                                        "(pbu_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
                         "peak_burst": fp64_burst, "ms_per_launch": {"min": float(min(ms)), "median": float(np.median(ms)),
                                                                     "max": float(max(ms))},
                         "hbm": {"achieved": bytes_launch / (ms_launch * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": bytes_launch / (ms_launch * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src}},
            "cpu_baseline": cpu,
            "ess": ess,
            "configs": configs,
        }
        prof = os.path.join(ROOT, "profiles", "traffic_r2.json")
        if not os.path.exists(prof):
            prof = os.path.join(ROOT, "profiles", "traffic_r1.json")
        if os.path.exists(prof):
            try:
                with open(prof) as f:
                    line["roofline"]["traffic"] = json.load(f).get("dram_bytes_per_launch")
            except (OSError, ValueError):
                pass
        print(json.dumps(_finite(line)), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chains", type=int, default=CHAINS_PER_GPU, help="chains per GPU")
    ap.add_argument("--iters", type=int, default=ITERS_PER_STEP, help="sampler iterations per bench step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block (the other BASELINE.json configurations)")
    ap.add_argument("--configs", default="1,1b,3,4,5", help="which of the other configurations to measure")
    ap.add_argument("--ref-iters", type=int, default=20000, help="reference arm: sampler iterations per chain and bench step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
