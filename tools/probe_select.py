"""Time the statistics passes of the Monte-Carlo propagation one by one (tools/probe_select.py [S])."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybitup_b200 import synthetic, _capi as capi                              # noqa: E402
from pybitup_b200.engine import ChainEngine, ProblemSpec, SamplerConfig      # noqa: E402
from pybitup_b200.propagation import DevicePropagation, percentile_ranks     # noqa: E402

S = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2500000
p = synthetic.arrhenius_problem()
with ChainEngine(ProblemSpec.from_dict(p), SamplerConfig(algo="RWMH"), 1) as eng:
    with DevicePropagation(eng, S) as dp:
        dp.evaluate_gaussian(p["x0"], 0.01 * np.abs(p["x0"]), seed=0, sample_offset=0)
        torch.cuda.synchronize()
        N = dp.n_points
        t0 = time.perf_counter()
        s, mn, mx = np.empty(N), np.empty(N), np.empty(N)
        capi.check(capi.lib.pbu_propagation_moments(dp._h, capi.dptr(s), capi.dptr(mn), capi.dptr(mx)))
        torch.cuda.synchronize()
        print("moments %.2f ms  (%.0f GB/s)" % ((time.perf_counter() - t0) * 1e3, S * N * 8 / (time.perf_counter() - t0) / 1e9))
        ranks = []
        for q in (2.5, 97.5):
            ranks.extend(percentile_ranks(S, q))
        r = np.ascontiguousarray(np.asarray(ranks, dtype=np.int64))
        first = C.c_int(0)
        capi.check(capi.lib.pbu_propagation_select_begin(dp._h, r.ctypes.data_as(C.POINTER(C.c_int64)), len(ranks), capi.dptr(mn), capi.dptr(mx), C.byref(first)))
        hist = np.empty((N, len(ranks), 256), dtype=np.uint64)
        hp = hist.ctypes.data_as(C.POINTER(C.c_uint64))
        print('first pass', first.value)
        for ps in range(first.value, 8):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            capi.check(capi.lib.pbu_propagation_select_pass(dp._h, ps, hp))
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            capi.check(capi.lib.pbu_propagation_select_update(dp._h, ps, hp))
            torch.cuda.synchronize()
            print("pass %d: %.2f ms (%.0f GB/s), update %.2f ms, nonzero bins per (row, rank): %.1f"
                  % (ps, (t1 - t0) * 1e3, S * N * 8 / (t1 - t0) / 1e9, (time.perf_counter() - t1) * 1e3, np.count_nonzero(hist) / (N * len(ranks))))
