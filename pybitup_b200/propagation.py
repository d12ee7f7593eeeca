"""Monte-Carlo uncertainty propagation on the device (K3).

Host-side mirror of the ``emulator == "None"`` branch of ``Propagation.propagate``
(``pybitup/solve_problem.py:467-647``): the per-sample ``run_model`` loop (:543-576) runs as one kernel over all
samples, and the statistics the reference writes (mean / min / max :596-637, the 2.5 / 97.5 percentiles :639-647)
are reduced on the device.  With samples sharded over GPUs (one process per GPU) the only exchange is a sum of the
per-point moment and histogram arrays over ranks -- ``reduce`` below -- where the reference pickles every evaluation
to rank 0 (:578-592).
"""
import ctypes as C

import numpy as np

from . import _capi as capi


def lerp_percentile(n_total, q, v_lo, v_hi):
    """np.percentile(..., method='linear') from the two bracketing order statistics (numpy's _lerp)."""
    h = (n_total - 1) * (q / 100.0)
    g = h - np.floor(h)
    diff = v_hi - v_lo
    out = v_lo + diff * g
    if g >= 0.5:
        out = v_hi - diff * (1.0 - g)
    return out


def percentile_ranks(n_total, q):
    h = (n_total - 1) * (q / 100.0)
    lo = int(np.floor(h))
    return lo, min(lo + 1, n_total - 1)


class DevicePropagation:
    """Evaluations of S samples of one problem on one GPU, resident in HBM."""

    def __init__(self, engine, n_samples):
        self.engine = engine
        self.n_samples = int(n_samples)
        self.n_points = engine.problem.n_points
        h = C.c_void_p()
        capi.check(capi.lib.pbu_propagation_create(engine._h, self.n_samples, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            capi.lib.pbu_propagation_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def evaluate(self, X):
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
        if X.shape != (self.n_samples, self.engine.d):
            raise ValueError("samples must be [S, d] = [%d, %d], got %s" % (self.n_samples, self.engine.d, X.shape))
        capi.check(capi.lib.pbu_propagation_eval_host(self._h, capi.dptr(X)))

    def evaluate_gaussian(self, mean, std, seed=0, sample_offset=0):
        mean = np.ascontiguousarray(np.asarray(mean, dtype=np.float64))
        std = np.ascontiguousarray(np.asarray(std, dtype=np.float64))
        capi.check(capi.lib.pbu_propagation_eval_gaussian(self._h, capi.dptr(mean), capi.dptr(std), int(seed), int(sample_offset)))

    def samples(self, first=0, count=None):
        """Rows [first, first + count) of the samples the evaluations belong to, [count, d] on the host: the rows handed
        to ``evaluate`` or drawn on the device by ``evaluate_gaussian``."""
        count = self.n_samples - first if count is None else count
        out = np.empty((count, self.engine.d))
        capi.check(capi.lib.pbu_propagation_get_samples(self._h, int(first), int(count), capi.dptr(out)))
        return out

    def evals(self, first=0, count=None):
        count = self.n_samples - first if count is None else count
        out = np.empty((count, self.n_points))
        capi.check(capi.lib.pbu_propagation_get_evals(self._h, int(first), int(count), capi.dptr(out)))
        return out

    # below this many samples per GPU the full radix select is cheap and the bracket of a subsample is loose
    BRACKET_MIN_SAMPLES = 1 << 18
    BRACKET_STRIDE = 64

    def statistics(self, percentiles=(2.5, 97.5), reduce=None, n_total=None, bracket=None):
        """mean, min, max and the percentiles per design point.

        reduce(array, op) sums ('sum') or takes the min/max ('min'/'max') of a numpy array over all ranks and returns
        it; None = single GPU.  n_total = number of samples over all ranks.  bracket: True / False forces / forbids the
        one-pass bracketed select (default: used from BRACKET_MIN_SAMPLES samples per GPU and for one or two
        percentiles); both paths return the same exact order statistics."""
        N = self.n_points
        n_total = self.n_samples if n_total is None else int(n_total)
        ident = (lambda a, op: a) if reduce is None else reduce
        ranks = []
        for q in percentiles:
            ranks.extend(percentile_ranks(n_total, q))
        r = np.ascontiguousarray(np.asarray(ranks, dtype=np.int64))
        T = len(ranks)
        nq = len(percentiles)
        if bracket is None:          # every rank must take the same branch
            bracket = bool(ident(np.array([float(self.n_samples >= self.BRACKET_MIN_SAMPLES)]), "min")[0] > 0.0)
        bracket = bracket and nq in (1, 2)
        s = mn = mx = None
        vals = None
        self.last_select = "radix"
        if bracket:
            s, mn, mx, vals = self._bracketed(percentiles, r, n_total, ident, reduce)
        if s is None:
            s, mn, mx = np.empty(N), np.empty(N), np.empty(N)
            capi.check(capi.lib.pbu_propagation_moments(self._h, capi.dptr(s), capi.dptr(mn), capi.dptr(mx)))
            s, mn, mx = ident(s, "sum"), ident(mn, "min"), ident(mx, "max")
        if vals is None:
            # the reduced min / max bound every key: the passes over the bytes they share are skipped
            first = C.c_int(0)
            gmn, gmx = np.ascontiguousarray(mn, dtype=np.float64), np.ascontiguousarray(mx, dtype=np.float64)
            capi.check(capi.lib.pbu_propagation_select_begin(self._h, r.ctypes.data_as(C.POINTER(C.c_int64)), T, capi.dptr(gmn), capi.dptr(gmx),
                                                             C.byref(first)))
            vals = self._select_passes(first.value, T, ident, reduce)
        out = dict(mean=s / n_total, min=mn, max=mx, n_samples=n_total)
        for i, q in enumerate(percentiles):
            out["p%g" % q] = np.array([lerp_percentile(n_total, q, vals[k, 2 * i], vals[k, 2 * i + 1]) for k in range(N)])
        if tuple(percentiles) == (2.5, 97.5):
            out["ci_lb"], out["ci_ub"] = out["p2.5"], out["p97.5"]
        return out

    def _select_passes(self, first, T, ident, reduce):
        """radix passes first..7 over whatever select_begin / bracket_select_begin armed; returns the values [N][T]"""
        N = self.n_points
        dev_sum = getattr(reduce, "device_sum", None)
        if reduce is None:          # one GPU: the histograms stay on the device, the passes are queued back to back
            for p in range(first, 8):
                capi.check(capi.lib.pbu_propagation_select_pass(self._h, p, None))
                capi.check(capi.lib.pbu_propagation_select_update(self._h, p, None))
        elif dev_sum is not None:   # NCCL: the histogram is all-reduced where it lies, on the engine's stream
            import torch
            hist = torch.empty((N, T, 256), dtype=torch.int64, device=torch.device("cuda", self.engine.device))
            hp = C.c_void_p(hist.data_ptr())
            for p in range(first, 8):
                capi.check(capi.lib.pbu_propagation_select_pass_dev(self._h, p, hp))
                dev_sum(hist)
                capi.check(capi.lib.pbu_propagation_select_update_dev(self._h, p, hp))
        else:                       # any other reducer: through host arrays
            hist = np.empty((N, T, 256), dtype=np.uint64)
            hp = hist.ctypes.data_as(C.POINTER(C.c_uint64))
            for p in range(first, 8):
                capi.check(capi.lib.pbu_propagation_select_pass(self._h, p, hp))
                tot = np.ascontiguousarray(ident(hist.view(np.int64), "sum")).view(np.uint64)
                capi.check(capi.lib.pbu_propagation_select_update(self._h, p, tot.ctypes.data_as(C.POINTER(C.c_uint64))))
        vals = np.empty((N, T))
        capi.check(capi.lib.pbu_propagation_select_result(self._h, capi.dptr(vals)))
        return vals

    def _bracketed(self, percentiles, ranks, n_total, ident, reduce):
        """One-pass path (include/pybitup_b200.h, 'Bracketed select').  Returns (sum, min, max, values); values is None
        when the brackets missed a wanted rank or the candidate buffer overflowed (the moments are still good).
        Three small reductions over the ranks: brackets (a, -b) by min; (sum, below, inside, overflow) by sum -- the
        counts are exact in float64 -- and (min, -max) by min."""
        N, nq, S = self.n_points, len(percentiles), self.n_samples
        stride = self.BRACKET_STRIDE
        fr = np.ascontiguousarray([q / 100.0 for q in percentiles], dtype=np.float64)
        brk = np.empty((N, nq, 2))
        capi.check(capi.lib.pbu_propagation_bracket_sample(self._h, capi.dptr(fr), nq, stride, capi.dptr(brk)))
        ab = ident(np.ascontiguousarray(np.stack([brk[..., 0], -brk[..., 1]])), "min")
        brk = np.ascontiguousarray(np.stack([ab[0], -ab[1]], axis=-1))
        m = max(S // stride, 1)
        hint = int(max(S * (2.0 * (6.0 * np.sqrt(m * f * (1.0 - f)) + 8.0) + 3.0) / m for f in fr))
        s, mn, mx = np.empty(N), np.empty(N), np.empty(N)
        below, inside = np.empty((N, nq), dtype=np.int64), np.empty((N, nq), dtype=np.int64)
        ovf = C.c_int(0)
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
        capi.check(capi.lib.pbu_propagation_bracket_pass(self._h, capi.dptr(brk), nq, hint, capi.dptr(s), capi.dptr(mn), capi.dptr(mx), ip(below), ip(inside),
                                                         C.byref(ovf)))
        tot = ident(np.concatenate([s, below.ravel().astype(np.float64), inside.ravel().astype(np.float64), [float(ovf.value)]]), "sum")
        s = np.ascontiguousarray(tot[:N])
        below = np.rint(tot[N:N + N * nq]).astype(np.int64).reshape(N, nq)
        inside = np.rint(tot[N + N * nq:N + 2 * N * nq]).astype(np.int64).reshape(N, nq)
        overflow = tot[-1] > 0.0
        ext = ident(np.concatenate([mn, -mx]), "min")
        mn, mx = np.ascontiguousarray(ext[:N]), np.ascontiguousarray(-ext[N:])
        rk = np.asarray(ranks, dtype=np.int64).reshape(nq, 2)
        inb = np.ascontiguousarray((rk[None, :, :] - below[:, :, None]).reshape(N, 2 * nq))
        ok = (not overflow) and bool(np.all(inb >= 0)) and bool(np.all(inb < np.repeat(inside, 2, axis=1)))
        if not ok:
            self.last_select = "radix (bracket missed)"
            return s, mn, mx, None
        first = C.c_int(0)
        capi.check(capi.lib.pbu_propagation_bracket_select_begin(self._h, ip(inb), capi.dptr(brk), C.byref(first)))
        self.last_select = "bracket"
        vals = self._select_passes(first.value, 2 * nq, ident, reduce)
        # a == b: the values between the brackets are all equal (they were counted, not collected)
        tie = np.repeat(brk[..., 0] == brk[..., 1], 2, axis=1)
        vals[tie] = np.repeat(brk[..., 0], 2, axis=1)[tie]
        return s, mn, mx, vals


def propagate(engine, X, percentiles=(2.5, 97.5), return_evals=False):
    """fun_eval statistics for samples X [S, d] on one GPU (the whole of solve_problem.py:543-647)."""
    with DevicePropagation(engine, len(X)) as dp:
        dp.evaluate(X)
        out = dp.statistics(percentiles)
        if return_evals:
            out["fun_eval"] = dp.evals()
    return out
