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

    def evals(self, first=0, count=None):
        count = self.n_samples - first if count is None else count
        out = np.empty((count, self.n_points))
        capi.check(capi.lib.pbu_propagation_get_evals(self._h, int(first), int(count), capi.dptr(out)))
        return out

    def statistics(self, percentiles=(2.5, 97.5), reduce=None, n_total=None):
        """mean, min, max and the percentiles per design point.

        reduce(array, op) sums ('sum') or takes the min/max ('min'/'max') of a numpy array over all ranks and returns
        it; None = single GPU.  n_total = number of samples over all ranks."""
        N = self.n_points
        n_total = self.n_samples if n_total is None else int(n_total)
        ident = (lambda a, op: a) if reduce is None else reduce
        s, mn, mx = np.empty(N), np.empty(N), np.empty(N)
        capi.check(capi.lib.pbu_propagation_moments(self._h, capi.dptr(s), capi.dptr(mn), capi.dptr(mx)))
        s, mn, mx = ident(s, "sum"), ident(mn, "min"), ident(mx, "max")
        ranks = []
        for q in percentiles:
            ranks.extend(percentile_ranks(n_total, q))
        r = np.ascontiguousarray(np.asarray(ranks, dtype=np.int64))
        T = len(ranks)
        # the reduced min / max bound every key: the passes over the bytes they share are skipped
        first = C.c_int(0)
        gmn, gmx = np.ascontiguousarray(mn, dtype=np.float64), np.ascontiguousarray(mx, dtype=np.float64)
        capi.check(capi.lib.pbu_propagation_select_begin(self._h, r.ctypes.data_as(C.POINTER(C.c_int64)), T, capi.dptr(gmn), capi.dptr(gmx),
                                                         C.byref(first)))
        hist = np.empty((N, T, 256), dtype=np.uint64)
        hp = hist.ctypes.data_as(C.POINTER(C.c_uint64))
        for p in range(first.value, 8):
            if reduce is None:          # one GPU: the histograms stay on the device, the passes are queued back to back
                capi.check(capi.lib.pbu_propagation_select_pass(self._h, p, None))
                capi.check(capi.lib.pbu_propagation_select_update(self._h, p, None))
                continue
            capi.check(capi.lib.pbu_propagation_select_pass(self._h, p, hp))
            tot = np.ascontiguousarray(ident(hist.view(np.int64), "sum")).view(np.uint64)
            capi.check(capi.lib.pbu_propagation_select_update(self._h, p, tot.ctypes.data_as(C.POINTER(C.c_uint64))))
        vals = np.empty((N, T))
        capi.check(capi.lib.pbu_propagation_select_result(self._h, capi.dptr(vals)))
        out = dict(mean=s / n_total, min=mn, max=mx, n_samples=n_total)
        for i, q in enumerate(percentiles):
            out["p%g" % q] = np.array([lerp_percentile(n_total, q, vals[k, 2 * i], vals[k, 2 * i + 1]) for k in range(N)])
        if tuple(percentiles) == (2.5, 97.5):
            out["ci_lb"], out["ci_ub"] = out["p2.5"], out["p97.5"]
        return out


def propagate(engine, X, percentiles=(2.5, 97.5), return_evals=False):
    """fun_eval statistics for samples X [S, d] on one GPU (the whole of solve_problem.py:543-647)."""
    with DevicePropagation(engine, len(X)) as dp:
        dp.evaluate(X)
        out = dp.statistics(percentiles)
        if return_evals:
            out["fun_eval"] = dp.evals()
    return out
