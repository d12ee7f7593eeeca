"""Many-chain diagnostics built on additive chain sums (K4, ``pbu_engine_chain_sums``).

The reference computes one chain's mean and covariance from its chain file at the end of the run
(``compute_covariance``, pybitup/metropolis_hastings_algorithms.py:295-307) and has no R-hat or ESS
(SURVEY.md §5).  With many chains the same quantities come from sums over chains that add up over GPUs.
"""
import numpy as np


def sums_layout(d):
    nt = d * (d + 1) // 2
    o = {"m": 0, "n": 1, "n_delta": slice(2, 2 + d), "delta": slice(2 + d, 2 + 2 * d),
         "delta2": slice(2 + 2 * d, 2 + 3 * d), "s2": slice(2 + 3 * d, 2 + 4 * d), "S": slice(2 + 4 * d, 2 + 4 * d + nt)}
    return o, 2 + 4 * d + nt


def chain_sums_numpy(count, mean, M2, mu):
    """The vector pbu_engine_chain_sums returns, from per-chain moments: count [C], mean [C, d], M2 [C, d, d]
    (sums of outer products of deviations).  Host restatement used by the CPU tests of the multi-rank path."""
    count = np.asarray(count, dtype=np.float64)
    C, d = mean.shape
    lay, L = sums_layout(d)
    out = np.zeros(L)
    act = count > 0
    dl = (mean - mu[None, :]) * act[:, None]
    out[0] = act.sum()
    out[1] = count.sum()
    out[lay["n_delta"]] = (count[:, None] * dl).sum(axis=0)
    out[lay["delta"]] = dl.sum(axis=0)
    out[lay["delta2"]] = (dl * dl).sum(axis=0)
    multi = count > 1
    s2 = np.where(multi[:, None], np.einsum("cii->ci", M2) / np.maximum(count - 1.0, 1.0)[:, None], 0.0)
    out[lay["s2"]] = s2.sum(axis=0)
    full = (M2 * act[:, None, None]).sum(axis=0) + np.einsum("c,ci,cj->ij", count, dl, dl)
    iu = np.triu_indices(d)
    out[lay["S"]] = full[iu]
    return out


def finalize(sums, mu, d):
    """Pooled mean / covariance and Gelman-Rubin R-hat from (all-reduced) chain sums taken about mu."""
    lay, _ = sums_layout(d)
    m, n = sums[0], sums[1]
    mean = mu + sums[lay["n_delta"]] / n
    shift = mean - mu
    S = np.zeros((d, d))
    S[np.triu_indices(d)] = sums[lay["S"]]
    S = S + np.triu(S, 1).T
    M2 = S - n * np.outer(shift, shift)                 # sum of outer products of deviations about the pooled mean
    cov = M2 / (n - 1.0)
    # Gelman-Rubin (chains of equal length n_c = n/m): W = mean within-chain variance, B/n_c = variance of chain means
    n_c = n / m
    W = sums[lay["s2"]] / m
    cm = sums[lay["delta"]] / m                         # mean of chain means - mu
    B_over_n = (sums[lay["delta2"]] - m * cm * cm) / (m - 1.0) if m > 1 else np.full(d, np.nan)
    var_plus = (n_c - 1.0) / n_c * W + B_over_n
    with np.errstate(divide="ignore", invalid="ignore"):
        rhat = np.sqrt(var_plus / W)
    return dict(n_chains=int(round(m)), n=int(round(n)), mean=mean, cov=cov, W=W, B_over_n=B_over_n, rhat=rhat)


def ess_batch_means(chain):
    """Effective sample size of one chain [T, d] by non-overlapping batch means (batch length ~ sqrt(T)),
    minimum over parameters.  The same estimator is applied to GPU and CPU chains in the benchmark."""
    chain = np.asarray(chain, dtype=np.float64)
    T, d = chain.shape
    b = max(1, int(np.floor(np.sqrt(T))))
    a = T // b
    if a < 2:
        return float("nan")
    x = chain[: a * b].reshape(a, b, d)
    bm = x.mean(axis=1)
    var_bm = bm.var(axis=0, ddof=1) * b
    var = chain[: a * b].var(axis=0, ddof=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        ess = a * b * var / var_bm
    return float(np.nanmin(ess))
