"""Posterior diagnostics and summaries with PyMC3 3.6 semantics [recalled, SURVEY §3.4/§8a].

Used by the host layer to write the summary CSV (ref: baghera_tool/gene_regression.py:120-127) and
the Gelman-Rubin log line (ref: baghera_tool/gene_regression.py:108-116).  Plain numpy: these run
once per invocation on a handful of scalar traces.
"""
from __future__ import annotations

import numpy as np


def gelman_rubin(x):
    """pymc3.diagnostics.gelman_rubin for one scalar variable. x: [n_draws, n_chains].
    B = n var(chain means, ddof=1); W = mean(chain vars, ddof=1); Rhat = sqrt(((n-1)/n W + B/n) / W)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    if x.shape[1] < 2:
        raise ValueError("Gelman-Rubin diagnostic requires multiple chains of the same length.")
    B = n * np.var(x.mean(axis=0), ddof=1)
    W = np.mean(np.var(x, axis=0, ddof=1))
    vhat = W * (n - 1) / n + B / n
    return float(np.sqrt(vhat / W))


def _autocov(x):
    """Autocovariance by FFT (biased, divided by n), like pymc3.stats.autocov."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    xc = x - x.mean()
    f = np.fft.rfft(xc, 2 * n)
    return np.fft.irfft(f * np.conjugate(f))[:n] / n


def effective_n(x):
    """pymc3.diagnostics.effective_n (3.6: Stan's estimator with Geyer's initial positive and monotone
    sequences) for one scalar variable. x: [n_draws, n_chains]."""
    x = np.asarray(x, dtype=np.float64)
    n_samples, nchain = x.shape
    if nchain < 2:
        raise ValueError("Calculation of effective sample size requires multiple chains of the same length.")
    acov = np.asarray([_autocov(x[:, c]) for c in range(nchain)])
    chain_mean = x.mean(axis=0)
    chain_var = acov[:, 0] * n_samples / (n_samples - 1.0)
    acov_t = acov[:, 1] * n_samples / (n_samples - 1.0)
    mean_var = np.mean(chain_var)
    var_plus = mean_var * (n_samples - 1.0) / n_samples
    var_plus += np.var(chain_mean, ddof=1)
    if not var_plus > 0:
        return float(nchain * n_samples)
    rho_hat_t = np.zeros(n_samples)
    rho_hat_even = 1.0
    rho_hat_t[0] = rho_hat_even
    rho_hat_odd = 1.0 - (mean_var - np.mean(acov_t)) / var_plus
    rho_hat_t[1] = rho_hat_odd
    max_t = 1
    t = 1
    while t < (n_samples - 2) and (rho_hat_even + rho_hat_odd) >= 0.0:
        rho_hat_even = 1.0 - (mean_var - np.mean(acov[:, t + 1])) / var_plus
        rho_hat_odd = 1.0 - (mean_var - np.mean(acov[:, t + 2])) / var_plus
        if (rho_hat_even + rho_hat_odd) >= 0:
            rho_hat_t[t + 1] = rho_hat_even
            rho_hat_t[t + 2] = rho_hat_odd
        max_t = t + 2
        t += 2
    t = 3
    while t <= max_t - 2:
        if (rho_hat_t[t + 1] + rho_hat_t[t + 2]) > (rho_hat_t[t - 1] + rho_hat_t[t]):
            rho_hat_t[t + 1] = (rho_hat_t[t - 1] + rho_hat_t[t]) / 2.0
            rho_hat_t[t + 2] = rho_hat_t[t + 1]
        t += 2
    ess = nchain * n_samples
    return float(ess / (-1.0 + 2.0 * np.sum(rho_hat_t)))


def mc_error(x, batches=100):
    """pymc3.stats.mc_error on the concatenated chains: std of batch means / sqrt(batches)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    if batches == 1 or x.shape[0] < batches:
        return float(np.std(x) / np.sqrt(x.shape[0]))
    keep = x.shape[0] - x.shape[0] % batches
    means = x[:keep].reshape(batches, -1).mean(axis=1)
    return float(np.std(means) / np.sqrt(batches))


def hpd(x, alpha=0.05):
    """pymc3.stats.hpd: the narrowest interval containing (1 - alpha) of the sorted draws."""
    x = np.sort(np.asarray(x, dtype=np.float64).ravel())
    n = x.shape[0]
    interval_idx_inc = int(np.floor((1.0 - alpha) * n))
    n_intervals = n - interval_idx_inc
    width = x[interval_idx_inc:] - x[:n_intervals]
    if len(width) == 0:
        raise ValueError("Too few elements for interval calculation")
    k = int(np.argmin(width))
    return float(x[k]), float(x[k + interval_idx_inc])


def summary(traces, extend_median_quantiles=True):
    """pm.summary(trace, varnames, extend=True, stat_funcs=[trace_median, trace_quantiles]) as the
    reference calls it (ref: baghera_tool/gene_regression.py:120-125, stat funcs :14-23).

    traces: ordered dict name -> array [n_draws, n_chains] (scalar variables).
    Columns: mean, sd, mc_error, hpd_2.5, hpd_97.5, median, 5, 50, 95, n_eff, Rhat (the last two only
    with more than one chain)."""
    import pandas as pd

    rows = []
    for name, x in traces.items():
        x = np.asarray(x, dtype=np.float64)
        flat = x.ravel(order="F")                      # chains concatenated, like trace.get_values(combine=True)
        lo, hi = hpd(flat, 0.05)
        row = {"mean": flat.mean(), "sd": flat.std(), "mc_error": mc_error(flat), "hpd_2.5": lo, "hpd_97.5": hi}
        if extend_median_quantiles:
            row["median"] = np.median(flat)
            q = np.percentile(flat, [5, 50, 95])
            row["5"], row["50"], row["95"] = q[0], q[1], q[2]
        if x.shape[1] > 1:
            row["n_eff"] = effective_n(x)
            row["Rhat"] = gelman_rubin(x)
        rows.append(pd.Series(row, name=name))
    return pd.DataFrame(rows)
