"""Oracle: chain summaries used by the reference's `table.jl` (TEST INFRASTRUCTURE).

`examples/logistic_regression/src/GAMC/table.jl:21-31`: per chain `acceptance(diagnostics)` and
`ess(chain, 2)` (per coordinate); efficiency = min over coordinates of the mean ESS / mean runtime.
Klara's `ess` is the IMSE-based effective sample size (Geyer's initial monotone sequence estimator);
restated here from its published algorithm (Geyer 1992), direct O(n*maxlag) autocovariances.
"""
from __future__ import annotations

import numpy as np

__all__ = ["acceptance", "autocov", "mcvar_imse", "ess"]


def acceptance(diagnostics):
    """Klara `acceptance`: mean of the `:accept` diagnostic over saved iterations (table.jl:22)."""
    d = np.asarray(diagnostics, dtype=bool).ravel()
    return float(np.mean(d))


def autocov(v, maxlag):
    """Biased (1/n) autocovariances of the demeaned series at lags 0..maxlag (StatsBase `autocov`)."""
    v = np.asarray(v, dtype=np.float64)
    n = v.shape[0]
    z = v - v.mean()
    out = np.empty(maxlag + 1)
    for k in range(maxlag + 1):
        out[k] = float(np.dot(z[: n - k], z[k:])) / n
    return out


def mcvar_imse(v, maxlag=None):
    """Monte Carlo variance of the mean by the initial monotone sequence estimator.

    Gamma_j = gamma_{2j} + gamma_{2j+1}; truncate at the first non-positive Gamma_j (j >= 1); enforce
    non-increasing; var = (-gamma_0 + 2 sum_j Gamma_j)/n.
    """
    v = np.asarray(v, dtype=np.float64)
    n = v.shape[0]
    if maxlag is None:
        maxlag = n - 1
    k = (maxlag - 1) // 2
    acv = autocov(v, 2 * k + 1)
    g = acv[0:2 * k + 2:2] + acv[1:2 * k + 2:2]
    m = k + 1
    for j in range(1, k + 1):
        if g[j] <= 0:
            m = j
            break
    g = g[:m].copy()
    for j in range(1, m):
        if g[j] > g[j - 1]:
            g[j] = g[j - 1]
    return (-acv[0] + 2.0 * float(np.sum(g))) / n


def ess(x, maxlag=None):
    """Klara `ess(x::Matrix, 2)`: per row (coordinate) n * mcvar_iid / mcvar_imse with mcvar_iid = var(v)/n
    (unbiased variance).  x is p x nsamples (the layout of `chain.value`) or a 1-D series."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[None, :]
    n = x.shape[1]
    if maxlag is None:
        maxlag = min(n - 1, 4000)  # direct autocov is O(n*maxlag); chains here decorrelate long before
    out = np.empty(x.shape[0])
    for i in range(x.shape[0]):
        v = x[i]
        out[i] = n * (np.var(v, ddof=1) / n) / mcvar_imse(v, maxlag)
    return out
