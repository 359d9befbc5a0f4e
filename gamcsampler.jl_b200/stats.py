"""Chain summaries on the host (product side; independent of oracle/): the `table.jl` pipeline of the
reference (examples/logistic_regression/src/GAMC/table.jl:21-31): acceptance rate, per-coordinate
IMSE effective sample size (Klara `ess`), efficiency = min ESS / time.  FFT autocovariances."""
from __future__ import annotations

import numpy as np


def _autocov_fft(v, maxlag):
    n = v.shape[0]
    z = v - v.mean()
    m = 1 << int(np.ceil(np.log2(2 * n)))
    f = np.fft.rfft(z, m)
    ac = np.fft.irfft(f * np.conj(f), m)[: maxlag + 1]
    return ac / n


def mcvar_imse(v, maxlag=None):
    """Geyer's initial monotone sequence estimator of the Monte Carlo variance of the mean."""
    v = np.asarray(v, dtype=np.float64)
    n = v.shape[0]
    if maxlag is None:
        maxlag = n - 1
    k = (maxlag - 1) // 2
    acv = _autocov_fft(v, 2 * k + 1)
    g = acv[0:2 * k + 2:2] + acv[1:2 * k + 2:2]
    m = k + 1
    nonpos = np.nonzero(g[1:] <= 0)[0]
    if nonpos.size:
        m = int(nonpos[0]) + 1
    g = np.minimum.accumulate(g[:m])
    return (-acv[0] + 2.0 * float(np.sum(g))) / n


def ess(x, maxlag=None):
    """Klara `ess(x, 2)`: per-coordinate n * (var/n) / mcvar_imse for a p x nsamples chain."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[None, :]
    n = x.shape[1]
    return np.array([n * (np.var(v, ddof=1) / n) / mcvar_imse(v, maxlag) for v in x])


def summary(chain_value, accept, runtime_s):
    """results[:rate], [:ess], [:time], [:efficiency] of table.jl:28-31 for one chain."""
    e = ess(chain_value)
    return {"rate": float(np.mean(accept)), "ess": e, "time": float(runtime_s), "efficiency": float(e.min() / runtime_s)}
