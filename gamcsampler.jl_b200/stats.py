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


# ---- the examples' on-disk formats (examples/logistic_regression/src/GAMC/simulations.jl:81-95, GAMC/table.jl:33-53) ----
def write_chain_csv(outdir, i, chain_value, diagnosticvalues):
    """`chainNN.csv` (p x nsamples, comma separated, as Julia's writedlm of chain.value) and `diagnosticsNN.csv`
    (one true/false per line, writedlm of vec(chain.diagnosticvalues))."""
    import os
    os.makedirs(outdir, exist_ok=True)
    np.savetxt(os.path.join(outdir, f"chain{i:02d}.csv"), np.asarray(chain_value), delimiter=",", fmt="%.17g")
    with open(os.path.join(outdir, f"diagnostics{i:02d}.csv"), "w") as f:
        for a in np.asarray(diagnosticvalues, dtype=bool).ravel():
            f.write("true\n" if a else "false\n")


def write_run_csvs(outdir, times, stepsizes, nupdates):
    """`times.csv`, `stepsizes.csv`, `nupdates.csv` (simulations.jl:93-95)."""
    import os
    os.makedirs(outdir, exist_ok=True)
    np.savetxt(os.path.join(outdir, "times.csv"), np.asarray(times, dtype=np.float64), fmt="%.17g")
    np.savetxt(os.path.join(outdir, "stepsizes.csv"), np.asarray(stepsizes, dtype=np.float64), fmt="%.17g")
    np.savetxt(os.path.join(outdir, "nupdates.csv"), np.asarray(nupdates, dtype=np.int64), fmt="%d")


def table(outdir, nchains):
    """GAMC/table.jl:21-53: read the chains back, mean acceptance, mean ESS per coordinate, mean time,
    efficiency = min mean-ESS / mean time; writes summary.csv and summary.txt and returns the results dict."""
    import os
    ratio, essizes = [], []
    for i in range(1, nchains + 1):
        with open(os.path.join(outdir, f"diagnostics{i:02d}.csv")) as f:
            ratio.append(np.mean([ln.strip() == "true" for ln in f if ln.strip()]))
        essizes.append(ess(np.loadtxt(os.path.join(outdir, f"chain{i:02d}.csv"), delimiter=",", ndmin=2)))
    res = {"rate": float(np.mean(ratio)), "ess": np.mean(np.array(essizes), axis=0), "time": float(np.mean(np.loadtxt(os.path.join(outdir, "times.csv"), ndmin=1)))}
    res["efficiency"] = float(res["ess"].min() / res["time"])
    np.savetxt(os.path.join(outdir, "summary.csv"), np.concatenate([[res["rate"]], res["ess"], [res["time"], res["efficiency"]]])[None, :], delimiter=",", fmt="%.17g")
    with open(os.path.join(outdir, "summary.txt"), "w") as f:
        f.write(" & ".join([f"{res['rate']:.2f}"] + [str(int(round(e))) for e in res["ess"]] + [f"{res['time']:.2f}", f"{res['efficiency']:.2f}"]) + "\n")
    return res
