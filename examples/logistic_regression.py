"""The reference's logistic-regression experiment (examples/logistic_regression/src/GAMC/simulations.jl + GAMC/table.jl)
through the B200 path: `nchains` GAMC chains of `nmcmc` iterations (burn-in `nburnin`), a chain is kept only if its
acceptance rate lies in (0.22, 0.37) (:80), chains / diagnostics / times / step sizes / update counts are written in the
reference's CSV formats and summarised as in table.jl.

The Swiss bank-note data of the reference lives inside Klara.jl (`dataset("swiss", ...)`, :17,22) and is not available
here; a synthetic stand-in of the same shape (N = 200, p = 4, standardised columns) is used unless --data X.csv y.csv is given.

  python examples/logistic_regression.py --outdir /tmp/gamc_out [--nchains 10 --nmcmc 110000 --nburnin 10000]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gamc_b200 as g  # noqa: E402


def standardised_synthetic(N=200, p=4, seed=1):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, p))
    X = (X - X.mean(0)) / X.std(0, ddof=1)                       # (covariates .- mean) ./ std  (:20)
    beta = np.array([-1.2, 1.5, 0.8, -0.6])[:p]
    y = (rng.random(N) < 1.0 / (1.0 + np.exp(-(X @ beta)))).astype(np.float64)
    return X, y


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--outdir", default="/tmp/gamc_b200_logistic/GAMC")
    ap.add_argument("--nchains", type=int, default=10)
    ap.add_argument("--nmcmc", type=int, default=110000)
    ap.add_argument("--nburnin", type=int, default=10000)
    ap.add_argument("--data", nargs=2, default=None)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    if a.data:
        X, y = np.loadtxt(a.data[0], delimiter=",", ndmin=2), np.loadtxt(a.data[1], delimiter=",")
    else:
        X, y = standardised_synthetic()
    npars = X.shape[1]
    eng = g.Engine(a.device)
    eng.target_logistic(X, y, 100.0)                              # λ = 100 (:69)
    sampler = g.GAMC(update=g.rand_exp_decay_update(10.0), driftstep=0.02, minorscale=0.001, c=0.01, am_mode=1)          # :50-55
    mcrange = g.BasicMCRange(nsteps=a.nmcmc, burnin=a.nburnin)                                                            # :57
    mctuner = g.GAMCTuner(g.VanillaMCTuner(verbose=False), g.VanillaMCTuner(verbose=False), g.AcceptanceRateMCTuner(0.35, verbose=False))   # :59
    rng = np.random.default_rng(0)
    times, stepsizes, nupdates = [], [], []
    i, attempts = 1, 0
    while i <= a.nchains and attempts < 10 * a.nchains:
        attempts += 1
        v0 = {"p": 3.0 * rng.standard_normal(npars)}             # rand(Normal(0, 3), npars)  (:69)
        job = g.BasicMCJob(g.LogisticModel(X, y, 100.0), sampler, mcrange, v0, tuner=mctuner, engine=eng, tape_seed=1000 + attempts)
        t0 = time.perf_counter()
        g.run(job)
        runtime = time.perf_counter() - t0
        chain = g.output(job)
        ratio = g.acceptance(chain)
        if 0.22 < ratio < 0.37:                                   # :80
            g.write_chain_csv(a.outdir, i, chain.value, chain.diagnosticvalues)
            times.append(runtime)
            stepsizes.append(job.sstate.tune.totaltune.step)
            nupdates.append(job.sstate.updatetensorcount)
            print(f"Iteration {i} of {a.nchains} completed with acceptance ratio {ratio:.4f} in {runtime:.2f} s")
            i += 1
        else:
            print(f"  chain rejected: acceptance ratio {ratio:.4f} outside (0.22, 0.37)")
    g.write_run_csvs(a.outdir, times, stepsizes, nupdates)
    res = g.table(a.outdir, len(times))
    print("rate", res["rate"], "ess", res["ess"], "time", res["time"], "efficiency (min ESS / s)", res["efficiency"])
    eng.close()


if __name__ == "__main__":
    main()
