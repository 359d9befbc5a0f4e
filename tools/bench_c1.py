"""BASELINE.json configs[0] ("C1"): the logistic-regression example as shipped -- 10 GAMC chains x 110 000 iterations (10 000
burn-in), N = 200 observations, p = 4, prior N(0, 100 I), init N(0, 3^2), acceptance window (0.22, 0.37)
(examples/logistic_regression/src/GAMC/simulations.jl:13-15,50-95; efficiency = min-coordinate mean ESS / mean time,
GAMC/table.jl:21-31).  The Swiss bank-note data lives inside Klara.jl and is not available: a synthetic stand-in of the same shape.

Three arms on the same data: (i) the batched persistent kernel (all chains at once, one launch), (ii) the single-chain path
(one chain, >= 3 launches per iteration: launch-bound at this size), (iii) the CPU oracle (NumPy restatement of the reference).
  python tools/bench_c1.py [--nchains 10 --nmcmc 110000 --nburnin 10000 --cpu-chains 1]
Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
import gamc_b200 as g  # noqa: E402
from logistic_regression import standardised_synthetic  # noqa: E402


def summarise(chains, seconds_per_chain):
    rates = [float(c.diagnosticvalues.mean()) for c in chains]
    ess = np.array([g.ess(c.value) for c in chains])
    keep = [i for i, r in enumerate(rates) if 0.22 < r < 0.37]
    mean_ess = ess[keep].mean(axis=0) if keep else ess.mean(axis=0)
    return {"acceptance_mean": float(np.mean(rates)), "chains_in_window": len(keep), "chains": len(chains),
            "mean_ess_per_coordinate": [float(x) for x in mean_ess], "min_ess": float(mean_ess.min()),
            "seconds_per_chain": seconds_per_chain, "efficiency_min_ess_per_sec": float(mean_ess.min() / seconds_per_chain)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nchains", type=int, default=10)
    ap.add_argument("--nmcmc", type=int, default=110000)
    ap.add_argument("--nburnin", type=int, default=10000)
    ap.add_argument("--cpu-chains", type=int, default=1)
    ap.add_argument("--many", type=int, default=4096, help="second batched run with this many chains (0 = skip): what one GPU does in the same time")
    a = ap.parse_args()
    import torch
    X, y = standardised_synthetic()
    p = X.shape[1]
    rng = np.random.default_rng(0)
    v0s = 3.0 * rng.standard_normal((max(a.nchains, a.many), p))          # rand(Normal(0, 3), npars)
    sampler = g.GAMC(update=g.rand_exp_decay_update(10.0), driftstep=0.02, minorscale=0.001, c=0.01, am_mode=1)
    tuner = g.GAMCTuner(g.VanillaMCTuner(), g.VanillaMCTuner(), g.AcceptanceRateMCTuner(0.35))
    mcrange = g.BasicMCRange(a.nmcmc, a.nburnin)
    model = g.LogisticModel(X, y, 100.0)
    out = {"workload": f"C1: logistic example as shipped, N=200, p=4, {a.nchains} chains x {a.nmcmc} iterations ({a.nburnin} burn-in), synthetic stand-in for the Swiss data"}

    # (i) batched persistent kernel
    def batched(nch):
        job = g.BatchedMCJob(model, sampler, g.BasicMCRange(200, 0), v0s[:nch], tuner=tuner, seed=1)
        job.run()                                                           # warm-up
        job.engine.close()
        job = g.BatchedMCJob(model, sampler, mcrange, v0s[:nch], tuner=tuner, seed=7)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        job.run()
        sec = time.perf_counter() - t0
        return job, sec
    job, sec = batched(a.nchains)
    chains = job.output()
    out["gpu_batched"] = {"seconds_total": sec, "chain_iters_per_sec": a.nchains * a.nmcmc / sec, **summarise(chains, sec / a.nchains),
                          "seconds_per_chain_if_run_alone": sec,
                          "note": "all chains advance concurrently in one kernel launch (one CTA per chain); seconds_per_chain = total / nchains"}
    job.engine.close()
    if a.many:
        nsteps_many = min(a.nmcmc, 20000)
        jobm = g.BatchedMCJob(model, sampler, g.BasicMCRange(nsteps_many, nsteps_many - 10), v0s[:a.many], tuner=tuner, seed=9)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        jobm.run()
        secm = time.perf_counter() - t0
        out["gpu_batched_many"] = {"chains": a.many, "nsteps": nsteps_many, "seconds_total": secm, "chain_iters_per_sec": a.many * nsteps_many / secm}
        jobm.engine.close()

    # (ii) single-chain path, one chain (launch-bound)
    eng = g.Engine(0)
    eng.target_logistic(X, y, 100.0)
    job1 = g.BasicMCJob(model, sampler, mcrange, {"p": v0s[0]}, tuner=tuner, engine=eng, tape_seed=11)
    t0 = time.perf_counter()
    job1.run()
    sec1 = time.perf_counter() - t0
    out["gpu_single_chain_path"] = {"seconds_per_chain": sec1, "iters_per_sec": a.nmcmc / sec1, **{k: v for k, v in summarise([job1.output()], sec1).items()
                                                                                              if k in ("acceptance_mean", "min_ess", "efficiency_min_ess_per_sec")}}
    eng.close()

    # (iii) CPU oracle (the reference's algorithm on the host)
    import oracle as o
    ochains, osecs = [], []
    for c in range(a.cpu_chains):
        osam = o.GAMCOracle(o.LogisticTarget(X, y, 100.0), v0s[c], a.nmcmc, a.nburnin, update="rand_exp_decay_update!", update_params=(10.0, 0.0),
                            driftstep=0.02, minorscale=0.001, c=0.01, tuner=o.GAMCTuner(o.VanillaMCTuner(), o.VanillaMCTuner(), o.AcceptanceRateMCTuner(0.35)))
        t0 = time.perf_counter()
        v, acc, _ = osam.run(o.RandomTape(a.nmcmc, p, seed=100 + c))
        osecs.append(time.perf_counter() - t0)
        ochains.append(g.Chain(v, acc.reshape(1, -1)))
    out["cpu_oracle"] = {"cores": 1, "kind": "port", "iters_per_sec": a.nmcmc / float(np.mean(osecs)), **summarise(ochains, float(np.mean(osecs)))}
    out["speedup_efficiency_batched_vs_cpu"] = out["gpu_batched"]["efficiency_min_ess_per_sec"] / out["cpu_oracle"]["efficiency_min_ess_per_sec"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
