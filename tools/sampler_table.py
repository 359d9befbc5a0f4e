"""The reference's sampler comparison (examples/logistic_regression/src/table.jl:9-17: acceptance rate, ESS, time and efficiency of
MALA / AM / SMMALA / GAMC, efficiency relative to MALA) on the C2-shaped synthetic problem, every sampler on the same kernels:

  python tools/sampler_table.py [--logN 20 --p 256 --steps 3000]

MALA = Klara's comparison sampler (`sampler_kind = 1`); "AM" and "SMMALA" are GAMC with `am_only_update!` / `smmala_only_update!`
(after the t0 = 3 SMMALA start of GAMC; Klara's own AM seeds its covariance differently)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gamc_b200 as g  # noqa: E402

SEED = 20260101


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logN", type=int, default=20)
    ap.add_argument("--p", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3000)
    a = ap.parse_args()
    N, p, K = 1 << a.logN, a.p, a.steps
    tt = g.synth_theta_true(SEED, p)
    theta0 = tt + 0.01 * np.random.Generator(np.random.PCG64(SEED + 1)).standard_normal(p)
    eng = g.Engine(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    eng.target_logistic_synth(SEED, 0, N, p, tt, 100.0)
    model = g.LogisticModel(synth={"seed": SEED, "N": N, "p": p, "theta_true": tt}, lam=100.0)
    gamc_tuner = g.GAMCTuner(g.VanillaMCTuner(), g.VanillaMCTuner(), g.AcceptanceRateMCTuner(0.35))
    cases = [
        ("MALA", g.MALA(0.02), g.AcceptanceRateMCTuner(0.574, score_k=3.0)),                      # MALA/simulations.jl:37-41
        ("AM", g.GAMC(update=g.am_only_update(), driftstep=0.02, minorscale=0.001, c=0.01, am_mode=2), gamc_tuner),
        ("SMMALA", g.GAMC(update=g.smmala_only_update(), driftstep=0.02, minorscale=0.001, c=0.01, am_mode=2), gamc_tuner),
        ("GAMC", g.GAMC(update=g.rand_exp_decay_update(10.0), driftstep=0.02, minorscale=0.001, c=0.01, am_mode=2), gamc_tuner),
    ]
    rows = []
    for name, sampler, tuner in cases:
        for nsteps in (64, K):                       # a short warm-up job, then the measured one
            job = g.BasicMCJob(model, sampler, g.BasicMCRange(nsteps, nsteps // 5), {"p": theta0}, tuner=tuner, engine=eng,
                               tape=g.RandomTape(nsteps, p, seed=SEED + 9))
            t = job.tape
            eng.tape_set(t.u_sched, t.u_mix, t.z, t.u_acc)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            eng.run(nsteps)
            e1.record(stream)
            torch.cuda.synchronize()
            job._done = nsteps
        sec = e0.elapsed_time(e1) * 1e-3
        chain = job.output()
        ess = g.ess(chain.value)
        rows.append({"sampler": name, "iters_per_sec": K / sec, "acceptance": g.acceptance(chain), "ess_min": float(ess.min()),
                     "ess_mean": float(ess.mean()), "seconds": sec, "efficiency": float(ess.min() / sec)})
    base = rows[0]["efficiency"]
    for r in rows:
        r["efficiency_vs_mala"] = r["efficiency"] / base if base > 0 else None
        print(json.dumps(r))
    eng.close()


if __name__ == "__main__":
    main()
