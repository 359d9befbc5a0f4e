"""The reference's radial-velocity experiment (examples/radial_velocity/src/{one_planet,two_planets}/GAMC/simulations.jl +
data_generation.jl) through the batched-chains kernel: simulated Keplerian observations, `nchains` independent GAMC chains
with `transform = H -> softabs(H, 1000.)` run at once in ONE persistent launch (the reference runs them one after the other),
a chain is kept if its acceptance rate lies in (0.22, 0.37), outputs in the reference's CSV formats.

  python examples/radial_velocity.py --outdir /tmp/gamc_rv [--planets 1 --nchains 64 --nmcmc 20000 --nburnin 2000 --num-obs 50]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gamc_b200 as g  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--outdir", default="/tmp/gamc_b200_rv/GAMC")
    ap.add_argument("--planets", type=int, default=1, choices=[1, 2])
    ap.add_argument("--nchains", type=int, default=64)
    ap.add_argument("--nmcmc", type=int, default=20000)
    ap.add_argument("--nburnin", type=int, default=2000)
    ap.add_argument("--num-obs", type=int, default=50)
    ap.add_argument("--data", default=None, help="time,obs,sigma CSV (e.g. the reference's data/one_planet.csv) instead of simulated data")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    os.makedirs(a.outdir, exist_ok=True)

    param_true = g.make_param_true_ex1() if a.planets == 1 else g.make_param_true_ex2()          # utils_ex.jl:48-68
    model = g.RVModel.from_csv(a.data, a.planets) if a.data else g.RVModel.simulate(param_true, num_obs=a.num_obs)   # data_generation.jl:13-28
    model.to_csv(os.path.join(a.outdir, "data.csv"))
    # param_perturb_scale of utils_ex.jl:23-46: per-parameter scales shrunk by sqrt(num_obs / 50)
    scale1 = np.array([0.0001, 0.01, 0.05, 0.05, np.pi * 0.01])
    scale = np.concatenate([np.tile(scale1, a.planets), [0.3]]) / np.sqrt(len(model.times) / 50.0)
    rng = np.random.default_rng(1)
    v0s = param_true + 0.01 * scale * rng.standard_normal((a.nchains, len(param_true)))           # simulations.jl:63

    sampler = g.GAMC(update=g.rand_exp_decay_update(10.0), transform=g.softabs(1000.0), driftstep=0.02, minorscale=0.001, c=0.01)
    tuner = g.GAMCTuner(g.VanillaMCTuner(), g.VanillaMCTuner(), g.AcceptanceRateMCTuner(0.35))
    t0 = time.perf_counter()
    job = g.BatchedMCJob(model, sampler, g.BasicMCRange(a.nmcmc, a.nburnin), v0s, tuner=tuner, device=a.device, seed=1)
    job.run()
    chains = job.output()
    runtime = time.perf_counter() - t0

    kept, times, steps, nupd = 0, [], [], []
    for c, chain in enumerate(chains):
        ratio = g.acceptance(chain)
        if 0.22 < ratio < 0.37:                                                                    # simulations.jl:75
            kept += 1
            g.write_chain_csv(a.outdir, kept, chain.value, chain.diagnosticvalues)
            st, _ = job.engine.batched_state(c)
            times.append(runtime / a.nchains)
            steps.append(st.step)
            nupd.append(st.updatetensorcount)
    g.write_run_csvs(a.outdir, times, steps, nupd)
    post_mean = np.mean([ch.value.mean(axis=1) for ch in chains], axis=0)
    print(f"{a.nchains} chains x {a.nmcmc} iterations in {runtime:.2f} s ({a.nchains * a.nmcmc / runtime:.0f} chain-iterations/s); "
          f"{kept} chains with acceptance in (0.22, 0.37)")
    print("posterior mean (all chains):", np.array2string(post_mean, precision=4))
    print("truth                      :", np.array2string(param_true, precision=4))
    if kept:
        res = g.table(a.outdir, kept)
        print("rate", res["rate"], "min ESS", np.min(res["ess"]), "efficiency (min ESS / s)", res["efficiency"])


if __name__ == "__main__":
    main()
