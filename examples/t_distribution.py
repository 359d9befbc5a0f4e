"""The reference's t-distribution experiment (examples/t_distribution/src/GAMC/simulations.jl) through the batched-chains kernels:
`nchains` independent GAMC chains of the n-dimensional t target (nu = 30, Sigma_ij = 0.9^|i-j|, `transform = H -> softabs(H, 1000.)`,
initial values N(0, 2^2), :17-59) run at once; a chain is kept if its acceptance rate lies in (0.22, 0.37) (:68); outputs in the
reference's CSV formats.  n <= 32 (the reference uses 20) runs in one persistent kernel; 33 <= n <= 1024 on the structured
large-dimension path (BASELINE config: n = 1024, 4096 chains per GPU).

  python examples/t_distribution.py --outdir /tmp/gamc_t [--dim 20 --nchains 10 --nmcmc 110000 --nburnin 10000]
  python examples/t_distribution.py --dim 1024 --nchains 256 --nmcmc 2000 --nburnin 1000
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
    ap.add_argument("--outdir", default="/tmp/gamc_b200_t/GAMC")
    ap.add_argument("--dim", type=int, default=20)
    ap.add_argument("--nchains", type=int, default=10)
    ap.add_argument("--nmcmc", type=int, default=110000)
    ap.add_argument("--nburnin", type=int, default=10000)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    os.makedirs(a.outdir, exist_ok=True)
    rng = np.random.default_rng(1)
    v0s = 2.0 * rng.standard_normal((a.nchains, a.dim))                                            # rand(Normal(0, 2), npars)  (:57)
    sampler = g.GAMC(update=g.rand_exp_decay_update(10.0), transform=g.softabs(1000.0), driftstep=0.02, minorscale=0.001, c=0.01)   # :37-43
    tuner = g.GAMCTuner(g.VanillaMCTuner(), g.VanillaMCTuner(), g.AcceptanceRateMCTuner(0.35))                                      # :47
    t0 = time.perf_counter()
    job = g.BatchedMCJob(g.MvTModel(n=a.dim, nu=30.0, c=0.9), sampler, g.BasicMCRange(a.nmcmc, a.nburnin), v0s, tuner=tuner, device=a.device, seed=1)
    job.run()
    chains = job.output()
    runtime = time.perf_counter() - t0
    state = job.engine.bigd_state if job.bigd else job.engine.batched_state
    kept, times, steps, nupd = 0, [], [], []
    for c, chain in enumerate(chains):
        ratio = g.acceptance(chain)
        if 0.22 < ratio < 0.37:                                                                    # :68
            kept += 1
            g.write_chain_csv(a.outdir, kept, chain.value, chain.diagnosticvalues)
            st, _ = state(c)
            times.append(runtime / a.nchains)
            steps.append(st.step)
            nupd.append(st.updatetensorcount)
    g.write_run_csvs(a.outdir, times, steps, nupd)
    means = np.mean([ch.value.mean(axis=1) for ch in chains], axis=0)
    print(f"{a.nchains} chains x {a.nmcmc} iterations of the {a.dim}-dimensional t target in {runtime:.2f} s "
          f"({a.nchains * a.nmcmc / runtime:.0f} chain-iterations/s, {'structured large-dimension path' if job.bigd else 'persistent batched kernel'}); "
          f"{kept} chains inside the acceptance window; |posterior mean| max {np.max(np.abs(means)):.3f} (truth 0)")
    if kept:
        res = g.table(a.outdir, kept)
        print("rate", res["rate"], "min ESS", float(np.min(res["ess"])), "efficiency (min ESS / s)", res["efficiency"])
    job.engine.close()


if __name__ == "__main__":
    main()
