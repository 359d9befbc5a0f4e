"""An arbitrary model through the generic (host-callback) target: the reference's t-distribution example
(examples/t_distribution/src/GAMC/simulations.jl:17-59) written the way a Klara user would -- closures for the log-target
and for (log-target, gradient, tensor), `transform = H -> softabs(H, 1000.)` -- with the sampler itself on the device.

One host round trip per evaluation: this is the drop-in for small, arbitrary models; the built-in targets
(`LogisticModel`, `MvTModel`, `RVModel`) keep the evaluation on the GPU.

  python examples/generic_target.py [--n 20 --nmcmc 20000 --nburnin 2000]
"""
import argparse
import math
import os
import sys
import time

import numpy as np
from scipy.special import gammaln

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gamc_b200 as g  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20)
    ap.add_argument("--nmcmc", type=int, default=20000)
    ap.add_argument("--nburnin", type=int, default=2000)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    n, nu = a.n, 30.0
    idx = np.arange(n)
    Sigma_t = (nu - 2.0) / nu * 0.9 ** np.abs(idx[:, None] - idx[None, :])          # (nu-2) Sigma / nu, Sigma_ij = 0.9^|i-j|  (:23-27)
    Sinv = np.linalg.inv(Sigma_t)
    const = gammaln((nu + n) / 2) - gammaln(nu / 2) - 0.5 * n * math.log(nu * math.pi) - 0.5 * np.linalg.slogdet(Sigma_t)[1]

    def plogtarget(p):                                                              # logpdf(MvTDist(nu, zeros(n), Sigma_t), p)  (:31)
        return const - 0.5 * (nu + n) * math.log1p(p @ Sinv @ p / nu)

    def puptotensorlogtarget(p):                                                    # what ForwardDiff gives the reference (:33)
        v = Sinv @ p
        q = p @ v
        s = (nu + n) / (nu + q)
        return plogtarget(p), -s * v, s * (Sinv - 2.0 * np.outer(v, v) / (nu + q))  # tensor = -Hessian

    model = g.GenericModel(n, puptotensorlogtarget, plogtarget)
    sampler = g.GAMC(update=g.rand_exp_decay_update(10.0), transform=g.softabs(1000.0), driftstep=0.02, minorscale=0.001, c=0.01)   # (:36-41)
    tuner = g.GAMCTuner(g.VanillaMCTuner(), g.VanillaMCTuner(), g.AcceptanceRateMCTuner(0.35))
    v0 = {"p": 2.0 * np.random.default_rng(1).standard_normal(n)}                   # rand(Normal(0, 2), n)  (:57)
    t0 = time.perf_counter()
    job = g.BasicMCJob(model, sampler, g.BasicMCRange(a.nmcmc, a.nburnin), v0, tuner=tuner, device=a.device, tape_seed=1)
    job.run()
    chain = job.output()
    dt = time.perf_counter() - t0
    ess = g.ess(chain.value)
    print(f"n={n}: {a.nmcmc} iterations in {dt:.2f} s ({a.nmcmc / dt:.0f} it/s, host-evaluated target), acceptance {g.acceptance(chain):.3f}, "
          f"min ESS {ess.min():.1f}, sample variance of coordinate 1: {chain.value[0].var():.3f} (target {nu / (nu - 2) * Sigma_t[0, 0]:.3f})")


if __name__ == "__main__":
    main()
