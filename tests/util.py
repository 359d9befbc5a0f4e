import numpy as np


def relerr(a, b):
    """Norm-wise relative error max|a-b| / max|b| (the 1e-10 FP64 tolerance of BASELINE.json's north star is
    applied to this quantity: single components of a gradient can be arbitrarily close to zero)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


def example_config(gamc, nsteps, burnin, am_mode=0, update=None, verbose=False):
    """The sampler configuration shipped with the reference's logistic example
    (examples/logistic_regression/src/GAMC/simulations.jl:50-59)."""
    sampler = gamc.GAMC(update=update if update is not None else gamc.rand_exp_decay_update(10.0),
                        driftstep=0.02, minorscale=0.001, c=0.01, am_mode=am_mode)
    tuner = gamc.GAMCTuner(gamc.VanillaMCTuner(verbose=verbose), gamc.VanillaMCTuner(verbose=verbose),
                           gamc.AcceptanceRateMCTuner(0.35, verbose=verbose))
    return sampler, tuner, gamc.BasicMCRange(nsteps=nsteps, burnin=burnin)
