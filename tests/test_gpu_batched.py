"""Batched-chains kernel (one persistent launch, target evaluated in the kernel) against the CPU oracle, chain by chain,
on shared per-chain random tapes: the t-distribution example of the reference (examples/t_distribution) with
`transform = softabs(H, 1000)`."""
import numpy as np
import pytest

from util import relerr

pytestmark = pytest.mark.gpu


def _oracle_chain(o, T, n, nu, theta0, nsteps, burnin, seed, update, params, transform, alpha):
    tgt = T.TTarget(n=n, nu=nu)
    tr = (lambda H: T.softabs(H, alpha)) if transform else None
    tuner = o.GAMCTuner(o.VanillaMCTuner(), o.VanillaMCTuner(), o.AcceptanceRateMCTuner(0.35))
    sam = o.GAMCOracle(tgt, theta0, nsteps, burnin, update=update, update_params=params, transform=tr,
                       driftstep=0.02, minorscale=0.001, c=0.01, tuner=tuner)
    v, a, k = sam.run(o.RandomTape(nsteps, n, seed=seed))
    return sam, v, a


@pytest.mark.parametrize("n,nchains,nsteps,update", [(20, 5, 800, "rand_exp_decay_update"), (7, 3, 600, "rand_update"),
                                                      (32, 2, 300, "smmala_only_update"), (12, 3, 500, "am_only_update")])
def test_batched_t_target_parity(gamc, engine, n, nchains, nsteps, update):
    import oracle as o
    from oracle import targets as T
    burnin = nsteps // 4
    rng = np.random.default_rng(11)
    v0s = 2.0 * rng.standard_normal((nchains, n))            # rand(Normal(0, 2), n), simulations.jl:57
    sched = getattr(gamc, update)()
    sampler = gamc.GAMC(update=sched, transform=gamc.softabs(1000.0), driftstep=0.02, minorscale=0.001, c=0.01)
    tuner = gamc.GAMCTuner(gamc.VanillaMCTuner(), gamc.VanillaMCTuner(), gamc.AcceptanceRateMCTuner(0.35))
    tape = gamc.BatchedRandomTape(nchains, nsteps, n, seed=100)
    job = gamc.BatchedMCJob(gamc.MvTModel(n=n, nu=30.0), sampler, gamc.BasicMCRange(nsteps, burnin), v0s, tuner=tuner, tape=tape, engine=engine)
    job.run()
    chains = job.output()
    for c in range(nchains):
        sam, v, a = _oracle_chain(o, T, n, 30.0, v0s[c], nsteps, burnin, 100 + c, sched.name, sched.params, True, 1000.0)
        assert np.array_equal(chains[c].diagnosticvalues.ravel(), a), f"chain {c}: accept/reject decisions diverged"
        assert relerr(chains[c].value, v) <= 1e-7
        st, th = engine.batched_state(c)
        assert st.count == nsteps and st.updatetensorcount == sam.updatetensorcount
        assert abs(st.step - sam.tune.totaltune.step) <= 1e-10 * sam.tune.totaltune.step
        assert relerr(th, sam.theta) <= 1e-7


def test_batched_philox_mode_statistics(gamc, engine):
    """Device-generated random numbers: many chains on the 20-dimensional t target; pooled posterior mean ~ 0."""
    n, nchains, nsteps = 20, 256, 3000
    rng = np.random.default_rng(3)
    sampler = gamc.GAMC(update=gamc.rand_exp_decay_update(10.0), transform=gamc.softabs(1000.0), driftstep=0.02, minorscale=0.001, c=0.01)
    tuner = gamc.GAMCTuner(gamc.VanillaMCTuner(), gamc.VanillaMCTuner(), gamc.AcceptanceRateMCTuner(0.35))
    job = gamc.BatchedMCJob(gamc.MvTModel(n=n), sampler, gamc.BasicMCRange(nsteps, 1000), 2.0 * rng.standard_normal((nchains, n)), tuner=tuner,
                            seed=7, engine=engine)
    job.run()
    chains = job.output()
    acc = np.mean([c.diagnosticvalues.mean() for c in chains])
    assert 0.1 < acc < 0.7
    pooled = np.concatenate([c.value for c in chains], axis=1)
    assert np.all(np.abs(pooled.mean(axis=1)) < 0.15)      # true mean 0, marginal sd ~ 1
    assert np.all(np.abs(pooled.std(axis=1) - 1.0) < 0.25)
    # distinct chains got distinct random streams
    assert not np.array_equal(chains[0].value, chains[1].value)
