"""GPU parity of the large-dimension t-target path (gamc_bigd_*: structured SMMALA, dense AM) against the dense CPU oracle
(oracle/targets.py TTarget + `eigh` softabs + GAMCOracle with its literal inv / cholesky / slogdet) on shared random tapes.
BASELINE config 4 is this target at d = 1024 with 4096 chains per GPU (examples/t_distribution/src/GAMC/simulations.jl:17-39,56-79)."""
import numpy as np
import pytest

from util import relerr

pytestmark = pytest.mark.gpu


def _run(gamc, n, nchains, nsteps, burnin, init_scale, seed, update=None, max_factors=0):
    import oracle as o
    from oracle import targets as T
    rng = np.random.default_rng(seed)
    if init_scale is None:          # the typical set of the target: x ~ N(0, Sigma) -- softabs then changes a single eigenvalue
        idx = np.arange(n)
        v0 = rng.multivariate_normal(np.zeros(n), 0.9 ** np.abs(idx[:, None] - idx[None, :]), size=nchains)
    else:
        v0 = init_scale * rng.standard_normal((nchains, n))
    sampler = gamc.GAMC(update=update if update is not None else gamc.rand_exp_decay_update(10.0), transform=gamc.softabs(1000.0),
                        driftstep=0.02, minorscale=0.001, c=0.01)
    tuner = gamc.GAMCTuner(gamc.VanillaMCTuner(), gamc.VanillaMCTuner(), gamc.AcceptanceRateMCTuner(0.35))
    tape = gamc.BatchedRandomTape(nchains, nsteps, n, seed=100 + seed)
    job = gamc.BatchedMCJob(gamc.MvTModel(n=n), sampler, gamc.BasicMCRange(nsteps, burnin), v0, tuner=tuner, tape=tape, max_factors=max_factors)
    assert job.bigd
    job.run()
    chains = job.output()
    tgt = T.TTarget(n, 30.0, 0.9)
    outs = []
    for c in range(nchains):
        osam = o.GAMCOracle(tgt, v0[c], nsteps, burnin, update=sampler.update.name, update_params=sampler.update.params,
                            transform=lambda H: T.softabs(H, 1000.0), driftstep=0.02, minorscale=0.001, c=0.01,
                            tuner=o.GAMCTuner(o.VanillaMCTuner(), o.VanillaMCTuner(), o.AcceptanceRateMCTuner(0.35)))
        v, a, _ = osam.run(o.RandomTape(nsteps, n, seed=100 + seed + c))
        outs.append((osam, v, a))
    return job, chains, outs


@pytest.mark.parametrize("n,nsteps,init_scale", [(40, 600, 2.0), (64, 500, 1.0), (130, 400, 2.0), (256, 300, 1.0), (96, 500, None), (300, 250, None)])
def test_bigd_trajectory_parity(gamc, n, nsteps, init_scale):
    """Step-for-step agreement with the dense oracle; N(0, 2^2) initial values (the reference's, simulations.jl:57) put many
    eigenvalues of the tensor below 21/alpha, so the softabs correction has tens of columns (several product factors at n = 130)."""
    job, chains, outs = _run(gamc, n, 3, nsteps, nsteps // 5, init_scale, seed=n)
    for c, (osam, v, a) in enumerate(outs):
        assert np.array_equal(chains[c].diagnosticvalues.ravel(), a), f"chain {c}: accept/reject decisions diverged"
        assert relerr(chains[c].value, v) <= 1e-6
        st, th = job.engine.bigd_state(c)
        if init_scale is None:
            assert st.error_pivot <= 3          # few eigenvalues changed by softabs: the four-column code paths ran
        assert st.count == nsteps and st.updatetensorcount == osam.updatetensorcount
        assert abs(st.step - osam.tune.totaltune.step) <= 1e-12 * osam.tune.totaltune.step
        assert abs(st.logtarget - osam.logtarget) <= 1e-8 * abs(osam.logtarget)
    job.engine.close()


@pytest.mark.parametrize("update", ["smmala_only_update", "am_only_update", "mod_update"])
def test_bigd_schedules(gamc, update):
    job, chains, outs = _run(gamc, 48, 2, 200, 0, 1.5, seed=7, update=getattr(gamc, update)())
    for c, (osam, v, a) in enumerate(outs):
        assert np.array_equal(chains[c].diagnosticvalues.ravel(), a)
        assert relerr(chains[c].value, v) <= 1e-6
    job.engine.close()


def test_bigd_d1024_short_trajectory(gamc):
    """The BASELINE dimension itself: d = 1024, two chains, 60 transitions from N(0, 2^2) initial values (about 127 eigenvalues
    below 21/alpha: four product factors) against the dense oracle (`eigh` of a 1024 x 1024 matrix per evaluation)."""
    job, chains, outs = _run(gamc, 1024, 2, 60, 0, 2.0, seed=3, update=gamc.mod_update(3))
    for c, (osam, v, a) in enumerate(outs):
        assert np.array_equal(chains[c].diagnosticvalues.ravel(), a)
        assert relerr(chains[c].value, v) <= 1e-6
        st, _ = job.engine.bigd_state(c)
        assert st.error_pivot > 64            # the softabs correction really had many columns
    job.engine.close()


def test_bigd_capacity_error(gamc):
    """A chain that needs more product factors than allocated fails loudly instead of truncating the correction."""
    rng = np.random.default_rng(0)
    v0 = 2.0 * rng.standard_normal((1, 400))
    sampler = gamc.GAMC(update=gamc.smmala_only_update(), transform=gamc.softabs(1000.0), driftstep=0.02, minorscale=0.001, c=0.01)
    with pytest.raises(gamc._lib.UnsupportedShape):
        gamc.BatchedMCJob(gamc.MvTModel(n=400), sampler, gamc.BasicMCRange(10, 0), v0, max_factors=1)


def test_bigd_philox_statistics(gamc):
    """Device-Philox mode at n = 64: many chains started in the typical set stay finite and centred, accept at a sensible rate and
    are distinct.  (No variance check against the target: the AM branch of the reference adapts its proposal covariance to the
    chain's own history without a Hastings correction, and the oracle shows the same transient contraction of x'Tx -- a property
    of the algorithm, reproduced step for step in the parity tests above.)"""
    n, nchains, nsteps, burnin = 64, 64, 3000, 1000
    rng = np.random.default_rng(2)
    idx = np.arange(n)
    Sigma = 0.9 ** np.abs(idx[:, None] - idx[None, :])
    v0 = rng.multivariate_normal(np.zeros(n), Sigma, size=nchains)
    sampler = gamc.GAMC(update=gamc.rand_exp_decay_update(10.0), transform=gamc.softabs(1000.0), driftstep=0.02, minorscale=0.001, c=0.01)
    tuner = gamc.GAMCTuner(gamc.VanillaMCTuner(), gamc.VanillaMCTuner(), gamc.AcceptanceRateMCTuner(0.35))
    job = gamc.BatchedMCJob(gamc.MvTModel(n=n), sampler, gamc.BasicMCRange(nsteps, burnin), v0, tuner=tuner, seed=5)
    job.run()
    chains = job.output()
    allv = np.concatenate([c.value for c in chains], axis=1)
    assert np.all(np.isfinite(allv))
    acc = np.mean([c.diagnosticvalues.mean() for c in chains])
    assert 0.1 < acc < 0.8
    assert abs(allv.mean()) < 0.1 and 0.05 < allv.var() < 2.0
    assert not np.array_equal(chains[0].value, chains[1].value)
    job.engine.close()
