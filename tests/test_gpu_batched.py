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


def _rv_perturb_scale(npl):
    """make_param_perturb_scale, examples/radial_velocity/src/utils_ex.jl:23-46."""
    return np.array([0.0001, 0.01, 0.05, 0.05, np.pi * 0.01] * npl + [0.3])


@pytest.mark.parametrize("name,npl,nchains,nsteps", [("one_planet", 1, 4, 500), ("two_planets", 2, 2, 300)])
def test_batched_rv_target_parity(gamc, engine, name, npl, nchains, nsteps):
    """Radial-velocity example on the data shipped with the reference (tests/golden/rv_fixtures.json): in-kernel dual-number
    gradient / Hessian + softabs against the NumPy dual-number oracle, chain by chain on shared tapes."""
    import json
    import os
    import oracle as o
    from oracle import targets as T
    with open(os.path.join(os.path.dirname(__file__), "golden", "rv_fixtures.json")) as f:
        d = json.load(f)[name]
    p = 5 * npl + 1
    truth = np.array(d["theta_true"])
    rng = np.random.default_rng(21)
    v0s = truth + 0.01 * _rv_perturb_scale(npl) * rng.standard_normal((nchains, p))     # one_planet/GAMC/simulations.jl:63
    burnin = nsteps // 5
    sched = gamc.rand_exp_decay_update(10.0)
    sampler = gamc.GAMC(update=sched, transform=gamc.softabs(1000.0), driftstep=0.02, minorscale=0.001, c=0.01)
    tuner = gamc.GAMCTuner(gamc.VanillaMCTuner(), gamc.VanillaMCTuner(), gamc.AcceptanceRateMCTuner(0.35))
    tape = gamc.BatchedRandomTape(nchains, nsteps, p, seed=300)
    model = gamc.RVModel(np.array(d["t"]), np.array(d["rv"]), np.array(d["sigma"]), nplanets=npl)
    job = gamc.BatchedMCJob(model, sampler, gamc.BasicMCRange(nsteps, burnin), v0s, tuner=tuner, tape=tape, engine=engine)
    job.run()
    chains = job.output()
    om = T.RVModel(d["t"], d["rv"], d["sigma"])
    for c in range(nchains):
        otuner = o.GAMCTuner(o.VanillaMCTuner(), o.VanillaMCTuner(), o.AcceptanceRateMCTuner(0.35))
        sam = o.GAMCOracle(om, v0s[c], nsteps, burnin, update=sched.name, update_params=sched.params, transform=lambda H: T.softabs(H, 1000.0),
                           driftstep=0.02, minorscale=0.001, c=0.01, tuner=otuner)
        v, a, k = sam.run(o.RandomTape(nsteps, p, seed=300 + c))
        assert np.array_equal(chains[c].diagnosticvalues.ravel(), a), f"chain {c}: accept/reject decisions diverged"
        # tolerance: the softabs metric of this target has condition number 1e6 (one planet) to 1e8 (two planets), so
        # rounding differences between the Jacobi (device) and LAPACK (oracle) eigen-decompositions are amplified accordingly
        assert relerr(chains[c].value, v) <= (1e-6 if npl == 1 else 2e-5)
        st, th = engine.batched_state(c)
        assert st.updatetensorcount == sam.updatetensorcount and st.count == nsteps


def test_batched_rv_c5_size_parity(gamc, engine):
    """BASELINE config 5 at its full size (N = 2^18 epochs, one planet): three SMMALA transitions (second-order duals in the kernel,
    block-cooperative reduction over 262 144 epochs) against the oracle's accept decisions, values and final log-target."""
    import oracle as o
    from oracle import targets as T
    rng = np.random.default_rng(11)
    N = 1 << 18
    t = np.sort(rng.uniform(0.0, 730.5, N))
    truth = T.make_param_true_ex1()
    sig = 2.0 * np.ones(N)
    obs = T.RVModel(t, np.zeros(N), sig).model_rv(truth) + 2.0 * rng.standard_normal(N)
    nsteps = 3
    sampler = gamc.GAMC(update=gamc.smmala_only_update(), transform=gamc.softabs(1000.0), driftstep=0.02, minorscale=0.001, c=0.01)
    tuner = gamc.GAMCTuner(gamc.VanillaMCTuner(), gamc.VanillaMCTuner(), gamc.AcceptanceRateMCTuner(0.35))
    v0s = truth + 3e-6 * _rv_perturb_scale(1) * rng.standard_normal((1, 6))
    tape = gamc.BatchedRandomTape(1, nsteps, 6, seed=19)
    job = gamc.BatchedMCJob(gamc.RVModel(t, obs, sig, 1), sampler, gamc.BasicMCRange(nsteps, 0), v0s, tuner=tuner, tape=tape, engine=engine)
    job.run()
    chains = job.output()
    sam = o.GAMCOracle(T.RVModel(t, obs, sig), v0s[0], nsteps, 0, update="smmala_only_update!", transform=lambda H: T.softabs(H, 1000.0), driftstep=0.02,
                       minorscale=0.001, c=0.01, tuner=o.GAMCTuner(o.VanillaMCTuner(), o.VanillaMCTuner(), o.AcceptanceRateMCTuner(0.35)))
    v, a, _ = sam.run(o.RandomTape(nsteps, 6, seed=19))
    assert np.array_equal(chains[0].diagnosticvalues.ravel(), a)
    assert relerr(chains[0].value, v) <= 1e-6
    st, _ = engine.batched_state(0)
    assert abs(st.logtarget - sam.logtarget) <= 1e-9 * abs(sam.logtarget)


def test_batched_rv_many_epochs_derivatives(gamc, engine):
    """Synthetic epochs (N = 20000, data_generation.jl:11-33 recipe): one SMMALA step from the truth must reproduce the
    oracle's log-target after the step, i.e. the block-cooperative reduction over epochs is exact to FP64 rounding."""
    import oracle as o
    from oracle import targets as T
    rng = np.random.default_rng(5)
    N = 20000
    t = np.sort(rng.uniform(0.0, 730.5, N))
    truth = T.make_param_true_ex1()
    sig = 2.0 * np.ones(N)
    obs = T.RVModel(t, np.zeros(N), sig).model_rv(truth) + 2.0 * rng.standard_normal(N)
    nsteps = 12
    sampler = gamc.GAMC(update=gamc.smmala_only_update(), transform=gamc.softabs(1000.0), driftstep=0.02, minorscale=0.001, c=0.01)
    tuner = gamc.GAMCTuner(gamc.VanillaMCTuner(), gamc.VanillaMCTuner(), gamc.AcceptanceRateMCTuner(0.35))
    v0s = truth + 1e-5 * _rv_perturb_scale(1) * rng.standard_normal((2, 6))
    tape = gamc.BatchedRandomTape(2, nsteps, 6, seed=9)
    job = gamc.BatchedMCJob(gamc.RVModel(t, obs, sig, 1), sampler, gamc.BasicMCRange(nsteps, 0), v0s, tuner=tuner, tape=tape, engine=engine)
    job.run()
    chains = job.output()
    om = T.RVModel(t, obs, sig)
    for c in range(2):
        sam = o.GAMCOracle(om, v0s[c], nsteps, 0, update="smmala_only_update!", transform=lambda H: T.softabs(H, 1000.0), driftstep=0.02,
                           minorscale=0.001, c=0.01, tuner=o.GAMCTuner(o.VanillaMCTuner(), o.VanillaMCTuner(), o.AcceptanceRateMCTuner(0.35)))
        v, a, _ = sam.run(o.RandomTape(nsteps, 6, seed=9 + c))
        assert np.array_equal(chains[c].diagnosticvalues.ravel(), a)
        assert relerr(chains[c].value, v) <= 1e-6
        st, _ = engine.batched_state(c)
        assert abs(st.logtarget - sam.logtarget) <= 1e-9 * abs(sam.logtarget)
