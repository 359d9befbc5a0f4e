"""GPU tests of the round-2 boundary additions (through the C ABI): user schedule function, restart from host state, monitored
extras, burn-in report records, device-side generic target, statistical agreement of GPU chains, the device tape generator,
chain read-out bounds, batched logistic chains (the as-shipped example sizes)."""
import ctypes as C

import numpy as np
import pytest

from util import relerr, example_config

pytestmark = pytest.mark.gpu


def _problem(N, p, seed=20260101):
    import oracle as o
    X, y, tt = o.synth_logistic(seed, N, p)
    return X, y, tt


def _oracle_sampler(o, X, y, theta0, nsteps, burnin, update="rand_exp_decay_update!", params=(10.0, 0.0), verbose=False, update_fn=None):
    otuner = o.GAMCTuner(o.VanillaMCTuner(verbose=verbose), o.VanillaMCTuner(verbose=verbose), o.AcceptanceRateMCTuner(0.35, verbose=verbose))
    return o.GAMCOracle(o.LogisticTarget(X, y, 100.0), theta0, nsteps, burnin, update=update, update_params=params, driftstep=0.02,
                        minorscale=0.001, c=0.01, tuner=otuner)


def test_user_schedule_function(gamc, engine):
    """`GAMC.update!::Function` (src/samplers/GAMC.jl:127): a closure instead of one of the nine descriptors.  The closure below
    restates rand_exp_decay_update!(a = 10), so the chain must be identical to the descriptor's and to the oracle's."""
    import math
    import oracle as o
    N, p, nsteps, burnin = 800, 11, 700, 100
    X, y, tt = _problem(N, p)
    theta0 = tt + 0.05 * np.random.default_rng(3).standard_normal(p)
    engine.target_logistic(X, y, 100.0)
    calls = []

    def my_update(i, tot, u):
        calls.append(i)
        return u <= math.exp(-10.0 * i / tot)

    chains = []
    for upd in (my_update, gamc.rand_exp_decay_update(10.0)):
        sampler = gamc.GAMC(update=upd, driftstep=0.02, minorscale=0.001, c=0.01, am_mode=1)
        tuner = gamc.GAMCTuner(gamc.VanillaMCTuner(), gamc.VanillaMCTuner(), gamc.AcceptanceRateMCTuner(0.35))
        job = gamc.BasicMCJob(gamc.LogisticModel(X, y, 100.0), sampler, gamc.BasicMCRange(nsteps, burnin), {"p": theta0}, tuner=tuner,
                              tape=gamc.RandomTape(nsteps, p, seed=5), engine=engine)
        job.run()
        chains.append(job.output())
    assert calls and min(calls) == 4 and max(calls) == nsteps          # only consulted for count > t0 (iterate/GAMC.jl:8-10)
    assert np.array_equal(chains[0].value, chains[1].value) and np.array_equal(chains[0].diagnosticvalues, chains[1].diagnosticvalues)
    osam = _oracle_sampler(o, X, y, theta0, nsteps, burnin)
    v, a, _ = osam.run(o.RandomTape(nsteps, p, seed=5))
    assert np.array_equal(chains[0].diagnosticvalues.ravel(), a) and relerr(chains[0].value, v) <= 1e-8
    # a schedule no descriptor expresses: SMMALA on the first 10 iterations of every block of 50
    sampler = gamc.GAMC(update=lambda i, tot, u: (i % 50) < 10, driftstep=0.02, minorscale=0.001, c=0.01)
    job = gamc.BasicMCJob(gamc.LogisticModel(X, y, 100.0), sampler, gamc.BasicMCRange(200, 0), {"p": theta0}, tape=gamc.RandomTape(200, p, seed=6), engine=engine)
    job.run()
    want = 3 + sum(1 for i in range(4, 201) if (i % 50) < 10)
    assert job.sstate.updatetensorcount == want


def test_set_state_restart_large_p(gamc, engine):
    """The same at p = 300 (the step kernels then need more than the default 48 KB of dynamic shared memory)."""
    N, p, nsteps, split = 2500, 300, 60, 42
    X, y, tt = _problem(N, p)
    theta0 = tt + 0.02 * np.random.default_rng(9).standard_normal(p)
    tape = gamc.RandomTape(nsteps, p, seed=13)
    engine.target_logistic(X, y, 100.0)
    sampler, tuner, rng_ = example_config(gamc, nsteps, 0, am_mode=1, update=gamc.mod_update(5))
    job = gamc.BasicMCJob(gamc.LogisticModel(X, y, 100.0), sampler, rng_, {"p": theta0}, tuner=tuner, tape=tape, engine=engine)
    job.run()
    full = job.output()
    job = gamc.BasicMCJob(gamc.LogisticModel(X, y, 100.0), sampler, rng_, {"p": theta0}, tuner=tuner, engine=engine)
    engine.tape_set(tape.u_sched[:split], tape.u_mix[:split], tape.z[:split], tape.u_acc[:split])
    engine.run(split)
    engine.sync()
    st = engine.state()
    assert st.pastupdatetensor == 0            # iterations 41, 42 of mod_update(5) are AM steps: the restart must carry the AM covariance
    saved = dict(theta=engine.array(0), m1=engine.array(3), m2=engine.array(4), C=engine.array(5))
    part1, acc1 = engine.chain_get(0, split)
    job = gamc.BasicMCJob(gamc.LogisticModel(X, y, 100.0), sampler, rng_, {"p": theta0}, tuner=tuner, engine=engine)
    engine.set_state(st, saved["theta"], saved["m1"], saved["m2"], saved["C"])
    engine.tape_set(tape.u_sched[split:], tape.u_mix[split:], tape.z[split:], tape.u_acc[split:])
    engine.run(nsteps - split)
    engine.sync()
    part2, acc2 = engine.chain_get(0, nsteps - split)
    assert np.array_equal(np.concatenate([acc1, acc2]), full.diagnosticvalues.ravel())
    assert relerr(np.concatenate([part1, part2], axis=1), full.value) <= 1e-9


@pytest.mark.parametrize("am_mode", [0, 1, 2, 3])
@pytest.mark.parametrize("split", [137, 150])
def test_set_state_restart(gamc, engine, am_mode, split):
    """gamc_set_state: stop a chain after `split` transitions, carry the state over the host, restart on a fresh sampler and
    finish -- the pieces must reproduce the uninterrupted chain (bit for bit with the literal AM recursion; to rounding when
    the AM covariance is carried in factored form, which the restart re-factorises)."""
    N, p, nsteps = 1500, 24, 300
    X, y, tt = _problem(N, p)
    theta0 = tt + 0.05 * np.random.default_rng(9).standard_normal(p)
    tape = gamc.RandomTape(nsteps, p, seed=13)
    engine.target_logistic(X, y, 100.0)
    sampler, tuner, rng_ = example_config(gamc, nsteps, 0, am_mode=am_mode)
    job = gamc.BasicMCJob(gamc.LogisticModel(X, y, 100.0), sampler, rng_, {"p": theta0}, tuner=tuner, tape=tape, engine=engine)
    job.run()
    full = job.output()
    # first part
    job = gamc.BasicMCJob(gamc.LogisticModel(X, y, 100.0), sampler, rng_, {"p": theta0}, tuner=tuner, engine=engine)
    engine.tape_set(tape.u_sched[:split], tape.u_mix[:split], tape.z[:split], tape.u_acc[:split])
    engine.run(split)
    engine.sync()
    st = engine.state()
    part1, acc1 = engine.chain_get(0, split)
    saved = dict(theta=engine.array(0), m1=engine.array(3), m2=engine.array(4), C=engine.array(5))
    assert st.count == split
    # restart
    job = gamc.BasicMCJob(gamc.LogisticModel(X, y, 100.0), sampler, rng_, {"p": theta0}, tuner=tuner, engine=engine)
    engine.set_state(st, saved["theta"], saved["m1"], saved["m2"], saved["C"])
    rest = nsteps - split
    engine.tape_set(tape.u_sched[split:], tape.u_mix[split:], tape.z[split:], tape.u_acc[split:])
    engine.run(rest)
    engine.sync()
    part2, acc2 = engine.chain_get(0, rest)
    assert engine.state().count == nsteps
    assert np.array_equal(np.concatenate([acc1, acc2]), full.diagnosticvalues.ravel())
    got = np.concatenate([part1, part2], axis=1)
    if am_mode == 0:
        assert np.array_equal(got, full.value)
    else:
        assert relerr(got, full.value) <= 1e-10


def test_monitored_extras_and_logtarget_chain(gamc, engine):
    """pstate.loglikelihood / logprior / gradloglikelihood / gradlogprior / tensorloglikelihood / tensorlogprior of the current
    point (the monitored extras iterate! copies at src/samplers/iterate/GAMC.jl:74-90,100-106) and the per-sample log-target."""
    import oracle as o
    N, p, nsteps = 900, 7, 120
    X, y, tt = _problem(N, p)
    theta0 = tt + 0.05 * np.random.default_rng(1).standard_normal(p)
    engine.target_logistic(X, y, 100.0)
    sampler, tuner, rng_ = example_config(gamc, nsteps, 20, update=gamc.smmala_only_update())
    job = gamc.BasicMCJob(gamc.LogisticModel(X, y, 100.0), sampler, rng_, {"p": theta0}, tuner=tuner, tape=gamc.RandomTape(nsteps, p, seed=2), engine=engine)
    job.run()
    chain = job.output()
    theta = chain.value[:, -1]
    tgt = o.LogisticTarget(X, y, 100.0)
    ll0, g0, G0 = o.logistic_partials(theta, X, y)
    lp0 = -0.5 * (theta @ theta / 100.0 + p * np.log(2 * np.pi * 100.0))
    s = job.sstate
    assert abs(s.loglikelihood - ll0) <= 1e-10 * abs(ll0) and abs(s.logprior - lp0) <= 1e-12 * abs(lp0)
    assert relerr(s.gradloglikelihood, g0) <= 1e-10 and relerr(s.gradlogprior, -theta / 100.0) <= 1e-14
    assert relerr(s.tensorloglikelihood, G0) <= 1e-10 and np.array_equal(s.tensorlogprior, np.eye(p) / 100.0)
    lts = engine.chain_get_logtarget(0, nsteps - 20)
    want = np.array([tgt.logtarget(chain.value[:, k]) for k in range(chain.value.shape[1])])
    assert relerr(lts, want) <= 1e-10


def test_burnin_report_records(gamc, engine, capsys):
    """The verbose burn-in report (src/samplers/iterate/GAMC.jl:199-292): one record per tuning event, equal to what the oracle's
    tuner held at that moment; the front-end prints the reference's lines from them."""
    import oracle as o
    N, p, nsteps, burnin = 300, 6, 1500, 500
    X, y, tt = _problem(N, p)
    theta0 = tt + 0.05 * np.random.default_rng(5).standard_normal(p)
    engine.target_logistic(X, y, 100.0)
    sampler, tuner, rng_ = example_config(gamc, nsteps, burnin, verbose=True)
    job = gamc.BasicMCJob(gamc.LogisticModel(X, y, 100.0), sampler, rng_, {"p": theta0}, tuner=tuner, tape=gamc.RandomTape(nsteps, p, seed=6), engine=engine)
    job.run()
    log = engine.tune_log()
    assert log.shape[0] == (burnin - 100) // 100 + 1        # totproposed starts at the period, tuning stops once it exceeds burnin
    assert np.array_equal(log[:, 0], 100.0 * np.arange(1, log.shape[0] + 1))
    assert np.all(log[:, 2] == 100) and np.all(log[:, 5] + log[:, 9] == 100)
    assert np.allclose(log[:, 3], log[:, 1] / log[:, 2])
    # the step after the last tuning event is the final step of the oracle run
    osam = _oracle_sampler(o, X, y, theta0, nsteps, burnin, verbose=True)
    osam.run(o.RandomTape(nsteps, p, seed=6))
    assert abs(log[-1, 12] - osam.tune.totaltune.step) <= 1e-12 * osam.tune.totaltune.step
    out = capsys.readouterr().out
    assert "Burnin iteration 100 out of 500..." in out and "  SMMALA: " in out and "  AM    : " in out and "running frequency" in out


def test_chain_get_bounds(gamc, engine):
    """Samples that were not produced yet cannot be read (the buffers are uninitialised device memory)."""
    N, p, nsteps = 300, 5, 60
    X, y, tt = _problem(N, p)
    engine.target_logistic(X, y, 100.0)
    sampler, tuner, rng_ = example_config(gamc, nsteps, 10)
    job = gamc.BasicMCJob(gamc.LogisticModel(X, y, 100.0), sampler, rng_, {"p": tt}, tuner=tuner, engine=engine)
    tape = gamc.RandomTape(nsteps, p, seed=1)
    with pytest.raises(gamc._lib.NotReady):
        job.output()
    engine.tape_set(tape.u_sched[:30], tape.u_mix[:30], tape.z[:30], tape.u_acc[:30])
    engine.run(30)
    v, a = engine.chain_get(0, 20)
    assert v.shape == (p, 20)
    with pytest.raises(gamc._lib.NotReady):
        engine.chain_get(0, 21)


def test_transform_on_builtin_target_is_refused(gamc, engine):
    X, y, tt = _problem(100, 3)
    sampler = gamc.GAMC(transform=gamc.softabs(1000.0), driftstep=0.02)
    with pytest.raises(NotImplementedError):
        gamc.BasicMCJob(gamc.LogisticModel(X, y, 100.0), sampler, gamc.BasicMCRange(10, 0), {"p": tt}, engine=engine)


def test_device_generic_target_gaussian_statistics(gamc, engine):
    """SURVEY.md 8f-3 with the model ON THE DEVICE (`gamc_target_device_callback`): torch enqueues the evaluation of a Gaussian
    target on the library's stream (no host synchronisation per evaluation).  Known-target check: the chain's mean and covariance
    match the exact ones within Monte-Carlo error, and the trajectory equals the oracle's on the same tape."""
    import torch
    import oracle as o
    p, nsteps, burnin = 6, 24000, 4000
    rng = np.random.default_rng(0)
    A = rng.standard_normal((p, p))
    Sigma = A @ A.T / p + np.eye(p)
    mu = rng.standard_normal(p)
    Prec = np.linalg.inv(Sigma)
    Prec = 0.5 * (Prec + Prec.T)
    dev = torch.device("cuda", 0)
    P_d = torch.tensor(Prec, dtype=torch.float64, device=dev)
    mu_d = torch.tensor(mu, dtype=torch.float64, device=dev)
    goff = (p + 2) & ~1
    n_out = goff + p * p

    def as_tensor(ptr, n):
        # wrap library-owned device memory (no copy): cuda array interface
        class _W:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}
        return torch.as_tensor(_W(), device=dev)

    def enqueue(d_theta, level, d_out, stream_ptr):
        ext = torch.cuda.ExternalStream(stream_ptr, device=dev)
        with torch.cuda.stream(ext):
            th = as_tensor(d_theta, p)
            out = as_tensor(d_out, n_out)
            d = th - mu_d
            Pd = P_d @ d
            out[0] = -0.5 * (d @ Pd)
            if level >= 1:
                out[1:1 + p] = -Pd
            if level >= 2:
                out[goff:goff + p * p] = P_d.t().reshape(-1)     # column-major; symmetric anyway

    engine.target_device_callback(p, enqueue)
    sampler = gamc.GAMC(update=gamc.rand_exp_decay_update(10.0), driftstep=1.0, minorscale=0.001, c=0.01, am_mode=0)
    tuner = gamc.GAMCTuner(gamc.VanillaMCTuner(), gamc.VanillaMCTuner(), gamc.AcceptanceRateMCTuner(0.35))
    theta0 = mu + 0.1 * rng.standard_normal(p)
    cfg = gamc.frontend.make_config(sampler, tuner, gamc.BasicMCRange(nsteps, burnin))
    engine.set_schedule_callback(None)
    engine.sampler_init(cfg, theta0)
    tape = gamc.RandomTape(nsteps, p, seed=77)
    engine.tape_set(tape.u_sched, tape.u_mix, tape.z, tape.u_acc)
    engine.run(nsteps)
    engine.sync()
    value, accept = engine.chain_get(0, nsteps - burnin)
    # oracle on the same tape
    class Gauss:
        def logtarget(self, th):
            d = th - mu
            return float(-0.5 * d @ Prec @ d)

        def uptotensorlogtarget(self, th):
            d = th - mu
            return float(-0.5 * d @ Prec @ d), -Prec @ d, Prec.copy()
    otuner = o.GAMCTuner(o.VanillaMCTuner(), o.VanillaMCTuner(), o.AcceptanceRateMCTuner(0.35))
    osam = o.GAMCOracle(Gauss(), theta0, nsteps, burnin, update="rand_exp_decay_update!", update_params=(10.0, 0.0), driftstep=1.0,
                        minorscale=0.001, c=0.01, tuner=otuner)
    v, a, _ = osam.run(o.RandomTape(nsteps, p, seed=77))
    assert np.array_equal(accept, a), "accept/reject decisions diverged"
    assert relerr(value, v) <= 1e-7
    # exact moments within Monte-Carlo error (ESS-based standard errors)
    ess = gamc.ess(value)
    se = np.sqrt(np.diag(Sigma) / ess)
    assert np.all(np.abs(value.mean(axis=1) - mu) <= 5 * se), (value.mean(axis=1) - mu, se)
    emp = np.cov(value)
    assert relerr(emp, Sigma) <= 0.25


def test_tape_generate_distribution(gamc, engine):
    """`gamc_tape_generate` (device Philox + Box-Muller, the throughput mode): Kolmogorov-Smirnov tests of the uniforms and the
    normals, moments, no correlation between neighbouring slots or coordinates, and distinct streams for distinct seeds."""
    from scipy import stats
    p, n = 9, 40000
    X, y, tt = _problem(64, p)
    engine.target_logistic(X, y, 100.0)
    engine.tape_generate(n, 12345)
    us, um, z, ua = engine.tape_get(0, n)
    for u in (us, um, ua):
        assert 0.0 <= u.min() and u.max() < 1.0 and stats.kstest(u, "uniform").pvalue > 1e-3
    assert ua.min() > 0.0                                   # log(u_acc) must be finite
    assert stats.kstest(z.ravel(), "norm").pvalue > 1e-3
    for j in range(p):
        assert stats.kstest(z[:, j], "norm").pvalue > 1e-4
    assert abs(z.mean()) < 4 / np.sqrt(z.size) and abs(z.var() - 1.0) < 0.02
    cc = np.corrcoef(z.T)
    assert np.max(np.abs(cc - np.eye(p))) < 5 / np.sqrt(n)
    assert abs(np.corrcoef(z[:-1, 0], z[1:, 0])[0, 1]) < 5 / np.sqrt(n)
    assert abs(np.corrcoef(us, um)[0, 1]) < 5 / np.sqrt(n) and abs(np.corrcoef(us, ua)[0, 1]) < 5 / np.sqrt(n)
    engine.tape_generate(100, 12346)
    assert not np.array_equal(engine.tape_get(0, 100)[2], z[:100])
    # odd p: the unpaired last Box-Muller output is dropped, not left uninitialised
    assert np.all(np.isfinite(z))


def test_logistic_posterior_statistics_gpu_vs_oracle(gamc, engine):
    """North star: 'posterior means and ESS must fall within Monte Carlo error'.  A GPU chain driven by the DEVICE-generated
    tape against an oracle chain driven by an independent host tape, on the same logistic posterior: coordinate means agree
    within 5 combined standard errors, acceptance rates and ESS are of the same size."""
    import oracle as o
    N, p, nsteps, burnin = 400, 4, 40000, 5000
    X, y, tt = _problem(N, p, seed=31)
    engine.target_logistic(X, y, 100.0)
    sampler, tuner, rng_ = example_config(gamc, nsteps, burnin, am_mode=2)
    job = gamc.BasicMCJob(gamc.LogisticModel(X, y, 100.0), sampler, rng_, {"p": tt}, tuner=tuner, engine=engine, tape_seed=2024)
    job.run()
    chain = job.output()
    osam = _oracle_sampler(o, X, y, tt, nsteps, burnin)
    v, a, _ = osam.run(o.RandomTape(nsteps, p, seed=4242))
    ess_g, ess_o = gamc.ess(chain.value), o.ess(v)
    se = np.sqrt(chain.value.var(axis=1) / ess_g + v.var(axis=1) / ess_o)
    assert np.all(np.abs(chain.value.mean(axis=1) - v.mean(axis=1)) <= 5 * se)
    assert abs(chain.diagnosticvalues.mean() - a.mean()) < 0.05
    assert 0.4 < ess_g.min() / ess_o.min() < 2.5


@pytest.mark.parametrize("nchains", [3])
def test_batched_logistic_parity(gamc, nchains):
    """The as-shipped logistic example sizes (N = 200, p = 4) as batched chains: every chain equals the oracle's chain on its
    own tape (examples/logistic_regression/src/GAMC/simulations.jl:13-37,50-61)."""
    import oracle as o
    N, p, nsteps, burnin = 200, 4, 1500, 300
    X, y, tt = _problem(N, p, seed=7)
    rng = np.random.default_rng(8)
    v0 = 3.0 * rng.standard_normal((nchains, p)) * 0.1 + tt
    sampler = gamc.GAMC(update=gamc.rand_exp_decay_update(10.0), driftstep=0.02, minorscale=0.001, c=0.01)
    tuner = gamc.GAMCTuner(gamc.VanillaMCTuner(), gamc.VanillaMCTuner(), gamc.AcceptanceRateMCTuner(0.35))
    tape = gamc.BatchedRandomTape(nchains, nsteps, p, seed=50)
    job = gamc.BatchedMCJob(gamc.LogisticModel(X, y, 100.0), sampler, gamc.BasicMCRange(nsteps, burnin), v0, tuner=tuner, tape=tape)
    job.run()
    chains = job.output()
    for c in range(nchains):
        osam = _oracle_sampler(o, X, y, v0[c], nsteps, burnin)
        v, a, _ = osam.run(o.RandomTape(nsteps, p, seed=50 + c))
        assert np.array_equal(chains[c].diagnosticvalues.ravel(), a), f"chain {c}: accept/reject decisions diverged"
        assert relerr(chains[c].value, v) <= 1e-8
    job.engine.close()
