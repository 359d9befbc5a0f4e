"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Tolerance: 1e-10 norm-wise relative error in FP64 (BASELINE.json north star) for log-target, gradient and
metric; trajectories on a shared random tape must agree step for step (identical accept decisions)."""
import numpy as np
import pytest

from util import relerr, example_config

pytestmark = pytest.mark.gpu

TOL = 1e-10


def _oracle():
    import oracle
    return oracle


@pytest.mark.parametrize("N,p", [(200, 4), (1000, 37), (33, 1), (5000, 256), (4097, 300), (777, 512), (36 * 50 + 1, 64)])
def test_eval_parity(engine, N, p):
    o = _oracle()
    X, y, tt = o.synth_logistic(20260101, N, p)
    rng = np.random.default_rng(7)
    theta = tt + 0.05 * rng.standard_normal(p)
    engine.target_logistic(X, y, 100.0)
    lt, g, G = engine.eval_upto_tensor(theta)
    tgt = o.LogisticTarget(X, y, 100.0)
    lt0, g0, G0 = tgt.uptotensorlogtarget(theta)
    assert abs(lt - lt0) <= TOL * abs(lt0)
    assert relerr(g, g0) <= TOL
    assert relerr(G, G0) <= TOL
    assert np.array_equal(G, G.T)  # exactly symmetric by construction
    assert abs(engine.eval_logtarget(theta) - lt0) <= TOL * abs(lt0)
    # likelihood-only partials (what a rank contributes to the all-reduce)
    ll, gl, Gl = engine.eval_partials(theta)
    ll0, gl0, Gl0 = o.logistic_partials(theta, X, y)
    assert abs(ll - ll0) <= TOL * abs(ll0)
    assert relerr(gl, gl0) <= TOL and relerr(Gl, Gl0) <= TOL


def test_eval_deterministic(engine):
    o = _oracle()
    X, y, tt = o.synth_logistic(3, 3000, 100)
    engine.target_logistic(X, y, 100.0)
    a = engine.eval_upto_tensor(tt)
    b = engine.eval_upto_tensor(tt)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_eval_extreme_eta(engine):
    """|eta| beyond the overflow point of the reference's naive log(1+exp(eta)) must stay finite."""
    o = _oracle()
    X, y, tt = o.synth_logistic(5, 500, 8)
    theta = 400.0 * np.ones(8)
    engine.target_logistic(X, y, 100.0)
    lt, g, G = engine.eval_upto_tensor(theta)
    lt0, g0, G0 = o.LogisticTarget(X, y, 100.0).uptotensorlogtarget(theta)
    assert np.isfinite(lt) and abs(lt - lt0) <= 1e-10 * abs(lt0)
    assert relerr(g, g0) <= 1e-10 and relerr(G, G0) <= 1e-9


@pytest.mark.parametrize("N,p,row0", [(1000, 20, 0), (513, 7, 100000)])
def test_synth_bitexact(engine, N, p, row0):
    o = _oracle()
    tt = o.logistic.synth_theta_true(11, p)
    X, y, _ = o.synth_logistic(11, row0 + N, p, row0=row0, nrows=N, theta_true=tt)
    engine.target_logistic_synth(11, row0, N, p, tt, 100.0)
    Xd, yd = engine.read_rows(0, N)
    assert np.array_equal(Xd, X)
    assert np.array_equal(yd, y)


def _run_both(gamc, engine, N, p, nsteps, burnin, seed, am_mode=0, update=None, theta_scale=0.05, verbose=False):
    o = _oracle()
    X, y, tt = o.synth_logistic(20260101 + p, N, p)
    rng = np.random.default_rng(seed)
    theta0 = tt + theta_scale * rng.standard_normal(p)
    sampler, tuner, rng_ = example_config(gamc, nsteps, burnin, am_mode=am_mode, update=update, verbose=verbose)
    tape = gamc.RandomTape(nsteps, p, seed=seed + 1)
    engine.target_logistic(X, y, 100.0)
    job = gamc.BasicMCJob(gamc.LogisticModel(X, y, 100.0), sampler, rng_, {"p": theta0}, tuner=tuner, tape=tape, engine=engine)
    job.run()
    chain = job.output()
    # oracle on the same tape
    otape = o.RandomTape(nsteps, p, seed=seed + 1)
    assert np.array_equal(otape.z, tape.z)
    otuner = o.GAMCTuner(o.VanillaMCTuner(verbose=verbose), o.VanillaMCTuner(verbose=verbose), o.AcceptanceRateMCTuner(0.35, verbose=verbose))
    osam = o.GAMCOracle(o.LogisticTarget(X, y, 100.0), theta0, nsteps, burnin, update=sampler.update.name,
                        update_params=sampler.update.params, driftstep=0.02, minorscale=0.001, c=0.01, tuner=otuner)
    v, a, k = osam.run(otape)
    return job, chain, osam, v, a, k


@pytest.mark.parametrize("am_mode", [0, 1, 2, 3])
@pytest.mark.parametrize("N,p,nsteps", [(200, 4, 3000), (500, 33, 1200), (2000, 64, 1000), (3000, 130, 1000), (6000, 256, 1000), (4000, 512, 1000),
                                        (2500, 300, 160), (2500, 352, 160), (2500, 353, 120)])
def test_trajectory_parity(gamc, engine, N, p, nsteps, am_mode):
    """Same host-drawn normals and uniforms => step-for-step agreement for >= 1000 iterations, including the BASELINE
    dimensions p = 256 and p = 512; p = 300 / 352 / 353 straddle the panel-buffer switch of the rank-one AM update.
    am_mode 0 = literal covariance recursion + Cholesky per AM step; 1 = rank-one Cholesky update (fast path);
    2 = the same update computed speculatively for both accept/reject outcomes while the data pass runs;
    3 = two consecutive AM steps share one pass over X (log-likelihood at the proposal and at both candidates of the next step)."""
    burnin = nsteps // 5
    job, chain, osam, v, a, k = _run_both(gamc, engine, N, p, nsteps, burnin, seed=42, am_mode=am_mode)
    assert chain.value.shape == v.shape
    assert np.array_equal(chain.diagnosticvalues.ravel(), a), "accept/reject decisions diverged"
    assert relerr(chain.value, v) <= 1e-8
    st = job.engine.state()
    assert st.count == nsteps == osam.count
    assert st.updatetensorcount == osam.updatetensorcount
    assert abs(st.step - osam.tune.totaltune.step) <= 1e-12 * osam.tune.totaltune.step
    assert st.total_totproposed == osam.tune.totaltune.totproposed
    assert relerr(job.engine.array(3), osam.lastmean) <= 1e-9   # lastmean


def test_trajectory_parity_verbose_counters(gamc, engine):
    job, chain, osam, v, a, k = _run_both(gamc, engine, 300, 6, 1500, 500, seed=5, verbose=True)
    assert np.array_equal(chain.diagnosticvalues.ravel(), a)
    st = job.engine.state()
    assert (st.smmala_totproposed, st.am_totproposed) == (osam.tune.smmalatune.totproposed, osam.tune.amtune.totproposed)
    assert (st.smmala_proposed, st.am_proposed) == (osam.tune.smmalatune.proposed, osam.tune.amtune.proposed)


@pytest.mark.parametrize("update", ["smmala_only_update", "am_only_update", "mod_update", "rand_update", "cos_update",
                                    "rand_lin_decay_update", "rand_pow_decay_update", "rand_quad_decay_update"])
def test_schedules_parity(gamc, engine, update):
    sched = getattr(gamc, update)()
    job, chain, osam, v, a, k = _run_both(gamc, engine, 250, 5, 400, 100, seed=9, update=sched)
    assert np.array_equal(chain.diagnosticvalues.ravel(), a)
    assert relerr(chain.value, v) <= 1e-8
    assert job.engine.state().updatetensorcount == osam.updatetensorcount


def test_state_arrays_parity(gamc, engine):
    """oldinvtensor / cholinvtensor / firstterm (one reverse Cholesky on the device) against the reference's
    literal inv + chol route in the oracle."""
    job, chain, osam, v, a, k = _run_both(gamc, engine, 400, 12, 40, 0, seed=3, update=gamc.smmala_only_update())
    eng = job.engine
    assert relerr(eng.array(5), osam.oldinvtensor) <= 1e-9
    assert relerr(eng.array(6), osam.cholinvtensor) <= 1e-9
    assert relerr(eng.array(7), osam.oldfirstterm) <= 1e-9
    assert relerr(eng.array(2), osam.G) <= 1e-10
    assert relerr(eng.array(1), osam.grad) <= 1e-10


@pytest.mark.parametrize("N,p,nsteps", [(400, 12, 25), (1500, 95, 10), (3000, 130, 12), (6000, 256, 9), (5000, 257, 6), (6000, 479, 4)])
def test_smmala_ratio_parity(gamc, engine, N, p, nsteps):
    """sstate.ratio of the last SMMALA transition (iterate/GAMC.jl:45-67).  The device forms the proposal-side quadratic
    form as a'G*a - eps a'grad* + eps^2/4 f*'grad* (a = theta - theta*, f* = G*^-1 grad*) from the column partials of
    k_land instead of the reference's literal (theta - mu*)'G*(theta - mu*); both must agree to rounding."""
    job, chain, osam, v, a, k = _run_both(gamc, engine, N, p, nsteps, 0, seed=21, update=gamc.smmala_only_update())
    st = job.engine.state()
    assert np.array_equal(chain.diagnosticvalues.ravel(), a)
    assert np.isfinite(osam.ratio)
    assert abs(st.ratio - osam.ratio) <= 1e-9 * max(1.0, abs(osam.ratio))
    assert abs(st.logtarget - osam.logtarget) <= 1e-10 * abs(osam.logtarget)


@pytest.mark.parametrize("am_mode", [0, 1, 2, 3])
def test_am_covariance_readout(gamc, engine, am_mode):
    """sstate.oldinvtensor after a run that ends in AM steps = the Haario covariance recursion seeded with inv(G)."""
    job, chain, osam, v, a, k = _run_both(gamc, engine, 400, 9, 60, 0, seed=4, am_mode=am_mode, update=gamc.mod_update(7))
    assert osam.last_kind == "am"
    assert np.array_equal(chain.diagnosticvalues.ravel(), a)
    assert relerr(job.engine.array(5), osam.oldinvtensor) <= 1e-9


def test_errors(gamc, engine):
    o = _oracle()
    X, y, tt = o.synth_logistic(1, 100, 3)
    engine.target_logistic(X, y, 100.0)
    with pytest.raises(AssertionError):
        gamc.GAMC(driftstep=-1.0)
    sampler, tuner, rng_ = example_config(gamc, 10, 0)
    with pytest.raises(AssertionError):   # non-finite initial value (GAMC.jl:168-170)
        gamc.BasicMCJob(gamc.LogisticModel(X, y, 100.0), sampler, rng_, {"p": np.array([np.nan, 0.0, 0.0])}, tuner=tuner, engine=engine)
    with pytest.raises(gamc._lib.UnsupportedShape):
        engine.target_logistic(np.zeros((10, 513)), np.zeros(10), 100.0)


@pytest.mark.parametrize("am_mode", [0, 1, 2, 3])
def test_am_before_enough_history_raises_posdef(gamc, engine, am_mode):
    """t0 = 1 lets an AM step run at t = 2/3, where Klara's covariance recursion has k = 0/1 and the proposal covariance
    is not positive definite: the reference throws from the MvNormal constructor (set_gmm, GAMC.jl:188); here the device
    raises GAMC_NOT_POSDEF, reported at the next synchronising call."""
    import oracle as o
    X, y, tt = o.synth_logistic(2, 200, 4)
    engine.target_logistic(X, y, 100.0)
    sampler = gamc.GAMC(update=gamc.am_only_update(), driftstep=0.02, minorscale=0.001, c=0.01, t0=1, am_mode=am_mode)
    tuner = gamc.GAMCTuner(gamc.VanillaMCTuner(), gamc.VanillaMCTuner(), gamc.AcceptanceRateMCTuner(0.35))
    job = gamc.BasicMCJob(gamc.LogisticModel(X, y, 100.0), sampler, gamc.BasicMCRange(10, 0), {"p": tt}, tuner=tuner,
                          tape=gamc.RandomTape(10, 4, seed=1), engine=engine)
    with pytest.raises(gamc._lib.PosDefError):
        job.run()
    # the oracle (the reference's semantics) fails at the same place
    osam = o.GAMCOracle(o.LogisticTarget(X, y, 100.0), tt, 10, 0, update="am_only_update!", driftstep=0.02, minorscale=0.001, c=0.01, t0=1)
    with pytest.raises((o.sampler.NotPosDef, ZeroDivisionError, FloatingPointError, np.linalg.LinAlgError)):
        with np.errstate(all="ignore"):
            osam.run(o.RandomTape(10, 4, seed=1))


@pytest.mark.parametrize("N,p,nsteps", [(300, 5, 2000), (3000, 100, 600)])
def test_mala_trajectory_parity(gamc, engine, N, p, nsteps):
    """Klara's MALA (the baseline of the reference's efficiency tables) on the fused log-lik + gradient pass."""
    import oracle as o
    X, y, tt = o.synth_logistic(77, N, p)
    theta0 = tt + 0.05 * np.random.default_rng(1).standard_normal(p)
    burnin = nsteps // 4
    engine.target_logistic(X, y, 100.0)
    tuner = gamc.AcceptanceRateMCTuner(0.574, score_k=3.0)          # MALA/simulations.jl:41
    tape = gamc.RandomTape(nsteps, p, seed=12)
    job = gamc.BasicMCJob(gamc.LogisticModel(X, y, 100.0), gamc.MALA(0.02), gamc.BasicMCRange(nsteps, burnin), {"p": theta0}, tuner=tuner,
                          tape=tape, engine=engine)
    job.run()
    chain = job.output()
    osam = o.MALAOracle(o.LogisticTarget(X, y, 100.0), theta0, nsteps, burnin, driftstep=0.02, tuner=o.AcceptanceRateMCTuner(0.574, score_k=3.0))
    v, a = osam.run(o.RandomTape(nsteps, p, seed=12))
    assert np.array_equal(chain.diagnosticvalues.ravel(), a)
    assert relerr(chain.value, v) <= 1e-9
    st = job.engine.state()
    assert abs(st.step - osam.tune.step) <= 1e-12 * osam.tune.step and st.count == nsteps


def test_speculative_am_is_bitwise_identical(gamc, engine):
    """am_mode 2 overlaps the factor update with the data pass, am_mode 3 evaluates two AM steps in one data pass (including across
    burn-in tuning events, which must not be paired); neither may change a single bit of the chain."""
    import oracle as o
    X, y, tt = o.synth_logistic(5, 5000, 96)
    theta0 = tt + 0.05 * np.random.default_rng(2).standard_normal(96)
    nsteps = 700
    engine.target_logistic(X, y, 100.0)
    out = []
    for mode in (1, 2, 3):
        sampler, tuner, rng_ = example_config(gamc, nsteps, 100, am_mode=mode, verbose=True)
        job = gamc.BasicMCJob(gamc.LogisticModel(X, y, 100.0), sampler, rng_, {"p": theta0}, tuner=tuner, tape=gamc.RandomTape(nsteps, 96, seed=4), engine=engine)
        job.run()
        st = engine.state()
        out.append((job.output(), st.step, st.am_totproposed, st.total_totproposed, engine.array(5)))
    for other in out[1:]:
        assert np.array_equal(out[0][0].value, other[0].value)
        assert np.array_equal(out[0][0].diagnosticvalues, other[0].diagnosticvalues)
        assert out[0][1:4] == other[1:4]
        assert np.array_equal(out[0][4], other[4])


def test_am_mode3_unpaired_fallback_is_bitwise_identical(gamc, engine, monkeypatch):
    """Below the size threshold am_mode 3 runs as am_mode 2 (speculative factor update on the side stream, one data pass per AM step):
    the same chain, bit for bit, as with the pairing forced on."""
    import oracle as o
    X, y, tt = o.synth_logistic(6, 4000, 48)
    theta0 = tt + 0.05 * np.random.default_rng(3).standard_normal(48)
    engine.target_logistic(X, y, 100.0)
    out = []
    for thr in ("0", "1e18"):
        monkeypatch.setenv("GAMC_AM_PAIR_MIN_BYTES", thr)
        sampler, tuner, rng_ = example_config(gamc, 400, 100, am_mode=3)
        job = gamc.BasicMCJob(gamc.LogisticModel(X, y, 100.0), sampler, rng_, {"p": theta0}, tuner=tuner, tape=gamc.RandomTape(400, 48, seed=5), engine=engine)
        job.run()
        out.append((job.output().value.copy(), engine.launch_count()))
    assert np.array_equal(out[0][0], out[1][0])


@pytest.mark.parametrize("am_mode", [0, 2])
def test_generic_target_callback_parity(gamc, engine, am_mode):
    """an arbitrary model through `gamc_target_callback` -- the host evaluates the Klara-style closures
    (here the t example with `transform = softabs(1000)`, examples/t_distribution/src/GAMC/simulations.jl:17-39), the
    device runs everything else.  Step-for-step agreement with the oracle on the same tape."""
    o = _oracle()
    from oracle import targets as T
    n, nsteps, burnin = 9, 1200, 200
    tgt = T.TTarget(n, 30.0, 0.9)
    theta0 = 0.3 * np.random.default_rng(7).standard_normal(n)
    sampler = gamc.GAMC(update=gamc.rand_exp_decay_update(10.0), transform=gamc.softabs(1000.0), driftstep=0.02, minorscale=0.001, c=0.01,
                        am_mode=am_mode)
    tuner = gamc.GAMCTuner(gamc.VanillaMCTuner(), gamc.VanillaMCTuner(), gamc.AcceptanceRateMCTuner(0.35))
    tape = gamc.RandomTape(nsteps, n, seed=11)
    model = gamc.GenericModel(n, tgt.uptotensorlogtarget, tgt.logtarget)
    job = gamc.BasicMCJob(model, sampler, gamc.BasicMCRange(nsteps, burnin), {"p": theta0}, tuner=tuner, tape=tape, engine=engine)
    job.run()
    chain = job.output()
    otuner = o.GAMCTuner(o.VanillaMCTuner(), o.VanillaMCTuner(), o.AcceptanceRateMCTuner(0.35))
    osam = o.GAMCOracle(tgt, theta0, nsteps, burnin, update="rand_exp_decay_update!", update_params=(10.0, 0.0),
                        transform=lambda H: T.softabs(H, 1000.0), driftstep=0.02, minorscale=0.001, c=0.01, tuner=otuner)
    v, a, k = osam.run(o.RandomTape(nsteps, n, seed=11))
    assert np.array_equal(chain.diagnosticvalues.ravel(), a), "accept/reject decisions diverged"
    assert relerr(chain.value, v) <= 1e-8
    st = job.engine.state()
    assert st.updatetensorcount == osam.updatetensorcount
    # a callable transform is applied on the host as well
    sampler2 = gamc.GAMC(update=gamc.smmala_only_update(), transform=lambda H: T.softabs(H, 1000.0), driftstep=0.02, minorscale=0.001, c=0.01)
    job2 = gamc.BasicMCJob(model, sampler2, gamc.BasicMCRange(50, 0), {"p": theta0}, tuner=tuner, tape=gamc.RandomTape(50, n, seed=12), engine=engine)
    job2.run()
    osam2 = o.GAMCOracle(tgt, theta0, 50, 0, update="smmala_only_update!", update_params=(), transform=lambda H: T.softabs(H, 1000.0),
                         driftstep=0.02, minorscale=0.001, c=0.01, tuner=o.GAMCTuner(o.VanillaMCTuner(), o.VanillaMCTuner(), o.AcceptanceRateMCTuner(0.35)))
    v2, a2, _ = osam2.run(o.RandomTape(50, n, seed=12))
    assert np.array_equal(job2.output().diagnosticvalues.ravel(), a2)
    assert relerr(job2.output().value, v2) <= 1e-8


def test_generic_target_callback_errors(gamc, engine):
    """A non-finite initial point raises the reference's initialisation error; a failing callback aborts the call."""
    sampler, tuner, rng_ = example_config(gamc, 10, 0)
    bad = gamc.GenericModel(3, lambda th: (float("nan"), np.zeros(3), np.eye(3)))
    with pytest.raises(gamc._lib.NonFiniteInitError):
        gamc.BasicMCJob(bad, sampler, rng_, {"p": np.zeros(3)}, tuner=tuner, tape=gamc.RandomTape(10, 3, seed=1), engine=engine)

    def boom(th):
        raise RuntimeError("model failure")
    with pytest.raises(AssertionError, match="callback"):
        gamc.BasicMCJob(gamc.GenericModel(3, boom), sampler, rng_, {"p": np.zeros(3)}, tuner=tuner, tape=gamc.RandomTape(10, 3, seed=1), engine=engine)


@pytest.mark.parametrize("am_mode", [0, 2])
def test_reuse_tensor_is_bitwise_identical(gamc, engine, am_mode):
    """gamc_config.reuse_tensor: skipping the re-evaluation of gradient and metric at a current point that no AM step has
    moved (src/samplers/iterate/GAMC.jl:19-31 always re-evaluates) must not change a single bit of the chain."""
    o = _oracle()
    N, p, nsteps = 3000, 40, 900
    X, y, tt = o.synth_logistic(5, N, p)
    theta0 = tt + 0.05 * np.random.default_rng(2).standard_normal(p)
    engine.target_logistic(X, y, 100.0)
    out = []
    for reuse in (False, True):
        sampler = gamc.GAMC(update=gamc.rand_exp_decay_update(10.0), driftstep=0.02, minorscale=0.001, c=0.01, am_mode=am_mode, reuse_tensor=reuse)
        tuner = gamc.GAMCTuner(gamc.VanillaMCTuner(), gamc.VanillaMCTuner(), gamc.AcceptanceRateMCTuner(0.35))
        job = gamc.BasicMCJob(gamc.LogisticModel(X, y, 100.0), sampler, gamc.BasicMCRange(nsteps, 0), {"p": theta0}, tuner=tuner,
                              tape=gamc.RandomTape(nsteps, p, seed=8), engine=engine)
        job.run()
        c = job.output()
        st = job.engine.state()
        out.append((c.value.copy(), c.diagnosticvalues.copy(), st.logtarget, st.step, st.updatetensorcount))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert out[0][2:] == out[1][2:]


@pytest.mark.parametrize("N,p", [(777, 33), (6000, 256), (5000, 130)])
def test_fused_level2_kernel_parity(gamc, engine, N, p, monkeypatch):
    """GAMC_FUSED_METRIC=1 (experimental, off by default): log-likelihood, gradient and metric from ONE pass over X
    (k_metric_fused) against the oracle and against the default two-pass route."""
    o = _oracle()
    X, y, tt = o.synth_logistic(77, N, p)
    theta = tt + 0.05 * np.random.default_rng(3).standard_normal(p)
    engine.target_logistic(X, y, 100.0)
    lt0, g0, G0 = engine.eval_upto_tensor(theta)
    monkeypatch.setenv("GAMC_FUSED_METRIC", "1")
    engine.target_logistic(X, y, 100.0)
    lt1, g1, G1 = engine.eval_upto_tensor(theta)
    monkeypatch.delenv("GAMC_FUSED_METRIC")
    ltr, gr, Gr = o.LogisticTarget(X, y, 100.0).uptotensorlogtarget(theta)
    assert abs(lt1 - ltr) <= 1e-10 * abs(ltr) and relerr(g1, gr) <= 1e-10 and relerr(G1, Gr) <= 1e-10
    assert abs(lt1 - lt0) <= 1e-12 * abs(lt0) and relerr(g1, g0) <= 1e-11 and relerr(G1, G0) <= 1e-12
    engine.target_logistic(X, y, 100.0)      # leave the engine on the default route


@pytest.mark.parametrize("N,p", [(200, 4), (777, 33), (6000, 256), (5000, 130), (40000, 200), (31, 64), (36 * 148 * 3 + 5, 256), (100001, 96)])
def test_coscheduled_level2_parity(gamc, engine, N, p, monkeypatch):
    """Default level-2 evaluation for p <= 256: k_datapass_co and the metric kernel run at the same time (one pass over X out of
    HBM, weights handed over through progress flags).  Against the oracle (1e-10), and against the two kernels run one after the
    other (GAMC_LEVEL2_CO=0): equal to rounding (1e-13; the data pass that shares the SM with the metric kernel sums eta and the gradient in
    a different order).  Repeated evaluations must give identical bits (deterministic); they also exercise the epoch
    counter of the flags; GAMC_CO_LAG_MB=0.01 forces the data pass to throttle on the metric frontier, =0 switches the throttle off."""
    o = _oracle()
    X, y, tt = o.synth_logistic(78, N, p)
    rng = np.random.default_rng(5)
    thetas = [tt + 0.05 * rng.standard_normal(p) for _ in range(3)]
    monkeypatch.setenv("GAMC_LEVEL2_CO", "0")
    engine.target_logistic(X, y, 100.0)
    serial = [engine.eval_upto_tensor(t) for t in thetas]
    monkeypatch.delenv("GAMC_LEVEL2_CO")
    tgt = o.LogisticTarget(X, y, 100.0)
    for lag in (None, "0.01", "0"):
        if lag is None:
            monkeypatch.delenv("GAMC_CO_LAG_MB", raising=False)
        else:
            monkeypatch.setenv("GAMC_CO_LAG_MB", lag)
        engine.target_logistic(X, y, 100.0)
        first = {}
        for rep in range(2):
            for it, (t, (lt0, g0, G0)) in enumerate(zip(thetas, serial)):
                lt1, g1, G1 = engine.eval_upto_tensor(t)
                if rep == 0:
                    first[it] = (lt1, g1, G1)
                else:
                    assert lt1 == first[it][0] and np.array_equal(g1, first[it][1]) and np.array_equal(G1, first[it][2])
                assert abs(lt1 - lt0) <= 1e-13 * abs(lt0) and relerr(G1, G0) <= 1e-13 and np.array_equal(G1, G1.T)
                assert relerr(g1, g0) <= 1e-13
                if rep == 0 and lag is None:
                    ltr, gr, Gr = tgt.uptotensorlogtarget(t)
                    assert abs(lt1 - ltr) <= 1e-10 * abs(ltr) and relerr(g1, gr) <= 1e-10 and relerr(G1, Gr) <= 1e-10
    monkeypatch.delenv("GAMC_CO_LAG_MB", raising=False)
    engine.target_logistic(X, y, 100.0)      # leave the engine on the default route
