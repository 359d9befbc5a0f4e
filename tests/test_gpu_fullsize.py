"""Parity at BASELINE.json's full size (C2: N = 2^20, p = 256) through size-independent properties -- the oracle cannot
evaluate 2.1 GB problems in seconds, so the CUDA path is checked against itself across decompositions, against the
oracle on row subsets of the device-resident data, and against calculus identities the model satisfies."""
import numpy as np
import pytest

from util import relerr

pytestmark = pytest.mark.gpu

N, P, SEED = 1 << 20, 256, 20260101


@pytest.fixture(scope="module")
def big(gamc):
    import oracle.logistic as ol
    tt = ol.synth_theta_true(SEED, P)
    eng = gamc.Engine(0)
    eng.target_logistic_synth(SEED, 0, N, P, tt, 100.0)
    yield eng, tt
    eng.close()


def test_device_data_matches_oracle_generator_on_row_subsets(big):
    import oracle as o
    eng, tt = big
    for r0 in (0, 123457, N - 300):
        Xd, yd = eng.read_rows(r0, 300)
        X, y, _ = o.synth_logistic(SEED, N, P, row0=r0, nrows=300, theta_true=tt)
        assert np.array_equal(Xd, X) and np.array_equal(yd, y)


def test_row_shards_add_up_to_the_whole(gamc, big):
    """Two handles holding the halves of the rows: partial sums add up to the whole-problem partials (what the all-reduce does)."""
    eng, tt = big
    theta = tt + 0.01 * np.random.default_rng(2).standard_normal(P)
    ll, g, G = eng.eval_partials(theta)
    parts = []
    for r0, n in ((0, N // 2 + 12345), (N // 2 + 12345, N - (N // 2 + 12345))):
        e2 = gamc.Engine(0)
        e2.target_logistic_synth(SEED, r0, n, P, tt, 100.0)
        parts.append(e2.eval_partials(theta))
        e2.close()
    assert abs(parts[0][0] + parts[1][0] - ll) <= 1e-11 * abs(ll)
    assert relerr(parts[0][1] + parts[1][1], g) <= 1e-11
    assert relerr(parts[0][2] + parts[1][2], G) <= 1e-11


def test_gradient_and_metric_are_the_derivatives_of_the_logtarget(big):
    """Directional finite differences: d/dh logtarget(theta + h v) = grad.v and, because the expected Fisher information of
    the logistic model equals minus the Hessian, -d/dh grad(theta + h v).v = v'Gv."""
    eng, tt = big
    rng = np.random.default_rng(3)
    theta = tt + 0.01 * rng.standard_normal(P)
    v = rng.standard_normal(P)
    v /= np.linalg.norm(v)
    lt, g, G = eng.eval_upto_tensor(theta)
    h = 1e-5
    ltp, gp, _ = eng.eval_upto_tensor(theta + h * v)
    ltm, gm, _ = eng.eval_upto_tensor(theta - h * v)
    assert abs((ltp - ltm) / (2 * h) - g @ v) <= 1e-6 * abs(g @ v)
    assert abs(-(gp - gm) @ v / (2 * h) - v @ G @ v) <= 1e-6 * (v @ G @ v)
    assert np.array_equal(G, G.T) and np.all(np.linalg.eigvalsh(G) > 0)
    assert abs(eng.eval_logtarget(theta) - lt) <= 1e-13 * abs(lt)


def test_oracle_on_a_subsample_matches_the_partial_of_those_rows(gamc, big):
    import oracle as o
    eng, tt = big
    theta = tt + 0.01 * np.random.default_rng(4).standard_normal(P)
    r0, n = 700001, 4096
    X, y = eng.read_rows(r0, n)
    e2 = gamc.Engine(0)
    e2.target_logistic_synth(SEED, r0, n, P, tt, 100.0)
    ll, g, G = e2.eval_partials(theta)
    e2.close()
    ll0, g0, G0 = o.logistic_partials(theta, X, y)
    assert abs(ll - ll0) <= 1e-10 * abs(ll0) and relerr(g, g0) <= 1e-10 and relerr(G, G0) <= 1e-10


def test_chain_runs_are_reproducible_and_independent_of_chunking(gamc, big):
    """Same tape => bit-identical chain whether `run` is called once or in three chunks (state fully device-resident)."""
    eng, tt = big
    theta0 = tt + 0.01 * np.random.default_rng(5).standard_normal(P)
    nsteps = 60
    tape = gamc.RandomTape(nsteps, P, seed=8)
    sampler = gamc.GAMC(update=gamc.rand_update(0.3), driftstep=0.02, minorscale=0.001, c=0.01, am_mode=1)
    tuner = gamc.GAMCTuner(gamc.VanillaMCTuner(), gamc.VanillaMCTuner(), gamc.AcceptanceRateMCTuner(0.35, period=10))
    out = []
    for chunks in ((60,), (7, 30, 23)):
        job = gamc.BasicMCJob(gamc.LogisticModel(lam=100.0), sampler, gamc.BasicMCRange(nsteps, 0), {"p": theta0}, tuner=tuner, tape=tape, engine=eng)
        lo = 0
        for c in chunks:
            eng.tape_set(tape.u_sched[lo:lo + c], tape.u_mix[lo:lo + c], tape.z[lo:lo + c], tape.u_acc[lo:lo + c])
            eng.run(c)
            lo += c
        eng.sync()
        out.append(eng.chain_get(0, nsteps))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert 0 < out[0][1].mean() < 1
