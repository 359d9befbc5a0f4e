"""The GAMC transition restatement pinned on identities it must satisfy, and the algebra the CUDA path relies on."""
import math

import numpy as np
import pytest

import oracle as o
from oracle import sampler as osamp
from util import relerr


def test_decay_functions_and_schedule_defaults():
    assert o.exp_decay(5, 100) == pytest.approx(math.exp(-0.5))
    assert o.lin_decay(10, 100) == pytest.approx(1 / (1 + 1e-3 * 10))
    assert o.pow_decay(50, 100) == pytest.approx(1e-3 ** 0.5)
    assert o.quad_decay(10, 100) == pytest.approx(1 / (1 + 1e-3 * 100))
    assert o.exp_decay(5, 100, 10.0, 0.25) == pytest.approx(0.75 * math.exp(-0.5) + 0.25)
    # the rand_* wrappers carry their own defaults (src/samplers/GAMC.jl:287-334)
    assert o.schedule_probability("rand_lin_decay_update!", 10, 100)[0] == pytest.approx(1 / (1 + 1e-2 * 10))
    assert o.schedule_probability("rand_quad_decay_update!", 10, 100)[0] == pytest.approx(1 / (1 + 1e-5 * 100))
    assert o.schedule_probability("mod_update!", 20, 100) == (1.0, False) and o.schedule_probability("mod_update!", 21, 100) == (0.0, False)
    assert o.schedule_probability("cos_update!", 5, 100)[0] == pytest.approx(0.2 * math.cos(10 * 5 * math.pi / 100) + 0.2)
    assert len(o.SCHEDULES) == 9


def test_haario_recursion_is_the_sample_covariance():
    """Klara covariance! as restated with the reference's argument roles reproduces np.cov."""
    rng = np.random.default_rng(0)
    xs = rng.standard_normal((40, 5)) @ rng.standard_normal((5, 5))
    means = [xs[: k + 1].mean(0) for k in range(40)]
    C = np.cov(xs[:3].T)
    for k in range(3, 40):     # add point x_{k+1} (index k): k points before
        C = o.covariance_update(C, k, xs[k], means[k], means[k - 1])
        assert relerr(C, np.cov(xs[: k + 1].T)) < 1e-10


def test_three_term_recursion_is_a_rank_one_update():
    """With m1 = (k m2 + x)/(k+1) the recursion equals ((k-1)/k) C + d d'/(k+1), d = x - m2, for ANY seed C
    (the identity behind am_mode = 1)."""
    rng = np.random.default_rng(1)
    p, k = 6, 7
    A = rng.standard_normal((p, p)); C = A @ A.T
    m2, x = rng.standard_normal(p), rng.standard_normal(p)
    m1 = (k * m2 + x) / (k + 1)
    lhs = o.covariance_update(C, k, x, m1, m2)
    d = x - m2
    assert relerr(lhs, (k - 1) / k * C + np.outer(d, d) / (k + 1)) < 1e-12


def test_closed_form_factor_of_identity_plus_rank_one():
    rng = np.random.default_rng(2)
    p = 33
    A = rng.standard_normal((p, p)); L = np.linalg.cholesky(A @ A.T + p * np.eye(p))
    u = rng.standard_normal(p); a = 0.93
    w = np.linalg.solve(a * L, u)
    s = np.concatenate([[1.0], 1.0 + np.cumsum(w ** 2)])
    K = np.tril(np.outer(w, w) / np.sqrt(s[:-1] * s[1:])[None, :], -1) + np.diag(np.sqrt(s[1:] / s[:-1]))
    assert relerr(K @ K.T, np.eye(p) + np.outer(w, w)) < 1e-12
    assert relerr((a * L) @ K, np.linalg.cholesky(a * a * (L @ L.T) + np.outer(u, u))) < 1e-12


def test_reverse_cholesky_gives_the_factor_of_the_inverse():
    """chol(Hermitian(inv(G)))' of the reference (iterate/GAMC.jl:30) == U^-T with G = U U', U upper."""
    rng = np.random.default_rng(3)
    p = 16
    A = rng.standard_normal((p, p)); G = A @ A.T + p * np.eye(p)
    J = np.eye(p)[::-1]
    U = J @ np.linalg.cholesky(J @ G @ J) @ J
    assert np.allclose(U, np.triu(U)) and relerr(U @ U.T, G) < 1e-13
    L = np.linalg.cholesky(np.linalg.inv(G))
    assert relerr(np.linalg.inv(U).T, L) < 1e-12
    z = rng.standard_normal(p); eps = 0.3
    d = math.sqrt(eps) * (L @ z)
    assert float(d @ G @ d) / eps == pytest.approx(float(z @ z), rel=1e-12)      # the Mahalanobis term is z'z
    assert np.linalg.slogdet(eps * np.linalg.inv(G))[1] == pytest.approx(p * math.log(eps) - 2 * np.sum(np.log(np.diag(U))), rel=1e-12)


def test_gmm_terms_cancel_exactly():
    """logpdf(q(.|theta), theta*) and logpdf(q(.|theta*), theta) are bit-identical (iterate/GAMC.jl:152,156)."""
    rng = np.random.default_rng(4)
    p = 9
    A = rng.standard_normal((p, p)); Lc = np.linalg.cholesky(0.02 * (A @ A.T + np.eye(p)))
    x, m = rng.standard_normal(p), rng.standard_normal(p)
    w = np.array([0.99, 0.01])
    assert o.gmm_logpdf(x, m, Lc, math.sqrt(0.001), w) == o.gmm_logpdf(m, x, Lc, math.sqrt(0.001), w)


class _Gauss:
    def __init__(self, mean, cov):
        self.mean, self.P = mean, np.linalg.inv(cov)

    def logtarget(self, x):
        d = x - self.mean
        return float(-0.5 * d @ self.P @ d)

    def uptotensorlogtarget(self, x):
        d = x - self.mean
        return float(-0.5 * d @ self.P @ d), -self.P @ d, self.P.copy()


@pytest.mark.parametrize("update", ["smmala_only_update!", "rand_update!", "am_only_update!"])
def test_gaussian_target_moments(update):
    """Known target: posterior mean / covariance within Monte Carlo error."""
    rng = np.random.default_rng(5)
    p = 3
    mean = np.array([1.0, -2.0, 0.5])
    A = rng.standard_normal((p, p)); cov = A @ A.T + np.eye(p)
    n = 30000
    tuner = o.GAMCTuner(o.VanillaMCTuner(), o.VanillaMCTuner(), o.AcceptanceRateMCTuner(0.35))
    sam = o.GAMCOracle(_Gauss(mean, cov), mean + 0.1, n, 3000, update=update, driftstep=1.0, minorscale=0.001, c=0.01, tuner=tuner)
    v, a, k = sam.run(o.RandomTape(n, p, seed=6))
    ess = o.ess(v)
    se = np.sqrt(np.diag(cov) / ess)
    assert np.all(np.abs(v.mean(1) - mean) < 5 * se)
    assert relerr(np.cov(v), cov) < 0.25
    assert 0.1 < a.mean() < 0.95


def test_first_t0_iterations_are_smmala_and_reeval_semantics():
    X, y, tt = o.synth_logistic(1, 100, 3)
    tgt = o.LogisticTarget(X, y, 100.0)
    sam = o.GAMCOracle(tgt, tt, 50, 0, update="am_only_update!", driftstep=0.02, minorscale=0.001, c=0.01)
    v, a, kinds = sam.run(o.RandomTape(50, 3, seed=1))
    assert list(kinds[:3]) == [1, 1, 1] and not kinds[3:].any()       # t0 = 3
    assert sam.updatetensorcount == 3
    # init evaluates once, the first iterate! re-evaluates at theta0 (pastupdatetensor starts false), then one per step
    assert sam.n_tensor_evals == 1 + 1 + 3 and sam.n_logtarget_evals == 47


def test_tuner_stops_one_period_before_burnin_and_step_rule():
    X, y, tt = o.synth_logistic(2, 150, 3)
    tuner = o.GAMCTuner(o.VanillaMCTuner(), o.VanillaMCTuner(), o.AcceptanceRateMCTuner(0.35, period=50))
    sam = o.GAMCOracle(o.LogisticTarget(X, y, 100.0), tt, 600, 300, update="rand_update!", driftstep=0.02, minorscale=0.001, c=0.01, tuner=tuner)
    steps = []
    tape = o.RandomTape(600, 3, seed=2)
    for i in range(600):
        sam.iterate(tape)
        steps.append(sam.tune.totaltune.step)
    changes = [i + 1 for i in range(1, 600) if steps[i] != steps[i - 1]]
    assert changes and all(c % 50 == 0 for c in changes)
    assert max(changes) <= 300 - 50 + 50     # totproposed seeded with the period: last tuning at totproposed <= burnin
    assert sam.tune.totaltune.totproposed >= 300
    assert o.logistic_rate_score(0.0) == 1.0 and o.logistic_rate_score(0.1) == pytest.approx(2 / (1 + math.exp(-0.7)))


def test_ess_on_ar1():
    rng = np.random.default_rng(7)
    rho, n = 0.6, 40000
    x = np.empty(n); x[0] = 0
    e = rng.standard_normal(n)
    for i in range(1, n):
        x[i] = rho * x[i - 1] + e[i]
    est = o.ess(x, maxlag=400)[0]
    assert est == pytest.approx(n * (1 - rho) / (1 + rho), rel=0.12)
    assert o.acceptance(np.array([[True, False, True, True]])) == 0.75


def test_notposdef_is_raised_like_chol():
    with pytest.raises(osamp.NotPosDef):
        osamp._chol_lower(np.array([[1.0, 2.0], [2.0, 1.0]]))


def test_smmala_ratio_expansion_identity():
    """The device evaluates the proposal-side quadratic form of the SMMALA ratio (iterate/GAMC.jl:58-67) as
    a'G*a - eps a'grad* + eps^2/4 f*'grad* with a = theta - theta*, f* = G*^-1 grad* (DESIGN.md 3C); here the identity itself,
    against the literal (theta - mu*)' G* (theta - mu*), mu* = theta* + eps/2 f*, on random well- and ill-conditioned metrics."""
    rng = np.random.default_rng(0)
    for p, cond in ((5, 1e1), (40, 1e4), (120, 1e7)):
        Q, _ = np.linalg.qr(rng.standard_normal((p, p)))
        G = (Q * np.geomspace(1.0, cond, p)) @ Q.T
        G = 0.5 * (G + G.T)
        theta, theta_p, grad = rng.standard_normal(p), rng.standard_normal(p), rng.standard_normal(p)
        eps = 0.02
        f = np.linalg.solve(G, grad)
        d2 = theta - (theta_p + 0.5 * eps * f)
        literal = d2 @ G @ d2
        a = theta - theta_p
        expanded = a @ G @ a - eps * (a @ grad) + 0.25 * eps * eps * (f @ grad)
        assert abs(expanded - literal) <= 1e-12 * abs(literal) * max(1.0, cond * 1e-6)
