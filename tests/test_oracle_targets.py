"""t-distribution and radial-velocity restatements against SciPy and the reference's shipped data fixtures."""
import json
import math
import os

import numpy as np
import pytest
from scipy import stats

from oracle import targets as T
from util import relerr

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_t_target_matches_scipy():
    tgt = T.TTarget(n=20, nu=30.0)
    rng = np.random.default_rng(0)
    for _ in range(5):
        x = 2.0 * rng.standard_normal(20)
        ref = stats.multivariate_t(loc=np.zeros(20), shape=tgt.Sigma_t, df=30.0).logpdf(x)
        assert tgt.logtarget(x) == pytest.approx(ref, rel=1e-12)


def test_t_target_derivatives():
    tgt = T.TTarget(n=8, nu=30.0)
    x = np.random.default_rng(1).standard_normal(8) * 3.0
    lt, g, G = tgt.uptotensorlogtarget(x)
    h = 1e-6
    fd = np.array([(tgt.logtarget(x + h * e) - tgt.logtarget(x - h * e)) / (2 * h) for e in np.eye(8)])
    assert relerr(g, fd) < 1e-7
    H = np.array([(tgt.uptotensorlogtarget(x + h * e)[1] - tgt.uptotensorlogtarget(x - h * e)[1]) / (2 * h) for e in np.eye(8)])
    assert relerr(G, -H) < 1e-6


def test_t_hessian_indefinite_far_out_and_softabs_fixes_it():
    tgt = T.TTarget(n=6, nu=30.0)
    x = 20.0 * np.ones(6)                      # q > nu: one negative eigenvalue of -H
    _, _, G = tgt.uptotensorlogtarget(x)
    assert np.linalg.eigvalsh(G).min() < 0
    S = T.softabs(G, 1000.0)
    lam = np.linalg.eigvalsh(S)
    assert lam.min() > 0
    assert np.allclose(np.sort(lam), np.sort(np.abs(np.linalg.eigvalsh(G))), rtol=1e-2, atol=2e-3)
    assert relerr(T.softabs(np.zeros((3, 3)), 1000.0), np.eye(3) / 1000.0) < 1e-12


def test_kepler_solver():
    M = np.linspace(-10, 20, 301)
    for ecc in (0.0, 0.2, 0.7, 0.95):
        E = T.kepler_laguerre(M, ecc)
        assert np.max(np.abs(E - ecc * np.sin(E) - np.mod(M, 2 * math.pi))) < 1e-8


@pytest.mark.parametrize("name,loglik,logprior,logtarget,chisq", [
    ("one_planet", -105.7088855237, -15.0927775295, -120.8016630532, 50.2092),
    ("two_planets", -100.4558566321, -27.5429285046, -127.9987851367, 39.7031),
])
def test_rv_fixtures(name, loglik, logprior, logtarget, chisq):
    """Data shipped with the reference (examples/radial_velocity/data/*.csv) at the truth vectors of utils_ex.jl:48-68
    (also listed in one_planet/meanplot.r:19, two_planets/meanplot.r:19-31); values recorded in SURVEY.md section 4."""
    with open(os.path.join(GOLD, "rv_fixtures.json")) as f:
        d = json.load(f)[name]
    m = T.RVModel(d["t"], d["rv"], d["sigma"])
    th = np.array(d["theta_true"])
    assert m.loglikelihood(th) == pytest.approx(loglik, abs=5e-10)
    assert m.logprior(th) == pytest.approx(logprior, abs=5e-10)
    assert m.logtarget(th) == pytest.approx(logtarget, abs=5e-10)
    assert m.loglikelihood(th) == pytest.approx(d["loglik"], rel=1e-14)
    res = (m.model_rv(th) - np.array(d["rv"])) / np.array(d["sigma"])
    assert float(res @ res) == pytest.approx(chisq, abs=1e-4)
    assert abs(res.std() - 1.0) < 0.2          # residuals consistent with sigma = 2 noise


def test_rv_truth_vectors_and_validity():
    assert np.allclose(T.make_param_true_ex1(), [3.931826, 3.044522, 0.141421, 0.141421, 1.570796, 1.0], atol=1e-6)
    assert len(T.make_param_true_ex2()) == 11
    th = T.make_param_true_ex1().copy()
    th[2], th[3] = 0.8, 0.7                      # h^2 + k^2 >= 1  => -Inf (stat_model.jl:21)
    m = T.RVModel([0.0, 1.0], [0.0, 0.0], [2.0, 2.0])
    assert m.loglikelihood(th) == -math.inf and m.logtarget(th) == -math.inf


def test_rv_dual_number_derivatives_match_finite_differences():
    """Exact (dual-number) gradient and tensor of the RV log-target against central differences of the plain restatement."""
    with open(os.path.join(GOLD, "rv_fixtures.json")) as f:
        d = json.load(f)["one_planet"]
    m = T.RVModel(d["t"], d["rv"], d["sigma"])
    th = np.array(d["theta_true"]) + 0.01 * np.array([0.0001, 0.01, 0.05, 0.05, 0.03, 0.3])
    lt, g, G = m.uptotensorlogtarget(th)
    assert lt == pytest.approx(m.logtarget(th), rel=1e-13)
    h = 1e-6
    fd = np.array([(m.logtarget(th + h * e) - m.logtarget(th - h * e)) / (2 * h) for e in np.eye(6)])
    assert relerr(g, fd) < 1e-6
    H = np.array([(m.uptotensorlogtarget(th + h * e)[1] - m.uptotensorlogtarget(th - h * e)[1]) / (2 * h) for e in np.eye(6)])
    assert relerr(G, -H) < 1e-6
    assert np.array_equal(G, G.T)
