"""The structured algebra of the large-dimension t-target path (tests/bigd_ref.py = what csrc/bigd.cuh computes) against the
dense oracle: tridiagonal inverse scale matrix, low-rank softabs correction from the secular equation, product-form reverse
Cholesky factor, and the quantities the SMMALA step takes from it.  CPU only."""
import math

import numpy as np
import pytest

import bigd_ref as R
from util import relerr


def _dense(n, x, alpha=1000.0):
    from oracle import targets as T
    tgt = T.TTarget(n, 30.0, 0.9)
    lt, g, H = tgt.uptotensorlogtarget(x)
    G = T.softabs(H, alpha)
    return tgt, lt, g, H, G


@pytest.mark.parametrize("n,scale", [(20, 2.0), (48, 2.0), (48, 0.7), (64, 1.0), (130, 2.0), (130, 1.0)])
def test_structured_point_matches_dense_oracle(n, scale):
    rng = np.random.default_rng(n)
    x = scale * rng.standard_normal(n)
    td, te, ld, le, logconst = R.tri_constants(n)
    tgt, lt, g, H, G = _dense(n, x)
    # the inverse scale matrix is exactly tridiagonal
    Tm = np.diag(td) + np.diag(te, 1) + np.diag(te, -1)
    assert relerr(Tm, tgt.Sinv) < 1e-11 and abs(logconst - tgt.const) < 1e-9 * abs(tgt.const)
    pt = R.structured_point(td, te, ld, le, logconst, 30.0, 1000.0, x)
    assert abs(pt["lt"] - lt) < 1e-11 * abs(lt) and relerr(pt["grad"], g) < 1e-11
    # secular roots = lowest eigenvalues of H / a
    lam = np.linalg.eigvalsh(H)
    r = len(pt["roots"])
    assert r >= 1 and relerr(pt["a"] * pt["roots"], lam[:r]) < 1e-9
    assert r == n or pt["a"] * 0 + lam[r] >= 21.0 / 1000.0 - 1e-12
    # G = a T + V S V'
    Gs = pt["a"] * Tm + (pt["V"] * pt["S"]) @ pt["V"].T
    assert relerr(Gs, G) < 1e-8
    d = rng.standard_normal(n)
    assert relerr(R.g_times(td, te, pt, d), G @ d) < 1e-8
    # reverse Cholesky factor: chol(inv(G)) z, G^-1 grad, logdet
    Ginv = np.linalg.inv(G)
    L = np.linalg.cholesky(0.5 * (Ginv + Ginv.T))
    z = rng.standard_normal(n)
    assert relerr(R.chol_inv_times(pt, z), L @ z) < 1e-7
    assert relerr(pt["fterm"], Ginv @ g) < 1e-7
    _, logdetG = np.linalg.slogdet(G)
    assert abs(2.0 * pt["sumlog"] - logdetG) < 1e-8 * max(1.0, abs(logdetG))
    # dense hand-over: chol(inv(G)) = reversal of Lm^-1
    X = pt["F"].dense_inverse()
    L0 = X[::-1, ::-1].T
    assert relerr(L0, L) < 1e-7


def test_many_low_eigenvalues_use_several_chunks():
    """Far from the mode a = (nu+n)/(nu+q) is small and many eigenvalues of the tensor fall below 21/alpha: the correction has
    more than 32 columns, i.e. several product factors."""
    n = 400
    rng = np.random.default_rng(5)
    x = 2.0 * rng.standard_normal(n)
    td, te, ld, le, logconst = R.tri_constants(n)
    tgt, lt, g, H, G = _dense(n, x)
    pt = R.structured_point(td, te, ld, le, logconst, 30.0, 1000.0, x)
    assert len(pt["F"].chunks) >= 2
    Ginv = np.linalg.inv(G)
    L = np.linalg.cholesky(0.5 * (Ginv + Ginv.T))
    z = rng.standard_normal(n)
    assert relerr(R.chol_inv_times(pt, z), L @ z) < 1e-6
    assert relerr(pt["fterm"], Ginv @ g) < 1e-6


def test_softabs_gap_formula():
    for lam in (-3.0, -1e-3, -1e-12, 0.0, 1e-12, 1e-4, 5e-3, 0.02, 0.5):
        want = (1.0 / 1000.0 if abs(1000.0 * lam) < 1e-8 else lam / math.tanh(1000.0 * lam)) - lam
        got = R.softabs_gap(lam, 1000.0)
        assert abs(got - want) <= 1e-12 * max(abs(want), 1e-300) + 1e-18
