"""Oracle: Bayesian logistic-regression target (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates `examples/logistic_regression/src/GAMC/simulations.jl:25-37` of the reference:

    ploglikelihood   :25-28   l(theta) = (X theta).y - sum(log(1+exp(X theta)))
    plogprior        :30      -0.5*(theta.theta/lambda + npars*log(2*pi*lambda))
    pgradlogtarget   :32      X'(y - 1/(1+exp(-X theta))) - theta/lambda
    ptensorlogtarget :34-37   X' diag(r(1-r)) X + I/lambda

X is N x p (any memory order; the product path uses Julia's column-major), y is N (0/1 as float64),
lambda is the prior *variance* (100 in the example, `:69`).

Documented deviation: `ploglikelihood` uses a numerically stable softplus; `ploglikelihood_naive` is the
literal reference formula (overflows for eta > ~709).  They agree to rounding for |eta| < 700.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "ploglikelihood", "ploglikelihood_naive", "plogprior", "pgradlogtarget", "ptensorlogtarget",
    "logistic_partials", "LogisticTarget", "synth_logistic", "philox4x32", "synth_X_block",
]


def _sigmoid(eta):
    """1/(1+exp(-eta)) evaluated without overflow (reference: `1./(1+exp.(-v[2]*p))`, simulations.jl:32,35)."""
    e = np.exp(-np.abs(eta))
    return np.where(eta >= 0, 1.0 / (1.0 + e), e / (1.0 + e))


def _softplus(eta):
    """log(1+exp(eta)) evaluated without overflow (reference: `log.(1+exp.(Xp))`, simulations.jl:27)."""
    return np.maximum(eta, 0.0) + np.log1p(np.exp(-np.abs(eta)))


def ploglikelihood_naive(theta, X, y):
    """Literal restatement of simulations.jl:25-28."""
    Xp = X @ theta
    return float(np.dot(Xp, y) - np.sum(np.log(1 + np.exp(Xp))))


def ploglikelihood(theta, X, y):
    """simulations.jl:25-28 with a stable softplus; summed with math.fsum-grade accuracy (pairwise in NumPy)."""
    Xp = X @ theta
    return float(np.sum(Xp * y - _softplus(Xp)))


def plogprior(theta, lam):
    """simulations.jl:30 (npars = len(theta))."""
    p = theta.shape[0]
    return float(-0.5 * (np.dot(theta, theta) / lam + p * np.log(2 * np.pi * lam)))


def pgradlogtarget(theta, X, y, lam):
    """simulations.jl:32."""
    return X.T @ (y - _sigmoid(X @ theta)) - theta / lam


def ptensorlogtarget(theta, X, lam):
    """simulations.jl:34-37 (expected Fisher information + prior precision)."""
    r = _sigmoid(X @ theta)
    w = r * (1 - r)
    G = (X * w[:, None]).T @ X
    G = 0.5 * (G + G.T)  # exact symmetry (the reference's gemm result is symmetric only to rounding)
    return G + np.eye(theta.shape[0]) / lam


def logistic_partials(theta, X, y, want_grad=True, want_tensor=True):
    """Likelihood-only partial sums over a row shard: (l_lik, g_lik, G_lik) with no prior terms.

    This is what one rank contributes to the all-reduce in the row-sharded layout: prior
    terms are added once after the reduction.  One fused pass (one X@theta) -- the arithmetic the CUDA
    kernels implement.
    """
    eta = X @ theta
    ll = float(np.sum(eta * y - _softplus(eta)))
    g = G = None
    if want_grad or want_tensor:
        r = _sigmoid(eta)
        if want_grad:
            g = X.T @ (y - r)
        if want_tensor:
            w = r * (1 - r)
            G = (X * w[:, None]).T @ X
            G = 0.5 * (G + G.T)
    return ll, g, G


class LogisticTarget:
    """The Klara closures `parameter.logtarget!` / `parameter.uptotensorlogtarget!` for this model
    (call sites `src/samplers/iterate/GAMC.jl:20,37,148`, `src/samplers/GAMC.jl:163`)."""

    def __init__(self, X, y, lam=100.0, reference_shaped=False):
        self.X = np.asarray(X, dtype=np.float64)
        self.y = np.asarray(y, dtype=np.float64)
        self.lam = float(lam)
        self.p = self.X.shape[1]
        # reference_shaped=True recomputes X@theta in each of the three callbacks and builds the N x p
        # temporary + full gemm, exactly as the Julia callbacks do (used for the CPU baseline timing).
        self.reference_shaped = reference_shaped

    def logtarget(self, theta):
        return ploglikelihood(theta, self.X, self.y) + plogprior(theta, self.lam)

    def uptogradlogtarget(self, theta):
        """Klara closure `parameter.uptogradlogtarget!` (what MALA calls): log-target and gradient."""
        if self.reference_shaped:
            return (ploglikelihood(theta, self.X, self.y) + plogprior(theta, self.lam), pgradlogtarget(theta, self.X, self.y, self.lam))
        ll, g, _ = logistic_partials(theta, self.X, self.y, want_tensor=False)
        return ll + plogprior(theta, self.lam), g - theta / self.lam

    def uptotensorlogtarget(self, theta):
        if self.reference_shaped:
            lt = ploglikelihood(theta, self.X, self.y) + plogprior(theta, self.lam)
            g = pgradlogtarget(theta, self.X, self.y, self.lam)
            G = ptensorlogtarget(theta, self.X, self.lam)
            return lt, g, G
        ll, g, G = logistic_partials(theta, self.X, self.y)
        lt = ll + plogprior(theta, self.lam)
        return lt, g - theta / self.lam, G + np.eye(self.p) / self.lam


# ----------------------------------------------------------------------------------------------------
# Synthetic data: counter-based, keyed by GLOBAL (row, col) so every sharding sees the
# same matrix.  Bit-exact twin of the device generator in gamcsampler.jl_b200/csrc/synth.cuh.
# ----------------------------------------------------------------------------------------------------
_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = 0x9E3779B9
_PHILOX_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    """Philox-4x32-10 (Salmon et al. 2011).  Inputs: uint32 arrays (broadcastable); returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64) & _MASK32
    c1 = np.asarray(c1, dtype=np.uint64) & _MASK32
    c2 = np.asarray(c2, dtype=np.uint64) & _MASK32
    c3 = np.asarray(c3, dtype=np.uint64) & _MASK32
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = _PHILOX_M0 * c0
        p1 = _PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + _PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


_IH_SCALE = np.sqrt(3.0) / 4294967296.0  # sqrt(3) / 2^32: Irwin-Hall(4) of U(0,1) has variance 1/3


def synth_X_block(seed, row0, nrows, p, col0=0):
    """X[row0:row0+nrows, col0:col0+p] of the synthetic design matrix.

    X[i,j] = (u0+u1+u2+u3 - 2^33) * sqrt(3)/2^32 with (u0..u3) = Philox4x32-10(ctr=(i_lo, i_hi, j, 0),
    key=(seed_lo, seed_hi)): a standardised Irwin-Hall(4) variate (mean 0, variance 1, support +-3.46) --
    pure integer arithmetic plus one exact int->double conversion and one multiply, hence bit-identical on
    CPU and GPU.  It stands in for the "standardised covariates" of simulations.jl:20.
    """
    rows = np.arange(row0, row0 + nrows, dtype=np.uint64)[:, None]
    cols = np.arange(col0, col0 + p, dtype=np.uint64)[None, :]
    u0, u1, u2, u3 = philox4x32(rows & _MASK32, rows >> np.uint64(32), cols, 0,
                                seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    s = u0.astype(np.int64) + u1.astype(np.int64) + u2.astype(np.int64) + u3.astype(np.int64) - (1 << 33)
    return s.astype(np.float64) * _IH_SCALE


def synth_uniform_rows(seed, row0, nrows, stream=1):
    """One U[0,1) double per global row (53-bit), counter (i_lo, i_hi, 0xFFFFFFFF, stream)."""
    rows = np.arange(row0, row0 + nrows, dtype=np.uint64)
    u0, u1, _, _ = philox4x32(rows & _MASK32, rows >> np.uint64(32), 0xFFFFFFFF, stream,
                              seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    bits = (u0.astype(np.uint64) << np.uint64(21)) | (u1.astype(np.uint64) >> np.uint64(11))  # 53 bits
    return bits.astype(np.float64) * (1.0 / 9007199254740992.0)


def synth_theta_true(seed, p):
    """theta_true[j] ~ N(0, 1/p) (host side, NumPy Generator(PCG64(seed)); passed to the device by value)."""
    return np.random.Generator(np.random.PCG64(seed)).standard_normal(p) / np.sqrt(p)


def synth_logistic(seed, N, p, row0=0, nrows=None, theta_true=None, order="F"):
    """Synthetic logistic-regression shard rows [row0, row0+nrows) of the global N x p problem.

    y[i] = 1{ u_i < sigma(x_i . theta_true) }.  Returns (X, y, theta_true).  eta is accumulated
    sequentially over columns j = 0..p-1 with a fused multiply-add free plain sum, as in the device
    generator; a borderline u_i within rounding of sigma(eta_i) could in principle flip a label between
    CPU and GPU (probability ~1e-16 per row) -- tests compare labels exactly and would flag it.
    """
    if nrows is None:
        nrows = N - row0
    if theta_true is None:
        theta_true = synth_theta_true(seed, p)
    X = np.empty((nrows, p), dtype=np.float64, order=order)
    blk = max(1, (1 << 22) // max(p, 1))
    for r in range(0, nrows, blk):
        n = min(blk, nrows - r)
        X[r:r + n, :] = synth_X_block(seed, row0 + r, n, p)
    eta = np.zeros(nrows)
    for j in range(p):  # sequential-in-j accumulation: the order the device generator uses
        eta += X[:, j] * theta_true[j]
    u = synth_uniform_rows(seed, row0, nrows)
    y = (u < _sigmoid(eta)).astype(np.float64)
    return X, y, theta_true
