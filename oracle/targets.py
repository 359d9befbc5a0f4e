"""Oracle: the other two example targets of the reference (TEST INFRASTRUCTURE -- see oracle/__init__.py).

  * multivariate-t target            examples/t_distribution/src/GAMC/simulations.jl:17-39
  * Keplerian radial-velocity target examples/radial_velocity/src/{kepler_eqn,rv_model_keplerian,rv_model_param,stat_model,utils_ex}.jl
  * softabs metric regulariser       Klara `softabs(H, alpha)` used at t_distribution/.../simulations.jl:39

The reference obtains gradient and Hessian of these targets by forward-mode autodiff (Klara `DiffOptions(mode=:forward,
order=2)`), tensor = -Hessian; here the t target has closed forms, the RV target uses complex-step / finite differences
of this restatement for validation only.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import gammaln

__all__ = ["toeplitz_cov", "TTarget", "softabs", "RVModel", "make_param_true_ex1", "make_param_true_ex2", "kepler_laguerre"]


# ------------------------------------------------------------------ multivariate t
def toeplitz_cov(n, c=0.9):
    """`C(n, c)` of t_distribution/.../simulations.jl:17-21: Sigma_ij = c^|i-j|."""
    idx = np.arange(n)
    return c ** np.abs(idx[:, None] - idx[None, :])


class TTarget:
    """log MvTDist(nu, 0, Sigma_t)(x), Sigma_t = (nu-2) Sigma / nu  (simulations.jl:23-31).

    Distributions.jl `logpdf(MvTDist(nu, mu, S), x)` =
        lgamma((nu+n)/2) - lgamma(nu/2) - (n/2) log(nu pi) - (1/2) logdet S - ((nu+n)/2) log1p(d' S^-1 d / nu).
    Closed-form derivatives (q = x' S^-1 x, v = S^-1 x):
        grad = -(nu+n) v / (nu+q),   -Hessian = (nu+n)/(nu+q) [S^-1 - 2 v v'/(nu+q)].
    """

    def __init__(self, n=20, nu=30.0, c=0.9):
        self.n, self.nu = int(n), float(nu)
        self.Sigma = toeplitz_cov(n, c)
        self.Sigma_t = (nu - 2.0) * self.Sigma / nu
        self.Sinv = np.linalg.inv(self.Sigma_t)
        _, self.logdet = np.linalg.slogdet(self.Sigma_t)
        self.const = (gammaln((nu + n) / 2.0) - gammaln(nu / 2.0) - 0.5 * n * math.log(nu * math.pi) - 0.5 * self.logdet)

    def logtarget(self, x):
        q = float(x @ self.Sinv @ x)
        return float(self.const - 0.5 * (self.nu + self.n) * math.log1p(q / self.nu))

    def uptotensorlogtarget(self, x):
        v = self.Sinv @ x
        q = float(x @ v)
        a = (self.nu + self.n) / (self.nu + q)
        lt = float(self.const - 0.5 * (self.nu + self.n) * math.log1p(q / self.nu))
        return lt, -a * v, a * (self.Sinv - 2.0 * np.outer(v, v) / (self.nu + q))


def softabs(H, alpha=1000.0):
    """Klara `softabs(H, alpha)`: Q diag(lambda coth(alpha lambda)) Q' from the symmetric eigendecomposition
    (-> |lambda| for |lambda| >> 1/alpha, -> 1/alpha at lambda = 0)."""
    lam, Q = np.linalg.eigh(0.5 * (H + H.T))
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        s = np.where(np.abs(alpha * lam) < 1e-8, 1.0 / alpha, lam / np.tanh(alpha * lam))
    return (Q * s) @ Q.T


# ------------------------------------------------------------------ Keplerian radial velocity
def kepler_laguerre(mean_anom, ecc, tol=1e-8, max_its=200):
    """`calc_ecc_anom_itterative_laguerre` (kepler_eqn.jl:43-56) with the Danby initial guess (:18-23) and the
    Laguerre update with n = 5 (:30-40).  Vectorised over mean_anom."""
    M = np.mod(np.asarray(mean_anom, dtype=np.float64), 2.0 * math.pi)
    k = 0.85
    E = np.where(M < math.pi, M + k * ecc, M - k * ecc)
    n = 5.0
    for _ in range(max_its):
        es, ec = ecc * np.sin(E), ecc * np.cos(E)
        F = (E - es) - M
        Fp = 1.0 - ec
        Fpp = es
        root = np.sqrt(np.abs((n - 1) * ((n - 1) * Fp * Fp - n * F * Fpp)))
        denom = np.where(Fp > 0, Fp + root, Fp - root)
        E_new = E - n * F / denom
        done = np.all(np.abs(E_new - E) < tol)
        E = E_new
        if done:
            break
    else:  # pragma: no cover
        raise AssertionError("Kepler solver did not converge")
    return E


def make_param_true_ex1():
    """utils_ex.jl:48-57 packed by rv_model_param.jl:19-44: [log1p(P), log1p(K), e cos w, e sin w, w + M0, C]."""
    P, K, e, w, M0, C = 50.0, 20.0, 0.2, math.pi / 4, math.pi / 4, 1.0
    return np.array([math.log1p(P), math.log1p(K), e * math.cos(w), e * math.sin(w), w + M0, C])


def make_param_true_ex2():
    """utils_ex.jl:59-68."""
    out = []
    for P, K in ((40.0, 30.0), (80.8, 30.0)):
        e, w, M0 = 0.2, math.pi / 4, math.pi / 4
        out += [math.log1p(P), math.log1p(K), e * math.cos(w), e * math.sin(w), w + M0]
    return np.array(out + [0.0])


class RVModel:
    """plogtarget = ploglikelihood + plogprior of stat_model.jl:18-63 (num_jitters = 0, rv_model_param.jl:5)."""

    def __init__(self, times, obs, sigma_obs):
        self.t = np.asarray(times, dtype=np.float64)
        self.o = np.asarray(obs, dtype=np.float64)
        self.s = np.asarray(sigma_obs, dtype=np.float64)

    @staticmethod
    def num_planets(theta):
        return len(theta) // 5

    @staticmethod
    def is_valid(theta):
        """rv_model_param.jl:83-96: every planet needs h^2 + k^2 < 1."""
        for pl in range(len(theta) // 5):
            h, k = theta[5 * pl + 2], theta[5 * pl + 3]
            if not (h * h + k * k < 1.0):
                return False
        return True

    def model_rv(self, theta):
        """calc_model_rv (rv_model_keplerian.jl:9-51): sum of Keplerian velocities + offset."""
        npl = self.num_planets(theta)
        z = np.zeros_like(self.t)
        for pl in range(npl):
            P = math.exp(theta[5 * pl]) - 1.0            # extract_period      rv_model_param.jl:45
            K = math.exp(theta[5 * pl + 1]) - 1.0        # extract_amplitude   :46
            h, k = theta[5 * pl + 2], theta[5 * pl + 3]  # e cos w, e sin w    :47-48
            M0 = theta[5 * pl + 4] - math.atan2(k, h)    # extract_PKhkM       :73-80
            ecc = math.sqrt(h * h + k * k)
            w = math.atan2(k, h)
            n = 2.0 * math.pi / P
            M = self.t * n - M0
            lam = w + M
            E = kepler_laguerre(M, ecc)
            assert 0.0 <= ecc < 1.0
            j = math.sqrt((1.0 - ecc) * (1.0 + ecc))
            pp_, q = (0.0, 0.0) if ecc == 0.0 else (ecc * np.sin(E), ecc * np.cos(E))
            a = K / (n / math.sqrt((1.0 - ecc) * (1.0 + ecc)))
            z = z + a * n / (1.0 - q) * (np.cos(lam + pp_) - k * q / (1.0 + j))
        return z + theta[5 * npl]

    def loglikelihood(self, theta):
        if not self.is_valid(theta):
            return -math.inf
        model = self.model_rv(theta)
        s2 = self.s ** 2
        chisq = float(np.sum((model - self.o) ** 2 / s2))
        lognorm = -0.5 * len(self.t) * math.log(2 * math.pi) - 0.5 * float(np.sum(np.log(s2)))
        return -0.5 * chisq + lognorm

    def logprior(self, theta):
        """stat_model.jl:39-61: modified Jeffreys priors on P and K (P0 = K0 = 1, max 10000)."""
        if not self.is_valid(theta):
            return -math.inf
        logp = -2.0 * math.log(2 * math.pi)
        for pl in range(self.num_planets(theta)):
            P = math.exp(theta[5 * pl]) - 1.0
            K = math.exp(theta[5 * pl + 1]) - 1.0
            logp += -math.log((1 + P) * math.log1p(10000.0) * (1 + K) * math.log1p(10000.0))
        return logp

    def logtarget(self, theta):
        return self.loglikelihood(theta) + self.logprior(theta)
