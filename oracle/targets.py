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


# ------------------------------------------------------------------ second-order forward-mode duals (vectorised over epochs)
class Dual2:
    """value v (N,), gradient g (N, m), Hessian h (N, m, m) of a scalar function of m variables, per epoch.
    Stands in for the reference's ForwardDiff sweep (Klara `DiffOptions(mode=:forward, order=2)`,
    examples/radial_velocity/src/one_planet/GAMC/simulations.jl:37)."""

    def __init__(self, v, g, h):
        self.v, self.g, self.h = v, g, h

    @staticmethod
    def const(c, N, m):
        return Dual2(np.full(N, float(c)) if np.isscalar(c) else np.asarray(c, dtype=np.float64), np.zeros((N, m)), np.zeros((N, m, m)))

    @staticmethod
    def var(c, k, N, m):
        d = Dual2.const(c, N, m)
        d.g[:, k] = 1.0
        return d

    def _lift(self, o):
        return o if isinstance(o, Dual2) else Dual2.const(o, self.v.shape[0], self.g.shape[1])

    def unary(self, f0, f1, f2):
        gg = self.g[:, :, None] * self.g[:, None, :]
        return Dual2(f0, f1[:, None] * self.g, f1[:, None, None] * self.h + f2[:, None, None] * gg)

    def __add__(self, o):
        o = self._lift(o)
        return Dual2(self.v + o.v, self.g + o.g, self.h + o.h)
    __radd__ = __add__

    def __neg__(self):
        return Dual2(-self.v, -self.g, -self.h)

    def __sub__(self, o):
        return self + (-self._lift(o))

    def __rsub__(self, o):
        return self._lift(o) - self

    def __mul__(self, o):
        o = self._lift(o)
        cross = self.g[:, :, None] * o.g[:, None, :]
        return Dual2(self.v * o.v, self.v[:, None] * o.g + o.v[:, None] * self.g,
                     self.v[:, None, None] * o.h + o.v[:, None, None] * self.h + cross + cross.transpose(0, 2, 1))
    __rmul__ = __mul__

    def recip(self):
        r = 1.0 / self.v
        return self.unary(r, -r * r, 2 * r ** 3)

    def __truediv__(self, o):
        return self * self._lift(o).recip()

    def __rtruediv__(self, o):
        return self._lift(o) * self.recip()


def d_sin(x):
    return x.unary(np.sin(x.v), np.cos(x.v), -np.sin(x.v))


def d_cos(x):
    return x.unary(np.cos(x.v), -np.sin(x.v), -np.cos(x.v))


def d_sqrt(x):
    s = np.sqrt(x.v)
    return x.unary(s, 0.5 / s, -0.25 / (s * x.v))


def d_exp(x):
    e = np.exp(x.v)
    return x.unary(e, e, e)


def d_atan2(y, x):
    r2 = x * x + y * y
    ay, ax = x / r2, -(y / r2)
    out = Dual2(np.arctan2(y.v, x.v), ay.v[:, None] * y.g + ax.v[:, None] * x.g,
                ay.v[:, None, None] * y.h + ax.v[:, None, None] * x.h + y.g[:, :, None] * ay.g[:, None, :] + x.g[:, :, None] * ax.g[:, None, :])
    out.h = 0.5 * (out.h + out.h.transpose(0, 2, 1))
    return out


def _rv_uptotensor(self, theta):
    """(logtarget, gradient, tensor = -Hessian) of the RV target with exact derivatives: forward-mode duals over the whole
    expression of rv_model_keplerian.jl:9-51; the eccentric anomaly is solved in plain floating point and then refined by
    two dual Newton steps (implicit differentiation of Kepler's equation; the reference differentiates through its Laguerre
    iterations, which converge to the same derivatives)."""
    theta = np.asarray(theta, dtype=np.float64)
    m = len(theta)
    if not self.is_valid(theta):
        return -math.inf, np.full(m, np.nan), np.full((m, m), np.nan)
    N = self.t.shape[0]
    npl = self.num_planets(theta)
    z = Dual2.var(theta[5 * npl], 5 * npl, N, m)                  # offset C
    for pl in range(npl):
        b = 5 * pl
        P = d_exp(Dual2.var(theta[b], b, N, m)) - 1.0
        K = d_exp(Dual2.var(theta[b + 1], b + 1, N, m)) - 1.0
        h = Dual2.var(theta[b + 2], b + 2, N, m)
        k = Dual2.var(theta[b + 3], b + 3, N, m)
        wm = Dual2.var(theta[b + 4], b + 4, N, m)
        ecc = d_sqrt(h * h + k * k)
        w = d_atan2(k, h)
        n = (2.0 * math.pi) / P
        M = n * Dual2.const(self.t, N, m) - wm + w                # M = t n - M0, M0 = wpM0 - w
        lam = w + M
        E0 = kepler_laguerre(M.v, float(ecc.v[0]), tol=1e-14)
        Mred = Dual2(np.mod(M.v, 2.0 * math.pi), M.g, M.h)
        E = Dual2.const(E0, N, m)
        for _ in range(2):
            E = E - (E - ecc * d_sin(E) - Mred) / (1.0 - ecc * d_cos(E))
        pp_, q = ecc * d_sin(E), ecc * d_cos(E)
        j = d_sqrt((1.0 - ecc) * (1.0 + ecc))
        z = z + K * j / (1.0 - q) * (d_cos(lam + pp_) - k * q / (1.0 + j))
    s2 = self.s ** 2
    r = z - Dual2.const(self.o, N, m)
    chi = r * r
    ll = -0.5 * float(np.sum(chi.v / s2)) - 0.5 * N * math.log(2 * math.pi) - 0.5 * float(np.sum(np.log(s2)))
    g = -0.5 * np.sum(chi.g / s2[:, None], axis=0)
    H = -0.5 * np.sum(chi.h / s2[:, None, None], axis=0)
    lp = self.logprior(theta)
    for pl in range(npl):                                          # log prior = const - sum(theta_P + theta_K): linear
        g[5 * pl] -= 1.0
        g[5 * pl + 1] -= 1.0
    G = -H
    return ll + lp, g, 0.5 * (G + G.T)


RVModel.uptotensorlogtarget = _rv_uptotensor
