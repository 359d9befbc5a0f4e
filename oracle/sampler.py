"""Oracle: the GAMC transition (AM <-> SMMALA), schedules and tuners (TEST INFRASTRUCTURE).

Restates `src/samplers/iterate/GAMC.jl:1-293` (one transition), `src/samplers/GAMC.jl:157-228` (init),
`:180-193` (GMM proposal), `:264-337` + `src/stats/schedule.jl:1-7` (schedules) and
`src/tuners/GAMCTuner.jl:1-23` of the reference, literally: explicit `inv`, `cholesky(inv)`, `slogdet`,
dense mat-vecs -- the reference's own operation sequence, so that the CUDA path (which uses one reverse
Cholesky factorisation instead) is checked against what the reference computes, not against itself.

Third-party semantics (Klara.jl 0.9.3, Distributions.jl 0.12.1 -- not under /root/reference, not runnable
here) are restated from their published algorithms and are DEFINED by this file; see SURVEY.md App. C.

Random numbers: the reference consumes Julia MersenneTwister draws (`randn(p)`, `rand()`), which cannot be
reproduced outside Julia.  The oracle and the device path both read a `RandomTape` with fixed slots per
iteration t = 1..nsteps: u_sched[t], u_mix[t], z[t, :], u_acc[t].  Deviation from the reference: the
accept uniform is always *available* on the tape (the reference only draws it if ratio <= 0,
`iterate/GAMC.jl:69,158`); decisions are identical.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

__all__ = [
    "MALAOracle", "RandomTape", "BasicMCTune", "VanillaMCTuner", "AcceptanceRateMCTuner", "GAMCTuner", "GAMCTune",
    "logistic_rate_score", "covariance_update", "recursive_mean", "gmm_logpdf", "GAMCOracle",
    "exp_decay", "lin_decay", "pow_decay", "quad_decay", "SCHEDULES", "schedule_probability",
]


# ------------------------------------------------------------------ schedules (src/stats/schedule.jl:1-7)
def exp_decay(i, tot, a=10.0, b=0.0):
    return (1 - b) * math.exp(-a * i / tot) + b


def lin_decay(i, tot, a=1e-3, b=0.0):
    return (1 - b) * (1 / (1 + a * i)) + b


def pow_decay(i, tot, a=1e-3, b=0.0):
    return (1 - b) * (a ** (i / tot)) + b


def quad_decay(i, tot, a=1e-3, b=0.0):
    return (1 - b) * (1 / (1 + a * i * i)) + b


def schedule_probability(name, i, tot, params=()):
    """Probability (or 0/1 decision) that iteration i runs the SMMALA branch.

    `src/samplers/GAMC.jl:264-337`.  Defaults of the rand_* wrappers differ from the decay functions'
    own defaults (a = 10 / 1e-2 / 1e-3 / 1e-5) -- kept as in the reference.  Returns (prob, is_random).
    """
    P = tuple(params)
    if name == "am_only_update!":
        return 0.0, False
    if name == "smmala_only_update!":
        return 1.0, False
    if name == "mod_update!":
        n = int(P[0]) if P else 10
        return (1.0 if i % n == 0 else 0.0), False
    if name == "rand_update!":
        return (P[0] if P else 0.5), True
    if name == "cos_update!":
        a, b, c = (list(P) + [10.0, 0.2, 0.2][len(P):])[:3]
        return b * math.cos(a * i * math.pi / tot) + c, True
    if name == "rand_exp_decay_update!":
        a, b = (list(P) + [10.0, 0.0][len(P):])[:2]
        return exp_decay(i, tot, a, b), True
    if name == "rand_lin_decay_update!":
        a, b = (list(P) + [1e-2, 0.0][len(P):])[:2]
        return lin_decay(i, tot, a, b), True
    if name == "rand_pow_decay_update!":
        a, b = (list(P) + [1e-3, 0.0][len(P):])[:2]
        return pow_decay(i, tot, a, b), True
    if name == "rand_quad_decay_update!":
        a, b = (list(P) + [1e-5, 0.0][len(P):])[:2]
        return quad_decay(i, tot, a, b), True
    raise ValueError(f"unknown schedule {name!r}")


SCHEDULES = (
    "am_only_update!", "smmala_only_update!", "mod_update!", "rand_update!", "cos_update!",
    "rand_exp_decay_update!", "rand_lin_decay_update!", "rand_pow_decay_update!", "rand_quad_decay_update!",
)


def schedule_decision(name, i, tot, params, u):
    """`rand(Bernoulli(prob))` == (u <= prob) [Distributions.jl: `rand() <= succprob(d)`]."""
    prob, is_random = schedule_probability(name, i, tot, params)
    if not is_random:
        return prob >= 1.0
    return bool(u <= prob)


# ------------------------------------------------------------------ tuners (Klara, restated)
def logistic_rate_score(x, k=7.0):
    """Klara `logistic_rate_score(x, k=7.)` = 2/(1+exp(-k x))."""
    return 2.0 / (1.0 + math.exp(-k * x))


@dataclass
class BasicMCTune:
    """Klara `BasicMCTune(step, accepted, proposed, totproposed)`; `rate` starts NaN."""
    step: float = 1.0
    accepted: int = 0
    proposed: int = 0
    totproposed: int = 0
    rate: float = float("nan")


@dataclass
class VanillaMCTuner:
    period: int = 100
    verbose: bool = False


@dataclass
class AcceptanceRateMCTuner:
    targetrate: float = 0.35
    score_k: float = 7.0
    period: int = 100
    verbose: bool = False

    def score(self, x):
        return logistic_rate_score(x, self.score_k)


@dataclass
class GAMCTuner:
    """`src/tuners/GAMCTuner.jl:15-23`: verbose is the OR of the three."""
    smmalatuner: VanillaMCTuner = field(default_factory=VanillaMCTuner)
    amtuner: VanillaMCTuner = field(default_factory=VanillaMCTuner)
    totaltuner: object = field(default_factory=VanillaMCTuner)

    @property
    def verbose(self):
        return bool(self.totaltuner.verbose or self.smmalatuner.verbose or self.amtuner.verbose)


@dataclass
class GAMCTune:
    """`src/tuners/GAMCTuner.jl:1-10`."""
    smmalatune: BasicMCTune
    amtune: BasicMCTune
    totaltune: BasicMCTune
    smmalafrequency: float = float("nan")
    amfrequency: float = float("nan")


# ------------------------------------------------------------------ Klara stats helpers (restated)
def recursive_mean(lastmean, k, x):
    """Klara `recursive_mean!(m, lastm, k, x)`: ((k-1)*lastm + x)/k  (iterate/GAMC.jl:195;
    same formula in examples/logistic_regression/src/meanplot.r:4-6)."""
    return ((k - 1) * lastmean + x) / k


def covariance_update(lastC, k, x, lastmean, secondlastmean):
    """Klara `covariance!(C, lastC, k, x, lastmean, secondlastmean)` (iterate/GAMC.jl:133-140):
    Haario et al. recursive empirical covariance
        C = ((k-1)*lastC + k*m2 m2' - (k+1)*m1 m1' + x x')/k,  m1 = lastmean, m2 = secondlastmean.
    """
    m1, m2 = lastmean, secondlastmean
    return ((k - 1) * lastC + k * np.outer(m2, m2) - (k + 1) * np.outer(m1, m1) + np.outer(x, x)) / k


def _hermitian_upper(A):
    """Julia `Hermitian(A)` (uplo = :U) materialised: mirror the upper triangle (iterate/GAMC.jl:142)."""
    U = np.triu(A)
    return U + np.triu(A, 1).T


def _mvnormal_logpdf_chol(x, mean, L):
    """log N(x; mean, L L') with L lower."""
    p = x.shape[0]
    d = x - mean
    s = _solve_lower(L, d)
    return -0.5 * (p * math.log(2 * math.pi) + 2.0 * float(np.sum(np.log(np.diag(L))))) - 0.5 * float(s @ s)


def _solve_lower(L, b):
    import scipy.linalg as sla
    return sla.solve_triangular(L, b, lower=True, check_finite=False)


def gmm_logpdf(x, mean, Lcore, sqrtminorscale, w):
    """`logpdf(MixtureModel([MvNormal(mean, eps*C), MvNormal(mean, sqrtminorscale)], w), x)`
    (`src/samplers/GAMC.jl:180-188`; Distributions: log-sum-exp over components).
    Lcore = lower Cholesky factor of eps*C; minor component covariance = sqrtminorscale^2 * I."""
    p = x.shape[0]
    lp1 = _mvnormal_logpdf_chol(x, mean, Lcore)
    d = x - mean
    lp2 = -0.5 * (p * math.log(2 * math.pi) + 2.0 * p * math.log(sqrtminorscale)) - 0.5 * float(d @ d) / sqrtminorscale ** 2
    terms = []
    if w[0] > 0:
        terms.append(math.log(w[0]) + lp1)
    if w[1] > 0:
        terms.append(math.log(w[1]) + lp2)
    m = max(terms)
    return m + math.log(sum(math.exp(t - m) for t in terms))


# ------------------------------------------------------------------ random tape
class RandomTape:
    """Host-drawn normals and uniforms with fixed per-iteration slots.

    Slot t (1-based iteration count) holds u_sched[t-1], u_mix[t-1], z[t-1, :], u_acc[t-1].
    """

    def __init__(self, nsteps, p, seed=0):
        rng = np.random.Generator(np.random.PCG64(seed))
        self.nsteps, self.p = int(nsteps), int(p)
        self.u_sched = rng.random(nsteps)
        self.u_mix = rng.random(nsteps)
        self.u_acc = rng.random(nsteps)
        # avoid log(0)
        self.u_acc[self.u_acc == 0.0] = np.finfo(np.float64).tiny
        self.z = rng.standard_normal((nsteps, p))


class NotPosDef(Exception):
    """Stands for Julia's PosDefException raised by `chol` / the MvNormal constructor."""


def _chol_lower(A):
    try:
        return np.linalg.cholesky(A)
    except np.linalg.LinAlgError as e:  # pragma: no cover - exercised in tests via crafted input
        raise NotPosDef(str(e))


# ------------------------------------------------------------------ the sampler
class GAMCOracle:
    """GAMC sampler + state (`src/samplers/GAMC.jl:3-151`) driven one `iterate!` at a time.

    target: object with `.logtarget(theta)` and `.uptotensorlogtarget(theta) -> (lt, grad, G)`
            (the Klara closures `parameter.logtarget!`, `parameter.uptotensorlogtarget!`).
    update: schedule name from SCHEDULES (+ `update_params`), as in `GAMC(update=...)`.
    transform: optional G -> G' (e.g. softabs), `src/samplers/GAMC.jl:128,164-166`.
    """

    def __init__(self, target, theta0, nsteps, burnin=0, update="rand_update!", update_params=(),
                 transform=None, driftstep=1.0, minorscale=1.0, c=0.05, t0=3, tuner=None):
        # GAMC inner constructor asserts, src/samplers/GAMC.jl:135-138
        assert driftstep > 0, "Drift step is not positive"
        assert 0 < minorscale, "Constant minorscale must be positive"
        assert 0 <= c, "Constant c must be non-negative"
        assert t0 > 0, "t0 is not positive"
        self.target, self.transform = target, transform
        self.update, self.update_params = update, tuple(update_params)
        self.driftstep, self.minorscale, self.c, self.t0 = float(driftstep), float(minorscale), float(c), int(t0)
        self.nsteps, self.burnin = int(nsteps), int(burnin)
        self.tuner = tuner if tuner is not None else GAMCTuner()
        self.p = int(np.asarray(theta0).shape[0])

        # initialize!  (src/samplers/GAMC.jl:157-176)
        self.theta = np.array(theta0, dtype=np.float64)
        self.logtarget, self.grad, self.G = self._upto_tensor(self.theta)
        assert math.isfinite(self.logtarget), "Log-target not finite: initial value out of support"
        assert np.all(np.isfinite(self.grad)), "Gradient of log-target not finite: initial values out of support"
        assert np.all(np.isfinite(self.G)), "Tensor of log-target not finite: initial values out of support"
        self.accept_diag = False

        # tuner_state  (src/samplers/GAMC.jl:195-200): 4th positional arg seeds totproposed with the period
        self.tune = GAMCTune(
            BasicMCTune(float("nan"), 0, 0, self.tuner.smmalatuner.period),
            BasicMCTune(float("nan"), 0, 0, self.tuner.amtuner.period),
            BasicMCTune(self.driftstep, 0, 0, self.tuner.totaltuner.period),
        )
        # sampler_state  (src/samplers/GAMC.jl:202-228)
        self.w = np.array([1 - self.c, self.c])
        self.sqrtminorscale = math.sqrt(self.minorscale)
        self.sqrttunestep = math.sqrt(self.driftstep)
        self.oldinvtensor = np.linalg.inv(self.G)
        self.cholinvtensor = _chol_lower(_hermitian_upper(self.oldinvtensor))
        self.oldfirstterm = self.oldinvtensor @ self.grad
        self.newinvtensor = np.full((self.p, self.p), np.nan)
        self.newfirstterm = np.full(self.p, np.nan)
        self.mu = np.full(self.p, np.nan)
        self.lastmean = self.theta.copy()
        self.secondlastmean = np.full(self.p, np.nan)
        self.presentupdatetensor, self.pastupdatetensor = True, False
        self.count, self.updatetensorcount = 0, 0
        self.ratio = float("nan")
        self.n_tensor_evals = 1  # bookkeeping for the bench (not a reference field)
        self.n_logtarget_evals = 0
        self.last_kind = None

    # -- closures
    def _upto_tensor(self, theta):
        lt, g, G = self.target.uptotensorlogtarget(theta)
        if self.transform is not None:
            G = self.transform(G)
        return float(lt), np.asarray(g, dtype=np.float64), np.asarray(G, dtype=np.float64)

    @property
    def _tuning(self):
        return self.tuner.verbose or isinstance(self.tuner.totaltuner, AcceptanceRateMCTuner)

    @property
    def _count_total(self):
        return self.tuner.totaltuner.verbose or isinstance(self.tuner.totaltuner, AcceptanceRateMCTuner)

    # -- one transition: src/samplers/iterate/GAMC.jl:1-293
    def iterate(self, tape: RandomTape):
        self.count += 1                                                             # :2
        t = self.count
        if self._tuning:                                                            # :4-6
            self.tune.totaltune.proposed += 1
        if t > self.t0:                                                             # :8-10
            self.presentupdatetensor = schedule_decision(
                self.update, t, self.nsteps, self.update_params, tape.u_sched[t - 1])
        eps = self.tune.totaltune.step
        z = tape.z[t - 1]
        logu = math.log(tape.u_acc[t - 1])

        if self.presentupdatetensor:                                                # :12 SMMALA branch
            self.last_kind = "smmala"
            self.updatetensorcount += 1                                             # :13
            if self.tuner.smmalatuner.verbose:
                self.tune.smmalatune.proposed += 1
            if not self.pastupdatetensor:                                           # :19-31
                self.logtarget, self.grad, self.G = self._upto_tensor(self.theta)
                self.n_tensor_evals += 1
                self.oldinvtensor = np.linalg.inv(self.G)                           # :26
                self.oldfirstterm = self.oldinvtensor @ self.grad                   # :28
                self.cholinvtensor = _chol_lower(_hermitian_upper(self.oldinvtensor))  # :30
            self.mu = self.theta + 0.5 * eps * self.oldfirstterm                    # :33
            prop = self.mu + self.sqrttunestep * (self.cholinvtensor @ z)           # :35
            try:
                lt_p, g_p, G_p = self._upto_tensor(prop)                            # :37-41
                self.n_tensor_evals += 1
                finite = math.isfinite(lt_p) and np.all(np.isfinite(g_p)) and np.all(np.isfinite(G_p))
            except (FloatingPointError, ZeroDivisionError):
                finite = False
            if finite:
                self.newinvtensor = np.linalg.inv(G_p)                              # :43
                ratio = lt_p - self.logtarget                                       # :45
                d = prop - self.mu
                ratio += 0.5 * (_logdet(eps * self.oldinvtensor) + float(d @ (self.G @ d)) / eps)  # :47-54
                self.newfirstterm = self.newinvtensor @ g_p                         # :56
                self.mu = prop + 0.5 * eps * self.newfirstterm                      # :58
                d = self.theta - self.mu
                ratio -= 0.5 * (_logdet(eps * self.newinvtensor) + float(d @ (G_p @ d)) / eps)    # :60-67
            else:
                ratio = -math.inf  # non-finite proposal => rejection (SURVEY.md 5 "Failure detection")
            self.ratio = ratio
            if ratio > 0 or ratio > logu:                                           # :69
                self.theta, self.grad, self.G, self.logtarget = prop.copy(), g_p, G_p, lt_p  # :70-98
                self.oldinvtensor = self.newinvtensor.copy()                        # :92
                self.cholinvtensor = _chol_lower(_hermitian_upper(self.newinvtensor))  # :94
                self.oldfirstterm = self.newfirstterm.copy()                        # :96
                self.accept_diag = True                                             # :108-112
                if self._count_total:
                    self.tune.totaltune.accepted += 1                               # :114-116
                if self.tuner.smmalatuner.verbose:
                    self.tune.smmalatune.accepted += 1
            else:
                self.accept_diag = False                                            # :121-126
        else:                                                                       # :128 AM branch
            self.last_kind = "am"
            if self.tuner.amtuner.verbose:
                self.tune.amtune.proposed += 1
            C = covariance_update(self.oldinvtensor, t - 2, self.theta, self.lastmean, self.secondlastmean)  # :133-140
            self.oldinvtensor = _hermitian_upper(C)                                 # :142
            Lcore = _chol_lower(eps * self.oldinvtensor)                            # set_gmm! :144 (MvNormal ctor)
            if tape.u_mix[t - 1] <= self.w[0]:                                      # rand(MixtureModel) :146
                prop = self.theta + Lcore @ z
            else:
                prop = self.theta + self.sqrtminorscale * z
            lt_p = float(self.target.logtarget(prop))                               # :148
            self.n_logtarget_evals += 1
            if math.isfinite(lt_p):
                ratio = lt_p - self.logtarget                                       # :150
                ratio -= gmm_logpdf(prop, self.theta, Lcore, self.sqrtminorscale, self.w)   # :152
                ratio += gmm_logpdf(self.theta, prop, Lcore, self.sqrtminorscale, self.w)   # :154-156
            else:
                ratio = -math.inf
            self.ratio = ratio
            if ratio > 0 or ratio > logu:                                           # :158
                self.theta, self.logtarget = prop.copy(), lt_p                      # :159-161
                self.accept_diag = True
                if self._count_total:
                    self.tune.totaltune.accepted += 1
                if self.tuner.amtuner.verbose:
                    self.tune.amtune.accepted += 1
            else:
                self.accept_diag = False

        self.secondlastmean = self.lastmean.copy()                                  # :193
        self.lastmean = recursive_mean(self.lastmean, t, self.theta)                # :195
        self.pastupdatetensor = self.presentupdatetensor                            # :197

        if self._tuning:                                                            # :199-292
            tt = self.tune.totaltune
            if tt.totproposed <= self.burnin and tt.proposed % self.tuner.totaltuner.period == 0:
                if self._count_total:
                    tt.rate = tt.accepted / tt.proposed                             # rate! :205
                if self.tuner.smmalatuner.verbose:
                    st = self.tune.smmalatune
                    self.tune.smmalafrequency = st.proposed / tt.proposed
                    st.rate = st.accepted / st.proposed if st.proposed else float("nan")
                if self.tuner.amtuner.verbose:
                    at = self.tune.amtune
                    self.tune.amfrequency = at.proposed / tt.proposed
                    at.rate = at.accepted / at.proposed if at.proposed else float("nan")
                if isinstance(self.tuner.totaltuner, AcceptanceRateMCTuner):        # tune! :270-273
                    tt.step *= self.tuner.totaltuner.score(tt.rate - self.tuner.totaltuner.targetrate)
                    self.sqrttunestep = math.sqrt(tt.step)
                for verbose, tn in ((self.tuner.smmalatuner.verbose, self.tune.smmalatune),
                                    (self.tuner.amtuner.verbose, self.tune.amtune)):
                    if verbose:                                                     # reset_burnin! :275-281
                        tn.totproposed += tn.proposed
                        tn.proposed, tn.accepted, tn.rate = 0, 0, float("nan")
                tt.totproposed += tt.proposed                                       # :283
                tt.proposed = 0                                                     # :285
                if self._count_total:                                               # :287-290
                    tt.accepted = 0
                    tt.rate = float("nan")
        return self.accept_diag

    # -- Klara `run(job)` + `output(job)`: save theta and :accept when i in burnin+1:nsteps
    def run(self, tape: RandomTape, nsteps=None):
        nsteps = self.nsteps if nsteps is None else nsteps
        nsave = max(0, nsteps - self.burnin)
        value = np.empty((self.p, nsave))
        accept = np.empty(nsave, dtype=bool)
        kinds = np.empty(nsteps, dtype=np.uint8)
        for i in range(1, nsteps + 1):
            acc = self.iterate(tape)
            kinds[i - 1] = 1 if self.last_kind == "smmala" else 0
            if i > self.burnin:
                value[:, i - self.burnin - 1] = self.theta
                accept[i - self.burnin - 1] = acc
        return value, accept, kinds


def _logdet(A):
    """Julia `logdet(A)` for a general dense matrix: LU based; sign must be +1 (else DomainError)."""
    sign, ld = np.linalg.slogdet(A)
    if sign <= 0:
        raise NotPosDef("logdet of a matrix with non-positive determinant")
    return float(ld)


class MALAOracle:
    """Klara's MALA sampler (the baseline of the reference's efficiency tables: `MALA(0.02)` with
    `AcceptanceRateMCTuner(0.574, score=x->logistic_rate_score(x, 3.))`, examples/logistic_regression/src/MALA/simulations.jl:37-41).
    Klara.jl is not vendored; restated from the published algorithm (Roberts & Tweedie 1996):
        mu = theta + eps/2 grad(theta);  theta* = mu + sqrt(eps) z;  mu* = theta* + eps/2 grad(theta*)
        log alpha = lt* - lt + |theta* - mu|^2/(2 eps) - |theta - mu*|^2/(2 eps)
    Tuning (every `period` proposals while totproposed <= burnin) as for the total tuner of GAMC.  `target` needs
    `.uptogradlogtarget(theta) -> (lt, grad)`."""

    def __init__(self, target, theta0, nsteps, burnin=0, driftstep=1.0, tuner=None):
        assert driftstep > 0, "Drift step is not positive"
        self.target, self.nsteps, self.burnin = target, int(nsteps), int(burnin)
        self.tuner = tuner if tuner is not None else VanillaMCTuner()
        self.theta = np.array(theta0, dtype=np.float64)
        self.p = self.theta.shape[0]
        self.logtarget, self.grad = target.uptogradlogtarget(self.theta)
        assert math.isfinite(self.logtarget) and np.all(np.isfinite(self.grad)), "initial values out of support"
        self.tune = BasicMCTune(float(driftstep), 0, 0, self.tuner.period)
        self.sqrttunestep = math.sqrt(driftstep)
        self.count = 0

    @property
    def _tuning(self):
        return self.tuner.verbose or isinstance(self.tuner, AcceptanceRateMCTuner)

    def iterate(self, tape):
        self.count += 1
        t = self.count
        if self._tuning:
            self.tune.proposed += 1
        eps = self.tune.step
        mu = self.theta + 0.5 * eps * self.grad
        prop = mu + self.sqrttunestep * tape.z[t - 1]
        lt_p, g_p = self.target.uptogradlogtarget(prop)
        if math.isfinite(lt_p) and np.all(np.isfinite(g_p)):
            mu_p = prop + 0.5 * eps * g_p
            d1, d2 = prop - mu, self.theta - mu_p
            ratio = lt_p - self.logtarget + float(d1 @ d1) / (2 * eps) - float(d2 @ d2) / (2 * eps)
        else:
            ratio = -math.inf
        acc = ratio > 0 or ratio > math.log(tape.u_acc[t - 1])
        if acc:
            self.theta, self.logtarget, self.grad = prop.copy(), lt_p, g_p
            if self._tuning:
                self.tune.accepted += 1
        if self._tuning:
            tt = self.tune
            if tt.totproposed <= self.burnin and tt.proposed % self.tuner.period == 0:
                tt.rate = tt.accepted / tt.proposed
                if isinstance(self.tuner, AcceptanceRateMCTuner):
                    tt.step *= self.tuner.score(tt.rate - self.tuner.targetrate)
                    self.sqrttunestep = math.sqrt(tt.step)
                tt.totproposed += tt.proposed
                tt.proposed, tt.accepted, tt.rate = 0, 0, float("nan")
        return acc

    def run(self, tape):
        nsave = max(0, self.nsteps - self.burnin)
        value = np.empty((self.p, nsave))
        accept = np.empty(nsave, dtype=bool)
        for i in range(1, self.nsteps + 1):
            acc = self.iterate(tape)
            if i > self.burnin:
                value[:, i - self.burnin - 1] = self.theta
                accept[i - self.burnin - 1] = acc
        return value, accept
