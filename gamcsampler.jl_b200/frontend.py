"""Host-side mirror of GAMCSampler.jl's public API for the hot path (Python stand-in for the Julia
front-end in julia/GAMCSamplerB200.jl; Julia is not available in this image).

Same names, keyword arguments, defaults and error behaviour as the reference:
  GAMC(update, transform, driftstep, minorscale, c, t0)                 src/samplers/GAMC.jl:126-151
  GAMCTuner(smmalatuner, amtuner, totaltuner)                           src/tuners/GAMCTuner.jl:15-23
  VanillaMCTuner / AcceptanceRateMCTuner / logistic_rate_score          Klara.jl (restated)
  nine schedules (Julia's trailing `!` dropped) and four decay functions src/samplers/GAMC.jl:264-337, src/stats/schedule.jl:1-7
  BasicMCRange / BasicMCJob / run / output / acceptance                 Klara.jl call sites in
                                                                        examples/logistic_regression/src/GAMC/simulations.jl:39-95
All arithmetic happens in the CUDA library; this module only builds the C config struct, draws the host
random tape and marshals arrays.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _lib as L
from .engine import Engine

__all__ = [
    "MALA", "GAMC", "GAMCTuner", "GAMCTune", "BasicMCTune", "VanillaMCTuner", "AcceptanceRateMCTuner", "logistic_rate_score",
    "am_only_update", "smmala_only_update", "mod_update", "rand_update", "cos_update", "rand_exp_decay_update",
    "rand_lin_decay_update", "rand_pow_decay_update", "rand_quad_decay_update",
    "exp_decay", "lin_decay", "pow_decay", "quad_decay",
    "BasicMCRange", "LogisticModel", "BasicMCJob", "Chain", "run", "output", "acceptance", "RandomTape",
    "synth_theta_true", "softabs", "Softabs", "MvTModel", "RVModel", "BatchedMCJob", "BatchedRandomTape", "GenericModel",
    "make_param_true_ex1", "make_param_true_ex2", "rv_model_velocity",
]


# ------------------------------------------------------------------ decay functions (src/stats/schedule.jl:1-7)
def exp_decay(i, tot, a=10.0, b=0.0):
    return (1 - b) * math.exp(-a * i / tot) + b


def lin_decay(i, tot, a=1e-3, b=0.0):
    return (1 - b) * (1 / (1 + a * i)) + b


def pow_decay(i, tot, a=1e-3, b=0.0):
    return (1 - b) * (a ** (i / tot)) + b


def quad_decay(i, tot, a=1e-3, b=0.0):
    return (1 - b) * (1 / (1 + a * i * i)) + b


# ------------------------------------------------------------------ schedules (src/samplers/GAMC.jl:264-337)
@dataclass(frozen=True)
class Schedule:
    """A schedule `update!` with its trailing parameters bound, e.g. the reference's
    `(sstate, pstate, i, tot) -> rand_exp_decay_update!(sstate, pstate, i, tot, 10.)`."""
    kind: int
    params: tuple = ()
    name: str = ""


def am_only_update():
    return Schedule(0, (), "am_only_update!")


def smmala_only_update():
    return Schedule(1, (), "smmala_only_update!")


def mod_update(n=10):
    return Schedule(2, (float(n),), "mod_update!")


def rand_update(p=0.5):
    return Schedule(3, (float(p),), "rand_update!")


def cos_update(a=10.0, b=0.2, c=0.2):
    return Schedule(4, (float(a), float(b), float(c)), "cos_update!")


def rand_exp_decay_update(a=10.0, b=0.0):
    return Schedule(5, (float(a), float(b)), "rand_exp_decay_update!")


def rand_lin_decay_update(a=1e-2, b=0.0):
    return Schedule(6, (float(a), float(b)), "rand_lin_decay_update!")


def rand_pow_decay_update(a=1e-3, b=0.0):
    return Schedule(7, (float(a), float(b)), "rand_pow_decay_update!")


def rand_quad_decay_update(a=1e-5, b=0.0):
    return Schedule(8, (float(a), float(b)), "rand_quad_decay_update!")


# ------------------------------------------------------------------ tuners
def logistic_rate_score(x, k=7.0):
    return 2.0 / (1.0 + math.exp(-k * x))


@dataclass
class VanillaMCTuner:
    period: int = 100
    verbose: bool = False


@dataclass
class AcceptanceRateMCTuner:
    targetrate: float = 0.35
    score_k: float = 7.0        # score = x -> logistic_rate_score(x, score_k)
    period: int = 100
    verbose: bool = False


@dataclass
class GAMCTuner:
    smmalatuner: VanillaMCTuner = field(default_factory=VanillaMCTuner)
    amtuner: VanillaMCTuner = field(default_factory=VanillaMCTuner)
    totaltuner: object = field(default_factory=VanillaMCTuner)

    @property
    def verbose(self):
        return bool(self.totaltuner.verbose or self.smmalatuner.verbose or self.amtuner.verbose)

    def __repr__(self):
        return "GAMCTuner"


@dataclass
class BasicMCTune:
    step: float = float("nan")
    accepted: int = 0
    proposed: int = 0
    totproposed: int = 0
    rate: float = float("nan")


@dataclass
class GAMCTune:
    smmalatune: BasicMCTune
    amtune: BasicMCTune
    totaltune: BasicMCTune
    smmalafrequency: float = float("nan")
    amfrequency: float = float("nan")


# ------------------------------------------------------------------ metric transform
@dataclass(frozen=True)
class Softabs:
    """`transform = H -> softabs(H, alpha)` (Klara softabs; examples/t_distribution/src/GAMC/simulations.jl:39)."""
    alpha: float = 1000.0


def softabs(alpha=1000.0):
    return Softabs(float(alpha))


# ------------------------------------------------------------------ sampler
class GAMC:
    """`GAMC(; update=rand_update!, transform=nothing, driftstep=1., minorscale=1., c=0.05, t0=3)`."""

    def __init__(self, update: Optional[Schedule] = None, transform=None, driftstep=1.0, minorscale=1.0, c=0.05, t0=3,
                 am_mode=0, reuse_tensor=False):
        assert driftstep > 0, "Drift step is not positive"
        assert 0 < minorscale, f"Constant minorscale must be positive, got {minorscale}"
        assert 0 <= c, "Constant c must be non-negative"
        assert t0 > 0, "t0 is not positive"
        if transform is not None and not isinstance(transform, Softabs) and not callable(transform):
            raise TypeError("transform must be softabs(alpha) or a callable H -> H' (the latter only with a GenericModel, whose "
                            "metric is evaluated on the host)")
        if update is not None and not isinstance(update, Schedule) and not callable(update):
            raise TypeError("update must be one of the nine schedules or a function (i, tot, u) -> Bool")
        # `update!::Function` (src/samplers/GAMC.jl:127): one of the nine schedules with its parameters bound, or any function
        # (i, tot, u) -> Bool, u being the iteration's schedule uniform (in place of the reference's in-line `rand(Bernoulli(.))`)
        self.update = update if update is not None else rand_update()
        self.transform = transform
        self.driftstep, self.minorscale, self.c, self.t0 = float(driftstep), float(minorscale), float(c), int(t0)
        self.am_mode = int(am_mode)
        self.reuse_tensor = bool(reuse_tensor)    # skip the re-evaluation at an unchanged current point (gamc_config.reuse_tensor)

    def __repr__(self):
        return (f"GAMC sampler: drift step = {self.driftstep}, scaling of stabilizing covariance = {self.minorscale}, "
                f"weight of stabilizing mixture component = {self.c}")


class MALA:
    """Klara's `MALA(driftstep)` (examples/logistic_regression/src/MALA/simulations.jl:37), the baseline the reference's
    efficiency tables are normalised by (examples/logistic_regression/src/table.jl:9-17).  Use with a plain tuner, e.g.
    `AcceptanceRateMCTuner(0.574, score_k=3.)` (:41)."""

    def __init__(self, driftstep=1.0):
        assert driftstep > 0, "Drift step is not positive"
        self.driftstep = float(driftstep)

    def __repr__(self):
        return f"MALA sampler: drift step = {self.driftstep}"


@dataclass
class BasicMCRange:
    nsteps: int
    burnin: int = 0


@dataclass
class LogisticModel:
    """`likelihood_model([Hyperparameter(:λ), Data(:X), Data(:y), p])` of the logistic example with its
    closed-form callbacks (examples/logistic_regression/src/GAMC/simulations.jl:25-48)."""
    X: Optional[np.ndarray] = None
    y: Optional[np.ndarray] = None
    lam: float = 100.0
    synth: Optional[dict] = None   # {"seed":..., "N":..., "p":..., "theta_true":..., "row0": 0, "nrows": N}


@dataclass
class GenericModel:
    """An arbitrary model given by its Klara-style closures: `uptotensorlogtarget(theta) -> (logtarget, grad,
    tensor)` (what `parameter.uptotensorlogtarget!` leaves in `pstate`, src/samplers/iterate/GAMC.jl:20,37) and, optionally,
    `logtarget(theta)` (`parameter.logtarget!`, :148; defaults to the first component of `uptotensorlogtarget`).  The host
    evaluates the model -- and the sampler's `transform` -- once per evaluation; the dense algebra, proposal, accept/reject,
    means, tuner and chain run on the device (`gamc_target_callback`)."""
    p: int
    uptotensorlogtarget: object
    logtarget: object = None


def _softabs_host(H, alpha):
    """Klara `softabs(H, alpha)` = Q diag(lambda coth(alpha lambda)) Q' (examples/t_distribution/src/GAMC/simulations.jl:39)."""
    lam, Q = np.linalg.eigh(0.5 * (H + H.T))
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        s = np.where(np.abs(alpha * lam) < 1e-8, 1.0 / alpha, lam / np.tanh(alpha * lam))
    return (Q * s) @ Q.T


def synth_theta_true(seed, p):
    """theta_true[j] ~ N(0, 1/p) of the synthetic logistic workloads: NumPy Generator(PCG64(seed))."""
    return np.random.Generator(np.random.PCG64(seed)).standard_normal(p) / np.sqrt(p)


class RandomTape:
    """Host-drawn normals and uniforms in fixed per-iteration slots."""

    def __init__(self, nsteps, p, seed=0):
        rng = np.random.Generator(np.random.PCG64(seed))
        self.u_sched = rng.random(nsteps)
        self.u_mix = rng.random(nsteps)
        self.u_acc = rng.random(nsteps)
        self.u_acc[self.u_acc == 0.0] = np.finfo(np.float64).tiny
        self.z = rng.standard_normal((nsteps, p))


def make_config(sampler, tuner, mcrange: BasicMCRange) -> L.Config:
    if isinstance(sampler, MALA):
        tt = tuner if not isinstance(tuner, GAMCTuner) else tuner.totaltuner
        cfg = L.Config()
        cfg.driftstep, cfg.minorscale, cfg.c, cfg.t0, cfg.schedule = sampler.driftstep, 1.0, 0.0, 1, 1
        for i in range(3):
            cfg.sched_params[i] = float("nan")
        cfg.total_tuner_kind = 1 if isinstance(tt, AcceptanceRateMCTuner) else 0
        cfg.targetrate = getattr(tt, "targetrate", float("nan"))
        cfg.score_k = getattr(tt, "score_k", 7.0)
        cfg.period = int(tt.period)
        cfg.verbose_total = int(bool(tt.verbose))
        cfg.nsteps, cfg.burnin = int(mcrange.nsteps), int(mcrange.burnin)
        cfg.sampler_kind = 1
        return cfg
    cfg = L.Config()
    cfg.driftstep, cfg.minorscale, cfg.c, cfg.t0 = sampler.driftstep, sampler.minorscale, sampler.c, sampler.t0
    is_fn = not isinstance(sampler.update, Schedule)          # a user function: installed with gamc_set_schedule_callback
    cfg.schedule = 1 if is_fn else sampler.update.kind
    pr = ([] if is_fn else list(sampler.update.params)) + [float("nan")] * 3
    for i in range(3):
        cfg.sched_params[i] = pr[i]
    tt = tuner.totaltuner
    cfg.total_tuner_kind = 1 if isinstance(tt, AcceptanceRateMCTuner) else 0
    cfg.targetrate = getattr(tt, "targetrate", float("nan"))
    cfg.score_k = getattr(tt, "score_k", 7.0)
    cfg.period = int(tt.period)
    cfg.verbose_total = int(bool(tt.verbose))
    cfg.verbose_smmala = int(bool(tuner.smmalatuner.verbose))
    cfg.verbose_am = int(bool(tuner.amtuner.verbose))
    cfg.nsteps, cfg.burnin = int(mcrange.nsteps), int(mcrange.burnin)
    cfg.am_mode = sampler.am_mode
    cfg.reuse_tensor = int(getattr(sampler, "reuse_tensor", False))
    return cfg


@dataclass
class Chain:
    value: np.ndarray              # p x nsamples (chain.value)
    diagnosticvalues: np.ndarray   # 1 x nsamples Bool (:accept)


class _SState:
    """View of job.sstate with the fields the examples read (simulations.jl:85-86)."""

    def __init__(self, eng: Engine):
        self._eng = eng

    def _refresh(self):
        return self._eng.state()

    @property
    def tune(self):
        s = self._refresh()
        return GAMCTune(
            BasicMCTune(float("nan"), s.smmala_accepted, s.smmala_proposed, s.smmala_totproposed),
            BasicMCTune(float("nan"), s.am_accepted, s.am_proposed, s.am_totproposed),
            BasicMCTune(s.step, s.total_accepted, s.total_proposed, s.total_totproposed, s.total_rate),
            s.smmalafrequency, s.amfrequency)

    @property
    def updatetensorcount(self):
        return self._refresh().updatetensorcount

    @property
    def count(self):
        return self._refresh().count

    @property
    def sqrttunestep(self):
        return self._refresh().sqrttunestep

    @property
    def ratio(self):
        return self._refresh().ratio

    @property
    def logtarget(self):
        return self._refresh().logtarget

    @property
    def loglikelihood(self):
        return self._eng.monitor()[0]

    @property
    def logprior(self):
        return self._eng.monitor()[1]

    def __getattr__(self, name):
        ids = {"lastmean": L.ARR_LASTMEAN, "secondlastmean": L.ARR_SECONDLASTMEAN, "oldinvtensor": L.ARR_OLDINVTENSOR,
               "cholinvtensor": L.ARR_CHOLINVTENSOR, "oldfirstterm": L.ARR_OLDFIRSTTERM,
               # monitored extras of the current point (iterate/GAMC.jl:74-90)
               "gradloglikelihood": L.ARR_GRADLOGLIKELIHOOD, "gradlogprior": L.ARR_GRADLOGPRIOR,
               "tensorloglikelihood": L.ARR_TENSORLOGLIKELIHOOD, "tensorlogprior": L.ARR_TENSORLOGPRIOR,
               "gradlogtarget": L.ARR_GRADLOGTARGET, "tensorlogtarget": L.ARR_TENSORLOGTARGET, "value": L.ARR_VALUE}
        if name in ids:
            return self._eng.array(ids[name])
        raise AttributeError(name)


class BasicMCJob:
    """`BasicMCJob(model, sampler, mcrange, v0, tuner=..., outopts=...)` (simulations.jl:71) for the GAMC sampler
    on one B200 (device=...) or on this rank's row shard of a multi-GPU job (comm=(uid, nranks, rank))."""

    def __init__(self, model: LogisticModel, sampler: GAMC, mcrange: BasicMCRange, v0, tuner: Optional[GAMCTuner] = None,
                 outopts=None, device=0, tape: Optional[RandomTape] = None, tape_seed=0, comm=None, engine: Optional[Engine] = None):
        self.model, self.sampler, self.range = model, sampler, mcrange
        self.tuner = tuner if tuner is not None else GAMCTuner()
        self.outopts = outopts or {"monitor": ["value"], "diagnostics": ["accept"]}
        self.engine = engine if engine is not None else Engine(device)
        eng = self.engine
        if isinstance(model, GenericModel):
            tr = getattr(sampler, "transform", None)

            def _eval(theta, level, _m=model, _tr=tr):
                if level == 0 and _m.logtarget is not None:
                    return float(_m.logtarget(theta)), None, None
                lt, grad, G = _m.uptotensorlogtarget(theta)
                if level >= 2 and _tr is not None:
                    G = _softabs_host(np.asarray(G, dtype=np.float64), _tr.alpha) if isinstance(_tr, Softabs) else _tr(G)
                return lt, grad, G

            eng.target_callback(model.p, _eval)
        elif getattr(sampler, "transform", None) is not None:
            # the reference applies sampler.transform to the metric at iterate/GAMC.jl:22-24,39-41; the single-chain logistic path
            # has no device transform (the logistic metric is positive definite and the example passes none), so refuse rather
            # than run with different semantics
            raise NotImplementedError("transform= is supported for GenericModel (host or device callback) and for the batched "
                                      "chains (BatchedMCJob); the built-in single-chain logistic target takes transform=None")
        elif engine is None:
            if model.synth is not None:
                s = model.synth
                eng.target_logistic_synth(s["seed"], s.get("row0", 0), s.get("nrows", s["N"]), s["p"], s["theta_true"], model.lam)
            else:
                eng.target_logistic(model.X, model.y, model.lam)
        if comm is not None:
            eng.comm_init(*comm)
        theta0 = np.asarray(v0["p"] if isinstance(v0, dict) else v0, dtype=np.float64)
        self.cfg = make_config(sampler, self.tuner, mcrange)
        if not isinstance(sampler, MALA):
            eng.set_schedule_callback(None if isinstance(sampler.update, Schedule) else sampler.update)
        eng.sampler_init(self.cfg, theta0)          # initialize! + sampler_state
        self.tape = tape
        self.tape_seed = tape_seed
        self.sstate = _SState(eng)
        self._done = 0

    def run(self, chunk=None):
        """Klara `run(job)`: nsteps transitions; samples after burn-in stay on the device until `output`."""
        n = self.range.nsteps - self._done
        eng = self.engine
        if self.tape is not None:
            t = self.tape
            lo, hi = self._done, self._done + n
            eng.tape_set(t.u_sched[lo:hi], t.u_mix[lo:hi], t.z[lo:hi], t.u_acc[lo:hi])
        else:
            eng.tape_generate(n, self.tape_seed)
        eng.run(n)
        eng.sync()
        self._done += n
        if isinstance(self.tuner, GAMCTuner) and self.tuner.verbose:
            self._print_burnin_report()
        return self

    def _print_burnin_report(self):
        """The burn-in report of src/samplers/iterate/GAMC.jl:218-268, printed from the records the device kept."""
        t = self.tuner
        for r in self.engine.tune_log()[getattr(self, "_reported", 0):]:
            print(f"Burnin iteration {int(r[0])} out of {self.range.burnin}...")
            if t.totaltuner.verbose:
                print(f"  Total : {int(r[1])}/{int(r[2])} ({100 * r[3]:.2f}%) acceptance rate")
            if t.smmalatuner.verbose:
                print(f"  SMMALA: {int(r[4])}/{int(r[5])} ({100 * r[6]:.2f}%) acceptance rate, {int(r[5])}/{int(r[2])} ({100 * r[7]:.2f}%) running frequency")
            if t.amtuner.verbose:
                print(f"  AM    : {int(r[8])}/{int(r[9])} ({100 * r[10]:.2f}%) acceptance rate, {int(r[9])}/{int(r[2])} ({100 * r[11]:.2f}%) running frequency")
        self._reported = len(self.engine.tune_log())

    def output(self) -> Chain:
        nsave = max(0, self.range.nsteps - self.range.burnin)
        value, accept = self.engine.chain_get(0, nsave)
        return Chain(value, accept.reshape(1, -1))


def run(job: BasicMCJob):
    return job.run()


def output(job: BasicMCJob) -> Chain:
    return job.output()


def acceptance(chain_or_diag) -> float:
    """Klara `acceptance(chain)` / `acceptance(::Matrix{Bool})`: mean of the :accept diagnostic."""
    d = chain_or_diag.diagnosticvalues if isinstance(chain_or_diag, Chain) else chain_or_diag
    return float(np.mean(np.asarray(d, dtype=bool)))


# ------------------------------------------------------------------ batched chains (small p)
@dataclass
class MvTModel:
    """`plogtarget(p, v) = logpdf(MvTDist(nu, zeros(n), (nu-2)*Sigma/nu), p)` with Sigma_ij = c^|i-j|
    (examples/t_distribution/src/GAMC/simulations.jl:17-33).  Model constants (inverse scale matrix, normalising
    constant) are set up once on the host; every evaluation happens in the kernel."""
    n: int = 20
    nu: float = 30.0
    c: float = 0.9

    def tridiagonal(self):
        """Diagonal and off-diagonal of the inverse scale matrix T = inv(Sigma_t): for Sigma_ij = c^|i-j| it is tridiagonal,
        T = nu / ((nu - 2)(1 - c^2)) tridiag(-c; 1, 1 + c^2, ..., 1 + c^2, 1; -c) -- what the large-dimension path works from."""
        n, s = self.n, self.nu / ((self.nu - 2.0) * (1.0 - self.c * self.c))
        td = np.full(n, s * (1.0 + self.c * self.c))
        td[0] = td[-1] = s
        if n == 1:
            td[0] = self.nu / (self.nu - 2.0)
        return td, np.full(max(n - 1, 0), -s * self.c)

    def constants(self):
        idx = np.arange(self.n)
        Sigma_t = (self.nu - 2.0) * (self.c ** np.abs(idx[:, None] - idx[None, :])) / self.nu
        Sinv = np.linalg.inv(Sigma_t)
        Sinv = 0.5 * (Sinv + Sinv.T)
        _, logdet = np.linalg.slogdet(Sigma_t)
        logconst = (math.lgamma((self.nu + self.n) / 2.0) - math.lgamma(self.nu / 2.0) - 0.5 * self.n * math.log(self.nu * math.pi) - 0.5 * logdet)
        return Sinv, logconst


@dataclass
class RVModel:
    """Keplerian radial-velocity model of examples/radial_velocity/src (stat_model.jl:18-63): data = epochs t, observed
    velocities obs, uncertainties sigma_obs (set_times / set_obs / set_sigma_obs); nplanets in {1, 2}."""
    times: np.ndarray
    obs: np.ndarray
    sigma_obs: np.ndarray
    nplanets: int = 1

    @classmethod
    def from_csv(cls, path, nplanets=1):
        """The reference's data files (`examples/radial_velocity/data/one_planet.csv`: rows `time,obs,sigma_obs`, written by
        `writedlm(..., zip(times, obs, sigma_obs), ',')`, one_planet/data_generation.jl:30-34)."""
        d = np.loadtxt(path, delimiter=",", ndmin=2)
        return cls(d[:, 0].copy(), d[:, 1].copy(), d[:, 2].copy(), nplanets)

    def to_csv(self, path):
        np.savetxt(path, np.column_stack([self.times, self.obs, self.sigma_obs]), delimiter=",", fmt="%.17g")

    @classmethod
    def simulate(cls, param_true, num_obs=50, observation_timespan=2 * 365.25, sigma_obs_scalar=2.0, seed=42):
        """`data_generation.jl` of the one- and two-planet examples (:13-28): epochs = timespan * sort(rand(num_obs)), observations =
        true model + sigma * randn (the jitter term is switched off in the reference, `num_jitters = 0`, rv_model_param.jl:5).
        Julia's `srand(42)` stream is not reproducible here; NumPy PCG64(seed) stands in."""
        rng = np.random.Generator(np.random.PCG64(seed))
        param_true = np.asarray(param_true, dtype=np.float64)
        times = observation_timespan * np.sort(rng.random(num_obs))
        sigma = sigma_obs_scalar * np.ones(num_obs)
        obs = rv_model_velocity(param_true, times) + sigma_obs_scalar * rng.standard_normal(num_obs)
        return cls(times, obs, sigma, len(param_true) // 5)


def make_param_true_ex1():
    """One-planet truth of utils_ex.jl:48-57 packed as rv_model_param.jl:19-44: [log1p(P), log1p(K), e cos w, e sin w, w + M0, C]."""
    P, K, e, w, M0, C = 50.0, 20.0, 0.2, math.pi / 4, math.pi / 4, 1.0
    return np.array([math.log1p(P), math.log1p(K), e * math.cos(w), e * math.sin(w), w + M0, C])


def make_param_true_ex2():
    """Two-planet truth of utils_ex.jl:59-68."""
    out = []
    for P, K in ((40.0, 30.0), (80.8, 30.0)):
        e, w, M0 = 0.2, math.pi / 4, math.pi / 4
        out += [math.log1p(P), math.log1p(K), e * math.cos(w), e * math.sin(w), w + M0]
    return np.array(out + [0.0])


def rv_model_velocity(theta, times):
    """`calc_model_rv` (rv_model_keplerian.jl:9-51) on the host, for data generation only: sum over planets of the Keplerian
    velocity K j / (1 - e cos E) [cos(w + M + e sin E) - k e cos E / (1 + j)], j = sqrt(1 - e^2), plus the offset; Kepler's
    equation by Newton iterations from the Danby guess (kepler_eqn.jl:18-23) to machine precision."""
    theta = np.asarray(theta, dtype=np.float64)
    t = np.asarray(times, dtype=np.float64)
    npl = len(theta) // 5
    z = np.zeros_like(t)
    for pl in range(npl):
        P, K = math.expm1(theta[5 * pl]), math.expm1(theta[5 * pl + 1])
        h, k = theta[5 * pl + 2], theta[5 * pl + 3]
        ecc, w = math.hypot(h, k), math.atan2(k, h)
        if not ecc < 1.0:
            raise ValueError("eccentricity must be below one (rv_model_param.jl:83-96)")
        M = t * (2.0 * math.pi / P) - (theta[5 * pl + 4] - w)
        Mr = np.mod(M, 2.0 * math.pi)
        E = np.where(Mr < math.pi, Mr + 0.85 * ecc, Mr - 0.85 * ecc)
        for _ in range(60):
            dE = (E - ecc * np.sin(E) - Mr) / (1.0 - ecc * np.cos(E))
            E = E - dE
            if np.max(np.abs(dE)) < 1e-15:
                break
        j = math.sqrt((1.0 - ecc) * (1.0 + ecc))
        q = ecc * np.cos(E)
        z = z + K * j / (1.0 - q) * (np.cos(w + M + ecc * np.sin(E)) - k * q / (1.0 + j))
    return z + theta[5 * npl]


class BatchedRandomTape:
    """Per-chain host tapes: chain c uses RandomTape(nsteps, p, seed + c)."""

    def __init__(self, nchains, nsteps, p, seed=0):
        tapes = [RandomTape(nsteps, p, seed + c) for c in range(nchains)]
        self.u_sched = np.stack([t.u_sched for t in tapes])
        self.u_mix = np.stack([t.u_mix for t in tapes])
        self.u_acc = np.stack([t.u_acc for t in tapes])
        self.z = np.stack([t.z for t in tapes])


class BatchedMCJob:
    """The reference's `while i <= nchains ... BasicMCJob(...); run(job)` loop (examples/t_distribution/src/GAMC/simulations.jl:56-79)
    as ONE persistent kernel: every chain is an independent GAMC job with its own initial value v0s[c].
    A t target of dimension n > 32 runs on the structured large-dimension path (gamc_bigd_*; n <= 1024)."""

    def __init__(self, model, sampler: GAMC, mcrange: BasicMCRange, v0s, tuner: Optional[GAMCTuner] = None, device=0,
                 tape: Optional[BatchedRandomTape] = None, seed=0, engine: Optional[Engine] = None, chain_offset=0, max_factors=0):
        self.model, self.sampler, self.range = model, sampler, mcrange
        self.tuner = tuner if tuner is not None else GAMCTuner()
        self.engine = engine if engine is not None else Engine(device)
        self.bigd = isinstance(model, MvTModel) and model.n > 32
        if self.bigd:
            if not isinstance(sampler.update, Schedule):
                raise NotImplementedError("the batched kernels decide the branch on the device: pass one of the nine schedules")
            tr = sampler.transform
            if tr is not None and not isinstance(tr, Softabs):
                raise NotImplementedError("the large-dimension t target supports transform=None or softabs(alpha)")
            td, te = model.tridiagonal()
            self.engine.bigd_target_mvt(td, te, model.nu)
            self.cfg = make_config(sampler, self.tuner, mcrange)
            self.engine.bigd_set_chain_offset(chain_offset)
            self.engine.bigd_init(self.cfg, np.asarray(v0s, dtype=np.float64), 1 if isinstance(tr, Softabs) else 0,
                                  tr.alpha if isinstance(tr, Softabs) else 0.0, max_factors)
            self.tape, self.seed = tape, seed
            self._done = 0
            return
        if isinstance(model, RVModel):
            self.engine.batched_target_rv(model.times, model.obs, model.sigma_obs, model.nplanets)
        elif isinstance(model, LogisticModel):
            self.engine.batched_target_logistic(model.X, model.y, model.lam)    # the as-shipped sizes (N = 200, p = 4)
        else:
            Sinv, logconst = model.constants()
            self.engine.batched_target_mvt(Sinv, model.nu, logconst)
        if not isinstance(sampler.update, Schedule):
            raise NotImplementedError("the batched kernel decides the branch on the device: pass one of the nine schedules")
        self.cfg = make_config(sampler, self.tuner, mcrange)
        self.engine.batched_set_chain_offset(chain_offset)     # this rank's first chain in a chains-sharded multi-GPU job
        tr = sampler.transform
        self.engine.batched_init(self.cfg, np.asarray(v0s, dtype=np.float64), 1 if isinstance(tr, Softabs) else 0,
                                 tr.alpha if isinstance(tr, Softabs) else 0.0)
        self.tape, self.seed = tape, seed
        self._done = 0

    def run(self):
        n = self.range.nsteps - self._done
        runner = self.engine.bigd_run if self.bigd else self.engine.batched_run
        if self.tape is not None:
            lo, hi = self._done, self._done + n
            t = self.tape
            runner(n, self.seed, {"u_sched": t.u_sched[:, lo:hi], "u_mix": t.u_mix[:, lo:hi], "u_acc": t.u_acc[:, lo:hi], "z": t.z[:, lo:hi]})
        else:
            runner(n, self.seed)
        self.engine.sync()
        self._done += n
        return self

    def output(self):
        """List of Chain objects (value p x nsamples, as `chain.value` of the reference)."""
        value, accept = self.engine.bigd_chain_get() if self.bigd else self.engine.batched_chain_get()
        return [Chain(np.ascontiguousarray(value[c].T), accept[c].reshape(1, -1)) for c in range(value.shape[0])]
