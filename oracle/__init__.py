"""CPU FP64 oracle for the GAMC hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This package restates, in NumPy, the arithmetic of papamarkou/GAMCSampler.jl's per-iteration hot path
(log-target / gradient / metric evaluation and the SMMALA / AM transition that consumes them).  Every
function cites the reference file:line it follows.

PARITY UNPINNED: the reference ships no golden vectors, known-answer tests or fixtures for this path
(`test/runtests.jl:1-5` is `@test 1 == 1`) and neither Julia nor Klara.jl 0.9.3 / Distributions.jl 0.12.1
(`REQUIRE:1-3`) can run in this image, so the oracle cannot be checked against outputs of the reference
itself.  It is anchored instead on closed-form identities, SciPy as an independent implementation, and the
radial-velocity CSV fixtures shipped with the reference (see tests/test_oracle_*.py, tests/golden/).

Only `tests/`, `__graft_entry__.smoke()` and the CPU-baseline / `--impl reference` legs of `bench.py`
may import this package.  The product package (`gamcsampler.jl_b200/`) never does.
"""
from .logistic import (  # noqa: F401
    ploglikelihood, ploglikelihood_naive, plogprior, pgradlogtarget, ptensorlogtarget,
    logistic_partials, LogisticTarget, synth_logistic,
)
from .sampler import (  # noqa: F401
    GAMCOracle, MALAOracle, RandomTape, BasicMCTune, VanillaMCTuner, AcceptanceRateMCTuner, GAMCTuner,
    logistic_rate_score, covariance_update, recursive_mean, gmm_logpdf,
    exp_decay, lin_decay, pow_decay, quad_decay, SCHEDULES, schedule_probability,
)
from .stats import ess, acceptance, mcvar_imse  # noqa: F401
