"""Generates the committed golden fixtures under tests/golden/ (run in the build container, where /root/reference
is mounted; the GPU box never reads /root/reference).

  rv_fixtures.json   the two 50-epoch radial-velocity data sets shipped with the reference
                     (examples/radial_velocity/data/{one,two}_planet(s).csv: t, rv, sigma) together with the truth vectors
                     of utils_ex.jl:48-68 and the log-likelihood / log-prior / log-target the oracle restatement gives there.
  logistic_kat.json  oracle outputs (log-target, gradient, metric) on a small seeded logistic problem, so that later
                     edits of the oracle cannot drift silently.
  i8_scheme.json     the integer route of the metric on the same problem (tests/i8_ref.py): column exponents, the six int8 digit
                     planes of diag(sqrt w) X and the metric the 21 digit products give -- pins the DEFINITION of the scheme
                     (fixed-point scale, balanced digits, truncation at a + b <= 7) that the device code implements.

The reference ships no expected outputs of its own (test/runtests.jl:1-5 is `@test 1 == 1`): these numbers pin our
reading of the model, nothing more (parity unpinned).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from oracle import targets  # noqa: E402

REF = "/root/reference/examples/radial_velocity/data"


def rv():
    out = {}
    for name, truth in (("one_planet", targets.make_param_true_ex1()), ("two_planets", targets.make_param_true_ex2())):
        d = np.loadtxt(os.path.join(REF, name + ".csv"), delimiter=",")
        m = targets.RVModel(d[:, 0], d[:, 1], d[:, 2])
        out[name] = {
            "t": d[:, 0].tolist(), "rv": d[:, 1].tolist(), "sigma": d[:, 2].tolist(), "theta_true": truth.tolist(),
            "loglik": m.loglikelihood(truth), "logprior": m.logprior(truth), "logtarget": m.logtarget(truth),
            "chisq": float(np.sum((m.model_rv(truth) - d[:, 1]) ** 2 / d[:, 2] ** 2)),
        }
    with open(os.path.join(HERE, "rv_fixtures.json"), "w") as f:
        json.dump(out, f)
    return out


def logistic():
    X, y, tt = oracle.synth_logistic(20260101, 64, 5)
    theta = tt + 0.1 * np.arange(5)
    lt, g, G = oracle.LogisticTarget(X, y, 100.0).uptotensorlogtarget(theta)
    out = {"seed": 20260101, "N": 64, "p": 5, "theta": theta.tolist(), "X": X.tolist(), "y": y.tolist(),
           "logtarget": lt, "grad": g.tolist(), "G": G.tolist()}
    with open(os.path.join(HERE, "logistic_kat.json"), "w") as f:
        json.dump(out, f)


def i8_scheme():
    sys.path.insert(0, os.path.dirname(HERE))
    import i8_ref
    X, y, tt = oracle.synth_logistic(20260101, 64, 5)
    theta = tt + 0.1 * np.arange(5)
    sig = 1.0 / (1.0 + np.exp(-(X @ theta)))
    w = sig * (1.0 - sig)
    planes, e = i8_ref.digit_planes(X, w, 6)
    G = i8_ref.metric_from_planes(planes, e, 7)
    out = {"seed": 20260101, "N": 64, "p": 5, "slices": 6, "smax": 7, "theta": theta.tolist(), "w": w.tolist(), "column_exponents": e.tolist(),
           "planes": planes.astype(int).tolist(), "G_lik": G.tolist()}
    with open(os.path.join(HERE, "i8_scheme.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "i8":      # (does not need /root/reference)
        i8_scheme()
        sys.exit(0)
    r = rv()
    for k, v in r.items():
        print(k, v["loglik"], v["logprior"], v["logtarget"], v["chisq"])
    logistic()
    i8_scheme()
