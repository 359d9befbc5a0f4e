"""A/B of the level-2 evaluation (log-likelihood + gradient + metric) on the C2-shaped synthetic problem:
two kernels one after the other (GAMC_LEVEL2_CO=0) against the co-scheduled pair (default), for a few throttle lags.
    python tools/level2_probe.py [--logN 20] [--p 256] [--reps 12]
Prints one JSON line per configuration: GPU milliseconds per evaluation of the data-parallel kernels (CUDA events on the
launching stream: `loglik_grad` + `metric` families; the co-scheduled pair is one scope booked under `metric`) and the
agreement of the results with the serial route."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gamc_b200 as g  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logN", type=int, default=20)
    ap.add_argument("--p", type=int, default=256)
    ap.add_argument("--reps", type=int, default=12)
    a = ap.parse_args()
    seed = 20260101
    N, p = 1 << a.logN, a.p
    tt = g.synth_theta_true(seed, p)
    rng = np.random.default_rng(1)
    thetas = [tt + 0.02 * rng.standard_normal(p) for _ in range(a.reps)]
    eng = g.Engine(0)
    base = None
    for label, env in [("serial", {"GAMC_LEVEL2_CO": "0"}), ("co lag 48 MB", {}), ("co lag 16 MB", {"GAMC_CO_LAG_MB": "16"}),
                       ("co lag 96 MB", {"GAMC_CO_LAG_MB": "96"}), ("co no throttle", {"GAMC_CO_LAG_MB": "0"}), ("serial again", {"GAMC_LEVEL2_CO": "0"})]:
        for k in ("GAMC_LEVEL2_CO", "GAMC_CO_LAG_MB"):
            os.environ.pop(k, None)
        os.environ.update(env)
        eng.target_logistic_synth(seed, 0, N, p, tt, 100.0)
        for t in thetas[:3]:
            eng.eval_upto_tensor(t)          # warm-up
        eng.profile_enable(True)
        eng.profile_reset()
        res = [eng.eval_upto_tensor(t) for t in thetas]
        eng.sync()
        prof = eng.profile_get()
        eng.profile_enable(False)
        n_lg, ms_lg = prof["loglik_grad"]
        n_mt, ms_mt = prof["metric"]
        out = {"config": label, "N": N, "p": p, "evals": len(thetas), "ms_loglik_grad": ms_lg / max(n_lg, 1), "ms_metric_scope": ms_mt / max(n_mt, 1),
               "ms_level2": (ms_lg + ms_mt) / len(thetas)}
        if base is None:
            base = res
        else:
            out["ll_relerr_max"] = max(abs(r[0] - b[0]) / abs(b[0]) for r, b in zip(res, base))
            out["metric_relerr_max"] = max(float(np.linalg.norm(r[2] - b[2]) / np.linalg.norm(b[2])) for r, b in zip(res, base))
            out["grad_relerr_max"] = max(float(np.linalg.norm(r[1] - b[1]) / np.linalg.norm(b[1])) for r, b in zip(res, base))
        print(json.dumps(out), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
