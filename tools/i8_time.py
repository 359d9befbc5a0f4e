"""A few level-2 evaluations on the C2-shaped synthetic problem (for `ncu` launch lists / captures of the integer-tensor-core metric):
  GAMC_METRIC_I8=1 python tools/i8_time.py [--rows 1048576 --p 256 --evals 4]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gamc_b200 as g  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1 << 20)
    ap.add_argument("--p", type=int, default=256)
    ap.add_argument("--evals", type=int, default=4)
    a = ap.parse_args()
    seed = 20260101
    tt = g.synth_theta_true(seed, a.p)
    eng = g.Engine(0)
    eng.target_logistic_synth(seed, 0, a.rows, a.p, tt, 100.0)
    eng.eval_upto_tensor(tt)
    eng.profile_enable(True)
    eng.profile_reset()
    t0 = time.perf_counter()
    for _ in range(a.evals):
        lt, gr, G = eng.eval_upto_tensor(tt)
    dt = (time.perf_counter() - t0) / a.evals
    print(f"route {eng.metric_route()}: {dt * 1e3:.3f} ms per evaluation (host loop), logtarget {lt:.6f}")
    for k, (n, ms) in eng.profile_get().items():
        if n:
            print(f"  {k:12s} {n:4d} launches  {ms / n:8.4f} ms each")
    eng.close()


if __name__ == "__main__":
    main()
