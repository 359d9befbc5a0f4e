"""GAMC_CO_DEBUG=1 python tools/co_debug.py [--logN 20] [--p 256]: a few level-2 evaluations with the placement / timing report of the
co-scheduled pair on stderr (debug aid)."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("GAMC_CO_DEBUG", "1")
import gamc_b200 as g  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--logN", type=int, default=20)
ap.add_argument("--p", type=int, default=256)
a = ap.parse_args()
seed = 20260101
tt = g.synth_theta_true(seed, a.p)
eng = g.Engine(0)
eng.target_logistic_synth(seed, 0, 1 << a.logN, a.p, tt, 100.0)
for i in range(4):
    eng.eval_upto_tensor(tt + 0.01 * i)
eng.close()
