"""Phase clocks of the single-CTA step kernels, measured in situ (debug build, not a benchmark):
  make -C gamcsampler.jl_b200/csrc dbg && GAMC_LIB=$PWD/gamcsampler.jl_b200/libgamc_b200_dbg.so python tools/phase_probe.py [--p 256]
Runs a short GAMC chain on the C2-shaped synthetic problem and prints the mean microseconds between the PHASE() marks
of k_smmala_pre / k_smmala_post / factor_slot / chol_blocked (clock64 at the SM clock nvidia-smi reports)."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gamc_b200 as g  # noqa: E402
from gamc_b200 import _lib  # noqa: E402

NAMES = {
    1: "post: landed finite", 2: "post: (after factor_slot)", 4: "post: ratio sums",
    5: "post: accept+tail", 16: "factor: entry", 17: "factor: chol epilogue", 18: "factor: trsv_lower", 19: "factor: trsv_lower_t",
    32: "pre(refactor): prior+factor tail", 33: "pre(no refactor): entry", 34: "pre: trsv_t + proposal",
    48: "chol: pass",
    52: "  warp0: team update of the diagonal block", 53: "  warp0: column chain", 54: "  warp0: publish", 55: "  warp4: inverter (whole pass)",
    56: "  warp1: trailing tasks + barrier", 57: "  warp1: panel solve (pipelined)",
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--p", type=int, default=256)
    ap.add_argument("--rows", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--mhz", type=float, default=1965.0)
    a = ap.parse_args()
    if "dbg" not in _lib.LIB_PATH:
        raise SystemExit("set GAMC_LIB to the debug build (make dbg)")
    lib = _lib.load()
    lib.gamc_debug_phases.argtypes = [C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.c_int]
    seed = 20260101
    tt = g.synth_theta_true(seed, a.p)
    eng = g.Engine(0)
    eng.target_logistic_synth(seed, 0, a.rows, a.p, tt, 100.0)
    sampler = g.GAMC(update=g.rand_exp_decay_update(10.0), driftstep=0.02, minorscale=0.001, c=0.01)
    tuner = g.GAMCTuner(g.VanillaMCTuner(), g.VanillaMCTuner(), g.AcceptanceRateMCTuner(0.35))
    job = g.BasicMCJob(g.LogisticModel(lam=100.0), sampler, g.BasicMCRange(a.steps, a.steps // 5), {"p": tt.copy()}, tuner=tuner,
                       tape=g.RandomTape(a.steps, a.p, seed=3), engine=eng)
    acc = (C.c_longlong * 64)()
    cnt = (C.c_longlong * 64)()
    lib.gamc_debug_phases(acc, cnt, 1)
    job.run()
    lib.gamc_debug_phases(acc, cnt, 0)
    print(f"p={a.p} rows={a.rows} steps={a.steps}  (us at {a.mhz} MHz)")
    for k in sorted(NAMES):
        if cnt[k]:
            print(f"  [{k:2d}] {NAMES[k]:36s} hits {cnt[k]:6d}  mean {acc[k] / cnt[k] / a.mhz:9.2f} us  total {acc[k] / a.mhz / 1e3:9.2f} ms")


if __name__ == "__main__":
    main()
