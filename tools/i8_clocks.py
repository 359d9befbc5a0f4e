"""Where the MMA issuer and the TMA producer of k_metric_i8 spend their time (debug build, not a benchmark):
  make -C gamcsampler.jl_b200/csrc dbg && GAMC_METRIC_I8=1 GAMC_LIB=$PWD/gamcsampler.jl_b200/libgamc_b200_dbg.so python tools/i8_clocks.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gamc_b200 as g  # noqa: E402
from gamc_b200 import _lib  # noqa: E402

NAMES = {58: "MMA warp: waiting for operand tiles", 59: "MMA warp: waiting for the epilogue (TMEM)", 60: "MMA warp: issue", 61: "MMA warp: total",
         62: "producer: waiting for free slots", 63: "producer: total"}


def main():
    lib = _lib.load()
    lib.gamc_debug_phases.argtypes = [C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.c_int]
    rows, p = int(os.environ.get("ROWS", 1 << 20)), int(os.environ.get("P", 256))
    tt = g.synth_theta_true(20260101, p)
    eng = g.Engine(0)
    eng.target_logistic_synth(20260101, 0, rows, p, tt, 100.0)
    eng.eval_upto_tensor(tt)
    acc = (C.c_longlong * 64)()
    cnt = (C.c_longlong * 64)()
    lib.gamc_debug_phases(acc, cnt, 1)
    n = 3
    for _ in range(n):
        eng.eval_upto_tensor(tt)
    lib.gamc_debug_phases(acc, cnt, 0)
    for k in sorted(NAMES):
        if cnt[k]:
            print(f"  {NAMES[k]:45s} {acc[k] / cnt[k] / 1e3:10.1f} k cycles per CTA and evaluation")
    eng.close()


if __name__ == "__main__":
    main()
