"""Step-by-step check of the integer-tensor-core metric on a GPU (development aid):
  GAMC_METRIC_I8=1 python tools/i8_probe.py [--N 4096 --p 256]
1. digit planes written by k_i8_slice against tests/i8_ref.py (bit for bit); 2. the likelihood metric of gamc_eval_partials against
the NumPy statement of the scheme and against the FP64 oracle; 3. timing of a level-2 evaluation with the FP64 route beside it."""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gamc_b200 as g  # noqa: E402
from gamc_b200 import _lib  # noqa: E402
import oracle as o  # noqa: E402
import i8_ref  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=4096)
    ap.add_argument("--p", type=int, default=256)
    ap.add_argument("--time-rows", type=int, default=0, help="also time a level-2 evaluation on a synthetic problem of this many rows")
    a = ap.parse_args()
    spec = os.environ.get("GAMC_METRIC_I8", "1")
    os.environ["GAMC_METRIC_I8"] = spec
    k, smax = (6, 7) if "," not in spec else tuple(int(v) for v in spec.split(","))
    lib = _lib.load()
    lib.gamc_debug_i8_read.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong]
    lib.gamc_debug_i8_read.restype = C.c_int
    X, y, tt = o.synth_logistic(20260101, a.N, a.p)
    theta = tt + 0.05 * np.random.default_rng(7).standard_normal(a.p)
    eng = g.Engine(0)
    eng.target_logistic(X, y, 100.0)
    ll, gl, Gl = eng.eval_partials(theta)
    ll0, gl0, Gl0 = o.logistic_partials(theta, X, y)
    info = (C.c_longlong * 4)()
    rc = lib.gamc_debug_i8_read(eng.h, 2, info, 32)
    print("i8 route active:", rc == 0, list(info))
    eta = X @ theta
    sig = 1.0 / (1.0 + np.exp(-eta))
    # the device weights: sigma (1 - sigma) with the device's sigmoid; compare planes through the reference built from the SAME weights
    ldn = int(info[2])
    buf = np.empty(k * a.p * ldn, dtype=np.int8)
    assert lib.gamc_debug_i8_read(eng.h, 0, buf.ctypes.data_as(C.c_void_p), buf.nbytes) == 0
    kc = int(info[3])
    planes_dev = buf.reshape(ldn // kc, k, a.p, kc).transpose(1, 2, 0, 3).reshape(k, a.p, ldn)[:, :, :a.N]
    w = sig * (1.0 - sig)
    planes_ref, e = i8_ref.digit_planes(X, w, k)
    ndiff = int(np.count_nonzero(planes_dev != planes_ref))
    print(f"digit planes: {ndiff} of {planes_ref.size} bytes differ (weights from the host sigmoid: a few last-digit differences are rounding of w)")
    cg = np.empty(a.p)
    assert lib.gamc_debug_i8_read(eng.h, 1, cg.ctypes.data_as(C.c_void_p), cg.nbytes) == 0
    print("column scales equal:", bool(np.array_equal(cg, np.ldexp(1.0, e))))
    G_from_dev_planes = i8_ref.metric_from_planes(planes_dev, e, smax)
    rel = lambda A, B: float(np.max(np.abs(A - B)) / np.max(np.abs(B)))
    print(f"G (device) vs NumPy sum over the DEVICE digit planes: {rel(Gl, G_from_dev_planes):.3e}   (tensor-core path exact => ~1e-16)")
    print(f"G (device) vs FP64 oracle:                            {rel(Gl, Gl0):.3e}")
    print(f"ll {abs(ll - ll0) / abs(ll0):.2e}  grad {rel(gl, gl0):.2e}  symmetric {bool(np.array_equal(Gl, Gl.T))}")
    if rel(Gl, G_from_dev_planes) > 1e-13:
        D = np.abs(Gl - G_from_dev_planes) / np.max(np.abs(Gl0))
        bad = np.argwhere(D > 1e-13)
        print("mismatching entries:", len(bad), "first:", bad[:8].tolist(), "rows hit:", sorted(set((bad[:, 0] // 8).tolist()))[:20])
    eng.close()
    if a.time_rows:
        for route in ("i8", "f64"):
            if route == "f64":
                os.environ.pop("GAMC_METRIC_I8", None)
            seed = 20260101
            t2 = g.synth_theta_true(seed, a.p)
            e2 = g.Engine(0)
            e2.target_logistic_synth(seed, 0, a.time_rows, a.p, t2, 100.0)
            e2.eval_upto_tensor(t2)
            e2.profile_enable(True)
            e2.profile_reset()
            t0 = time.perf_counter()
            for _ in range(10):
                e2.eval_upto_tensor(t2)
            dt = (time.perf_counter() - t0) / 10
            print(route, f"level-2 evaluation incl. host read-back {dt * 1e3:.3f} ms; families:", e2.profile_get())
            e2.close()


if __name__ == "__main__":
    main()
