"""Out-of-bounds-write check without compute-sanitizer (closed on this GPU pool): runs the small jobs of tools/sanitizer_job.py
against the guard build of the library (`make -C gamcsampler.jl_b200/csrc guard`: every device allocation sits between two 4 KB
guard zones filled with a pattern) and counts overwritten guard bytes before every handle is destroyed.
  python tools/guard_check.py        # prints GUARD_CHECK ... PASS / FAIL"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["GAMC_LIB"] = os.path.join(ROOT, "gamcsampler.jl_b200", "libgamc_b200_guard.so")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gamc_b200 as g  # noqa: E402
import sanitizer_job as jobs  # noqa: E402

lib = g._lib.load()
lib.gamc_debug_guard_check.restype = C.c_longlong
lib.gamc_debug_guard_check.argtypes = [C.POINTER(C.c_longlong)]
results = []
_close = g.Engine.close


def checked_close(self):
    if getattr(self, "h", None):
        n = C.c_longlong(0)
        bad = lib.gamc_debug_guard_check(C.byref(n))
        results.append((int(bad), int(n.value)))
    _close(self)


g.Engine.close = checked_close
for which in ("single", "batched", "bigd"):
    before = len(results)
    getattr(jobs, which)()
    bad = sum(r[0] for r in results[before:])
    print(f"GUARD_CHECK {which}: {len(results) - before} handles checked, {max((r[1] for r in results[before:]), default=0)} live allocations at most, "
          f"{bad} guard bytes overwritten")
total = sum(r[0] for r in results)
print(f"GUARD_CHECK {'PASS' if total == 0 and results else 'FAIL'}")
sys.exit(0 if total == 0 and results else 1)
