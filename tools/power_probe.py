"""Power / clock behaviour of the level-2 evaluation: an SMMALA-only chain (two level-2 evaluations per step) for a few seconds with
NVML sampled every 20 ms, for the serial route (GAMC_LEVEL2_CO=0) and the co-scheduled pair.
    python tools/power_probe.py [--steps 400]"""
import argparse, json, os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gamc_b200 as g  # noqa: E402
import pynvml  # noqa: E402


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(0)
        self.rows, self.stop = [], False

    def run(self):
        while not self.stop:
            self.rows.append((time.time(), pynvml.nvmlDeviceGetPowerUsage(self.h) / 1e3, pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM),
                              pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)))
            time.sleep(0.02)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--p", type=int, default=256)
    a = ap.parse_args()
    seed, N, p = 20260101, 1 << 20, a.p
    tt = g.synth_theta_true(seed, p)
    eng = g.Engine(0)
    for label, env in [("serial", {"GAMC_LEVEL2_CO": "0"}), ("co-scheduled", {}), ("co-scheduled, no throttle", {"GAMC_CO_LAG_MB": "0"}), ("serial again", {"GAMC_LEVEL2_CO": "0"})]:
        for k in ("GAMC_LEVEL2_CO", "GAMC_CO_LAG_MB"):
            os.environ.pop(k, None)
        os.environ.update(env)
        eng.target_logistic_synth(seed, 0, N, p, tt, 100.0)
        sampler = g.GAMC(update=g.smmala_only_update(), driftstep=0.02, minorscale=0.001, c=0.01)
        tuner = g.GAMCTuner(g.VanillaMCTuner(), g.VanillaMCTuner(), g.AcceptanceRateMCTuner(0.35))
        job = g.BasicMCJob(g.LogisticModel(lam=100.0), sampler, g.BasicMCRange(a.steps, a.steps // 5), {"p": tt.copy()}, tuner=tuner,
                           tape=g.RandomTape(a.steps, p, seed=3), engine=eng)
        smp = Sampler()
        smp.start()
        time.sleep(0.3)
        t0 = time.time()
        job.run()
        eng.sync()
        t1 = time.time()
        smp.stop = True
        smp.join()
        rows = [r for r in smp.rows if t0 + 0.3 <= r[0] <= t1 - 0.05]
        pw = np.array([r[1] for r in rows]); ck = np.array([r[2] for r in rows])
        reasons = 0
        for r in rows:
            reasons |= r[3]
        print(json.dumps({"config": label, "steps": a.steps, "ms_per_step": 1e3 * (t1 - t0) / a.steps, "samples": len(rows),
                          "power_w_mean": float(pw.mean()) if len(pw) else None, "power_w_max": float(pw.max()) if len(pw) else None,
                          "sm_mhz_mean": float(ck.mean()) if len(ck) else None, "sm_mhz_min": float(ck.min()) if len(ck) else None,
                          "throttle_reasons_or": hex(reasons)}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
