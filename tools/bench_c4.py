"""BASELINE.json configs[3] ("C4"): the t-distribution example scaled to d = 1024, 4096 independent GAMC chains per GPU, chains
sharded across GPUs (no collective).  Target and sampler as in examples/t_distribution/src/GAMC/simulations.jl:17-59: nu = 30,
Sigma_ij = 0.9^|i-j|, transform = softabs(1000), rand_exp_decay_update!(a = 10), driftstep 0.02, minorscale 0.001, c 0.01,
AcceptanceRateMCTuner(0.35), initial values N(0, 2^2) (:57).

  python tools/bench_c4.py [--dim 1024] [--chains 4096] [--steps 400] [--cpu-steps 6]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_c4.py     # chains x 8
Prints one JSON line: chain-iterations per second (device-timed), time per kernel family, and the dense CPU oracle's rate."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gamc_b200 as g  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--chains", type=int, default=4096, help="chains PER GPU")
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--cpu-steps", type=int, default=6)
    ap.add_argument("--init-scale", type=float, default=2.0, help="initial values N(0, scale^2) (reference: 2)")
    ap.add_argument("--max-factors", type=int, default=0)
    ap.add_argument("--init", default="ref", choices=["ref", "typical"],
                    help="ref: i.i.d. N(0, scale^2) as in the reference (far from the mode: ~127 eigenvalues below 21/alpha at d = 1024); "
                         "typical: draws from N(0, Sigma), the regime a long run spends its time in (one such eigenvalue)")
    a = ap.parse_args()
    import torch
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.dim
    rng = np.random.default_rng(1000 + rank)
    v0s = a.init_scale * rng.standard_normal((a.chains, n))
    if a.init == "typical":                                      # AR(1) recursion: x ~ N(0, Sigma), Sigma_ij = 0.9^|i-j|
        z = rng.standard_normal((a.chains, n))
        v0s = np.empty_like(z)
        v0s[:, 0] = z[:, 0]
        for i in range(1, n):
            v0s[:, i] = 0.9 * v0s[:, i - 1] + np.sqrt(1.0 - 0.81) * z[:, i]
    sampler = g.GAMC(update=g.rand_exp_decay_update(10.0), transform=g.softabs(1000.0), driftstep=0.02, minorscale=0.001, c=0.01)
    tuner = g.GAMCTuner(g.VanillaMCTuner(), g.VanillaMCTuner(), g.AcceptanceRateMCTuner(0.35))
    eng = g.Engine(local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    off = rank * a.chains
    keep = 4                                                     # saved samples per chain (the chain buffer is [chains][keep][d])
    job = g.BatchedMCJob(g.MvTModel(n=n), sampler, g.BasicMCRange(12, 8), v0s, tuner=tuner, seed=1, engine=eng, chain_offset=off, max_factors=a.max_factors)
    job.run()                                                    # warm-up
    t_init0 = time.perf_counter()
    job = g.BatchedMCJob(g.MvTModel(n=n), sampler, g.BasicMCRange(a.steps, a.steps - keep), v0s, tuner=tuner, seed=1, engine=eng, chain_offset=off,
                         max_factors=a.max_factors)
    torch.cuda.synchronize()
    init_s = time.perf_counter() - t_init0
    eng.profile_enable(True)
    eng.profile_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    eng.bigd_run(a.steps, 1)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    prof = eng.profile_get()
    eng.profile_enable(False)
    chains = job.output()                                        # raises if any chain failed
    acc = float(np.mean([c.diagnosticvalues.mean() for c in chains]))
    nsm = [eng.bigd_state(c)[0].updatetensorcount for c in range(min(a.chains, 64))]
    rlow = [eng.bigd_state(c)[0].error_pivot for c in range(min(a.chains, 64))]
    first = chains[0].value[:4, 0].tolist()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        fs = [None] * world
        dist.all_gather_object(fs, first)
        assert len({tuple(x) for x in fs}) == world, "ranks drew identical random streams"
        if rank != 0:
            eng.close(); dist.destroy_process_group(); return
    fam = {k: {"launches": v[0], "ms": round(v[1], 3)} for k, v in prof.items() if v[0]}
    out = {"metric": "gamc_chain_iters_per_sec", "value": world * a.chains * a.steps / (ms * 1e-3), "unit": "chain-iters/s", "n_gpus": world,
           "config": {"workload": f"C4: t-distribution example, d={n}, nu=30, softabs(1000), {a.chains} chains per GPU, GAMC rand_exp_decay(a=10), "
                                  f"initial values {'N(0, %g^2) i.i.d. (reference)' % a.init_scale if a.init == 'ref' else 'N(0, Sigma) (typical set)'}, {a.steps} transitions", "parallelism": f"chains sharded x{world}, no collective"},
           "ms_total": ms, "ms_per_step_all_chains": ms / a.steps, "steps": a.steps, "acceptance_last_samples": acc, "dtype": "f64",
           "smmala_steps_per_chain_mean": float(np.mean(nsm)), "softabs_changed_eigenvalues_at_end_mean": float(np.mean(rlow)),
           "init_seconds": init_s, "kernel_families": fam,
           "hbm_state_bytes_per_gpu": int(a.chains * (n * n * 8 + 2 * (eng._cfg_ncmax if hasattr(eng, "_cfg_ncmax") else 5) * n * 32 * 8 * 2))}
    if a.cpu_steps > 0:
        import oracle as o
        from oracle import targets as T
        tgt = T.TTarget(n=n, nu=30.0)
        sam = o.GAMCOracle(tgt, v0s[0], a.cpu_steps, 0, update="mod_update!", update_params=(3,), transform=lambda H: T.softabs(H, 1000.0),
                           driftstep=0.02, minorscale=0.001, c=0.01, tuner=o.GAMCTuner(o.VanillaMCTuner(), o.VanillaMCTuner(), o.AcceptanceRateMCTuner(0.35)))
        t0 = time.perf_counter()
        sam.run(o.RandomTape(a.cpu_steps, n, seed=1))
        cpu = a.cpu_steps / (time.perf_counter() - t0)
        out["cpu_baseline"] = {"value": cpu, "unit": "chain-iters/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": f"{a.cpu_steps} dense oracle transitions (mod_update! n=3) of one chain at d={n}: eigh softabs, inv, cholesky"}
    print(json.dumps(out))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
