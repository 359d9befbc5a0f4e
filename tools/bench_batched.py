"""Throughput of the batched-chains kernel on the reference's t-distribution example scaled out to many chains
(BASELINE.json configs[3] family; d <= 32 here -- d = 1024 needs the dense large-p batched path).
  python tools/bench_batched.py [--dim 20] [--chains 4096] [--steps 2000]
Prints one JSON line: chain-iterations per second (device-timed, Philox tape generated in the kernel) and the CPU oracle's rate."""
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
    ap.add_argument("--dim", dest="n", type=int, default=20, help="dimension of the t target")
    ap.add_argument("--chains", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--cpu-steps", type=int, default=1500)
    ap.add_argument("--workload", default="t", choices=["t", "rv"])
    ap.add_argument("--logN", type=int, default=18, help="rv: log2 of the number of epochs")
    a = ap.parse_args()
    import torch
    rng = np.random.default_rng(0)
    if a.workload == "rv":
        return main_rv(a, rng, torch)
    # chains sharded across ranks (torchrun): a.chains is PER GPU, no data-path collective
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rng = np.random.default_rng(rank)
    v0s = 2.0 * rng.standard_normal((a.chains, a.n))
    sampler = g.GAMC(update=g.rand_exp_decay_update(10.0), transform=g.softabs(1000.0), driftstep=0.02, minorscale=0.001, c=0.01)
    tuner = g.GAMCTuner(g.VanillaMCTuner(), g.VanillaMCTuner(), g.AcceptanceRateMCTuner(0.35))
    eng = g.Engine(local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    rngk = g.BasicMCRange(a.steps, a.steps // 5)
    off = rank * a.chains
    job = g.BatchedMCJob(g.MvTModel(n=a.n), sampler, g.BasicMCRange(50, 0), v0s, tuner=tuner, seed=1, engine=eng, chain_offset=off)
    job.run()                                           # warm-up
    job = g.BatchedMCJob(g.MvTModel(n=a.n), sampler, rngk, v0s, tuner=tuner, seed=1, engine=eng, chain_offset=off)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    eng.batched_run(a.steps, 1)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    first_sample = job.output()[0].value[:, 0].copy()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        fs = [None] * world
        dist.all_gather_object(fs, first_sample.tolist())
        assert len({tuple(x) for x in fs}) == world, "ranks drew identical random streams"
        if rank != 0:
            eng.close(); dist.destroy_process_group(); return
    chains = job.output()
    acc = float(np.mean([c.diagnosticvalues.mean() for c in chains]))
    # CPU oracle: one chain
    import oracle as o
    from oracle import targets as T
    tgt = T.TTarget(n=a.n, nu=30.0)
    sam = o.GAMCOracle(tgt, v0s[0], a.cpu_steps, 0, update="rand_exp_decay_update!", update_params=(10.0, 0.0), transform=lambda H: T.softabs(H, 1000.0),
                       driftstep=0.02, minorscale=0.001, c=0.01, tuner=o.GAMCTuner(o.VanillaMCTuner(), o.VanillaMCTuner(), o.AcceptanceRateMCTuner(0.35)))
    t0 = time.perf_counter()
    sam.run(o.RandomTape(a.cpu_steps, a.n, seed=1))
    cpu = a.cpu_steps / (time.perf_counter() - t0)
    print(json.dumps({"metric": "gamc_chain_iters_per_sec", "value": world * a.chains * a.steps / (ms * 1e-3), "unit": "chain-iters/s", "n_gpus": world,
                      "config": {"workload": f"t-distribution example, n={a.n}, nu=30, softabs(1000), {a.chains} chains per GPU, GAMC rand_exp_decay(a=10)",
                                 "parallelism": f"chains sharded x{world}, no collective"},
                      "ms_total": ms, "steps": a.steps, "acceptance": acc, "dtype": "f64",
                      "cpu_baseline": {"value": cpu, "unit": "chain-iters/s", "cores": 1, "kind": "port", "sample": f"{a.cpu_steps} oracle transitions of one chain"}}))
    eng.close()


def main_rv(a, rng, torch):
    """BASELINE.json configs[4]: Keplerian RV model on synthetic observations (N = 2^18 epochs), one planet, many chains."""
    import oracle as o
    from oracle import targets as T
    N = 1 << a.logN
    truth = T.make_param_true_ex1()
    t = np.sort(rng.uniform(0.0, 730.5, N))                   # one_planet/data_generation.jl:16-18
    sig = 2.0 * np.ones(N)                                    # :20-21
    obs = T.RVModel(t, np.zeros(N), sig).model_rv(truth) + 2.0 * rng.standard_normal(N)   # :26
    scale = np.array([0.0001, 0.01, 0.05, 0.05, np.pi * 0.01, 0.3])
    v0s = truth + 0.01 * scale * rng.standard_normal((a.chains, 6)) / np.sqrt(N / 50.0)
    sampler = g.GAMC(update=g.rand_exp_decay_update(10.0), transform=g.softabs(1000.0), driftstep=0.02, minorscale=0.001, c=0.01)
    tuner = g.GAMCTuner(g.VanillaMCTuner(), g.VanillaMCTuner(), g.AcceptanceRateMCTuner(0.35))
    eng = g.Engine(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    model = g.RVModel(t, obs, sig, 1)
    job = g.BatchedMCJob(model, sampler, g.BasicMCRange(8, 0), v0s, tuner=tuner, seed=1, engine=eng)
    job.run()
    job = g.BatchedMCJob(model, sampler, g.BasicMCRange(a.steps, a.steps // 5), v0s, tuner=tuner, seed=1, engine=eng)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    eng.batched_run(a.steps, 1)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    chains = job.output()
    acc = float(np.mean([c.diagnosticvalues.mean() for c in chains]))
    st, _ = eng.batched_state(0)
    if a.cpu_steps == 0:     # --cpu-steps 0: device measurement only
        print(json.dumps({"metric": "gamc_chain_iters_per_sec", "value": a.chains * a.steps / (ms * 1e-3), "unit": "chain-iters/s", "n_gpus": 1,
                          "ms_total": ms, "steps": a.steps, "acceptance": acc, "epoch_evals_per_sec": a.chains * a.steps * N / (ms * 1e-3)}))
        eng.close()
        return
    # CPU oracle, one chain, a few transitions at full N
    om = T.RVModel(t, obs, sig)
    ncpu = 12
    sam = o.GAMCOracle(om, v0s[0], ncpu, 0, update="mod_update!", update_params=(4,), transform=lambda H: T.softabs(H, 1000.0), driftstep=0.02,
                       minorscale=0.001, c=0.01, tuner=o.GAMCTuner(o.VanillaMCTuner(), o.VanillaMCTuner(), o.AcceptanceRateMCTuner(0.35)))
    t0 = time.perf_counter()
    sam.run(o.RandomTape(ncpu, 6, seed=1))
    cpu = ncpu / (time.perf_counter() - t0)
    print(json.dumps({"metric": "gamc_chain_iters_per_sec", "value": a.chains * a.steps / (ms * 1e-3), "unit": "chain-iters/s", "n_gpus": 1,
                      "config": {"workload": f"radial-velocity example, one planet, N=2^{a.logN} epochs, softabs(1000), {a.chains} chains, GAMC rand_exp_decay(a=10)"},
                      "ms_total": ms, "steps": a.steps, "acceptance": acc, "n_smmala_chain0": int(st.updatetensorcount), "dtype": "f64",
                      "epoch_evals_per_sec": a.chains * a.steps * N / (ms * 1e-3),
                      "cpu_baseline": {"value": cpu, "unit": "chain-iters/s", "cores": os.cpu_count(), "kind": "port",
                                       "sample": f"{ncpu} oracle transitions (mod_update! n=4: 6 SMMALA / 6 AM) of one chain at full N"}}))
    eng.close()


if __name__ == "__main__":
    main()
