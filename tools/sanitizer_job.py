"""Small jobs that touch every kernel family, for `compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitizer_job.py [which]`.
which = single (p = 37 logistic chain, all three am_modes, MALA) | batched (t / RV / logistic persistent kernels) | bigd (structured t target) |
multi (run under torchrun with 2 ranks: peer-memory reductions)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gamc_b200 as g  # noqa: E402


def cfg(nsteps, am_mode):
    sampler = g.GAMC(update=g.rand_exp_decay_update(10.0), driftstep=0.02, minorscale=0.001, c=0.01, am_mode=am_mode)
    tuner = g.GAMCTuner(g.VanillaMCTuner(), g.VanillaMCTuner(), g.AcceptanceRateMCTuner(0.35))
    return sampler, tuner, g.BasicMCRange(nsteps, nsteps // 5)


def single(eng=None, rank=0, world=1):
    rng = np.random.default_rng(0)
    N, p, nsteps = 1500, 37, 50
    tt = g.synth_theta_true(1, p)
    eng = eng or g.Engine(int(os.environ.get("LOCAL_RANK", "0")))
    r0, r1 = (rank * N) // world, ((rank + 1) * N) // world
    eng.target_logistic_synth(1, r0, r1 - r0, p, tt, 100.0)
    if world > 1:
        from gamc_b200.distributed import init_comm
        init_comm(eng, rank, world)
    for am_mode in (0, 1, 2):
        sampler, tuner, rngk = cfg(nsteps, am_mode)
        job = g.BasicMCJob(g.LogisticModel(lam=100.0), sampler, rngk, {"p": tt + 0.05 * rng.standard_normal(p) * 0 + 0.01}, tuner=tuner,
                           tape=g.RandomTape(nsteps, p, seed=2), engine=eng)
        job.run()
        job.output()
    if world == 1:
        job = g.BasicMCJob(g.LogisticModel(lam=100.0), g.MALA(0.02), g.BasicMCRange(30, 5), {"p": tt}, tuner=g.AcceptanceRateMCTuner(0.574, score_k=3.0),
                           tape=g.RandomTape(30, p, seed=3), engine=eng)
        job.run()
        eng.eval_upto_tensor(tt)
    eng.close()


def batched():
    rng = np.random.default_rng(1)
    sampler = g.GAMC(update=g.rand_exp_decay_update(10.0), transform=g.softabs(1000.0), driftstep=0.02, minorscale=0.001, c=0.01)
    tuner = g.GAMCTuner(g.VanillaMCTuner(), g.VanillaMCTuner(), g.AcceptanceRateMCTuner(0.35))
    job = g.BatchedMCJob(g.MvTModel(n=20), sampler, g.BasicMCRange(60, 10), 2.0 * rng.standard_normal((4, 20)), tuner=tuner, seed=1)
    job.run(); job.output(); job.engine.close()
    X = rng.standard_normal((200, 4)); y = (rng.random(200) < 0.5).astype(np.float64)
    s2 = g.GAMC(update=g.rand_exp_decay_update(10.0), driftstep=0.02, minorscale=0.001, c=0.01)
    job = g.BatchedMCJob(g.LogisticModel(X, y, 100.0), s2, g.BasicMCRange(80, 10), 0.1 * rng.standard_normal((3, 4)), tuner=tuner, seed=2)
    job.run(); job.output(); job.engine.close()
    m = g.RVModel.simulate(g.make_param_true_ex1(), num_obs=50, seed=3)
    v0 = g.make_param_true_ex1() + 1e-4 * rng.standard_normal((2, 6))
    job = g.BatchedMCJob(m, sampler, g.BasicMCRange(30, 5), v0, tuner=tuner, seed=3)
    job.run(); job.output(); job.engine.close()


def bigd():
    rng = np.random.default_rng(2)
    sampler = g.GAMC(update=g.mod_update(3), transform=g.softabs(1000.0), driftstep=0.02, minorscale=0.001, c=0.01)
    tuner = g.GAMCTuner(g.VanillaMCTuner(), g.VanillaMCTuner(), g.AcceptanceRateMCTuner(0.35))
    for n, scale in ((40, 2.0), (130, 2.0)):
        job = g.BatchedMCJob(g.MvTModel(n=n), sampler, g.BasicMCRange(24, 4), scale * rng.standard_normal((3, n)), tuner=tuner, seed=4)
        job.run(); job.output(); job.engine.close()


def multi():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    single(rank=rank, world=world)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "single"
    {"single": single, "batched": batched, "bigd": bigd, "multi": multi}[which]()
    print(f"SANITIZER_JOB {which} done")
