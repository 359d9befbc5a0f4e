"""Multi-GPU parity check (run under torchrun, one rank per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py
Row-shards a synthetic logistic problem over N ranks, and checks (i) the all-reduced target / gradient / metric
against the single-GPU evaluation of the whole problem and against the CPU oracle, (ii) that every rank runs the
identical GAMC trajectory (replicated dense algebra on bit-identical reduced inputs), equal to the oracle's."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gamc_b200 as g  # noqa: E402
import oracle as o  # noqa: E402
from gamc_b200.distributed import init_comm, shard_rows  # noqa: E402


def relerr(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, p, seed = 20000 + 7, 96, 20260101
    tt = o.logistic.synth_theta_true(seed, p)
    theta = tt + 0.03 * np.random.default_rng(1).standard_normal(p)
    r0, r1 = shard_rows(N, world, rank)
    eng = g.Engine(local)
    eng.target_logistic_synth(seed, r0, r1 - r0, p, tt, 100.0)
    init_comm(eng, rank, world)
    lt, grad, G = eng.eval_upto_tensor(theta)
    # oracle on the whole problem
    X, y, _ = o.synth_logistic(seed, N, p, theta_true=tt)
    lt0, g0, G0 = o.LogisticTarget(X, y, 100.0).uptotensorlogtarget(theta)
    ok = abs(lt - lt0) <= 1e-10 * abs(lt0) and relerr(grad, g0) <= 1e-10 and relerr(G, G0) <= 1e-10
    # sharded partial of this rank against the oracle's partial of the same rows
    ll, gl, Gl = eng.eval_partials(theta)
    ll0, gl0, Gl0 = o.logistic_partials(theta, X[r0:r1], y[r0:r1])
    ok = ok and abs(ll - ll0) <= 1e-10 * abs(ll0) and relerr(gl, gl0) <= 1e-10 and relerr(Gl, Gl0) <= 1e-10
    # trajectory: identical on every rank and equal to the oracle
    nsteps = 600
    for am_mode in (0, 1, 2):
        sampler = g.GAMC(update=g.rand_exp_decay_update(10.0), driftstep=0.02, minorscale=0.001, c=0.01, am_mode=am_mode)
        tuner = g.GAMCTuner(g.VanillaMCTuner(), g.VanillaMCTuner(), g.AcceptanceRateMCTuner(0.35))
        tape = g.RandomTape(nsteps, p, seed=3)
        job = g.BasicMCJob(g.LogisticModel(lam=100.0), sampler, g.BasicMCRange(nsteps, 100), {"p": theta}, tuner=tuner, tape=tape, engine=eng)
        job.run()
        chain = job.output()
        osam = o.GAMCOracle(o.LogisticTarget(X, y, 100.0), theta, nsteps, 100, update="rand_exp_decay_update!", update_params=(10.0, 0.0),
                            driftstep=0.02, minorscale=0.001, c=0.01,
                            tuner=o.GAMCTuner(o.VanillaMCTuner(), o.VanillaMCTuner(), o.AcceptanceRateMCTuner(0.35)))
        v, a, _ = osam.run(o.RandomTape(nsteps, p, seed=3))
        ok = ok and np.array_equal(chain.diagnosticvalues.ravel(), a) and relerr(chain.value, v) <= 1e-8
        # bitwise agreement across ranks
        t = torch.from_numpy(np.ascontiguousarray(chain.value)).cuda()
        ref = t.clone()
        dist.broadcast(ref, src=0)
        ok = ok and bool(torch.equal(t, ref))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"MULTIGPU_CHECK world={world} {'PASS' if flag.item() == 1 else 'FAIL'} relerr_G={relerr(G, G0):.2e} relerr_g={relerr(grad, g0):.2e} "
              f"am_scalar_allreduce={'peer-mailbox' if eng.comm_peer_path() & 1 else 'nccl'} "
              f"metric_allreduce={'peer-landing' if eng.comm_peer_path() & 2 else 'nccl'}")
    eng.close()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
