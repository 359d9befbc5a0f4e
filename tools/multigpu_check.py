"""Multi-GPU parity check (run under torchrun, one rank per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py [mode]
mode = parity (default) | fallback (the same job with the peer paths switched off: NCCL carries the reductions) |
       timeout (a rank stops early: the others must report GAMC_NCCL_ERROR after the configured wait instead of hanging)
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


def run_chain(eng, theta, p, nsteps, am_mode, n_run=None):
    sampler = g.GAMC(update=g.rand_exp_decay_update(10.0), driftstep=0.02, minorscale=0.001, c=0.01, am_mode=am_mode)
    tuner = g.GAMCTuner(g.VanillaMCTuner(), g.VanillaMCTuner(), g.AcceptanceRateMCTuner(0.35))
    tape = g.RandomTape(nsteps, p, seed=3)
    job = g.BasicMCJob(g.LogisticModel(lam=100.0), sampler, g.BasicMCRange(nsteps, 100), {"p": theta}, tuner=tuner, tape=tape, engine=eng)
    if n_run is None:
        job.run()
        return job.output()
    eng.tape_set(tape.u_sched, tape.u_mix, tape.z, tape.u_acc)
    eng.run(n_run)
    eng.sync()
    return None


def mode_fallback(rank, world, local):
    """GAMC_NO_P2P / GAMC_NO_P2P_LAND: NCCL carries the reductions.  Same accept decisions as the peer-memory paths, chain equal to
    rounding (NCCL's summation order is not the rank order), bit-identical across ranks within each path."""
    N, p, seed, nsteps = 20000 + 7, 96, 20260101, 400
    tt = o.logistic.synth_theta_true(seed, p)
    theta = tt + 0.03 * np.random.default_rng(1).standard_normal(p)
    r0, r1 = shard_rows(N, world, rank)
    ok, chains, paths = True, [], []
    for env in ({}, {"GAMC_NO_P2P_LAND": "1"}, {"GAMC_NO_P2P": "1"}):
        for k in ("GAMC_NO_P2P", "GAMC_NO_P2P_LAND"):
            os.environ.pop(k, None)
        os.environ.update(env)
        eng = g.Engine(local)
        eng.target_logistic_synth(seed, r0, r1 - r0, p, tt, 100.0)
        init_comm(eng, rank, world)
        chain = run_chain(eng, theta, p, nsteps, 2)
        paths.append(eng.comm_peer_path())
        t = torch.from_numpy(np.ascontiguousarray(chain.value)).cuda()
        ref = t.clone()
        dist.broadcast(ref, src=0)
        ok = ok and bool(torch.equal(t, ref))
        chains.append(chain)
        dist.barrier()
        eng.close()
    ok = ok and paths == [3, 1, 0]
    for c in chains[1:]:
        ok = ok and np.array_equal(c.diagnosticvalues, chains[0].diagnosticvalues) and relerr(c.value, chains[0].value) <= 1e-10
    return ok, f"paths={paths}"


def mode_timeout(rank, world, local):
    """Rank `world-1` stops after 5 transitions; the others wait `timeout` seconds inside the step kernel, then report
    GAMC_NCCL_ERROR at the next synchronising call."""
    import time
    N, p, seed, nsteps = 4000, 24, 20260101, 40
    tt = o.logistic.synth_theta_true(seed, p)
    r0, r1 = shard_rows(N, world, rank)
    eng = g.Engine(local)
    eng.target_logistic_synth(seed, r0, r1 - r0, p, tt, 100.0)
    init_comm(eng, rank, world)
    eng.comm_set_timeout(1.5)
    quitter = rank == world - 1
    t0 = time.perf_counter()
    ok, msg = True, ""
    try:
        run_chain(eng, tt, p, nsteps, 2, n_run=5 if quitter else nsteps)
        ok = quitter
    except g._lib.NcclError as e:
        waited = time.perf_counter() - t0
        ok = (not quitter) and 1.0 <= waited <= 20.0 and "timed out" in str(e)
        msg = f"rank {rank}: NcclError after {waited:.1f} s ({e})"
    dist.barrier()
    eng.close()
    return ok, msg


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mode = sys.argv[1] if len(sys.argv) > 1 else "parity"
    if mode in ("fallback", "timeout"):
        ok, msg = (mode_fallback if mode == "fallback" else mode_timeout)(rank, world, local)
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if msg and (rank == 0 or not ok):
            print(msg)
        if rank == 0:
            print(f"MULTIGPU_CHECK[{mode}] world={world} {'PASS' if flag.item() == 1 else 'FAIL'}")
        dist.destroy_process_group()
        sys.exit(0 if flag.item() == 1 else 1)
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
    for am_mode in (0, 1, 2, 3):
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
