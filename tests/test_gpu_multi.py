"""Multi-GPU tests (need >= 2 B200 on the box; skipped otherwise): tools/multigpu_check.py under torchrun, one rank per GPU --
row-sharded evaluation and chains against the single-GPU result and the CPU oracle, the NCCL fallback of the peer-memory
reductions, and the bounded wait for a rank that stopped."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _torchrun(n, mode, port, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "multigpu_check.py"), mode]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, **(env or {})))


@pytest.mark.parametrize("mode,port", [("parity", 29531), ("fallback", 29532), ("timeout", 29533)])
def test_two_ranks(mode, port):
    if _ngpus() < 2:
        pytest.skip("needs two GPUs")
    out = _torchrun(2, mode, port)
    assert out.returncode == 0 and "PASS" in out.stdout, (out.stdout[-1500:], out.stderr[-1500:])


def test_two_ranks_integer_metric_route():
    """The same parity job with the metric forced onto the integer tensor cores on every rank (GAMC_METRIC_I8=1; the test sizes are
    below the automatic threshold): every rank cuts its own row shard into digit planes with its own column exponents, and the
    reduced metric still equals the single-GPU evaluation of all rows and the oracle."""
    if _ngpus() < 2:
        pytest.skip("needs two GPUs")
    out = _torchrun(2, "parity", 29536, env={"GAMC_METRIC_I8": "1"})
    assert out.returncode == 0 and "PASS" in out.stdout, (out.stdout[-1500:], out.stderr[-1500:])


def test_all_ranks_parity():
    n = _ngpus()
    if n < 4:
        pytest.skip("needs four or more GPUs")
    out = _torchrun(n, "parity", 29534)
    assert out.returncode == 0 and "PASS" in out.stdout, (out.stdout[-1500:], out.stderr[-1500:])


def test_chains_sharded_across_ranks():
    """SURVEY 8e, independent chains: every rank owns a block of chains, no collective; the device random streams are keyed by the
    GLOBAL chain index, so ranks draw distinct numbers (tools/bench_c4.py asserts it) and the aggregate rate is the sum."""
    import json
    if _ngpus() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29535",
           os.path.join(ROOT, "tools", "bench_c4.py"), "--dim", "64", "--chains", "16", "--steps", "40", "--cpu-steps", "0"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-1500:])
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["n_gpus"] == 2 and d["value"] > 0 and "chains sharded x2" in d["config"]["parallelism"]
