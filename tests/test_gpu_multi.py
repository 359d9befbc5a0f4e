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


def _torchrun(n, mode, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "multigpu_check.py"), mode]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)


@pytest.mark.parametrize("mode,port", [("parity", 29531), ("fallback", 29532), ("timeout", 29533)])
def test_two_ranks(mode, port):
    if _ngpus() < 2:
        pytest.skip("needs two GPUs")
    out = _torchrun(2, mode, port)
    assert out.returncode == 0 and "PASS" in out.stdout, (out.stdout[-1500:], out.stderr[-1500:])


def test_all_ranks_parity():
    n = _ngpus()
    if n < 4:
        pytest.skip("needs four or more GPUs")
    out = _torchrun(n, "parity", 29534)
    assert out.returncode == 0 and "PASS" in out.stdout, (out.stdout[-1500:], out.stderr[-1500:])
