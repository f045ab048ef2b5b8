"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): launches
tests/multi_rank_check.py under torchrun, one rank per GPU, NCCL halo between sub-domains."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _run(nproc, kind, cells, steps, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           str(ROOT / "tests" / "multi_rank_check.py"), kind, str(cells), str(steps)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "MULTI_RANK_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("kind,cells,steps", [("lj", 12, 100), ("eam", 8, 60)])
def test_two_ranks_match_oracle(kind, cells, steps):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, kind, cells, steps, 29531)


def test_eight_ranks_match_oracle():
    if _ngpu() < 8:
        pytest.skip("needs 8 GPUs")
    _run(8, "lj", 16, 100, 29532)
