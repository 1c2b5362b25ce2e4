"""GPU tests of the distributed backend (pChASEGPU) through the reference's distributed C interface.

On the 1-GPU box the whole distributed code path (layouts, NCCL communicators of size 1, alternating A^H / A HEMM,
all-gather redistribution, cached residual block, replicated Lanczos) runs on a 1 x 1 grid with block and block-cyclic
layouts; with >= 2 GPUs the same checks run under torchrun on real grids (tests/dist_check.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc, extra=()):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + nproc), os.path.join(ROOT, "tests", "dist_check.py"),
           *extra]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-4000:] + p.stderr[-4000:]
    assert "FAIL" not in p.stdout
    return p.stdout


def test_single_rank_grid_matches_reference_traces():
    out = _run(1)
    assert out.count("ok   ") >= 8


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_ranks():
    _run(2)


@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs 4 GPUs")
def test_four_ranks():
    _run(4)
