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


def test_local_block_generated_in_place_matches_the_copied_hand_over():
    """chase_b200_dist_device_matrix_ / _mark_device_matrix_ (blocks that fit only once in HBM: C4 on 2 GPUs): filling
    the solver's own buffer with bench_dist.fill_local_block gives the same solve as generating the block separately and
    handing it over with chase_b200_dist_load_device_matrix_."""
    import numpy as np

    from chase_b200 import bench_dist as bd
    from chase_b200 import dist as cd

    import ctypes

    import chase_b200

    world = cd.World(0, 1, 0)
    N, nev, nex, nb = 1800, 60, 30, 64
    gr, gc = cd.global_indices(N, 1, nb, 0), cd.global_indices(N, 1, nb, 0)
    res = []
    try:
        for in_place in (False, True):
            s = cd.PChASE(world, N, nev, nex, np.complex128, grid=(1, 1), major="R", mb=nb, nb=nb)
            if in_place:
                ptr, ld = s.device_matrix()
                lam = bd.fill_local_block(ptr, ld, N, gr, gc, True, "cuda:0", chunk=500)
                s.mark_device_matrix()
            else:
                At, lam = bd.local_block(N, gr, gc, True, "cuda:0", transposed=True)
                s.load_device_matrix(At.data_ptr(), len(gr))
                del At
            res.append(s.solve(copy=True))
            s.finalize()
    finally:
        # the device hand-over switches the process-global "matrix resident" mode on: later solves in this process
        # must read their host matrices again
        chase_b200.lib().chase_b200_set_matrix_resident_(ctypes.byref(ctypes.c_int(0)))
        world.close()
    a, b = res
    assert a.iterations == b.iterations and a.filtered_vecs == b.filtered_vecs
    assert np.max(np.abs(a.ritzv[:nev] - lam[:nev]) / lam[:nev]) < 1e-10
    assert np.max(np.abs(a.ritzv[:nev] - b.ritzv[:nev])) < 1e-12
