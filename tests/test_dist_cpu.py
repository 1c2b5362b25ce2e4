"""CPU checks of the distributed backend's host logic: the index maps of the reference's block / block-cyclic layouts
(linalg/distMatrix/distMatrix.hpp:44-67 numroc, :1992-2039 block), the process grid (grid/mpiGrid2D.hpp:402-430), the
all-gather based layout change, and the world_size-2 launcher plumbing over gloo (no GPU, no NCCL calls)."""
import ctypes
import os

import numpy as np
import pytest

from chase_b200 import dist as cd
from chase_b200 import lib


def ref_numroc(n, nb, iproc, nprocs):
    """ScaLAPACK NUMROC, isrcproc = 0 (restated from the reference's chase::numroc)."""
    nblocks = n // nb
    loc = (nblocks // nprocs) * nb
    extra = nblocks % nprocs
    if iproc < extra:
        loc += nb
    elif iproc == extra:
        loc += n % nb
    return loc


def ref_block(n, nprocs, p):
    ln = n // nprocs if n % nprocs == 0 else min(n, n // nprocs + 1)
    off = p * ln
    size = ln if p < nprocs - 1 else n - (nprocs - 1) * ln
    return off, size


@pytest.mark.parametrize("N,nprocs,nb", [(1001, 4, 64), (1001, 2, 64), (256, 2, 32), (120000, 4, 64), (100, 3, 7),
                                         (64, 2, 64), (63, 2, 64)])
def test_block_cyclic_matches_numroc(N, nprocs, nb):
    allidx = []
    for p in range(nprocs):
        assert cd.local_size(N, nprocs, nb, p) == ref_numroc(N, nb, p, nprocs)
        g = cd.global_indices(N, nprocs, nb, p)
        assert len(g) == ref_numroc(N, nb, p, nprocs)
        assert all((x // nb) % nprocs == p for x in g[:: max(1, len(g) // 50)])
        assert np.all(np.diff(g) > 0)
        allidx.append(g)
    assert np.array_equal(np.sort(np.concatenate(allidx)), np.arange(N))


@pytest.mark.parametrize("N,nprocs", [(1001, 4), (256, 2), (20000, 4), (20000, 2), (7, 1), (1001, 3)])
def test_block_layout_matches_reference(N, nprocs):
    for p in range(nprocs):
        off, size = ref_block(N, nprocs, p)
        assert cd.local_size(N, nprocs, 0, p) == size
        g = cd.global_indices(N, nprocs, 0, p)
        assert np.array_equal(g, off + np.arange(size))


def test_grid_coords_and_dims():
    assert [cd.grid_dims(n) for n in (1, 2, 4, 8, 6)] == [(1, 1), (2, 1), (2, 2), (4, 2), (3, 2)]
    # row-major: rank = i * c + j ; column-major: rank = j * r + i   (MPI_Cart_create order, mpiGrid2D.hpp:402-430)
    for rank in range(8):
        assert cd.grid_coords(4, 2, "R", rank) == (rank // 2, rank % 2)
        assert cd.grid_coords(4, 2, "C", rank) == (rank % 4, rank // 4)
    with pytest.raises(ValueError):
        cd.grid_coords(2, 4, "R", 0)  # the reference requires row_dim >= col_dim


@pytest.mark.parametrize("N,r,mb,c,nb", [(1001, 2, 0, 2, 0), (1001, 4, 64, 2, 64), (300, 2, 0, 1, 0), (300, 3, 16, 2, 0),
                                         (257, 2, 32, 2, 8)])
def test_redistribution_map_col_to_row_layout(N, r, mb, c, nb):
    """Simulate: every grid row holds its rows of a vector x; all-gather (stride = padded max piece); the map must
    deliver exactly the rows of the column distribution."""
    f = lib().chase_b200_redistribution_map
    f.argtypes = [ctypes.c_longlong, ctypes.c_int, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int,
                  ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p]
    x = np.arange(N, dtype=np.float64) * 1.5 + 3
    stride = (max(cd.local_size(N, r, mb, p) for p in range(r)) + 15) // 16 * 16
    stacked = np.full(r * stride, np.nan)
    for p in range(r):
        g = cd.global_indices(N, r, mb, p)
        stacked[p * stride:p * stride + len(g)] = x[g]
    for pd in range(c):
        gd = cd.global_indices(N, c, nb, pd)
        out = np.full(len(gd), -1, dtype=np.int64)
        f(N, r, mb, stride, c, nb, pd, out.ctypes.data_as(ctypes.c_void_p))
        assert np.array_equal(stacked[out], x[gd])
    # and into global order (destination = one process owning everything)
    out = np.full(N, -1, dtype=np.int64)
    f(N, r, mb, stride, 1, 0, 0, out.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(stacked[out], x)


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the id exchange of chase_b200.dist.World (rank 0 creates, everybody receives) with a stand-in id
        ident = [bytes(range(128))] if rank == 0 else [None]
        dist.broadcast_object_list(ident, src=0)
        assert ident[0] == bytes(range(128))
        # 2 x 1 grid over a Clement matrix: the local blocks of all ranks tile the global matrix exactly
        N = 301
        r, c = cd.grid_dims(world)
        i, j = cd.grid_coords(r, c, "R", rank)
        pieces = {}
        for mb in (0, 32):
            gr, gc = cd.global_indices(N, r, mb, i), cd.global_indices(N, c, mb, j)
            gathered = [None] * world
            dist.all_gather_object(gathered, (i, j, gr.tolist(), gc.tolist()))
            cover = np.zeros((N, N), dtype=int)
            for (_, _, rr, cc) in gathered:
                cover[np.ix_(rr, cc)] += 1
            pieces[mb] = bool(np.all(cover == 1))
        q.put((rank, pieces))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_plumbing():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(all(v.values()) for _, v in res)


def test_sequence_generator_has_the_stated_spectrum_and_fixed_eigenvectors():
    """chase_b200.bench_dist.sequence_spectrum / local_block(lam=...): the correlated-sequence matrices of BASELINE
    config C3 (scripts/run_dist.py --seq) keep Q and perturb the known spectrum by 1e-4 relative per step."""
    from chase_b200 import bench_dist as bd

    N = 200
    idx = np.arange(N)
    A0, lam0 = bd.local_block(N, idx, idx, True, "cpu")
    lam2 = bd.sequence_spectrum(N, 2)
    A2, lam = bd.local_block(N, idx, idx, True, "cpu", lam=lam2)
    assert np.array_equal(lam, lam2) and np.array_equal(lam0, bd.sequence_spectrum(N, 0))
    assert 0 < np.max(np.abs(lam2 / lam0 - 1)) < 1e-3
    w, Q = np.linalg.eigh(A0.numpy())
    assert np.max(np.abs(w - lam0)) < 1e-11
    # same eigenvectors: Q^H A2 Q is diagonal with the perturbed spectrum
    D = Q.conj().T @ A2.numpy() @ Q
    assert np.max(np.abs(D - np.diag(np.diag(D)))) < 1e-10
    assert np.max(np.abs(np.sort(np.diag(D).real) - np.sort(lam2))) < 1e-11
