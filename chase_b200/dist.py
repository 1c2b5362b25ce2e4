"""Host-side mirror of the reference's distributed C interface (``p?chase_init_`` / ``p?chase_init_blockcyclic_`` /
``p?chase_`` / ``p?chase_finalize_``, ``/root/reference/interface/chase_c_interface.h:61-195``) for one process per
GPU, plus the launcher-side bootstrap the reference did with MPI (``grid/mpiGrid2D.hpp:449-485``): rank 0 makes an
NCCL unique id, ``torch.distributed`` ships it, every rank joins.

``torch.distributed`` is plumbing only (id exchange, barriers, max-over-ranks of timings); every collective on the
solver's data path is issued by the native library on its own NCCL communicators.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from ._lib import lib
from .solver import _PFX, _REAL, STAT_NAMES, SolveResult, _i, _p

ID_BYTES = 128


def grid_dims(nranks: int) -> tuple[int, int]:
    """r x c with r >= c, as square as possible (what MPI_Dims_create gives; the reference requires r >= c,
    grid/mpiGrid2D.hpp:209-211): 1 -> 1x1, 2 -> 2x1, 4 -> 2x2, 8 -> 4x2."""
    c = int(np.floor(np.sqrt(nranks)))
    while nranks % c:
        c -= 1
    return nranks // c, c


def local_size(N, nprocs, nb, p):
    f = lib().chase_b200_local_size
    f.restype = ctypes.c_longlong
    f.argtypes = [ctypes.c_longlong, ctypes.c_int, ctypes.c_longlong, ctypes.c_int]
    return int(f(N, nprocs, nb, p))


def global_indices(N, nprocs, nb, p) -> np.ndarray:
    n = local_size(N, nprocs, nb, p)
    out = np.zeros(max(n, 1), dtype=np.int64)
    f = lib().chase_b200_global_indices
    f.argtypes = [ctypes.c_longlong, ctypes.c_int, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p]
    f(N, nprocs, nb, p, out.ctypes.data_as(ctypes.c_void_p))
    return out[:n]


def grid_coords(dim0, dim1, major, rank):
    i, j = ctypes.c_int(-1), ctypes.c_int(-1)
    rc = lib().chase_b200_grid_coords(int(dim0), int(dim1), ctypes.c_char(major.encode()), int(rank),
                                      ctypes.byref(i), ctypes.byref(j))
    if rc != 0:
        raise ValueError("invalid grid")
    return i.value, j.value


class World:
    """The handle that plays the role of the reference's MPI communicator."""

    def __init__(self, rank: int | None = None, size: int | None = None, device: int | None = None):
        import torch
        import torch.distributed as dist

        self.rank = int(os.environ.get("RANK", "0")) if rank is None else rank
        self.size = int(os.environ.get("WORLD_SIZE", "1")) if size is None else size
        self.device = int(os.environ.get("LOCAL_RANK", "0")) if device is None else device
        torch.cuda.set_device(self.device)
        self._own_pg = False
        if self.size > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("cpu:gloo,cuda:nccl", rank=self.rank, world_size=self.size)
            self._own_pg = True
        L = lib()
        ident = [None]
        if self.rank == 0:
            buf = ctypes.create_string_buffer(ID_BYTES)
            if L.chase_b200_comm_unique_id(buf) != 0:
                raise RuntimeError("chase_b200: ncclGetUniqueId failed")
            ident = [buf.raw]
        if self.size > 1:
            dist.broadcast_object_list(ident, src=0)
        self.handle = ctypes.c_void_p()
        rc = L.chase_b200_comm_init(self.rank, self.size, ctypes.c_char_p(ident[0]), self.device,
                                    ctypes.byref(self.handle))
        if rc != 0:
            raise RuntimeError("chase_b200: communicator init failed")

    def barrier(self):
        import torch.distributed as dist

        if self.size > 1:
            dist.barrier()

    def max(self, x: float) -> float:
        import torch
        import torch.distributed as dist

        if self.size == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def close(self):
        if self.handle:
            lib().chase_b200_comm_free(self.handle)
            self.handle = ctypes.c_void_p()


class PChASE:
    """p?chase_init_[blockcyclic_] / p?chase_ / p?chase_finalize_ on this rank's local pieces (host numpy buffers)."""

    def __init__(self, world: World, N: int, nev: int, nex: int, H_loc: np.ndarray, grid=None, major="R", mb=0, nb=0,
                 V_loc: np.ndarray | None = None, pseudo: bool = False):
        self.world = world
        self.N, self.nev, self.nex, self.nevex = int(N), int(nev), int(nex), int(nev + nex)
        self.pseudo = bool(pseudo)  # p?chase_init_pseudo_[blockcyclic_]: 2 (nev+nex) columns
        self.ncols = (2 if self.pseudo else 1) * self.nevex
        self.grid = grid or grid_dims(world.size)
        self.major = major
        self.mb, self.nb = int(mb), int(nb)
        r, c = self.grid
        self.i, self.j = grid_coords(r, c, major, world.rank)
        self.m = local_size(N, r, mb, self.i)
        self.n = local_size(N, c, nb, self.j)
        if isinstance(H_loc, np.dtype) or isinstance(H_loc, type):
            # no host matrix: the block will be handed over on the device (load_device_matrix)
            self._dtype = np.dtype(H_loc)
            H_loc = None
        else:
            H_loc = np.asarray(H_loc)
            if H_loc.shape != (self.m, self.n):
                raise ValueError(f"local block must be {self.m} x {self.n}, got {H_loc.shape}")
            if not H_loc.flags.f_contiguous:
                H_loc = np.asfortranarray(H_loc)
            self._dtype = H_loc.dtype
        self.pfx = _PFX[self._dtype]
        self.rdt = _REAL[self.pfx]
        self.H = H_loc
        if self.pseudo and self.pfx not in ("c", "z"):
            raise ValueError("pseudo-Hermitian problems are complex")
        if V_loc is None:
            V_loc = np.zeros((max(self.m, 1), self.ncols), dtype=self._dtype, order="F")
        assert V_loc.flags.f_contiguous and V_loc.shape == (max(self.m, 1), self.ncols)
        self.V = V_loc
        self.ritzv = np.zeros(self.ncols, dtype=self.rdt)
        self._lib = lib()
        flag = ctypes.c_int(0)
        ldh = ctypes.byref(ctypes.c_int(max(self.m, 1)))
        gm = ctypes.c_char_p(major.encode())
        comm = ctypes.byref(world.handle)
        ps = "pseudo_" if self.pseudo else ""
        if mb == 0 and nb == 0:
            getattr(self._lib, f"p{self.pfx}chase_init_{ps}")(
                _i(N), _i(nev), _i(nex), _i(self.m), _i(self.n), self._hp(), ldh, _p(self.V), _p(self.ritzv),
                _i(r), _i(c), gm, comm, ctypes.byref(flag))
        else:
            getattr(self._lib, f"p{self.pfx}chase_init_{ps}blockcyclic_")(
                _i(N), _i(nev), _i(nex), _i(mb), _i(nb), self._hp(), ldh, _p(self.V), _p(self.ritzv), _i(r), _i(c),
                gm, _i(0), _i(0), comm, ctypes.byref(flag))
        if flag.value != 1:
            raise RuntimeError("chase_b200: p?chase_init_ failed")
        self._alive = True

    def _hp(self):
        return _p(self.H) if self.H is not None else ctypes.c_void_p(0)

    def load_device_matrix(self, dev_ptr: int, ld: int):
        """Hand over this rank's block as a column-major device array (e.g. a torch tensor's data_ptr()).

        The library copies on its own (non-blocking) stream, which does not wait for torch's streams: the producer
        of the block is synchronised here first."""
        import torch

        torch.cuda.synchronize()
        rc = self._lib.chase_b200_dist_load_device_matrix_(ctypes.c_char_p(self.pfx.encode()), ctypes.c_void_p(dev_ptr),
                                                           ctypes.byref(ctypes.c_longlong(ld)))
        if rc != 0:
            raise RuntimeError("chase_b200: device matrix hand-over failed")

    def device_matrix(self):
        """(device pointer, leading dimension) of the solver's own column-major local block, for callers that generate
        the block in place (it then exists only once in HBM); call mark_device_matrix() when it is filled."""
        ptr, ld = ctypes.c_void_p(), ctypes.c_longlong()
        rc = self._lib.chase_b200_dist_device_matrix_(ctypes.c_char_p(self.pfx.encode()), ctypes.byref(ptr), ctypes.byref(ld))
        if rc != 0:
            raise RuntimeError("chase_b200: no active distributed solver")
        return ptr.value, ld.value

    def mark_device_matrix(self):
        import torch

        torch.cuda.synchronize()
        if self._lib.chase_b200_dist_mark_device_matrix_(ctypes.c_char_p(self.pfx.encode())) != 0:
            raise RuntimeError("chase_b200: no active distributed solver")

    def row_indices(self):
        return global_indices(self.N, self.grid[0], self.mb, self.i)

    def col_indices(self):
        return global_indices(self.N, self.grid[1], self.nb, self.j)

    def solve(self, deg=20, tol=None, mode="R", opt="S", qr="C", trace=False, copy=True) -> SolveResult:
        L = self._lib
        if tol is None:
            tol = 1e-10 if self.rdt == np.float64 else 1e-5
        L.chase_b200_trace_enable_(_i(1 if trace else 0))
        tolc = ctypes.c_double(tol) if self.rdt == np.float64 else ctypes.c_float(tol)
        getattr(L, f"p{self.pfx}chase_")(
            _i(deg), ctypes.byref(tolc), ctypes.c_char_p(mode.encode()), ctypes.c_char_p(opt.encode()),
            ctypes.c_char_p(qr.encode()))
        resid = np.zeros(self.nevex, dtype=self.rdt)
        getattr(L, f"p{self.pfx}chase_get_resid_")(_p(resid))
        st = np.zeros(16)
        L.chase_b200_get_stats_(_p(st), _i(16))
        if st[15] != 0:
            msg = "chase_b200: distributed solve failed"
            if trace:  # the call trace up to the failure is the best diagnostic there is
                n = L.chase_b200_trace_copy_(None, 0)
                buf = ctypes.create_string_buffer(n + 1)
                L.chase_b200_trace_copy_(buf, n + 1)
                tail = [t[:160] for t in buf.value.decode().splitlines()][-25:]
                msg += "; last calls:\n  " + "\n  ".join(tail)
            raise RuntimeError(msg)
        res = SolveResult(self.ritzv.copy(), resid, self.V.copy(order="F") if copy else self.V,
                          dict(zip(STAT_NAMES, st.tolist())))
        if trace:
            n = L.chase_b200_trace_copy_(None, 0)
            buf = ctypes.create_string_buffer(n + 1)
            L.chase_b200_trace_copy_(buf, n + 1)
            res.trace = buf.value.decode().splitlines()
        n = L.chase_b200_qr_log_copy_(None, 0)
        buf = ctypes.create_string_buffer(n + 1)
        L.chase_b200_qr_log_copy_(buf, n + 1)
        res.qr_log = buf.value.decode().splitlines()
        return res

    def finalize(self):
        if self._alive:
            flag = ctypes.c_int(1)
            getattr(self._lib, f"p{self.pfx}chase_finalize_")(ctypes.byref(flag))
        self._alive = False

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.finalize()
