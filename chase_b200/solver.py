"""Host-side mirror of the reference's sequential C interface.

Every call goes through the C ABI of ``include/chase_c_interface.h`` with HOST
(numpy) buffers, exactly as a C or Fortran application would call the
reference (``/root/reference/interface/chase_c_interface.h:17-41``):

    s = ChASE(H, nev, nex)            # ?chase_init_   (H: column-major host matrix)
    r = s.solve(deg=20, tol=1e-10)    # ?chase_        ('R'/'A', 'S'/'N', 'C'/'H')
    s.finalize()                      # ?chase_finalize_

Pseudo-Hermitian (BSE) problems use the ``?chase_init_pseudo_`` / ``?chase_pseudo_`` pair of the same interface
(``chase_c_interface.h:42-58``): ``ChASE(H, nev, nex, pseudo=True)``; V then has ``2 (nev+nex)`` columns.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field

import numpy as np

from ._lib import lib

_PFX = {np.dtype(np.float32): "s", np.dtype(np.float64): "d", np.dtype(np.complex64): "c", np.dtype(np.complex128): "z"}
_REAL = {"s": np.float32, "d": np.float64, "c": np.float32, "z": np.float64}
STAT_NAMES = [
    "iterations", "filtered_vecs", "hemm_calls", "swaps", "t_all", "t_initvecs", "t_lanczos", "t_filter", "t_qr",
    "t_rr", "t_resid", "gflop_filter", "gflop_total", "heev_sweeps", "gather_passes", "error",
]


@dataclass
class SolveResult:
    ritzv: np.ndarray  # nev+nex Ritz values (first nev ascending)
    resid: np.ndarray  # nev+nex residual norms
    V: np.ndarray  # N x (nev+nex), column-major
    stats: dict = field(default_factory=dict)
    trace: list = field(default_factory=list)
    qr_log: list = field(default_factory=list)

    @property
    def iterations(self):
        return int(self.stats["iterations"])

    @property
    def filtered_vecs(self):
        return int(self.stats["filtered_vecs"])


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _i(v):
    return ctypes.byref(ctypes.c_int(int(v)))


class ChASE:
    """One process-global solver per scalar type (like the reference: not re-entrant)."""

    _active: dict = {}

    def __init__(self, H: np.ndarray, nev: int, nex: int, V: np.ndarray | None = None, pseudo: bool = False):
        H = np.asarray(H)
        if H.ndim != 2 or H.shape[0] != H.shape[1]:
            raise ValueError("H must be square")
        if not H.flags.f_contiguous:
            H = np.asfortranarray(H)
        self.pfx = _PFX[H.dtype]
        self.rdt = _REAL[self.pfx]
        self.H = H
        self.N, self.nev, self.nex = H.shape[0], int(nev), int(nex)
        self.nevex = self.nev + self.nex
        self.pseudo = bool(pseudo)
        if self.pseudo and self.pfx not in ("c", "z"):
            raise ValueError("pseudo-Hermitian problems are complex (cchase_init_pseudo_ / zchase_init_pseudo_)")
        self.ncols = (2 if self.pseudo else 1) * self.nevex
        if V is None:
            V = np.zeros((self.N, self.ncols), dtype=H.dtype, order="F")
        else:
            V = np.asfortranarray(V, dtype=H.dtype)
            assert V.shape == (self.N, self.ncols)
        self.V = V
        self.ritzv = np.zeros(self.ncols, dtype=self.rdt)
        self._lib = lib()
        flag = ctypes.c_int(0)
        getattr(self._lib, f"{self.pfx}chase_init_pseudo_" if self.pseudo else f"{self.pfx}chase_init_")(
            _i(self.N), _i(self.nev), _i(self.nex), _p(self.H), _i(self.N), _p(self.V), _p(self.ritzv),
            ctypes.byref(flag),
        )
        if flag.value != 1:
            raise RuntimeError("chase_b200: ?chase_init_ failed: " + self._last_error())
        self._alive = True
        ChASE._active[self.pfx] = id(self)

    # unified setters of the reference interface (chase_c_interface.h:217-239)
    def set(self, **kw):
        L = self._lib
        for k, v in kw.items():
            if k in ("decaying_rate", "upperb_scale_rate"):
                getattr(L, f"chase_set_{k}_")(ctypes.byref(ctypes.c_float(v)))
            elif k == "tol":
                L.chase_set_tol_(ctypes.byref(ctypes.c_double(v)))
            elif k == "sym_check":
                L.chase_enable_sym_check_(_i(v))
            else:
                getattr(L, f"chase_set_{k}_")(_i(v))
        return self

    def solve(self, deg: int = 20, tol: float | None = None, mode: str = "R", opt: str = "S", qr: str = "C",
              trace: bool = False, copy: bool = True) -> SolveResult:
        L = self._lib
        if tol is None:
            tol = 1e-10 if self.rdt == np.float64 else 1e-5
        L.chase_b200_trace_enable_(_i(1 if trace else 0))
        tolc = ctypes.c_double(tol) if self.rdt == np.float64 else ctypes.c_float(tol)
        getattr(L, f"{self.pfx}chase_pseudo_" if self.pseudo else f"{self.pfx}chase_")(
            _i(deg), ctypes.byref(tolc), ctypes.c_char_p(mode.encode()), ctypes.c_char_p(opt.encode()),
            ctypes.c_char_p(qr.encode()),
        )
        resid = np.zeros(self.nevex, dtype=self.rdt)
        getattr(L, f"{self.pfx}chase_get_resid_")(_p(resid))
        st = np.zeros(16)
        L.chase_b200_get_stats_(_p(st), _i(16))
        if st[15] != 0:
            raise RuntimeError("chase_b200: solve failed: " + self._last_error())
        # copy=False hands out the caller-owned V itself (what a C caller sees), avoiding an N x (nev+nex) host copy
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

    def _last_error(self):
        f = self._lib.chase_b200_last_error_copy_
        f.restype = ctypes.c_size_t
        f.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        n = f(None, 0)
        buf = ctypes.create_string_buffer(n + 1)
        f(buf, n + 1)
        return buf.value.decode()

    def finalize(self):
        # the native solver is a per-type singleton: only its current owner may destroy it
        if self._alive and ChASE._active.get(self.pfx) == id(self):
            flag = ctypes.c_int(1)
            getattr(self._lib, f"{self.pfx}chase_finalize_")(ctypes.byref(flag))
            ChASE._active.pop(self.pfx, None)
        self._alive = False

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.finalize()

    def __del__(self):
        try:
            self.finalize()
        except Exception:
            pass
