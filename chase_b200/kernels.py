"""Kernel-level C ABI (include/chase_b200_kernels.h) on torch CUDA tensors.

Matrices are column-major: a torch tensor of shape (cols, ld) (row-major) is the
column-major (ld x cols) matrix the kernels expect.  ``colmajor``/``to_numpy``
convert from/to numpy.  PyTorch is only the device-memory allocator here.
"""
from __future__ import annotations

import ctypes

import numpy as np

from ._lib import lib

SFX = {"float32": "s", "float64": "d", "complex64": "c", "complex128": "z"}


def _sfx(t):
    return SFX[str(t.dtype).replace("torch.", "")]


def colmajor(a: np.ndarray, ld: int | None = None, device="cuda"):
    """numpy (rows x cols) -> torch tensor (cols, ld) holding the column-major matrix."""
    import torch

    rows, cols = a.shape
    ld = ld or rows
    buf = np.zeros((cols, ld), dtype=a.dtype)
    buf[:, :rows] = a.T
    return torch.from_numpy(buf).to(device)


def to_numpy(t, rows: int) -> np.ndarray:
    return t.cpu().numpy()[:, :rows].T.copy()


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(rc, what):
    if rc < 0:
        raise RuntimeError(f"chase_b200 kernel launcher {what} failed with {rc}")
    return rc


def _c(z):
    z = complex(z)
    return ctypes.c_double(z.real), ctypes.c_double(z.imag)


def gemm(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, uplo=0, ws=None):
    f = getattr(lib(), f"chase_b200_gemm_{_sfx(C)}")
    ar, ai = _c(alpha)
    br, bi = _c(beta)
    wsb = ws.numel() * ws.element_size() if ws is not None else 0
    return _chk(f(int(ta), int(tb), ctypes.c_int64(M), ctypes.c_int64(N), ctypes.c_int64(K), ar, ai, _ptr(A),
                  ctypes.c_int64(lda), _ptr(B), ctypes.c_int64(ldb), br, bi, _ptr(C), ctypes.c_int64(ldc), int(uplo),
                  _ptr(ws), ctypes.c_size_t(wsb), _stream()), "gemm")


def hemm(n, k, alpha, A, lda, B, ldb, beta, C, ldc, shift=0.0, theta=None):
    f = getattr(lib(), f"chase_b200_hemm_{_sfx(C)}")
    ar, ai = _c(alpha)
    br, bi = _c(beta)
    return _chk(f(ctypes.c_int64(n), ctypes.c_int64(k), ar, ai, _ptr(A), ctypes.c_int64(lda), _ptr(B),
                  ctypes.c_int64(ldb), br, bi, _ptr(C), ctypes.c_int64(ldc), ctypes.c_double(shift), _ptr(theta),
                  _stream()), "hemm")


def hemm_rect(ta, M, K, k, alpha, A, lda, B, ldb, beta, C, ldc):
    """C(M x k) <- alpha op(A) B(K x k) + beta C; op(A) = A (M x K) or A^H (A stored K x M) for ta."""
    f = getattr(lib(), f"chase_b200_hemm_rect_{_sfx(C)}")
    ar, ai = _c(alpha)
    br, bi = _c(beta)
    return _chk(f(int(ta), ctypes.c_int64(M), ctypes.c_int64(K), ctypes.c_int64(k), ar, ai, _ptr(A),
                  ctypes.c_int64(lda), _ptr(B), ctypes.c_int64(ldb), br, bi, _ptr(C), ctypes.c_int64(ldc),
                  _stream()), "hemm_rect")


def tf32_lo(A, cplx: bool):
    """lo part of the TF32 split of the FP32 (or interleaved complex FP32) device array A (same shape)."""
    import torch

    L = lib()
    lo = torch.empty_like(A)
    ld = A.shape[-1]
    cols = A.numel() // ld
    _chk(L.chase_b200_tf32_register(_ptr(A), _ptr(lo), ctypes.c_int64(ld), ctypes.c_int64(ld), ctypes.c_int64(cols), 2,
                                    ctypes.c_void_p(0), ctypes.c_size_t(0)), "tf32_register")
    _chk(L.chase_b200_tf32_sync(ctypes.c_char(b"c" if cplx else b"s"), _ptr(A), _stream()), "tf32_sync")
    L.chase_b200_tf32_unregister(_ptr(A))
    return lo


def hemm_tf32(M, K, k, alpha, A, Alo, lda, B, ldb, beta, C, ldc, shift=0.0, theta=None, sflip=0, terms=3):
    """C(M x k) <- alpha S (A^s)^H S B + beta C - alpha shift_j B on the tcgen05 kind::tf32 kernel; A^s is K x M."""
    import torch

    L = lib()
    sfx = _sfx(C)
    es = C.element_size()
    nbytes = L.chase_b200_hemm_tf32_scratch_bytes(K, k, es)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=C.device)
    f = getattr(L, f"chase_b200_hemm_tf32_{sfx}")
    ar, ai = _c(alpha)
    br, bi = _c(beta)
    rc = f(ctypes.c_int64(M), ctypes.c_int64(K), ctypes.c_int64(k), ar, ai, _ptr(A), _ptr(Alo), ctypes.c_int64(lda),
           _ptr(B), ctypes.c_int64(ldb), br, bi, _ptr(C), ctypes.c_int64(ldc), ctypes.c_double(shift), _ptr(theta),
           ctypes.c_int64(sflip), int(terms), _ptr(scratch), ctypes.c_size_t(nbytes), _stream())
    torch.cuda.current_stream().synchronize()  # scratch is a temporary
    return _chk(rc, "hemm_tf32")


def tri_pack(n, G, ldg, P, lower):
    f = getattr(lib(), f"chase_b200_tri_pack_{_sfx(G)}")
    return _chk(f(ctypes.c_int64(n), _ptr(G), ctypes.c_int64(ldg), _ptr(P), int(lower), _stream()), "tri_pack")


def tri_unpack(n, P, G, ldg, lower):
    f = getattr(lib(), f"chase_b200_tri_unpack_{_sfx(G)}")
    return _chk(f(ctypes.c_int64(n), _ptr(P), _ptr(G), ctypes.c_int64(ldg), int(lower), _stream()), "tri_unpack")


def potrf(n, G, ldg, info):
    f = getattr(lib(), f"chase_b200_potrf_{_sfx(G)}")
    return _chk(f(ctypes.c_int64(n), _ptr(G), ctypes.c_int64(ldg), _ptr(info), _stream()), "potrf")


def trsm(rows, n, R, ldr, V, ldv, X, ldx):
    import torch

    f = getattr(lib(), f"chase_b200_trsm_{_sfx(V)}")
    nbytes = lib().chase_b200_trsm_ws_bytes(n, V.element_size())
    ws = torch.zeros(nbytes, dtype=torch.uint8, device=V.device)
    return _chk(f(ctypes.c_int64(rows), ctypes.c_int64(n), _ptr(R), ctypes.c_int64(ldr), _ptr(V), ctypes.c_int64(ldv),
                  _ptr(X), ctypes.c_int64(ldx), _ptr(ws), ctypes.c_size_t(nbytes), _stream()), "trsm")


def shift_abstrace(n, G, ldg, scale, shift_out=None):
    f = getattr(lib(), f"chase_b200_shift_abstrace_{_sfx(G)}")
    return _chk(f(ctypes.c_int64(n), _ptr(G), ctypes.c_int64(ldg), ctypes.c_double(scale), _ptr(shift_out), _stream()),
                "shift_abstrace")


def heev(n, G, ldg, Z, ldz):
    """-> (w ascending as numpy float64, sweeps, rc); eigenvectors in Z."""
    import torch

    f = getattr(lib(), f"chase_b200_heev_{_sfx(G)}")
    nbytes = lib().chase_b200_heev_ws_bytes(n, 1 if G.is_complex() else 0)
    ws = torch.zeros(nbytes, dtype=torch.uint8, device=G.device)
    w = np.zeros(n, dtype=np.float64)
    sweeps = ctypes.c_int(0)
    rc = f(ctypes.c_int64(n), _ptr(G), ctypes.c_int64(ldg), _ptr(Z), ctypes.c_int64(ldz),
           w.ctypes.data_as(ctypes.c_void_p), _ptr(ws), ctypes.c_size_t(nbytes), ctypes.byref(sweeps), _stream())
    _chk(rc, "heev")
    return w, sweeps.value, rc


def colnorms(rows, cols, X, ldx, out, take_sqrt=True):
    f = getattr(lib(), f"chase_b200_colnorms_{_sfx(X)}")
    return _chk(f(ctypes.c_int64(rows), ctypes.c_int64(cols), _ptr(X), ctypes.c_int64(ldx), _ptr(out),
                  int(bool(take_sqrt)), _stream()), "colnorms")


def lacpy(rows, cols, src, lds, dst, ldd):
    f = getattr(lib(), f"chase_b200_lacpy_{_sfx(src)}")
    return _chk(f(ctypes.c_int64(rows), ctypes.c_int64(cols), _ptr(src), ctypes.c_int64(lds), _ptr(dst),
                  ctypes.c_int64(ldd), _stream()), "lacpy")


def gather_cols(rows, cnt, scols, dcols, src, lds, dst, ldd):
    f = getattr(lib(), f"chase_b200_gather_cols_{_sfx(src)}")
    return _chk(f(ctypes.c_int64(rows), int(cnt), _ptr(scols), _ptr(dcols), _ptr(src), ctypes.c_int64(lds), _ptr(dst),
                  ctypes.c_int64(ldd), _stream()), "gather_cols")


def gemv_conjt(rows, cols, A, lda, X, ldx, nv, Y, ldy):
    f = getattr(lib(), f"chase_b200_gemv_conjt_{_sfx(A)}")
    return _chk(f(ctypes.c_int64(rows), ctypes.c_int64(cols), _ptr(A), ctypes.c_int64(lda), _ptr(X),
                  ctypes.c_int64(ldx), int(nv), _ptr(Y), ctypes.c_int64(ldy), _stream()), "gemv_conjt")


def lanczos_step(rows, nv, k, M, v0, v1, v2, ld, d, e, rbeta):
    f = getattr(lib(), f"chase_b200_lanczos_step_{_sfx(v1)}")
    return _chk(f(ctypes.c_int64(rows), int(nv), int(k), int(M), _ptr(v0), _ptr(v1), _ptr(v2), ctypes.c_int64(ld),
                  _ptr(d), _ptr(e), _ptr(rbeta), _stream()), "lanczos_step")


def normalize_cols(rows, cols, X, ldx):
    f = getattr(lib(), f"chase_b200_normalize_cols_{_sfx(X)}")
    return _chk(f(ctypes.c_int64(rows), ctypes.c_int64(cols), _ptr(X), ctypes.c_int64(ldx), _stream()),
                "normalize_cols")


def rng_normal(rows, cols, X, ldx, seed):
    f = getattr(lib(), f"chase_b200_rng_normal_{_sfx(X)}")
    return _chk(f(ctypes.c_int64(rows), ctypes.c_int64(cols), _ptr(X), ctypes.c_int64(ldx), ctypes.c_uint64(seed),
                  _stream()), "rng_normal")


def herm_check(n, A, lda, tol, bad):
    f = getattr(lib(), f"chase_b200_herm_check_{_sfx(A)}")
    return _chk(f(ctypes.c_int64(n), _ptr(A), ctypes.c_int64(lda), ctypes.c_double(tol), _ptr(bad), _stream()),
                "herm_check")


def shift_diag(n, A, lda, c):
    f = getattr(lib(), f"chase_b200_shift_diag_{_sfx(A)}")
    return _chk(f(ctypes.c_int64(n), _ptr(A), ctypes.c_int64(lda), ctypes.c_double(c), _stream()), "shift_diag")


def herm_mirror(n, A, lda, from_upper):
    f = getattr(lib(), f"chase_b200_herm_mirror_{_sfx(A)}")
    return _chk(f(ctypes.c_int64(n), _ptr(A), ctypes.c_int64(lda), int(from_upper), _stream()), "herm_mirror")


def tridiag_eig(n, batch, d, e, ldde, w, Z):
    return _chk(lib().chase_b200_tridiag_eig(int(n), int(batch), _ptr(d), _ptr(e), int(ldde), _ptr(w), _ptr(Z),
                                             _stream()), "tridiag_eig")


def hemm_path(code, n, k, lda, ldb, ldc):
    return lib().chase_b200_hemm_path(int(code), ctypes.c_int64(n), ctypes.c_int64(k), ctypes.c_int64(lda),
                                      ctypes.c_int64(ldb), ctypes.c_int64(ldc))


def dmma_peak(iters=20000):
    return lib().chase_b200_dmma_peak(int(iters), _stream())


# ---- pseudo-Hermitian (BSE) helpers ----------------------------------------------------------------------------
def scale_rows(nrows, cols, X, ldx, a, row0=0):
    """X[row0:row0+nrows, :cols] *= a."""
    f = getattr(lib(), f"chase_b200_scale_rows_{_sfx(X)}")
    p = ctypes.c_void_p(X.data_ptr() + row0 * X.element_size())
    return _chk(f(ctypes.c_int64(nrows), ctypes.c_int64(cols), p, ctypes.c_int64(ldx), ctypes.c_double(a), _stream()),
                "scale_rows")


def kconj(rows, cols, src, lds, dst, ldd):
    f = getattr(lib(), f"chase_b200_kconj_{_sfx(src)}")
    return _chk(f(ctypes.c_int64(rows), ctypes.c_int64(cols), _ptr(src), ctypes.c_int64(lds), _ptr(dst),
                  ctypes.c_int64(ldd), _stream()), "kconj")


def lanczos_pseudo_norm(rows, nv, ke, M, v1, v2, ld, e, bnorm):
    f = getattr(lib(), f"chase_b200_lanczos_pseudo_norm_{_sfx(v1)}")
    return _chk(f(ctypes.c_int64(rows), int(nv), int(ke), int(M), _ptr(v1), _ptr(v2), ctypes.c_int64(ld), _ptr(e),
                  _ptr(bnorm), _stream()), "lanczos_pseudo_norm")


def lanczos_pseudo_step(rows, nv, k, M, v0, v1, v2, ld, d, bnorm):
    f = getattr(lib(), f"chase_b200_lanczos_pseudo_step_{_sfx(v1)}")
    return _chk(f(ctypes.c_int64(rows), int(nv), int(k), int(M), _ptr(v0), _ptr(v1), _ptr(v2), ctypes.c_int64(ld),
                  _ptr(d), _ptr(bnorm), _stream()), "lanczos_pseudo_step")


def hhqr(rows, n, A, lda, Q, ldq):
    """Householder QR: Q <- orthonormal factor of A (rows x n), A <- R and the reflectors."""
    import torch

    L = lib()
    L.chase_b200_hhqr_ws_bytes.restype = ctypes.c_size_t
    L.chase_b200_hhqr_ws_bytes.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
    nbytes = L.chase_b200_hhqr_ws_bytes(rows, n, A.element_size())
    ws = torch.zeros(nbytes, dtype=torch.uint8, device=A.device)
    f = getattr(L, f"chase_b200_hhqr_{_sfx(A)}")
    return _chk(f(ctypes.c_int64(rows), ctypes.c_int64(n), _ptr(A), ctypes.c_int64(lda), _ptr(Q), ctypes.c_int64(ldq),
                  _ptr(ws), ctypes.c_size_t(nbytes), _stream()), "hhqr")
