"""TEST INFRASTRUCTURE — CPU restatement of the reference algorithm (numpy/scipy).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` leg of
``bench.py`` may import this module; the product path never does.

It restates, in plain numpy, the Hermitian ChASE solve of the reference (and, at the end of the file, the
pseudo-Hermitian / BSE solve: algorithm.inc:1834-2220 with ChASECPU<T, PseudoHermitianMatrix<T>>):

* driver      /root/reference/algorithm/algorithm.inc:1376-1788 (``solve``),
              :942-1009 (``filter``), :136-193 (``calc_degrees``),
              :519-578 (``locking``), :1067-1214 (``lanczos`` + DoS)
* backend     /root/reference/Impl/chase_cpu/chase_cpu.hpp:291-851 (ChASECPU)
* kernels     /root/reference/linalg/internal/cpu/cholqr1.hpp:49-196,
              lanczos.hpp:45-209, rayleighRitz.hpp:60-120, residuals.hpp:45-80

Third-party arithmetic the reference delegates to BLAS/LAPACK (gemm, herk,
potrf, trsm, heevd, stemr; vendor unpinned by the reference, OpenBLAS 0.3.15 in
``oracle/_ref``) is done here with numpy/scipy (whatever BLAS they bundle).

Pinned: ``tests/test_oracle_vs_reference.py`` checks this restatement against
the golden call traces of the unmodified reference CPU solver
(``tests/golden/*.json``, produced by ``oracle/_ref/chase_ref_cpu_*`` through
``tests/golden/make_golden.py``): identical iteration count, filtered-vector
count, HEMM schedule, QR variants and lock counts; eigenvalues to 1e-10.
"""
from __future__ import annotations

import ctypes
import math
import os
from dataclasses import dataclass, field

import numpy as np
import scipy.linalg as sla

_HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------
# libstdc++ mt19937 + normal_distribution stream (chase_cpu.hpp:296-309)
# --------------------------------------------------------------------------
def _mtlib():
    path = os.path.join(_HERE, "_build", "libmtnormal.so")
    if not os.path.exists(path):
        import subprocess

        subprocess.check_call(["make", "-C", _HERE, "helpers"], stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(path)
    lib.mt_normal_fill_skip.argtypes = [ctypes.c_uint, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]
    return lib


def mt_normal(seed: int, n: int, skip: int = 0) -> np.ndarray:
    out = np.empty(n, dtype=np.float64)
    _mtlib().mt_normal_fill_skip(seed, skip, n, out.ctypes.data)
    return out


def init_vectors(N: int, ncols: int, dtype, seed: int = 1337) -> np.ndarray:
    """Column-major N x ncols start block exactly as ChASECPU::initVecs fills it."""
    dtype = np.dtype(dtype)
    if dtype.kind == "c":
        s = mt_normal(seed, 2 * N * ncols)
        # getRandomT<complex>(f) = complex(f(), f()); g++ evaluates the two
        # calls right-to-left, so the FIRST draw lands in the imaginary part
        # (verified against chase_ref_cpu_z --initvecs-only in the tests).
        v = (s[1::2] + 1j * s[0::2]).astype(dtype)
    else:
        v = mt_normal(seed, N * ncols).astype(dtype)
    return np.asfortranarray(v.reshape(ncols, N).T)


# --------------------------------------------------------------------------
# matrices (tests/noinput.cpp:67-74; examples/2_input_output.cpp:250-262)
# --------------------------------------------------------------------------
def clement(N: int, dtype=np.float64) -> np.ndarray:
    H = np.zeros((N, N), dtype=dtype, order="F")
    i = np.arange(N - 1)
    v = np.sqrt((i * (N + 1 - i)).astype(np.float64))
    H[i + 1, i] = v
    H[i, i + 1] = v
    return H


def uniform_spectrum(N: int, dmax: float = 100.0, eps: float = 1e-4) -> np.ndarray:
    k = np.arange(N, dtype=np.float64)
    return dmax * (eps + k * (1.0 - eps) / float(N))


def uniform_diag(N: int, dtype=np.float64) -> np.ndarray:
    H = np.zeros((N, N), dtype=dtype, order="F")
    H[np.arange(N), np.arange(N)] = uniform_spectrum(N)
    return H


def dense_from_spectrum(lam: np.ndarray, dtype=np.float64, seed: int = 7, nrefl: int = 3) -> np.ndarray:
    """A = Q diag(lam) Q^H with Q a product of `nrefl` Householder reflectors."""
    N = lam.shape[0]
    dtype = np.dtype(dtype)
    rng = np.random.default_rng(seed)
    A = np.zeros((N, N), dtype=dtype, order="F")
    A[np.arange(N), np.arange(N)] = lam
    for _ in range(nrefl):
        v = rng.standard_normal(N)
        if dtype.kind == "c":
            v = v + 1j * rng.standard_normal(N)
        v = (v / np.linalg.norm(v)).astype(dtype)
        w = A @ v
        s = np.vdot(v, w)
        A -= 2 * np.outer(v, w.conj())
        A -= 2 * np.outer(w, v.conj())
        A += 4 * s * np.outer(v, v.conj())
    A = 0.5 * (A + A.conj().T)
    return np.asfortranarray(A)


def perturb_hermitian(H: np.ndarray, stream: np.ndarray, perturb: float) -> int:
    """In-place element-wise perturbation of oracle/ref_driver.cpp (mirrors
    tests/noinput.cpp:120-134); returns the number of stream values consumed."""
    N = H.shape[0]
    cplx = np.iscomplexobj(H)
    pos = 0
    for i in range(1, N):
        cnt = i - 1
        if cnt <= 0:
            continue
        if cplx:
            blk = stream[pos:pos + 2 * cnt]
            e = (blk[0::2] + 1j * blk[1::2]) * perturb
            pos += 2 * cnt
        else:
            e = stream[pos:pos + cnt] * perturb
            pos += cnt
        H[1:i, i] += e
        H[i, 1:i] += np.conj(e)
    return pos


# --------------------------------------------------------------------------
# kernels (linalg/internal/cpu)
# --------------------------------------------------------------------------
def _eps(dtype):
    return np.finfo(np.dtype(dtype).char.lower() if np.dtype(dtype).kind == "c" else dtype).eps


def _real_dtype(dtype):
    return np.zeros(1, dtype=dtype).real.dtype


def _chol_step(V: np.ndarray, shift: float = 0.0) -> int:
    """one round: G = V^H V (+shift I), R = chol(G) upper, V <- V R^-1. cholqr1.hpp:49-80"""
    G = V.conj().T @ V
    if shift:
        G[np.diag_indices_from(G)] += shift
    try:
        Rm = sla.cholesky(G, lower=False, check_finite=False)
    except sla.LinAlgError:
        return 1
    V[...] = sla.solve_triangular(Rm, V.conj().T, trans="C", lower=False, check_finite=False).conj().T
    return 0


def cholqr1(V):
    return _chol_step(V)


def cholqr2(V):
    info = _chol_step(V)
    if info:
        return info
    return _chol_step(V)


def shifted_cholqr2(V):
    """cholqr1.hpp:136-196: shift = sqrt(m) * sum|G_ii| * eps (double), 10 * sum|G_ii| * eps (float)."""
    m = V.shape[0]
    G = V.conj().T @ V
    nrmf = float(np.sum(np.abs(np.diag(G))))
    rd = _real_dtype(V.dtype)
    if rd == np.float32:
        shift = 10.0 * nrmf * float(np.finfo(np.float32).eps)
    else:
        shift = math.sqrt(float(m)) * nrmf * float(np.finfo(np.float64).eps)
    info = _chol_step(V, shift)
    if info:
        return info
    _chol_step(V)
    return _chol_step(V)


def householder_qr(V):
    Q, _ = np.linalg.qr(V)
    V[...] = Q


@dataclass
class Trace:
    calls: list = field(default_factory=list)
    swaps: int = 0
    hemm_calls: int = 0
    filtered_vecs: int = 0
    iterations: int = 0

    def add(self, s):
        self.calls.append(s)


class OracleBackend:
    """numpy mirror of ChASECPU (chase_cpu.hpp), Hermitian case only."""

    def __init__(self, H: np.ndarray, nev: int, nex: int, V0: np.ndarray | None = None):
        self.H = H  # never modified: Shift is tracked in self.shift and applied as A - cI
        self.N = H.shape[0]
        self.nev, self.nex, self.nevex = nev, nex, nev + nex
        self.dtype = H.dtype
        self.rdtype = _real_dtype(H.dtype)
        self.V1 = np.zeros((self.N, self.nevex), dtype=self.dtype, order="F") if V0 is None else np.asfortranarray(V0.copy())
        self.V2 = np.zeros_like(self.V1)
        self.ritzv = np.zeros(self.nevex, dtype=self.rdtype)
        self.resid = np.zeros(self.nevex, dtype=self.rdtype)
        self.locked = 0
        self.trace = Trace()
        self.qr_variants = []

    # -- ChaseBase virtuals --------------------------------------------------
    def Start(self):
        self.locked = 0

    def initVecs(self, random: bool):
        if random:
            self.V1 = init_vectors(self.N, self.nevex, self.dtype)
        self.V2[...] = self.V1

    def Shift(self, c, isunshift=False):
        # chase_cpu.hpp:392-397 adds c to the diagonal of H itself.
        self.H[np.arange(self.N), np.arange(self.N)] += self.dtype.type(c)

    def HEMM(self, block, alpha, beta, offset_left, offset_right=0):
        ncols = block - offset_right if offset_right < block else 0
        self.trace.hemm_calls += 1
        self.trace.filtered_vecs += ncols
        if ncols:
            s = slice(offset_left + self.locked, offset_left + self.locked + ncols)
            a = self.dtype.type(alpha)
            b = self.dtype.type(beta)
            self.V2[:, s] = a * (self.H @ self.V1[:, s]) + b * self.V2[:, s]
        self.V1, self.V2 = self.V2, self.V1

    def QR(self, fixednev, cond):
        # chase_cpu.hpp:597-781
        self.V2[:, : self.locked] = self.V1[:, : self.locked]
        dbl = self.rdtype == np.float64
        upper = 1e8 if dbl else 1e4
        lower = 2e1 if dbl else 1e1
        Vw = self.V1
        work = Vw
        if not dbl:
            # QR_DOUBLE_PRECISION only affects the Householder fallback on the CPU backend
            pass
        if getattr(self, "disable_cholqr", False) and cond != 1.0:
            # CHASE_DISABLE_CHOLQR=1 / qr == 'H': Householder QR in every iteration (chase_cpu.hpp:670-690)
            householder_qr(work)
            self.qr_variants.append("householder")
            self.V1[:, : self.locked] = self.V2[:, : self.locked]
            return
        if cond > upper:
            info = shifted_cholqr2(work)
            self.qr_variants.append("shifted2")
        elif cond < lower:
            info = cholqr1(work)
            self.qr_variants.append("chol1")
        else:
            info = cholqr2(work)
            self.qr_variants.append("chol2")
        if info != 0:
            householder_qr(work)
            self.qr_variants[-1] += "+householder"
        self.V1[:, : self.locked] = self.V2[:, : self.locked]

    def RR(self, block):
        # rayleighRitz.hpp:60-120 : W = A^H Q ; G = W^H Q ; heevd(lower) ; V2 = Q Z ; swap
        s = slice(self.locked, self.locked + block)
        Q = self.V1[:, s]
        W = self.H.conj().T @ Q
        G = W.conj().T @ Q
        w, Z = sla.eigh(G, lower=True, driver="evd", check_finite=False)
        self.ritzv[s] = w.astype(self.rdtype)
        self.V2[:, s] = Q @ Z.astype(self.dtype)
        self.V1, self.V2 = self.V2, self.V1

    def Resd(self):
        # residuals.hpp:45-80 on the nevex-locked trailing columns
        s = slice(self.locked, self.nevex)
        V = self.V1[:, s]
        W = self.H @ V - V * self.ritzv[s].astype(self.dtype)
        self.V2[:, s] = W
        self.resid[s] = np.linalg.norm(W, axis=0).astype(self.rdtype)

    def Swap(self, i, j):
        self.trace.swaps += 1
        self.V1[:, [i, j]] = self.V1[:, [j, i]]

    def Lock(self, n):
        self.locked += n

    def Lanczos(self, M, numvec):
        """lanczos.hpp:45-209 (multi-vector). Returns upperb, Theta, Tau, ritzV(last run)."""
        N = self.N
        dt = self.dtype
        v1 = np.array(self.V1[:, :numvec], order="F", copy=True)
        v0 = np.zeros_like(v1)
        d = np.zeros((M, numvec), dtype=self.rdtype)
        e = np.zeros((M, numvec), dtype=self.rdtype)
        v1 /= np.linalg.norm(v1, axis=0).astype(self.rdtype)
        r_beta = np.zeros(numvec, dtype=self.rdtype)
        AH = self.H.conj().T
        for k in range(M):
            self.V1[:, k] = v1[:, numvec - 1]
            v2 = AH @ v1
            alpha = np.einsum("ij,ij->j", v1.conj(), v2)
            v2 -= v1 * alpha
            d[k, :] = alpha.real
            if k > 0:
                v2 -= v0 * r_beta.astype(dt)
            r_beta = np.linalg.norm(v2, axis=0).astype(self.rdtype)
            if k == M - 1:
                break
            v2 *= (1.0 / r_beta).astype(dt)
            e[k, :] = r_beta
            v0, v1 = v1, v2
        self.V1[:, :numvec] = v1
        Theta = np.zeros((numvec, M), dtype=self.rdtype)
        Tau = np.zeros((numvec, M), dtype=self.rdtype)
        ritzV = None
        for i in range(numvec):
            w, Z = sla.eigh_tridiagonal(d[:, i].astype(np.float64), e[: M - 1, i].astype(np.float64), lapack_driver="stemr")
            Theta[i, :] = w
            Tau[i, :] = np.abs(Z[0, :]) ** 2
            ritzV = Z
        upperb = max(max(abs(Theta[i, 0]), abs(Theta[i, M - 1])) + abs(r_beta[i]) for i in range(numvec))
        return self.rdtype.type(upperb), Theta, Tau, ritzV

    def LanczosDos(self, idx, m, ritzV):
        # chase_cpu.hpp:376-390
        self.V2[:, :idx] = self.V1[:, :m] @ ritzV[:, :idx].astype(self.dtype)
        self.V1[:, :m] = self.V2[:, :m]


# --------------------------------------------------------------------------
# driver (algorithm.inc)
# --------------------------------------------------------------------------
@dataclass
class Config:
    tol: float = 1e-10
    deg: int = 20
    max_deg: int = 36
    deg_extra: int = 2
    max_iter: int = 25
    lanczos_iter: int = 25
    num_lanczos: int = 4
    opt: bool = True
    approx: bool = False
    decaying_rate: float = 1.0

    @staticmethod
    def for_dtype(dtype):
        if _real_dtype(dtype) == np.float32:
            return Config(tol=1e-5, deg=10, max_deg=18, lanczos_iter=12)
        return Config()


def _lanczos_dos(be: OracleBackend, cfg: Config, lanczos_iter: int, random: bool):
    """algorithm.inc:1067-1214; returns upperb and fills be.ritzv when random."""
    N, nevex = be.N, be.nevex
    numvec, m = cfg.num_lanczos, lanczos_iter
    rd = be.rdtype
    if not random:
        raise NotImplementedError  # handled by caller
    upperb, Theta, Tau, ritzV = be.Lanczos(m, numvec)
    be.trace.add(f"Lanczos {m} {numvec} {float(upperb)!r}")
    Th = Theta.reshape(-1).astype(np.float64)
    Ta = Tau.reshape(-1).astype(np.float64)
    ThS = np.sort(Th)
    lam = rd.type(ThS[0])
    sigma = 0.25
    threshold = 2 * sigma * sigma / 10
    search = float(nevex) / float(N)
    lowerb = rd.type(0)
    prev = 0.0
    n = numvec * m
    for i in range(n - 1):
        x = ThS[i]
        curr = 0.0
        for j in range(n):
            if x < Th[j] - threshold:
                pass
            elif x > Th[j] + threshold:
                curr += Ta[j]
            else:
                curr += Ta[j] * 0.5 * (1 + math.erf((x - Th[j]) / math.sqrt(2 * sigma * sigma)))
        curr /= numvec
        if curr > search:
            if abs(curr - search) < abs(prev - search):
                lowerb = rd.type(ThS[i + 1] if i + 1 < n else ThS[i])
            else:
                lowerb = rd.type(ThS[i])
            break
        prev = curr
    idx = 0
    last = Theta[numvec - 1]
    for i in range(m):
        if last[i] > lowerb:
            idx = i - 1
            break
    if idx > 0:
        be.trace.add(f"LanczosDos {idx} {m}")
        be.LanczosDos(idx, m, ritzV)
    rv = be.ritzv
    for i in range(max(idx, 0)):
        rv[i] = last[i]
    for i in range(max(idx, 0), nevex - 1):
        rv[i] = lam
    rv[nevex - 1] = lowerb
    for i in range(1, idx):
        j = i * (nevex // idx)
        be.Swap(i, j)
        rv[i], rv[j] = rv[j], rv[i]
    return upperb


def _lanczos_single(be: OracleBackend, m: int):
    """lanczos.hpp:231-330 single-vector variant (approx mode): upper bound only."""
    dt = be.dtype
    rd = be.rdtype
    v1 = be.V1[:, 0].copy()
    v1 /= rd.type(np.linalg.norm(v1))
    v0 = np.zeros_like(v1)
    d = np.zeros(m, dtype=rd)
    e = np.zeros(m, dtype=rd)
    AH = be.H.conj().T
    r_beta = rd.type(0)
    for k in range(m):
        v2 = AH @ v1
        alpha = np.vdot(v1, v2)
        v2 -= alpha * v1
        d[k] = alpha.real
        if k > 0:
            v2 -= dt.type(r_beta) * v0
        r_beta = rd.type(np.linalg.norm(v2))
        if k == m - 1:
            break
        v2 *= dt.type(1.0 / r_beta)
        e[k] = r_beta
        v0, v1 = v1, v2
    w = sla.eigh_tridiagonal(d.astype(np.float64), e[: m - 1].astype(np.float64), eigvals_only=True, lapack_driver="stemr")
    return rd.type(max(abs(w[0]), abs(w[-1])) + abs(r_beta))


def _calc_degrees(be, cfg, unconverged, nex, upperb, lowerb, tol, ritzv, resid, degrees, locked):
    rd = be.rdtype
    c = (upperb + lowerb) / rd.type(2)
    e = (upperb - lowerb) / rd.type(2)
    for i in range(unconverged - nex):
        t = (ritzv[i] - c) / e
        sq = np.sqrt(np.abs(t * t - 1))
        rho = max(abs(t - sq), abs(t + sq))
        dg = int(math.ceil(abs(math.log(float(resid[i]) / tol) / math.log(float(rho)))))
        if rd == np.float32:
            dg = max(dg, 8)
        degrees[i] = min(dg + cfg.deg_extra, cfg.max_deg)
    for i in range(unconverged - nex, unconverged):
        degrees[i] = degrees[unconverged - 1 - nex]
    for i in range(unconverged):
        degrees[i] += degrees[i] % 2
    for j in range(unconverged - 1):
        for k in range(j, unconverged):
            if degrees[k] < degrees[j]:
                degrees[k], degrees[j] = degrees[j], degrees[k]
                ritzv[k], ritzv[j] = ritzv[j], ritzv[k]
                resid[k], resid[j] = resid[j], resid[k]
                be.Swap(k + locked, j + locked)
    return int(degrees[unconverged - 1])


def _filter(be, unprocessed, deg, degrees, lambda_1, lower, upper):
    rd = be.rdtype
    c = (upper + lower) / rd.type(2)
    e = (upper - lower) / rd.type(2)
    sigma_1 = e / (lambda_1 - c)
    sigma = sigma_1
    be.trace.add(f"Shift {float(-c)!r} 0")
    be.Shift(-c)
    alpha = sigma_1 / e
    off = 0
    num_mult = 0
    dpos = 0
    be.trace.add(f"HEMM {unprocessed} {off}")
    be.HEMM(unprocessed, alpha, 0.0, off)
    num_mult += 1
    while unprocessed >= 0 and degrees[dpos] <= num_mult:
        dpos += 1
        unprocessed -= 1
        off += 1
    for _ in range(2, deg + 1):
        sigma_new = rd.type(1.0) / (rd.type(2.0) / sigma_1 - sigma)
        alpha = rd.type(2.0) * sigma_new / e
        beta = -sigma * sigma_new
        be.trace.add(f"HEMM {unprocessed} {off}")
        be.HEMM(unprocessed, alpha, beta, off)
        sigma = sigma_new
        num_mult += 1
        while unprocessed != 0 and degrees[dpos] <= num_mult:
            dpos += 1
            unprocessed -= 1
            off += 1
    be.Shift(+c, True)


def _locking(be, unconverged, tol, ritzv, resid, residLast, locked):
    index = sorted(range(unconverged), key=lambda a: ritzv[a])  # std::sort on distinct keys
    converged = 0
    early = 0
    for k in range(unconverged):
        j = index[k]
        if resid[j] <= tol or (resid[j] >= residLast[j] and resid[j] < 100.0 * tol):
            if resid[j] > tol:
                early += 1
            if j != converged:
                resid[j], resid[converged] = resid[converged], resid[j]
                residLast[j], residLast[converged] = residLast[converged], residLast[j]
                ritzv[j], ritzv[converged] = ritzv[converged], ritzv[j]
                be.Swap(j + locked, converged + locked)
            converged += 1
    return converged, early


def solve(be: OracleBackend, cfg: Config) -> Trace:
    """algorithm.inc:1376-1788 (Hermitian branch)."""
    rd = be.rdtype
    N, nev, nex, nevex = be.N, be.nev, be.nex, be.nevex
    tr = be.trace
    be.Start()
    unconverged = nevex
    fmax = np.finfo(rd).max
    residLast_ = np.full(nevex, fmax, dtype=rd)
    be.resid[:] = fmax
    deg = cfg.deg + cfg.deg % 2
    deg = min(deg, cfg.max_deg)
    degrees_ = [deg] * nevex
    random = not cfg.approx
    be.initVecs(random)
    if random:
        tr.add("QR 0 1")
        be.QR(0, 1.0)
    lanczos_iter = min(nevex, min(N // 2, cfg.lanczos_iter))
    if 2 * (lanczos_iter // 2) < lanczos_iter:
        lanczos_iter -= 1
        cfg.lanczos_iter = lanczos_iter
    if random:
        upperb = _lanczos_dos(be, cfg, lanczos_iter, True)
    else:
        upperb = _lanczos_single(be, lanczos_iter)
    locked = 0
    iteration = 0
    ritzv_all, resid_all = be.ritzv, be.resid
    lowerb = rd.type(np.max(ritzv_all[:unconverged])) * rd.type(cfg.decaying_rate)
    lam = rd.type(np.min(ritzv_all[:nevex]))
    tol = cfg.tol
    while unconverged > nex and iteration < cfg.max_iter:
        ritzv = ritzv_all[locked:]
        resid = resid_all[locked:]
        residLast = residLast_[locked:]
        degrees = degrees_[locked:]
        if np.all(resid[:unconverged] <= 0.5):
            lowerb = ritzv[unconverged - 1]
        if lowerb > upperb:
            lowerb = upperb
        residLast[:unconverged] = np.minimum(residLast[:unconverged], resid[:unconverged])
        if cfg.opt and iteration != 0:
            deg = _calc_degrees(be, cfg, unconverged, nex, upperb, lowerb, tol, ritzv, resid, degrees, locked)
            degrees_[locked:] = degrees
        tr.add(f"iter {iteration} lambda {float(lam)!r} lowerb {float(lowerb)!r} upperb {float(upperb)!r} unconv {unconverged}")
        _filter(be, unconverged, deg, degrees, lam, lowerb, upperb)
        cc = (upperb + lowerb) / rd.type(2)
        ee = (upperb - lowerb) / rd.type(2)
        t_1 = (ritzv_all[0] - cc) / ee
        t_k = (ritzv[0] - cc) / ee
        with np.errstate(invalid="ignore"):
            rho_1 = max(abs(t_1 - np.sqrt(t_1 * t_1 - 1)), abs(t_1 + np.sqrt(t_1 * t_1 - 1)))
            rho_k = max(abs(t_k - np.sqrt(t_k * t_k - 1)), abs(t_k + np.sqrt(t_k * t_k - 1)))
        dmax = max(degrees[: nevex - locked])
        cond = rd.type(float(rho_k) ** degrees[0] * float(rho_1) ** (dmax - degrees[0]))
        tr.add(f"QR {locked} {float(cond)!r}")
        be.QR(locked, cond)
        be.RR(unconverged)
        tr.add(f"RR {unconverged}")
        be.Resd()
        new_converged, _ = _locking(be, unconverged - nex, tol, ritzv, resid, residLast, locked)
        tr.add(f"Lock {new_converged}")
        be.Lock(new_converged)
        locked += new_converged
        unconverged -= new_converged
        iteration += 1
    # final sort of the first nev pairs (algorithm.inc:1726-1774)
    perm = sorted(range(nev), key=lambda i: ritzv_all[i])
    visited = [False] * nev
    for i in range(nev):
        if visited[i] or perm[i] == i:
            continue
        cyc = []
        cur = i
        while not visited[cur]:
            visited[cur] = True
            cyc.append(cur)
            cur = perm[cur]
        t_r, t_s = ritzv_all[i], resid_all[i]
        for k in range(len(cyc) - 1):
            ritzv_all[cyc[k]] = ritzv_all[cyc[k + 1]]
            resid_all[cyc[k]] = resid_all[cyc[k + 1]]
        ritzv_all[cyc[-1]] = t_r
        resid_all[cyc[-1]] = t_s
        for k in range(len(cyc) - 1):
            be.Swap(cyc[k], cyc[k + 1])
    tr.iterations = iteration
    return tr


def solve_problem(H: np.ndarray, nev: int, nex: int, cfg: Config | None = None, V0: np.ndarray | None = None):
    """Convenience wrapper: returns (ritzv, resid, V, trace, backend). H is not modified."""
    H = np.asfortranarray(H.copy())
    cfg = cfg or Config.for_dtype(H.dtype)
    be = OracleBackend(H, nev, nex, V0)
    tr = solve(be, cfg)
    return be.ritzv.copy(), be.resid.copy(), be.V1.copy(), tr, be


# Stand-alone single-kernel oracles used by the kernel-level parity tests ----
def gemm_filter_step(A, B, C, alpha, beta, shift):
    """C <- alpha*(A - shift*I) B + beta*C  (one filter step, algorithm.inc:985-990)."""
    return alpha * (A @ B - shift * B) + beta * C


def residual_norms(A, V, theta):
    """residuals.hpp:45-80."""
    return np.linalg.norm(A @ V - V * theta, axis=0)


# --------------------------------------------------------------------------
# Pseudo-Hermitian (BSE) path — kernel-level restatements.
#
# H = [[A, B], [-conj(B), -conj(A)]] with S = diag(I, -I) and S H Hermitian positive definite.  The functions below
# restate the BACKEND arithmetic the CUDA path has to reproduce and are pinned against the reference's golden spectra
# (tests/golden/bse_fixtures/eigs_*.bin) in tests/test_pseudo_cpu.py; the DRIVER (algorithm.inc:1834-2220 solve_pseudo
# and helpers) is restated further down (solve_pseudo) and pinned against golden traces of the unmodified reference
# (oracle/_ref/chase_ref_cpu_p{z,c}, tests/golden/pseudo_*.json).  Independently, the new C++ driver is pinned
# bit-for-bit against the reference's own driver by oracle/xcheck_driver.cpp.
# --------------------------------------------------------------------------
def bse_matrix(N: int, dtype=np.complex128, seed: int = 11, lam_min: float = 1.0, lam_max: float = 100.0,
               coupling: float = 0.3, nrefl: int = 3):
    """Synthetic pseudo-Hermitian matrix with an exactly known spectrum, computable block-wise.

    H = U [[a, b], [-conj(b), -a]] U^H with U = diag(Q, conj(Q)), Q a product of `nrefl` Householder reflectors,
    a, b diagonal, a_i > |b_i|: eigenvalues +-sqrt(a_i^2 - |b_i|^2).  Returns (H, positive eigenvalues ascending).
    """
    assert N % 2 == 0
    k = N // 2
    dtype = np.dtype(dtype)
    rng = np.random.default_rng(seed)
    lam = lam_min + (lam_max - lam_min) * (np.arange(k) / max(k - 1, 1))
    phase = np.exp(2j * np.pi * rng.random(k))
    b = coupling * lam * phase
    a = np.sqrt(lam**2 + np.abs(b) ** 2)
    Q = np.eye(k, dtype=np.complex128)
    for _ in range(nrefl):
        v = rng.standard_normal(k) + 1j * rng.standard_normal(k)
        v /= np.linalg.norm(v)
        Q -= 2 * np.outer(Q @ v, v.conj())
    A = (Q * a) @ Q.conj().T
    A = 0.5 * (A + A.conj().T)
    B = (Q * b) @ Q.T
    B = 0.5 * (B + B.T)
    H = np.empty((N, N), dtype=np.complex128, order="F")
    H[:k, :k] = A
    H[:k, k:] = B
    H[k:, :k] = -B.conj()
    H[k:, k:] = -A.conj()
    return np.asfortranarray(H.astype(dtype)), lam


def flip_lower_half(X: np.ndarray) -> np.ndarray:
    """S X: rows [N/2, N) negated (cpu/utils.hpp:100-111 flipLowerHalfMatrixSign)."""
    Y = X.copy()
    Y[X.shape[0] // 2:] *= -1
    return Y


def k_conjugate(X: np.ndarray) -> np.ndarray:
    """K-conjugate partner vectors: conj of the half-swapped block (chase_cpu.hpp:592-625 ApplyKconjugate)."""
    h = X.shape[0] // 2
    return np.conj(np.vstack([X[h:], X[:h]]))


def hemm_h2_step(H, V1, V2, alpha, beta, gamma):
    """V2 <- alpha H (H V1) + beta V2 + gamma V1 (chase_cpu.hpp:545-590 HEMM_H2)."""
    return alpha * (H @ (H @ V1)) + beta * V2 + gamma * V1


def rayleigh_ritz_v2(H: np.ndarray, Q: np.ndarray):
    """cpu/rayleighRitz.hpp:284-392: returns (ritz values [n], Ritz vectors of the first n/2 values [N x n/2]).

    A = Q^H S H Q = L L^H;  M = -L^-1 (I - 2 Q2^H Q2) L^-H;  heevd(M) ascending w;  ritz = 1 / (-w);
    X = L^-H Z, first n/2 columns normalised, V = Q X[:, :n/2].
    """
    N, n = Q.shape
    k = N // 2
    Aq = Q.conj().T @ flip_lower_half(H @ Q)
    L = np.linalg.cholesky(0.5 * (Aq + Aq.conj().T))
    M = np.eye(n, dtype=Q.dtype) - 2.0 * (Q[k:].conj().T @ Q[k:])
    M = sla.solve_triangular(L, M, lower=True)
    M = sla.solve_triangular(L, M.conj().T, lower=True).conj().T
    M = -M
    w, Z = np.linalg.eigh(0.5 * (M + M.conj().T))
    ritz = 1.0 / (-w)
    X = sla.solve_triangular(L.conj().T, Z, lower=False)
    X[:, : n // 2] /= np.linalg.norm(X[:, : n // 2], axis=0)
    return ritz, Q @ X[:, : n // 2]


def lanczos_pseudo(H: np.ndarray, V: np.ndarray, M: int, numvec: int):
    """cpu/lanczos.hpp:332-533 (multi-vector Lanczos in the S H inner product).

    V: N x >=max(M, numvec) start block (modified like the reference: column k <- k-th Lanczos vector of the last
    run, then the first numvec columns <- the final vectors).  Returns (Theta [numvec*M], Tau [numvec*M],
    ritzV [M x M of the last run], d, e).
    """
    N = H.shape[0]
    rdt = _real_dtype(H.dtype)
    d = np.zeros((M, numvec), dtype=rdt)
    e = np.zeros((M, numvec), dtype=rdt)
    v0 = np.zeros((N, numvec), dtype=H.dtype)
    v1 = V[:, :numvec].copy()
    v2 = H @ v1
    Sv = flip_lower_half(v2)
    beta = np.einsum("ij,ij->j", v1.conj(), Sv)
    beta = 1.0 / np.sqrt(beta)
    v1 = v1 * beta
    v2 = v2 * beta
    for k in range(M):
        V[:, k] = v1[:, numvec - 1]
        alpha = np.einsum("ij,ij->j", v2.conj(), Sv)
        alpha = -alpha * beta
        v2 = v2 + v1 * alpha
        alpha = -alpha
        d[k] = alpha.real
        if k == M - 1:
            break
        beta = -1.0 / beta
        v2 = v2 + v0 * beta
        beta = -beta
        v0, v1 = v1, v2
        v2 = H @ v1
        Sv = flip_lower_half(v2)
        beta = np.sqrt(np.einsum("ij,ij->j", v1.conj(), Sv))
        e[k] = beta.real
        beta = 1.0 / beta
        v1 = v1 * beta
        v2 = v2 * beta
    V[:, :numvec] = v1
    Theta = np.zeros(numvec * M, dtype=rdt)
    Tau = np.zeros(numvec * M, dtype=rdt)
    ritzV = None
    for i in range(numvec):
        w, Z = sla.eigh_tridiagonal(d[:, i].astype(np.float64), e[: M - 1, i].astype(np.float64))
        Theta[i * M:(i + 1) * M] = w
        Tau[i * M:(i + 1) * M] = np.abs(Z[0, :]) ** 2
        ritzV = Z
    return Theta, Tau, ritzV, d, e


def qr_pseudo(V: np.ndarray, locked: int, qr=None) -> np.ndarray:
    """chase_cpu.hpp:627-781 for the pseudo-Hermitian layout [locked+ | active (2u) | locked-]: the active columns
    are orthonormalised against S [locked+ locked-] and among themselves; locked columns are returned unchanged."""
    ncols = V.shape[1]
    W = np.hstack([flip_lower_half(V[:, :locked]), flip_lower_half(V[:, ncols - locked:]), V[:, locked:ncols - locked]])
    W = np.asfortranarray(W)
    (qr or cholqr2)(W)
    out = V.copy()
    out[:, locked:ncols - locked] = W[:, 2 * locked:]
    return out


# --------------------------------------------------------------------------
# Pseudo-Hermitian (BSE) path — backend mirror and driver restatement (complex128 / complex64 storage, decisions in
# the matching real type).  Pinned by tests/test_pseudo_cpu.py against the golden traces of the unmodified reference
# (tests/golden/pseudo_*.json): identical iteration count, filtered-vector count, HEMM_H2 schedule, lock counts,
# ApplyKconjugate sequence and swap count; eigenvalues to 1e-10.
# --------------------------------------------------------------------------
class OracleBackendPseudo:
    """numpy mirror of ChASECPU<T, PseudoHermitianMatrix<T>> (chase_cpu.hpp:74-92, 291-851): 2 (nev+nex) columns laid
    out [locked+ | active | K-conjugates of active | K-conjugates of locked+]."""

    def __init__(self, H: np.ndarray, nev: int, nex: int):
        self.H = H
        self.N = H.shape[0]
        self.nev, self.nex, self.nevex = nev, nex, nev + nex
        self.nc = 2 * self.nevex
        self.dtype = H.dtype
        self.rdtype = _real_dtype(H.dtype)
        self.V1 = np.zeros((self.N, self.nc), dtype=self.dtype, order="F")
        self.V2 = np.zeros_like(self.V1)
        self.ritzv = np.zeros(self.nc, dtype=self.rdtype)
        self.resid = np.zeros(self.nc, dtype=self.rdtype)
        self.locked = 0
        self.trace = Trace()
        self.qr_variants = []

    def Start(self):
        self.locked = 0

    def initVecs(self, random: bool):
        # chase_cpu.hpp:291-326: the CPU stream on all 2 (nev+nex) columns, lower block damped by T(0.001)
        if random:
            self.V1 = init_vectors(self.N, self.nc, self.dtype)
            self.V1[self.N // 2:] *= self.dtype.type(0.001)
        self.V2[...] = self.V1

    def HEMM_H2(self, block, alpha, beta, gamma, offset_left, offset_right=0):
        # chase_cpu.hpp:545-590; columns [locked + offset_left, + block - offset_right) as the reference multiplies
        # them (the trailing offset_left of those are K-conjugate slots rewritten by ApplyKconjugate afterwards)
        ncols = block - offset_right if offset_right < block else 0
        self.trace.hemm_calls += 1
        self.trace.filtered_vecs += 2 * ncols
        self.trace.add(f"HEMM_H2 {block} {offset_left}")
        if ncols:
            s = slice(offset_left + self.locked, offset_left + self.locked + ncols)
            t = self.dtype.type
            self.V2[:, s] = t(alpha) * (self.H @ (self.H @ self.V1[:, s])) + t(beta) * self.V2[:, s]
            self.V2[:, s] += t(gamma) * self.V1[:, s]
        self.V1, self.V2 = self.V2, self.V1

    def ApplyKconjugate(self, block):
        # chase_cpu.hpp:592-625
        self.trace.add(f"ApplyK {block}")
        c2 = self.nc - self.locked - block
        self.V1[:, c2:c2 + block] = k_conjugate(self.V1[:, self.locked:self.locked + block])

    def QR(self, fixednev, cond):
        # chase_cpu.hpp:627-781 (pseudo-Hermitian branches)
        self.trace.add(f"QR {fixednev} {cond!r}")
        L, nc = self.locked, self.nc
        W = np.asfortranarray(np.hstack([flip_lower_half(self.V1[:, :L]), flip_lower_half(self.V1[:, nc - L:]),
                                         self.V1[:, L:nc - L]]))
        dbl = self.rdtype == np.float64
        upper, lower = (1e8, 2e1) if dbl else (1e4, 1e1)
        if cond > upper:
            info = shifted_cholqr2(W)
            self.qr_variants.append("shifted2")
        elif cond < lower:
            info = cholqr1(W)
            self.qr_variants.append("chol1")
        else:
            info = cholqr2(W)
            self.qr_variants.append("chol2")
        if info != 0:
            householder_qr(W)
            self.qr_variants[-1] += "+householder"
        self.V1[:, L:nc - L] = W[:, 2 * L:]
        # both panels keep the locked columns (RR swaps the panels); with nothing locked the second panel also holds
        # the orthonormal block, which LanczosDos reads back (chase_cpu.hpp:757-774)
        self.V2[:, :L] = self.V1[:, :L]
        self.V2[:, nc - L:] = self.V1[:, nc - L:]
        if L == 0:
            self.V2[...] = self.V1

    def RR(self, block):
        # cpu/rayleighRitz.hpp:284-392 on the 2 block active columns; writes 2 block Ritz values, block vectors
        s = slice(self.locked, self.locked + 2 * block)
        ritz, X = rayleigh_ritz_v2(self.H, self.V1[:, s])
        self.ritzv[s] = ritz.astype(self.rdtype)
        self.V2[:, self.locked:self.locked + block] = X
        self.V1, self.V2 = self.V2, self.V1

    def Resd(self):
        # chase_cpu.hpp:803-817: first-half active columns only
        s = slice(self.locked, self.nevex)
        V = self.V1[:, s]
        W = self.H @ V - V * self.ritzv[s].astype(self.dtype)
        self.resid[s] = np.linalg.norm(W, axis=0).astype(self.rdtype)

    def Swap(self, i, j):
        self.trace.swaps += 1
        self.V1[:, [i, j]] = self.V1[:, [j, i]]

    def Lock(self, n):
        self.trace.add(f"Lock {n}")
        self.locked += n

    def Lanczos(self, M, numvec):
        Theta, Tau, ritzV, _, _ = lanczos_pseudo(self.H, self.V1, M, numvec)
        return Theta.astype(self.rdtype), Tau.astype(self.rdtype), ritzV

    def LanczosDos(self, idx, m, ritzV):
        self.trace.add(f"LanczosDos {idx} {m}")
        self.V2[:, :idx] = self.V1[:, :m] @ ritzV[:, :idx].astype(self.dtype)
        self.V1[:, :m] = self.V2[:, :m]


def _cheb_rho_c(t):
    q = np.sqrt(complex(t * t - 1, 0))
    return max(abs(complex(t, 0) - q), abs(complex(t, 0) + q))


def _cluster_factors(ritzv, resid, tol, unconverged, nex, upperb, lowerb):
    """algorithm.inc:19-135 detect_eigenvalue_clusters (double arithmetic)."""
    na = unconverged - nex
    thr = abs(upperb - lowerb) * 1e-6
    mean_res = float(np.sum(resid[:na])) / na
    weight = np.minimum(1.0 + np.log(1.0 + resid[:na] / (mean_res + 1e-14)), 2.5)
    f = np.ones(na)
    for i in range(na):
        dist = np.abs(ritzv[i] - ritzv[:na])
        near = (dist < thr)
        near[i] = False
        spatial = 1.0
        if near.any():
            density = 0.0
            for j in np.nonzero(near)[0]:
                density += weight[j] / (dist[j] + 1e-14)
            spatial = 1.0 + math.log(1.0 + density * 0.1)
        c = spatial * weight[i]
        if near.sum() > 2 and resid[i] > 2.0 * mean_res:
            c *= 1.2
        if resid[i] > 10.0 * tol:
            c *= 1.15
        f[i] = min(3.0, max(0.5, c))
    raw = f.copy()
    for i in range(1, na - 1):
        f[i] = 0.25 * raw[i - 1] + 0.5 * raw[i] + 0.25 * raw[i + 1]
    return np.minimum(3.0, np.maximum(0.5, f))


def _calc_degrees_pseudo_h2(be, cfg, unconverged, nex, upperb, lowerb, tol, ritzv, resid, residLast, degrees, locked):
    """algorithm.inc:196-317 (cluster-aware degrees on, the reference default); arrays are views from `locked` on."""
    factors = _cluster_factors(ritzv, resid, tol, unconverged, nex, upperb, lowerb)
    c, e = (upperb + lowerb) / 2, (upperb - lowerb) / 2
    cap = cfg.max_deg
    for i in range(unconverged):
        rho = _cheb_rho_c((ritzv[i] * ritzv[i] - c) / e)
        if not math.isfinite(rho) or rho <= 1:
            deg = cap
        else:
            steps = math.log(resid[i] / tol) / math.log(rho)
            if not math.isfinite(steps):
                deg = cap
            else:
                deg = int(math.ceil(abs(steps)))
                deg = int(deg * (factors[i] if i < len(factors) else 1.0))
                if resid[i] <= tol * 10.0:
                    if abs(resid[i] - residLast[i]) / (resid[i] + 1e-14) < 0.1:
                        deg += 6
                if abs(ritzv[i]) < abs(upperb - lowerb) * 0.1:
                    deg += 2
                deg = min(deg + cfg.deg_extra, cap)
        degrees[i] = deg + (deg % 2)
    for j in range(unconverged - 1):
        for k in range(j, unconverged):
            if degrees[k] < degrees[j]:
                degrees[[k, j]] = degrees[[j, k]]
                ritzv[[k, j]] = ritzv[[j, k]]
                resid[[k, j]] = resid[[j, k]]
                be.Swap(k + locked, j + locked)
    return int(degrees[:unconverged].max())


def _filter_h2(be, unconverged, degrees, lambda_1, lower, upper):
    """algorithm.inc:1012-1064."""
    if lower >= upper:
        lower, upper = upper, lower
    c, e = (upper + lower) / 2, (upper - lower) / 2
    sigma_1 = e / (lambda_1 - c)
    sigma = sigma_1
    deg_max = int(degrees[:unconverged].max())
    a1 = sigma_1 / e
    be.HEMM_H2(unconverged, a1, 0.0, -a1 * c, 0, 0)
    s = 0
    for t in range(2, deg_max + 1):
        if s >= unconverged:
            break
        tau = 1.0 / (2.0 / sigma_1 - sigma)
        alpha = 2.0 * tau / e
        be.HEMM_H2(unconverged, alpha, -(sigma * tau), -alpha * c, s, 0)
        sigma = tau
        while s < unconverged and degrees[s] <= t:
            s += 1


def _locking_pseudo(be, unconverged, nex, tol, ritzv, resid, residLast, locked, iteration, early):
    """algorithm.inc:730-816 locking_pseudo_v3 with the identity index the driver passes."""
    resid_in = resid[:2 * unconverged].copy()
    open_idx = []
    converged = 0
    for k in range(unconverged - nex):
        j = k
        early_lock = resid[j] > tol and resid[j] >= residLast[k] and resid[j] <= 1000.0 * tol and iteration >= 4
        if resid[j] <= tol or early_lock:
            if early_lock:
                early.append(float(resid[j]))
            if j != converged:
                resid[[j, converged]] = resid[[converged, j]]
                ritzv[[j, converged]] = ritzv[[converged, j]]
                be.Swap(j + locked, converged + locked)
            converged += 1
        else:
            open_idx.append(j)
    open_idx += list(range(unconverged - nex, unconverged))
    for i in range(converged, unconverged):
        residLast[i] = resid_in[open_idx[i - converged]]
    return converged


def _lanczos_for_h2(be, cfg, N, numvec, m, nevex, ritzv_):
    """algorithm.inc:1217-1373 (mode = true).  Returns b_sup = (max |theta|)^2."""
    Theta, Tau, ritzV = be.Lanczos(m, numvec)
    nt = numvec * m
    ThetaSorted = np.sort(Theta.astype(np.float64))
    sigma = 0.25
    thresh = 2 * sigma * sigma / 10
    absT = np.abs(Theta)
    i_min = int(np.argmin(absT))  # first minimum, like the reference's strict '<' scan
    mu_1 = Theta[i_min] * Theta[i_min]
    b_sup = absT.max() ** 2
    search = min(1.0, max(0.0, (N / 2 - cfg_nev(be) - be.nex - 1) / N))
    lam_q = ThetaSorted[nt - 1]
    prev = 0.0
    Td, Taud = Theta.astype(np.float64), Tau.astype(np.float64)
    for i in range(nt):
        x = ThetaSorted[i] - Td
        g = 0.5 * (1 + np.array([math.erf(v / math.sqrt(2 * sigma * sigma)) for v in x]))
        curr = float(np.sum(np.where(x < -thresh, 0.0, np.where(x > thresh, Taud, Taud * g)))) / numvec
        if curr > search:
            if abs(curr - search) < abs(prev - search):
                lam_q = ThetaSorted[i]
            else:
                lam_q = ThetaSorted[i - 1] if i > 0 else ThetaSorted[i]
            break
        prev = curr
        lam_q = ThetaSorted[i]
    lam_q = be.rdtype.type(lam_q)
    mu_q = lam_q * lam_q
    last = Theta[(numvec - 1) * m:]
    idx = 0
    for i in range(m):
        if last[i] > lam_q:
            idx = i - 1
            break
        idx = i + 1
    idx = max(idx, 0)
    if idx > 0:
        be.LanczosDos(idx, m, ritzV)
    ritzv_[:idx] = last[:idx] * last[:idx]
    ritzv_[idx:nevex - 1] = mu_1
    ritzv_[nevex - 1] = mu_q
    for i in range(1, idx):
        j = i * (nevex // idx)
        be.Swap(i, j)
        ritzv_[[i, j]] = ritzv_[[j, i]]
    return be.rdtype.type(b_sup)


def cfg_nev(be):
    return be.nev


def solve_pseudo(be: OracleBackendPseudo, cfg: Config, upperb_scale: float = 1.0) -> Trace:
    """algorithm.inc:1834-2220 (random start vectors)."""
    N, nev, nex, nevex = be.N, be.nev, be.nex, be.nevex
    tr = be.trace
    be.Start()
    unconverged = nevex
    deg = min(cfg.deg + cfg.deg % 2, cfg.max_deg)
    degrees_ = np.zeros(2 * nevex, dtype=np.int64)
    degrees_[:unconverged] = deg
    tol = cfg.tol
    big = np.finfo(be.rdtype).max
    be.resid[:] = big
    residLast_ = np.full(2 * nevex, big, dtype=be.rdtype)
    be.initVecs(True)
    be.QR(0, 1.0)
    lanczos_iter = min(nevex, N // 2, cfg.lanczos_iter)
    lanczos_iter -= lanczos_iter % 2
    upperb = _lanczos_for_h2(be, cfg, N, cfg.num_lanczos, lanczos_iter, nevex, be.ritzv)
    lambda_1 = be.ritzv[:nevex - 1].min()
    lower = be.ritzv[nevex - 1]
    b_sup = upperb * upperb_scale if upperb > 0 else upperb / upperb_scale
    next_lower = lower
    lower = lower * cfg.decaying_rate
    locked = iteration = 0
    early = []
    while locked < nev and unconverged > 0 and iteration < cfg.max_iter:
        ritzv, resid, residLast, degrees = be.ritzv[locked:], be.resid[locked:], residLast_[locked:], degrees_[locked:]
        if iteration > 0:
            next_lower = next_lower * next_lower
            if lambda_1 < next_lower < lower:
                lower = next_lower
        if cfg.opt and iteration != 0:
            deg = _calc_degrees_pseudo_h2(be, cfg, unconverged, nex, b_sup, lower, tol, ritzv, resid, residLast,
                                          degrees, locked)
        _filter_h2(be, unconverged, degrees, lambda_1, lower, b_sup)
        be.ApplyKconjugate(unconverged)
        cc, ee = (b_sup + lower) / 2, (b_sup - lower) / 2
        if ee <= 0:
            ee = abs(lower - b_sup) / 2
        t_1 = (lambda_1 - cc) / ee
        t_k = (ritzv[0] * ritzv[0] - cc) / ee if iteration > 0 else t_1
        rho_1, rho_k = _cheb_rho_c(t_1), _cheb_rho_c(t_k)
        deg_max = int(degrees[:unconverged].max())
        with np.errstate(over="ignore"):
            cond = be.rdtype.type(rho_k) ** be.rdtype.type(degrees[0]) * be.rdtype.type(rho_1) ** be.rdtype.type(
                deg_max - degrees[0])
        be.QR(locked, float(cond))
        be.RR(unconverged)
        tr.iterations += 1
        be.Resd()
        order = sorted(range(unconverged), key=lambda a: ritzv[a])
        next_lower = ritzv[order[int(unconverged * 0.95) - 1]] * cfg.decaying_rate
        new_conv = _locking_pseudo(be, unconverged, nex, tol, ritzv, resid, residLast, locked, iteration, early)
        if new_conv > 0:
            be.ApplyKconjugate(new_conv)
        be.Lock(new_conv)
        locked += new_conv
        unconverged -= new_conv
        iteration += 1
    # positive Ritz values first (ascending), then the rest (ascending); swaps along the permutation cycles
    n = locked + unconverged
    rz, rs = be.ritzv, be.resid
    perm = sorted(range(n), key=lambda i: (not (rz[i] > 0), rz[i]))
    visited = [False] * n
    for i in range(n):
        if visited[i] or perm[i] == i:
            continue
        cyc, cur = [], i
        while not visited[cur]:
            visited[cur] = True
            cyc.append(cur)
            cur = perm[cur]
        r0, s0 = rz[i], rs[i]
        for k in range(len(cyc) - 1):
            rz[cyc[k]] = rz[cyc[k + 1]]
            rs[cyc[k]] = rs[cyc[k + 1]]
            be.Swap(cyc[k], cyc[k + 1])
        rz[cyc[-1]], rs[cyc[-1]] = r0, s0
    return tr


def solve_problem_pseudo(H: np.ndarray, nev: int, nex: int, cfg: Config | None = None):
    """-> (ritzv[:nev+nex], resid[:nev+nex], V (all 2 (nev+nex) columns), trace, backend)."""
    H = np.asfortranarray(H.copy())
    cfg = cfg or Config.for_dtype(H.dtype)
    be = OracleBackendPseudo(H, nev, nex)
    tr = solve_pseudo(be, cfg)
    return be.ritzv[:nev + nex].copy(), be.resid[:nev + nex].copy(), be.V1.copy(), tr, be


# ---- TF32 operand splitting of the single-precision filter product (chase_b200/csrc/hemm_tf32.cuh) ---------------------
# TEST INFRASTRUCTURE: a statement of the operand preparation of the tcgen05 kind::tf32 kernel that replaces the
# cublasSgemm / cublasCgemm of the reference's ChASEGPU::HEMM (external/cublaspp/cublaspp.hpp:563, 623).  The tensor core
# reads the upper 19 bits of an FP32 container (hi = truncation); the kernel keeps lo = rna_tf32(x - hi) beside it and
# forms A B ~ A_hi B_hi + A_hi B_lo + A_lo B_hi (+ A_lo B_lo).
def tf32_trunc(x: np.ndarray) -> np.ndarray:
    """What kind::tf32 sees of an FP32 value: the low 13 mantissa bits dropped."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)
    return b.view(np.float32)


def tf32_rna(x: np.ndarray) -> np.ndarray:
    """cvt.rna.tf32.f32: round to nearest, ties away from zero, to 10 explicit mantissa bits."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((b + np.uint64(0x1000)) & np.uint64(0xFFFFE000)).astype(np.uint32)
    out = r.view(np.float32).copy()
    bad = ~np.isfinite(np.asarray(x, dtype=np.float32))
    out[bad] = np.asarray(x, dtype=np.float32)[bad]
    return out


def tf32_split(x: np.ndarray):
    """(hi, lo) as the kernel uses them: hi = tf32_trunc(x) (implicit, the stored value itself), lo = rna(x - hi)."""
    x = np.asarray(x, dtype=np.float32)
    hi = tf32_trunc(x)
    lo = tf32_rna((x - hi).astype(np.float32))
    return hi, lo


def gemm_tf32_split(A: np.ndarray, B: np.ndarray, terms: int = 3) -> np.ndarray:
    """A @ B from the split operands, products accumulated exactly (float64): isolates the error of the 3- / 4-term
    splitting from the accumulation error of the hardware."""
    Ah, Al = tf32_split(A)
    Bh, Bl = tf32_split(B)
    f = np.float64
    C = Ah.astype(f) @ Bh.astype(f) + Ah.astype(f) @ Bl.astype(f) + Al.astype(f) @ Bh.astype(f)
    if terms >= 4:
        C = C + Al.astype(f) @ Bl.astype(f)
    return C
