"""GPU parity tests of every kernel behind include/chase_b200_kernels.h, called through the C ABI
(ctypes) on torch device buffers and compared with the numpy statement of the reference op
(oracle/chase_oracle.py for the filter step / residual norms; LAPACK via numpy/scipy otherwise).

Tolerances: FP64 kernels 1e-12 relative to the operand norms (north_star asks 1e-10 on eigenvalues),
FP32 storage 1e-5 (north_star: 1e-4)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import chase_oracle as co  # noqa: E402  (checker only)

DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
TOL = {"s": 2e-5, "d": 1e-12, "c": 2e-5, "z": 1e-12}
REF = os.environ.get("CHASE_REFERENCE_ROOT", "/root/reference")


def K():
    from chase_b200 import kernels

    return kernels


def rnd(rng, shape, t):
    a = rng.standard_normal(shape)
    if t in "cz":
        a = a + 1j * rng.standard_normal(shape)
    return a.astype(DT[t])


def relerr(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("t", ["d", "z", "s", "c"])
@pytest.mark.parametrize("ta,tb", [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("M,N,Kd", [(257, 131, 77), (128, 128, 16), (1, 1, 1), (300, 5, 1000), (64, 200, 33)])
def test_gemm_all_ops(t, ta, tb, M, N, Kd):
    k = K()
    rng = np.random.default_rng(M * 7 + N * 3 + Kd + ta * 2 + tb)
    A = rnd(rng, (Kd, M) if ta else (M, Kd), t)
    B = rnd(rng, (N, Kd) if tb else (Kd, N), t)
    C = rnd(rng, (M, N), t)
    alpha, beta = (0.7 - 0.2j, -0.3 + 0.5j) if t in "cz" else (0.7, -0.3)
    opA = A.conj().T if ta else A
    opB = B.conj().T if tb else B
    ref = alpha * (opA.astype(np.complex128 if t in "cz" else np.float64) @ opB) + beta * C
    lda, ldb, ldc = A.shape[0] + 3, B.shape[0] + 1, M + 5
    dA, dB, dC = k.colmajor(A, lda), k.colmajor(B, ldb), k.colmajor(C, ldc)
    k.gemm(ta, tb, M, N, Kd, alpha, dA, lda, dB, ldb, beta, dC, ldc)
    torch.cuda.synchronize()
    out = k.to_numpy(dC, M)
    assert relerr(out, ref) < TOL[t]
    # padding rows of C untouched
    assert np.all(dC.cpu().numpy()[:, M:] == 0)


@pytest.mark.parametrize("t", ["d", "z"])
def test_gemm_beta_zero_ignores_nan_and_splitk(t):
    k = K()
    rng = np.random.default_rng(5)
    M = N = 96
    Kd = 9000  # deep product -> split-K path with a workspace
    A = rnd(rng, (Kd, M), t)
    B = rnd(rng, (Kd, N), t)
    C = np.full((M, N), np.nan, dtype=DT[t])
    dA, dB, dC = k.colmajor(A), k.colmajor(B), k.colmajor(C)
    ws = torch.zeros(64 << 20, dtype=torch.uint8, device="cuda")
    k.gemm(1, 0, M, N, Kd, 1.0, dA, Kd, dB, Kd, 0.0, dC, M, 0, ws)
    torch.cuda.synchronize()
    out = k.to_numpy(dC, M)
    ref = A.conj().T @ B
    assert relerr(out, ref) < 1e-12
    # upper-only variant (Gram matrix of CholQR): strictly-lower tiles may be skipped, upper must match
    dG = k.colmajor(np.zeros((M, N), dtype=DT[t]))
    k.gemm(1, 0, M, N, Kd, 1.0, dA, Kd, dA, Kd, 0.0, dG, M, 1, ws)
    torch.cuda.synchronize()
    G = k.to_numpy(dG, M)
    refG = A.conj().T @ A
    iu = np.triu_indices(M)
    assert np.linalg.norm(G[iu] - refG[iu]) / np.linalg.norm(refG[iu]) < 1e-12


@pytest.mark.parametrize("t", ["d", "z", "s", "c"])
@pytest.mark.parametrize("n,kcols", [(1001, 140), (512, 64), (777, 3), (2048, 256)])
def test_hemm_filter_step_matches_oracle(t, n, kcols):
    """One Chebyshev filter step C <- alpha (A - cI) B + beta C (oracle: gemm_filter_step)."""
    k = K()
    rng = np.random.default_rng(n + kcols)
    A = rnd(rng, (n, n), t)
    A = ((A + A.conj().T) / 2).astype(DT[t])
    B = rnd(rng, (n, kcols), t)
    C = rnd(rng, (n, kcols), t)
    alpha, beta, shift = 0.013, -0.42, 3.7
    wide = np.complex128 if t in "cz" else np.float64
    ref = co.gemm_filter_step(A.astype(wide), B.astype(wide), C.astype(wide), alpha, beta, shift)
    ld = (n + 15) // 16 * 16
    dA, dB, dC = k.colmajor(A, ld), k.colmajor(B, ld), k.colmajor(C, ld)
    k.hemm(n, kcols, alpha, dA, ld, dB, ld, beta, dC, ld, shift)
    torch.cuda.synchronize()
    assert relerr(k.to_numpy(dC, n), ref) < TOL[t]
    # A must not be modified (the shift is folded into the epilogue)
    assert np.array_equal(k.to_numpy(dA, n), A)


@pytest.mark.parametrize("t", ["d", "z", "s", "c"])
def test_hemm_residual_block_and_norms(t):
    """W = A V - V diag(theta); ||W_j|| (oracle: residual_norms; reference residuals.cu)."""
    k = K()
    n, kc = 600, 37
    rng = np.random.default_rng(11)
    A = rnd(rng, (n, n), t)
    A = ((A + A.conj().T) / 2).astype(DT[t])
    V = rnd(rng, (n, kc), t)
    theta = rng.standard_normal(kc)
    wide = np.complex128 if t in "cz" else np.float64
    ref = co.residual_norms(A.astype(wide), V.astype(wide), theta)
    ld = 608
    dA, dV = k.colmajor(A, ld), k.colmajor(V, ld)
    dW = torch.zeros_like(dV)
    dth = torch.from_numpy(theta).cuda()
    out = torch.zeros(kc, dtype=torch.float64, device="cuda")
    k.hemm(n, kc, 1.0, dA, ld, dV, ld, 0.0, dW, ld, 0.0, dth)
    k.colnorms(n, kc, dW, ld, out, True)
    torch.cuda.synchronize()
    assert np.max(np.abs(out.cpu().numpy() - ref) / ref) < TOL[t] * 10


@pytest.mark.parametrize("t", ["d", "z", "s", "c"])
@pytest.mark.parametrize("ta", [0, 1])
@pytest.mark.parametrize("M,Kd,kcols", [(640, 512, 96), (515, 1001, 37), (1001, 300, 130), (96, 64, 5)])
def test_hemm_rect_local_block(t, ta, M, Kd, kcols):
    """Distributed filter step on a rectangular local block: C <- alpha op(A) B + beta C with op = N or ^H
    (reference nccl/hemm.hpp:325-332, 382-389: cublasTgemm(OP_C | OP_N) on the local block)."""
    k = K()
    rng = np.random.default_rng(M + 7 * Kd + ta)
    A = rnd(rng, (Kd, M) if ta else (M, Kd), t)  # stored shape
    B = rnd(rng, (Kd, kcols), t)
    C = rnd(rng, (M, kcols), t)
    alpha, beta = 0.37, -1.25
    wide = np.complex128 if t in "cz" else np.float64
    opA = A.astype(wide).conj().T if ta else A.astype(wide)
    ref = alpha * (opA @ B.astype(wide)) + beta * C.astype(wide)
    lda = (A.shape[0] + 15) // 16 * 16
    ldb = (Kd + 15) // 16 * 16
    ldc = (M + 15) // 16 * 16
    dA, dB, dC = k.colmajor(A, lda), k.colmajor(B, ldb), k.colmajor(C, ldc)
    k.hemm_rect(ta, M, Kd, kcols, alpha, dA, lda, dB, ldb, beta, dC, ldc)
    torch.cuda.synchronize()
    assert relerr(k.to_numpy(dC, M), ref) < TOL[t]


@pytest.mark.parametrize("t", ["d", "z"])
@pytest.mark.parametrize("ta", [0, 1])
def test_hemm_streamk_many_tiles(t, ta):
    """>= 148 output tiles: the stream-K schedule splits tiles across CTAs (head part parked in scratch, tail part
    adds it).  Checked against a float64 matmul, twice (the second launch reuses the scratch slots), plus bitwise
    reproducibility of the two launches."""
    k = K()
    M, Kd, kcols = (4100, 3000, 650) if t == "d" else (3000, 2100, 330)
    rng = np.random.default_rng(5 + ta)
    A = rnd(rng, (Kd, M) if ta else (M, Kd), t)
    B = rnd(rng, (Kd, kcols), t)
    C = rnd(rng, (M, kcols), t)
    alpha, beta = 0.37, -1.25
    opA = A.conj().T if ta else A
    ref = alpha * (opA @ B) + beta * C
    lda = (A.shape[0] + 15) // 16 * 16
    ldb = (Kd + 15) // 16 * 16
    ldc = (M + 15) // 16 * 16
    dA, dB = k.colmajor(A, lda), k.colmajor(B, ldb)
    outs = []
    for _ in range(2):
        dC = k.colmajor(C, ldc)
        k.hemm_rect(ta, M, Kd, kcols, alpha, dA, lda, dB, ldb, beta, dC, ldc)
        torch.cuda.synchronize()
        outs.append(k.to_numpy(dC, M))
        assert relerr(outs[-1], ref) < TOL[t]
    assert np.array_equal(outs[0], outs[1])


def test_hemm_streamk_square_with_shift():
    k = K()
    n, kcols = 4500, 700
    rng = np.random.default_rng(3)
    A = rnd(rng, (n, n), "d")
    A = (A + A.T) / 2
    B = rnd(rng, (n, kcols), "d")
    C = rnd(rng, (n, kcols), "d")
    ref = co.gemm_filter_step(A, B, C, 0.02, -0.4, 1.5)
    ld = (n + 15) // 16 * 16
    dA, dB, dC = k.colmajor(A, ld), k.colmajor(B, ld), k.colmajor(C, ld)
    k.hemm(n, kcols, 0.02, dA, ld, dB, ld, -0.4, dC, ld, 1.5)
    torch.cuda.synchronize()
    assert relerr(k.to_numpy(dC, n), ref) < 1e-12


def _fixture(name):
    p = os.path.join(os.path.dirname(__file__), "golden", "qr_fixtures", name)
    return p if os.path.exists(p) else None


@pytest.mark.parametrize("t", ["d", "z", "s", "c"])
@pytest.mark.parametrize("n", [50, 200, 333])
def test_potrf_trsm_cholqr_orthogonality(t, n):
    """CholQR building blocks: G = V^H V (upper), R = chol(G), Q = V R^-1 -> ||Q^H Q - I||_F / sqrt(n) small
    (the reference's assertion in tests/linalg/internal/cuda/cholqr.cpp:52-160)."""
    k = K()
    rows = 1000
    rng = np.random.default_rng(n)
    V = rnd(rng, (rows, n), t)
    ld, ldg = 1008, (n + 15) // 16 * 16
    dV = k.colmajor(V, ld)
    dG = torch.zeros((n, ldg), dtype=dV.dtype, device="cuda")
    k.gemm(1, 0, n, n, rows, 1.0, dV, ld, dV, ld, 0.0, dG, ldg, 1)
    info = torch.zeros(4, dtype=torch.int32, device="cuda")
    k.potrf(n, dG, ldg, info)
    torch.cuda.synchronize()
    assert int(info[0]) == 0
    R = np.triu(k.to_numpy(dG, n))
    wide = np.complex128 if t in "cz" else np.float64
    Rref = np.linalg.cholesky(V.astype(wide).conj().T @ V.astype(wide)).conj().T
    assert relerr(R, Rref) < TOL[t] * 50
    dX = torch.zeros_like(dV)
    k.trsm(rows, n, dG, ldg, dV, ld, dX, ld)
    torch.cuda.synchronize()
    Q = k.to_numpy(dX, rows).astype(wide)
    orth = np.linalg.norm(Q.conj().T @ Q - np.eye(n)) / np.sqrt(n)
    eps = np.finfo(np.float32 if t in "sc" else np.float64).eps
    assert orth < 200 * eps * n ** 0.5
    assert relerr(Q @ Rref, V.astype(wide)) < TOL[t] * 50


def test_potrf_reports_first_bad_pivot():
    k = K()
    n = 100
    G = np.eye(n)
    G[70, 70] = -1.0
    dG = k.colmajor(G, 112)
    info = torch.zeros(4, dtype=torch.int32, device="cuda")
    k.potrf(n, dG, 112, info)
    torch.cuda.synchronize()
    assert int(info[0]) == 71  # LAPACK convention: 1-based index of the failing leading minor


def test_shift_abstrace():
    k = K()
    n = 77
    rng = np.random.default_rng(3)
    G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    G = (G + G.conj().T)
    dG = k.colmajor(G.astype(np.complex128), 80)
    s = torch.zeros(1, dtype=torch.float64, device="cuda")
    k.shift_abstrace(n, dG, 80, 1e-3, s)
    torch.cuda.synchronize()
    exp = 1e-3 * np.sum(np.abs(np.diag(G)))
    assert abs(float(s[0]) - exp) < 1e-12 * exp
    out = k.to_numpy(dG, n)
    assert np.allclose(np.diag(out), np.diag(G) + exp, rtol=1e-14)
    assert np.array_equal(out - np.diag(np.diag(out)), G - np.diag(np.diag(G)))


@pytest.mark.parametrize("t", ["d", "z", "s", "c"])
@pytest.mark.parametrize("n", [1, 2, 7, 64, 141, 400])
def test_heev_matches_lapack(t, n):
    """Hermitian eigensolver vs LAPACK heevd on Q diag(0.1 (i+1)) Q^H (the reference's own RR test matrix,
    tests/linalg/internal/cuda/rayleighRitz.cpp:55-131, tolerance 100 eps there)."""
    k = K()
    rng = np.random.default_rng(n)
    wide = np.complex128 if t in "cz" else np.float64
    X = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if t in "cz" else 0)
    Q, _ = np.linalg.qr(X)
    lam = 0.1 * (np.arange(n) + 1)
    G = (Q * lam) @ Q.conj().T
    G = ((G + G.conj().T) / 2).astype(DT[t])
    ldg = (n + 15) // 16 * 16
    # only the LOWER triangle may be referenced: poison the strict upper part
    Gp = G.copy()
    Gp[np.triu_indices(n, 1)] = 777.0
    dG = k.colmajor(Gp, ldg)
    dZ = torch.zeros_like(dG)
    w, sweeps, rc = k.heev(n, dG, ldg, dZ, ldg)
    assert rc == 0 and sweeps <= 20
    wref = np.linalg.eigvalsh(G.astype(wide))
    eps = np.finfo(np.float32 if t in "sc" else np.float64).eps
    assert np.max(np.abs(w - wref)) < 100 * eps * max(1.0, np.max(np.abs(wref)))
    assert np.all(np.diff(w) >= 0)
    Z = k.to_numpy(dZ, n).astype(wide)
    Gw = G.astype(wide)
    assert np.linalg.norm(Gw @ Z - Z * w) / np.linalg.norm(Gw) < 50 * eps * np.sqrt(n)
    assert np.linalg.norm(Z.conj().T @ Z - np.eye(n)) < 50 * eps * n


def test_heev_clustered_and_degenerate():
    k = K()
    n = 200
    rng = np.random.default_rng(9)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam = np.concatenate([np.full(50, 1.0), 1.0 + 1e-9 * np.arange(50), np.linspace(-5, 5, 100)])
    G = (Q * lam) @ Q.T
    G = (G + G.T) / 2
    dG = k.colmajor(G, 208)
    dZ = torch.zeros_like(dG)
    w, sweeps, rc = k.heev(n, dG, 208, dZ, 208)
    assert rc == 0
    assert np.max(np.abs(w - np.sort(lam))) < 1e-13 * 5
    Z = k.to_numpy(dZ, n)
    assert np.linalg.norm(Z.T @ Z - np.eye(n)) < 1e-12
    assert np.linalg.norm(G @ Z - Z * w) < 1e-12 * np.linalg.norm(G)


@pytest.mark.parametrize("t,n", [("d", 1400), ("z", 700), ("d", 1029)])
def test_heev_large_projected_matrix(t, n):
    """Rayleigh-Ritz sized problems (BASELINE C2: nev+nex = 1400): dense spectrum 0.01..7 and the nearly diagonal
    matrix of late iterations; eigenvalues to 1e-12 relative to ||G||, orthogonality and residual at n eps level."""
    k = K()
    rng = np.random.default_rng(n)
    lam = np.linspace(0.01, 7.0, n)
    X = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if t == "z" else 0)
    Q, _ = np.linalg.qr(X)
    E = rng.standard_normal((n, n)) * 1e-5
    for G in ((Q * lam) @ Q.conj().T, np.diag(lam) + E + E.T):
        G = ((G + G.conj().T) / 2).astype(DT[t])
        ldg = (n + 15) // 16 * 16
        dG = k.colmajor(G, ldg)
        dZ = torch.zeros_like(dG)
        w, sweeps, rc = k.heev(n, dG, ldg, dZ, ldg)
        assert rc == 0 and sweeps <= 25
        wref = np.linalg.eigvalsh(G)
        assert np.max(np.abs(w - wref)) < 1e-12 * 7.0
        Z = k.to_numpy(dZ, n)
        assert np.linalg.norm(Z.conj().T @ Z - np.eye(n)) < 1e-11
        assert np.linalg.norm(G @ Z - Z * w) / np.linalg.norm(G) < 5e-12


def test_tridiag_eig_batched():
    k = K()
    M, batch = 25, 4
    rng = np.random.default_rng(1)
    d = rng.standard_normal((batch, M))
    e = np.abs(rng.standard_normal((batch, M)))
    dd, de = torch.from_numpy(d).cuda(), torch.from_numpy(e).cuda()
    w = torch.zeros((batch, M), dtype=torch.float64, device="cuda")
    Z = torch.zeros((batch, M, M), dtype=torch.float64, device="cuda")
    k.tridiag_eig(M, batch, dd, de, M, w, Z)
    torch.cuda.synchronize()
    import scipy.linalg as sla

    for b in range(batch):
        wr, Zr = sla.eigh_tridiagonal(d[b], e[b, : M - 1])
        assert np.max(np.abs(w[b].cpu().numpy() - wr)) < 1e-13 * max(1, np.max(np.abs(wr)))
        Zb = Z[b].cpu().numpy().T  # stored column-major
        # first components squared are what the DoS estimate consumes (tau)
        assert np.max(np.abs(Zb[0, :] ** 2 - Zr[0, :] ** 2)) < 1e-12


@pytest.mark.parametrize("t", ["d", "z", "s", "c"])
def test_gemv_conjt_and_lanczos_step(t):
    k = K()
    n, nv, M = 1001, 4, 10
    rng = np.random.default_rng(2)
    A = rnd(rng, (n, n), t)
    X = rnd(rng, (n, nv), t)
    ld = 1008
    dA, dX = k.colmajor(A, ld), k.colmajor(X, ld)
    dY = torch.zeros_like(dX)
    k.gemv_conjt(n, n, dA, ld, dX, ld, nv, dY, ld)
    torch.cuda.synchronize()
    wide = np.complex128 if t in "cz" else np.float64
    ref = A.astype(wide).conj().T @ X.astype(wide)
    assert relerr(k.to_numpy(dY, n), ref) < TOL[t]

    # one fused step, k = 1 (uses v0 and the previous beta)
    v0, v1, v2 = rnd(rng, (n, nv), t), rnd(rng, (n, nv), t), rnd(rng, (n, nv), t)
    d0, d1, d2 = k.colmajor(v0, ld), k.colmajor(v1, ld), k.colmajor(v2, ld)
    dd = torch.zeros((nv, M), dtype=torch.float64, device="cuda")
    de = torch.zeros((nv, M), dtype=torch.float64, device="cuda")
    rb = torch.tensor([0.5, 1.5, 2.0, 0.25], dtype=torch.float64, device="cuda")
    rb0 = rb.cpu().numpy().copy()
    k.lanczos_step(n, nv, 1, M, d0, d1, d2, ld, dd, de, rb)
    torch.cuda.synchronize()
    V0, V1, V2 = v0.astype(wide), v1.astype(wide), v2.astype(wide)
    alpha = np.einsum("ij,ij->j", V1.conj(), V2)
    W = V2 - V1 * alpha - V0 * rb0
    beta = np.linalg.norm(W, axis=0)
    tol = TOL[t] * 20
    assert np.allclose(dd.cpu().numpy()[:, 1], alpha.real, rtol=tol, atol=tol * np.abs(alpha).max())
    assert np.allclose(rb.cpu().numpy(), beta, rtol=tol)
    assert np.allclose(de.cpu().numpy()[:, 1], beta, rtol=tol)
    assert relerr(k.to_numpy(d2, n), W / beta) < tol


@pytest.mark.parametrize("t", ["d", "c"])
def test_copies_gather_normalize(t):
    k = K()
    n, m = 515, 23
    rng = np.random.default_rng(4)
    X = rnd(rng, (n, m), t)
    ld = 528
    dX = k.colmajor(X, ld)
    dY = torch.zeros_like(dX)
    k.lacpy(n, m, dX, ld, dY, ld)
    perm = rng.permutation(m).astype(np.int32)
    dst = np.arange(m, dtype=np.int32)
    dZ = torch.zeros_like(dX)
    k.gather_cols(n, m, torch.from_numpy(perm).cuda(), torch.from_numpy(dst).cuda(), dX, ld, dZ, ld)
    k.normalize_cols(n, m, dY, ld)
    torch.cuda.synchronize()
    assert np.array_equal(k.to_numpy(dZ, n), X[:, perm])
    Y = k.to_numpy(dY, n)
    assert np.allclose(np.linalg.norm(Y, axis=0), 1.0, rtol=1e-5)
    assert relerr(Y, X / np.linalg.norm(X, axis=0)) < TOL[t] * 5


def test_rng_normal_statistics_and_determinism():
    k = K()
    n, m = 4096, 64
    a = torch.zeros((m, n), dtype=torch.float64, device="cuda")
    b = torch.zeros((m, n), dtype=torch.float64, device="cuda")
    k.rng_normal(n, m, a, n, 24141)
    k.rng_normal(n, m, b, n, 24141)
    torch.cuda.synchronize()
    x = a.cpu().numpy().ravel()
    assert np.array_equal(x, b.cpu().numpy().ravel())
    assert abs(x.mean()) < 0.01 and abs(x.std() - 1.0) < 0.01
    assert abs(np.mean(x ** 4) - 3.0) < 0.1


def test_herm_check_shift_mirror():
    k = K()
    n = 300
    rng = np.random.default_rng(8)
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    H = (A + A.conj().T) / 2
    dH = k.colmajor(H, 304)
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    k.herm_check(n, dH, 304, 1e-12, bad)
    torch.cuda.synchronize()
    assert int(bad[0]) == 0
    dA = k.colmajor(A, 304)
    k.herm_check(n, dA, 304, 1e-12, bad)
    torch.cuda.synchronize()
    assert int(bad[0]) > n
    k.herm_mirror(n, dA, 304, 1)
    k.shift_diag(n, dH, 304, -2.5)
    torch.cuda.synchronize()
    M = k.to_numpy(dA, n)
    assert np.array_equal(np.tril(M, -1), np.triu(A, 1).conj().T)
    assert np.allclose(k.to_numpy(dH, n), H - 2.5 * np.eye(n))


@pytest.mark.parametrize("cond", ["10", "1e4", "ill"])
@pytest.mark.parametrize("t,name", [("d", "double"), ("z", "cdouble"), ("s", "float"), ("c", "cfloat")])
def test_reference_qr_fixtures(t, name, cond):
    """The reference's own CholQR fixtures (tests/linalg/internal/QR_matrices, 100x50): plain CholQR is
    orthogonal to 15 eps on cond_10 / cond_1e4 after two rounds, and potrf must FAIL on cond_ill
    (tests/linalg/internal/cuda/cholqr.cpp:52-160); shifted CholQR2 then succeeds within 10 eps."""
    p = _fixture(f"matrix_{name}_cond_{cond}.bin")
    if p is None:
        pytest.skip("fixture not generated")
    k = K()
    V = np.fromfile(p, dtype=DT[t]).reshape(50, 100).T.copy()
    rows, n, ld, ldg = 100, 50, 112, 64
    eps = np.finfo(np.float32 if t in "sc" else np.float64).eps
    wide = np.complex128 if t in "cz" else np.float64

    def round_(dV, shift_scale=None):
        dG = torch.zeros((n, ldg), dtype=dV.dtype, device="cuda")
        k.gemm(1, 0, n, n, rows, 1.0, dV, ld, dV, ld, 0.0, dG, ldg, 1)
        if shift_scale is not None:
            k.shift_abstrace(n, dG, ldg, shift_scale)
        info = torch.zeros(4, dtype=torch.int32, device="cuda")
        k.potrf(n, dG, ldg, info)
        torch.cuda.synchronize()
        if int(info[0]) != 0:
            return int(info[0]), dV
        dX = torch.zeros_like(dV)
        k.trsm(rows, n, dG, ldg, dV, ld, dX, ld)
        return 0, dX

    dV = k.colmajor(V, ld)
    if cond == "ill":
        info, _ = round_(dV.clone())
        if t in "dz":
            assert info > 0
        scale = np.sqrt(rows) * eps if t in "dz" else 10 * eps
        info, dQ = round_(dV.clone(), scale)
        assert info == 0
        for _ in range(2):
            info, dQ = round_(dQ)
            assert info == 0
        lim = 10
    else:
        info, dQ = round_(dV.clone())
        assert info == 0
        info, dQ = round_(dQ)
        assert info == 0
        lim = 15
    Q = k.to_numpy(dQ, rows).astype(wide)
    orth = np.linalg.norm(Q.conj().T @ Q - np.eye(n)) / np.sqrt(n)
    assert orth < lim * eps


@pytest.mark.parametrize("t", ["s", "d", "c", "z"])
@pytest.mark.parametrize("lower", [0, 1])
def test_packed_triangle_roundtrip(t, lower):
    """tri_pack / tri_unpack = LAPACK 'U' / 'L' packed storage (reference cuda/lacpy.cu:837-, 956-): the payload of
    the distributed Gram allreduce."""
    k = K()
    n, ldg = 77, 80
    rng = np.random.default_rng(5)
    G = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if t in "cz" else 0)
    G = G.astype(DT[t])
    dG = k.colmajor(G, ldg)
    dP = torch.zeros(n * (n + 1) // 2, dtype=dG.dtype, device="cuda")
    k.tri_pack(n, dG, ldg, dP, lower)
    P = dP.cpu().numpy()
    ref = np.concatenate([G[j:, j] if lower else G[: j + 1, j] for j in range(n)])
    assert np.array_equal(P, ref)
    dH = torch.full_like(dG, 9.0)
    k.tri_unpack(n, dP, dH, ldg, lower)
    H = k.to_numpy(dH, n)
    mask = np.tril(np.ones((n, n), bool)) if lower else np.triu(np.ones((n, n), bool))
    assert np.array_equal(H[mask], G[mask])
    assert np.all(H[~mask] == 9.0)
