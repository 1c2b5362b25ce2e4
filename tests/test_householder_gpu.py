"""Householder QR: the reference's fallback when CholQR breaks down or qr == 'H'
(cuda::houseHoulderQR = cusolver geqrf + orgqr/ungqr, linalg/internal/cuda/cholqr.hpp:524-556; CPU: LAPACK through
cpu/cholqr1.hpp:199-215).  Kernel against LAPACK (numpy.linalg.qr uses the same geqrf + orgqr conventions) on random
matrices and on the reference's own ill-conditioned QR fixtures; full solves with Householder in every iteration
(CHASE_DISABLE_CHOLQR=1) against golden traces of the unmodified reference CPU solver run the same way."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import chase_oracle as co  # noqa: E402  (checker only)
from tests.golden_util import DT, GOLDEN, load, parse_trace  # noqa: E402

DTK = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}


def K():
    from chase_b200 import kernels

    return kernels


def _rnd(rng, shape, t):
    a = rng.standard_normal(shape)
    if t in "cz":
        a = a + 1j * rng.standard_normal(shape)
    return a.astype(DTK[t])


@pytest.mark.parametrize("t", ["d", "z", "s", "c"])
@pytest.mark.parametrize("rows,n", [(100, 50), (513, 97), (2000, 320), (64, 64), (40, 1)])
def test_hhqr_matches_lapack(t, rows, n):
    k = K()
    rng = np.random.default_rng(rows + n)
    A = _rnd(rng, (rows, n), t)
    lda, ldq = rows + 5, rows + 3
    dA, dQ = k.colmajor(A, lda), k.colmajor(np.zeros_like(A), ldq)
    k.hhqr(rows, n, dA, lda, dQ, ldq)
    torch.cuda.synchronize()
    wide = np.complex128 if t in "cz" else np.float64
    Q = k.to_numpy(dQ, rows).astype(wide)
    R = np.triu(k.to_numpy(dA, rows).astype(wide)[:n])
    eps = np.finfo(np.float32 if t in "sc" else np.float64).eps
    assert np.linalg.norm(Q.conj().T @ Q - np.eye(n)) < 40 * eps * np.sqrt(n)
    assert np.linalg.norm(Q @ R - A) < 40 * eps * np.linalg.norm(A)
    assert np.all(np.abs(np.diag(R).imag) == 0)  # LAPACK convention: real diagonal of R
    Qr, Rr = np.linalg.qr(A.astype(wide))
    # same reflectors => same factors (not just up to signs); conditioning of random A is mild
    assert np.linalg.norm(Q - Qr) < 2e3 * eps * np.sqrt(n)
    assert np.all(dQ.cpu().numpy()[:, rows:] == 0)


@pytest.mark.parametrize("cond", ["10", "1e4", "ill"])
@pytest.mark.parametrize("t,name", [("d", "double"), ("z", "cdouble"), ("s", "float"), ("c", "cfloat")])
def test_hhqr_on_reference_qr_fixtures(t, name, cond):
    """Orthogonality to O(eps) independent of the conditioning — what CholQR cannot deliver on cond_ill
    (tests/linalg/internal/cuda/cholqr.cpp:52-160)."""
    p = os.path.join(GOLDEN, "qr_fixtures", f"matrix_{name}_cond_{cond}.bin")
    if not os.path.exists(p):
        pytest.skip("fixture missing")
    k = K()
    V = np.fromfile(p, dtype=DTK[t]).reshape(50, 100).T.copy()
    rows, n, ld = 100, 50, 112
    dV, dQ = k.colmajor(V, ld), k.colmajor(np.zeros_like(V), ld)
    k.hhqr(rows, n, dV, ld, dQ, ld)
    torch.cuda.synchronize()
    wide = np.complex128 if t in "cz" else np.float64
    Q = k.to_numpy(dQ, rows).astype(wide)
    eps = np.finfo(np.float32 if t in "sc" else np.float64).eps
    assert np.linalg.norm(Q.conj().T @ Q - np.eye(n)) / np.sqrt(n) < 10 * eps
    # same column space: projecting V onto Q loses nothing
    Vw = V.astype(wide)
    assert np.linalg.norm(Vw - Q @ (Q.conj().T @ Vw)) < 100 * eps * np.linalg.norm(Vw)


@pytest.mark.parametrize("name", ["hhqr_clement_d_N300", "hhqr_clement_z_N256"])
def test_solve_with_householder_matches_reference_trace(name, monkeypatch):
    import chase_b200

    g = load(name)
    p = g["problems"][0]
    H = co.clement(g["N"], DT[g["type"]])
    monkeypatch.setenv("CHASE_DISABLE_CHOLQR", "1")
    with chase_b200.ChASE(H, g["nev"], g["nex"]) as s:
        res = s.solve(deg=g["deg"], tol=g["tol"], trace=True)
    monkeypatch.delenv("CHASE_DISABLE_CHOLQR")
    ref, got = parse_trace(p["trace"]), parse_trace(res.trace)
    nev = g["nev"]
    assert all(q == "householder" for q in res.qr_log[1:]) and len(res.qr_log) == res.iterations + 1
    assert res.iterations == p["iterations"]
    assert res.filtered_vecs == p["filtered_vecs"]
    assert [(b, o) for (b, o, _, _) in got["hemm"]] == [(b, o) for (b, o, _, _) in ref["hemm"]]
    assert got["locks"] == ref["locks"]
    refv = np.array(p["ritzv"][:nev])
    assert np.max(np.abs(res.ritzv[:nev] - refv) / np.abs(refv)) < 1e-10
    assert np.all(res.resid[:nev] < 100 * g["tol"])


def test_qr_char_H_selects_householder():
    """?chase_(..., qr = 'H') = SetCholQR(false) (chase_c_interface.cpp:455)."""
    import chase_b200

    H = co.clement(200, np.float64)
    with chase_b200.ChASE(H, 20, 10) as s:
        res = s.solve(deg=16, tol=1e-10, qr="H")
    assert "householder" in res.qr_log
    assert np.allclose(res.ritzv[:20], -200 + 2 * np.arange(20), atol=1e-8)
