"""Pins oracle/chase_oracle.py (numpy restatement) against the golden traces of the
UNMODIFIED reference CPU solver (tests/golden/*.json, see make_golden.py)."""
import numpy as np
import pytest

from oracle import chase_oracle as co
from tests.golden_util import DT, load, parse_trace


def _matrix(g):
    dt = DT[g["type"]]
    if g["matrix"] == "clement":
        return co.clement(g["N"], dt)
    if g["matrix"] == "uniform":
        return co.uniform_diag(g["N"], dt)
    raise ValueError(g["matrix"])


def _oracle_trace(tr):
    hemm = [tuple(int(x) for x in c.split()[1:3]) for c in tr.calls if c.startswith("HEMM")]
    qr = [(int(c.split()[1]), float(c.split()[2])) for c in tr.calls if c.startswith("QR")]
    locks = [int(c.split()[1]) for c in tr.calls if c.startswith("Lock")]
    return hemm, qr, locks


@pytest.mark.parametrize(
    "name",
    ["c1_clement_d_N1001", "serial_clement_d_N256", "serial_clement_z_N256", "c2s_uniform_d_N2000", "noopt_clement_d_N300"],
)
def test_oracle_matches_reference_trace(name):
    g = load(name)
    p = g["problems"][0]
    ref = parse_trace(p["trace"])
    cfg = co.Config.for_dtype(DT[g["type"]])
    cfg.tol, cfg.deg, cfg.opt = g["tol"], g["deg"], bool(g["opt"])
    rv, rs, V, tr, be = co.solve_problem(_matrix(g), g["nev"], g["nex"], cfg)
    hemm, qr, locks = _oracle_trace(tr)
    assert tr.iterations == p["iterations"]
    assert tr.filtered_vecs == p["filtered_vecs"]
    assert hemm == [(b, o) for (b, o, _, _) in ref["hemm"]]  # identical degree schedule
    assert locks == ref["locks"]
    assert [q[0] for q in qr] == [q[0] for q in ref["qr"]]
    for (_, c1), (_, c2) in zip(qr, ref["qr"]):
        assert c1 == pytest.approx(c2, rel=1e-6)
    nev = g["nev"]
    refv = np.array(p["ritzv"][:nev])
    # north_star tolerance: eigenvalues within 1e-10 relative in FP64
    assert np.max(np.abs(rv[:nev] - refv) / np.abs(refv)) < 1e-10
    assert np.all(rs[:nev] <= g["tol"] * 100)  # early locking allows < 100 tol (algorithm.inc:543-544)
    assert tr.swaps == p["swaps"]


def test_clement_known_answer():
    """Known-answer: lowest Clement eigenvalues are -N, -N+2, ... (tests/noinput.cpp matrix)."""
    g = load("c1_clement_d_N1001")
    rv = np.array(g["problems"][0]["ritzv"][:100])
    assert np.allclose(rv, -1001 + 2 * np.arange(100), atol=1e-9)


def test_start_vectors_match_libstdcxx_stream():
    v = co.init_vectors(8, 3, np.float64)
    s = co.mt_normal(1337, 24)
    assert np.array_equal(v.T.reshape(-1), s)
    z = co.init_vectors(8, 3, np.complex128)
    s = co.mt_normal(1337, 48)
    assert np.array_equal(z.T.reshape(-1).imag, s[0::2])
    assert np.array_equal(z.T.reshape(-1).real, s[1::2])


@pytest.mark.parametrize("name", ["hhqr_clement_d_N300", "hhqr_clement_z_N256"])
def test_oracle_with_householder_qr_matches_reference_trace(name):
    """Goldens produced with CHASE_DISABLE_CHOLQR=1 (Householder QR in every iteration, chase_cpu.hpp:670-690)."""
    g = load(name)
    p = g["problems"][0]
    ref = parse_trace(p["trace"])
    cfg = co.Config.for_dtype(DT[g["type"]])
    cfg.tol, cfg.deg, cfg.opt = g["tol"], g["deg"], bool(g["opt"])
    H = np.asfortranarray(_matrix(g).copy())
    be = co.OracleBackend(H, g["nev"], g["nex"])
    be.disable_cholqr = True
    tr = co.solve(be, cfg)
    hemm, qr, locks = _oracle_trace(tr)
    assert tr.iterations == p["iterations"]
    assert tr.filtered_vecs == p["filtered_vecs"]
    assert hemm == [(b, o) for (b, o, _, _) in ref["hemm"]]
    assert locks == ref["locks"]
    assert be.qr_variants[0] == "chol1" and set(be.qr_variants[1:]) == {"householder"}
    nev = g["nev"]
    refv = np.array(p["ritzv"][:nev])
    assert np.max(np.abs(be.ritzv[:nev] - refv) / np.abs(refv)) < 1e-10


@pytest.mark.parametrize("name", ["seq_clement_d_N400", "seq_clement_z_N400"])
def test_oracle_sequence_with_approximate_start_matches_reference_trace(name):
    """tests/noinput.cpp-style sequence (oracle/ref_driver.cpp --seq 3): problem 0 from random vectors, the following
    ones perturbed element-wise and solved in mode 'A' from the previous eigenvectors / Ritz values."""
    g = load(name)
    dt = DT[g["type"]]
    N, nev, nex = g["N"], g["nev"], g["nex"]
    H = np.asfortranarray(co.clement(N, dt))
    be = co.OracleBackend(H, nev, nex)
    nseq = len(g["problems"])
    stream = co.mt_normal(1337, (2 if g["type"] == "z" else 1) * nseq * N * N)
    pos = 0
    for idx, p in enumerate(g["problems"]):
        ref = parse_trace(p["trace"])
        cfg = co.Config.for_dtype(dt)
        cfg.tol, cfg.deg, cfg.opt, cfg.approx = g["tol"], g["deg"], bool(g["opt"]), idx > 0
        be.trace = co.Trace()
        tr = co.solve(be, cfg)
        hemm, qr, locks = _oracle_trace(tr)
        assert tr.iterations == p["iterations"], idx
        assert tr.filtered_vecs == p["filtered_vecs"], idx
        assert hemm == [(b, o) for (b, o, _, _) in ref["hemm"]], idx
        assert locks == ref["locks"], idx
        refv = np.array(p["ritzv"][:nev])
        assert np.max(np.abs(be.ritzv[:nev] - refv) / np.abs(refv)) < 1e-10
        if idx + 1 < nseq:
            pos += co.perturb_hermitian(be.H, stream[pos:], 1e-4)
