"""CPU tests of the host-side pieces that need no GPU: the start-vector stream (against fixtures written by the
unmodified reference CPU binary, `chase_ref_cpu_<t> --initvecs-only`) and the oracle helpers."""
import ctypes
import os

import numpy as np
import pytest

from oracle import chase_oracle as co

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}


@pytest.mark.parametrize("t", ["d", "z", "s", "c"])
def test_start_vectors_equal_reference_binary_output(t):
    from chase_b200 import lib

    ref = np.fromfile(os.path.join(GOLD, f"initvecs_{t}_N16_m3.bin"), dtype=DT[t]).reshape(3, 16).T
    V = np.zeros((20, 3), dtype=DT[t], order="F")  # ldv = 20 > N
    f = getattr(lib(), f"chase_b200_start_vectors_{t}")
    f(ctypes.c_int64(16), ctypes.c_int64(3), V.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(20))
    assert np.array_equal(V[:16], ref)
    assert np.all(V[16:] == 0)
    # and the numpy oracle produces the same block
    assert np.array_equal(co.init_vectors(16, 3, DT[t]), ref)


def test_round_robin_pairing_covers_every_pair_once():
    """The Jacobi tournament ordering used by the device eigensolver (csrc/jacobi.cuh: rr_pair)."""
    for np_ in (2, 4, 6, 26, 100):
        seen = set()
        m = np_ - 1
        for r in range(np_ - 1):
            used = set()
            for a in range(np_ // 2):
                p, q = (m, r) if a == 0 else ((r + a) % m, (r - a + m) % m)
                assert p != q and p not in used and q not in used
                used |= {p, q}
                seen.add((min(p, q), max(p, q)))
            assert len(used) == np_
        assert len(seen) == np_ * (np_ - 1) // 2


def test_hemm_tile_remap_is_a_bijection_with_wave_locality():
    """csrc/hemm_tma.cuh: hemm_tile_remap — the stream-K spans walk a virtual tile index; the map to the raster index
    must be a bijection for every (tiles, CTAs) and must put the s-th tiles of all CTAs next to each other."""
    from chase_b200 import lib

    f = lib().chase_b200_hemm_tile_remap
    f.restype = ctypes.c_longlong
    f.argtypes = [ctypes.c_longlong] * 3
    for G in (1, 2, 7, 148):
        for T in list(range(1, 40)) + [147, 148, 149, 295, 296, 1727, 1728, 157 * 22, 5000]:
            r = [f(v, T, G) for v in range(T)]
            assert sorted(r) == list(range(T)), (T, G)
    # C2 shape: 1727 tiles on 148 CTAs; position s of every CTA -> 148 consecutive raster tiles
    T, G = 1727, 148
    for s in range(T // G):
        block = sorted(f((c * T) // G + s, T, G) for c in range(G))
        assert block == list(range(s * G, (s + 1) * G))


@pytest.mark.parametrize("hybrid", ["0", "1"])
@pytest.mark.parametrize("ntiles,nkt", [(1727, 1250), (148, 40), (149, 7), (295, 64), (296, 64), (300, 3), (3454, 500),
                                        (100, 10), (5, 33)])
def test_hemm_schedule_covers_every_k_block_once_and_orders_the_handover(hybrid, ntiles, nkt):
    """csrc/hemm_tma.cuh: HemmWalk / hemm_schedule replayed on the host for all 148 CTAs: every (tile, k-block) is
    executed exactly once; a tile has at most two parts; the head part (low k, parked) belongs to CTA b and is the
    FIRST part b executes, the tail part belongs to CTA b + 1 (deadlock-free hand-over); with the hybrid schedule all
    CTAs enter the whole-tile phase after the same number of k-blocks (k-aligned waves)."""
    import subprocess
    import sys

    code = f"""
import ctypes, numpy as np, json
from chase_b200 import lib
f = lib().chase_b200_hemm_walk
f.restype = ctypes.c_longlong
f.argtypes = [ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong]
ntiles, nkt, sms = {ntiles}, {nkt}, 148
cover = np.zeros((ntiles, nkt), dtype=np.int32)
parts_of = {{}}
lead = []
for b in range(sms):
    buf = np.zeros(3 * 64, dtype=np.int64)
    n = f(ntiles, nkt, sms, b, buf.ctypes.data, 64)
    if n < 0:
        continue
    assert n <= 64
    ps = buf[:3 * n].reshape(n, 3)
    sk = 0
    for i, (t, k0, k1) in enumerate(ps):
        assert 0 <= t < ntiles and 0 <= k0 < k1 <= nkt
        cover[t, k0:k1] += 1
        parts_of.setdefault(int(t), []).append((b, i, int(k0), int(k1), n))
        if not (k0 == 0 and k1 == nkt):
            sk = i + 1
    whole = [i for i, (t, k0, k1) in enumerate(ps) if k0 == 0 and k1 == nkt]
    lead.append(int(sum(k1 - k0 for (t, k0, k1) in ps[:sk])))
assert cover.min() == 1 and cover.max() == 1
for t, pl in parts_of.items():
    assert len(pl) <= 2
    if len(pl) == 2:
        head = min(pl, key=lambda x: x[2]); tail = max(pl, key=lambda x: x[2])
        assert head[2] == 0 and head[3] == tail[2] and tail[3] == nkt
        assert tail[0] == head[0] + 1      # tail owner is the next CTA
        assert head[1] == 0                # the head is the first thing its CTA does
print(json.dumps(lead))
"""
    env = dict(os.environ, CHASE_B200_HEMM_HYBRID=hybrid)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True,
                         cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert out.returncode == 0, out.stderr[-2000:]


@pytest.mark.parametrize("n", [1, 2, 5, 49, 64, 120])
def test_host_tridiagonal_eigensolver_matches_lapack(n):
    """host/tridiag_host.hpp (Lanczos with more than 48 steps; the reference uses LAPACK ?stemr on the host,
    cuda/lanczos.hpp:270-299): eigenvalues ascending, orthonormal eigenvectors, T Z = Z diag(w)."""
    import scipy.linalg as sla

    from chase_b200 import lib

    rng = np.random.default_rng(n)
    d = rng.standard_normal(n)
    e = np.abs(rng.standard_normal(max(n - 1, 1))) + 0.1
    w = np.zeros(n)
    Z = np.zeros((n, n), order="F")
    f = lib().chase_b200_tridiag_eig_host
    rc = f(ctypes.c_int(n), d.ctypes.data_as(ctypes.c_void_p), e.ctypes.data_as(ctypes.c_void_p),
           w.ctypes.data_as(ctypes.c_void_p), Z.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    T = np.diag(d) + (np.diag(e[: n - 1], 1) + np.diag(e[: n - 1], -1) if n > 1 else 0)
    wr = np.linalg.eigvalsh(T) if n > 1 else d.copy()
    scale = max(np.abs(wr).max(), 1.0)
    assert np.all(np.diff(w) >= 0)
    assert np.max(np.abs(w - wr)) < 1e-13 * scale * n
    assert np.linalg.norm(Z.T @ Z - np.eye(n)) < 1e-12 * n
    assert np.linalg.norm(T @ Z - Z * w) < 1e-12 * scale * n
    if n > 1:
        wl, Zl = sla.eigh_tridiagonal(d, e[: n - 1])
        assert np.max(np.abs(np.abs(Z[0]) - np.abs(Zl[0]))) < 1e-9  # the DoS weights |z_0k|^2 agree


def test_tf32_split_statement_of_the_fp32_filter_product():
    """oracle.tf32_split / gemm_tf32_split (statement of the operand preparation of csrc/hemm_tf32.cuh): hi is a TF32
    value, hi + lo reproduces x to 2^-21, one TF32 product is 1e-3 accurate, three partial products reach 2^-20 of the
    dot-product scale (a bias: the dropped lo*lo term has the sign of the product) and four reach the FP32 level -- why the
    filter uses 3 terms and the Rayleigh-Ritz / residual products 4."""
    from oracle import chase_oracle as co

    rng = np.random.default_rng(1)
    x = (rng.standard_normal(100000) * np.exp(rng.uniform(-20, 20, 100000))).astype(np.float32)
    hi, lo = co.tf32_split(x)
    assert np.all((hi.view(np.uint32) & 0x1FFF) == 0) and np.all((lo.view(np.uint32) & 0x1FFF) == 0)
    assert np.all(np.abs(hi) <= np.abs(x)) and np.all(np.sign(lo) * np.sign(x) >= 0)  # truncation: lo has x's sign
    rel = np.abs((hi.astype(np.float64) + lo.astype(np.float64)) - x) / np.abs(x)
    assert rel.max() <= 2.0 ** -21
    M, K, N = 64, 4096, 48
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = rng.standard_normal((K, N)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    scale = np.sqrt(K)  # RMS of an entry of the result
    e1 = np.abs(co.tf32_trunc(A).astype(np.float64) @ co.tf32_trunc(B).astype(np.float64) - ref).max() / scale
    e3 = np.abs(co.gemm_tf32_split(A, B, 3) - ref).max() / scale
    e4 = np.abs(co.gemm_tf32_split(A, B, 4) - ref).max() / scale
    e32 = np.abs((A @ B).astype(np.float64) - ref).max() / scale
    assert 1e-4 < e1 < 1e-2  # plain TF32: not FP32-accurate
    assert e3 < 2e-5 and e4 < 1e-6 and e4 < e3
    assert e32 < 2e-5  # FP32 GEMM on the same data, for scale
