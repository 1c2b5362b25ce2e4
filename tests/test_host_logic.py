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
