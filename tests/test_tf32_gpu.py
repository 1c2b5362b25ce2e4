"""The tcgen05 kind::tf32 filter product (chase_b200/csrc/hemm_tf32.cuh) against an FP64 statement of

    C <- alpha S (A^s)^H S B + beta C - alpha shift_j B

through the kernel C ABI (chase_b200_hemm_tf32_{s,c}), for float and complex<float>, ragged shapes, rectangular blocks,
the pseudo-Hermitian sign flip and per-column shifts.  Tolerance: 3e-5 of the result's RMS, i.e. the error level of a
plain FP32 GEMM (cuBLAS SGEMM measures 2e-6 .. 1.5e-5 on the same inputs); a TF32-only product would be at 1e-3."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

TOL = 3e-5


def _run(cplx, M, K, kc, alpha, beta, shift=0.0, sflip=0, terms=3, theta=False, seed=0):
    from chase_b200 import kernels as k

    g = torch.Generator(device="cuda").manual_seed(seed)
    dt = torch.complex64 if cplx else torch.float32
    wide = torch.complex128 if cplx else torch.float64
    ldk, ldm = (K + 15) // 16 * 16, (M + 15) // 16 * 16

    def rnd(r, c):
        x = torch.randn((r, c), generator=g, device="cuda", dtype=torch.float32)
        if cplx:
            x = torch.complex(x, torch.randn((r, c), generator=g, device="cuda", dtype=torch.float32))
        return x

    A = torch.zeros((M, ldk), dtype=dt, device="cuda")  # the stored matrix, K x M column-major
    A[:, :K] = rnd(M, K)
    B = torch.zeros((kc, ldk), dtype=dt, device="cuda")
    B[:, :K] = rnd(kc, K)
    C = torch.full((kc, ldm), 7.0, dtype=dt, device="cuda")
    C[:, :M] = rnd(kc, M)
    C0 = C.clone()
    th = torch.linspace(-1.0, 2.0, kc, dtype=torch.float64, device="cuda") if theta else None
    Alo = k.tf32_lo(A, cplx)
    k.hemm_tf32(M, K, kc, alpha, A, Alo, ldk, B, ldk, beta, C, ldm, shift=shift, theta=th, sflip=sflip, terms=terms)
    torch.cuda.synchronize()
    Aw, Bw, Cw = A[:, :K].to(wide), B[:, :K].to(wide), C0[:, :M].to(wide)
    Bs = Bw.clone()
    if sflip:
        Bs[:, sflip:] *= -1
    P = Bs @ Aw.conj().T
    if sflip:
        P[:, sflip:] *= -1
    ref = alpha * P + beta * Cw
    if theta:
        ref = ref - alpha * th[:, None] * Bw[:, :M]
    elif shift != 0.0:
        ref = ref - alpha * shift * Bw[:, :M]
    got = C[:, :M].to(wide)
    scale = float(torch.linalg.norm(ref)) / np.sqrt(ref.numel())
    err = float((got - ref).abs().max()) / scale
    assert torch.equal(C[:, M:], C0[:, M:]), "padding rows written"
    return err


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("terms", [3, 4])
def test_tf32_square_filter_step(cplx, terms):
    assert _run(cplx, 1000, 1000, 300, 0.7, -0.3, shift=1.5, terms=terms) < TOL


@pytest.mark.parametrize("cplx,M,K,kc", [(False, 256, 256, 8), (False, 640, 2000, 130), (False, 129, 515, 257),
                                          (True, 384, 1500, 70), (True, 131, 700, 129), (False, 2048, 6000, 700)])
def test_tf32_rectangular_and_ragged(cplx, M, K, kc):
    assert _run(cplx, M, K, kc, 1.0, 0.0) < TOL


def test_tf32_long_accumulation_chain_stays_fp32_accurate():
    """K = 16384: 2048 k-steps.  A single TMEM accumulation chain drifts by ~7e-4 here (truncating adds); the chunked
    drain keeps the error at the FP32 level."""
    assert _run(False, 256, 16384, 64, 1.0, 0.0) < TOL
    assert _run(True, 256, 8192, 40, 1.0, 0.0) < TOL


def test_tf32_pseudo_hermitian_sign_flip_and_column_shifts():
    assert _run(True, 512, 512, 100, 1.0, 0.5, shift=-2.0, sflip=256, terms=4) < TOL
    assert _run(True, 600, 600, 90, 1.0, 0.0, theta=True, terms=4) < TOL
    assert _run(False, 600, 600, 90, 0.5, 0.25, theta=True, terms=4) < TOL


def test_fp32_solves_use_the_tcgen05_kernel():
    """FP32 problems through ?chase_ now run their HEMMs on the kind::tf32 kernel (no FP64 copy of the matrix):
    eigenvalues to 1e-4 of the reference CPU solver's FP32 runs, residuals below tolerance, iteration count within one
    of the reference's (FP32 runs of the reference itself differ by that much between BLAS builds: every residual near
    a ceil() boundary of the degree formula flips a degree)."""
    import chase_b200
    from oracle import chase_oracle as co
    from tests.golden_util import DT, load

    for name in ("serial_clement_s_N256", "serial_clement_c_N256"):
        g = load(name)
        H = co.clement(g["N"], DT[g["type"]])
        p = g["problems"][0]
        with chase_b200.ChASE(H, g["nev"], g["nex"]) as s:
            res = s.solve(deg=g["deg"], tol=g["tol"])
        nev = g["nev"]
        refv = np.array(p["ritzv"][:nev])
        assert np.max(np.abs(res.ritzv[:nev] - refv) / np.abs(refv)) < 1e-4
        assert np.all(res.resid[:nev] < 100 * g["tol"])
        assert abs(res.iterations - p["iterations"]) <= 1, (name, res.iterations, p["iterations"])


@pytest.mark.parametrize("t", ["d", "z"])
def test_mixed_precision_filter_reaches_double_precision_results(t):
    """chase_b200_set_mixed_precision_ (the reference's ENABLE_MIXED_PRECISION, pchase_gpu.hpp:785-881): the first
    filters of a double-precision problem run in single precision on the tcgen05 kernel (residuals > 1e-3), the rest in
    double; the converged eigenpairs must be as good as a pure double-precision solve's."""
    import ctypes

    import chase_b200
    from oracle import chase_oracle as co

    N, nev, nex = 1500, 80, 40
    lam = co.uniform_spectrum(N)
    H = co.dense_from_spectrum(lam, np.float64 if t == "d" else np.complex128)
    L = chase_b200.lib()
    with chase_b200.ChASE(H, nev, nex) as s:
        ref = s.solve()
        assert L.chase_b200_last_sp_filter_cols_() == 0
        L.chase_b200_set_mixed_precision_(ctypes.byref(ctypes.c_int(1)))
        try:
            res = s.solve()
            sp_cols = L.chase_b200_last_sp_filter_cols_()
        finally:
            L.chase_b200_set_mixed_precision_(ctypes.byref(ctypes.c_int(0)))
        again = s.solve()
    assert sp_cols >= (nev + nex) * 20  # at least the first filter (degree 20 on every column)
    assert sp_cols < res.filtered_vecs  # ... and not all of them
    assert np.max(np.abs(res.ritzv[:nev] - lam[:nev]) / lam[:nev]) < 1e-10
    assert np.all(res.resid[:nev] < 1e-8)
    V = res.V[:, :nev]
    assert np.all(np.linalg.norm(H @ V - V * res.ritzv[:nev], axis=0) < 1e-8)
    # switching it off again restores the double-precision run bit for bit
    assert again.filtered_vecs == ref.filtered_vecs and np.array_equal(again.ritzv, ref.ritzv)
