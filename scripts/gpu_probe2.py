"""Round-2 GPU probe (run under gpurun): A/B of HEMM kernel variants against same-box cuBLAS.

  * ragged last-column tiles: CHASE_B200_HEMM_SPLIT=0 (one launch, default tile) vs 1 (ragged columns in a second launch on a narrow tile)
  * filter widths of a C2 solve (k = 1400, 1342, 609, 419) and the complex C4-like width
  * distributed local blocks (M x K rectangular, op(A) = A and A^H) with the hybrid schedule on/off
  * Lanczos gemv_conjT: GB/s against MEASURED_PEAKS.json
Writes gpurun_out/probe2.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chase_b200 import kernels as k  # noqa: E402


def timeit(fn, warm=2, it=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e-3


out = {"gpu": torch.cuda.get_device_name(0), "dmma_peak_tflops": max(k.dmma_peak(40000) for _ in range(3)) / 1e12}
print(out, flush=True)
what = sys.argv[1:] or ["square", "rect", "gemv"]

if "square" in what:
    res = []
    for t, n, kc in [("d", 20000, 1400), ("d", 20000, 1342), ("d", 20000, 1300), ("d", 20000, 609), ("d", 20000, 419),
                     ("d", 20000, 200), ("z", 12000, 1400), ("z", 12000, 1342), ("z", 12000, 419)]:
        dt = torch.float64 if t == "d" else torch.complex128
        f = 1 if t == "d" else 4
        ld = (n + 15) // 16 * 16
        A = torch.randn((n, ld), dtype=dt, device="cuda")
        B = torch.randn((kc, ld), dtype=dt, device="cuda")
        C = torch.randn((kc, ld), dtype=dt, device="cuda")
        flops = 2.0 * f * n * n * kc
        r = dict(type=t, n=n, k=kc)
        for alt in ("0", "1"):  # ragged columns: single launch vs split launch
            os.environ["CHASE_B200_HEMM_SPLIT"] = alt
            tk = timeit(lambda: k.hemm(n, kc, 0.5, A, ld, B, ld, -0.25, C, ld, 1.0))
            r[f"alt{alt}_ms"] = tk * 1e3
            r[f"alt{alt}_tflops"] = flops / tk / 1e12
        Cb = torch.empty_like(C)
        tb = timeit(lambda: torch.matmul(B, A, out=Cb))
        r["cublas_ms"], r["cublas_tflops"] = tb * 1e3, flops / tb / 1e12
        print(r, flush=True)
        res.append(r)
        del A, B, C, Cb
    out["hemm_square"] = res
    os.environ.pop("CHASE_B200_HEMM_SPLIT", None)

if "rect" in what:
    # local blocks of C2 on 2x1 / 2x2 / 4x2 grids, and of C4 (z N=120000) on 4x2 scaled to fit quickly
    res = []
    for t, M, K, kc in [("d", 10000, 20000, 1400), ("d", 10000, 10000, 1400), ("d", 5000, 10000, 1400),
                        ("z", 7504, 15008, 1400)]:
        dt = torch.float64 if t == "d" else torch.complex128
        f = 1 if t == "d" else 4
        ldm, ldk = (M + 15) // 16 * 16, (K + 15) // 16 * 16
        A = torch.randn((K, ldm), dtype=dt, device="cuda")  # column-major M x K
        Bk = torch.randn((kc, ldk), dtype=dt, device="cuda")  # K x k
        Bm = torch.randn((kc, ldm), dtype=dt, device="cuda")  # M x k
        flops = 2.0 * f * M * K * kc
        r = dict(type=t, M=M, K=K, k=kc)
        for hyb in ("0", "1"):
            os.environ["CHASE_B200_HEMM_HYBRID"] = hyb
            # op(A) = A:  C(M x k) = A (M x K) B(K x k)
            t0 = timeit(lambda: k.hemm_rect(0, M, K, kc, 1.0, A, ldm, Bk, ldk, 0.0, Bm, ldm))
            # op(A) = A^H: C(K x k) = A^H (K x M) B(M x k)
            t1 = timeit(lambda: k.hemm_rect(1, K, M, kc, 1.0, A, ldm, Bm, ldm, 0.0, Bk, ldk))
            r[f"hybrid{hyb}_n_tflops"] = flops / t0 / 1e12
            r[f"hybrid{hyb}_h_tflops"] = flops / t1 / 1e12
        os.environ.pop("CHASE_B200_HEMM_HYBRID", None)
        At = A[:, :M]  # (K, M) row-major view == A^T
        o1 = torch.empty((kc, M), dtype=dt, device="cuda")
        tb = timeit(lambda: torch.matmul(Bk[:, :K], At, out=o1))
        r["cublas_n_tflops"] = flops / tb / 1e12
        print(r, flush=True)
        res.append(r)
        del A, Bk, Bm, o1
    out["hemm_rect"] = res

if "gemv" in what:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    res = []
    for t, n in [("d", 20000), ("z", 12000), ("d", 40000)]:
        dt = torch.float64 if t == "d" else torch.complex128
        ld = (n + 15) // 16 * 16
        A = torch.randn((n, ld), dtype=dt, device="cuda")
        X = torch.randn((4, ld), dtype=dt, device="cuda")
        Y = torch.zeros((4, ld), dtype=dt, device="cuda")
        tk = timeit(lambda: k.gemv_conjt(n, n, A, ld, X, ld, 4, Y, ld), warm=3, it=10)
        gb = n * ld * A.element_size() / 1e9
        r = dict(type=t, n=n, ms=tk * 1e3, gbs=gb / tk, frac_of_measured_copy_peak=(gb / tk) / peaks.get("hbm_gbs", 6545.9))
        print(r, flush=True)
        res.append(r)
        del A, X, Y
    out["gemv_conjT"] = res

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe2.json"), "w"), indent=1)
