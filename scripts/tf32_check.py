"""GPU check of the tcgen05 kind::tf32 filter product (csrc/hemm_tf32.cuh): accuracy against an FP64 statement of
C <- alpha S A^H S B + beta C - alpha shift B, against plain FP32 (cuBLAS SGEMM/CGEMM, TF32 off), and speed against
cuBLAS.  Writes gpurun_out/tf32_check.json.  Run under `timeout`: a wrong barrier protocol hangs."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chase_b200 import kernels as k  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
out = {"gpu": torch.cuda.get_device_name(0), "cases": [], "perf": []}


def case(cplx, M, K, kc, alpha, beta, shift=0.0, sflip=0, terms=3, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    dt = torch.complex64 if cplx else torch.float32
    ldk, ldm = (K + 15) // 16 * 16, (M + 15) // 16 * 16

    def rnd(r, c):
        x = torch.randn((r, c), generator=g, device="cuda", dtype=torch.float32)
        if cplx:
            x = torch.complex(x, torch.randn((r, c), generator=g, device="cuda", dtype=torch.float32))
        return x

    A = torch.zeros((M, ldk), dtype=dt, device="cuda")  # column-major K x M
    A[:, :K] = rnd(M, K)
    B = torch.zeros((kc, ldk), dtype=dt, device="cuda")  # K x k
    B[:, :K] = rnd(kc, K)
    C = torch.zeros((kc, ldm), dtype=dt, device="cuda")  # M x k
    C[:, :M] = rnd(kc, M)
    C0 = C.clone()
    Alo = k.tf32_lo(A, cplx)
    k.hemm_tf32(M, K, kc, alpha, A, Alo, ldk, B, ldk, beta, C, ldm, shift=shift, sflip=sflip, terms=terms)
    torch.cuda.synchronize()
    # FP64 statement in torch's row-major view: C^T (k x M) = alpha (S B)^T conj(A^s) S + ...
    wide = torch.complex128 if cplx else torch.float64
    Aw, Bw, Cw = A[:, :K].to(wide), B[:, :K].to(wide), C0[:, :M].to(wide)
    Bs = Bw.clone()
    if sflip:
        Bs[:, sflip:] *= -1
    P = Bs @ Aw.conj().T  # (k x M): P[n, m] = sum_kk B[kk, n] conj(A^s[kk, m])
    if sflip:
        P[:, sflip:] *= -1
    ref = alpha * P + beta * Cw
    if shift != 0.0:
        ref = ref - alpha * shift * Bw[:, :M]
    got = C[:, :M].to(wide)
    scale = torch.linalg.norm(ref) / np.sqrt(ref.numel())
    err = float((got - ref).abs().max() / scale)
    # what plain FP32 arithmetic gives for the product itself
    P32 = (B[:, :K] @ A[:, :K].conj().T).to(wide)
    err32 = float((P32 - Bw @ Aw.conj().T).abs().max() / (torch.linalg.norm(P32) / np.sqrt(P32.numel())))
    r = dict(cplx=cplx, M=M, K=K, k=kc, alpha=alpha, beta=beta, shift=shift, sflip=sflip, terms=terms, max_err_rel_rms=err,
             fp32_matmul_err_rel_rms=err32, untouched_padding=bool(torch.equal(C[:, M:], C0[:, M:])))
    print(r, flush=True)
    out["cases"].append(r)
    return err


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


what = sys.argv[1:] or ["acc", "perf"]
if "acc" in what:
    case(False, 256, 256, 64, 1.0, 0.0)
    case(False, 256, 256, 64, 1.0, 0.0, terms=4)
    case(False, 1000, 1000, 300, 0.7, -0.3, shift=1.5)
    case(False, 1000, 1000, 300, 0.7, -0.3, shift=1.5, terms=4)
    case(False, 640, 2000, 130, 1.0, 0.0)  # rectangular block, A^H B
    case(True, 256, 256, 64, 1.0, 0.0)
    case(True, 1000, 1000, 200, 0.7, -0.3, shift=1.5)
    case(True, 1000, 1000, 200, 0.7, -0.3, shift=1.5, terms=4)
    case(True, 512, 512, 100, 1.0, 0.5, shift=-2.0, sflip=256, terms=4)
    case(True, 384, 1500, 70, 1.0, 0.0)
    case(False, 4096, 4096, 1400, 1.0, 0.0)
    case(False, 4096, 4096, 1400, 1.0, 0.0, terms=4)

if "perf" in what:
    for cplx, n, kc in [(False, 20000, 1400), (False, 20000, 419), (True, 20000, 1400), (True, 40000, 1400)]:
        dt = torch.complex64 if cplx else torch.float32
        f = 4 if cplx else 1
        ld = (n + 15) // 16 * 16
        A = torch.randn((n, ld), dtype=dt, device="cuda")
        B = torch.randn((kc, ld), dtype=dt, device="cuda")
        C = torch.zeros((kc, ld), dtype=dt, device="cuda")
        Alo = k.tf32_lo(A, cplx)
        flops = 2.0 * f * n * n * kc
        r = dict(cplx=cplx, n=n, k=kc)
        for terms in (3, 4):
            t = timeit(lambda: k.hemm_tf32(n, n, kc, 0.5, A, Alo, ld, B, ld, 0.0, C, ld, shift=1.0, terms=terms))
            r[f"tf32x{terms}_ms"], r[f"tf32x{terms}_tflops"] = t * 1e3, flops / t / 1e12
        Cb = torch.empty((kc, n), dtype=dt, device="cuda")
        tb = timeit(lambda: torch.matmul(B[:, :n], A[:, :n], out=Cb))
        r["cublas_fp32_ms"], r["cublas_fp32_tflops"] = tb * 1e3, flops / tb / 1e12
        torch.backends.cuda.matmul.allow_tf32 = True
        tb = timeit(lambda: torch.matmul(B[:, :n], A[:, :n], out=Cb))
        torch.backends.cuda.matmul.allow_tf32 = False
        r["cublas_tf32x1_ms"], r["cublas_tf32x1_tflops"] = tb * 1e3, flops / tb / 1e12
        print(r, flush=True)
        out["perf"].append(r)
        del A, B, C, Alo, Cb
        torch.cuda.empty_cache()

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tf32_check.json"), "w"), indent=1)
