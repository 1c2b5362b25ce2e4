"""GPU probe (run under gpurun): FP64 tensor-pipe peak, HEMM kernel rate at BASELINE shapes and the same-box
cuBLAS DGEMM/ZGEMM rate (torch.matmul) for comparison.  Writes gpurun_out/probe.json."""
import json
import os
import sys
import time

import numpy as np
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


out = {"gpu": torch.cuda.get_device_name(0)}
out["dmma_peak_tflops"] = [k.dmma_peak(40000) / 1e12 for _ in range(3)]
print(out, flush=True)

shapes = [("d", 20000, 1400), ("d", 20000, 512), ("z", 12000, 1400), ("d", 8192, 1400)]
if len(sys.argv) > 1:
    shapes = [(a.split(",")[0], int(a.split(",")[1]), int(a.split(",")[2])) for a in sys.argv[1:]]
res = []
for t, n, kc in shapes:
    dt = torch.float64 if t == "d" else torch.complex128
    f = 1 if t == "d" else 4
    ld = (n + 15) // 16 * 16
    A = torch.randn((n, ld), dtype=dt, device="cuda")
    B = torch.randn((kc, ld), dtype=dt, device="cuda")
    C = torch.randn((kc, ld), dtype=dt, device="cuda")
    path = k.hemm_path(1 if t == "d" else 3, n, kc, ld, ld, ld)
    tk = timeit(lambda: k.hemm(n, kc, 0.5, A, ld, B, ld, -0.25, C, ld, 1.0))
    # cuBLAS on the same shape: C^T(k x n) = B^T(k x n) A^T(n x n) in torch's row-major view
    Cb = torch.empty_like(C)
    tb = timeit(lambda: torch.matmul(B, A, out=Cb))
    flops = 2.0 * f * n * n * kc
    r = dict(type=t, n=n, k=kc, path=path, ours_ms=tk * 1e3, ours_tflops=flops / tk / 1e12, cublas_ms=tb * 1e3,
             cublas_tflops=flops / tb / 1e12)
    print(r, flush=True)
    res.append(r)
    del A, B, C, Cb
out["hemm"] = res
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
