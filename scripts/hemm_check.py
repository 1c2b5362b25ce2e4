"""Correctness + timing of the filter HEMM under its schedule variants (env read at every launch):
plain stream-K walk, virtual tile order (default), hybrid data-parallel + stream-K tail.
usage: python scripts/hemm_check.py [n] [k]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chase_b200 import kernels as K  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
kc = int(sys.argv[2]) if len(sys.argv) > 2 else 1400
ld = (n + 15) // 16 * 16
out = {"n": n, "k": kc}
for t, dt in (("d", torch.float64), ("z", torch.complex128)):
    nn = n if t == "d" else min(n, 12000)
    l2 = (nn + 15) // 16 * 16
    A = torch.randn((nn, l2), dtype=dt, device="cuda")
    B = torch.randn((kc, l2), dtype=dt, device="cuda")
    C0 = torch.randn((kc, l2), dtype=dt, device="cuda")
    # column-major views: A is (nn x nn) with ld l2 -> reference in torch on the transposed storage
    ref = 0.5 * (B[:, :nn] @ A[:, :nn] - 1.0 * B[:, :nn]) - 0.25 * C0[:, :nn]  # (C^T = B^T A^T): rows = columns of C
    for mode, env in (("default", {}), ("plain", {"CHASE_B200_HEMM_REMAP": "0", "CHASE_B200_HEMM_HYBRID": "0"}),
                      ("remap", {"CHASE_B200_HEMM_REMAP": "1", "CHASE_B200_HEMM_HYBRID": "0"}),
                      ("hybrid", {"CHASE_B200_HEMM_REMAP": "1", "CHASE_B200_HEMM_HYBRID": "1"})):
        for key in ("CHASE_B200_HEMM_REMAP", "CHASE_B200_HEMM_HYBRID"):
            os.environ.pop(key, None)
        os.environ.update(env)
        C = C0.clone()
        K.hemm(nn, kc, 0.5, A, l2, B, l2, -0.25, C, l2, 1.0)
        torch.cuda.synchronize()
        err = float((C[:, :nn] - ref).norm() / ref.norm())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            K.hemm(nn, kc, 0.5, A, l2, B, l2, -0.25, C, l2, 1.0)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        f = 1 if t == "d" else 4
        out[f"{t}_{mode}"] = {"rel_err": err, "ms": ms, "tflops": 2 * f * nn * nn * kc / ms / 1e9}
    del A, B, C0, ref
print(json.dumps(out))
