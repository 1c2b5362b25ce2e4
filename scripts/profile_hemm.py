"""Runs the filter HEMM at a BASELINE shape a few times (for `ncu -k regex:hemm_tma`).
usage: python scripts/profile_hemm.py [d|z] [n] [k] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chase_b200 import kernels as k  # noqa: E402

t = sys.argv[1] if len(sys.argv) > 1 else "d"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
kc = int(sys.argv[3]) if len(sys.argv) > 3 else 1400
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dt = torch.float64 if t == "d" else torch.complex128
ld = (n + 15) // 16 * 16
A = torch.randn((n, ld), dtype=dt, device="cuda")
B = torch.randn((kc, ld), dtype=dt, device="cuda")
C = torch.randn((kc, ld), dtype=dt, device="cuda")
for _ in range(reps):
    k.hemm(n, kc, 0.5, A, ld, B, ld, -0.25, C, ld, 1.0)
torch.cuda.synchronize()
print("done", t, n, kc)
