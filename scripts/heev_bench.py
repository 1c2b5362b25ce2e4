"""Times the dense Hermitian eigensolver (chase_b200_heev) at Rayleigh-Ritz sizes.
usage: [CHASE_B200_HEEV=1|2] python scripts/heev_bench.py [d|z] n [n ...]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chase_b200 import kernels as k  # noqa: E402

t = sys.argv[1]
for n in [int(x) for x in sys.argv[2:]]:
    rng = np.random.default_rng(n)
    lam = np.linspace(0.01, 7.0, n)
    X = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if t == "z" else 0)
    Q, _ = np.linalg.qr(X)
    E = rng.standard_normal((n, n)) * 1e-4
    for name, G in (("dense", (Q * lam) @ Q.conj().T), ("neardiag", np.diag(lam) + E + E.T)):
        G = (G + G.conj().T) / 2
        ldg = (n + 15) // 16 * 16
        dG = k.colmajor(G, ldg)
        dZ = torch.zeros_like(dG)
        k.heev(n, dG, ldg, dZ, ldg)
        torch.cuda.synchronize()
        t0 = time.time()
        w, sweeps, rc = k.heev(n, dG, ldg, dZ, ldg)
        torch.cuda.synchronize()
        dt = time.time() - t0
        err = np.max(np.abs(w - np.linalg.eigvalsh(G)))
        print(f"method={os.environ.get('CHASE_B200_HEEV', '0')} type={t} n={n} {name}: {dt * 1e3:.1f} ms, "
              f"{sweeps} sweeps, rc={rc}, max eig err {err:.2e}", flush=True)
